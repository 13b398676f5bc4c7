"""GPU parity of cases that were added after the round's GPU budget was spent (oracle.cases.UNVALIDATED_ON_GPU).
Their CPU-side checks are green (state_dict keys, same-seed init, oracle pinned against the reference class, golden);
the hardware run is pending, so these are xfail(strict=False): a pass shows up as XPASS, a failure does not turn the suite
red.  Move a name out of UNVALIDATED_ON_GPU once it has passed on a B200 and it joins the regular parametrisations."""
import pytest

from oracle.cases import UNVALIDATED_ON_GPU

pytestmark = [pytest.mark.gpu, pytest.mark.xfail(strict=False, reason="not yet run on hardware (GPU budget of round 1 spent)")]


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("name", sorted(UNVALIDATED_ON_GPU))
def test_engine_parity_unvalidated(name, dtype):
    import test_gpu_parity as tg          # the tests directory is on sys.path (rootdir conftest / prepend import mode)
    tg.test_engine_matches_oracle_and_golden(name, dtype)


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("name", sorted(UNVALIDATED_ON_GPU))
def test_module_parity_unvalidated(name, dtype):
    import test_modules as tm
    tm.test_module_forward_backward_vs_oracle(name, dtype)
