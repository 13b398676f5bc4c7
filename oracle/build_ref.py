"""TEST INFRASTRUCTURE ONLY - byte-compile the reference translator modules into oracle/_ref/.

    python -m oracle.build_ref          (build container: needs /root/reference; called by __graft_entry__.build())

The reference is 100 % Python: its "compiled" form is CPython bytecode.  This recipe imports the reference classes through
`oracle/ref_shims.py` (which records every file of the reference tree that gets imported), and writes one sourceless bytecode file
per such file to `oracle/_ref/<same relative path>.bc` (pyc format; the gpurun snapshot drops `*.pyc`).  `oracle/_ref/` is git-ignored (no reference source or derived file
enters history) but NOT gpurun-ignored, so on the GPU box - where /root/reference does not exist - `ref_shims` imports the very
same classes from these files: `bench.py --impl reference`, `bench.py`'s `gpu_baseline` leg and the `requires_reference`
tests then execute the unmodified reference code, not a restatement.
"""
from __future__ import annotations

import os
import py_compile
import shutil
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import ref_shims as rs  # noqa: E402

EXTRA = ["HOI/evaluation/pnr/metrics.py", "HHI/utils/ttm/utils.py", "HOI/evaluation/lta/lta_metrics.py"]


def main() -> int:
    if rs.REFERENCE_ROOT == rs.COMPILED_ROOT:
        print("build_ref: no reference source tree here; keeping", rs.COMPILED_ROOT)
        return 0
    rs.LOADED_FILES.clear()
    rs.load_hhi()
    rs.load_hoi()
    files = set(rs.LOADED_FILES)
    for rel in EXTRA:        # metric post-processing functions the (f)-3 oracles are pinned against (imported by path)
        f = os.path.join(rs.REFERENCE_ROOT, rel)
        if os.path.exists(f):
            files.add(f)
    if os.path.isdir(rs.COMPILED_ROOT):
        shutil.rmtree(rs.COMPILED_ROOT)
    n = 0
    for f in sorted(files):
        rel = os.path.relpath(f, rs.REFERENCE_ROOT)
        dst = os.path.join(rs.COMPILED_ROOT, rel[:-3] + rs.COMPILED_SUFFIX)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        py_compile.compile(f, cfile=dst, dfile=rel, doraise=True)
        n += 1
    # the loaders test for these directories
    for d in ("HHI/models", "HOI/models"):
        os.makedirs(os.path.join(rs.COMPILED_ROOT, d), exist_ok=True)
    print(f"build_ref: {n} reference modules byte-compiled into {rs.COMPILED_ROOT}")
    return n


if __name__ == "__main__":
    main()
