"""TTM evaluation post-processing on the device (HHI/utils/ttm/utils.py:45-80, `PostProcessor`; used by
HHI/tasks/ttm/video_task.py:41-50).

The reference keeps every minibatch's logits, and whenever the segment id changes it concatenates them, takes
`softmax(mean(0))[1]` and reads the score and four scalars back with `.item()` - a device synchronisation per segment.  This
class has the same `update(outputs, targets)` / `save()` interface and writes the same two csv files, but only records the
minibatches (logits stay on the device) and scores ALL segments in one launch (`egot2_segment_softmax_mean`) followed by one
device -> host copy when `save()` (or `results()`) is called.
"""
from __future__ import annotations

import os
from typing import List, Optional

import torch

from . import _lib as L
from .engine import _stream


def segment_scores(logits: torch.Tensor, seg_offsets: torch.Tensor) -> torch.Tensor:
    """softmax(mean over rows [off[s], off[s+1]) of logits) for every segment: (n_seg, n_cls) fp32 on the device."""
    if not logits.is_cuda:
        raise L.Egot2Error("egot2_b200.ttm_postprocess runs on CUDA tensors only (no CPU fallback)")
    dev = logits.device
    lg = logits.to(torch.float32).contiguous()
    off = seg_offsets.to(device=dev, dtype=torch.int32).contiguous()
    n_seg = off.numel() - 1
    out = torch.empty((n_seg, lg.shape[1]), device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        L.call("egot2_segment_softmax_mean", n_seg, lg.shape[1], lg.data_ptr(), off.data_ptr(), out.data_ptr(), _stream())
    return out


class PostProcessor:
    def __init__(self, args, run_evaluation=None):
        self.exp_path = args.exp_path
        self.save_path = f'{self.exp_path}/tmp'
        self.groundtruth: List[list] = []
        self.prediction: List[list] = []
        self.groundtruthfile = f'{self.save_path}/gt.csv.rank.{args.rank}'
        self.predctionfile = f'{self.save_path}/pred.csv.rank.{args.rank}'
        self._run_evaluation = run_evaluation          # the reference's utils.ttm.metrics.run_evaluation (not part of this path)
        self._batches: List[tuple] = []                # (segid, logits on the device, start, end, label) per minibatch
        self._merged = 0

    def update(self, outputs, targets):
        """utils.py:57-69: one minibatch; consecutive minibatches with the same `uid:index` form one segment."""
        segid = targets[0][0] + ':' + str(int(targets[-1]))
        self._batches.append((segid, outputs.detach(), int(targets[-3]), int(targets[-2]), int(targets[2])))

    def _merge_all(self):
        """Everything recorded since the last merge: runs of equal segment ids -> one (ground truth, prediction) row each
        (utils.py:71-80), one kernel launch and one read-back for all of them."""
        todo = self._batches[self._merged:]
        if not todo:
            return
        runs, offs, rows = [], [0], 0
        for segid, out, start, end, label in todo:
            if runs and runs[-1][0] == segid:
                runs[-1][1] = min(runs[-1][1], start)
                runs[-1][2] = max(runs[-1][2], end)
            else:
                if runs:
                    offs.append(rows)
                runs.append([segid, start, end, label])
            rows += out.shape[0]
        offs.append(rows)
        logits = torch.cat([b[1] for b in todo], dim=0)
        scores = segment_scores(logits, torch.tensor(offs, dtype=torch.int32))[:, 1].cpu().tolist()
        for (segid, start, end, label), score in zip(runs, scores):
            uid, idx = segid.split(':')
            self.groundtruth.append([uid, idx, start, end, label])
            self.prediction.append([uid, idx, start, end, 1, score])
        self._merged = len(self._batches)

    def results(self):
        self._merge_all()
        return self.groundtruth, self.prediction

    def save(self):
        """utils.py:82-94: the same two csv files (no header, no index)."""
        import pandas as pd
        os.makedirs(self.save_path, exist_ok=True)
        self._merge_all()
        for f in (self.groundtruthfile, self.predctionfile):
            if os.path.exists(f):
                os.remove(f)
        pd.DataFrame(self.groundtruth).to_csv(self.groundtruthfile, index=False, header=None)
        pd.DataFrame(self.prediction).to_csv(self.predctionfile, index=False, header=None)

    def get_mAP(self):
        """utils.py:96-116: concatenates the per-rank files and hands them to the reference's evaluation script, which is
        outside the translator path; pass it in as `run_evaluation`."""
        import glob
        import shutil
        if self._run_evaluation is None:
            raise L.Egot2Error("PostProcessor.get_mAP needs the reference's utils.ttm.metrics.run_evaluation (pass run_evaluation=...)")
        merge_path = f'{self.exp_path}/result'
        os.makedirs(merge_path, exist_ok=True)
        out = []
        for name in ("gt", "pred"):
            dst = f'{merge_path}/{name}.csv'
            with open(dst, "w") as fo:
                for part in sorted(glob.glob(f'{self.save_path}/{name}.csv.rank.*')):
                    fo.write(open(part).read())
            out.append(dst)
        shutil.rmtree(self.save_path)
        return self._run_evaluation(*out)
