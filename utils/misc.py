"""Host-side logging helpers with the names the reference's main_acgan.py imports (main_acgan.py:13).
Not on the hot path; written independently (dataclass-style meter, one-hot via F.one_hot)."""
from dataclasses import dataclass

import torch
import torch.nn.functional as F


@dataclass
class AverageMeter:
    """Weighted running mean. `val` is the last sample, `avg` the mean so far (API of reference utils/misc.py:3-18)."""
    val: float = 0.0
    sum: float = 0.0
    count: int = 0

    @property
    def avg(self):
        return self.sum / self.count if self.count else 0.0

    def reset(self):
        self.val, self.sum, self.count = 0.0, 0.0, 0

    def update(self, val, n=1):
        self.val = val
        self.sum = self.sum + val * n
        self.count = self.count + n


@torch.no_grad()
def accuracy(output, target, topk=(1,)):
    """precision@k in percent for each k (API of reference utils/misc.py:20-32)."""
    ranked = output.topk(max(topk), dim=1).indices              # (B, maxk)
    hits = ranked.eq(target.reshape(-1, 1))                     # (B, maxk) bool
    scale = 100.0 / target.shape[0]
    return [hits[:, :k].any(dim=1).float().sum().reshape(1) * scale for k in topk]


def to_one_hot(y, n_class):
    """int64 labels (B,) -> float32 one-hot (B, n_class) (API of reference utils/misc.py:34-36)."""
    return F.one_hot(y.reshape(-1), n_class).to(torch.float32)
