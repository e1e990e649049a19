"""One launch of each tcgen05 fitting GEMM shape at chunk size (for ncu): L0 fwd, hidden fwd, hidden bwd, L0 bwd."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

ge.load_package()
from deepmd_kit_b200 import ops  # noqa: E402
from deepmd_kit_b200.model import FittingNet  # noqa: E402

dev = torch.device("cuda:0")
n = int(os.environ.get("N", 131072))
net = FittingNet(1600, (240, 240, 240), True, 3, torch.float64, dev)
assert net.prepare_tc(6)
g = torch.Generator().manual_seed(5)
d = (torch.randn(n, 1600, generator=g, dtype=torch.float64) * 0.05).to(dev)
xs, ex = ops.split_i8_rows(d, 6)
for _ in range(int(os.environ.get("REPS", 2))):
    e, gd = net.forward_backward_tc(xs, ex, n)
torch.cuda.synchronize()
print("ok", float(e.sum()))
