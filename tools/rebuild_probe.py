"""Where does a neighbour-list rebuild spend its time? (host wall clock with syncs, per phase)"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as g
g.load_package()
from deepmd_kit_b200 import ops
from deepmd_kit_b200.model import DeepPotB200, SeAConfig, SeAModel

ncopy = int(sys.argv[1]) if len(sys.argv) > 1 else 20
dev = torch.device("cuda:0")
model = SeAModel(SeAConfig(), torch.float64, dev)
coord, atype, box = g.water_box(ncopy, 0.01)
c = torch.as_tensor(coord).to(dev); t = torch.as_tensor(atype).to(dev)
dp = DeepPotB200(model)
names = ["normalize_coord", "copy_coord", "build_nlist"]
orig = {n: getattr(ops, n) for n in names}
log = []
def wrap(n):
    def f(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = orig[n](*a, **k)
        torch.cuda.synchronize(); log.append((n, (time.perf_counter() - t0) * 1e3))
        return r
    return f
for n in names: setattr(ops, n, wrap(n))
ob = dp.build_neighbors
def bn(*a, **k):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = ob(*a, **k)
    torch.cuda.synchronize(); log.append(("build_neighbors_total", (time.perf_counter() - t0) * 1e3))
    return r
dp.build_neighbors = bn
for s in range(45):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    dp.eval_device(c, t, box)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) * 1e3
    if log:
        print(f"step {s}: {dt:.1f} ms  " + "  ".join(f"{n}={v:.1f}" for n, v in log) +
              f"  mem_alloc={torch.cuda.memory_allocated()/2**30:.1f}G reserved={torch.cuda.memory_reserved()/2**30:.1f}G", flush=True)
        log.clear()
    elif s % 10 == 5:
        print(f"step {s}: {dt:.1f} ms", flush=True)
