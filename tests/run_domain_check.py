"""torchrun worker: N-rank spatial decomposition vs. the single-GPU evaluation of the same global box.
Launched by tests/test_gpu_multi.py (and usable by hand:
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/run_domain_check.py)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    g.load_package()
    from deepmd_kit_b200.domain import DomainDeepPot, proc_grid
    from deepmd_kit_b200.model import DeepPotB200, SeAConfig, SeAModel

    ncopy = int(os.environ.get("DPB_NCOPY", "2"))
    dtype = torch.float64
    model = SeAModel(SeAConfig(), dtype, dev)
    grid = proc_grid(world)
    dp = DomainDeepPot(model, grid, skin=2.0, nlist_every=10)
    coord, atype, box = dp.make_local_water(g.water_box, ncopy, 0.01)
    c_d = torch.as_tensor(coord).to(dev)
    t_d = torch.as_tensor(atype).to(dev)
    e, f, v, _ = dp.eval_device(c_d, t_d, box)
    e2, f2, v2, _ = dp.eval_device(c_d, t_d, box)  # list reuse (ago > 0)
    assert torch.allclose(f, f2, rtol=0, atol=1e-12)
    # gather the global system on every rank and evaluate it on one GPU
    cs = [None] * world
    ts = [None] * world
    dist.all_gather_object(cs, coord)
    dist.all_gather_object(ts, atype)
    gc, gt = np.concatenate(cs), np.concatenate(ts)
    off = np.cumsum([0] + [len(x) for x in ts])
    ref = DeepPotB200(model, skin=2.0)
    er, fr, vr, _ = ref.eval_device(torch.as_tensor(gc).to(dev), torch.as_tensor(gt).to(dev), box)
    fr_mine = fr[off[rank]:off[rank + 1]]
    ferr = float((f - fr_mine).abs().max() / fr.abs().max())
    eerr = float(abs(e - er) / abs(er))
    verr = float((v - vr).abs().max() / vr.abs().max())
    print(f"[rank {rank}/{world}] grid {grid} nloc {len(atype)} nghost {dp.plan.nghost} "
          f"rel.err energy {eerr:.2e} force {ferr:.2e} virial {verr:.2e}", flush=True)
    assert eerr < 1e-10 and ferr < 1e-10 and verr < 1e-10
    # public host API on every rank
    eh, fh, vh = dp.eval(coord.reshape(1, -1), box.reshape(1, 9), atype)
    assert abs(eh[0, 0] - float(er)) < 1e-10 * abs(float(er))
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("DOMAIN_CHECK_OK", flush=True)


if __name__ == "__main__":
    main()
