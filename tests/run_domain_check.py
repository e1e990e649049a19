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


def diagnose(dp, coord, atype, box, rank, world, cs, ts, rex, ref, off, dev):
    """Where does the decomposed evaluation differ? ghost set vs brute force, per-atom energies."""
    from deepmd_kit_b200.domain import rank_to_coords

    nloc = len(atype)
    c_d = torch.as_tensor(coord).to(dev)
    ext = dp.halo_forward(c_d).cpu().numpy()
    ext_t = dp.state.ext_type.cpu().numpy()
    allc, allt = np.concatenate(cs), np.concatenate(ts)
    L = np.diag(box)
    me = np.array(rank_to_coords(rank, dp.grid))
    lo, hi = me * L / np.array(dp.grid), (me + 1) * L / np.array(dp.grid)
    rc = dp.model.cfg.rcut + dp.skin
    want = []
    for sx in (-1, 0, 1):
        for sy in (-1, 0, 1):
            for sz in (-1, 0, 1):
                img = allc + np.array([sx, sy, sz]) * L
                inside = np.all((img >= lo - rc) & (img < hi + rc), axis=1)
                core = np.all((img >= lo) & (img < hi), axis=1)
                sel = inside & ~core
                want.append(np.concatenate([img[sel], allt[sel, None].astype(np.float64)], axis=1))
    want = np.concatenate(want)
    got = np.concatenate([ext[nloc:], ext_t[nloc:, None].astype(np.float64)], axis=1)
    key = lambda a: sorted(map(tuple, np.round(a, 7)))
    kg, kw = key(got), key(want)
    print(f"[diag {rank}] ghosts got {len(kg)} want {len(kw)} equal {kg == kw}", flush=True)
    if kg != kw:
        sg, sw = set(kg), set(kw)
        print(f"[diag {rank}] missing {len(sw - sg)} extra {len(sg - sw)} e.g. {list(sw - sg)[:3]} / {list(sg - sw)[:3]}",
              flush=True)
    e, f, v, ex = dp.eval_device(c_d, torch.as_tensor(atype).to(dev), box)
    ea = ex["atom_energy"].cpu().numpy()
    ea_ref = rex["atom_energy"].cpu().numpy()[off[rank]:off[rank + 1]]
    bad = np.nonzero(np.abs(ea - ea_ref) > 1e-9)[0]
    print(f"[diag {rank}] atoms with different energy: {len(bad)} of {nloc}; max diff {np.abs(ea - ea_ref).max():.3e}",
          flush=True)
    if len(bad):
        print(f"[diag {rank}] positions (fraction of brick) {((coord[bad[:6]] - lo) / (hi - lo)).round(3).tolist()}", flush=True)
    # neighbour counts
    nl = ex["nlist"].cpu().numpy()
    nl_ref = rex["nlist"].cpu().numpy()[off[rank]:off[rank + 1]]
    cnt, cnt_ref = (nl >= 0).sum(1), (nl_ref >= 0).sum(1)
    print(f"[diag {rank}] formatted neighbour counts differ for {int((cnt != cnt_ref).sum())} atoms", flush=True)


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import datetime

    dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))
    g.load_package()
    from deepmd_kit_b200.domain import DomainDeepPot, proc_grid
    from deepmd_kit_b200.model import DeepPotB200, SeAConfig, SeAModel

    ncopy = int(os.environ.get("DPB_NCOPY", "2"))
    dtype = torch.float64
    if os.environ.get("DPB_MODEL", "se_a") == "se_atten":  # config 5: DPA-1 strip mode with the pair-indexed gate
        from deepmd_kit_b200.atten import SeAttenConfig, SeAttenModel

        model = SeAttenModel(SeAttenConfig(), dtype, dev)
    elif os.environ.get("DPB_MODEL") == "dpa1_attn":  # DPA-1 with two attention layers (slab-wise, no CUDA graph)
        from deepmd_kit_b200.atten import SeAttenConfig, SeAttenModel

        model = SeAttenModel(SeAttenConfig(attn_layer=2), dtype, dev)
    else:
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
    er, fr, vr, rex = ref.eval_device(torch.as_tensor(gc).to(dev), torch.as_tensor(gt).to(dev), box)
    _ = rex
    fr_mine = fr[off[rank]:off[rank + 1]]
    if os.environ.get("DPB_DIAG"):
        diagnose(dp, coord, atype, box, rank, world, cs, ts, _, ref, off, dev)
    ferr = float((f - fr_mine).abs().max() / fr.abs().max())
    eerr = float(abs(e - er) / abs(er))
    verr = float((v - vr).abs().max() / vr.abs().max())
    print(f"[rank {rank}/{world}] grid {grid} nloc {len(atype)} nghost {dp.plan.nghost} "
          f"rel.err energy {eerr:.2e} force {ferr:.2e} virial {verr:.2e}", flush=True)
    assert eerr < 1e-10 and ferr < 1e-10 and verr < 1e-10
    # public host API on every rank
    eh, fh, vh = dp.eval(coord.reshape(1, -1), box.reshape(1, 9), atype)
    assert abs(eh[0, 0] - float(er)) < 1e-10 * abs(float(er))
    # atom migration: the whole system drifts by more than skin / 2 (atoms cross brick faces and the periodic
    # boundary), every rank hands its leavers to their new owners, and the decomposed evaluation still equals the
    # single-GPU one of the same moved system, atom by atom (global ids travel as payload)
    drift = np.array([3.1, -2.7, 1.3])
    ids = torch.arange(off[rank], off[rank + 1], dtype=torch.int64, device=dev)
    moved = c_d + torch.as_tensor(drift, dtype=dtype, device=dev)
    assert dp.needs_rebuild(moved, t_d, box)
    c_m, t_m, (ids_m,) = dp.exchange_atoms(moved, t_d, box, payload=(ids,))
    n_m = torch.tensor([t_m.numel()], dtype=torch.int64, device=dev)
    dist.all_reduce(n_m)
    assert int(n_m.item()) == len(gt)
    em_, fm, vm, _ = dp.eval_device(c_m, t_m, box)
    er2, fr2, vr2, _ = ref.eval_device(torch.as_tensor(gc + drift).to(dev), torch.as_tensor(gt).to(dev), box)
    ferr2 = float((fm - fr2[ids_m]).abs().max() / fr2.abs().max())
    eerr2 = float(abs(em_ - er2) / abs(er2))
    verr2 = float((vm - vr2).abs().max() / vr2.abs().max())
    changed = int((ids_m.numel() != ids.numel()) or not torch.equal(ids_m, ids))
    print(f"[rank {rank}/{world}] after migration: nloc {t_m.numel()} (owner set changed: {changed}) "
          f"rel.err energy {eerr2:.2e} force {ferr2:.2e} virial {verr2:.2e}", flush=True)
    assert eerr2 < 1e-10 and ferr2 < 1e-10 and verr2 < 1e-10
    assert world == 1 or changed
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("DOMAIN_CHECK_OK", flush=True)


if __name__ == "__main__":
    main()
