"""N>1 host logic of the spatial decomposition on CPU: world_size 2 (and 4), gloo backend.
Checks the halo plan against a brute-force enumeration of periodic images and the reverse
(ghost force) exchange against a direct fold.  The CUDA pack/unpack kernels are replaced by
their torch one-liners here (tests only); the plan, the counts and the message order are the
product code."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import __graft_entry__ as g


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, grid, rc, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g.load_package()
        from deepmd_kit_b200.domain import DomainDeepPot, rank_to_coords

        class CpuDomain(DomainDeepPot):
            def _pack(self, coord, plan):
                return coord.index_select(0, plan.sendlist.long()) + plan.shift

            def _unpack_add(self, force, buf, plan):
                return force.index_add_(0, plan.sendlist.long(), buf)

        dp = CpuDomain(None, grid, skin=0.0)
        rng = np.random.default_rng(100 + rank)
        Lb = np.array([9.0, 8.0, 7.5])  # brick edge
        me = np.array(rank_to_coords(rank, grid))
        nloc = 150 + 7 * rank
        coord = rng.uniform(0, 1, size=(nloc, 3)) * Lb + me * Lb
        box = np.diag(Lb * np.array(grid))
        from deepmd_kit_b200.domain import HaloPlan

        c = torch.as_tensor(coord)
        dp.plan = HaloPlan(c, box, grid, rank, rc)
        ext = dp.halo_forward(c).numpy()
        # brute force: every periodic image of every atom of every rank inside my brick + shell
        sizes = [None] * world
        dist.all_gather_object(sizes, coord)
        allc = np.concatenate(sizes)
        owner_off = np.cumsum([0] + [len(s) for s in sizes])
        lo, hi = me * Lb, (me + 1) * Lb
        want = []
        for sx in (-1, 0, 1):
            for sy in (-1, 0, 1):
                for sz in (-1, 0, 1):
                    img = allc + np.array([sx, sy, sz]) @ box
                    inside = np.all((img >= lo - rc) & (img < hi + rc), axis=1)
                    core = np.all((img >= lo) & (img < hi), axis=1)
                    want.append(img[inside & ~core])
        want = np.concatenate(want)
        got = ext[nloc:]
        key = lambda a: sorted(map(tuple, np.round(a, 9)))
        assert len(got) == len(want), (len(got), len(want))
        assert key(got) == key(want)
        assert np.array_equal(ext[:nloc], coord)
        # reverse halo: every ghost carries weight 1 -> each owner ends with 1 + (number of its images in use)
        f_ext = torch.ones(len(ext), 3, dtype=torch.float64)
        f = dp.halo_reverse(f_ext, nloc).numpy()
        tot = torch.tensor([f.sum()])
        dist.all_reduce(tot)
        n_ext = torch.tensor([float(len(ext))])
        dist.all_reduce(n_ext)
        assert abs(tot.item() - 3 * n_ext.item()) < 1e-9  # nothing lost, nothing duplicated
        # per-atom: count of my images selected anywhere == f - 1
        cnt = np.zeros(nloc)
        for k, off in enumerate(dp.plan.send_off[:-1]):
            idx = dp.plan.sendlist[off:dp.plan.send_off[k + 1]].numpy()
            np.add.at(cnt, idx, 1)
        assert np.allclose(f[:, 0], 1 + cnt)
        # ghost types travel with the same plan
        t = torch.full((nloc,), rank, dtype=torch.int32)
        gt = torch.empty(dp.plan.nghost, dtype=torch.int32)
        dp.plan.exchange(t.index_select(0, dp.plan.sendlist.long()), gt)
        assert set(gt.tolist()) <= set(range(world))
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,grid", [(2, (2, 1, 1)), (4, (2, 2, 1))])
def test_halo_plan_gloo(tmp_path, world, grid):
    g.build()
    port = _free_port()
    mp.spawn(_worker, args=(world, port, grid, 2.5, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))


def test_proc_grid():
    g.load_package()
    from deepmd_kit_b200.domain import coords_to_rank, proc_grid, rank_to_coords

    assert proc_grid(1) == (1, 1, 1) and proc_grid(2) == (2, 1, 1) and proc_grid(4) == (2, 2, 1)
    assert proc_grid(8) == (2, 2, 2)
    for w in (2, 4, 8, 6):
        gr = proc_grid(w)
        assert gr[0] * gr[1] * gr[2] == w
        for r in range(w):
            assert coords_to_rank(rank_to_coords(r, gr), gr) == r
