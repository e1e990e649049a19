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


def _migrate_worker(rank, world, port, grid, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g.load_package()
        from deepmd_kit_b200.domain import HaloPlan, migrate_atoms, rank_to_coords

        rng = np.random.default_rng(7 + rank)
        Lb = np.array([9.0, 8.0, 7.5])
        me = np.array(rank_to_coords(rank, grid))
        box = np.diag(Lb * np.array(grid))
        nloc = 400 + 11 * rank
        coord = rng.uniform(0, 1, size=(nloc, 3)) * Lb + me * Lb
        ids = np.arange(nloc, dtype=np.int64) + 100000 * rank
        atype = (ids % 3).astype(np.int32)
        vel = np.stack([ids * 0.5, -ids * 0.25, ids * 1e-3], 1)
        moved = coord + rng.uniform(-2.5, 2.5, size=coord.shape)  # crosses faces, edges, corners and the periodic wrap
        c2, t2, (id2, v2) = migrate_atoms(torch.as_tensor(moved), torch.as_tensor(atype), box, grid, rank, None,
                                          payload=(torch.as_tensor(ids), torch.as_tensor(vel)))
        c2, t2, id2, v2 = c2.numpy(), t2.numpy(), id2.numpy(), v2.numpy()
        assert t2.dtype == np.int32 and id2.dtype == np.int64 and c2.shape == (len(id2), 3)
        lo, hi = me * Lb, (me + 1) * Lb
        assert np.all((c2 >= lo - 1e-9) & (c2 < hi + 1e-9)), "an atom is outside its new owner's brick"
        assert np.array_equal(t2, (id2 % 3).astype(np.int32))
        assert np.allclose(v2, np.stack([id2 * 0.5, -id2 * 0.25, id2 * 1e-3], 1))
        # nothing lost, nothing duplicated; every atom sits at its wrapped position
        got = [None] * world
        dist.all_gather_object(got, (id2, c2))
        sent = [None] * world
        dist.all_gather_object(sent, (ids, moved))
        all_ids = np.concatenate([x[0] for x in got])
        assert len(all_ids) == len(set(all_ids.tolist())) == sum(len(x[0]) for x in sent)
        want = {int(i): p for ii, pp in sent for i, p in zip(ii, pp)}
        Lg = np.diag(box)
        for i, p in zip(id2, c2):
            w = want[int(i)] - np.floor(want[int(i)] / Lg) * Lg
            assert np.allclose(p, w, atol=1e-9)
        # the halo plan accepts the migrated atoms (it refuses atoms outside the brick)
        HaloPlan(torch.as_tensor(c2), box, grid, rank, 2.0)
        # an atom that skipped a whole brick is an error, not a silent misplacement
        if grid[0] >= 3 and rank == 0:
            far = coord.copy()
            far[0, 0] += 2 * Lb[0]
            with pytest.raises(ValueError):
                migrate_atoms(torch.as_tensor(far), torch.as_tensor(atype), box, grid, rank, None)
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,grid", [(2, (2, 1, 1)), (4, (2, 2, 1))])
def test_atom_migration_gloo(tmp_path, world, grid):
    g.build()
    port = _free_port()
    mp.spawn(_migrate_worker, args=(world, port, grid, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))


def test_atom_migration_single_rank_wraps():
    g.load_package()
    from deepmd_kit_b200.domain import migrate_atoms

    box = np.diag([5.0, 6.0, 7.0])
    c = torch.tensor([[5.5, -0.5, 3.0], [1.0, 2.0, 14.5]], dtype=torch.float64)
    t = torch.tensor([1, 0], dtype=torch.int32)
    c2, t2, (q,) = migrate_atoms(c, t, box, (1, 1, 1), 0, None, payload=(torch.tensor([7, 8]),))
    assert torch.allclose(c2, torch.tensor([[0.5, 5.5, 3.0], [1.0, 2.0, 0.5]], dtype=torch.float64))
    assert t2.tolist() == [1, 0] and q.tolist() == [7, 8]


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
