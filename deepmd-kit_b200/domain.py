"""Spatial domain decomposition for multi-GPU force evaluation: one rank per GPU, each owning a brick
of the periodic box, ghost-atom coordinate halo forward and ghost-force halo back every step.

This is the role LAMMPS' Comm plays around PairDeepMD::compute in the reference
(source/lmp/pair_deepmd.cpp:217-229 swap plan hand-off, :482-488 reverse_comm of ghost forces;
source/lib/include/neighbor_list.h:29-57 carries the plan in InputNlist), re-done for NVSwitch: all
peers are one hop away at full bandwidth, so the 6 staged swaps collapse to ONE grouped NCCL
send/recv over the 26 neighbour directions (<= 7 distinct peers on 2x2x2, self-images copied
locally), bracketed by the dpb200 pack / unpack-accumulate kernels.  Energy and virial leave as
one 10-scalar all-reduce.

Atom migration (`migrate_atoms`, `DomainDeepPot.exchange_atoms`): when the list is rebuilt, atoms whose
wrapped coordinate has left the brick are handed to the rank that owns their new position (LAMMPS
`Comm::exchange`, which runs before `Comm::borders` on every reneighbouring step; the reference sees its
result as a changed nlocal / ilist, source/lmp/pair_deepmd.cpp:217-229).  Between rebuilds an atom may sit
outside its brick by up to skin / 2: the ghost shell is selected with rcut + skin.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import ops
from .model import DeepPotB200, NeighborState, SeAModel, type_partition

DIRS: List[Tuple[int, int, int]] = [(a, b, c) for a in (-1, 0, 1) for b in (-1, 0, 1) for c in (-1, 0, 1)
                                    if (a, b, c) != (0, 0, 0)]


def proc_grid(world: int) -> Tuple[int, int, int]:
    """Near-cubic factorisation, x fastest-growing: 1->(1,1,1) 2->(2,1,1) 4->(2,2,1) 8->(2,2,2)."""
    g = [1, 1, 1]
    n = world
    d = 0
    f = 2
    while n > 1:
        while n % f:
            f += 1
        g[d % 3] *= f
        n //= f
        d += 1
    g.sort(reverse=True)
    return tuple(g)


def rank_to_coords(rank: int, grid: Sequence[int]) -> Tuple[int, int, int]:
    gx, gy, gz = grid
    return (rank // (gy * gz), (rank // gz) % gy, rank % gz)


def coords_to_rank(c: Sequence[int], grid: Sequence[int]) -> int:
    gx, gy, gz = grid
    return ((c[0] % gx) * gy + (c[1] % gy)) * gz + (c[2] % gz)


def migrate_atoms(coord: torch.Tensor, atype: torch.Tensor, box, grid, rank: int, group=None,
                  payload: Sequence[torch.Tensor] = ()):
    """Hand every atom to the rank whose brick contains its periodically wrapped position.

    coord [n, 3], atype [n], payload = per-atom tensors ([n] or [n, k]; ids, velocities, ...) that travel with the
    atoms.  Returns (coord, atype, payload) of this rank afterwards: the atoms that stayed (original order, wrapped
    into the box) followed by the arrivals in direction order.  One grouped exchange over the 26 neighbour
    directions, like the halo; an atom that moved further than the neighbouring brick is an error (the caller
    rebuilt too rarely).  Device-agnostic torch code: the gloo tests run it on the CPU."""
    grid = tuple(int(x) for x in grid)
    world = grid[0] * grid[1] * grid[2]
    dev, dt = coord.device, coord.dtype
    c = coord.reshape(-1, 3)
    n = c.shape[0]
    b = torch.as_tensor(np.asarray(box, np.float64).reshape(3, 3), device=dev)
    rec = torch.linalg.inv(b)
    s = c.to(torch.float64) @ rec
    wrap = torch.floor(s)
    s = s - wrap
    c = (c.to(torch.float64) - wrap @ b).to(dt)
    me = rank_to_coords(rank, grid)
    gt = torch.as_tensor(grid, dtype=torch.float64, device=dev)
    bi = torch.clamp(torch.floor(s * gt).to(torch.int64), min=torch.zeros(3, dtype=torch.int64, device=dev),
                     max=torch.as_tensor(grid, dtype=torch.int64, device=dev) - 1)
    delta = bi - torch.as_tensor(me, dtype=torch.int64, device=dev)
    for d in range(3):  # minimum image on the process grid: -1, 0, +1 (+1 when both are the same peer)
        gd = grid[d]
        dd = torch.remainder(delta[:, d], gd)
        dd = torch.where(dd > gd // 2, dd - gd, dd)
        if gd == 2:
            dd = torch.abs(dd)
        delta[:, d] = dd
    if n and bool((delta.abs() > 1).any()):
        raise ValueError(f"rank {rank}: an atom moved past the neighbouring brick between two list rebuilds")
    cols = [c.to(torch.float64), atype.reshape(-1, 1).to(torch.float64)]
    shapes = []
    for q in payload:
        q2 = q.reshape(n, -1)
        shapes.append((q.dtype, tuple(q.shape[1:])))
        cols.append(q2.to(torch.float64))
    rows = torch.cat(cols, 1).contiguous()
    width = rows.shape[1]
    code = (delta[:, 0] + 1) * 9 + (delta[:, 1] + 1) * 3 + (delta[:, 2] + 1)  # 13 = stays
    send, dests, srcs = [], [], []
    for (dx, dy, dz) in DIRS:
        k = (dx + 1) * 9 + (dy + 1) * 3 + (dz + 1)
        send.append(rows[code == k])
        dests.append(coords_to_rank((me[0] + dx, me[1] + dy, me[2] + dz), grid))
        srcs.append(coords_to_rank((me[0] - dx, me[1] - dy, me[2] - dz), grid))
    keep = rows[code == 13]
    cnt = torch.tensor([int(x.shape[0]) for x in send], dtype=torch.int64, device=dev)
    if world > 1:
        allc = torch.empty(world * len(DIRS), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(allc, cnt, group=group)
        allc = allc.reshape(world, len(DIRS)).cpu()
    else:
        allc = cnt.reshape(1, -1).cpu()
    parts = [keep]
    p2p = []
    for k in range(len(DIRS)):
        nrecv = int(allc[srcs[k], k])
        if dests[k] == rank and srcs[k] == rank:  # a periodic self-image: the wrap above already placed the atom
            parts.append(send[k])
            continue
        if send[k].shape[0]:
            p2p.append(dist.P2POp(dist.isend, send[k].contiguous(), dests[k], group=group))
        if nrecv:
            buf = torch.empty((nrecv, width), dtype=torch.float64, device=dev)
            p2p.append(dist.P2POp(dist.irecv, buf, srcs[k], group=group))
            parts.append(buf)
    if p2p:
        for r in dist.batch_isend_irecv(p2p):
            r.wait()
    rows = torch.cat(parts, 0)
    new_c = rows[:, :3].to(dt).contiguous()
    new_t = rows[:, 3].round().to(atype.dtype).contiguous()
    out, col = [], 4
    for (qdt, tail) in shapes:
        w = int(np.prod(tail)) if tail else 1
        q = rows[:, col:col + w]
        q = q.round().to(qdt) if not qdt.is_floating_point else q.to(qdt)
        out.append(q.reshape((rows.shape[0],) + tail).contiguous())
        col += w
    return new_c, new_t, out


class HaloPlan:
    """Send lists per direction, receive segments, peers.  Device-agnostic (torch ops only) so the
    N>1 host logic is testable with the gloo backend on CPU."""

    def __init__(self, coord: torch.Tensor, box, grid, rank, rc: float, group=None, slack: float = 0.0):
        """`slack` (a length): how far outside its brick an owned atom may sit (half the skin)."""
        self.grid, self.rank, self.group = tuple(grid), rank, group
        self.world = grid[0] * grid[1] * grid[2]
        dev, dt = coord.device, coord.dtype
        b = np.asarray(box, np.float64).reshape(3, 3)
        vol = abs(np.linalg.det(b))
        rec = np.linalg.inv(b)  # fractional s = r @ rec
        face = np.array([vol / np.linalg.norm(np.cross(b[(d + 1) % 3], b[(d + 2) % 3])) for d in range(3)])
        me = rank_to_coords(rank, grid)
        lo = np.array([me[d] / grid[d] for d in range(3)])
        hi = np.array([(me[d] + 1) / grid[d] for d in range(3)])
        w = rc / face
        for d in range(3):
            if w[d] > 1.0 / grid[d] + 1e-12:
                raise ValueError(f"halo width {rc} exceeds the brick width along axis {d}: use fewer ranks on that axis")
        c = coord.reshape(-1, 3)
        s = c.to(torch.float64) @ torch.as_tensor(rec, device=dev)
        if c.shape[0]:
            tol = torch.as_tensor(slack / face + 1e-9, device=dev)
            out = ((s < torch.as_tensor(lo, device=dev) - tol) | (s >= torch.as_tensor(hi, device=dev) + tol)).any()
            if bool(out):
                raise ValueError(f"rank {rank}: some atoms lie outside this rank's brick by more than {slack} "
                                 "(hand them to their owners first: DomainDeepPot.exchange_atoms)")
        lists, shifts, dests, srcs = [], [], [], []
        for (dx, dy, dz) in DIRS:
            m = torch.ones(c.shape[0], dtype=torch.bool, device=dev)
            wrap = [0, 0, 0]
            for d, dd in enumerate((dx, dy, dz)):
                if dd == 1:
                    m &= s[:, d] >= hi[d] - w[d]
                    wrap[d] = 1 if me[d] + 1 >= grid[d] else 0
                elif dd == -1:
                    m &= s[:, d] < lo[d] + w[d]
                    wrap[d] = -1 if me[d] - 1 < 0 else 0
            idx = torch.nonzero(m).reshape(-1).to(torch.int32)
            sh = -(wrap[0] * b[0] + wrap[1] * b[1] + wrap[2] * b[2])
            lists.append(idx)
            shifts.append(torch.as_tensor(sh, dtype=dt, device=dev).expand(idx.numel(), 3))
            dests.append(coords_to_rank((me[0] + dx, me[1] + dy, me[2] + dz), grid))
            srcs.append(coords_to_rank((me[0] - dx, me[1] - dy, me[2] - dz), grid))
        self.dests, self.srcs = dests, srcs
        self.send_counts = [int(l.numel()) for l in lists]
        self.sendlist = torch.cat(lists) if lists else torch.zeros(0, dtype=torch.int32, device=dev)
        self.shift = torch.cat(shifts).contiguous()
        cnt = torch.tensor(self.send_counts, dtype=torch.int64, device=dev)
        if self.world > 1:
            allc = torch.empty(self.world * len(DIRS), dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(allc, cnt, group=group)
            allc = allc.reshape(self.world, len(DIRS)).cpu()
        else:
            allc = cnt.reshape(1, -1).cpu()
        # direction k arrives from srcs[k], which sent it as ITS direction k
        self.recv_counts = [int(allc[srcs[k], k]) for k in range(len(DIRS))]
        self.nsend = int(sum(self.send_counts))
        self.nghost = int(sum(self.recv_counts))
        self.send_off = np.concatenate([[0], np.cumsum(self.send_counts)]).astype(int)
        self.recv_off = np.concatenate([[0], np.cumsum(self.recv_counts)]).astype(int)

    # one grouped exchange: sendbuf segments -> dests, recvbuf segments <- srcs (reverse=True mirrors it)
    def exchange(self, sendbuf: torch.Tensor, recvbuf: torch.Tensor, reverse: bool = False):
        p2p = []
        s_off, r_off = (self.recv_off, self.send_off) if reverse else (self.send_off, self.recv_off)
        to, frm = (self.srcs, self.dests) if reverse else (self.dests, self.srcs)
        for k in range(len(DIRS)):
            a, bnd = s_off[k], s_off[k + 1]
            ra, rb = r_off[k], r_off[k + 1]
            if to[k] == self.rank and frm[k] == self.rank:
                if rb > ra:
                    recvbuf[ra:rb].copy_(sendbuf[a:bnd])
                continue
            if bnd > a:
                p2p.append(dist.P2POp(dist.isend, sendbuf[a:bnd], to[k], group=self.group))
            if rb > ra:
                p2p.append(dist.P2POp(dist.irecv, recvbuf[ra:rb], frm[k], group=self.group))
        if p2p:
            for r in dist.batch_isend_irecv(p2p):
                r.wait()


class DomainDeepPot(DeepPotB200):
    """DeepPotB200 over a brick decomposition: `eval_device(local coords)` returns the global energy
    and virial (all-reduced) and the forces on this rank's own atoms."""

    def __init__(self, model: SeAModel, grid, skin: float = 2.0, nlist_every: int = 10, group=None):
        super().__init__(model, skin, nlist_every, use_graph=False)  # NCCL exchanges stay eager
        self.grid = tuple(grid)
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.plan = None

    def make_local_water(self, water_box_fn, ncopy: int, jitter: float):
        """This rank's brick of the weak-scaling benchmark box: ncopy^3 replicas of the 192-atom frame per
        rank, global box = grid * brick."""
        c, t, b = water_box_fn(ncopy, jitter, seed=20260101 + self.rank)
        me = rank_to_coords(self.rank, self.grid)
        L = np.diag(b)
        c = c + np.array(me) * L
        return c, t, b * np.array(self.grid)[:, None]

    def make_local_copper(self, ncell: int, jitter: float = 0.05, a0: float = 3.615, seed: int = 20260102):
        """This rank's brick of ONE global FCC box of ncell^3 conventional cells (BASELINE config 3, strong scaling:
        the global box is fixed, every rank generates the cells of its own brick only).  Cells are split as evenly
        as the grid allows; the brick faces sit on cell boundaries."""
        me = rank_to_coords(self.rank, self.grid)
        lo = [ncell * me[d] // self.grid[d] for d in range(3)]
        hi = [ncell * (me[d] + 1) // self.grid[d] for d in range(3)]
        base = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]]) * a0
        cells = np.stack(np.meshgrid(*[np.arange(lo[d], hi[d]) for d in range(3)], indexing="ij"), -1).reshape(-1, 3) * a0
        coord = (cells[:, None, :] + base[None]).reshape(-1, 3)
        if jitter > 0:
            coord = coord + np.random.default_rng(seed + 7919 * self.rank).normal(scale=jitter, size=coord.shape)
        L = ncell * a0
        # (atoms pushed across a face by the jitter stay with this rank: the halo plan allows skin / 2 of slack;
        #  the brick bounds of the plan are fractions me/grid of the box, so uneven cell splits must not occur)
        for d in range(3):
            if ncell % self.grid[d]:
                raise ValueError(f"ncell = {ncell} is not divisible by the process grid {self.grid}")
        return coord, np.zeros(len(coord), np.int32), np.eye(3) * L

    def exchange_atoms(self, coord, atype, box, payload: Sequence[torch.Tensor] = ()):
        """migrate_atoms for this rank's brick; the next eval_device rebuilds the halo plan and the list (the local
        atom set changed).  An MD driver calls this whenever `needs_rebuild(...)` says so, before eval_device."""
        out = migrate_atoms(coord.reshape(-1, 3), atype, box, self.grid, self.rank, self.group, payload)
        self.state = None
        return out

    def needs_rebuild(self, coord, atype, box) -> bool:
        """True on every rank when any rank's list is stale (collective)."""
        return self._any_rank_stale(coord, atype, box)

    # -- host-side pieces (overridable for CPU tests) -----------------------------------------
    def _pack(self, coord, plan):
        return ops.halo_pack(coord, plan.sendlist, plan.shift)

    def _unpack_add(self, force, buf, plan):
        return ops.halo_unpack_add(force, buf, plan.sendlist)

    def halo_forward(self, coord: torch.Tensor) -> torch.Tensor:
        plan = self.plan
        nloc = coord.shape[0]
        ext = torch.empty((nloc + plan.nghost, 3), dtype=coord.dtype, device=coord.device)
        ext[:nloc].copy_(coord)
        plan.exchange(self._pack(coord, plan), ext[nloc:])
        return ext

    def halo_reverse(self, force_ext: torch.Tensor, nloc: int) -> torch.Tensor:
        plan = self.plan
        buf = torch.empty((plan.nsend, 3), dtype=force_ext.dtype, device=force_ext.device)
        plan.exchange(force_ext[nloc:].contiguous(), buf, reverse=True)
        f = force_ext[:nloc].contiguous()
        return self._unpack_add(f, buf, plan)

    def build_neighbors(self, coord, atype, box):
        m = self.model
        rc = m.cfg.rcut + self.skin
        c = coord.reshape(-1, 3)
        nloc = atype.numel()
        self.plan = HaloPlan(c, box, self.grid, self.rank, rc, self.group, slack=0.5 * self.skin)
        ext_c = self.halo_forward(c)
        gt = torch.empty(self.plan.nghost, dtype=torch.int32, device=c.device)
        self.plan.exchange(atype.to(torch.int32).index_select(0, self.plan.sendlist.long()), gt)
        ext_t = torch.cat([atype.to(torch.int32), gt]).contiguous()
        self.state = None
        numneigh, rows = ops.build_nlist(ext_c, nloc, rc, ext_t, cache=self._cache)
        perm, ranges, inv = self._type_partition(atype)
        ref = ops._buf(self._cache, "ref_coord", (nloc, 3), c.dtype, c.device)
        ref.copy_(c)
        self.state = NeighborState(nloc, ext_t, None, None, numneigh, rows, perm, ranges, type_inv=inv,
                                   chunks=self._chunks(atype),
                                   box=np.array(box, dtype=np.float64).reshape(9).copy(), ref_coord=ref)
        return self.state

    def _any_rank_stale(self, coord, atype, box) -> bool:
        """The halo plan and the raw list are rebuilt by ALL ranks together: the local staleness test of
        DeepPotB200 (cell changed, an atom moved more than skin / 2, reuse cap reached) is OR-reduced."""
        stale = self._list_is_stale(coord, atype, box)
        if dist.is_initialized() and dist.get_world_size(self.group) > 1:
            flag = torch.tensor([1 if stale else 0], dtype=torch.int32, device=coord.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=self.group)
            stale = bool(flag.item())
        return stale

    def eval_device(self, coord, atype, box, atom_virial=False, fused=True):
        st = self.state
        if self._any_rank_stale(coord, atype, box):
            st = self.build_neighbors(coord, atype, box)
        c = coord.reshape(-1, 3)
        ext_c = self.halo_forward(c)
        self._last_ext_coord = ext_c
        st.ago += 1
        if st.chunks is not None:  # per-atom intermediates exceed the device memory: slabs of centre atoms
            e, f_ext, virial, ex = self.model.evaluate_chunked(ext_c, st.ext_type, st.numneigh, st.rows, None, st.nloc,
                                                               st.chunks, atom_virial=atom_virial)
        else:
            e, f_ext, virial, ex = self.model.evaluate(ext_c, st.ext_type, st.numneigh, st.rows, None, st.nloc,
                                                       st.type_perm, st.type_ranges, atom_virial=atom_virial,
                                                       fused=fused, type_inv=st.type_inv)
        force = self.halo_reverse(f_ext, st.nloc)
        red = torch.cat([e.reshape(1), virial.reshape(9)])
        if dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(red, group=self.group)
        if atom_virial and ex.get("atom_virial") is not None:
            av_ext = ex["atom_virial"].reshape(-1, 9)
            av = av_ext[:st.nloc].contiguous()
            for q in range(3):  # 9 components as three 3-vectors through the same reverse halo
                part = self.halo_reverse(av_ext[:, 3 * q:3 * q + 3].contiguous(), st.nloc)
                av[:, 3 * q:3 * q + 3] = part
            ex["atom_virial"] = av.reshape(-1)
        return red[0], force, red[1:], ex
