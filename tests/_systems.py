"""Shared synthetic systems for the parity tests (no reference files read at run time)."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def six_atom_system(cpu, rc=6.0, dtype=np.float64):
    """The 6-atom fixture of the reference lib tests (test_env_mat_a.cc:26-36,176-194):
    box 13 A, periodic images copied with `rc`, raw list built with `rc`."""
    g = golden("env_mat_a.json")["TestEnvMatA"]
    posi = np.array(g["posi"], dtype=np.float64).reshape(-1, 3)
    atype = np.array(g["atype"], dtype=np.int32)
    box = np.diag([13.0, 13.0, 13.0])
    coord, at, mapping = cpu.copy_coord(posi, atype, box, rc)
    nloc = len(atype)
    numneigh, rows = cpu.build_nlist(coord, nloc, rc)
    return dict(coord=coord.astype(dtype), atype=at, mapping=mapping, nloc=nloc, numneigh=numneigh, rows=rows,
                box=box, posi=posi)


def water_like_box(ncopy=2, seed=0, jitter=0.05, dtype=np.float64):
    """A small periodic O/H box: 4x4x4 simple-cubic sites per unit (spacing 3.1 A) with one O and
    two H per site, replicated `ncopy` times per axis (exact replication => distance ties), then
    Gaussian jitter (0 keeps the ties)."""
    rng = np.random.default_rng(seed)
    a = 3.1
    nsite = 3
    base_o = np.stack(np.meshgrid(*[np.arange(nsite)] * 3, indexing="ij"), -1).reshape(-1, 3) * a + 0.4
    dirs = rng.normal(size=(len(base_o), 2, 3))
    dirs /= np.linalg.norm(dirs, axis=-1, keepdims=True)
    base_h = base_o[:, None, :] + 0.97 * dirs
    unit = np.concatenate([base_o, base_h.reshape(-1, 3)])
    utype = np.concatenate([np.zeros(len(base_o), np.int32), np.ones(2 * len(base_o), np.int32)])
    L = nsite * a
    shifts = np.stack(np.meshgrid(*[np.arange(ncopy)] * 3, indexing="ij"), -1).reshape(-1, 3) * L
    coord = (unit[None] + shifts[:, None]).reshape(-1, 3)
    atype = np.tile(utype, len(shifts))
    if jitter > 0:
        coord = coord + rng.normal(scale=jitter, size=coord.shape)
    box = np.diag([L * ncopy] * 3)
    return coord.astype(dtype), atype, box.astype(dtype)


def extended_system(cpu, coord, atype, box, rc_list, dtype=np.float64):
    """normalize -> periodic ghost copy -> raw list (strict cutoff rc_list)."""
    c = cpu.normalize_coord(np.asarray(coord, dtype=dtype), box)
    ext_c, ext_t, mapping = cpu.copy_coord(c, atype, box, rc_list)
    nloc = len(atype)
    numneigh, rows = cpu.build_nlist(ext_c, nloc, rc_list)
    return dict(coord=ext_c, atype=ext_t, mapping=mapping, nloc=nloc, numneigh=numneigh, rows=rows, box=box)


def random_table(nspline, M, rng, dtype=np.float64):
    return (rng.normal(size=(nspline, M * 6)) * np.array([1, 1, 0.5, 0.2, 0.1, 0.05] * M)).astype(dtype)
