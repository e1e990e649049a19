"""`dp compress` table builder for se_e2_a (type_one_side) — the data producer of the hot path.

Restates deepmd/utils/tabulate.py:67-147 (grid and nspline), :275-371 (quintic Hermite
coefficients from value / first / second derivative at both ends of every interval), :468-486
(table range from davg/dstd and the closest pair distance) and the embedding-net evaluation of
deepmd/utils/tabulate_math.py:353-470 (tanh MLP, resnet doubling when the width doubles, no
`idt`), with the info vector of deepmd/pt/model/descriptor/se_a.py:725-736.  Runs in float64 on
whatever device the weights live on; it is an offline step (once per model), not part of the
timed path.
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import numpy as np
import torch


class EmbeddingNet:
    """tanh MLP 1 -> neuron[0] -> ... with the reference's default init (mlp.py:149-163) and
    skip connections (same width: +x, doubled width: +concat(x, x))."""

    def __init__(self, neuron: Sequence[int], seed: int, resnet_dt: bool = False):
        if resnet_dt:
            raise NotImplementedError("model compression presumes resnet_dt = false for the embedding net")
        g = torch.Generator().manual_seed(seed)
        self.weights: List[torch.Tensor] = []
        self.biases: List[torch.Tensor] = []
        n_in = 1
        for n_out in neuron:
            w = torch.empty(n_in, n_out, dtype=torch.float64).normal_(0.0, 1.0 / math.sqrt(n_in + n_out), generator=g)
            b = torch.empty(n_out, dtype=torch.float64).normal_(0.0, 1.0, generator=g)
            self.weights.append(w)
            self.biases.append(b)
            n_in = n_out

    def to(self, device):
        self.weights = [w.to(device) for w in self.weights]
        self.biases = [b.to(device) for b in self.biases]
        return self

    def value_and_derivatives(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """x: [n] -> (v, dv/dx, d2v/dx2), each [n, neuron[-1]] (tabulate_math.py:353-470)."""
        h = x.reshape(-1, 1).to(torch.float64)
        d1 = torch.ones_like(h)
        d2 = torch.zeros_like(h)
        for w, b in zip(self.weights, self.biases):
            a = torch.tanh(h @ w + b)
            u = d1 @ w
            q = d2 @ w
            s = 1.0 - a * a
            a1 = s * u
            a2 = s * (q - 2.0 * a * u * u)
            n_in, n_out = w.shape
            if n_out == n_in:
                a, a1, a2 = a + h, a1 + d1, a2 + d2
            elif n_out == 2 * n_in:
                a = a + torch.cat([h, h], 1)
                a1 = a1 + torch.cat([d1, d1], 1)
                a2 = a2 + torch.cat([d2, d2], 1)
            h, d1, d2 = a, a1, a2
        return h, d1, d2

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        return self.value_and_derivatives(x)[0]


def spline5_switch(r: float, rmin: float, rmax: float) -> float:
    if r < rmin:
        return 1.0
    if r < rmax:
        uu = (r - rmin) / (rmax - rmin)
        return uu * uu * uu * (-6 * uu * uu + 15 * uu - 10) + 1
    return 0.0


def env_mat_range(davg: np.ndarray, dstd: np.ndarray, min_nbor_dist: float, rcut_smth: float, rcut: float):
    """deepmd/utils/tabulate.py:468-486 + :94-96.  davg / dstd: the statistics of neighbour SLOT 0, [ntypes, 4] (the
    reference indexes `davg[:, 0]` on its [ntypes, nnei, 4] arrays, i.e. all four components of the first slot): per
    centre type floor(-davg/dstd) and ceil((sw/rmin - davg)/dstd) element-wise, then the minimum / maximum over the
    components.  With the usual statistics (davg = 0, smaller dstd on the angular components) the ANGULAR components
    set the upper boundary -- checked against the table the reference's own `enable_compression` builds
    (tests/golden/dpa1_strip.json: lower -1, upper 15 for the water statistics)."""
    davg = np.asarray(davg, np.float64).reshape(len(davg), -1)
    dstd = np.asarray(dstd, np.float64).reshape(len(dstd), -1)
    sw = spline5_switch(min_nbor_dist, rcut_smth, rcut)
    lower = np.floor(-davg / dstd)
    upper = np.ceil(((1.0 / min_nbor_dist) * sw - davg) / dstd)
    return lower.min(axis=1), upper.max(axis=1)


def build_table(net: EmbeddingNet, lower: float, upper: float, stride0: float, stride1: float, extrapolate: float,
                device="cpu") -> torch.Tensor:
    """tabulate.py:108-126 + 275-371: [nspline, 6*M] float64, coefficients innermost."""
    xx = np.arange(lower, upper, stride0, dtype=np.float64)
    xx = np.append(xx, np.arange(upper, extrapolate * upper, stride1, dtype=np.float64))
    xx = np.append(xx, np.array([extrapolate * upper], dtype=np.float64))
    nspline = int((upper - lower) / stride0 + (extrapolate * upper - upper) / stride1)
    v, d, d2 = net.value_and_derivatives(torch.as_tensor(xx, device=device))
    M = v.shape[1]
    tt = torch.full((nspline, 1), stride1, dtype=torch.float64, device=v.device)
    tt[: int((upper - lower) / stride0)] = stride0
    v0, v1 = v[:nspline], v[1:nspline + 1]
    d0, d1 = d[:nspline], d[1:nspline + 1]
    s0, s1 = d2[:nspline], d2[1:nspline + 1]
    hh = v1 - v0
    tab = torch.zeros(nspline, M, 6, dtype=torch.float64, device=v.device)
    tab[:, :, 0] = v0
    tab[:, :, 1] = d0
    tab[:, :, 2] = 0.5 * s0
    tab[:, :, 3] = (1 / (2 * tt * tt * tt)) * (20 * hh - (8 * d1 + 12 * d0) * tt - (3 * s0 - s1) * tt * tt)
    tab[:, :, 4] = (1 / (2 * tt * tt * tt * tt)) * (-30 * hh + (14 * d1 + 16 * d0) * tt + (3 * s0 - 2 * s1) * tt * tt)
    tab[:, :, 5] = (1 / (2 * tt * tt * tt * tt * tt)) * (12 * hh - 6 * (d1 + d0) * tt + (s1 - s0) * tt * tt)
    return tab.reshape(nspline, 6 * M)


def compress_se_a(nets: Sequence[EmbeddingNet], davg: np.ndarray, dstd: np.ndarray, sel: Sequence[int],
                  min_nbor_dist: float, rcut_smth: float, rcut: float, stride0: float = 0.01, stride1: float = 0.1,
                  extrapolate: float = 5.0, check_frequency: float = -1.0, device="cpu"):
    """type_one_side compression: one table per NEIGHBOUR type (net `filter_-1_net_<ii>`), range
    over all centre types with sel > 0 (tabulate.py:118-131).  Returns (tables, infos)."""
    lower, upper = env_mat_range(np.asarray(davg, np.float64), np.asarray(dstd, np.float64), min_nbor_dist, rcut_smth,
                                 rcut)
    idx = [s > 0 for s in sel]
    uu = float(np.max(upper[idx]))
    ll = float(np.min(lower[idx]))
    tables, infos = [], []
    for net in nets:
        tables.append(build_table(net, ll, uu, stride0, stride1, extrapolate, device))
        infos.append(torch.tensor([ll, uu, uu * extrapolate, stride0, stride1, check_frequency], dtype=torch.float64))
    return tables, infos
