"""TEST INFRASTRUCTURE ONLY — the model data (weights, statistics, tables) of the benchmark's reference arm, built
WITHOUT importing the product package (so that `bench.py --impl reference` loads none of the product's native
code).  Same seeds and the same initialisation recipe as deepmd_kit_b200.model.SeAModel / FittingNet (reference
default init deepmd/pt/model/network/mlp.py:149-163: weights N(0, 1/sqrt(in + out)), biases N(0, 1), idt
N(0.1, 0.001)); the table builder is the pure-torch data producer deepmd-kit_b200/compress.py (the `dp compress`
restatement), loaded by file path.  tests/test_host_cpu.py checks that both arms hold bit-identical weights and tables.
"""
from __future__ import annotations

import importlib.util
import math
import os
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WATER_STATS = [(0.05033, 0.13984, 0.08580), (0.04810, 0.12388, 0.07672)]


def _compress_module():
    spec = importlib.util.spec_from_file_location("_dpb200_compress", os.path.join(ROOT, "deepmd-kit_b200", "compress.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class RefFittingNet:
    """y = tanh(x W + b) * idt (+ x); energy head (deepmd/pt/model/network/mlp.py)."""

    def __init__(self, dim_in, neuron, resnet_dt, seed, dtype):
        g = torch.Generator().manual_seed(seed)
        self.layers = []
        n_in = dim_in
        for n_out in neuron:
            w = torch.empty(n_in, n_out, dtype=torch.float64).normal_(0.0, 1.0 / math.sqrt(n_in + n_out), generator=g)
            b = torch.empty(n_out, dtype=torch.float64).normal_(0.0, 1.0, generator=g)
            idt = torch.empty(n_out, dtype=torch.float64).normal_(0.1, 0.001, generator=g) if resnet_dt else None
            self.layers.append((w.to(dtype), b.to(dtype), None if idt is None else idt.to(dtype)))
            n_in = n_out
        w = torch.empty(n_in, 1, dtype=torch.float64).normal_(0.0, 1.0 / math.sqrt(n_in + 1), generator=g)
        b = torch.empty(1, dtype=torch.float64).normal_(0.0, 1.0, generator=g)
        self.head = (w.to(dtype), b.to(dtype))

    def __call__(self, x):
        for w, b, idt in self.layers:
            y = torch.tanh(torch.addmm(b, x, w))
            if idt is not None:
                y = y * idt
            if w.shape[0] == w.shape[1]:
                y = y + x
            elif w.shape[1] == 2 * w.shape[0]:
                y = y + torch.cat([x, x], 1)
            x = y
        return torch.addmm(self.head[1], x, self.head[0]).reshape(-1)


class RefWaterModel:
    """examples/water/se_e2_a/input.json architecture, random-init (seed 1), compressed: what oracle.pipeline.evaluate
    needs (cfg, dtype, davg, dstd, tables, infos, M, fit, bias_atom_e)."""

    def __init__(self, dtype=torch.float64, seed=1, min_nbor_dist=0.9):
        cz = _compress_module()
        sel = (46, 92)
        sec = [0, 46, 138]
        self.cfg = SimpleNamespace(ntypes=2, sel=sel, sec=sec, nnei=138, rcut=6.0, rcut_smth=0.5, neuron=(25, 50, 100),
                                   axis_neuron=16, fitting_neuron=(240, 240, 240), fitting_resnet_dt=True, seed=seed)
        self.dtype = dtype
        davg = np.zeros((2, 138, 4))
        dstd = np.ones((2, 138, 4))
        for t, (a0, s0, s1) in enumerate(WATER_STATS):
            davg[t, :, 0] = a0
            dstd[t, :, 0] = s0
            dstd[t, :, 1:] = s1
        self.davg = torch.as_tensor(davg.reshape(2, -1), dtype=dtype)
        self.dstd = torch.as_tensor(dstd.reshape(2, -1), dtype=dtype)
        embed = [cz.EmbeddingNet(self.cfg.neuron, seed + 17 * t) for t in range(2)]
        tables, infos = cz.compress_se_a(embed, davg[:, 0, :], dstd[:, 0, :], sel, min_nbor_dist, 0.5, 6.0, 0.01, 0.1, 5.0)
        self.tables = [t.to(dtype).contiguous() for t in tables]
        self.infos = [i.to(dtype) for i in infos]
        self.M = 100
        self.fit = [RefFittingNet(1600, self.cfg.fitting_neuron, True, seed + 101 * t, dtype) for t in range(2)]
        self.bias_atom_e = torch.zeros(2, dtype=dtype)


class RefAttnModel:
    """DPA-1 se_atten_v2 WITH attention layers (strip, smooth, attn_layer 2, attn 128, dotr, sel 120; the architecture
    of deepmd_kit_b200.atten.SeAttenConfig(attn_layer=2)), random-init: what oracle.pipeline_atten.evaluate_layers
    needs.  Used by the CPU timing legs of bench.py (`--workload dpa1_attn`): the arithmetic does not depend on the
    weight values, so this model has its own seeds (it is NOT weight-identical to the product's random model; parity
    for this path is pinned by tests/golden/dpa1_attn.npz instead)."""

    def __init__(self, dtype=torch.float64, seed=1, attn_layer=2):
        cz = _compress_module()
        nnei, M, tebd_dim, attn, nt = 120, 100, 8, 128, 2
        self.cfg = SimpleNamespace(ntypes=nt, nsel=nnei, sel=(nnei,), sec=[0, nnei], nnei=nnei, rcut=6.0, rcut_smth=0.5,
                                   neuron=(25, 50, 100), axis_neuron=16, tebd_dim=tebd_dim, fitting_neuron=(240, 240, 240),
                                   fitting_resnet_dt=True, seed=seed, attn_layer=attn_layer, attn=attn, attn_dotr=True,
                                   attn_normalize=True, scaling_factor=1.0, ln_eps=1e-5, attnw_shift=20.0)
        self.dtype = dtype
        self.M = M
        davg = np.zeros((nt, nnei, 4))
        dstd = np.ones((nt, nnei, 4))
        for t, (a0, s0, s1) in enumerate(WATER_STATS):
            davg[t, :, 0] = a0
            dstd[t, :, 0] = s0
            dstd[t, :, 1:] = s1
        self.davg = torch.as_tensor(davg.reshape(nt, -1), dtype=dtype)
        self.dstd = torch.as_tensor(dstd.reshape(nt, -1), dtype=dtype)
        self.embed = cz.EmbeddingNet(self.cfg.neuron, seed)
        g = torch.Generator().manual_seed(seed + 7)

        def normal(*shape, std=1.0):
            return torch.empty(*shape, dtype=torch.float64).normal_(0.0, std, generator=g)

        self.tebd = torch.zeros(nt + 1, tebd_dim, dtype=torch.float64)
        self.tebd[:nt] = normal(nt, tebd_dim)
        self.tt_full = 0.1 * normal((nt + 1) ** 2, M)  # the strip net's output table (values irrelevant for timing)
        self.attn_scaling = float(attn ** -0.5)
        self.attn_layers = []
        for _ in range(attn_layer):
            self.attn_layers.append(dict(in_w=normal(M, 3 * attn, std=1.0 / math.sqrt(M + 3 * attn)), in_b=normal(3 * attn),
                                         out_w=normal(attn, M, std=1.0 / math.sqrt(attn + M)), out_b=normal(M),
                                         ln_w=torch.ones(M, dtype=torch.float64), ln_b=torch.zeros(M, dtype=torch.float64)))
        self.dim_d = M * self.cfg.axis_neuron
        self.dim_in = (self.dim_d + tebd_dim + 15) // 16 * 16
        self.fit = RefFittingNet(self.dim_in, self.cfg.fitting_neuron, True, seed + 101, torch.float64)
        self.bias_atom_e = torch.zeros(nt, dtype=torch.float64)
