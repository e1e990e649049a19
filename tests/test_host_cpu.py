"""CPU-only checks of the boundary and the host logic (no compute calls into CUDA)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import __graft_entry__ as g

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def pkg():
    g.build()
    return g.load_package()


def test_c_abi_exports_every_declared_symbol(pkg):
    header = open(os.path.join(ROOT, "include", "dpb200.h")).read()
    # expand the DECL macros by substituting both suffixes
    names = set(re.findall(r"\b(dpb200_[a-z0-9_]+)\s*\(", header))
    for macro_name in re.findall(r"\b(dpb200_[a-z0-9_]+)_##SUF", header):
        names.add(macro_name + "_f32")
        names.add(macro_name + "_f64")
    names = {n for n in names if not n.endswith("_")}
    lib = ctypes.CDLL(pkg._lib.LIB_PATH)
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, missing
    assert len(names) >= 35
    assert set(pkg._lib.lib().exported()) <= names
    assert lib.dpb200_abi_version() == 1


def test_no_cpu_fallback(pkg):
    """CPU tensors are rejected loudly, and without a device the compute entry points fail."""
    t = torch.zeros(2, 12, dtype=torch.float64)
    with pytest.raises(RuntimeError, match="no CPU implementation"):
        torch.ops.deepmd.tabulate_fusion_se_a(t, torch.zeros(6, dtype=torch.float64), torch.zeros(4, 1, dtype=torch.float64),
                                              torch.zeros(1, 4, 4, dtype=torch.float64), 2)
    with pytest.raises(RuntimeError, match="no CPU implementation"):
        pkg.ops.prod_force_a(torch.zeros(1, 8), torch.zeros(1, 24), torch.zeros(1, 2, dtype=torch.int32), 1, 1, 2)
    with pytest.raises(ValueError):
        pkg.ops._op_tabulate_fusion_se_a(torch.zeros(2), torch.zeros(6), torch.zeros(4, 1), torch.zeros(1, 4, 4), 2)


def test_product_does_not_import_oracle():
    pk = os.path.join(ROOT, "deepmd-kit_b200")
    for dirpath, _, files in os.walk(pk):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("oracle/", "").lower() or f == "__init__.py" and False, \
                    f"{f} mentions the oracle"


def test_compress_table_reproduces_the_embedding_net(pkg):
    """SURVEY 8c self-check: the tabulated quintic agrees with the net it was built from (stride 0.01:
    well below the 10-decimal contract of source/tests/pt/test_model_compression_se_a.py:22-24 is not
    reachable by a quintic, the reference asserts 1e-10 on energies only; we check 1e-8 per channel)."""
    from deepmd_kit_b200.compress import EmbeddingNet, build_table

    net = EmbeddingNet((25, 50, 100), seed=3)
    lower, upper, s0, s1, ex = -1.0, 9.0, 0.01, 0.1, 5.0
    tab = build_table(net, lower, upper, s0, s1, ex).reshape(-1, 100, 6).numpy()
    nfirst = int((upper - lower) / s0)
    assert tab.shape[0] == int((upper - lower) / s0 + (ex * upper - upper) / s1)
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(lower, upper, 200), rng.uniform(upper, ex * upper - 0.2, 200)])
    want = net(torch.as_tensor(x)).numpy()
    idx = np.where(x < upper, ((x - lower) / s0).astype(int), nfirst + ((x - upper) / s1).astype(int))
    base = np.where(x < upper, lower + idx * s0, upper + (idx - nfirst) * s1)
    dx = (x - base)[:, None]
    a = tab[idx]
    got = a[..., 0] + (a[..., 1] + (a[..., 2] + (a[..., 3] + (a[..., 4] + a[..., 5] * dx) * dx) * dx) * dx) * dx
    assert np.abs(got[:200] - want[:200]).max() < 1e-8
    assert np.abs(got[200:] - want[200:]).max() < 1e-5
    # derivative continuity at a knot: value and slope of adjacent rows agree
    r = 37
    t = s0
    end = a = tab[r]
    v_end = end[:, 0] + (end[:, 1] + (end[:, 2] + (end[:, 3] + (end[:, 4] + end[:, 5] * t) * t) * t) * t) * t
    assert np.abs(v_end - tab[r + 1][:, 0]).max() < 1e-12


def test_mesh_decoding_and_type_partition(pkg):
    from deepmd_kit_b200.model import type_partition

    ops = pkg.ops
    assert ops.decode_mesh(torch.zeros(0, dtype=torch.int32), 3)[0] == "nopbc"
    assert ops.decode_mesh(torch.zeros(6, dtype=torch.int32), 3)[0] == "pbc"
    with pytest.raises(ValueError):
        ops.decode_mesh(torch.zeros(7, dtype=torch.int32), 3)
    # inline list (source/op/tf/prod_env_mat_multi_device.cc:2813-2828)
    ilist = np.array([0, 1, 2], np.int32)
    numneigh = np.array([2, 0, 1], np.int32)
    neigh = np.array([1, 2, 0], np.int32)
    head = np.zeros(16, np.int32)
    head[1] = 3
    mode, il, nn, ne = ops.decode_mesh(torch.as_tensor(np.concatenate([head, ilist, numneigh, neigh])), 3)
    assert mode == "list" and il.tolist() == [0, 1, 2] and nn.tolist() == [2, 0, 1] and ne.tolist() == [1, 2, 0]
    # LAMMPS-style host pointers packed into 16 ints (source/api_cc/src/common.cc:786-798)
    rows = [np.array([1, 2], np.int32), np.zeros(0, np.int32), np.array([0], np.int32)]
    first = (ctypes.POINTER(ctypes.c_int) * 3)(*[r.ctypes.data_as(ctypes.POINTER(ctypes.c_int)) for r in rows])
    mesh = np.zeros(16, np.int32)
    mesh[1] = 3
    for k, addr in ((4, ilist.ctypes.data), (8, numneigh.ctypes.data), (12, ctypes.addressof(first))):
        mesh[k:k + 2] = np.frombuffer(np.uint64(addr).tobytes(), dtype=np.int32)
    mode, il, nn, ne = ops.decode_mesh(torch.as_tensor(mesh), 3)
    assert mode == "list" and il.tolist() == [0, 1, 2] and nn.tolist() == [2, 0, 1] and ne.tolist() == [1, 2, 0]
    perm, ranges = type_partition(torch.tensor([1, 0, 1, 1, 0], dtype=torch.int32), 2)
    assert perm.tolist() == [1, 4, 0, 2, 3] and ranges == [(0, 2), (2, 5)]


def test_model_tables_and_cpu_reference_pipeline(pkg, port):
    """The CPU pipeline that bench.py times as the reference arm: translation invariance and
    F = -dE/dx by central differences (size-independent physical checks of the composition)."""
    from deepmd_kit_b200.model import SeAConfig, SeAModel
    from oracle import pipeline

    coord, atype, box = g.water_box(1)
    cfg = SeAConfig()
    m = SeAModel(cfg, torch.float64, "cpu")
    # (lower -1, upper 15: the range the reference's own `dp compress` produces for these statistics -- the angular
    #  components of slot 0 set the upper boundary; tests/test_reference_model.py pins it to the reference's table)
    assert m.tables[0].shape == (2200, 600) and m.infos[0].tolist()[:5] == [-1.0, 15.0, 75.0, 0.01, 0.1]
    lists = pipeline.build_lists(port, coord, atype, box, 8.0)
    e0, f0, v0, ex = pipeline.evaluate(port, m, lists)
    assert np.abs(f0.sum(0)).max() < 1e-12
    assert np.abs(v0.reshape(3, 3) - v0.reshape(3, 3).T).max() < 1e-10
    h = 1e-5
    for i, d in ((0, 0), (77, 2)):
        es = []
        for sgn in (1, -1):
            c = coord.copy()
            c[i, d] += sgn * h
            es.append(pipeline.evaluate(port, m, pipeline.build_lists(port, c, atype, box, 8.0))[0])
        assert abs(-(es[0] - es[1]) / (2 * h) - f0[i, d]) < 1e-7


def test_split_i8_cols_reconstructs_the_weights(pkg):
    """Host-side operand split of the int8 tensor-core GEMM (model.split_i8_cols): balanced base-256 digits,
    most significant first, per-column exponent; the reconstruction error is below 2^-(7+8(ns-1)) of the
    column scale and the order-truncated product matches the fp64 product."""
    from deepmd_kit_b200.model import split_i8_cols

    torch.manual_seed(0)
    w = torch.randn(64, 24, dtype=torch.float64) * torch.logspace(-3, 2, 24, dtype=torch.float64)[None, :]
    w[:, 3] = 0.0  # an all-zero column must not produce NaN / overflow
    for ns in (2, 6, 7):
        sl, ce = split_i8_cols(w, ns)
        assert sl.shape == (ns, 64, 24) and sl.dtype == torch.int8 and ce.dtype == torch.int32
        assert int(sl.to(torch.int32).abs().max()) <= 128
        scale = torch.ldexp(torch.ones(24, dtype=torch.float64), ce)
        rec = sum(sl[s].double() * 2.0 ** (-7 - 8 * s) for s in range(ns)) * scale[None, :]
        err = ((rec - w).abs() / scale[None, :]).max().item()
        assert err <= 2.0 ** (-7 - 8 * (ns - 1)) * 0.51, (ns, err)
    # K-concatenated order sums == exact integer products, recombined
    ns = 6
    x = torch.randn(5, 64, dtype=torch.float64)
    xs, xe = split_i8_cols(x.t().contiguous(), ns)  # per-row scale of x == per-column scale of x^T
    xs = xs.permute(0, 2, 1).contiguous()  # [ns, 5, 64]
    sl, ce = split_i8_cols(w, ns)
    acc = torch.zeros(5, 24, dtype=torch.float64)
    for d in range(ns):
        od = sum(xs[i].to(torch.int64) @ sl[d - i].to(torch.int64) for i in range(d + 1))
        assert int(od.abs().max()) < 2 ** 31  # fits the tensor cores' int32 accumulators
        acc += od.double() * 2.0 ** (-14 - 8 * d)
    got = acc * torch.ldexp(torch.ones(5, dtype=torch.float64), xe)[:, None] * torch.ldexp(
        torch.ones(24, dtype=torch.float64), ce)[None, :]
    want = x @ w
    tol = 2e-13 * (x.abs().amax(1, keepdim=True) * w.abs().amax(0, keepdim=True) * 8)
    assert bool(((got - want).abs() <= tol).all())


def test_tf32_split_is_exact_to_2pow22(pkg):
    from deepmd_kit_b200.model import split_tf32_weight

    torch.manual_seed(1)
    w = torch.randn(1000, dtype=torch.float32) * torch.logspace(-8, 8, 1000)
    hi, lo = split_tf32_weight(w)
    for part in (hi, lo):  # 10 explicit mantissa bits: the low 13 bits of the fp32 pattern are clear
        assert int((part.view(torch.int32) & 0x1FFF).abs().max()) == 0
    assert float(((hi.double() + lo.double() - w.double()).abs() / w.double().abs()).max()) <= 2.0 ** -21


def test_compressed_coefficient_gates(pkg):
    """Host-side gates of DPB200_TAB_COMPRESSED_COEF: dp-compress tables of the benchmark models qualify (fp64 and fp32
    forms), an arbitrary table with O(1) high-order coefficients and coarse strides does not."""
    from deepmd_kit_b200 import ops
    from deepmd_kit_b200.model import COPPER_CONFIG, SeAConfig, SeAModel

    for cfg in (SeAConfig(), SeAConfig(**COPPER_CONFIG)):
        m = SeAModel(cfg, torch.float64, "cpu")
        assert m.coef_flags is None  # never enabled off the GPU path
        for t, i in zip(m.tables64, m.infos):
            assert ops.compressed_coef_flags(t, i) & 1
            assert ops.compressed_coef_flags_f32(t, i) == 1
            # the fp16 scale of a5 keeps the largest stride-0 coefficient inside the fp16 range
            k = ((ops.compressed_coef_flags(t, i) >> 8) & 0xFF)
            k = k - 256 if k >= 128 else k
            first = int((float(i[1]) - float(i[0])) / float(i[3]))
            a5 = t.reshape(t.shape[0], -1, 6)[:first, :, 5].abs().max().item()
            assert 2.0 ** 12 <= a5 * 2.0 ** k < 2.0 ** 15
    rng = np.random.default_rng(0)
    info = np.array([-0.4, 2.0, 6.0, 0.05, 0.5, -1.0])
    nrow = int((info[1] - info[0]) / info[3]) + int((info[2] - info[1]) / info[4]) + 1
    bad = torch.as_tensor(rng.normal(size=(nrow, 100 * 6)))
    assert ops.compressed_coef_flags(bad, info) == 0
    assert ops.compressed_coef_flags_f32(bad, info) == 0
    assert ops.compressed_coef_flags(torch.zeros(nrow, 600, dtype=torch.float64), info) == 0


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` prints ONE JSON line with the contract's keys (CPU only, tiny sample)."""
    import json
    import subprocess
    import sys

    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup",
                        "1", "--cpu-ncopy", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "us/step/atom" and d["higher_is_better"] is False
    assert d["value"] > 0 and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["natoms"] == 1536000 and "workload" in d["config"]


def test_bench_reference_arm_attention_workload():
    """`bench.py --impl reference --workload dpa1_attn`: the dense CPU path of DPA-1 with attention layers (checker +
    oracle/refmodel.RefAttnModel, no product native code) prints the same contract line."""
    import json
    import subprocess
    import sys

    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "dpa1_attn",
                        "--steps", "1", "--warmup", "0", "--cpu-ncopy", "1", "--ncopy", "6"], capture_output=True, text=True,
                       timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "us/step/atom" and d["higher_is_better"] is False
    assert d["value"] > 0 and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    assert d["config"]["natoms"] == 192 * 6 ** 3 and d["config"]["sample_natoms"] == 192
    assert d["config"]["bench_workload"] == "dpa1_attn" and d["product_native_loaded"] is False


def test_reference_arm_model_equals_product_model(pkg):
    """bench.py --impl reference builds its model data without importing the product (oracle/refmodel.py): weights,
    statistics and tables must be bit-identical to the product's SeAModel, or the two arms would time different work."""
    from deepmd_kit_b200.model import SeAConfig, SeAModel
    from oracle.refmodel import RefWaterModel

    a = SeAModel(SeAConfig(), torch.float64, "cpu")
    b = RefWaterModel(torch.float64)
    assert torch.equal(a.davg, b.davg) and torch.equal(a.dstd, b.dstd)
    for ta, tb in zip(a.tables, b.tables):
        assert torch.equal(ta, tb)
    for ia, ib in zip(a.infos, b.infos):
        assert torch.equal(ia, ib)
    for fa, fb in zip(a.fit, b.fit):
        for (w0, b0, i0), (w1, b1, i1) in zip(fa.layers, fb.layers):
            assert torch.equal(w0, w1) and torch.equal(b0, b1) and torch.equal(i0, i1)
        assert torch.equal(fa.head[0], fb.head[0]) and torch.equal(fa.head[1], fb.head[1])
    x = torch.randn(7, 1600, dtype=torch.float64)
    assert torch.equal(fa(x), fb(x))


def test_decomposed_copper_box_covers_the_global_box(pkg):
    """Strong-scaling workload (BASELINE config 3): the bricks generated rank by rank tile the global FCC box exactly
    and every atom lies inside its rank's brick."""
    import __graft_entry__ as g
    from deepmd_kit_b200.domain import DomainDeepPot, rank_to_coords

    grid = (2, 2, 1)
    want, _, box = g.copper_box(4, jitter=0.0)
    got = []
    for r in range(4):
        dd = DomainDeepPot(None, grid)
        dd.rank = r
        c, t, b = dd.make_local_copper(4, jitter=0.0)
        assert np.array_equal(b, box) and len(t) == len(c) and not t.any()
        me = np.array(rank_to_coords(r, grid))
        L = np.diag(box)
        assert ((c >= me * L / np.array(grid) - 1e-9) & (c < (me + 1) * L / np.array(grid) - 1e-9)).all()
        got.append(c)
    got = np.concatenate(got)
    key = lambda a: sorted(map(tuple, np.round(a, 6)))
    assert key(got) == key(want)
    with pytest.raises(ValueError):
        DomainDeepPot(None, (3, 1, 1)).make_local_copper(4)
