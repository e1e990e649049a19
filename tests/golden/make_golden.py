#!/usr/bin/env python
"""Extract the literal golden vectors of the reference's own unit tests for the hot path
 into small JSON fixtures under tests/golden/.

Run in the build container (the reference tree is mounted read-only at /root/reference):

    python tests/golden/make_golden.py [--ref /root/reference]

Only *data* (numeric literals of `std::vector<T> name = {...};` members of the gtest
fixtures) is extracted — no reference code is copied.  Each fixture records the file and
line it came from so a reviewer can check it.
"""
import argparse
import json
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))

# (reference file, [fixture classes]) -> output json
SOURCES = {
    "env_mat_a.json": ("source/lib/tests/test_env_mat_a.cc", ["TestEnvMatA"]),
    "fmt_nlist.json": ("source/lib/tests/test_fmt_nlist.cc", ["TestFormatNlist", "TestFormatNlistShortSel"]),
    "tabulate_se_a.json": ("source/lib/tests/test_tabulate_se_a.cc", ["TestTabulateSeA"]),
    "prod_force_a.json": ("source/lib/tests/test_prod_force_a.cc", ["TestProdForceA"]),
    "prod_virial_a.json": ("source/lib/tests/test_prod_virial_a.cc", ["TestProdVirialA"]),
    "prod_force_grad_a.json": ("source/lib/tests/test_prod_force_grad_a.cc", ["TestProdForceGradA"]),
    "prod_virial_grad_a.json": ("source/lib/tests/test_prod_virial_grad_a.cc", ["TestProdVirialGradA"]),
    "coord.json": ("source/lib/tests/test_coord.cc", ["TestNormCoord", "TestCopyCoord", "TestCopyCoordMoreCell"]),
    "neighbor_list.json": ("source/lib/tests/test_neighbor_list.cc", ["TestNeighborList"]),
}

VEC_RE = re.compile(r"std::vector<\s*(double|int|float)\s*>\s+(\w+)\s*=\s*\{([^;]*?)\}\s*;", re.S)
SCALAR_RE = re.compile(r"^\s*(?:const\s+|static\s+|constexpr\s+)*(double|int|float)\s+(\w+)\s*=\s*([-+0-9.eE]+)\s*;", re.M)
NESTED_RE = re.compile(r"std::vector<\s*std::vector<\s*int\s*>\s*>\s+(\w+)\s*=\s*\{(.*?)\}\s*;", re.S)


def class_bodies(text):
    """yield (class name, body text, first line number) for every gtest fixture class."""
    for m in re.finditer(r"class\s+(\w+)\s*:\s*public\s+::testing::Test\s*\{", text):
        depth, i = 1, m.end()
        while depth and i < len(text):
            depth += {"{": 1, "}": -1}.get(text[i], 0)
            i += 1
        yield m.group(1), text[m.end():i], text.count("\n", 0, m.start()) + 1


def parse_numbers(body, kind):
    toks = [t for t in re.split(r"[\s,]+", re.sub(r"//[^\n]*", "", body)) if t]
    if kind == "int":
        return [int(t, 0) if not t.startswith("0") or t == "0" else int(t) for t in toks]
    return [float(t) for t in toks]


def extract(ref_root, rel, classes):
    text = open(os.path.join(ref_root, rel)).read()
    out = {"_source": rel}
    for name, body, line in class_bodies(text):
        if name not in classes:
            continue
        entry = {"_line": line}
        for m in VEC_RE.finditer(body):
            kind, var, lit = m.groups()
            if not lit.strip():
                continue
            try:
                entry[var] = parse_numbers(lit, kind)
            except ValueError:
                continue  # non-literal initialiser
        for m in NESTED_RE.finditer(body):
            var, lit = m.groups()
            rows = re.findall(r"\{([^{}]*)\}", lit)
            entry[var] = [parse_numbers(r, "int") for r in rows]
        for m in SCALAR_RE.finditer(body):
            kind, var, lit = m.groups()
            entry.setdefault(var, int(lit) if kind == "int" else float(lit))
        out[name] = entry
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    args = ap.parse_args()
    for fname, (rel, classes) in SOURCES.items():
        data = extract(args.ref, rel, classes)
        missing = [c for c in classes if c not in data]
        path = os.path.join(HERE, fname)
        with open(path, "w") as f:
            json.dump(data, f, separators=(",", ":"))
        sizes = {c: {k: (len(v) if isinstance(v, list) else v) for k, v in data[c].items()} for c in classes if c in data}
        print(f"{fname}: {sizes} missing={missing}")
    water192(args.ref)


def water192(ref):
    """BASELINE config 1: frame 0 of examples/water/data/data_0 (192 atoms, cubic box)."""
    import numpy as np

    d = os.path.join(ref, "examples", "water", "data", "data_0")
    coord = np.load(os.path.join(d, "set.000", "coord.npy"))[0].astype(np.float64)
    box = np.load(os.path.join(d, "set.000", "box.npy"))[0].astype(np.float64)
    atype = [int(x) for x in open(os.path.join(d, "type.raw")).read().split()]
    out = {"_source": "examples/water/data/data_0/set.000 frame 0 (float32 data cast to float64)",
           "coord": coord.tolist(), "box": box.tolist(), "atype": atype}
    with open(os.path.join(HERE, "water192.json"), "w") as f:
        json.dump(out, f)
    print(f"water192.json: {len(atype)} atoms, box {box.reshape(3, 3).diagonal()}")


if __name__ == "__main__":
    main()
