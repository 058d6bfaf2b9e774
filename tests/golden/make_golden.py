"""Regenerates the golden fixtures in this directory from the reference.

Run in the build container only (it reads /root/reference, which does not
exist on the GPU box):   python tests/golden/make_golden.py

Outputs
  r_design_100x2.json    the R MaxProLHD design and the values the reference
                         pins on it (python/comparison_r.py:14-115,131 and
                         README.md:89-90).
  pymaxpro_vectors.json  designs scored by the reference author's own NumPy
                         restatement, python/pymaxpro.py:7-62, imported and
                         executed here (maxpro_criterion + calculate_true_maxpro).
"""
import ast
import importlib.util
import json
import os
import re

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def r_design():
    src = open(os.path.join(REF, "python/comparison_r.py")).read()
    tree = ast.parse(src)
    vals = {}
    for node in tree.body:
        if isinstance(node, ast.Assign) and isinstance(node.targets[0], ast.Name):
            name = node.targets[0].id
            if name in ("R_DESIGN", "R_MEASURE"):
                vals[name] = ast.literal_eval(node.value)
    design = [[float(t) for t in line.split(" ")] for line in vals["R_DESIGN"].split("\n")]
    readme = open(os.path.join(REF, "README.md")).read()
    m = re.search(r"R calculation: [0-9.]+ *\nRust calculation: ([0-9.]+)", readme)  # README.md:89-90
    rust_value = float(m.group(1)) if m else None
    assert len(design) == 100 and all(len(r) == 2 for r in design)
    return {
        "source": "python/comparison_r.py:14-115 (R_DESIGN, R_MEASURE); README.md:89-90 (Rust value)",
        "design": design,
        "r_measure": vals["R_MEASURE"],
        "r_rel_tol": 1e-7,
        "rust_value_readme": rust_value,
    }


def pymaxpro_vectors():
    spec = importlib.util.spec_from_file_location("pymaxpro", os.path.join(REF, "python/pymaxpro.py"))
    pm = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(pm)
    rng = np.random.default_rng(42)  # SEED = 42, python/comparisons.py:8
    cases = []
    for (n, d) in [(2, 2), (5, 2), (10, 2), (10, 4), (50, 3), (100, 4), (30, 10), (64, 1), (100, 5)]:
        # a random LHD of the reference's shape (lhd.rs:117-128): permuted strata + in-stratum jitter
        x = np.empty((n, d))
        for k in range(d):
            x[:, k] = (rng.permutation(n) + rng.random(n)) / n
        s = float(pm.maxpro_criterion(x))
        psi = float(pm.calculate_true_maxpro(s, n, d))
        cases.append({"n": n, "d": d, "design_hex": [[v.hex() for v in row] for row in x.tolist()],
                      "maxpro_sum": s, "maxpro_criterion": psi})
    return {"source": "python/pymaxpro.py:7-62 executed on seeded NumPy LHDs (np.random.default_rng(42))",
            "rel_tol": 1e-12, "cases": cases}


if __name__ == "__main__":
    json.dump(r_design(), open(os.path.join(HERE, "r_design_100x2.json"), "w"), indent=0)
    json.dump(pymaxpro_vectors(), open(os.path.join(HERE, "pymaxpro_vectors.json"), "w"), indent=0)
    print("golden fixtures written")
