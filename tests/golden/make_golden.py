"""Generates the golden fixtures in this directory from the CPU oracle (the reference itself cannot be
built in this image and ships no vectors — SURVEY.md §8c).  Run: python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from tests import common  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = {"c1_n50_w": ("c1", 0.005, "0.3"), "c1r_n200": ("c1r", 0.02, "zero"), "c2_n300_w": ("c2", 0.003, "0.4"),
         "c3_n500_w": ("c3", 0.0005, "0.3")}

for key, (name, scale, weights) in CASES.items():
    case = common.make_case(name, scale, weights)
    orc = common.oracle_for(O, case)
    f, g, H = orc.kantorovich(case["w"], mode=1)
    H.sort_indices()
    cfg = case["cfg"]
    meta = dict(kind=cfg["kind"], n=cfg.get("n", 0), m=cfg.get("m", 0))
    np.savez_compressed(os.path.join(HERE, key + ".npz"), vx=cfg["vx"], vy=cfg["vy"], tri=cfg["tri"], abc=case["abc"],
                        rho=cfg["rho"], X=case["X"], w=case["w"], f=f, g=g, H_data=H.data, H_indices=H.indices,
                        H_indptr=H.indptr, kind=meta["kind"], n=meta["n"], m=meta["m"],
                        counters=np.array([orc.counters()[k] for k in O.COUNTER_NAMES]))
    print(key, "N", case["N"], "nnz", H.nnz, "f", f)
