#!/usr/bin/env python
"""Golden values of the BASELINE.json configs AT THEIR STATED SIZE, produced by the CPU oracle.

    python tests/golden/make_config_golden.py [C1 C2 C3 C4 C4@2 C4@4 C4@8 C4g C5@8 ...]   ->  tests/golden/config_golden.json

For every case: iterations, converged, final_residual (as float.hex, so the comparison is bit-exact) and, per
row-block shard of the solution (chunk partition asm.rs:46-57 with the case's shard count), the SHA-256 of the
little-endian f64 bytes plus the oracle's canonical sum.  The GPU parity tests (tests/test_gpu_configs_full.py)
and bench.py's `parity` block compare against this file, so no oracle run is needed on the GPU box for the
big cases.  Entries are merged into an existing file; re-running a case overwrites it.

Oracle cost (8 cores): C1 5 s, C3 10 s, C4 52 s per shard count, C2 64 s, C4g ~10 min, C5@8 (two restart
cycles only, max_iters = 100) ~15 min.
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
OUT = os.path.join(HERE, "config_golden.json")

TOL = 1e-8
MAX_ITERS = 20000


def shard_digests(x, p):
    import oracle_ffi as o
    out = []
    for r in range(p):
        lo, hi = o.partition_range(x.size, p, r)
        s = np.ascontiguousarray(x[lo:hi])
        out.append({"lo": int(lo), "hi": int(hi), "sha256": hashlib.sha256(s.tobytes()).hexdigest(), "csum": float(o.csum(s)).hex()})
    return out


def run_case(name):
    import oracle_ffi as o
    from kryst_b200 import stencils
    cfg_name, _, shards = name.partition("@")
    p = int(shards) if shards else 1
    cfg = stencils.CONFIGS[cfg_name]
    max_iters = MAX_ITERS
    note = None
    if cfg_name == "C5":           # two restart cycles on the CPU, not the full solve
        max_iters = 2 * cfg["restart"]
        note = "truncated: max_iters = 2 restart cycles"
    o.use_all_cores()
    t0 = time.perf_counter()
    A = o.stencil(cfg["kind"], cfg["N"])
    b = o.spmv(A, np.ones(A.n))
    x0 = np.zeros(A.n)
    if cfg["solver"] == "pcg":
        rc, x, st, _ = o.pcg(A, o.OPc.jacobi(A), b, x0, TOL, max_iters, nshards=p)
    elif cfg["solver"] == "bicgstab":
        rc, x, st = o.bicgstab(A, o.OPc.jacobi(A), b, x0, TOL, max_iters, variant=o.BICG_TEXTBOOK, nshards=p)
    else:
        pc = o.OPc.ilu0(A, nblocks=p)
        rc, x, st = o.gmres(A, pc, b, x0, cfg["restart"], TOL, max_iters, mode=o.MODE_LEFT, variant=o.GMRES_CGS2, nshards=p)
    dt = time.perf_counter() - t0
    e = {"config": cfg_name, "shards": p, "kind": cfg["kind"], "N": cfg["N"], "n": int(A.n), "nnz": int(A.nnz), "solver": cfg["solver"],
         "pc": cfg["pc"], "restart": cfg.get("restart"), "tol": TOL, "max_iters": max_iters, "status": int(rc),
         "iterations": int(st.iterations), "converged": bool(st.converged), "final_residual": float(st.final_residual).hex(),
         "final_residual_dec": float(st.final_residual), "max_abs_err_vs_ones": float(np.abs(x - 1.0).max()),
         "x_shards": shard_digests(x, p), "oracle_seconds": round(dt, 1), "oracle_threads": o.num_threads()}
    if note:
        e["note"] = note
    return e


def main():
    names = sys.argv[1:] or ["C1", "C3", "C4", "C2"]
    for nm in names:
        e = run_case(nm)
        data = json.load(open(OUT)) if os.path.exists(OUT) else {}     # re-read: several generator processes may run
        data[nm] = e
        json.dump(data, open(OUT + ".tmp", "w"), indent=1, sort_keys=True)
        os.replace(OUT + ".tmp", OUT)
        print(nm, e["iterations"], e["final_residual_dec"], "%.1fs" % e["oracle_seconds"], flush=True)


if __name__ == "__main__":
    main()
