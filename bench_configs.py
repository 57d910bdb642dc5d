#!/usr/bin/env python
"""bench_configs.py — auxiliary measurements for the other BASELINE.json configs (C1, C2, C3, C4g, C5-shard).

Not the driver contract (that is bench.py).  For each requested config: one solve to rtol 1e-8 with b = A*1,
x0 = 0 on device-resident vectors, timed with CUDA events on the library stream, plus a profiled solve for the
per-kernel-class times.  Prints one JSON line per config:
  iterations, it/s, algorithmic bytes/iteration (SURVEY §8d byte model), fraction of the measured HBM peak.
Usage: python bench_configs.py C1 C3 C4g [--reps 3]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def bytes_per_iter(cfg, n, nnz):
    b_spmv = 12 * nnz + 4 * (n + 1) + 16 * n
    s = cfg["solver"]
    if s == "pcg":
        return b_spmv + 88 * n
    if s == "bicgstab":
        return 2 * b_spmv + 128 * n + 48 * n
    m = cfg["restart"]
    b_ilu = 12 * nnz + 8 * (n + 1) + 40 * n
    return b_spmv + b_ilu + 12 * (m + 1) * n + 32 * n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="+")
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--max-iters", type=int, default=20000)
    ap.add_argument("--scale", type=int, default=0, help="override N (debug)")
    args = ap.parse_args()
    import numpy as np
    import torch
    import kryst_b200 as kb
    from kryst_b200 import stencils, parallel
    from bench import peaks
    peak, src = peaks()
    rank, world, local = parallel.dist_env()
    torch.cuda.set_device(local)
    ctx = kb.Context(local)
    if world > 1:          # under torchrun: row-block shards, block-Jacobi ILU(0) per GPU
        import torch.distributed as dist
        dist.init_process_group("gloo")
        parallel.init_comm(ctx)
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local))
    for name in args.configs:
        cfg = dict(stencils.CONFIGS[name])
        N = args.scale or cfg["N"]
        t0 = time.perf_counter()
        n_glob, lo, hi, rp, ci, v = parallel.shard_stencil(cfg["kind"], N, world, rank)
        n = hi - lo
        nnz = int(rp[-1])
        A = kb.DeviceCsr.from_csr_shard(n_glob, lo, hi, rp, ci, v, ctx) if world > 1 else kb.DeviceCsr.from_csr(n, n, rp, ci, v, ctx)
        del rp, ci, v
        t_up = time.perf_counter() - t0
        t0 = time.perf_counter()
        pc = (kb.Jacobi() if cfg["pc"] == "jacobi" else kb.Ilu0()).setup(A)
        ctx.synchronize()
        t_pc = time.perf_counter() - t0
        ones = torch.ones(n, dtype=torch.float64, device="cuda")
        b = torch.zeros(n, dtype=torch.float64, device="cuda")
        x = torch.zeros(n, dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        A.matvec(ones, b)
        if cfg["solver"] == "pcg":
            solver = kb.PcgSolver(1e-8, args.max_iters)
            solver.record_history = False
        elif cfg["solver"] == "bicgstab":
            solver = kb.BiCgStabSolver(1e-8, args.max_iters, textbook=True)
        else:
            solver = kb.GmresSolver(cfg["restart"], 1e-8, args.max_iters)
        best = None
        for rep in range(args.reps + 1):
            x.zero_()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            st = solver.solve(A, pc, b, x)
            e1.record(stream)
            e1.synchronize()
            ms = e0.elapsed_time(e1)
            if rep > 0 and (best is None or ms < best):
                best = ms
        err = float((x - 1.0).abs().max().item())
        ctx.profile_reset()
        solver.flags = getattr(solver, "flags", 0) | kb.api.KB_FLAG_PROFILE
        x.zero_()
        torch.cuda.synchronize()
        solver.solve(A, pc, b, x)
        prof = ctx.profile()
        bi = bytes_per_iter(cfg, n, nnz)
        its = st.iterations
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        if rank != 0:
            del A, pc, b, x, ones
            torch.cuda.empty_cache()
            continue
        line = {"config": name, "gpus": world, "kind": cfg["kind"], "N": N, "n": n, "nnz": nnz, "solver": cfg["solver"], "pc": cfg["pc"],
                "iterations": its, "converged": st.converged, "final_residual": st.final_residual, "max_abs_err_vs_ones": err,
                "solve_ms": best, "it_per_s": its / (best * 1e-3), "bytes_per_iter_model": bi,
                "achieved_gbs": bi * its / (best * 1e-3) / 1e9, "frac_of_peak": bi * its / (best * 1e-3) / 1e9 / peak, "peak_gbs": peak,
                "upload_s": t_up, "pc_setup_s": t_pc,
                "per_class": {k: {"launches": v["launches"], "avg_ms": v["ms"] / v["launches"], "total_ms": v["ms"]} for k, v in prof.items()}}
        print(json.dumps(line), flush=True)
        del A, pc, b, x, ones
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
