#!/usr/bin/env python
"""bench.py — Krylov iterations/s on BASELINE.json's headline workload.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload C4]

Workload (config.workload): C4 = CG + Jacobi on the 3-D 7-point Poisson matrix 256^3 (16.7 M unknowns,
117 M nnz), b = A*1, x0 = 0, rtol 1e-8 — BASELINE.json configs[3], the case the north-star target is
quoted on; it fits one GPU and is row-partitioned over N GPUs (strong scaling).
A *step* is one complete solve(A, pc, b, x).  `value` = iterations/s with b/x resident in HBM;
`e2e` = the same through the host-slice API (pinned host b and x copied H2D, x copied D2H, every step).
`roofline` is for the dominant kernel (fused SpMV + p.Ap): algorithmic bytes 12 nnz + 4 (n+1) + 16 n
per launch over its CUDA-event duration measured in a profiled solve on the library stream.
`cpu_baseline` times the CPU oracle (restatement of the reference's Rayon path) on a bounded sample.
`parity` compares this run's iteration count, final residual and solution bits with the oracle's committed results for
the same config and shard count (tests/golden/config_golden.json, made by tests/golden/make_config_golden.py).
`configs` (N=1: C4g, C1, C2, C3; N=8: C5) repeats the measurement for the other BASELINE.json configs, each with
its own byte model, roofline fraction, per-kernel-class times, parity and a bounded CPU-oracle sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (kind, N, solver, description)
    "C4": ("poisson3d", 256, "pcg", "CG+Jacobi, 3D 7-pt Poisson 256^3 (16.7M unknowns), f64, rtol 1e-8"),
    "C1": ("poisson2d", 512, "pcg", "CG+Jacobi, 2D 5-pt Poisson 512^2 (262k unknowns), f64, rtol 1e-8"),
    "C4s": ("poisson3d", 96, "pcg", "CG+Jacobi, 3D 7-pt Poisson 96^3 (smoke size), f64, rtol 1e-8"),
}
TOL = 1e-8
MAX_ITERS = 20000


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.lines, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        # "under load": the upper half of the samples (the sampler also sees idle gaps between steps)
        load = sm[len(sm) // 2:] if sm else []
        med = load[len(load) // 2] if load else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def bytes_per_iter(cfg, n, nnz):
    """SURVEY §8d algorithmic bytes per (inner) iteration."""
    b_spmv = 12 * nnz + 4 * (n + 1) + 16 * n
    if cfg["solver"] == "pcg":
        return b_spmv + 88 * n
    if cfg["solver"] == "bicgstab":
        return 2 * b_spmv + 128 * n + 48 * n
    m = cfg["restart"]
    return b_spmv + (12 * nnz + 8 * (n + 1) + 40 * n) + 12 * (m + 1) * n + 32 * n


def golden_entry(name, world):
    p = os.path.join(ROOT, "tests", "golden", "config_golden.json")
    try:
        g = json.load(open(p))
    except Exception:
        return None
    return g.get(name if world == 1 else "%s@%d" % (name, world))


def parity_block(name, world, rank, stats, x_dev):
    """Compare with the oracle's committed result for the same config and shard count: iterations, final residual
    (bit-equal) and the SHA-256 of this rank's slice of the solution.  ok = all ranks agree with the oracle."""
    import hashlib
    import torch
    g = golden_entry(name, world)
    if g is None:
        return {"ok": None, "note": "no golden entry for %s at %d shard(s)" % (name, world)}
    x = x_dev.detach().cpu().numpy()
    sh = g["x_shards"][rank]
    x_ok = hashlib.sha256(x.tobytes()).hexdigest() == sh["sha256"]
    mine = (stats.iterations == g["iterations"] and float(stats.final_residual).hex() == g["final_residual"]
            and bool(stats.converged) == g["converged"])
    ok = bool(mine and x_ok)
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([1.0 if ok else 0.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        ok = bool(t.item() == 1.0)
    return {"ok": ok, "iterations": int(stats.iterations), "oracle_iterations": g["iterations"],
            "final_residual": float(stats.final_residual), "oracle_final_residual": g["final_residual_dec"],
            "x_bit_equal": bool(x_ok) if world == 1 else ok, "max_abs_err_vs_ones": float((x_dev - 1.0).abs().max().item()),
            "oracle": "tests/golden/config_golden.json (oracle/kryst_oracle.cpp, nshards=%d%s)" % (world, ", " + g["note"] if g.get("note") else "")}


def cpu_sample(name, world=1, budget_s=6.0):
    """Bounded CPU-oracle sample of config `name` (the first iterations of the same solve), all host cores."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import oracle_ffi as o
    from kryst_b200 import stencils
    cfg = stencils.CONFIGS[name]
    cores = o.use_all_cores()
    A = o.stencil(cfg["kind"], cfg["N"])
    b = o.spmv(A, np.ones(A.n))
    x0 = np.zeros(A.n)
    if cfg["solver"] == "pcg":
        pc = o.OPc.jacobi(A)
        run = lambda k: o.pcg(A, pc, b, x0, TOL, k)[2]
        unit, probe = 1, 2
    elif cfg["solver"] == "bicgstab":
        pc = o.OPc.jacobi(A)
        run = lambda k: o.bicgstab(A, pc, b, x0, TOL, k, variant=o.BICG_TEXTBOOK)[2]
        unit, probe = 1, 2
    else:
        pc = o.OPc.ilu0(A, nblocks=world)          # block-Jacobi ILU(0), one block per shard (asm.rs:46-57)
        m = cfg["restart"]
        run = lambda k: o.gmres(A, pc, b, x0, m, TOL, k, mode=o.MODE_LEFT, variant=o.GMRES_CGS2, nshards=world)[2]
        unit, probe = m, m            # whole restart cycles: the cost of an inner step grows with j
    t0 = time.perf_counter()
    st = run(probe)
    t_probe = time.perf_counter() - t0
    per = t_probe / max(1, int(st.iterations))
    k = int(max(1, min(MAX_ITERS // unit, budget_s / max(per * unit, 1e-9)))) * unit
    if k > probe:
        t0 = time.perf_counter()
        st = run(k)
        t_probe = time.perf_counter() - t0
    its = int(st.iterations)
    what = {"pcg": "PCG+Jacobi", "bicgstab": "BiCGStab+Jacobi (textbook)", "gmres": "GMRES(%s)+ILU(0), whole restart cycles" % cfg.get("restart")}[cfg["solver"]]
    return {"value": its / t_probe, "unit": "it/s", "cores": cores, "kind": "port",
            "sample": "first %d iterations of the same solve, %s (oracle/kryst_oracle.cpp, OpenMP; setup excluded)" % (its, what)}


def measure_config(name, ctx, stream, world, rank, reps=2, A=None, with_cpu=True):
    """One BASELINE.json config on the GPU(s): solve to rtol 1e-8 from x0 = 0 on device-resident vectors, CUDA events on
    the library stream, best of `reps` after one warm-up; then a profiled solve for the per-class times."""
    import torch
    import kryst_b200 as kb
    from kryst_b200 import stencils, parallel
    cfg = stencils.CONFIGS[name]
    peak, peak_src = peaks()
    t0 = time.perf_counter()
    own = A is None
    n_glob = stencils.dim(cfg["kind"], cfg["N"])
    lo, hi = kb.partition_range(n_glob, world, rank)
    if own:
        _, rp, ci, v = stencils.stencil(cfg["kind"], cfg["N"], lo, hi)
        A = kb.DeviceCsr.from_csr_shard(n_glob, lo, hi, rp, ci, v, ctx) if world > 1 else kb.DeviceCsr.from_csr(n_glob, n_glob, rp, ci, v, ctx)
        del rp, ci, v
    n, nnz = hi - lo, A.nnz()
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
    del ones
    if cfg["solver"] == "pcg":
        solver = kb.PcgSolver(TOL, MAX_ITERS)
        solver.record_history = False
    elif cfg["solver"] == "bicgstab":
        solver = kb.BiCgStabSolver(TOL, MAX_ITERS, textbook=True)
    else:
        solver = kb.GmresSolver(cfg["restart"], TOL, MAX_ITERS)
        if os.environ.get("KB_BENCH_BLOCK_ORTH") == "1":     # tuning runs only: the opt-in block-orthogonalisation variant (parity is vs CGS2 goldens)
            solver.with_block_orthogonalisation()
    best, st = None, None
    for rep in range(reps + 1):
        x.zero_()
        torch.cuda.synchronize()
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        st = solver.solve(A, pc, b, x)
        e1.record(stream)
        e1.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        if rep > 0 and (best is None or ms < best):
            best = ms
    par = parity_block(name, world, rank, st, x)
    ctx.profile_reset()
    solver.flags = getattr(solver, "flags", 0) | kb.api.KB_FLAG_PROFILE
    x.zero_()
    torch.cuda.synchronize()
    solver.solve(A, pc, b, x)
    prof = ctx.profile()
    bi = bytes_per_iter(cfg, n, nnz)
    its = st.iterations
    line = {"config": name, "gpus": world, "workload": "%s %s N=%d, %s + %s" % (name, cfg["kind"], cfg["N"], cfg["solver"] + ("(%d)" % cfg["restart"] if "restart" in cfg else ""), cfg["pc"]),
            "n": n_glob, "nnz_rank0": nnz, "iterations": its, "converged": bool(st.converged), "final_residual": st.final_residual,
            "solve_ms": best, "value": its / (best * 1e-3), "unit": "it/s",
            "roofline": {"bound": "hbm", "scope": "whole (inner) iteration, SURVEY 8d byte model", "bytes_per_iteration_per_gpu": bi,
                         "achieved": bi * its / (best * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": bi * its / (best * 1e-3) / 1e9 / peak,
                         "peak_source": peak_src},
            "parity": par, "upload_s": round(t_up, 3), "pc_setup_s": round(t_pc, 3),
            "per_class_ms": {k: {"launches": v["launches"], "avg_ms": v["ms"] / v["launches"], "per_iteration_ms": v["ms"] / max(1, its)} for k, v in prof.items()}}
    pc.close()
    if own:
        A.close()
    del b, x
    torch.cuda.empty_cache()
    if with_cpu and rank == 0:
        try:
            line["cpu_baseline"] = cpu_sample(name, world)
        except Exception as e:       # a reported baseline must never take the GPU numbers down with it
            line["cpu_baseline"] = {"error": repr(e)}
    return line


def run_reference(args):
    """Reference arm: the reference's own CPU algorithm for this path.  kryst is pure Rust and no Rust
    toolchain exists in the image, so this is the oracle port (oracle/kryst_oracle.cpp, OpenMP on all host
    cores) — the one other place bench.py may execute oracle/."""
    rank, world, _ = dist_env()
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import oracle_ffi as o
    kind, N, solver, desc = WORKLOADS[args.workload]
    A = o.stencil(kind, N)
    b = o.spmv(A, np.ones(A.n))
    pc = o.OPc.jacobi(A)
    cores = o.use_all_cores()        # torchrun sets OMP_NUM_THREADS=1 for its workers
    # bounded sample: S iterations of the same solve (from x0 = 0) per step
    t0 = time.perf_counter()
    o.pcg(A, pc, b, np.zeros(A.n), TOL, 2)
    t2 = (time.perf_counter() - t0) / 3.0        # ~ per-iteration cost incl. the initial residual SpMV
    budget = 120.0 / max(1, args.steps + args.warmup)
    S = int(max(3, min(MAX_ITERS, budget / max(t2, 1e-6))))
    for _ in range(args.warmup):
        o.pcg(A, pc, b, np.zeros(A.n), TOL, S)
    its, t0 = 0, time.perf_counter()
    for _ in range(args.steps):
        rc, x, st, _h = o.pcg(A, pc, b, np.zeros(A.n), TOL, S)
        its += int(st.iterations)
    S = its // max(1, args.steps)
    dt = time.perf_counter() - t0
    val = its / dt
    line = {
        "impl": "reference", "metric": "krylov_iterations_per_sec", "value": val, "unit": "it/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload + ": " + desc, "n": A.n, "nnz": A.nnz, "rtol": TOL,
                   "note": "reference = CPU oracle port of kryst's Rayon path (Rust toolchain absent); every step is a fresh solve() call, so "
                           "it pays the reference's per-call allocation and first touch of five n-vectors (pcg.rs:117-149) for %d iterations" % S},
        "cpu_baseline": {"value": val, "unit": "it/s", "cores": cores, "kind": "port",
                         "sample": "%d PCG+Jacobi iterations from x0=0 per step (of the full solve)" % S},
        "e2e": {"value": val, "unit": "it/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def cpu_baseline(args, kind, N):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import oracle_ffi as o
    o.use_all_cores()
    A = o.stencil(kind, N)
    b = o.spmv(A, np.ones(A.n))
    pc = o.OPc.jacobi(A)
    t0 = time.perf_counter()
    o.pcg(A, pc, b, np.zeros(A.n), TOL, 2)
    t2 = (time.perf_counter() - t0) / 3.0
    S = int(max(3, min(MAX_ITERS, 12.0 / max(t2, 1e-6))))
    t0 = time.perf_counter()
    rc, x, st, _h = o.pcg(A, pc, b, np.zeros(A.n), TOL, S)
    dt = time.perf_counter() - t0
    return {"value": int(st.iterations) / dt, "unit": "it/s", "cores": o.num_threads(), "kind": "port",
            "sample": "first %d PCG+Jacobi iterations of the same solve (oracle/kryst_oracle.cpp, OpenMP)" % int(st.iterations)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="kryst_b200")
    ap.add_argument("--workload", default="C4", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--configs", default="auto", help="comma list of other BASELINE configs to append (auto: N=1 -> C4g,C1,C2,C3; N=8 -> C5; none)")
    ap.add_argument("--pcg-variant", default="literal", choices=["literal", "fused", "pipelined"],
                    help="literal = pcg.rs recurrences (the headline); fused = single-reduction extension, pipelined = Ghysels-Vanroose (SURVEY 8(f3))")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import kryst_b200 as kb
    from kryst_b200 import stencils

    rank, world, local = dist_env()
    if world != args.gpus:
        if args.gpus != 1 or world != 1:
            raise SystemExit("launch with torchrun --nproc-per-node %d (WORLD_SIZE=%d)" % (args.gpus, world))
    kind, N, solver, desc = WORKLOADS[args.workload]
    torch.cuda.set_device(local)
    ctx = kb.Context(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        uid = [kb.Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(rank, world, uid[0])

    n = stencils.dim(kind, N)
    lo, hi = kb.partition_range(n, world, rank)
    _, rp, ci, v = stencils.stencil(kind, N, lo, hi)
    nnz_local = int(rp[-1])
    if world > 1:
        A = kb.DeviceCsr.from_csr_shard(n, lo, hi, rp, ci, v, ctx)
    else:
        A = kb.DeviceCsr.from_csr(n, n, rp, ci, v, ctx)
    del rp, ci, v
    nloc = hi - lo
    pc = kb.Jacobi().setup(A)

    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local))
    # b = A * 1 (computed once, on the device); x0 = 0
    ones = torch.ones(nloc, dtype=torch.float64, device="cuda")
    b_dev = torch.zeros(nloc, dtype=torch.float64, device="cuda")
    x_dev = torch.zeros(nloc, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    A.matvec(ones, b_dev)
    del ones
    b_host = torch.empty(nloc, dtype=torch.float64).pin_memory()
    x_host = torch.empty(nloc, dtype=torch.float64).pin_memory()
    b_host.copy_(b_dev)
    torch.cuda.synchronize()
    b_np, x_np = b_host.numpy(), x_host.numpy()

    solver_obj = kb.PcgSolver(TOL, MAX_ITERS)
    solver_obj.record_history = False
    fused = args.pcg_variant != "literal"
    if args.pcg_variant == "fused":
        solver_obj.with_fused_reduction(True)
    elif args.pcg_variant == "pipelined":
        solver_obj.with_pipelined(True)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """barrier+sync, CUDA events on the library stream around `steps` calls, max over ranks."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        its = 0
        for _ in range(steps):
            its += fn()
        e1.record(stream)
        e1.synchronize()
        wall = time.perf_counter() - t0
        barrier()
        ms = max(e0.elapsed_time(e1), 0.0)
        sec = max(ms * 1e-3, wall)          # the calls are synchronous; never report less than the wall time
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([sec], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sec = float(t.item())
        return sec, its

    def step_device():
        x_dev.zero_()
        torch.cuda.current_stream().synchronize()
        return solver_obj.solve(A, pc, b_dev, x_dev).iterations

    def step_host():
        x_host.zero_()          # x0 = 0 in the pinned host buffer (x_np is its numpy view; numpy's own fill is 5x slower: one thread)
        return solver_obj.solve(A, pc, b_np, x_np).iterations

    # warm-up (>= 3), both paths
    W = max(args.warmup, 3)
    for _ in range(W):
        step_device()
    step_host()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = ctx.launch_count()
    sec, its = timed(step_device, args.steps)
    launches = ctx.launch_count() - l0
    sec_e2e, its_e2e = timed(step_host, args.steps)
    clocks = sampler.stop() if rank == 0 else None

    # dominant-kernel roofline: one profiled solve (every launch bracketed by CUDA events on the library stream)
    ctx.profile_reset()
    solver_obj.flags = kb.api.KB_FLAG_PROFILE
    x_dev.zero_()
    torch.cuda.synchronize()
    prof_its = solver_obj.solve(A, pc, b_dev, x_dev).iterations
    solver_obj.flags = 0
    prof = ctx.profile()
    peak, peak_src = peaks()
    spmv = prof.get("spmv", {"launches": 1, "ms": float("nan")})
    spmv_ms = spmv["ms"] / max(spmv["launches"], 1)
    b_spmv = stencils.spmv_bytes(nloc, nnz_local)
    achieved = b_spmv / (spmv_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "spmv_traffic.json")
    traffic_src = None
    if world == 1 and os.path.exists(tp):     # the ncu capture is of the N=1 launch; a shard's launch was never captured -> null
        try:
            traffic = json.load(open(tp)).get(args.workload, {}).get("dram_bytes_per_launch")
            traffic_src = "ncu --set full capture of this kernel at N=1 (profiles/spmv_traffic.json), not re-measured in this run"
        except Exception:
            traffic = None
    total_ms = sum(v["ms"] for v in prof.values())
    iter_bytes = b_spmv + {"literal": 88, "fused": 96, "pipelined": 168}[args.pcg_variant] * nloc   # PCG+Jacobi per iteration (fused variant: DESIGN §4)
    roofline = {"bound": "hbm", "kernel": "kb_spmv_bulk<PcgAp> (bulk-async staged CSR SpMV fused with p.Ap)", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": b_spmv, "avg_launch_ms": spmv_ms,
                "share_of_step": spmv["ms"] / total_ms if total_ms else None,
                "iteration": {"bytes": iter_bytes, "achieved": iter_bytes * world * its / sec / args.gpus / 1e9,
                              "frac": iter_bytes * its / sec / 1e9 / peak},
                "per_class_ms": {k: v["ms"] / v["launches"] for k, v in prof.items()}}

    # parity of the headline solve against the oracle's committed result for this shard count (outside the timed region)
    x_dev.zero_()
    torch.cuda.synchronize()
    st_par = solver_obj.solve(A, pc, b_dev, x_dev)
    parity = parity_block(args.workload, world, rank, st_par, x_dev) if not fused else {"ok": None, "note": "fused variant: see tests"}

    # the other BASELINE.json configs (never allowed to take the headline down)
    wanted = {"auto": {1: ["C4g", "C1", "C2", "C3"], 8: ["C5"]}.get(world, []), "none": []}.get(args.configs, [c for c in args.configs.split(",") if c])
    extra = []
    for name in wanted:
        try:
            share = A if (name == "C4g" and args.workload == "C4") else None      # same matrix: no second upload
            extra.append(measure_config(name, ctx, stream, world, rank, A=share, with_cpu=not args.no_cpu_baseline))
        except Exception as e:
            extra.append({"config": name, "error": repr(e)})
        if world > 1:
            import torch.distributed as dist
            dist.barrier()

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and world == 1:     # reported at N=1 only (the other ranks would idle meanwhile)
            cpu = cpu_baseline(args, kind, N)
        line = {
            "metric": "krylov_iterations_per_sec", "value": its / sec, "unit": "it/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": W, "ms_per_step": 1e3 * sec / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload + ": " + desc, "n": n, "nnz_rank0": nnz_local, "rtol": TOL,
                       "iterations_per_solve": its // args.steps, "pcg_variant": args.pcg_variant, "parallelism": "row-block x%d" % world,
                       "l2": "inputs larger than L2 (per-iteration working set %.2f GB)" % (iter_bytes / 1e9)},
            "clocks": clocks,
            "e2e": {"value": its_e2e / sec_e2e, "unit": "it/s", "h2d_bytes_per_step": 16 * nloc, "d2h_bytes_per_step": 8 * nloc,
                    "ms_per_step": 1e3 * sec_e2e / args.steps},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "parity": parity,
            "configs": extra,
        }
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        A.close()
        ctx.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
