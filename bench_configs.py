#!/usr/bin/env python
"""bench_configs.py — one JSON line per requested BASELINE.json config (C1, C2, C3, C4, C4g; C5 under torchrun x8).

Not the driver contract (that is bench.py, whose `configs` array carries the same measurements); a CLI for
tuning runs:   python bench_configs.py C4g C2 [--reps 3] [--no-cpu]
Each line: iterations, it/s, SURVEY 8d byte model, fraction of the measured HBM peak, per-kernel-class times,
parity against tests/golden/config_golden.json and (rank 0) a bounded CPU-oracle sample.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="+")
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    import torch
    import kryst_b200 as kb
    from kryst_b200 import parallel
    import bench
    rank, world, local = parallel.dist_env()
    torch.cuda.set_device(local)
    ctx = kb.Context(local)
    if world > 1:          # under torchrun: row-block shards, block-Jacobi ILU(0) per GPU
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        parallel.init_comm(ctx)
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local))
    for name in args.configs:
        line = bench.measure_config(name, ctx, stream, world, rank, reps=args.reps, with_cpu=not args.no_cpu)
        if rank == 0:
            print(json.dumps(line), flush=True)
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
    if world > 1:
        import torch.distributed as dist
        ctx.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
