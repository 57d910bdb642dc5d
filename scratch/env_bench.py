"""Tuning runs in one process: python scratch/env_bench.py 'C1@KB_PCG_RESIDENT=0' 'C3@KB_SPMV_XTILE=1,KB_XT_CFG=0' ...
(environment knobs are read when the operator / preconditioner is created or at solve time)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import kryst_b200 as kb
import bench
ctx = kb.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", 0))
for arg in sys.argv[1:]:
    name, _, envs = arg.partition("@")
    sets = dict(e.split("=", 1) for e in envs.split(",") if e)
    old = {k: os.environ.get(k) for k in sets}
    os.environ.update(sets)
    try:
        line = bench.measure_config(name, ctx, stream, 1, 0, reps=3, with_cpu=False)
        keep = {k: line[k] for k in ("config", "iterations", "value", "solve_ms")}
        keep["run"] = arg
        keep["parity_ok"] = line["parity"].get("ok")
        keep["frac"] = line["roofline"]["frac"]
        keep["us_per_iteration"] = 1e3 * line["solve_ms"] / max(1, line["iterations"])
        keep["per_class_ms"] = {k: [v["launches"], round(v["per_iteration_ms"], 5)] for k, v in line["per_class_ms"].items()}
        print(json.dumps(keep), flush=True)
    except Exception as e:
        print(json.dumps({"run": arg, "error": repr(e)}), flush=True)
    for k, v in old.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v
