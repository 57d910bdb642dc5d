"""Tuning run for kb_spmv_xtile: C3 (27-point BiCGStab) and C4 (7-point PCG) with the x tiles staged or not."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import kryst_b200 as kb
import bench
ctx = kb.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", 0))
runs = [("C3", "0", "0"), ("C3", "1", "0"), ("C3", "1", "1"), ("C4", "1", "0"), ("C4", "1", "1"), ("C4", "0", "0")]
if len(sys.argv) > 1:
    runs = [tuple(a.split(":")) for a in sys.argv[1:]]
for name, mode, cfg in runs:
    os.environ["KB_SPMV_XTILE"] = mode
    os.environ["KB_XT_CFG"] = cfg
    try:
        line = bench.measure_config(name, ctx, stream, 1, 0, reps=2, with_cpu=False)
        line["xtile"] = {"mode": mode, "cfg": cfg}
        keep = {k: line[k] for k in ("config", "xtile", "iterations", "value", "solve_ms", "parity")}
        keep["frac"] = line["roofline"]["frac"]
        keep["spmv_ms"] = line["per_class_ms"].get("spmv", {}).get("avg_ms")
        keep["per_class_ms"] = {k: round(v["per_iteration_ms"], 5) for k, v in line["per_class_ms"].items()}
        print(json.dumps(keep), flush=True)
    except Exception as e:
        print(json.dumps({"config": name, "xtile": [mode, cfg], "error": repr(e)}), flush=True)
        break
