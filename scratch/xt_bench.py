"""Tuning run for kb_spmv_xtile: configs with the x tiles staged or not.  args: NAME:mode:cfg[:prod] ..."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import kryst_b200 as kb
import bench
ctx = kb.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", 0))
runs = [tuple(a.split(":")) for a in sys.argv[1:]]
for r in runs:
    name, mode, cfg = r[:3]
    os.environ["KB_SPMV_XTILE"] = mode
    os.environ["KB_XT_CFG"] = cfg
    if len(r) > 3:
        os.environ["KB_SPMV_PROD"] = r[3]
    else:
        os.environ.pop("KB_SPMV_PROD", None)
    try:
        line = bench.measure_config(name, ctx, stream, 1, 0, reps=2, with_cpu=False)
        keep = {k: line[k] for k in ("config", "iterations", "value", "solve_ms")}
        keep["run"] = ":".join(r)
        keep["parity_ok"] = line["parity"].get("ok")
        keep["frac"] = line["roofline"]["frac"]
        keep["spmv_ms"] = line["per_class_ms"].get("spmv", {}).get("avg_ms")
        keep["per_class_ms"] = {k: round(v["per_iteration_ms"], 5) for k, v in line["per_class_ms"].items()}
        print(json.dumps(keep), flush=True)
    except Exception as e:
        print(json.dumps({"run": ":".join(r), "error": repr(e)}), flush=True)
        break
