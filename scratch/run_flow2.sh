#!/bin/bash
timeout 120 python scratch/lean_probe.py 16 16 2048 2>&1 | tail -1
show() { python - "$1" "$2" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).readline())
    print(sys.argv[2], round(d['value'],1), d['iterations'], d['parity']['ok'], {k:round(v['avg_ms'],3) for k,v in d['per_class_ms'].items()})
except Exception as e: print(sys.argv[2], 'failed', e)
P
}
timeout 300 python bench_configs.py C4g --reps 2 --no-cpu > gpurun_out/flow2_c4g.jsonl 2> gpurun_out/flow2_c4g.err
show gpurun_out/flow2_c4g.jsonl "C4g"
timeout 300 python bench_configs.py C2 --reps 2 --no-cpu > gpurun_out/flow2_c2.jsonl 2> gpurun_out/flow2_c2.err
show gpurun_out/flow2_c2.jsonl "C2"
