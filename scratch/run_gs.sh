#!/bin/bash
show() { python - "$1" "$2" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).readline())
    print(sys.argv[2], round(d['value'],1), d['iterations'], d['parity']['ok'], {k:round(v['avg_ms'],3) for k,v in d['per_class_ms'].items()})
except Exception as e: print(sys.argv[2], 'failed', e)
P
}
timeout 300 python bench_configs.py C4g --reps 3 --no-cpu > gpurun_out/gs_c4g.jsonl 2> gpurun_out/gs_c4g.err
show gpurun_out/gs_c4g.jsonl "C4g"
KB_GS_FUSE=0 timeout 300 python bench_configs.py C4g --reps 2 --no-cpu > gpurun_out/gs_c4g_f0.jsonl 2> gpurun_out/gs_c4g_f0.err
show gpurun_out/gs_c4g_f0.jsonl "C4g 4-sweep"
