#!/bin/bash
mkdir -p gpurun_out
show() { python - "$1" "$2" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).readline())
    print(sys.argv[2], round(d['value'],1), d['iterations'], d['parity']['ok'], {k:round(v['avg_ms'],3) for k,v in d['per_class_ms'].items()})
except Exception as e: print(sys.argv[2], 'failed', e)
P
}
for lag in 5 6 7 8 10; do
  KB_MARCH_LAG=$lag timeout 300 python bench_configs.py C4g --reps 2 --no-cpu > gpurun_out/lean7_c4g_l$lag.jsonl 2> gpurun_out/lean7_c4g_l$lag.err
  show gpurun_out/lean7_c4g_l$lag.jsonl "C4g lag=$lag"
done
for lag in 6 8; do
KB_MARCH_LAG=$lag timeout 300 python bench_configs.py C2 --reps 2 --no-cpu > gpurun_out/lean7_c2_$lag.jsonl 2> gpurun_out/lean7_c2.err
show gpurun_out/lean7_c2_$lag.jsonl "C2 lag=$lag"
done
KB_MARCH_LAG=6 KB_MARCH_TRACE=1 python scratch/march_probe.py poisson3d 256 2 2>&1 | grep -v "first-step\|entry us\|end   us\|^ \[" | tail -10
