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
for lag in 5 6 12; do
  KB_TRSV_MARCH=1 KB_MARCH_LAG=$lag timeout 300 python bench_configs.py C4g --reps 2 --no-cpu > gpurun_out/lag_c4g_$lag.jsonl 2> gpurun_out/lag_c4g_$lag.err
  show gpurun_out/lag_c4g_$lag.jsonl "C4g lag=$lag"
  KB_MARCH_LAG=$lag timeout 300 python bench_configs.py C2 --reps 2 --no-cpu > gpurun_out/lag_c2_$lag.jsonl 2> gpurun_out/lag_c2_$lag.err
  show gpurun_out/lag_c2_$lag.jsonl "C2 lag=$lag"
done
