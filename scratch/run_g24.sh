#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_ilu_gmres.py -x -q -k "march or slab" 2>&1 | tail -3
KB_MARCH_GROUP=24 timeout 120 python scratch/lean_probe.py 16 16 2048 2>&1 | tail -1
timeout 120 python scratch/lean_probe.py 16 16 2048 2>&1 | tail -1
show() { python - "$1" "$2" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).readline())
    print(sys.argv[2], round(d['value'],1), d['iterations'], d['parity']['ok'], {k:round(v['avg_ms'],3) for k,v in d['per_class_ms'].items()})
except Exception as e: print(sys.argv[2], 'failed', e)
P
}
for g in 24 2; do
KB_MARCH_GROUP=$g timeout 300 python bench_configs.py C4g --reps 2 --no-cpu > gpurun_out/g24_c4g_$g.jsonl 2> gpurun_out/g24_c4g_$g.err
show gpurun_out/g24_c4g_$g.jsonl "C4g group=$g"
done
KB_MARCH_GROUP=24 KB_MARCH_LAG=8 timeout 300 python bench_configs.py C4g --reps 2 --no-cpu > gpurun_out/g24_c4g_l8.jsonl 2> gpurun_out/g24_c4g_l8.err
show gpurun_out/g24_c4g_l8.jsonl "C4g group=24 lag=8"
