#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ilu_gmres.py -x -q -k "march or slab or ilu0_factors" > gpurun_out/flow_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/flow_pytest.log
tail -6 gpurun_out/flow_pytest.log
if grep -q "failed\|rc=124" gpurun_out/flow_pytest.log; then exit 1; fi
timeout 120 python scratch/lean_probe.py 16 16 2048 2>&1 | tail -1
show() { python - "$1" "$2" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).readline())
    print(sys.argv[2], round(d['value'],1), d['iterations'], d['parity']['ok'], {k:round(v['avg_ms'],3) for k,v in d['per_class_ms'].items()})
except Exception as e: print(sys.argv[2], 'failed', e)
P
}
for lag in 6 10; do
  KB_MARCH_LAG=$lag timeout 300 python bench_configs.py C4g --reps 2 --no-cpu > gpurun_out/flow_c4g_l$lag.jsonl 2> gpurun_out/flow_c4g_l$lag.err
  show gpurun_out/flow_c4g_l$lag.jsonl "C4g lag=$lag"
  KB_MARCH_LAG=$lag timeout 300 python bench_configs.py C2 --reps 2 --no-cpu > gpurun_out/flow_c2_$lag.jsonl 2> gpurun_out/flow_c2.err
  show gpurun_out/flow_c2_$lag.jsonl "C2 lag=$lag"
done
KB_MARCH_TRACE=1 timeout 120 python scratch/march_probe.py poisson3d 256 2 2>&1 | grep -v "first-step\|entry us\|end   us\|^ \[" | tail -10
