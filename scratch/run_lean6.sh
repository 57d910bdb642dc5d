#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ilu_gmres.py -x -q -k "march or slab or ilu0_factors" > gpurun_out/lean6_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/lean6_pytest.log
tail -4 gpurun_out/lean6_pytest.log
if grep -q "failed\|rc=124" gpurun_out/lean6_pytest.log; then exit 1; fi
KB_LEAN_DBG=0 python scratch/lean_probe.py 16 16 2048 2>&1 | tail -1
show() { python - "$1" "$2" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).readline())
    print(sys.argv[2], round(d['value'],1), d['iterations'], d['parity']['ok'], {k:round(v['avg_ms'],3) for k,v in d['per_class_ms'].items()})
except Exception as e: print(sys.argv[2], 'failed', e)
P
}
for lag in 14 10 18; do
  KB_MARCH_LAG=$lag timeout 300 python bench_configs.py C4g --reps 2 --no-cpu > gpurun_out/lean6_c4g_l$lag.jsonl 2> gpurun_out/lean6_c4g_l$lag.err
  show gpurun_out/lean6_c4g_l$lag.jsonl "C4g lag=$lag"
done
timeout 300 python bench_configs.py C2 --reps 2 --no-cpu > gpurun_out/lean6_c2.jsonl 2> gpurun_out/lean6_c2.err
show gpurun_out/lean6_c2.jsonl "C2"
KB_MARCH_TRACE=1 python scratch/march_probe.py poisson3d 256 2 2>&1 | grep -v "first-step\|entry us\|end   us\|^ \[" | tail -14
KB_MARCH_TRACE=1 python scratch/march_probe.py convdiff2d 1024 4 2>&1 | grep -v "first-step\|entry us\|end   us\|^ \[" | tail -12
