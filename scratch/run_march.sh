#!/bin/bash
# first GPU check of the lean march kernel: parity tests, then C4g / C2 with several group shapes
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ilu_gmres.py -x -q -k "march or slab or ilu0_factors" > gpurun_out/lean_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/lean_pytest.log
tail -5 gpurun_out/lean_pytest.log
if grep -q "failed\|error\|rc=124" gpurun_out/lean_pytest.log; then exit 1; fi
for g in 2 24 4; do
  KB_TRSV_MARCH=1 KB_MARCH_GROUP=$g timeout 300 python bench_configs.py C4g --reps 2 --no-cpu > gpurun_out/lean_c4g_g$g.jsonl 2> gpurun_out/lean_c4g_g$g.err
  python - <<P
import json
try:
    d=json.loads(open('gpurun_out/lean_c4g_g$g.jsonl').readline())
    print('C4g g=$g', round(d['value'],1), d['iterations'], d['parity']['ok'], {k:round(v['avg_ms'],3) for k,v in d['per_class_ms'].items()})
except Exception as e: print('C4g g=$g failed', e)
P
done
for g in 4 2; do
  KB_MARCH_GROUP=$g timeout 300 python bench_configs.py C2 --reps 2 --no-cpu > gpurun_out/lean_c2_g$g.jsonl 2> gpurun_out/lean_c2_g$g.err
  python - <<P
import json
try:
    d=json.loads(open('gpurun_out/lean_c2_g$g.jsonl').readline())
    print('C2 g=$g', round(d['value'],1), d['iterations'], d['parity']['ok'], {k:round(v['avg_ms'],3) for k,v in d['per_class_ms'].items()})
except Exception as e: print('C2 g=$g failed', e)
P
done
