#!/bin/bash
# full GPU suite, the contract bench line, and a launch list of the C4g solve kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/full_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/full_pytest.log
tail -6 gpurun_out/full_pytest.log
timeout 900 python bench.py > gpurun_out/full_bench.json 2> gpurun_out/full_bench.err
echo "bench rc=$?"
python - <<'P'
import json
try:
    d=json.loads(open('gpurun_out/full_bench.json').read().strip().splitlines()[-1])
    print(d['value'], d['e2e']['value'], d['roofline']['frac'], d.get('parity',{}).get('ok'))
    for c in d.get('configs',[]):
        print(c.get('config'), round(c.get('value',0),1), c.get('roofline',{}).get('frac'), c.get('parity',{}).get('ok'), c.get('cpu_baseline',{}).get('value'))
except Exception as e: print('bench parse failed', e)
P
tail -3 gpurun_out/full_bench.err
timeout 500 ncu --clock-control none --metrics gpu__time_duration.sum -k "regex:kb_spmv|kb_gs|kb_trsv|kb_tile_kernel|kb_arnoldi|kb_gmres|kb_block" -s 100 -c 400 --csv --log-file gpurun_out/r02b_launches_c4g.csv python bench_configs.py C4g --no-cpu --reps 0 > gpurun_out/r02b_ncu_c4g_list.log 2>&1
tail -2 gpurun_out/r02b_ncu_c4g_list.log
