#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ilu_gmres.py -x -q -k "gmres" > gpurun_out/block_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/block_pytest.log
tail -4 gpurun_out/block_pytest.log
show() { python - "$1" "$2" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).readline())
    print(sys.argv[2], round(d['value'],1), d['iterations'], d['converged'], d['final_residual'], d['parity'].get('ok'), {k:round(v['per_iteration_ms'],3) for k,v in d['per_class_ms'].items()})
except Exception as e: print(sys.argv[2], 'failed', e)
P
}
for c in C4g C2; do
KB_BENCH_BLOCK_ORTH=1 timeout 300 python bench_configs.py $c --reps 2 --no-cpu > gpurun_out/block_$c.jsonl 2> gpurun_out/block_$c.err
show gpurun_out/block_$c.jsonl "$c block"
done
