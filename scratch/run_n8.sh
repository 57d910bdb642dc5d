#!/bin/bash
# 8 GPUs: world=8 parity worker, C4 bench (default / PDL / lazy halo), C5
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py -x -q -k "8" > gpurun_out/n8_dist.log 2>&1
echo "dist rc=$?" >> gpurun_out/n8_dist.log
tail -3 gpurun_out/n8_dist.log
run() { # name, env...
  name=$1; shift
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline --configs none 2> gpurun_out/n8_$name.err | grep '^{' > gpurun_out/n8_$name.json
  python -c "import sys,json; d=json.loads(open('gpurun_out/n8_$name.json').read()); print('N=8 $name', round(d['value'],1), round(d['e2e']['value'],1), d['config']['iterations_per_solve'], d['parity']['ok'], d['roofline']['per_class_ms'])"
}
run default KB_DUMMY=1
run pdl KB_PDL=1
run lazy KB_HALO_LAZY=1
run default2 KB_DUMMY=1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 bench_configs.py C5 --reps 2 --no-cpu > gpurun_out/n8_c5.jsonl 2> gpurun_out/n8_c5.err
python - <<'P'
import json
try:
    d=json.loads(open('gpurun_out/n8_c5.jsonl').readline())
    print('C5', round(d['value'],1), d['iterations'], d['converged'], d['parity'].get('ok'), d['roofline']['frac'], {k:round(v['per_iteration_ms'],3) for k,v in d['per_class_ms'].items()})
except Exception as e: print('C5 failed', e)
P
tail -3 gpurun_out/n8_c5.err
