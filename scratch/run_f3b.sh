#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_first.py -x -q -k "pipelined" 2>&1 | tail -2
for v in pipelined; do
python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline --pcg-variant $v 2> gpurun_out/f3_b1_$v.err | grep '^{' > gpurun_out/f3_b1_$v.json
python -c "import sys,json; d=json.loads(open('gpurun_out/f3_b1_$v.json').read()); print('N=1 $v', round(d['value'],1), round(d['e2e']['value'],1), d['config']['iterations_per_solve'], d['roofline']['per_class_ms'], d['roofline']['iteration'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --pcg-variant $v 2> gpurun_out/f3_b2_$v.err | grep '^{' > gpurun_out/f3_b2_$v.json
python -c "import sys,json; d=json.loads(open('gpurun_out/f3_b2_$v.json').read()); print('N=2 $v', round(d['value'],1), round(d['e2e']['value'],1), d['config']['iterations_per_solve'], d['roofline']['per_class_ms'])"
done
tail -3 gpurun_out/f3_b1_pipelined.err gpurun_out/f3_b2_pipelined.err
