#!/bin/bash
# usage: bench2.sh N
N=$1
for v in literal fused; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline --pcg-variant $v 2>&1 | grep '^{' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', d['n_gpus'], d['value'], d['e2e']['value'], d['config']['iterations_per_solve'], d['roofline']['per_class_ms'], d['gpu_launches'])"
done
