#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py -x -q -k "2" > gpurun_out/lazy_dist.log 2>&1
echo "dist rc=$?" >> gpurun_out/lazy_dist.log
tail -3 gpurun_out/lazy_dist.log
for lz in 1 0 1 0; do
KB_HALO_LAZY=$lz python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --configs none 2> gpurun_out/lazy_b2_$lz.err | grep '^{' > gpurun_out/lazy_b2_$lz.json
python -c "import sys,json; d=json.loads(open('gpurun_out/lazy_b2_$lz.json').read()); print('N=2 lazy=$lz', round(d['value'],1), round(d['e2e']['value'],1), d['config']['iterations_per_solve'], d['parity']['ok'], d['roofline']['per_class_ms'])"
done
