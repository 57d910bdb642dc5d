#!/bin/bash
# 2 GPUs: single-GPU tests of the new variants, 2-GPU parity worker, variants at N=1 (rank 0 GPU) and N=2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_first.py -x -q -k "pcg" > gpurun_out/f3_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/f3_pytest.log
tail -4 gpurun_out/f3_pytest.log
timeout 900 python -m pytest tests/test_gpu_dist.py -x -q -k "2" > gpurun_out/f3_dist.log 2>&1
echo "dist rc=$?" >> gpurun_out/f3_dist.log
tail -15 gpurun_out/f3_dist.log
for v in literal fused pipelined; do
python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline --pcg-variant $v 2> gpurun_out/f3_b1_$v.err | grep '^{' > gpurun_out/f3_b1_$v.json
python -c "import sys,json; d=json.loads(open('gpurun_out/f3_b1_$v.json').read()); print('N=1 $v', round(d['value'],1), round(d['e2e']['value'],1), d['config']['iterations_per_solve'], d['roofline']['per_class_ms'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --pcg-variant $v 2> gpurun_out/f3_b2_$v.err | grep '^{' > gpurun_out/f3_b2_$v.json
python -c "import sys,json; d=json.loads(open('gpurun_out/f3_b2_$v.json').read()); print('N=2 $v', round(d['value'],1), round(d['e2e']['value'],1), d['config']['iterations_per_solve'], d['roofline']['per_class_ms'])"
done
