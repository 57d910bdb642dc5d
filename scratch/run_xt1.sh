#!/bin/bash
# call 1: parity of the staged-x SpMV, then C3 / C4 with and without it
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
date +%s > gpurun_out/xt1_t0
timeout 420 python -m pytest tests/test_gpu_xtile.py -x -q > gpurun_out/xt1_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/xt1_tests.log
tail -5 gpurun_out/xt1_tests.log
timeout 400 python scratch/xt_bench.py > gpurun_out/xt1_bench.jsonl 2> gpurun_out/xt1_bench.err
echo "bench rc=$?"
cat gpurun_out/xt1_bench.jsonl
tail -3 gpurun_out/xt1_bench.err
date +%s > gpurun_out/xt1_t1
