#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 200 python -m pytest tests/test_gpu_xtile.py -x -q > $O/xt7_tests.log 2>&1
echo "tests rc=$?" >> $O/xt7_tests.log
tail -3 $O/xt7_tests.log
timeout 300 python scratch/env_bench.py 'C4@KB_SPMV_XTILE=1,KB_XT_CFG=3' 'C4@' 'C4g@KB_SPMV_XTILE=1,KB_XT_CFG=3' 'C2@KB_SPMV_XTILE=1,KB_XT_CFG=3' 'C1@KB_SPMV_XTILE=1,KB_XT_CFG=3,KB_PCG_RESIDENT=0' > $O/xt7_bench.jsonl 2> $O/xt7_bench.err
cat $O/xt7_bench.jsonl; tail -3 $O/xt7_bench.err
