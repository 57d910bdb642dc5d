#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 200 python -m pytest tests/test_gpu_pcg_resident.py -x -q > $O/xt4_tests.log 2>&1
echo "tests rc=$?" >> $O/xt4_tests.log
tail -3 $O/xt4_tests.log
timeout 200 python scratch/res_ablate.py 0 1 2 4 3 7 8 > $O/xt4_ablate.jsonl 2> $O/xt4_ablate.err
cat $O/xt4_ablate.jsonl; tail -3 $O/xt4_ablate.err
ABL_N=256 timeout 100 python scratch/res_ablate.py 0 7 >> $O/xt4_ablate.jsonl 2>> $O/xt4_ablate.err
tail -2 $O/xt4_ablate.jsonl
