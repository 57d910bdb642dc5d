#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 200 python -m pytest tests/test_gpu_pcg_resident.py "tests/test_gpu_first.py::test_pcg_bit_exact" -x -q > $O/xt5_tests.log 2>&1
echo "tests rc=$?" >> $O/xt5_tests.log
tail -3 $O/xt5_tests.log
timeout 200 python scratch/res_ablate.py 0 7 3 > $O/xt5_ablate.jsonl 2> $O/xt5_ablate.err
cat $O/xt5_ablate.jsonl; tail -3 $O/xt5_ablate.err
timeout 200 python scratch/env_bench.py 'C1@' > $O/xt5_bench.jsonl 2> $O/xt5_bench.err
cat $O/xt5_bench.jsonl; tail -3 $O/xt5_bench.err
