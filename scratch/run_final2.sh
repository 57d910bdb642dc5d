#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 300 python -m pytest tests -m gpu -q > $O/final2_tests.log 2>&1
echo "tests rc=$?" >> $O/final2_tests.log
tail -4 $O/final2_tests.log
timeout 120 python bench.py --configs none --no-cpu-baseline > $O/final2_bench.json 2> $O/final2_bench.err; echo "bench rc=$?"
cut -c1-1100 $O/final2_bench.json; tail -2 $O/final2_bench.err
