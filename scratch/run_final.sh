#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 400 python -m pytest tests -m gpu -q --durations=12 > $O/final_tests.log 2>&1
echo "tests rc=$?" >> $O/final_tests.log
tail -22 $O/final_tests.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > $O/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/final_smoke.log
timeout 240 python bench.py --configs C1,C3 --no-cpu-baseline > $O/final_bench.json 2> $O/final_bench.err; echo "bench rc=$?"
cut -c1-1500 $O/final_bench.json; tail -2 $O/final_bench.err
