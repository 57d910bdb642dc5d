#!/bin/bash
# call 2: parity of all stage geometries, C3 sweep, ncu of the staged-x SpMV on C3
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_xtile.py tests/test_gpu_variants.py -x -q > $O/xt2_tests.log 2>&1
echo "tests rc=$?" >> $O/xt2_tests.log
tail -4 $O/xt2_tests.log
timeout 300 python scratch/xt_bench.py C3:1:0 C3:1:0:0 C3:1:2 C3:1:2:0 C3:1:3 C3:1:3:0 > $O/xt2_bench.jsonl 2> $O/xt2_bench.err
echo "bench rc=$?"
cat $O/xt2_bench.jsonl
tail -3 $O/xt2_bench.err
NCU="ncu --clock-control none"
KB_SPMV_XTILE=1 KB_XT_CFG=0 timeout 300 $NCU --set full --import-source on -k regex:kb_spmv_xtile -s 4 -c 1 -f -o $O/r02_prof_spmv_xtile_c3 python bench_configs.py C3 --no-cpu --reps 0 > $O/xt2_ncu.log 2>&1
echo "ncu rc=$?"
ls -la $O/*.ncu-rep
