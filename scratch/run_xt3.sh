#!/bin/bash
# call 3: resident PCG parity + C1, final staged-x SpMV parity + ncu evidence on C3
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 400 python -m pytest tests/test_gpu_pcg_resident.py tests/test_gpu_xtile.py "tests/test_gpu_first.py::test_pcg_bit_exact" -x -q > $O/xt3_tests.log 2>&1
echo "tests rc=$?" >> $O/xt3_tests.log
tail -6 $O/xt3_tests.log
timeout 300 python scratch/env_bench.py 'C1@KB_PCG_RESIDENT=0' 'C1@KB_PCG_RESIDENT=1' 'C3@' > $O/xt3_bench.jsonl 2> $O/xt3_bench.err
echo "bench rc=$?"
cat $O/xt3_bench.jsonl
tail -3 $O/xt3_bench.err
NCU="ncu --clock-control none"
timeout 300 $NCU --set full --import-source on -k regex:kb_spmv_xtile -s 4 -c 1 -f -o $O/r02_prof_spmv_xtile_c3 python bench_configs.py C3 --no-cpu --reps 1 > $O/xt3_ncu.log 2>&1
echo "ncu rc=$?"
timeout 300 $NCU --metrics gpu__time_duration.sum -s 10 -c 300 --csv --log-file $O/r02_launches_c3.csv python bench_configs.py C3 --no-cpu --reps 1 > $O/xt3_ncu2.log 2>&1
echo "ncu launches rc=$?"
ls -la $O/*.ncu-rep $O/r02_launches_c3.csv
