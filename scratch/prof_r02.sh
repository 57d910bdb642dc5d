#!/bin/bash
# round-2 ncu evidence: one launch list of the headline command + full captures of the kernels VERDICT r1 asked for
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out
NCU="ncu --clock-control none"
timeout 400 $NCU --metrics gpu__time_duration.sum -s 40 -c 300 --csv --log-file $O/r02_launches_c4.csv python bench.py --steps 2 --warmup 1 --configs none --no-cpu-baseline > $O/r02_ncu_bench.log 2>&1
timeout 400 $NCU --set full --import-source on -k regex:kb_spmv_bulk -s 6 -c 1 -f -o $O/r02_prof_spmv_c4 python bench.py --steps 1 --warmup 1 --configs none --no-cpu-baseline > $O/r02_ncu_spmv.log 2>&1
timeout 500 $NCU --set full --import-source on -k "regex:kb_trsv_tiles|kb_gs_dot|kb_gs_fused|GsUpdateOp" -s 75 -c 10 -f -o $O/r02_prof_c4g python bench_configs.py C4g --no-cpu --reps 0 > $O/r02_ncu_c4g.log 2>&1
timeout 400 $NCU --set full --import-source on -k regex:kb_trsv_march -s 4 -c 2 -f -o $O/r02_prof_march_c2 python bench_configs.py C2 --no-cpu --reps 0 > $O/r02_ncu_c2.log 2>&1
timeout 400 $NCU --set full --import-source on -k regex:kb_spmv_bulk -s 4 -c 1 -f -o $O/r02_prof_spmv_c3 python bench_configs.py C3 --no-cpu --reps 0 > $O/r02_ncu_c3.log 2>&1
ls -la $O/*.ncu-rep
tail -3 $O/r02_ncu_c4g.log
