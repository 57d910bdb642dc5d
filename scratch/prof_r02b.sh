#!/bin/bash
# round-2 (second half) ncu evidence: the lean pencil-march solves on C4g / C2, C4g launch list
set -x
O=gpurun_out
NCU="ncu --clock-control none"
timeout 500 $NCU --metrics gpu__time_duration.sum -s 200 -c 400 --csv --log-file $O/r02b_launches_c4g.csv python bench_configs.py C4g --no-cpu --reps 0 > $O/r02b_ncu_c4g_list.log 2>&1
timeout 600 $NCU --set full --import-source on -k "regex:kb_trsv_lean" -s 20 -c 2 -f -o $O/r02b_prof_lean_c4g python bench_configs.py C4g --no-cpu --reps 0 > $O/r02b_ncu_c4g.log 2>&1
timeout 400 $NCU --set full --import-source on -k "regex:kb_trsv_lean" -s 8 -c 2 -f -o $O/r02b_prof_lean_c2 python bench_configs.py C2 --no-cpu --reps 0 > $O/r02b_ncu_c2.log 2>&1
ls -la $O/*.ncu-rep | tail -4
tail -3 $O/r02b_ncu_c4g.log
