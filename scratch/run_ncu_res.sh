#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
timeout 100 ncu --clock-control none --set full --import-source on -k regex:kb_pcg_resident -c 1 -f -o $O/r02_prof_pcg_resident_c1 python bench_configs.py C1 --no-cpu --reps 1 > $O/ncu_res.log 2>&1
echo "ncu rc=$?"; tail -4 $O/ncu_res.log; ls -la $O/r02_prof_pcg_resident_c1.ncu-rep
