#!/bin/bash
run() { # cfg shape
  KB_TILES_SHAPE=$2 timeout 100 python bench_configs.py $1 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', '$2', d['iterations'], d['converged'], round(d['it_per_s'],1), round(d['frac_of_peak'],3), 'trsv', round(d['per_class']['trsv']['avg_ms'],4))"
}
timeout 200 python -m pytest tests/test_gpu_ilu_gmres.py -m gpu -x -q 2>&1 | tail -2
run C4g 8,8,8
run C4g 16,8,8
run C4g 8,8,16
run C4g 4,8,8
run C2 32,32,1
run C2 64,16,1
