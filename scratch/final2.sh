#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for c in C4g C2; do timeout 100 python bench_configs.py $c 2>/dev/null | tail -1; done > gpurun_out/configs_v3.jsonl
python -c "
import json
for l in open('gpurun_out/configs_v3.jsonl'):
    d=json.loads(l); print(d['config'], d['iterations'], d['converged'], round(d['it_per_s'],1), round(d['frac_of_peak'],3))"
python bench.py --no-cpu-baseline --steps 3 2>/dev/null | tail -1 | cut -c1-200
