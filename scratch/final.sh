#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -2
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py 2>gpurun_out/bench.err | tail -1 > gpurun_out/bench_final.json; tail -c 600 gpurun_out/bench_final.json; echo
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_ref.json; tail -c 400 gpurun_out/bench_ref.json; echo
for c in C1 C3 C4g C2; do python bench_configs.py $c 2>/dev/null | tail -1; done > gpurun_out/configs_v2.jsonl
wc -l gpurun_out/configs_v2.jsonl
