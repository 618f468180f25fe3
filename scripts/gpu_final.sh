#!/bin/bash
# What the driver runs at round end, in the same order: full GPU test suite in ONE process, smoke(), the reference arm, the bench.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m bioreason_b200.build > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
echo "=== pytest -m gpu (single process)"; timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider 2>&1 | tail -6
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "=== bench"; timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_final.log 2>&1; tail -1 gpurun_out/bench_final.log | cut -c1-400
