#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m bioreason_b200.build > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
echo "=== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -15
echo "=== trainer test"; timeout 600 python -m pytest tests/test_gpu_trainer.py -m gpu -q -x -rA 2>&1 | tail -25
echo "=== bench small"; timeout 900 python bench.py --text small --dna small --dna-len 100 --text-len 120 --completion 24 --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | tail -12
echo "=== bench full (c)"; timeout 1500 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_c_first.log 2>&1; tail -15 gpurun_out/bench_c_first.log
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
