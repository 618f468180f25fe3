#!/bin/bash
# Round-2 closing run on the GPU box: full GPU test suite, smoke, ncu evidence, final bench line.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/r02_targets python scripts/ncu_targets.py > gpurun_out/r02_ncu_targets.log 2>&1
ncu -i gpurun_out/r02_targets.ncu-rep --page raw --csv 2>/dev/null > gpurun_out/r02_targets_raw.csv
python scripts/ncu_summary.py < gpurun_out/r02_targets_raw.csv > gpurun_out/r02_ncu_targets.txt 2>&1; cat gpurun_out/r02_ncu_targets.txt | cut -c1-150
C=16
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -c 20000 --csv \
   --log-file gpurun_out/r02_launches_c${C}_final.csv python bench.py --steps 1 --warmup 1 --completion $C --no-cpu-baseline --no-sweep > gpurun_out/r02_bench_under_ncu.log 2>&1
python scripts/summarize_launches.py gpurun_out/r02_launches_c${C}_final.csv > gpurun_out/r02_launches_c${C}_final_summary.txt 2>&1; head -24 gpurun_out/r02_launches_c${C}_final_summary.txt
gzip -f gpurun_out/r02_launches_c${C}_final.csv
timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tail -1 > gpurun_out/r02_bench_n1_final.json
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n1_final.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['phases'].get('gpu_phase_s'), d.get('secondary_resident_rows'), d['cpu_baseline']['value'])"
