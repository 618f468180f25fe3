#!/bin/bash
# Round-2 ncu evidence: (1) one full capture of every hot kernel at config (c) shapes, (2) the launch list of a short bench step.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/r02_targets python scripts/ncu_targets.py > gpurun_out/r02_ncu_targets.log 2>&1
ls -la gpurun_out/r02_targets.ncu-rep | awk '{print $5, $9}'
C=${PROFILE_C:-16}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -c ${PROFILE_MAX:-20000} --csv \
   --log-file gpurun_out/r02_launches_c${C}.csv python bench.py --steps 1 --warmup 1 --completion $C --no-cpu-baseline > gpurun_out/r02_bench_under_ncu.log 2>&1
python scripts/summarize_launches.py gpurun_out/r02_launches_c${C}.csv > gpurun_out/r02_launches_c${C}_summary.txt 2>&1; head -40 gpurun_out/r02_launches_c${C}_summary.txt
gzip -f gpurun_out/r02_launches_c${C}.csv
