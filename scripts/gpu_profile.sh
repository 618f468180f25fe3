#!/bin/bash
# Bounded ncu passes over the bench command (reduced completion length so the launch count stays tractable under ncu).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m bioreason_b200.build > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
C=${PROFILE_C:-8}
echo "=== launch list (completion=$C)"
timeout ${PROFILE_TIMEOUT:-700} ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -c ${PROFILE_MAX:-12000} --csv \
   --log-file gpurun_out/launches_c${C}.csv python bench.py --steps 1 --warmup 1 --completion $C --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_c${C}.csv > gpurun_out/launches_c${C}_summary.txt 2>&1; head -32 gpurun_out/launches_c${C}_summary.txt
for K in ${PROFILE_KERNELS:-skinny_tc5_kernel gemm_tc5_kernel attn_fwd_kernel attn_bwd_kernel decode_fused_kernel}; do
  echo "=== full capture $K"
  timeout 400 ncu --set full --clock-control none --import-source on --graph-profiling node -k regex:$K -s ${PROFILE_SKIP:-40} -c 2 -f -o gpurun_out/prof_$K \
     python bench.py --steps 1 --warmup 1 --completion 4 --no-cpu-baseline > gpurun_out/prof_$K.log 2>&1
  ls -la gpurun_out/prof_$K.ncu-rep 2>/dev/null | awk '{print $5, $9}'
done
