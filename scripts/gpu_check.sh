#!/bin/bash
# Run each GPU test group in its own process (a CUDA fault poisons the context) with a hard timeout.
# Usage (under gpurun): bash scripts/gpu_check.sh [pytest -k expressions...]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu_info.txt 2>&1
python -m bioreason_b200.build > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; exit 1; }
TGROUPS=("$@")
if [ ${#TGROUPS[@]} -eq 0 ]; then
  TGROUPS=("grpo or advantages" "gemm_plain" "gemm_strided or gemm_epilogues" "lmhead")
fi
rc=0
i=0
for g in "${TGROUPS[@]}"; do
  i=$((i+1))
  echo "=== group $i: $g"
  timeout -k 10 ${GROUP_TIMEOUT:-420} python -m pytest tests -m gpu -q -x -rA -k "$g" -p no:cacheprovider > gpurun_out/test_$i.log 2>&1
  r=$?
  tail -n 25 gpurun_out/test_$i.log
  echo "=== group $i exit $r"
  [ $r -ne 0 ] && rc=1
done
exit $rc
