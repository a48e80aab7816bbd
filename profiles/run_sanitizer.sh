#!/bin/bash
# compute-sanitizer evidence: memcheck + racecheck over smoke() and the EMD / pairwise / contraction
# kernels.  Run on the GPU box:  bash profiles/run_sanitizer.sh   -> gpurun_out/sanitizer_{memcheck,racecheck}_*.log
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  for part in smoke pairwise emd lsap; do
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 9 \
        python profiles/sanitize_target.py $part > gpurun_out/sanitizer_${tool}_${part}.log 2>&1
    echo "$tool $part rc=$?" | tee -a gpurun_out/sanitizer_summary.txt
    tail -n 4 gpurun_out/sanitizer_${tool}_${part}.log | tee -a gpurun_out/sanitizer_summary.txt
  done
done
