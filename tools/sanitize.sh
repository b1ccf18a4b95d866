#!/bin/bash
# compute-sanitizer over one small batch per kernel arm (tools/sanitize_case.py); logs under gpurun_out/sanitize_*.log.
# usage (on the GPU box): bash tools/sanitize.sh [tool ...]   tools: memcheck racecheck synccheck initcheck
mkdir -p gpurun_out
python tools/sanitize_case.py > gpurun_out/sanitize_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/sanitize_plain.log; exit 1; }
for tool in ${@:-memcheck racecheck synccheck}; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 --log-file gpurun_out/sanitize_$tool.log \
      python tools/sanitize_case.py > gpurun_out/sanitize_${tool}_stdout.log 2>&1
  echo "$tool rc=$?"; tail -n 3 gpurun_out/sanitize_$tool.log; grep -c " ok" gpurun_out/sanitize_${tool}_stdout.log
done
