#!/bin/bash
# compute-sanitizer over small invocations of every kernel family.  ONE tool per GPU call
# (B200_PROFILING.md: the four tools in one call once left a GPU unusable):
#     gpurun -- 'tools/sanitize.sh memcheck'      (then racecheck, initcheck, synccheck in calls of their own)
# The plain run goes first; the sanitizer only runs if it exits 0.  Logs: gpurun_out/sanitize_<tool>.log
set -u
tool=${1:-memcheck}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/sanitize_cases.py > gpurun_out/sanitize_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/sanitize_plain.log; exit 1; }
timeout 1500 compute-sanitizer --tool "$tool" --launch-timeout 120 --error-exitcode 3 python tools/sanitize_cases.py > "gpurun_out/sanitize_${tool}.log" 2>&1
rc=$?
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|all sanitizer cases ran" "gpurun_out/sanitize_${tool}.log" | tail -3
echo "compute-sanitizer --tool $tool: exit $rc"
exit $rc
