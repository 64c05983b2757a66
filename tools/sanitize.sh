#!/bin/bash
# One compute-sanitizer tool per gpurun call (B200_PROFILING.md): tools/sanitize.sh memcheck|racecheck|synccheck
#   gpurun --timeout 900 -- 'bash tools/sanitize.sh racecheck'
# The plain run comes first (a faulting program must not be run under the tool); the summary lands in gpurun_out/.
set -u
tool=${1:-memcheck}
mkdir -p gpurun_out
python tools/sanitize_cases.py > gpurun_out/sanitize_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/sanitize_plain.log; exit 1; }
timeout 1500 compute-sanitizer --tool "$tool" --print-limit 20 --error-exitcode 9 \
    python tools/sanitize_cases.py > "gpurun_out/sanitize_${tool}.log" 2>&1
rc=$?
echo "compute-sanitizer --tool $tool exit code $rc"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_cases:|Error|hazard" "gpurun_out/sanitize_${tool}.log" | head -40
exit $rc
