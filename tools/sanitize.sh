#!/usr/bin/env bash
# compute-sanitizer over the small-shape cases (one tool per invocation; run each under its own gpurun call):
#   tools/sanitize.sh memcheck   |   tools/sanitize.sh racecheck
# Writes gpurun_out/sanitizer_<tool>.log; copy the summaries to profiles/.
set -uo pipefail
TOOL="${1:-memcheck}"
cd "$(dirname "${BASH_SOURCE[0]}")/.."
mkdir -p gpurun_out
timeout 900 python tools/sanitize_cases.py > gpurun_out/sanitize_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/sanitize_plain.log; exit 1; }
timeout 1500 compute-sanitizer --tool "$TOOL" --error-exitcode 1 python tools/sanitize_cases.py > "gpurun_out/sanitizer_${TOOL}.log" 2>&1
rc=$?
tail -4 "gpurun_out/sanitizer_${TOOL}.log"
exit $rc
