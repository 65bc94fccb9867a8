#!/usr/bin/env bash
# compute-sanitizer over the smallest invocation of every kernel family (scripts/sanitize_cases.py).
# ONE tool per gpurun call (B200_PROFILING.md): bash scripts/sanitize.sh memcheck | racecheck | synccheck | initcheck
# The plain run goes first; the tool only runs if it exits 0.  Output: gpurun_out/sanitize_<tool>.log
set -uo pipefail
TOOL="${1:-memcheck}"
mkdir -p gpurun_out
python scripts/sanitize_cases.py > gpurun_out/sanitize_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/sanitize_plain.log; exit 1; }
timeout 1500 compute-sanitizer --tool "$TOOL" --print-limit 20 python scripts/sanitize_cases.py > "gpurun_out/sanitize_${TOOL}.log" 2>&1
echo "compute-sanitizer $TOOL rc=$?"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|SANITIZE CASES OK|Error|error" "gpurun_out/sanitize_${TOOL}.log" | head -30
