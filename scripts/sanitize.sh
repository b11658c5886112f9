#!/bin/bash
# compute-sanitizer targets (race / memory checking).  Run on a GPU box from the repository root:
#   bash scripts/sanitize.sh            -> gpurun_out/sanitize_memcheck.log, gpurun_out/sanitize_synccheck.log
# memcheck: out-of-bounds / misaligned accesses of every kernel family on edge-case shapes; synccheck: invalid barrier usage.
# (racecheck does not model mbarrier / TMA completion, which order every shared-memory hand-off of the pipelines: not run.)
set -u
mkdir -p gpurun_out
for tool in memcheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 python scripts/sanitize_smoke.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool: exit $? ; $(grep -c "========= " gpurun_out/sanitize_$tool.log) sanitizer lines ; $(grep -E "ERROR SUMMARY" gpurun_out/sanitize_$tool.log | tail -n 1)"
done
