#!/bin/bash
# compute-sanitizer over a small run of every kernel family (race / memory checks).  Run on a GPU box:
#   tools/sanitize.sh [memcheck|racecheck|initcheck|synccheck]   (default: memcheck, then racecheck)
# Small worlds (the tools slow kernels down 10-100x): 33 envs, 64 x 64 images, short episodes so that the standby pipeline,
# the masked terminal-observation raster and the scanline raster's fallback pass all run.  Exit code != 0 on any finding.
set -u
cd "$(dirname "$0")/.."
tools=${1:-"memcheck racecheck"}
rc=0
for tool in $tools; do
  for task in edge surface balance push roll; do
    echo "== compute-sanitizer --tool $tool : $task"
    TG_SANITIZE_TASK=$task compute-sanitizer --tool "$tool" --error-exitcode 9 --print-limit 5 python tools/sanitize_run.py 2>&1 | tail -4
    r=${PIPESTATUS[0]}
    if [ "$r" != "0" ]; then rc=$r; fi
  done
done
exit $rc
