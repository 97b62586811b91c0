#!/bin/bash
# compute-sanitizer over every kernel family (race detection / sanitizers).  One log per tool under gpurun_out/; the
# summary lines ("ERROR SUMMARY", hazards) go to gpurun_out/sanitize_summary.txt — copy that into profiles/ when it is clean.
# usage: scripts/sanitize.sh [tools] [targets]   e.g.  scripts/sanitize.sh "memcheck racecheck" solver,unet
set -u
TOOLS=${1:-"memcheck racecheck synccheck initcheck"}
TARGETS=${2:-"solver,unet,vae,clip"}
mkdir -p gpurun_out
SUM=gpurun_out/sanitize_summary.txt; : > $SUM
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in $TOOLS; do
  echo "== compute-sanitizer --tool $tool ($TARGETS)" | tee -a $SUM
  extra=""
  [ "$tool" = "racecheck" ] && extra="--racecheck-report all"
  timeout 900 $CS --tool $tool $extra --error-exitcode 86 --print-limit 30 python tests/tools/sanitize_target.py $TARGETS > gpurun_out/sanitize_$tool.log 2>&1
  echo "rc=$?" | tee -a $SUM
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|Uninitialized|sanitize target done|solver \(|unet rel|^vae|clip logits" gpurun_out/sanitize_$tool.log | cut -c1-240 | sort | uniq -c | sort -rn | head -30 | tee -a $SUM
done
