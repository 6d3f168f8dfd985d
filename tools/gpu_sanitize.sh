#!/bin/bash
mkdir -p gpurun_out
for TOOL in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $TOOL --print-limit 20 python tools/sanitize_small.py > gpurun_out/sanitize_$TOOL.log 2>&1
  echo "$TOOL rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|SANITIZE_RUN|Error|hazard" gpurun_out/sanitize_$TOOL.log | head -8
done
