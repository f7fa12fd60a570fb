#!/bin/bash
OUT=gpurun_out/sanitize
mkdir -p $OUT
for tool in memcheck initcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize_small.py > $OUT/$tool.txt 2>&1
  echo "== $tool"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize run done" $OUT/$tool.txt; grep -B2 -A12 "Uninitialized\|Invalid\|hazard" $OUT/$tool.txt | head -60
done
