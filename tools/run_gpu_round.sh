mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  STB_NO_GRAPH=1 timeout 600 compute-sanitizer --tool $tool python tools/sanitize_target.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize target|Error|hazard" | head -8
done > gpurun_out/r02_sanitizer.txt 2>&1
cat gpurun_out/r02_sanitizer.txt
