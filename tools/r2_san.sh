set -x
O=gpurun_out/r2san
mkdir -p $O
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize_smoke.py > $O/$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "^ok|ERROR SUMMARY|RACECHECK SUMMARY|error|Error" $O/$tool.log | tail -8
done
