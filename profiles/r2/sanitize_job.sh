# compute-sanitizer evidence: memcheck, racecheck, synccheck over profiles/tools/sanitize_cases.py
mkdir -p gpurun_out/sanitizer
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python profiles/tools/sanitize_cases.py > gpurun_out/sanitizer/$tool.log 2>&1
  echo "$tool exit $?" >> gpurun_out/sanitizer/$tool.log
  tail -4 gpurun_out/sanitizer/$tool.log
done
