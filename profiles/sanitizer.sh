#!/bin/bash
# compute-sanitizer memcheck + racecheck on the smoke configuration (usage on the GPU box: bash profiles/sanitizer.sh <out-prefix>)
OUT=${1:-gpurun_out/r02_sanitizer}
for tool in ${TOOLS:-memcheck racecheck}; do
  timeout 1200 compute-sanitizer --tool $tool --log-file ${OUT}_${tool}.full.log python -c "import __graft_entry__ as g; g.smoke()" > ${OUT}_${tool}.out 2>&1
  echo "$tool rc=$?" >> ${OUT}_${tool}.out
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY" ${OUT}_${tool}.full.log | tail -2
  (head -60 ${OUT}_${tool}.full.log; echo ...; tail -5 ${OUT}_${tool}.full.log) > ${OUT}_${tool}.log
  rm -f ${OUT}_${tool}.full.log
done
