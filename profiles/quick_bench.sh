#!/bin/bash
# quick GPU iteration: the GPU test suite, then a short device-resident bench with the per-phase table
# usage (through gpurun): bash profiles/quick_bench.sh <tag> [pytest -k filter | "none"]
TAG=${1:-x}
FILT=${2:-}
mkdir -p gpurun_out
if [ "$FILT" != "none" ]; then
  if [ -n "$FILT" ]; then python -m pytest tests -m gpu -x -q -k "$FILT" 2>&1 | tail -5; else python -m pytest tests -m gpu -x -q 2>&1 | tail -5; fi
fi
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary > gpurun_out/${TAG}.json 2> gpurun_out/${TAG}.err
echo rc=$?
python - <<P
import json
l=[x for x in open('gpurun_out/${TAG}.json') if x.startswith('{')]
if l:
    d=json.loads(l[-1])
    print(d['value'], d['ms_per_step'], d['parity'], d['roofline']['frac'], d.get('latency'))
    for k,v in sorted(d['phases_ms'].items(), key=lambda kv:-kv[1]['ms_per_step']): print('  ',k, v['ms_per_step'])
else:
    print(open('gpurun_out/${TAG}.err').read()[-3000:])
P
