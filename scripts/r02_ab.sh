#!/bin/bash
# A/B of the precision switches on the 64 x 1 s step.  Usage: r02_ab.sh TAG
TAG=${1:-x}
mkdir -p gpurun_out
for cfg in "1 1" "0 1" "1 0" "0 0"; do
  set -- $cfg
  WV_EXACT_BITS=$1 WV_EXACT_MASK=$2 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-profile > gpurun_out/bench_${TAG}_b$1m$2.json 2> gpurun_out/bench_${TAG}_b$1m$2.err
  python -c "import json;d=json.load(open('gpurun_out/bench_${TAG}_b$1m$2.json'));print('exact_bits=$1 exact_mask=$2 value',round(d['value']),'e2e',round(d['e2e']['value']),'ms',round(d['ms_per_step'],3),'recheck/step',d['config'].get('detector_rechecked_clips_per_step'))"
done
