#!/bin/bash
# Free experiments on the 64 x 1 s step: whole-step CUDA graphs, sub-batch sizes (L2 residency), PDL.  Usage: r02_sweep.sh TAG
TAG=${1:-x}
mkdir -p gpurun_out
run() {  # name, env..., -- bench args
  name=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-profile "$@" > gpurun_out/sweep_${TAG}_$name.json 2> gpurun_out/sweep_${TAG}_$name.err
  python -c "import json;d=json.load(open('gpurun_out/sweep_${TAG}_$name.json'));print('$name value',round(d['value']),'e2e',round(d['e2e']['value']),'ms',round(d['ms_per_step'],3))" || tail -3 gpurun_out/sweep_${TAG}_$name.err
}
run base X=1 --
run graph WV_GRAPH_MAX_SAMPLES=4000000 --
run pdl WV_PDL=1 --
run chunk32 X=1 -- --chunk-seconds 32
run chunk16 X=1 -- --chunk-seconds 16
run chunk8 X=1 -- --chunk-seconds 8
run fastL WV_EXACT_MASK=0 --
run fastLD WV_EXACT_MASK=0 WV_EXACT_BITS=0 --
