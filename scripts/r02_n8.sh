#!/bin/bash
# 8-GPU runs: the default workload (64 x 1 s per GPU, weak scaling) and BASELINE config 3 (512 x 10 s sharded over 8 ranks,
# BER / MIoU counters all-reduced over NCCL and checked against rank 0's recomputation from the ranks' seeds).
mkdir -p gpurun_out
N=${1:-8}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@"; }
run --steps 10 --warmup 3 > gpurun_out/bench_n${N}.json 2> gpurun_out/bench_n${N}.err; echo "default exit=$?"
run --steps 4 --warmup 3 --clips 64 --seconds 10 > gpurun_out/bench_n${N}_config3.json 2> gpurun_out/bench_n${N}_config3.err; echo "config3 exit=$?"
for f in gpurun_out/bench_n${N}.json gpurun_out/bench_n${N}_config3.json; do
  python -c "import json;d=json.load(open('$f'));print('$f', 'value',round(d['value']),'e2e',round(d['e2e']['value']),'ms',round(d['ms_per_step'],3),'check',d['quality']['counters_check'] and d['quality']['counters_check']['equal'])" || tail -5 ${f%.json}.err
done
