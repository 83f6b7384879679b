#!/bin/bash
# default workload at N GPUs (weak scaling, 64 x 1 s per GPU)
N=$1; mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n${N}.json 2> gpurun_out/bench_n${N}.err; echo "exit=$?"
python -c "import json;d=json.load(open('gpurun_out/bench_n${N}.json'));print('N=$N value',round(d['value']),'e2e',round(d['e2e']['value']),'ms',round(d['ms_per_step'],3),'check',d['quality']['counters_check'] and d['quality']['counters_check']['equal'])" || tail -5 gpurun_out/bench_n${N}.err
