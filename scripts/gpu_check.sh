#!/bin/bash
# Standard GPU battery: parity tests, per-launch table, bench (PDL on / off).  Usage: gpu_check.sh TAG
TAG=${1:-x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest_exit=$?"
tail -3 gpurun_out/pytest_$TAG.log
python scripts/op_table.py > gpurun_out/op_table_$TAG.log 2>&1; head -1 gpurun_out/op_table_$TAG.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench_exit=$?"
python -c "import json;d=json.load(open('gpurun_out/bench_$TAG.json'));print('value',d['value'],'e2e',d['e2e']['value'],'ms',d['ms_per_step'])"
WV_EPI_GROUPS=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-profile > gpurun_out/bench_${TAG}_nopdl.json 2> gpurun_out/bench_${TAG}_nopdl.err
python -c "import json;d=json.load(open('gpurun_out/bench_${TAG}_nopdl.json'));print('nogroups value',d['value'],'e2e',d['e2e']['value'],'ms',d['ms_per_step'])"
