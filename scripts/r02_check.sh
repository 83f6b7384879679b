#!/bin/bash
# Round-2 GPU battery: precise-path report, GPU tests, per-launch table, short bench.  Usage: r02_check.sh TAG [pytest-args]
TAG=${1:-x}
shift
mkdir -p gpurun_out
python scripts/precise_report.py > gpurun_out/precise_$TAG.md 2> gpurun_out/precise_$TAG.err; echo "precise_exit=$?"
cat gpurun_out/precise_$TAG.md | cut -c1-400; tail -5 gpurun_out/precise_$TAG.err
timeout 900 python -m pytest tests -m gpu -x -q "$@" > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest_exit=$?"
tail -15 gpurun_out/pytest_$TAG.log
python scripts/op_table.py > gpurun_out/op_table_$TAG.log 2>&1; head -1 gpurun_out/op_table_$TAG.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench_exit=$?"
python -c "import json;d=json.load(open('gpurun_out/bench_$TAG.json'));print('value',d['value'],'e2e',d['e2e']['value'],'ms',d['ms_per_step'])"
