#!/bin/bash
# Round-2 ncu evidence: launch list of one 64 x 1 s step (time + DRAM bytes) and --set full captures of the largest
# generator launches (indices = position among the gemm_sm100_kernel launches of one generator pass, scripts/op_table.py order).
# The reports are read on the box (scripts/ncu_read.py -> text); only the last one is kept (gpurun_out is capped at 64 MiB).
TAG=${1:-r02}
mkdir -p gpurun_out
python -c "import sys; sys.path.insert(0,'.'); import bench; print(bench.build_id())" > gpurun_out/build_id_$TAG.txt
bash scripts/ncu_launches.sh $TAG
for spec in "4 1" "31 1" "44 3" "52 2"; do
  set -- $spec
  ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_sm100_kernel \
      -s $1 -c $2 -f -o gpurun_out/prof_${TAG}_s$1 python scripts/ncu_target.py --which g > gpurun_out/ncu_full_${TAG}_s$1.log 2>&1
  echo "full s=$1 c=$2 exit=$?"
  python scripts/ncu_read.py gpurun_out/prof_${TAG}_s$1.ncu-rep > gpurun_out/ncu_read_${TAG}_s$1.txt 2>&1
  [ "$1" != "52" ] && rm -f gpurun_out/prof_${TAG}_s$1.ncu-rep
done
du -sh gpurun_out
