#!/bin/bash
# Knob sweep on the 64 x 1 s step: sum of per-launch minima (scripts/op_table.py, 5 reps) per setting.
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" python scripts/op_table.py --reps 5 > gpurun_out/knob_$name.log 2>&1; echo "$name: $(head -1 gpurun_out/knob_$name.log)"; }
run base X=1
run cg2kb3 WV_CG2_MIN_KB=3
run cg2kb6 WV_CG2_MIN_KB=6
run pairkb3 WV_PAIR_MIN_KB=3
run mg2 WV_MATH_GROUPS=2
run mg3 WV_MATH_GROUPS=3
run rows6_64_96_128 WV_ROWS6_BN=64,96,128
run rows6_off WV_ROWS6_BN=0
run specfuse256 WV_SPEC_FUSE_MAXC=256
run restma3 WV_RES_TMA_MIN_STAGES=3
run restma6 WV_RES_TMA_MIN_STAGES=6
run epi4 WV_EPI_GROUPS=4
run ldy16 WV_LDY_ALIGN=16
run prefuse WV_PRE_FUSE=1
run noevict WV_EVICT_FIRST=0
run noserp WV_SERPENTINE=0
