#!/bin/bash
# per-opcode cycle profile of the interpreter (A/B library built with -DEB_OPPROF into elementary_b200/libelem_b200_prof.so)
mkdir -p gpurun_out
for st in 0 3; do ELEM_B200_LIB=$PWD/elementary_b200/libelem_b200_prof.so python tools/opprof.py 1250 $st | tee gpurun_out/opprof_s$st.txt; done
