#!/bin/bash
# operand-starvation probe: role counters of single layers under forced tile shapes and at smaller batches
mkdir -p gpurun_out; out=gpurun_out/probe2.txt; : > $out
run() { echo "### $*" >> $out; env "$@" SMAPB_ROLES=1 timeout 120 python tools/conv_micro.py l3_c2 l3_c1 l4_c2 up4_1x1 l1_c1 >> $out 2>&1; }
run X=1
run SMAPB_FORCE_TILE=128,2
run SMAPB_FORCE_TILE=128,1
run CONV_MICRO_BATCH=4
run CONV_MICRO_BATCH=2
cat $out
