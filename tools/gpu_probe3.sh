#!/bin/bash
mkdir -p gpurun_out; out=gpurun_out/probe3.txt; : > $out
SMAPB_TIMELINE=1 timeout 120 python tools/conv_micro.py l3_c2 l3_c1 l3_c3 l4_c2 l2_c3 l1_c3 l1_c1 >> $out 2>&1
echo "### forced 128,2" >> $out
SMAPB_FORCE_TILE=128,2 SMAPB_TIMELINE=1 timeout 120 python tools/conv_micro.py l3_c2 l3_c1 l4_c2 >> $out 2>&1
cat $out
