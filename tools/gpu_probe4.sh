#!/bin/bash
mkdir -p gpurun_out; out=gpurun_out/probe4.txt; : > $out
SMAPB_TIMELINE=1 timeout 120 python tools/conv_micro.py l3_c2 l3_c1 l1_c2 2>&1 | grep -a timeline >> $out
SMAPB_FORCE_TILE=128,2 SMAPB_TIMELINE=1 timeout 120 python tools/conv_micro.py l3_c1 l4_c2 2>&1 | grep -a timeline >> $out
cat $out
