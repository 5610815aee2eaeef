#!/bin/bash
mkdir -p gpurun_out; out=gpurun_out/halo.txt; : > $out
timeout 300 python -m pytest tests/test_conv_gpu.py -x -q >> $out 2>&1; echo "pytest conv rc=$?" >> $out
echo "### l1_c2 default (halo)" >> $out
SMAPB_ROLES=1 SMAPB_TIMELINE=1 timeout 120 python tools/conv_micro.py l1_c2 >> $out 2>&1
echo "### l1_c2 forced 64,2" >> $out
SMAPB_FORCE_TILE=64,2 SMAPB_ROLES=1 timeout 120 python tools/conv_micro.py l1_c2 >> $out 2>&1
tail -40 $out
