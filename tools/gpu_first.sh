#!/bin/bash
# first GPU contact: conv unit tests, association parity, backbone parity (each under its own timeout)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 600 python -m pytest tests/test_conv_gpu.py -x -q -m gpu > gpurun_out/conv.log 2>&1; echo "conv rc=$?" >> gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_assoc_gpu.py -q -m gpu > gpurun_out/assoc.log 2>&1; echo "assoc rc=$?" >> gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_backbone_gpu.py -x -q -m gpu > gpurun_out/backbone.log 2>&1; echo "backbone rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -30 gpurun_out/conv.log; tail -30 gpurun_out/assoc.log; tail -30 gpurun_out/backbone.log
