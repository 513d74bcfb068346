#!/usr/bin/env bash
# round-2 GPU batch I: checkpoint-driven chain walk + blocked (coalesced) checkpoint layout
set -u
G=gpurun_out
mkdir -p $G
timeout 300 python tools/k1_probe.py > $G/i_k1_probe.txt 2>&1
timeout 300 python tools/shape_bench.py > $G/i_shapes.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $G/i_gpu_tests.log 2>&1; echo "pytest rc=$?" >> $G/i_gpu_tests.log
cat $G/i_k1_probe.txt $G/i_shapes.txt; tail -3 $G/i_gpu_tests.log
