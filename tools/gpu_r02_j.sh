#!/usr/bin/env bash
# round-2 GPU batch J: segment-based chain walk (K1) + low-chip-rate linear path (K2, config 5)
set -u
G=gpurun_out
mkdir -p $G
timeout 300 python tools/k1_probe.py > $G/j_k1_probe.txt 2>&1
timeout 300 python tools/shape_bench.py > $G/j_shapes.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $G/j_gpu_tests.log 2>&1; echo "pytest rc=$?" >> $G/j_gpu_tests.log
cat $G/j_k1_probe.txt $G/j_shapes.txt; tail -3 $G/j_gpu_tests.log
