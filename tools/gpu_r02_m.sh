#!/usr/bin/env bash
# round-2 GPU batch M: chain walk with prefetched binade constants (K1 + host carrier advance)
set -u
G=gpurun_out
mkdir -p $G
timeout 300 python tools/k1_probe.py > $G/m_k1_probe.txt 2>&1
timeout 300 python tools/shape_bench.py > $G/m_shapes.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $G/m_gpu_tests.log 2>&1; echo "pytest rc=$?" >> $G/m_gpu_tests.log
bash tools/cli_long.sh > $G/m_cli_long.txt 2>&1
cat $G/m_k1_probe.txt $G/m_shapes.txt; tail -3 $G/m_gpu_tests.log; cat $G/m_cli_long.txt
