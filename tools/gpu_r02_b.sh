#!/usr/bin/env bash
# round-2 GPU batch B: new chain walk (K1) + lean K2 with state prefetch only in the 128-register build
set -u
G=gpurun_out
mkdir -p $G
timeout 900 python -m pytest tests -m gpu -x -q > $G/b_gpu_tests.log 2>&1; echo "pytest rc=$?" >> $G/b_gpu_tests.log
timeout 300 python tools/shape_bench.py > $G/b_shapes.txt 2>&1
cp gps_sdr_sim_b200/libgpusim.so variants/libgpusim_main.so
timeout 300 python tools/variant_bench.py main > $G/b_variants.txt 2>&1
timeout 600 python tools/fuzz_parity.py 600 7 > $G/b_fuzz.txt 2>&1; echo "fuzz rc=$?" >> $G/b_fuzz.txt
tail -3 $G/b_gpu_tests.log; cat $G/b_shapes.txt $G/b_variants.txt; tail -4 $G/b_fuzz.txt
