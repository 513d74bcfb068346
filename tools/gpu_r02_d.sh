#!/usr/bin/env bash
# round-2 GPU batch D: state of the tree after the re-entry (lean K2, new K1 walk, new bench.py)
set -u
G=gpurun_out
mkdir -p $G variants
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $G/d_smi.txt 2>&1
timeout 300 python tools/shape_bench.py > $G/d_shapes.txt 2>&1
cp gps_sdr_sim_b200/libgpusim.so variants/libgpusim_main.so
timeout 300 python tools/variant_bench.py main > $G/d_variants.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $G/d_gpu_tests.log 2>&1; echo "pytest rc=$?" >> $G/d_gpu_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > $G/d_bench.json 2> $G/d_bench.err; echo "bench rc=$?" >> $G/d_bench.err
cat $G/d_shapes.txt $G/d_variants.txt; tail -3 $G/d_gpu_tests.log; tail -5 $G/d_bench.err; tail -c 3000 $G/d_bench.json
