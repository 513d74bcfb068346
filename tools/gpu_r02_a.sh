#!/usr/bin/env bash
# round-2 GPU batch A: parity of the lean kernel + variant timing + shapes + bench baseline
set -u
G=gpurun_out
mkdir -p $G
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $G/a_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $G/a_gpu_tests.log 2>&1; echo "pytest rc=$?" >> $G/a_gpu_tests.log
cp gps_sdr_sim_b200/libgpusim.so variants/libgpusim_main.so
timeout 600 python tools/variant_bench.py main,u2,nopf > $G/a_variants.txt 2>&1
timeout 300 python tools/shape_bench.py > $G/a_shapes.txt 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 > $G/a_bench.json 2> $G/a_bench.err
tail -3 $G/a_gpu_tests.log; cat $G/a_variants.txt; cat $G/a_shapes.txt; tail -c 1500 $G/a_bench.json
