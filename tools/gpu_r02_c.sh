#!/usr/bin/env bash
# round-2 GPU batch C: pipelined-constants chain walk (K1), CLI start-up breakdown
set -u
G=gpurun_out
mkdir -p $G
timeout 300 python tools/shape_bench.py > $G/c_shapes.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $G/c_gpu_tests.log 2>&1; echo "pytest rc=$?" >> $G/c_gpu_tests.log
D=oracle/_ref/data; H=integration/_build/gps-sdr-sim-gpu-int
{ for i in 1 2 3; do /usr/bin/time -f "wall=%e s" env GPUSIM_VERBOSE=2 $H -e $D/brdc3540.14n -u $D/circle.csv -s 2600000 -b 16 -d 300 -o /dev/null 2>&1 | tr '\r' '\n' | grep -E "gpusim|wall=|Process time"; echo ---; done; } > $G/c_cli.txt 2>&1
bash tools/cli_breakdown.sh >> $G/c_cli.txt 2>&1
cat $G/c_shapes.txt; tail -3 $G/c_gpu_tests.log; cat $G/c_cli.txt
