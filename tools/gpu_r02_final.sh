#!/usr/bin/env bash
# round-2 final check on one B200: the driver's own sequence (GPU tests, smoke, bench) + a long fuzz sweep
set -u
G=gpurun_out
mkdir -p $G
timeout 900 python -m pytest tests -m gpu -x -q > $G/final_gpu_tests.log 2>&1; echo "pytest rc=$?" >> $G/final_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $G/final_smoke.log 2>&1; echo "smoke rc=$?" >> $G/final_smoke.log
timeout 900 python bench.py > $G/final_bench.json 2> $G/final_bench.err; echo "bench rc=$?" >> $G/final_bench.err
timeout 1500 python tools/fuzz_parity.py 3000 7 > $G/r02_fuzz3000.txt 2>&1; echo "fuzz rc=$?" >> $G/r02_fuzz3000.txt
tail -3 $G/final_gpu_tests.log; tail -2 $G/final_smoke.log; tail -2 $G/final_bench.err; wc -l $G/final_bench.json; head -c 300 $G/final_bench.json; echo; tail -3 $G/r02_fuzz3000.txt
