#!/usr/bin/env bash
# round-2 GPU batch G: tabulated chain walk (K1)
set -u
G=gpurun_out
mkdir -p $G variants
timeout 300 python tools/shape_bench.py > $G/g_shapes.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $G/g_gpu_tests.log 2>&1; echo "pytest rc=$?" >> $G/g_gpu_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-ncu --no-configs > $G/g_bench.json 2> $G/g_bench.err
timeout 600 python bench.py --steps 10 --warmup 3 --pipeline 0 --no-ncu --no-configs > $G/g_bench_p0.json 2> $G/g_bench_p0.err
cat $G/g_shapes.txt; tail -3 $G/g_gpu_tests.log; python -c "
import json
for f in ('g_bench.json','g_bench_p0.json'):
    try:
        d=json.load(open('$G/'+f)); print(f, d['value'], d['ms_per_step'], d['kernels'], d['e2e']['value'])
    except Exception as e: print(f, 'ERR', e)
"
