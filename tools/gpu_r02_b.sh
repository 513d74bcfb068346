#!/usr/bin/env bash
# round-2 GPU batch B: new chain walk (K1) + lean K2 + new bench.py
set -u
G=gpurun_out
mkdir -p $G
timeout 900 python -m pytest tests -m gpu -x -q > $G/b_gpu_tests.log 2>&1; echo "pytest rc=$?" >> $G/b_gpu_tests.log
timeout 300 python tools/shape_bench.py > $G/b_shapes.txt 2>&1
cp gps_sdr_sim_b200/libgpusim.so variants/libgpusim_main.so
timeout 300 python tools/variant_bench.py main > $G/b_variants.txt 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > $G/b_bench.json 2> $G/b_bench.err; echo "bench rc=$?" >> $G/b_bench.err
timeout 600 python bench.py --steps 10 --warmup 3 --pipeline 0 --no-ncu --no-configs > $G/b_bench_p0.json 2> $G/b_bench_p0.err
timeout 600 python tools/fuzz_parity.py 600 7 > $G/b_fuzz.txt 2>&1; echo "fuzz rc=$?" >> $G/b_fuzz.txt
tail -3 $G/b_gpu_tests.log; cat $G/b_shapes.txt $G/b_variants.txt; tail -4 $G/b_fuzz.txt; tail -5 $G/b_bench.err; python -c "
import json
for f in ('b_bench.json','b_bench_p0.json'):
    try:
        d=json.load(open('$G/'+f)); print(f, d['value'], d['ms_per_step'], d['kernels'], d['e2e']['value'], d['data'][:40]); print(json.dumps(d['roofline'])[:1500]); print(json.dumps(d.get('configs'))[:3000]); print(d['output_check'])
    except Exception as e: print(f, 'ERR', e)
"
