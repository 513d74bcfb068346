#!/usr/bin/env bash
# round-2 GPU batch L: no 112-register builds (pipeline = tail overlap only); bench p1/p0; CLI wall clock + start-up timers; long CLI runs
set -u
G=gpurun_out
mkdir -p $G
timeout 900 python -m pytest tests -m gpu -x -q > $G/l_gpu_tests.log 2>&1; echo "pytest rc=$?" >> $G/l_gpu_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-ncu --no-configs > $G/l_bench.json 2> $G/l_bench.err
timeout 600 python bench.py --steps 10 --warmup 3 --pipeline 0 --no-ncu --no-configs > $G/l_bench_p0.json 2> $G/l_bench_p0.err
D=oracle/_ref/data; H=integration/_build/gps-sdr-sim-gpu-int
{ for i in 1 2 3; do /usr/bin/time -f "wall=%e s" env GPUSIM_VERBOSE=2 $H -e $D/brdc3540.14n -u $D/circle.csv -s 2600000 -b 16 -d 300 -o /dev/null 2>&1 | tr '\r' '\n' | grep -E "gpusim|wall=|Process time"; echo ---; done; } > $G/l_cli.txt 2>&1
bash tools/cli_breakdown.sh >> $G/l_cli.txt 2>&1
bash tools/cli_long.sh > $G/l_cli_long.txt 2>&1
tail -3 $G/l_gpu_tests.log; python -c "
import json
for f in ('l_bench.json','l_bench_p0.json'):
    try:
        d=json.load(open('$G/'+f)); print(f, d['value'], d['ms_per_step'], d['kernels'], d['e2e']['value'])
    except Exception as e: print(f, 'ERR', e)
"; cat $G/l_cli.txt $G/l_cli_long.txt
