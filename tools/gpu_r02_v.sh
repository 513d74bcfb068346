#!/usr/bin/env bash
# round-2 verification after the bench import fix: bench N=1 (full line), reference arm, e2e sub-batch sweep
set -u
G=gpurun_out
mkdir -p $G
timeout 900 python bench.py > $G/v_bench_n1.json 2> $G/v_bench_n1.err; echo "bench rc=$?" >> $G/v_bench_n1.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $G/v_bench_ref.json 2> $G/v_bench_ref.err; echo "ref rc=$?" >> $G/v_bench_ref.err
timeout 300 python tools/e2e_sweep.py > $G/v_e2e_sweep.txt 2>&1; echo "sweep rc=$?" >> $G/v_e2e_sweep.txt
tail -2 $G/v_bench_n1.err; head -c 400 $G/v_bench_n1.json; echo; tail -2 $G/v_bench_ref.err; head -c 300 $G/v_bench_ref.json; echo; cat $G/v_e2e_sweep.txt
