#!/usr/bin/env bash
# round-2 GPU batch P: everything profiles/r02_* is made from (bench lines, launch list, ncu captures, fuzz sweep)
set -u
G=gpurun_out
mkdir -p $G
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $G/p_smi.txt 2>&1
timeout 900 python bench.py --steps 40 --warmup 3 > $G/r02_bench_n1.json 2> $G/r02_bench_n1.err; echo "bench rc=$?" >> $G/r02_bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $G/r02_bench_n1_reference_arm.json 2> $G/r02_bench_ref.err
timeout 600 python bench.py --steps 2 --warmup 3 --no-ncu --no-configs > /dev/null 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $G/r02_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-ncu --no-configs > $G/p_ncu_list.log 2>&1
cap() { # name, kernel regex, profile_one args...
  local name=$1 rx=$2; shift 2
  timeout 120 python tools/profile_one.py "$@" > $G/p_${name}.txt 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s 1 -c 1 -o /tmp/r02_${name} -f python tools/profile_one.py "$@" > $G/p_${name}_ncu.log 2>&1
  # the reports are ~12 MB each and gpurun brings back 64 MiB: export the two pages the summaries are made from
  ncu -i /tmp/r02_${name}.ncu-rep --page raw --csv > $G/r02_${name}_raw.csv 2>/dev/null
  ncu -i /tmp/r02_${name}.ncu-rep --page source --csv 2>/dev/null | gzip -9 > $G/r02_${name}_src.csv.gz
  tail -1 $G/p_${name}.txt
}
cap k2_lean_sc08 k2_lean 8 1 2999 0 0 0
cap k2_lean_sc16 k2_lean 16 1 2999 0 0 0
cap k2_lean_sc01_s16 k2_lean 1 1 1560 0 0 0 100000 10
cap k2_lean_lin_sc16 k2_lean 16 1 512 0 0 0 2000000 11
cap k1_chain k1_chain 8 1 2999 0 0 0
cap k2_synth_float k2_synth 8 1 2999 0 1 0
cap k1_chain_float k1_chain 8 1 2999 0 1 0
timeout 1200 python tools/fuzz_parity.py 1500 20260219 > $G/r02_fuzz.txt 2>&1; echo "fuzz rc=$?" >> $G/r02_fuzz.txt
tail -3 $G/r02_bench_n1.err; tail -c 600 $G/r02_bench_n1_reference_arm.json; tail -4 $G/r02_fuzz.txt; du -sh $G; ls -la $G | tail -40
