#!/usr/bin/env bash
# round-2 final check on one B200: the driver's own sequence (GPU tests, smoke, bench, reference arm), the launch list and
# the ncu capture of the synthesis kernel the profiles are made from, and a fuzz sweep (with rows by reference into
# device-built navigation frames)
set -u
G=gpurun_out
mkdir -p $G
timeout 900 python -m pytest tests -m gpu -x -q > $G/final_gpu_tests.log 2>&1; echo "pytest rc=$?" >> $G/final_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $G/final_smoke.log 2>&1; echo "smoke rc=$?" >> $G/final_smoke.log
timeout 900 python bench.py > $G/final_bench.json 2> $G/final_bench.err; echo "bench rc=$?" >> $G/final_bench.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $G/final_bench_ref.json 2> $G/final_bench_ref.err; echo "ref rc=$?" >> $G/final_bench_ref.err
timeout 600 python bench.py --steps 2 --warmup 3 --no-ncu --no-configs > /dev/null 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $G/r02_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-ncu --no-configs > $G/final_ncu_list.log 2>&1
cap() { # name, kernel regex, profile_one args...
  local name=$1 rx=$2; shift 2
  timeout 120 python tools/profile_one.py "$@" > $G/p_${name}.txt 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s 1 -c 1 -o /tmp/r02_${name} -f python tools/profile_one.py "$@" > $G/p_${name}_ncu.log 2>&1
  ncu -i /tmp/r02_${name}.ncu-rep --page raw --csv > $G/r02_${name}_raw.csv 2>/dev/null
  ncu -i /tmp/r02_${name}.ncu-rep --page source --csv 2>/dev/null | gzip -9 > $G/r02_${name}_src.csv.gz
  tail -1 $G/p_${name}.txt
}
cap k2_lean_sc08 k2_lean 8 1 2999 0 0 0
timeout 1500 python tools/fuzz_parity.py ${FUZZ_CASES:-1500} 20261019 > $G/r02_fuzz_final.txt 2>&1; echo "fuzz rc=$?" >> $G/r02_fuzz_final.txt
tail -3 $G/final_gpu_tests.log; tail -2 $G/final_smoke.log; tail -2 $G/final_bench.err; wc -l $G/final_bench.json; head -c 300 $G/final_bench.json; echo; tail -c 300 $G/final_bench_ref.json; echo; tail -3 $G/r02_fuzz_final.txt
