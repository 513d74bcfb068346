#!/usr/bin/env bash
# round-2 GPU batch E: ncu source-level profile of k2_lean (SC08, bench shape) + channel-loop variants
set -u
G=gpurun_out
mkdir -p $G variants
cp gps_sdr_sim_b200/libgpusim.so variants/libgpusim_main.so
timeout 600 python tools/variant_bench.py main,u2,nopf > $G/e_variants.txt 2>&1
timeout 120 python tools/profile_one.py 8 1 2999 0 0 0 > $G/e_profile_one.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k2_lean -s 1 -c 1 -o $G/r02_k2_lean_sc08 -f python tools/profile_one.py 8 1 2999 0 0 0 > $G/e_ncu.log 2>&1
cat $G/e_variants.txt $G/e_profile_one.txt; tail -3 $G/e_ncu.log; ls -la $G
