#!/usr/bin/env bash
# Refresh profiles/ from the artefacts a GPU run left in gpurun_out/ (see the gpurun command in
# profiles/README.md).  usage: tools/make_profiles.sh r01
set -euo pipefail
R=${1:-r01}
G=gpurun_out
P=profiles
SC=$((2999*260000*13))
mkdir -p $P build
cp $G/r01_launches.csv $P/${R}_bench_launches.csv
cp $G/bench_r01.json $P/${R}_bench_n1.json
ncu -i $G/r01_k2_synth_sc08.ncu-rep --page source --csv 2>/dev/null > build/k2_src.csv
python tools/ncu_regions.py build/k2_src.csv $SC > build/k2_regions.txt
python - <<PY
import json,subprocess,csv
out=subprocess.run(["ncu","-i","$G/r01_k2_synth_sc08.ncu-rep","--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines())); h,u,v=rows[0],rows[1],rows[2]
def val(k):
    x=float(v[h.index(k)]); un=u[h.index(k)]
    return x*{'Gbyte':1e9,'Mbyte':1e6,'Kbyte':1e3,'byte':1}.get(un,1)
rd,wr=val('dram__bytes_read.sum'),val('dram__bytes_write.sum')
issue=float(v[h.index('smsp__issue_active.avg.pct_of_peak_sustained_active')])
inst=float(v[h.index('smsp__inst_executed.sum')])
json.dump({"k2_synth_sc08_dram_bytes_per_launch": int(rd+wr), "dram_bytes_read": int(rd), "dram_bytes_write": int(wr),
           "issue_active_pct": issue, "thread_instructions_per_sample_channel": inst*32/(2999*260000*13),
           "algorithmic_bytes_per_launch": 2999*260000*2,
           "source": "profiles/${R}_k2_synth_sc08_ncu.md (ncu --set full, one launch of k2_synth<AccF32x2,8,32,0,1>, 2999 epochs x 13 channels)"},
          open("$P/traffic.json","w"),indent=1)
PY
{
echo "# $R — k2_synth<AccF32x2, 8, 32, 0, 1> (the 112-register build bench.py's back-to-back steps run): ncu --set full, one launch, bench workload"; echo
echo "Command (B200, driver 580, CUDA 12.9): \`ncu --set full --clock-control none --import-source on -k regex:k2_synth -s 1 -c 1 python tools/profile_one.py 8 1 2999\`"
echo "(2999 epochs x 13 channels x 260 000 samples, 8-bit IQ = the bench.py workload; the same command ran first without ncu.)"; echo
python tools/ncu_summary.py $G/r01_k2_synth_sc08.ncu-rep; echo
echo "## Executed instructions and stall samples (ncu source page, tools/ncu_regions.py)"; echo; echo '```'; grep -v "^F2I marks\|^prologue\|^fast loop\|^wrap loop" build/k2_regions.txt; echo '```'
} > $P/${R}_k2_synth_sc08_ncu.md
{
echo "# $R — k1_chain<0>: ncu --set full, one launch, bench workload"; echo
echo "\`ncu --set full --clock-control none --import-source on -k regex:k1_chain -s 1 -c 1 python tools/profile_one.py 8 1 2999\`"; echo
python tools/ncu_summary.py $G/r01_k1_chain.ncu-rep
} > $P/${R}_k1_chain_ncu.md
if [ -f $G/k2_r01b.ncu-rep ]; then
{
echo "# $R — k2_synth<AccF32x2, 8, 32, 0, 0> (the 128-register build a single call runs): ncu --set full, one launch, bench workload"; echo
echo "\`ncu --set full --clock-control none --import-source on -k regex:k2_synth -s 1 -c 1 python tools/profile_one.py 8 1 2999 0 0 0\`"; echo
python tools/ncu_summary.py $G/k2_r01b.ncu-rep
} > $P/${R}_k2_synth_sc08_128reg_ncu.md
fi
if [ -f $G/r01_k2_synth_float.ncu-rep ]; then
{
echo "# $R — k2_synth<AccF32x2, 8, 32, 2, false> (FLOAT_CARR_PHASE hosts: double carrier phase, 512 threads x runs of 32): ncu --set full, one launch, bench workload shape"; echo
echo "\`ncu --set full --clock-control none --import-source on -k regex:k2_synth -s 1 -c 1 python tools/profile_one.py 8 1 2999 0 1 0\`"; echo
python tools/ncu_summary.py $G/r01_k2_synth_float.ncu-rep
} > $P/${R}_k2_synth_float_ncu.md
fi
{
echo "# $R — launch list of \`python bench.py --steps 2 --warmup 3\` under ncu"; echo
echo "\`ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv\` (cold-cache, serialised: shares, not absolutes). Raw list: ${R}_bench_launches.csv"; echo
python - <<PY
import json
b=json.load(open("$G/bench_r01.json"))
k2=b["kernels"]["k2_synth_ms"]; st=b["ms_per_step"]
print("Reading it next to the bench line: ncu runs one kernel at a time, so no call ever finds its predecessor in flight and the library takes its non-overlapped route (chain kernel, then the 128-register synthesis kernel \`<…, 0, 0>\`), and K2 is about nine tenths of a serialised step.  In the un-profiled timed region the steps are issued back to back, the chain kernel of step i+1 runs beside the synthesis kernel of step i (112-register build \`<…, 0, 1>\`, profiled in ${R}_k2_synth_sc08_ncu.md) and the launch stream carries nothing but K2 launches: their average duration, %.2f ms, is %.0f %% of the %.2f ms step (bench.py \`kernels\`; the last step's K2, with no chain kernel beside it any more, takes %.2f ms).\n" % (k2, 100*k2/st, st, b["kernels"].get("k2_synth_last_step_ms", k2)))
PY
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open('$G/r01_launches.csv')) if len(r)>5]
hdr=rows[0]; ix={h:i for i,h in enumerate(hdr)}
big=collections.defaultdict(lambda:[0,0.0]); small=collections.defaultdict(lambda:[0,0.0])
# bench.py launches in order: (3 warm-up + 2 timed) device-resident steps = 5 x (k1, k2), then the e2e passes
for i,r in enumerate(rows[1:]):
    name=r[ix['Kernel Name']].split('(')[0]; v=float(r[ix['Metric Value']].replace(',',''))/1e6
    tgt = big if i < 10 else small
    tgt[name][0]+=1; tgt[name][1]+=v
print("| phase | kernel | launches | total ms | ms / launch | share of phase |\n|---|---|---|---|---|---|")
for label,d in (("device-resident steps (value): 3 warm-up + 2 timed, whole 2999-epoch table per launch",big),("e2e steps: one chain launch per table, 64 MiB sub-batches of synthesis (2 warm-up + 2 timed passes)",small)):
    tot=sum(v for _,v in d.values())
    for k,(n,v) in sorted(d.items(),key=lambda kv:-kv[1][1]): print(f"| {label} | \`{k}\` | {n} | {v:.3f} | {v/n:.3f} | {100*v/tot:.1f} % |")
PY
} > $P/${R}_bench_launch_shares.md
echo "profiles refreshed:"; ls $P
