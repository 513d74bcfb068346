#!/usr/bin/env bash
# round-2 GPU batch S (8 GPUs): scaling of bench.py at N = 2, 4, 8 (ranks spread over both host roots, link-aware e2e split),
# host-link probe, config 5 at full size on 8 GPUs, CLI start-up timers
set -u
G=gpurun_out
mkdir -p $G
nvidia-smi topo -m > $G/s_topo.txt 2>&1
for N in 2 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
      bench.py --gpus $N --steps 20 --warmup 3 --no-ncu --no-configs > $G/r02_bench_n$N.json 2> $G/s_bench_n$N.err
  echo "N=$N rc=$?"; tail -c 300 $G/s_bench_n$N.err
done
timeout 300 python tools/pcie_probe_multi.py 8 > $G/r02_pcie_8gpu.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 tools/bench_config5.py > $G/r02_config5_full_n8.json 2> $G/s_c5.err
D=oracle/_ref/data; H=integration/_build/gps-sdr-sim-gpu-int
TIMEFORMAT="wall=%R s"
{ for i in 1 2 3; do time env GPUSIM_VERBOSE=2 $H -e $D/brdc3540.14n -u $D/circle.csv -s 2600000 -b 16 -d 300 -o /dev/null 2>&1 | tr '\r' '\n' | grep -E "gpusim|Process time"; echo ---; done; } > $G/s_cli.txt 2>&1
python -c "
import json
for n in (2,4,8):
    try:
        d=json.load(open('$G/r02_bench_n%d.json'%n)); print(n, d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e'].get('link_rates_gbs'), d['config']['devices'])
    except Exception as e: print(n, 'ERR', e)
"; tail -5 $G/r02_pcie_8gpu.txt; tail -c 600 $G/r02_config5_full_n8.json; cat $G/s_cli.txt
