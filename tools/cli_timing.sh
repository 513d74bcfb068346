#!/usr/bin/env bash
# Wall-clock of the reference CLI with the libgpusim binding (output to /dev/null), for 1..N GPUs,
# next to the unmodified reference on the same argv (shorter duration, CPU).  usage: tools/cli_timing.sh [max_gpus]
D=oracle/_ref/data; H=integration/_build/gps-sdr-sim-gpu-int; R=oracle/_ref/gps-sdr-sim-int
MAXG=${1:-1}
TIMEFORMAT="%R"
run() { # label, env..., -- argv
  local label="$1"; shift
  local t; t=$( { time env "$@" > /dev/null 2> /tmp/cli_err.txt; } 2>&1 )
  echo "$label wall=${t}s $(tr '\r' '\n' < /tmp/cli_err.txt | grep -E 'Process time|ERROR' | tail -1)"
}
for n in $(seq 1 $MAXG); do
  run "20MSps-16bit-60s gpus=$n" GPUSIM_DEVICES=$n GPUSIM_BATCH_EPOCHS=32 $H -e $D/brdc3540.14n -l 30.286502,120.032669,100 -d 60 -s 20000000 -b 16 -o /dev/null
done
for n in $(seq 1 $MAXG); do
  run "2.6MSps-16bit-300s(circle) gpus=$n" GPUSIM_DEVICES=$n GPUSIM_BATCH_EPOCHS=128 $H -e $D/brdc3540.14n -u $D/circle.csv -s 2600000 -b 16 -o /dev/null
done
run "reference-CPU 20MSps-16bit-3s" $R -e $D/brdc3540.14n -l 30.286502,120.032669,100 -d 3 -s 20000000 -b 16 -o /dev/null
run "reference-CPU 2.6MSps-16bit-30s(circle)" $R -e $D/brdc3540.14n -u $D/circle.csv -d 30 -s 2600000 -b 16 -o /dev/null
