# Long runs through the reference CLI with the libgpusim binding: where the host row pre-pass matters.
# usage: bash tools/cli_long.sh      (B200 box; output to /dev/null)
D=oracle/_ref/data
TIMEFORMAT="%R"
run() { # label, binary, threads, args...
  local label=$1 bin=$2 thr=$3; shift 3
  local t=$( { time env GPUSIM_HOST_THREADS=$thr GPUSIM_VERBOSE=1 integration/_build/$bin -e $D/brdc3540.14n "$@" -o /dev/null > /dev/null 2> gpurun_out/cli_long.err; } 2>&1 )
  echo "$label host_threads=$thr wall=$t s   [$(grep 'gpusim hook' gpurun_out/cli_long.err | tail -1)]"
}
nproc
for thr in 1 16; do
  run "int   2.6MSps 1-bit  -d 20000 (200k epochs, 13 GB)" gps-sdr-sim-gpu-int $thr -l 30.286502,120.032669,100 -s 2600000 -b 1 -d 20000
  run "int   2.6MSps 8-bit  -d 6000  (60k epochs, 31 GB)" gps-sdr-sim-gpu-int $thr -l 30.286502,120.032669,100 -s 2600000 -b 8 -d 6000
  run "float 2.6MSps 16-bit -d 1500  (15k epochs, 15.6 GB)" gps-sdr-sim-gpu-float $thr -l 30.286502,120.032669,100 -s 2600000 -b 16 -d 1500
  run "float 2.6MSps 1-bit  -d 3000  (30k epochs)" gps-sdr-sim-gpu-float $thr -l 30.286502,120.032669,100 -s 2600000 -b 1 -d 3000
done
