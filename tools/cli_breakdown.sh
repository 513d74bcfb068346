D=oracle/_ref/data; H=integration/_build/gps-sdr-sim-gpu-int
TIMEFORMAT="%R"
for d in 0.3 30 100 300; do t=$( { time $H -e $D/brdc3540.14n -u $D/circle.csv -s 2600000 -b 16 -d $d -o /dev/null > /dev/null 2>&1; } 2>&1 ); echo "circle 2.6MSps -d $d wall=$t"; done
for d in 0.3 10 60; do t=$( { time env GPUSIM_BATCH_EPOCHS=32 $H -e $D/brdc3540.14n -l 30.286502,120.032669,100 -s 20000000 -b 16 -d $d -o /dev/null > /dev/null 2>&1; } 2>&1 ); echo "static 20MSps -d $d wall=$t"; done
t=$( { time python -c "import ctypes; l=ctypes.CDLL('libcudart.so.12'); p=ctypes.c_void_p(); l.cudaMalloc(ctypes.byref(p), 1024); l.cudaDeviceSynchronize()" ; } 2>&1 ); echo "bare cuda init (python+cudart) wall=$t"
