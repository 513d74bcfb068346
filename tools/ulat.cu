// Dependent-chain latencies of the instructions on the serial paths of the synthesis kernels (sm_100a).
// One warp, one chain, clock64() around 4096 dependent steps.   nvcc -arch=sm_100a -O3 --fmad=false -o ulat ulat.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE> __global__ void chain(double *out, long long *cycles, double x, double d, int n)
{
    double c = x;
    uint32_t u = (uint32_t)n;
    long long t0 = clock64();
    for (int i = 0; i < n; i++) {
#pragma unroll
        for (int j = 0; j < 16; j++) {
            if (MODE == 0) {            // DADD
                c = __dadd_rn(c, d);
            } else if (MODE == 1) {     // DADD + compare + select (carrier wrap as the compiler writes it)
                c = __dadd_rn(c, d);
                double w = __dadd_rn(c, -512.0);
                c = c >= 512.0 ? w : c;
            } else if (MODE == 2) {     // DADD + integer test of (c-512) + select
                c = __dadd_rn(c, d);
                double w = __dadd_rn(c, -512.0);
                c = __double2hiint(w) >= 0 ? w : c;
            } else if (MODE == 3) {     // DADD.RD to the index + LOP3 + IMAD + LDS-free
                c = __dadd_rn(c, d);
                u ^= (uint32_t)__double2loint(__dadd_rd(c, 4503599627370496.0));
            } else if (MODE == 4) {     // IMAD chain
                u = u * 3u + 1u;
            } else if (MODE == 5) {     // FFMA chain
                float f = __int_as_float(u);
                f = __fmaf_rn(f, 1.0001f, 1.0f);
                u = __float_as_int(f);
            }
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = c + (double)u;
    if (threadIdx.x == 0)
        *cycles = t1 - t0;
}

int main()
{
    double *out; long long *cyc, h;
    cudaMalloc(&out, 32 * sizeof(double)); cudaMalloc(&cyc, sizeof(long long));
    const int n = 256;
    const char *names[] = {"DADD", "DADD + DADD/DSETP + FSEL (wrap)", "DADD + DADD + ISETP + FSEL", "DADD (+ DADD.RD side)", "IMAD", "FFMA"};
#define RUN(M) chain<M><<<1, 32>>>(out, cyc, 1.0, 0.37, n); chain<M><<<1, 32>>>(out, cyc, 1.0, 0.37, n); cudaDeviceSynchronize(); \
    cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost); printf("%-36s %7.2f cycles per step\n", names[M], (double)h / (n * 16.0));
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5)
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
