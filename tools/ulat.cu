// Dependent-chain latencies of the instructions on the serial paths of the synthesis kernels (sm_100a).
// One warp, one chain, clock64() around 4096 dependent steps.   nvcc -arch=sm_100a -O3 --fmad=false -o ulat ulat.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE> __global__ void chain(double *out, long long *cycles, double x, double d, int n)
{
    __shared__ double tab[8 * 32];
    for (int i = threadIdx.x; i < 8 * 32; i += 32)
        tab[i] = 1e-9 * (double)(i >> 5);
    __syncwarp();
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
            } else if (MODE == 6) {     // DFMA chain
                c = __fma_rn(c, 0.999999, d);
            } else if (MODE == 7) {     // DADD + a DADD that is predicated OFF (predicate known long before): does it cost its latency?
                c = __dadd_rn(c, d);
                asm volatile("{\n\t.reg .pred p;\n\tsetp.lt.s32 p, %1, 0;\n\t@!p bra SKIP7;\n\tadd.rn.f64 %0, %0, 0dC080000000000000;\nSKIP7:\n\t}"
                             : "+d"(c) : "r"(n));
            } else if (MODE == 8) {     // the double-carrier step as shipped: DADD, ISETP on the upper word, predicated DADD (never taken here)
                c = __dadd_rn(c, d);
                asm volatile("{\n\t.reg .pred p;\n\tsetp.lt.s32 p, %1, 0x40800000;\n\t@p bra SKIP8;\n\tadd.rn.f64 %0, %0, 0dC080000000000000;\nSKIP8:\n\t}"
                             : "+d"(c) : "r"(__double2hiint(c)));
            } else if (MODE == 10) {    // threshold form: wrap decided from the value BEFORE the add (DSETP beside the DADD), predicated DADD
                const double c0 = c;
                c = __dadd_rn(c, d);
                asm volatile("{\n\t.reg .pred p;\n\tsetp.lt.f64 p, %1, %2;\n\t@p bra SKIP10;\n\tadd.rn.f64 %0, %0, 0dC080000000000000;\nSKIP10:\n\t}"
                             : "+d"(c) : "d"(c0), "d"(x * 400.0));
            } else if (MODE == 11) {    // threshold form with a 64-bit INTEGER compare of the bit patterns (both values >= 0)
                const long long c0 = __double_as_longlong(c);
                c = __dadd_rn(c, d);
                asm volatile("{\n\t.reg .pred p;\n\tsetp.lt.s64 p, %1, %2;\n\t@p bra SKIP11;\n\tadd.rn.f64 %0, %0, 0dC080000000000000;\nSKIP11:\n\t}"
                             : "+d"(c) : "l"(c0), "l"(__double_as_longlong(x * 400.0)));
            } else if (MODE == 12) {    // threshold form, select: u = c + d, w = u - 512, c = (c0 >= T) ? w : u
                const double c0 = c;
                const double u = __dadd_rn(c, d);
                const double w = __dadd_rn(u, -512.0);
                c = c0 >= x * 400.0 ? w : u;
            } else if (MODE == 13) {    // threshold form, upper words only (exact whenever the upper words differ)
                const int c0 = __double2hiint(c);
                c = __dadd_rn(c, d);
                asm volatile("{\n\t.reg .pred p;\n\tsetp.lt.s32 p, %1, %2;\n\t@p bra SKIP13;\n\tadd.rn.f64 %0, %0, 0dC080000000000000;\nSKIP13:\n\t}"
                             : "+d"(c) : "r"(c0), "r"(__double2hiint(x * 400.0)));
            } else if (MODE == 9) {     // DADD + exponent extract + clamp + shared-memory load of the next addend (the chain kernel's lookup)
                c = __dadd_rn(c, d);
                int bi = (int)((unsigned long long)__double_as_longlong(c) >> 52) & 7;
                c = __dadd_rn(c, tab[bi * 32 + threadIdx.x]);
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
    const char *names[] = {"DADD", "DADD + DADD/DSETP + FSEL (wrap)", "DADD + DADD + ISETP + FSEL", "DADD (+ DADD.RD side)", "IMAD", "FFMA", "DFMA", "DADD + predicated-off DADD", "DADD + ISETP + @P DADD (carrier step)", "DADD + exponent + LDS.64 + DADD", "threshold: DSETP(c0,T) || DADD, @P DADD", "threshold: ISETP.64(c0,T) || DADD, @P DADD", "threshold: DADD, DADD, DSETP(c0,T), FSEL", "threshold: ISETP.32(hi) || DADD, @P DADD"};
#define RUN(M) chain<M><<<1, 32>>>(out, cyc, 1.0, 0.37, n); chain<M><<<1, 32>>>(out, cyc, 1.0, 0.37, n); cudaDeviceSynchronize(); \
    cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost); printf("%-36s %7.2f cycles per step\n", names[M], (double)h / (n * 16.0));
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(10) RUN(11) RUN(12) RUN(13)
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
