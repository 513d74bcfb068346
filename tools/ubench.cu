// ubench.cu - pipe-rate probes for the instructions the synthesis kernel is made of (B200, sm_100a).
// Not part of the product.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench tools/ubench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 2048
#define UNROLL 16

template <int MODE>
__global__ void __launch_bounds__(512, 1) probe(uint64_t *sink, double dseed, int iseed)
{
    extern __shared__ uint64_t sm[];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = (uint64_t)i * 0x9e3779b97f4a7c15ull;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    double a[UNROLL];
    uint64_t u[UNROLL];
    int32_t w[UNROLL];
    for (int i = 0; i < UNROLL; i++) { a[i] = dseed + i + threadIdx.x; u[i] = iseed + i; w[i] = iseed * (i + 1) + threadIdx.x; }
    const double d = dseed * 0.39;
    const uint64_t gg = ((uint64_t)__float_as_uint(3.0f) << 32) | __float_as_uint(3.0f);
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < UNROLL; i++) {
            if (MODE == 0) a[i] = __dadd_rn(a[i], d);                                  // DADD
            if (MODE == 1) { asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(u[i]) : "l"(gg), "l"(gg)); }  // FFMA2
            if (MODE == 2) { long long r; asm volatile("mad.wide.s32 %0, %1, %2, %3;" : "=l"(r) : "r"(w[i]), "r"(iseed), "l"((long long)u[i])); u[i] = r; } // IMAD.WIDE acc
            if (MODE == 3) { w[i] = (int32_t)((const uint32_t *)sm)[((w[i] >> 7) & 0x3fe0) | lane]; }  // dependent LDS.32 conflict free
            if (MODE == 4) { u[i] = sm[((u[i] >> 9) & 0x1ff0) | (lane & 15)]; }          // dependent LDS.64, 16 replicas
            if (MODE == 5) { w[i] = w[i] * iseed + it; }                                  // IMAD
            if (MODE == 6) { w[i] = (w[i] << 3) ^ (w[i] >> 5); }                          // SHF/LOP
            if (MODE == 7) { asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+r"(w[i]) : "r"(iseed)); } // FFMA
            if (MODE == 8) { w[i] = (int32_t)__umulhi((uint32_t)w[i], 0x10001u + it); }           // IMAD.HI
            if (MODE == 9) { a[i] = __dadd_rd(a[i], d); }                                        // DADD.RM
            if (MODE == 10) { asm volatile("prmt.b32 %0, %0, %1, 0x5432;" : "+r"(w[i]) : "r"(iseed)); } // PRMT
            if (MODE == 11) { long long r; asm volatile("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(w[i]), "r"(it << 9), "l"((long long)gg)); w[i] = (int32_t)(r >> 32); } // IMAD.WIDE, high word used
            if (MODE == 12) { a[i] = __fma_rn(a[i], d, dseed); }                                  // DFMA
            if (MODE == 13) { w[i] = w[i] ^ (it & 0x80000000); }                                  // LOP3
            if (MODE == 14) { w[i] = (uint32_t)w[i] >> 23; }                                        // SHF
            if (MODE == 15) { w[i] = w[i] + iseed; }                                                // IADD
            if (MODE == 17) { long long r; asm volatile("mad.wide.s32 %0, %1, %2, %3;" : "=l"(r) : "r"(iseed), "n"(0x50000), "l"((long long)u[i])); u[i] = r; } // IMAD.WIDE, immediate multiplier
            if (MODE == 18) { long long r; asm volatile("mad.wide.s32 %0, %1, %2, %3;" : "=l"(r) : "r"(w[i]), "n"(0x50000), "l"((long long)gg)); uint32_t a2; asm volatile("lop3.b32 %0, %1, 0xff80, %2, 0xea;" : "=r"(a2) : "r"((uint32_t)(r >> 32)), "r"(lane * 8)); w[i] = (int32_t)a2; } // IMAD.WIDE + LOP3 address pair
            if (MODE == 19) { int32_t v; asm volatile("ld.shared.s8 %0, [%1];" : "=r"(v) : "r"((uint32_t)(w[i] & 0x3ff))); w[i] = v + it; }  // dependent LDS.S8, warp-uniform-ish address
            if (MODE == 20) { u[i] = sm[((u[i] >> 9) & 0x1ff0) | (lane & 15)]; asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(u[(i + 1) % UNROLL]) : "l"(u[i]), "l"(gg)); } // LDS.64 + FFMA2
            if (MODE == 21) { asm volatile("{.reg .b16 h0, h1; mov.b32 {h0, h1}, %1; fma.rn.f32.f16 %0, h0, h1, %0;}" : "+f"(*(float *)&w[i]) : "r"(iseed)); } // mixed f32 += f16*f16
            if (MODE == 16) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x78;" : "+r"(w[i]) : "r"(iseed), "r"(it)); w[i] = w[i] * 128 + iseed; } // LOP3 + IMAD pair (two pipes)
        }
    }
    uint64_t acc = 0;
    for (int i = 0; i < UNROLL; i++) acc += (uint64_t)__double_as_longlong(a[i]) + u[i] + (uint32_t)w[i];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
void run(const char *name, uint64_t *sink, int sms)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    probe<MODE><<<sms, 512, 65536>>>(sink, 1.0, 3);
    cudaEventRecord(e0);
    probe<MODE><<<sms, 512, 65536>>>(sink, 1.0, 3);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)sms * 512 * ITERS * UNROLL;
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%-28s %8.3f ms  %8.1f Gop/s  %6.1f lane-ops/clk/SM @%d MHz nominal  err=%s\n", name, ms, ops / ms / 1e6,
           ops / (ms * 1e-3) / sms / (clk * 1e3), clk / 1000, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint64_t *sink; cudaMalloc(&sink, (size_t)sms * 512 * 8);
    printf("SMs=%d, 512 threads/SM (16 warps), %d-way ILP\n", sms, UNROLL);
    run<0>("DADD", sink, sms);
    run<1>("FFMA2 (f32x2)", sink, sms);
    run<2>("IMAD.WIDE accumulate", sink, sms);
    run<3>("LDS.32 dependent, no conflict", sink, sms);
    run<4>("LDS.64 dependent, 16 replicas", sink, sms);
    run<5>("IMAD", sink, sms);
    run<6>("SHF+SHF+LOP3", sink, sms);
    run<7>("FFMA", sink, sms);
    run<8>("IMAD.HI.U32", sink, sms);
    run<9>("DADD.RM", sink, sms);
    run<10>("PRMT", sink, sms);
    run<11>("IMAD.WIDE (hi word)", sink, sms);
    run<12>("DFMA", sink, sms);
    run<13>("LOP3", sink, sms);
    run<14>("SHF", sink, sms);
    run<15>("IADD", sink, sms);
    run<16>("LOP3+IMAD pair", sink, sms);
    run<17>("IMAD.WIDE imm multiplier", sink, sms);
    run<18>("IMAD.WIDE imm + LOP3 addr", sink, sms);
    run<19>("LDS.S8 dependent", sink, sms);
    run<20>("LDS.64 + FFMA2", sink, sms);
    run<21>("FMA f32 += f16*f16 (mixed)", sink, sms);
    return 0;
}
