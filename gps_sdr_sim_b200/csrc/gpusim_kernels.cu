// gpusim_kernels.cu - hand-written sm_100a kernels of the sample-synthesis path.
//
//   k1_chain    one thread per (epoch, channel): exact code-phase checkpoints
//               (replaces the loop-carried dependency of gpssim.c:2212-2218)
//   k2_synth    one thread per chunk of consecutive samples: carrier table x C/A chip x
//               data bit x gain summed over channels, rounded and packed
//               (replaces gpssim.c:2192-2263 and the formatter at :2266-2288)
//   k2_generic  same result, no range assumptions, scalar stores
//
// The arithmetic lives in gpusim_core.h; this file is the mapping onto threads, shared
// memory and global memory.  Not a contraction: no tensor cores.  The bound is
// instruction issue (INT / FP64 / LDS per sample and channel) below an HBM-write roof.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>

#include "gpusim_kernels.h"

namespace gpusim {

// Function attributes belong to (device, kernel): set them the first time a kernel is launched on a
// device, not on every launch.  One flag word per kernel instantiation, one bit per device.
template <void (*K)(DeviceJob)>
static cudaError_t set_attr_once(cudaFuncAttribute attr, int value)
{
    static std::atomic<unsigned long long> done{0ull};
    int dev = 0;
    cudaError_t err = cudaGetDevice(&dev);
    if (err != cudaSuccess)
        return err;
    const unsigned long long bit = dev < 64 ? 1ull << dev : 0ull;
    if (bit && (done.load(std::memory_order_acquire) & bit))
        return cudaSuccess;
    err = cudaFuncSetAttribute(K, attr, value);
    if (err == cudaSuccess && bit)
        done.fetch_or(bit, std::memory_order_release);
    return err;
}

// ------------------------------------------------------------------------------------
// K0 - navigation data words on the device
// ------------------------------------------------------------------------------------
// one thread per (frame, subframe): 10 words chained through their parity; the six subframes of a frame
// do not depend on each other (nav_subframe)
__global__ void __launch_bounds__(128) k0_navmsg(const NavFrame *frames, int n_frames, uint32_t *dwrd)
{
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int f = gid / kNavSubframes;
    if (f >= n_frames)
        return;
    nav_build_subframe(frames[f], gid - f * kNavSubframes, dwrd + (size_t)f * kNavWords);
}

// one thread per uploaded row: rows that came with (frame, iword, ibit) instead of data bits get their 32
// data bits from the frame's words
__global__ void __launch_bounds__(128) k0_navbits(DevRow *rows, int n_rows, const uint32_t *dwrd)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows)
        return;
    const uint32_t flags = rows[r].flags;
    if (!(flags & kRowNavRef))
        return;
    const uint32_t ref = rows[r].nav_bits;
    rows[r].nav_bits = nav_row_bits(dwrd + (size_t)(ref >> 11) * kNavWords, (int)((ref >> 5) & 63u), (int)(ref & 31u));
    rows[r].flags = (uint16_t)(flags & ~kRowNavRef);
}

// one thread per ephemeris: the five subframes eph2sbf() makes of it (gpssim.c:490-665)
__global__ void __launch_bounds__(64) k0_eph2sbf(const NavEph *eph, int n_eph, NavIono iono, uint32_t *sbf)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_eph)
        nav_eph_subframes(eph[i], iono, sbf + (size_t)i * kNavSbfWords);
}

// k0_navmsg for frames that name their subframes by ephemeris index
__global__ void __launch_bounds__(128) k0_navmsg_ref(const NavFrameRef *frames, int n_frames, const uint32_t *sbf, uint32_t *dwrd)
{
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int f = gid / kNavSubframes;
    if (f >= n_frames)
        return;
    nav_build_subframe_ref(frames[f], sbf, gid - f * kNavSubframes, dwrd + (size_t)f * kNavWords);
}

cudaError_t launch_eph2sbf(const NavEph *eph, int n_eph, const NavIono &iono, uint32_t *sbf, cudaStream_t stream)
{
    if (n_eph <= 0)
        return cudaSuccess;
    k0_eph2sbf<<<(n_eph + 63) / 64, 64, 0, stream>>>(eph, n_eph, iono, sbf);
    return cudaGetLastError();
}

cudaError_t launch_navmsg_ref(const NavFrameRef *frames, int n_frames, const uint32_t *sbf, uint32_t *dwrd, cudaStream_t stream)
{
    if (n_frames <= 0)
        return cudaSuccess;
    const int threads = 128, total = n_frames * kNavSubframes;
    k0_navmsg_ref<<<(total + threads - 1) / threads, threads, 0, stream>>>(frames, n_frames, sbf, dwrd);
    return cudaGetLastError();
}

cudaError_t launch_navmsg(const NavFrame *frames, int n_frames, uint32_t *dwrd, cudaStream_t stream)
{
    if (n_frames <= 0)
        return cudaSuccess;
    const int threads = 128, total = n_frames * kNavSubframes;
    k0_navmsg<<<(total + threads - 1) / threads, threads, 0, stream>>>(frames, n_frames, dwrd);
    return cudaGetLastError();
}

cudaError_t launch_navbits(DevRow *rows, int n_rows, const uint32_t *dwrd, cudaStream_t stream)
{
    if (n_rows <= 0)
        return cudaSuccess;
    const int threads = 128;
    k0_navbits<<<(n_rows + threads - 1) / threads, threads, 0, stream>>>(rows, n_rows, dwrd);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------
// K1
// ------------------------------------------------------------------------------------
// binade constants of one chain (phase_chain_tab) in shared memory: [field][binade][lane], conflict-free
struct ChainTabSmem {
    double *base; // this lane's column
    __device__ __forceinline__ void set(int b, double delta, double lim, double r)
    {
        base[b * 32] = delta;
        base[(kChainTabN + b) * 32] = lim;
        base[(2 * kChainTabN + b) * 32] = r;
    }
    __device__ __forceinline__ double delta(int b) const { return base[b * 32]; }
    __device__ __forceinline__ double lim(int b) const { return base[(kChainTabN + b) * 32]; }
    __device__ __forceinline__ double r(int b) const { return base[(2 * kChainTabN + b) * 32]; }
};

template <bool kReplay>
__global__ void __launch_bounds__(32) k1_chain(DeviceJob job)
{
    // One warp = ONE channel slot over 32 consecutive epochs.  The host re-derives the code phase
    // from the pseudorange every epoch, but the signal is continuous, so the same satellite starts
    // consecutive epochs at almost the same code phase: the 32 chains of a warp cross binades and
    // wrap nearly in lockstep and the walk (a serial, latency-bound loop) hardly diverges.
    __shared__ double tab_s[3 * kChainTabN * 32];
    if (blockIdx.x == 0 && threadIdx.x == 0)
        *job.work_counter = 0; // K2 of the same job runs after this kernel on the same stream
    const int groups = (job.n_epochs + 31) / 32;
    const int blk = blockIdx.x % (groups * kMaxChan);
    const bool carrier = blockIdx.x >= groups * kMaxChan; // FLOAT hosts: second half of the grid
    const int k = blk % kMaxChan;
    const int e = (blk / kMaxChan) * 32 + threadIdx.x;
    const bool has_chain = e < job.n_epochs && k < job.nch[min(e, job.n_epochs - 1)];
    if (!has_chain)
        return;
    ChainTabSmem tab;
    tab.base = tab_s + threadIdx.x;
    const size_t row = (size_t)e * kMaxChan + k;
    if (carrier) {
        // the double carrier phase of a FLOAT_CARR_PHASE host (gpssim.c:2245-2250), scaled by 512
        double *cc = job.ck_c + ck_index(job.ck_e0 + e, k, 0, job.kc);
        auto emit_c = [&](int j, double x, int) { cc[(size_t)j << 5] = x; };
        const int last = ((job.n_samples - 1) / job.chunk) * job.chunk;
        if (kReplay) {
            double x = job.cph0[row];
            for (int n = 0; n <= last; n++) {
                if (n % job.chunk == 0)
                    cc[(size_t)(n / job.chunk) << 5] = x;
                x = carrier_step(x, job.dc[row]);
            }
        } else {
            carrier_chain(job.cph0[row], job.dc[row], last, job.chunk, tab, emit_c);
        }
        return;
    }
    const double d = job.rows[row].d;
    const double x0 = job.x0[row];
    double *cx = job.ck_x + ck_index(job.ck_e0 + e, k, 0, job.kc);
    uint16_t *cw = job.ck_w + ck_index(job.ck_e0 + e, k, 0, job.kc);
    auto emit = [&](int j, double x, int wraps) {
        cx[(size_t)j << 5] = x;
        cw[(size_t)j << 5] = (uint16_t)wraps;
    };
    if (kReplay)
        code_chain_replay(x0, d, job.n_samples, job.chunk, emit);
    else
        code_chain(x0, d, job.n_samples, job.chunk, tab, emit);
}

cudaError_t launch_chain(const DeviceJob &job, ChainAlgo algo, cudaStream_t stream)
{
    // Latency-bound serial chains, few of them: one warp per block spreads them over all SMs.
    const int threads = 32;
    const int blocks = ((job.n_epochs + 31) / 32) * kMaxChan * (job.carrier_float ? 2 : 1);
    if (blocks == 0)
        return cudaSuccess;
    // Same shared-memory carve-out as the synthesis kernel: an SM cannot host two kernels that ask
    // for different L1/shared splits, and this kernel is meant to run beside the previous call's K2.
    const cudaError_t carve =
        algo == ChainAlgo::Replay
            ? set_attr_once<k1_chain<true>>(cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared)
            : set_attr_once<k1_chain<false>>(cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    if (carve != cudaSuccess)
        return carve;
    if (algo == ChainAlgo::Replay)
        k1_chain<true><<<blocks, threads, 0, stream>>>(job);
    else
        k1_chain<false><<<blocks, threads, 0, stream>>>(job);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------
// K2 tuned
// ------------------------------------------------------------------------------------
#ifndef GS_K2_THREADS
#define GS_K2_THREADS 512
#endif
constexpr int kK2Threads = GS_K2_THREADS;

int synth_threads() { return kK2Threads; }

// shared memory: [carrier table 64 KB][negw 33*33 u32, padded][x f64 [C][T]][phs u32 [C][T]][meta u32 [C][T]]
constexpr size_t kSmemLut = (size_t)kLutBytes;
constexpr size_t kSmemNegw = ((size_t)(kCaPrns * kCaWords + 32) * sizeof(uint32_t) + 15) & ~(size_t)15; // + lane table

size_t synth_smem_bytes(int max_active, int threads)
{
    if (max_active < 1)
        max_active = 1;
    return kSmemLut + kSmemNegw + (size_t)max_active * threads * 16;
}

// Shared-memory views of one block.  Per-thread channel state lives in shared memory (it does not
// fit in registers next to 32 accumulators).  It is addressed with 32-bit shared-window addresses so
// that the per-channel address update is a single integer add.  With T threads per block, channel k,
// thread t:
//   integer carrier: k*16T + t*8 : code phase f64 ; + 8T : (carrier phase << 7, meta)
//   double carrier : k*24T + t*8 : code phase f64 ; + 8T : 512*carr_phase f64 ; + 16T : (meta, fp32 bits of dataBit*gain)
// meta = icode | bitk<<8 | (dataBit*gain)<<16.
template <class A>
struct K2Smem {
    typename A::tab_t *lut; // replicated carrier table
    uint32_t *negw;         // inverted C/A chips, [33][35]
    uint32_t negw_s;        // the same as a shared-window address (no generic-to-shared conversion per use)
    uint32_t state;         // shared-window address of this thread's slot for channel 0
};
// FLOAT_CARR_PHASE kernel: 24 B of state per channel and thread.  Two geometries (CF = 1, 2):
//   CF = 2  512 threads, 128 registers, runs of 32 samples - the fast one (6.2 ms on the bench shape), but the
//           state of 512 threads only fits next to the tables for <= 13 active channels (227 KB per block);
//   CF = 1  384 threads (3 warps per SM sub-partition, the most that can have more than 128 registers: the
//           register file is per sub-partition, 16384 / (4 warps x 32 lanes) = 128), runs of 16 samples
//           (7.2 ms) - tables with 14..16 active channels and the 16-sample kernel of low sample rates.
// (Measured on B200: 384 threads with runs of 32: 8.3 ms; 512 threads with runs of 16: 9.1 ms, spills.)
#ifndef GS_K2_THREADS_FLOAT
#define GS_K2_THREADS_FLOAT 384
#endif
constexpr int kK2ThreadsFloat = GS_K2_THREADS_FLOAT;
constexpr int kFloatWideMaxChan = 13;
template <int CF> struct K2Geom {
    static constexpr int kThreads = CF == 1 ? kK2ThreadsFloat : kK2Threads;
    static constexpr uint32_t kStride = (CF ? 24u : 16u) * kThreads; // bytes between channels
    static constexpr uint32_t kSecond = 8u * kThreads;               // (phs, meta) / 512*carr_phase
    static constexpr uint32_t kMeta = CF ? 16u * kThreads : 8u * kThreads + 4u;
};

// Register cap of the synthesis kernel: one block owns the SM (128 registers at 512 threads, 168 at 384).
// (Round 1 also shipped 112-register builds of the 512-thread kernels that left room for the next call's chain
// kernel on the same SM; with the chain kernel at 0.2 ms instead of 0.53 ms they no longer pay - a 4 % slower
// synthesis kernel to hide 5 % - and are gone.)
constexpr int k2_max_regs(int cf)
{
    return cf == 1 ? ((65536 / kK2ThreadsFloat) > 255 ? 255 : (65536 / kK2ThreadsFloat) / 8 * 8) : 128;
}

size_t synth_smem_bytes_float(int max_active)
{
    return kSmemLut + kSmemNegw + (size_t)std::max(1, max_active) * (max_active <= kFloatWideMaxChan ? K2Geom<2>::kStride : K2Geom<1>::kStride);
}

__device__ __forceinline__ double lds_f64(uint32_t a)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint2 lds_u32x2(uint32_t a)
{
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_f64(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v)); }
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v)); }
__device__ __forceinline__ void sts_u32x2(uint32_t a, uint32_t v0, uint32_t v1)
{
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(v0), "r"(v1));
}

__device__ __forceinline__ uint4 ldg_u32x4(const uint4 *p)
{
    uint4 v;
    asm volatile("ld.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
// The row of the next channel through an explicit ld.global: with a plain C++ load ptxas re-derived the shared-window
// base (S2UR + UMOV + ULEA) in every trip of the channel loop and spilled 28 bytes; with this form it does neither.
#ifndef GS_LEAN_ASM_ROW
#define GS_LEAN_ASM_ROW 1
#endif
#if GS_LEAN_ASM_ROW
#define GS_LEAN_ROW_LOAD(p) ldg_u32x4(p)
#else
#define GS_LEAN_ROW_LOAD(p) (*(p))
#endif

// chip_window() through a 32-bit shared-window address: two LDS and one funnel shift
__device__ __forceinline__ uint32_t chip_window_s(uint32_t negw_s, uint32_t prn, int c0)
{
    const uint32_t a = negw_s + (prn * (uint32_t)kCaWords + (uint32_t)(c0 >> 5)) * 4u;
    return funnel_l(lds_u32(a + 4u), lds_u32(a), (uint32_t)c0 & 31u);
}

// SR consecutive samples of one thread, all channels, packed and stored.
// Lanes of a warp vote per channel on whether any of them may reach the 1023-chip wrap inside
// the run; only then the (longer) wrap-aware loop is taken for that channel.
// rows4: this epoch's rows as uint4 pairs; dcs: this epoch's carrier steps (double carrier only);
// cthr_mask: kCthrMask, or 0 to force the wrap-aware loop.
template <class A, int FMT, int SR, int CF>
__device__ __forceinline__ void synth_run(const K2Smem<A> &sm, const uint4 *rows4, const double *dcs, const int nc,
                                          const int ncw, const bool live, const unsigned mask,
                                          const uint32_t lane_off, const uint32_t cthr_mask, uint8_t *dst)
{
    typedef K2Geom<CF> G;
    typename A::acc_t acc[SR];
#pragma unroll
    for (int j = 0; j < SR; j++)
        acc[j] = A::init();

    uint32_t sa = sm.state;
    // The channel loop is software pipelined: the row and the code phase of channel k+1 are loaded
    // while channel k is being generated (each channel's state slot is private to the thread, so the
    // early loads are safe).  Without it every channel begins with a load -> convert -> load chain
    // that only other warps could hide, and there are just four warps per scheduler.  (Prefetching
    // more - carrier phase, chip words - measured slower: the kernel sits at the register limit.)
    uint4 r0n = make_uint4(0, 0, 0, 0); // d, steps, cthr | prn<<10 | woff<<16
    double xn = 0.0, dcn = 0.0; // (dcn: double carrier only - its load is a trip to L2 like the row's)
    if (live && nc > 0) {
        r0n = CF ? rows4[0] : GS_LEAN_ROW_LOAD(rows4); // (double carrier: the explicit ld.global measures 0.6 % slower, 5.421 -> 5.454 ms)
        xn = lds_f64(sa);
        if (CF)
            dcn = dcs[0];
    }
#pragma unroll 1
    for (int k = 0; k < ncw; k++, sa += G::kStride) {
        const bool act = live && k < nc;
        const uint4 r0 = r0n;
        const double x = xn;
        const double dc = dcn;
        if (live && k + 1 < nc) {
            r0n = CF ? rows4[2 * k + 2] : GS_LEAN_ROW_LOAD(rows4 + 2 * k + 2);
            xn = lds_f64(sa + G::kStride);
            if (CF)
                dcn = dcs[k + 1];
        }
        // The chip window (two LDS + funnel shift) and floor(x): the double-carrier kernel is latency bound and
        // gains 3 % from ONE conversion shared by the wrap test and the window and from 32-bit shared
        // addresses (6.17 -> 5.99 ms); the integer-carrier kernel is issue bound and LOSES 1.4 % with that
        // shape (4.17 -> 4.23 ms: the window loads move in front of the branch on the vote), so it keeps the plain form.
        int c0 = 0;
        double magic = 0.0;
        bool wrap = false;
        if (CF) {
            magic = floor_magic(x, c0); // floor(x) and 2^52 - floor(x): two FP64 adds, no F2I / I2F (x = 0 on idle lanes)
            wrap = act && c0 >= (int)(r0.w & cthr_mask);
        } else if (act) {
            wrap = (int)x >= (int)(r0.w & cthr_mask);
        }
        const bool any_wrap = __any_sync(mask, wrap);
        if (act) {
            const double d = __hiloint2double((int)r0.y, (int)r0.x);
            const uint32_t prn = (r0.w >> 10) & 63u;
            const uint32_t win_f = CF ? chip_window_s(sm.negw_s, prn, c0) : 0u;
            const uint32_t *nw = sm.negw + prn * kCaWords;
            if (!CF) {
                ChanState st;
                st.x = x;
                const uint2 pm = lds_u32x2(sa + G::kSecond);
                st.phs = pm.x;
                if (!any_wrap) {
                    synth_fast<A, SR>(acc, st, d, r0.z, meta_sgain(pm.y), chip_window(nw, (int)x), sm.lut, lane_off);
                } else {
                    const uint4 r1 = rows4[2 * k + 1]; // ph0s, gain, nav_bits, icode0 | flags<<16
                    st.icode = meta_icode(pm.y);
                    st.bitk = meta_bitk(pm.y);
                    synth_wrap<A, SR>(acc, st, d, r0.z, (int32_t)r1.y, r1.z, chip_window(nw, (int)x), sm.lut, lane_off);
                    sts_u32(sa + G::kMeta, pack_meta(st.icode, st.bitk, data_sign(r1.z, st.bitk) * (int32_t)r1.y));
                }
                sts_f64(sa, st.x);
                sts_u32(sa + G::kSecond, st.phs);
            } else {
                ChanStateF st;
                st.x = x;
                st.cph = lds_f64(sa + G::kSecond);
                // rising and falling carrier phase have their own loops (one wrap test each)
                if (!any_wrap) {
                    const uint32_t gbits = lds_u32(sa + G::kMeta + 4u); // fp32 bits of dataBit*gain
                    if (dc < 0.0)
                        synth_fast_f<A, SR, true>(acc, st, d, dc, gbits, win_f, magic, sm.lut, lane_off);
                    else
                        synth_fast_f<A, SR, false>(acc, st, d, dc, gbits, win_f, magic, sm.lut, lane_off);
                } else {
                    const uint32_t meta = lds_u32(sa + G::kMeta);
                    const uint4 r1 = rows4[2 * k + 1];
                    st.icode = meta_icode(meta);
                    st.bitk = meta_bitk(meta);
                    if (dc < 0.0)
                        synth_wrap_f<A, SR, true>(acc, st, d, dc, (int32_t)r1.y, r1.z, win_f, sm.lut, lane_off);
                    else
                        synth_wrap_f<A, SR, false>(acc, st, d, dc, (int32_t)r1.y, r1.z, win_f, sm.lut, lane_off);
                    const int sg = data_sign(r1.z, st.bitk) * (int32_t)r1.y;
                    sts_u32x2(sa + G::kMeta, pack_meta(st.icode, st.bitk, sg), A::gain_bits(sg));
                }
                sts_f64(sa, st.x);
                sts_f64(sa + G::kSecond, st.cph);
            }
        }
    }
    if (live)
        store_run<A, FMT, SR>(dst, acc);
}

template <class A, int FMT, int S, int CF>
__global__ void __maxnreg__(k2_max_regs(CF)) k2_synth(DeviceJob job)
{
    typedef K2Geom<CF> G;
    constexpr int T = G::kThreads;
    extern __shared__ __align__(16) unsigned char smem[];
    typedef typename A::tab_t tab_t;
    K2Smem<A> sm;
    sm.lut = reinterpret_cast<tab_t *>(smem);
    sm.negw = reinterpret_cast<uint32_t *>(smem + kSmemLut);
    sm.negw_s = (uint32_t)__cvta_generic_to_shared(sm.negw);
    uint32_t *lane_tab = sm.negw + kCaPrns * kCaWords;

    const int tid = threadIdx.x;
    sm.state = (uint32_t)__cvta_generic_to_shared(smem + kSmemLut + kSmemNegw) + (uint32_t)tid * 8u;
    // replicated carrier table: entry i, replica r at byte i*128 + r*sizeof(tab_t); a lane always
    // reads its own replica, so no lookup ever has a bank conflict
    {
        const tab_t *src = sizeof(tab_t) == 4 ? reinterpret_cast<const tab_t *>(job.lut_wide)
                                              : reinterpret_cast<const tab_t *>(job.lut_f32);
        constexpr int kPerEntry = 128 / (int)sizeof(tab_t);
        for (int i = tid; i < kLutEntries * kPerEntry; i += T)
            sm.lut[i] = src[i / kPerEntry];
    }
    for (int i = tid; i < kCaPrns * kCaWords; i += T)
        sm.negw[i] = job.negw[i];
    if (tid < 32)
        lane_tab[tid] = (uint32_t)(tid & A::kLaneMask) << A::kLaneShift;
    __syncthreads();

    // this lane's replica offset, read back from shared memory so that ptxas cannot see its
    // value range and keeps (x & 0xff80) | lane_off as ONE LOP3
    const uint32_t lane_off = *reinterpret_cast<volatile uint32_t *>(lane_tab + (tid & 31));
    const int lane = tid & 31;
    const uint32_t cthr_mask = job.force_wrap_path ? 0u : kCthrMask;
    constexpr int kBytesPer8 = (FMT == 16) ? 32 : (FMT == 8) ? 16 : 2;

    // Persistent warps: every warp repeatedly claims a unit of 32 chunks, one per lane.  No
    // block-level synchronisation after the tables are built, so a warp in the wrap-aware loop
    // never holds up the others, and the tail of the grid is one unit long.
    //
    // Which 32 chunks form a unit decides how often the wrap-aware loop runs: a run is slow for a
    // channel as soon as ONE lane is near that channel's 1023-chip wrap.  In the aligned layout
    // (ppe > 0) the lanes of a unit are the same slot of 32 consecutive code periods (1 ms of
    // samples apart), so all lanes see every channel at (almost) the same code phase and the wrap
    // hits them in the same run; otherwise a unit is 32 consecutive chunks.
    for (;;) {
        unsigned int unit = 0;
        if (lane == 0)
            unit = atomicAdd(job.work_counter, 1u);
        unit = __shfl_sync(0xffffffffu, unit, 0);
        if (unit >= (unsigned int)job.n_units)
            break;
        int e, jc;
        bool valid;
        if (job.ppe > 0) {
            const unsigned int pg = unit / (unsigned int)job.q;
            const unsigned int slot = unit - pg * (unsigned int)job.q;
            const long long period = (long long)pg * 32 + lane;
            valid = period < (long long)job.n_epochs * job.ppe;
            e = (int)(period / job.ppe);
            jc = (int)(period - (long long)e * job.ppe) * job.q + (int)slot;
        } else {
            const long long gid = (long long)unit * 32 + lane;
            valid = gid < (long long)job.n_epochs * job.kc;
            e = (int)(gid / job.kc);
            jc = (int)(gid - (long long)e * job.kc);
        }
        const unsigned mask = __ballot_sync(0xffffffffu, valid);
        if (valid) {
            const int n0 = jc * job.chunk;
            const int nrun = min(job.chunk, job.n_samples - n0);
            const DevRow *rows = job.rows + (size_t)e * kMaxChan;
            const uint4 *rows4 = reinterpret_cast<const uint4 *>(rows);
            const double *dcs = CF ? job.dc + (size_t)e * kMaxChan : nullptr;
            const int nc = job.nch[e];

            // chunk-start state of every channel
            uint32_t sa = sm.state;
            for (int k = 0; k < nc; k++, sa += G::kStride) {
                const size_t ck = ck_index(job.ck_e0 + e, k, jc, job.kc);
                const DevRow r = rows[k];
                const int ic = (int)r.icode0 + (int)job.ck_w[ck];
                const int bitk = ic / 20;
                const uint32_t meta = pack_meta(ic - bitk * 20, bitk, data_sign(r.nav_bits, bitk) * r.gain);
                sts_f64(sa, job.ck_x[ck]);
                if (!CF) {
                    sts_u32x2(sa + G::kSecond, r.ph0s + (uint32_t)n0 * (uint32_t)r.steps, meta);
                } else {
                    sts_f64(sa + G::kSecond, job.ck_c[ck]);
                    sts_u32x2(sa + G::kMeta, meta, A::gain_bits(data_sign(r.nav_bits, bitk) * r.gain));
                }
            }

            // warp-wide channel and run counts: every participating lane runs the same number of loop
            // iterations and takes part in every vote (the last chunk of an epoch may be shorter,
            // epochs may have different numbers of satellites)
            const int ncw = (int)__reduce_max_sync(mask, (unsigned)nc);
            const int full = nrun / S;
            const int tail8 = (nrun - full * S) / 8;
            const int full_w = (int)__reduce_max_sync(mask, (unsigned)full);
            const int tail_w = (int)__reduce_max_sync(mask, (unsigned)tail8);
            uint8_t *outp = job.out + (size_t)e * job.epoch_bytes + (size_t)(n0 / 8) * kBytesPer8;

            for (int i = 0; i < full_w; i++)
                synth_run<A, FMT, S, CF>(sm, rows4, dcs, nc, ncw, i < full, mask, lane_off, cthr_mask,
                                         outp + (size_t)i * (S / 8) * kBytesPer8);
            for (int i = 0; i < tail_w; i++)
                synth_run<A, FMT, 8, CF>(sm, rows4, dcs, nc, ncw, i < tail8, mask, lane_off, cthr_mask,
                                         outp + ((size_t)full * (S / 8) + i) * kBytesPer8);
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------
// K2 integer carrier ("lean"): the same work decomposition and inner loop as k2_synth, with the
// per-channel-per-run prologue cut down (42 instructions of overhead per 32 samples):
//   * per-thread channel state is ONE 16-byte slot {x f64, carr_phase<<7 | icode0+wraps, fp32 bits of
//     dataBit*gain}: one LDS.128 per run and channel, no unpacking or conversion of the gain;
//   * the chip window is ONE LDS.64 from a table of overlapping 64-bit windows at a byte offset that
//     comes ready-made in the row; the funnel shift takes its count modulo 32 by itself;
//   * floor(x) and the window magic are two FP64 adds (floor_magic), no F2I / I2F;
//   * sample 0 of a run needs no chip advance;
//   * warps whose 32 lanes all work (same channel count, same run count - all but the last unit of
//     an epoch group) run a loop without per-lane activity predicates.
// shared memory: [carrier table 64 KB][chip windows 33*34*8][lane table][state 16 B x channels x threads]
// ------------------------------------------------------------------------------------
constexpr size_t kSmemWin64 = ((size_t)kCaPrns * kCaWinBytes + 32 * sizeof(uint32_t) + 15) & ~(size_t)15;
// Geometry of the lean kernel: threads per block, samples per run, register cap.  512 x 32 x 128 is what ships;
// the macros exist for tools/build_variant.py experiments (e.g. 640 threads x runs of 24 x 96 registers).
#ifndef GS_LEAN_THREADS
#define GS_LEAN_THREADS GS_K2_THREADS
#endif
#ifndef GS_LEAN_S
#define GS_LEAN_S 32
#endif
#ifndef GS_LEAN_REGS
#define GS_LEAN_REGS 128
#endif
constexpr int kLeanThreads = GS_LEAN_THREADS;
constexpr uint32_t kLeanStride = 16u * (uint32_t)kLeanThreads;
#ifndef GS_LEAN_UNROLL
#define GS_LEAN_UNROLL 1
#endif
#ifndef GS_LEAN_PF_STATE
#define GS_LEAN_PF_STATE 1
#endif
constexpr bool kLeanPrefetchState = GS_LEAN_PF_STATE != 0; // also prefetch the next channel's state slot (the row always is)
constexpr int kLeanUnroll = GS_LEAN_UNROLL; // channel-loop unroll factor (2 would drop the software pipeline's register moves)

size_t lean_smem_bytes(int max_active) { return kSmemLut + kSmemWin64 + (size_t)std::max(1, max_active) * kLeanStride; }

__device__ __forceinline__ uint4 lds_u32x4(uint32_t a)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_u32x4(uint32_t a, uint32_t v0, uint32_t v1, uint32_t v2, uint32_t v3)
{
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v0), "r"(v1), "r"(v2), "r"(v3));
}


struct LeanSmem {
    const uint64_t *lut; // replicated carrier table (float2 entries)
    uint32_t win_s;      // shared-window address of the chip-window table
    uint32_t state;      // shared-window address of this thread's slot for channel 0
};

template <int FMT, int SR, bool UNI, bool PF, int NB>
__device__ __forceinline__ void lean_run(const LeanSmem &sm, const uint4 *rows4, const int nc, const int ncw,
                                         const bool live, const unsigned mask, const uint32_t lane_off,
                                         const uint32_t cthr_mask, const double lin_rinv, uint8_t *dst)
{
    typedef AccF32x2 A;
    A::acc_t acc[SR];
#pragma unroll
    for (int j = 0; j < SR; j++)
        acc[j] = A::init();

    uint32_t sa = sm.state;
    // software pipeline over channels: row and state of channel k+1 are in flight while channel k is generated
    uint4 r0n = make_uint4(0, 0, 0, 0), stn = make_uint4(0, 0, 0, 0);
    if (UNI || (live && nc > 0)) {
        r0n = GS_LEAN_ROW_LOAD(rows4);
        if (PF)
            stn = lds_u32x4(sa);
    }
#pragma unroll kLeanUnroll
    for (int k = 0; k < ncw; k++, sa += kLeanStride) {
        const bool act = UNI || (live && k < nc);
        const uint4 r0 = r0n; // d, steps, cthr | prn<<10 | woff<<16
        uint4 st = stn;       // x, phase word, gain bits
        if (!PF && act)
            st = lds_u32x4(sa);
        if (UNI ? (k + 1 < ncw) : (live && k + 1 < nc)) {
            r0n = GS_LEAN_ROW_LOAD(rows4 + 2 * k + 2);
            if (PF)
                stn = lds_u32x4(sa + kLeanStride);
        }
        double x = __hiloint2double((int)st.y, (int)st.x);
        int c0;
        const double magic = floor_magic(x, c0); // idle lanes: x = 0
        const bool wrap = act && c0 >= (int)(r0.w & cthr_mask);
        const bool any_wrap = __any_sync(mask, wrap);
        if (act) {
            const double d = __hiloint2double((int)r0.y, (int)r0.x);
            const uint2 ww = lds_u32x2(sm.win_s + (r0.w >> 16) + (((uint32_t)c0 >> 5) << 3)); // {word i+1, word i}
            const uint32_t win = funnel_l_wrap(ww.x, ww.y, (uint32_t)c0);
            uint32_t phs = st.z;
            // low chip rate (NB > 0): when every lane's run stays inside one binade, the chips of the run come
            // from the exact linear model (synth_lin) instead of two FP64 adds and a shift per sample
            const bool all_lin = NB > 0 && !any_wrap && __all_sync(UNI ? 0xffffffffu : __activemask(), lin_ok(c0, NB));
            if (!any_wrap) {
                if (all_lin)
                    synth_lin<A, SR, (NB > 0 ? NB : 1)>(acc, x, phs, d, r0.z, st.w, win, c0, lin_rinv, sm.lut, lane_off);
                else
                    synth_fast_g<A, SR>(acc, x, phs, d, r0.z, st.w, win, magic, sm.lut, lane_off);
                // (8- and 4-byte stores at a 16-byte lane stride are 2- and 4-way bank conflicts: 6.6 % of the
                // kernel's shared-memory wavefronts, ncu r02.  One 16-byte store of the whole slot has none, but
                // keeps the gain word alive across the run: spills, 4.06 -> 4.15 ms.  The conflicts are cheaper.)
                sts_f64(sa, x);
                sts_u32(sa + 8u, phs);
            } else {
                const uint4 r1 = rows4[2 * k + 1]; // ph0s, gain, nav_bits, icode0 | flags<<16
                ChanState cs;
                cs.x = x;
                cs.phs = phs;
                const int ic = lean_ic(phs);
                cs.bitk = ic / 20;
                cs.icode = ic - cs.bitk * 20;
                synth_wrap<A, SR>(acc, cs, d, r0.z, (int32_t)r1.y, r1.z, win, sm.lut, lane_off);
                const uint32_t gb = A::gain_bits(data_sign(r1.z, cs.bitk) * (int32_t)r1.y);
                sts_u32x4(sa, (uint32_t)__double2loint(cs.x), (uint32_t)__double2hiint(cs.x),
                          lean_phase_word(cs.phs, cs.bitk * 20 + cs.icode), gb);
            }
        }
    }
    if (live)
        store_run<A, FMT, SR>(dst, acc);
}

template <int FMT, int S, int NB>
__global__ void __maxnreg__(GS_LEAN_REGS) k2_lean(DeviceJob job)
{
    typedef AccF32x2 A;
    constexpr int T = kLeanThreads;
    extern __shared__ __align__(16) unsigned char smem[];
    LeanSmem sm;
    uint64_t *lut = reinterpret_cast<uint64_t *>(smem);
    uint32_t *win = reinterpret_cast<uint32_t *>(smem + kSmemLut);
    uint32_t *lane_tab = win + kCaPrns * kCaWin64 * 2;
    sm.lut = lut;
    sm.win_s = (uint32_t)__cvta_generic_to_shared(win);

    const int tid = threadIdx.x;
    sm.state = (uint32_t)__cvta_generic_to_shared(smem + kSmemLut + kSmemWin64) + (uint32_t)tid * 16u;
    // replicated carrier table: entry i, replica r at byte i*128 + r*8 (a lane always reads its own replica)
    for (int i = tid; i < kLutEntries * 16; i += T)
        lut[i] = job.lut_f32[i >> 4];
    // chip windows: entry (prn, i) = {inverted chips word i+1, word i}
    for (int i = tid; i < kCaPrns * kCaWin64; i += T) {
        const int prn = i / kCaWin64, w = i - prn * kCaWin64;
        win[2 * i] = job.negw[prn * kCaWords + w + 1];
        win[2 * i + 1] = job.negw[prn * kCaWords + w];
    }
    if (tid < 32)
        lane_tab[tid] = (uint32_t)(tid & A::kLaneMask) << A::kLaneShift;
    __syncthreads();

    const uint32_t lane_off = *reinterpret_cast<volatile uint32_t *>(lane_tab + (tid & 31)); // opaque to ptxas (see k2_synth)
    const int lane = tid & 31;
    const uint32_t cthr_mask = job.force_wrap_path ? 0u : kCthrMask;
    constexpr int kBytesPer8 = (FMT == 16) ? 32 : (FMT == 8) ? 16 : 2;
    constexpr bool PF = kLeanPrefetchState; // the next channel's state slot is prefetched too (measured: 4.14 -> 4.06 ms)

    for (;;) {
        unsigned int unit = 0;
        if (lane == 0)
            unit = atomicAdd(job.work_counter, 1u);
        unit = __shfl_sync(0xffffffffu, unit, 0);
        if (unit >= (unsigned int)job.n_units)
            break;
        int e, jc;
        bool valid;
        if (job.ppe > 0) {
            const unsigned int pg = unit / (unsigned int)job.q;
            const unsigned int slot = unit - pg * (unsigned int)job.q;
            const long long period = (long long)pg * 32 + lane;
            valid = period < (long long)job.n_epochs * job.ppe;
            e = (int)(period / job.ppe);
            jc = (int)(period - (long long)e * job.ppe) * job.q + (int)slot;
        } else {
            const long long gid = (long long)unit * 32 + lane;
            valid = gid < (long long)job.n_epochs * job.kc;
            e = (int)(gid / job.kc);
            jc = (int)(gid - (long long)e * job.kc);
        }
        const unsigned mask = __ballot_sync(0xffffffffu, valid);
        if (valid) {
            const int n0 = jc * job.chunk;
            const int nrun = min(job.chunk, job.n_samples - n0);
            const DevRow *rows = job.rows + (size_t)e * kMaxChan;
            const uint4 *rows4 = reinterpret_cast<const uint4 *>(rows);
            const int nc = job.nch[e];

            // chunk-start state of every channel
            uint32_t sa = sm.state;
            for (int k = 0; k < nc; k++, sa += kLeanStride) {
                const size_t ck = ck_index(job.ck_e0 + e, k, jc, job.kc);
                const DevRow r = rows[k];
                const int ic = (int)r.icode0 + (int)job.ck_w[ck];
                const double x = job.ck_x[ck];
                sts_u32x4(sa, (uint32_t)__double2loint(x), (uint32_t)__double2hiint(x),
                          lean_phase_word(r.ph0s + (uint32_t)n0 * (uint32_t)r.steps, ic),
                          A::gain_bits(data_sign(r.nav_bits, ic / 20) * r.gain));
            }

            const int ncw = (int)__reduce_max_sync(mask, (unsigned)nc);
            const int full = nrun / S;
            const int tail8 = (nrun - full * S) / 8;
            const int full_w = (int)__reduce_max_sync(mask, (unsigned)full);
            const int tail_w = (int)__reduce_max_sync(mask, (unsigned)tail8);
            uint8_t *outp = job.out + (size_t)e * job.epoch_bytes + (size_t)(n0 / 8) * kBytesPer8;
            // every lane works, on the same number of channels and runs: no activity predicates needed
            const bool uni = mask == 0xffffffffu && __all_sync(mask, nc == ncw && full == full_w);

            if (uni) {
                for (int i = 0; i < full_w; i++)
                    lean_run<FMT, S, true, PF, NB>(sm, rows4, nc, ncw, true, mask, lane_off, cthr_mask, NB > 0 ? job.lin_rinv : 0.0,
                                           outp + (size_t)i * (S / 8) * kBytesPer8);
            } else {
                for (int i = 0; i < full_w; i++)
                    lean_run<FMT, S, false, PF, NB>(sm, rows4, nc, ncw, i < full, mask, lane_off, cthr_mask, NB > 0 ? job.lin_rinv : 0.0,
                                            outp + (size_t)i * (S / 8) * kBytesPer8);
            }
            for (int i = 0; i < tail_w; i++)
                lean_run<FMT, 8, false, PF, NB>(sm, rows4, nc, ncw, i < tail8, mask, lane_off, cthr_mask, NB > 0 ? job.lin_rinv : 0.0,
                                        outp + ((size_t)full * (S / 8) + i) * kBytesPer8);
        }
        __syncwarp();
    }
}

template <int FMT, int S, int NB = 0>
static cudaError_t launch_lean(const DeviceJob &job, cudaStream_t stream)
{
    const size_t smem = lean_smem_bytes(job.max_active);
    cudaError_t err = set_attr_once<k2_lean<FMT, S, NB>>(cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::min<size_t>(lean_smem_bytes(kMaxChan), 232448));
    if (err != cudaSuccess)
        return err;
    const long long warps_per_block = kLeanThreads / 32;
    const int blocks = (int)std::min<long long>(std::max(1, job.sm_count), ((long long)job.n_units + warps_per_block - 1) / warps_per_block);
    k2_lean<FMT, S, NB><<<blocks, kLeanThreads, smem, stream>>>(job);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------
// K2 generic
// ------------------------------------------------------------------------------------
template <int FMT, bool CF>
__global__ void __launch_bounds__(128) k2_generic(DeviceJob job)
{
    const long long total = (long long)job.n_epochs * job.kc;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= total)
        return;
    const int e = (int)(gid / job.kc);
    const int jc = (int)(gid - (long long)e * job.kc);
    const int n0 = jc * job.chunk;
    const int nrun = min(job.chunk, job.n_samples - n0);
    const DevRow *rows = job.rows + (size_t)e * kMaxChan;
    const int nc = job.nch[e];

    GenericChan ch[kMaxChan];
    for (int k = 0; k < nc; k++) {
        const size_t ck = ck_index(job.ck_e0 + e, k, jc, job.kc);
        const DevRow r = rows[k];
        const int ic = (int)r.icode0 + (int)job.ck_w[ck];
        ch[k].x = job.ck_x[ck];
        ch[k].d = r.d;
        ch[k].phs = r.ph0s + (uint32_t)n0 * (uint32_t)r.steps;
        ch[k].steps = r.steps;
        ch[k].cph = CF ? job.ck_c[ck] : 0.0;
        ch[k].dc = CF ? job.dc[(size_t)e * kMaxChan + k] : 0.0;
        ch[k].gain = r.gain;
        ch[k].icode = ic % 20;
        ch[k].bitk = ic / 20;
        ch[k].nav_bits = r.nav_bits;
        ch[k].negw = job.negw + (size_t)row_prn(r) * kCaWords;
    }

    uint8_t *base = job.out + (size_t)e * job.epoch_bytes;
    uint32_t byte = 0;
    for (int n = 0; n < nrun; n++) {
        int i16, q16;
        generic_sample<CF>(ch, nc, job.sin16, job.cos16, i16, q16);
        const int s = n0 + n;
        if (FMT == 16) {
            reinterpret_cast<uint32_t *>(base)[s] = ((uint32_t)i16 & 0xffffu) | ((uint32_t)q16 << 16);
        } else if (FMT == 8) {
            reinterpret_cast<uint16_t *>(base)[s] =
                (uint16_t)(((uint32_t)(i16 >> 4) & 0xffu) | (((uint32_t)(q16 >> 4) & 0xffu) << 8));
        } else {
            byte = (byte << 2) | (i16 > 0 ? 2u : 0u) | (q16 > 0 ? 1u : 0u);
            if ((s & 3) == 3) {
                if ((s >> 2) < job.n_samples / 4) // the reference writes floor(N/4) bytes (gpssim.c:2276)
                    base[s >> 2] = (uint8_t)byte;
                byte = 0;
            }
        }
    }
}

template <class A, int FMT, int S, int CF>
static cudaError_t launch_tuned_a(const DeviceJob &job, cudaStream_t stream)
{
    constexpr int T = K2Geom<CF>::kThreads;
    const size_t smem = kSmemLut + kSmemNegw + (size_t)std::max(1, job.max_active) * K2Geom<CF>::kStride;
    // the most this instantiation can ever ask for (CF = 2 only runs with <= kFloatWideMaxChan channels)
    constexpr size_t smem_max = kSmemLut + kSmemNegw + (size_t)(CF == 2 ? kFloatWideMaxChan : kMaxChan) * K2Geom<CF>::kStride;
    cudaError_t err = set_attr_once<k2_synth<A, FMT, S, CF>>(cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
    if (err != cudaSuccess)
        return err;
    const long long units = job.n_units;
    const long long warps_per_block = T / 32;
    const int blocks = (int)std::min<long long>(std::max(1, job.sm_count), (units + warps_per_block - 1) / warps_per_block);
    k2_synth<A, FMT, S, CF><<<blocks, T, smem, stream>>>(job);
    return cudaGetLastError();
}

template <int FMT, int S>
static cudaError_t launch_tuned(const DeviceJob &job, cudaStream_t stream)
{
    // Double carrier: 512 threads x runs of 32 samples when the per-thread state of all channels fits
    // (<= 13 active channels), else 384 threads x runs of 16 (K2Geom).
    if (job.carrier_float) {
        if (S == 32 && job.max_active <= kFloatWideMaxChan && !job.float_narrow)
            return launch_tuned_a<AccF32x2, FMT, 32, 2>(job, stream);
        return launch_tuned_a<AccF32x2, FMT, 16, 1>(job, stream);
    }
    if (job.accum == 1 && job.lean) {
        if (S == 32 && job.lin_nb == 2)
            return launch_lean<FMT, 32, 2>(job, stream);
        if (S == 32 && job.lin_nb == 4)
            return launch_lean<FMT, 32, 4>(job, stream);
        return launch_lean<FMT, (S == 32 ? GS_LEAN_S : S)>(job, stream);
    }
    return job.accum == 1 ? launch_tuned_a<AccF32x2, FMT, S, 0>(job, stream)
                          : launch_tuned_a<AccWide, FMT, S, 0>(job, stream);
}

template <int FMT>
static cudaError_t launch_generic(const DeviceJob &job, cudaStream_t stream)
{
    const long long total = (long long)job.n_epochs * job.kc;
    const int threads = 128;
    const int blocks = (int)((total + threads - 1) / threads);
    if (job.carrier_float)
        k2_generic<FMT, true><<<blocks, threads, 0, stream>>>(job);
    else
        k2_generic<FMT, false><<<blocks, threads, 0, stream>>>(job);
    return cudaGetLastError();
}

cudaError_t launch_synth(const DeviceJob &job, SynthKernel which, cudaStream_t stream)
{
    if (job.n_epochs == 0)
        return cudaSuccess;
#ifdef GS_ONLY_LEAN8 // development builds (SASS inspection): one instantiation, compiles in seconds
    return launch_lean<8, GS_LEAN_S>(job, stream);
#else
#define GS_DISPATCH(FMTV)                                                        \
    if (which == SynthKernel::Tuned32) return launch_tuned<FMTV, 32>(job, stream); \
    if (which == SynthKernel::Tuned16) return launch_tuned<FMTV, 16>(job, stream); \
    return launch_generic<FMTV>(job, stream);
    if (job.fmt == 16) { GS_DISPATCH(16) }
    if (job.fmt == 8) { GS_DISPATCH(8) }
    GS_DISPATCH(1)
#undef GS_DISPATCH
#endif
}

} // namespace gpusim
