// gpusim_core.h - the per-thread algorithms of the device path.
//
// Everything in this header is plain arithmetic on values passed in; the CUDA
// kernels in gpusim_kernels.cu are thin wrappers that bind these functions to
// threads, shared memory and global memory.  The functions are __host__
// __device__ so that tests/emu/ can run exactly the same code thread by thread
// on the CPU and compare it with the oracle before any GPU time is spent (the
// emulation is test infrastructure; the product only ever launches the kernels).
//
// Reference semantics reproduced here (file:line into /root/reference/gpssim.c):
//   :2199-2209  carrier table index, +-1 * +-1 * table * gain, sum over channels
//   :2212-2238  code_phase += f_code*delt ; wrap at 1023 ; icode/ibit ; data bit
//   :2241       chip = ca[(int)code_phase]
//   :2252       carr_phase += carr_phasestep   (integer-carrier branch)
//   :2258-2263  (acc+64)>>7 -> short I,Q
//   :2266-2288  SC01 / SC08 / SC16 packing
#ifndef GPUSIM_CORE_H
#define GPUSIM_CORE_H

#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define GS_HD __host__ __device__ __forceinline__
#else
#define GS_HD inline
#endif

namespace gpusim {

constexpr int kMaxChan = 16;
constexpr int kCaLen = 1023;
constexpr int kLutEntries = 512;
constexpr int kLutReplicas = 32;     // one copy per lane: every LDS is conflict free
constexpr int kCaWords = 33;         // per PRN, see gpusim_tables.h
constexpr int kCaPrns = 33;          // index by prn 0..32 (0 unused)

// ---- exact IEEE-754 double steps (never contracted into FMA) -----------------------
GS_HD double dadd(double a, double b)
{
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, b);
#else
    volatile double r = a + b;
    return r;
#endif
}
GS_HD double dmul(double a, double b)
{
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    volatile double r = a * b;
    return r;
#endif
}
GS_HD uint64_t dbits(double x)
{
#ifdef __CUDA_ARCH__
    return (uint64_t)__double_as_longlong(x);
#else
    uint64_t u;
    memcpy(&u, &x, 8);
    return u;
#endif
}
GS_HD double dfrombits(uint64_t u)
{
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)u);
#else
    double x;
    memcpy(&x, &u, 8);
    return x;
#endif
}
// floor(x) - c0 for 0 <= x < 2^31, where magic = 2^52 - c0: a round-down add puts
// floor(x)-c0 into the low mantissa word.  One FP64-pipe instruction, no conversion.
GS_HD uint32_t chips_since(double x, double magic)
{
#ifdef __CUDA_ARCH__
    return (uint32_t)__double2loint(__dadd_rd(x, magic));
#else
    int c0 = (int)(4503599627370496.0 - magic);
    return (uint32_t)((int)x - c0);
#endif
}
GS_HD uint32_t funnel_l(uint32_t lo, uint32_t hi, uint32_t sh) // (hi:lo << sh) >> 32, sh in 0..31
{
#ifdef __CUDA_ARCH__
    return __funnelshift_l(lo, hi, sh);
#else
    return sh ? (hi << sh) | (lo >> (32 - sh)) : hi;
#endif
}
GS_HD int64_t mad_wide(int32_t a, int32_t b, int64_t c)
{
#ifdef __CUDA_ARCH__
    long long r;
    asm("mad.wide.s32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"((long long)c));
    return r;
#else
    return (int64_t)a * (int64_t)b + c;
#endif
}

// ---- device-side row: one (epoch, active channel), 32 bytes -------------------------
struct alignas(16) DevRow {
    double d;          // RN(f_code*delt): the addend of gpssim.c:2212
    uint32_t ph0;      // carr_phase at epoch start                     (INT mode)
    int32_t step;      // carr_phasestep, gpssim.c:2176
    int32_t gain;      // gain[i], gpssim.c:2186
    uint32_t nav_bits; // next 32 data bits, MSB first
    uint16_t prn;
    uint16_t icode0;   // chan[i].icode at epoch start
    uint32_t flags;    // bit0: outside the tuned kernel's ranges
};
static_assert(sizeof(DevRow) == 32, "DevRow layout");

constexpr uint32_t kRowNeedsGeneric = 1u;
constexpr int kTunedMaxGain = 255; // 16 ch * 250 * 255 + 64 < 2^20, see kAccBias

// ---- packed accumulator of the tuned kernel -----------------------------------------
// One 32-bit table word T = cos*2^21 + sin; multiplier g = dataBit*gain*2^4.
// T*g = (cos*dataBit*gain)<<25 + (sin*dataBit*gain)<<4 accumulates I and Q in ONE
// 64-bit multiply-add.  Both fields carry a bias of 2^20+64 so they never go negative
// (no borrow between fields) and the +64 of "(acc+64)>>7" is already inside:
//   bits 32..45 of acc = ((i_acc+64)>>7) + 8192,  bits 11..24 = ((q_acc+64)>>7) + 8192.
constexpr int kAccShiftQ = 4;
constexpr int kAccShiftI = 25;
constexpr int64_t kFieldBias = (1 << 20) + 64;
constexpr int64_t kAccBias = (kFieldBias << kAccShiftI) + (kFieldBias << kAccShiftQ);
GS_HD int32_t lut_word(int cosv, int sinv) { return cosv * (1 << 21) + sinv; }
GS_HD int32_t acc_i_biased(int64_t acc) { return (int32_t)(acc >> 32); }                   // I16 + 8192
GS_HD int32_t acc_q_biased(int64_t acc) { return (int32_t)(((uint32_t)acc >> 11) & 0x3fffu); } // Q16 + 8192

// =====================================================================================
// K1 - code-phase chain.  The reference advances code_phase by N rounded double adds per
// epoch and channel (gpssim.c:2212-2218); the samples of an epoch can only be generated
// in parallel if the exact value is known at the start of every thread's chunk.  This
// walks the chain without doing N adds: inside one binade [2^e, 2^(e+1)) below the wrap,
// every "x += d" moves the 53-bit significand of x by the same integer q = RN(d/ulp(x))
// (round-half-even ties: only from an even significand, where the increment is again
// constant), so k steps are one 64-bit multiply-add.  Only steps that cross a binade
// edge or the 1023-chip wrap are executed as real IEEE additions.  ~14 jumps per code
// period instead of ~2600 adds (2.6 MS/s); bit-exact by construction, and checked against
// the plain replay in tests/test_chain.py.
//
// emit(j, x, wraps) is called for sample indices j*every, j = 0 .. ceil(n_total/every)-1.
// =====================================================================================
template <class Emit>
GS_HD void code_chain(double x, const double d, const int n_total, const int every, Emit emit)
{
    const uint64_t db = dbits(d);
    const int ed = (int)((db >> 52) & 0x7ff) - 1023;
    const uint64_t dm = (db & 0xfffffffffffffull) | (1ull << 52);
    const int last = ((n_total - 1) / every) * every;
    int n = 0, next = 0, j = 0, wraps = 0;

    for (;;) {
        if (n == next) {
            emit(j++, x, wraps);
            if (n >= last)
                break;
            next += every;
        }
        const uint64_t xb = dbits(x);
        const int ex = (int)((xb >> 52) & 0x7ff) - 1023;
        const int shift = ex - ed;
        if (shift >= 1 && shift <= 52) {
            uint64_t m = (xb & 0xfffffffffffffull) | (1ull << 52);
            const uint64_t q0 = dm >> shift;
            const uint64_t rem = dm & ((1ull << shift) - 1);
            const uint64_t half = 1ull << (shift - 1);
            uint64_t q = q0 + (rem > half ? 1u : 0u);
            bool ok = true;
            if (rem == half) {      // exact tie: round-half-even depends on the parity of m
                if (m & 1)
                    ok = false;     // one real step makes m even
                else
                    q = q0 + (q0 & 1);
            }
            if (ok) {
                // stay strictly inside the binade and strictly below the 1023 wrap
                const uint64_t lim = (ex == 9) ? ((uint64_t)kCaLen << 43) : (1ull << 53);
                const uint64_t room = lim - 1 - m;
                uint64_t k = room / q;
                const uint64_t to_next = (uint64_t)(next - n);
                if (k > to_next)
                    k = to_next;
                if (k > 0) {
                    m += k * q;
                    x = dfrombits(((uint64_t)(ex + 1023) << 52) | (m & 0xfffffffffffffull));
                    n += (int)k;
                    if (n == next)
                        continue;
                }
            }
        }
        // one genuine step of gpssim.c:2212-2218
        x = dadd(x, d);
        if (x >= (double)kCaLen) {
            x = dadd(x, -(double)kCaLen);
            wraps++;
        }
        n++;
    }
}

// Plain replay of the same chain (N dependent adds); kept as the in-tree cross-check of
// code_chain() on the device ("chain=replay" option) and used by nothing else.
template <class Emit>
GS_HD void code_chain_replay(double x, const double d, const int n_total, const int every, Emit emit)
{
    int wraps = 0, j = 0;
    for (int n = 0; n < n_total; n++) {
        if (n % every == 0)
            emit(j++, x, wraps);
        x = dadd(x, d);
        if (x >= (double)kCaLen) {
            x = dadd(x, -(double)kCaLen);
            wraps++;
        }
    }
}

// ---- data bit of a row after `bitk` bit periods --------------------------------------
GS_HD int data_sign(uint32_t nav_bits, int bitk) // +1 / -1, gpssim.c:2236
{
    return bitk < 32 ? (int)((nav_bits >> (31 - bitk)) & 1u) * 2 - 1 : -1;
}

// =====================================================================================
// K2, tuned inner loops: S consecutive samples of ONE channel added into acc[0..S).
//
//   x      code phase of the first sample (chips)          d      per-sample addend
//   phs    carr_phase << 7 (table index = phs >> 23)        steps  carr_phasestep << 7
//   g      dataBit*gain << kAccShiftQ
//   negw   this PRN's chips, INVERTED (bit set = chip 0 = codeCA -1), 33 words, MSB first
//   lut    this lane's replica of the packed table: entry i at lut[i*kLutReplicas]
//
// The chip sign is folded into the carrier phase: negating (cos,sin) equals adding half
// a cycle, i.e. flipping bit 31 of phs (sin[i^256] == -sin[i] for this table), so a
// sample costs: 2 FP64 adds (advance x; floor(x)-c0 via a round-down magic add), a shift
// of the chip window, one LOP3, one shift, one conflict-free LDS and one 64-bit IMAD.
// =====================================================================================
struct ChanState {
    double x;
    uint32_t phs;
    int32_t icode; // 0..19
    int32_t bitk;  // data bits consumed since the row
};

template <int S>
GS_HD void synth_fast(int64_t (&acc)[S], ChanState &st, const double d, const uint32_t steps,
                      const int32_t g, const uint32_t *negw, const int32_t *lut)
{
    double x = st.x;
    uint32_t phs = st.phs;
    const int c0 = (int)x;
    const uint32_t win = funnel_l(negw[(c0 >> 5) + 1], negw[c0 >> 5], (uint32_t)c0 & 31u);
    const double magic = 4503599627370496.0 - (double)c0;
#pragma unroll
    for (int j = 0; j < S; j++) {
        const uint32_t adv = chips_since(x, magic);
        const uint32_t e = phs ^ ((win << adv) & 0x80000000u);
        acc[j] = mad_wide(lut[(e >> 23) * kLutReplicas], g, acc[j]);
        x = dadd(x, d);
        phs += steps;
    }
    st.x = x;
    st.phs = phs;
}

// Same samples, but the 1023-chip wrap (and with it the icode / data-bit walk of
// gpssim.c:2214-2238) may happen inside the run.
template <int S>
GS_HD void synth_wrap(int64_t (&acc)[S], ChanState &st, const double d, const uint32_t steps,
                      const int32_t gain, const uint32_t nav_bits, const uint32_t *negw,
                      const int32_t *lut)
{
    double x = st.x;
    uint32_t phs = st.phs;
    int c0 = (int)x;
    uint32_t win = funnel_l(negw[(c0 >> 5) + 1], negw[c0 >> 5], (uint32_t)c0 & 31u);
    double magic = 4503599627370496.0 - (double)c0;
    int32_t g = data_sign(nav_bits, st.bitk) * gain * (1 << kAccShiftQ);
#pragma unroll
    for (int j = 0; j < S; j++) {
        const uint32_t adv = chips_since(x, magic);
        const uint32_t e = phs ^ ((win << adv) & 0x80000000u);
        acc[j] = mad_wide(lut[(e >> 23) * kLutReplicas], g, acc[j]);
        x = dadd(x, d);
        phs += steps;
        if (x >= (double)kCaLen) {
            x = dadd(x, -(double)kCaLen);
            if (++st.icode >= 20) {
                st.icode = 0;
                st.bitk++;
                g = data_sign(nav_bits, st.bitk) * gain * (1 << kAccShiftQ);
            }
            win = negw[0];
            magic = 4503599627370496.0;
        }
    }
    st.x = x;
    st.phs = phs;
}

// ---- output packing (gpssim.c:2258-2288) from the biased accumulator -----------------
GS_HD uint32_t pack_sc16(int64_t acc) // little-endian short I, short Q
{
    const uint32_t i16 = (uint32_t)(acc_i_biased(acc) - 8192) & 0xffffu;
    const uint32_t q16 = (uint32_t)(acc_q_biased(acc) - 8192) & 0xffffu;
    return i16 | (q16 << 16);
}
GS_HD uint32_t pack_sc08(int64_t acc) // (signed char)(short>>4) for I then Q, in the low 16 bits
{
    const uint32_t i8 = (uint32_t)((acc_i_biased(acc) >> 4) - 512) & 0xffu;
    const uint32_t q8 = (uint32_t)((acc_q_biased(acc) >> 4) - 512) & 0xffu;
    return i8 | (q8 << 8);
}
GS_HD uint32_t pack_sc01(int64_t acc) // 2 bits: (I>0)<<1 | (Q>0)
{
    return ((acc_i_biased(acc) > 8192) ? 2u : 0u) | ((acc_q_biased(acc) > 8192) ? 1u : 0u);
}

// S consecutive samples of one thread -> their bytes at dst (16-byte aligned for SC16 / SC08,
// 4-byte aligned for SC01).  16-byte vector stores.
GS_HD void store16(uint8_t *dst, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
#ifdef __CUDA_ARCH__
    *reinterpret_cast<uint4 *>(dst) = make_uint4(a, b, c, d);
#else
    const uint32_t w[4] = {a, b, c, d};
    memcpy(dst, w, 16);
#endif
}

template <int FMT, int S>
GS_HD void store_run(uint8_t *dst, const int64_t (&acc)[S])
{
    if (FMT == 16) {
#pragma unroll
        for (int q = 0; q < S / 4; q++)
            store16(dst + 16 * q, pack_sc16(acc[4 * q]), pack_sc16(acc[4 * q + 1]),
                    pack_sc16(acc[4 * q + 2]), pack_sc16(acc[4 * q + 3]));
    } else if (FMT == 8) {
#pragma unroll
        for (int q = 0; q < S / 8; q++) {
            uint32_t w[4];
#pragma unroll
            for (int h = 0; h < 4; h++)
                w[h] = pack_sc08(acc[8 * q + 2 * h]) | (pack_sc08(acc[8 * q + 2 * h + 1]) << 16);
            store16(dst + 16 * q, w[0], w[1], w[2], w[3]);
        }
    } else {
        // byte b holds samples 4b..4b+3 as I0 Q0 I1 Q1 I2 Q2 I3 Q3, MSB first (gpssim.c:2268-2274)
#pragma unroll
        for (int q = 0; q < S / 16; q++) {
            uint32_t w = 0;
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const int s0 = 16 * q + 4 * b;
                const uint32_t byte = (pack_sc01(acc[s0]) << 6) | (pack_sc01(acc[s0 + 1]) << 4) |
                                      (pack_sc01(acc[s0 + 2]) << 2) | pack_sc01(acc[s0 + 3]);
                w |= byte << (8 * b);
            }
#ifdef __CUDA_ARCH__
            *reinterpret_cast<uint32_t *>(dst + 4 * q) = w;
#else
            memcpy(dst + 4 * q, &w, 4);
#endif
        }
    }
}

// =====================================================================================
// Generic exact path: one sample, one channel at a time, any samples_per_epoch, any
// gain, any f_code*delt.  Slow (scalar stores, table reads from global memory); used when
// a table violates the tuned kernel's ranges and as the on-device cross-check of it.
// =====================================================================================
struct GenericChan {
    double x, d;
    uint32_t ph;
    int32_t step, gain, icode, bitk;
    uint32_t nav_bits;
    const uint32_t *negw;
};

GS_HD void generic_sample(GenericChan *ch, int nc, const int16_t *sin512, const int16_t *cos512,
                          int &i16, int &q16)
{
    int i_acc = 0, q_acc = 0;
    for (int k = 0; k < nc; k++) {
        GenericChan &c = ch[k];
        const int chip = (int)c.x;
        const int neg = (int)((c.negw[chip >> 5] >> (31 - (chip & 31))) & 1u);
        const int sgn = (neg ? -1 : 1) * data_sign(c.nav_bits, c.bitk);
        const int it = (int)((c.ph >> 16) & 0x1ffu);
        i_acc += sgn * (int)cos512[it] * c.gain;
        q_acc += sgn * (int)sin512[it] * c.gain;
        c.x = dadd(c.x, c.d);
        if (c.x >= (double)kCaLen) {
            c.x = dadd(c.x, -(double)kCaLen);
            if (++c.icode >= 20) {
                c.icode = 0;
                c.bitk++;
            }
        }
        c.ph += (uint32_t)c.step;
    }
    i16 = (int)(short)((i_acc + 64) >> 7);
    q16 = (int)(short)((q_acc + 64) >> 7);
}

} // namespace gpusim
#endif
