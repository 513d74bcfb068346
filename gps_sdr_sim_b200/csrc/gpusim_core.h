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

// Kernel-shape switches, each measured on B200 on the bench workload (tools/variant_bench.py):
//   GS_ADDR_IMAD        table address = (phase >> 23) * 128 + replica base: the shift on the ALU pipe, the
//                       multiply-add on the FMA pipe, instead of shift + LOP3 (both ALU).  4.303 -> 4.240 ms.
//   GS_CHIP_GAIN_INT    integer carrier: the chip sign flips the sign bit of the fp32 gain (the broadcast operand
//                       of FFMA2) instead of the top bit of the carrier phase.  No gain (4.315 ms): off.
//   (double carrier: the chip sign always rides on the gain - the table lookup then depends on the carrier chain
//                       only, one instruction less on the kernel's critical path: 7.74 -> 7.16 ms in round 1.)
#ifndef GS_ADDR_IMAD
#define GS_ADDR_IMAD 1
#endif
#ifndef GS_CHIP_GAIN_INT
#define GS_CHIP_GAIN_INT 0
#endif
//   GS_PACK_FMA         16/8-bit output from the fp32 accumulators: (acc+64)>>7 (and >>4 more for 8-bit) taken by ONE
//                       round-down FFMA2 per sample (I and Q) and the bytes gathered with PRMT - 2 / 2.5 instructions
//                       per sample instead of 4 / 4.5 shifts and masks on the ALU pipe.
#ifndef GS_PACK_FMA
#define GS_PACK_FMA 1
#endif
//   GS_CARRIER_WRAP_PRED  double carrier: the wrap of the carrier phase as a predicated add (see carrier_step_signed)
#ifndef GS_CARRIER_WRAP_PRED
#define GS_CARRIER_WRAP_PRED 1
#endif

namespace gpusim {

constexpr int kMaxChan = 16;
constexpr int kCaLen = 1023;
constexpr int kLutEntries = 512;
constexpr int kLutBytes = 512 * 128;  // replicated carrier table: 512 entries, 128 bytes apart
constexpr int kCaWords = 35;         // per PRN, see gpusim_tables.h
constexpr int kCaPrns = 33;          // index by prn 0..32 (0 unused)
constexpr int kCaWin64 = 34;         // per PRN: 64-bit windows {word i+1, word i}, i = 0..33 (one LDS.64 per chip window)
constexpr int kCaWinBytes = kCaWin64 * 8;
constexpr uint32_t kCthrMask = 0x3ffu;
constexpr double kCarrMod = 512.0;  // FLOAT_CARR_PHASE hosts: the device keeps 512 * carr_phase

// ---- exact IEEE-754 double steps (never contracted into FMA) -----------------------
// Host: the sum / product must be rounded to double on its own, never fused with a neighbouring
// operation.  On x86-64 an empty asm that claims to modify the SSE register is such a barrier without
// the store + load of a volatile (the host's carrier advance spends its time in these).
#if !defined(__CUDA_ARCH__) && defined(__x86_64__) && defined(__GNUC__)
#define GS_ROUND_HERE(r) __asm__ volatile("" : "+x"(r))
#else
#define GS_ROUND_HERE(r) do { volatile double gs_v_ = (r); (r) = gs_v_; } while (0)
#endif
GS_HD double dadd(double a, double b)
{
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, b);
#else
    double r = a + b;
    GS_ROUND_HERE(r);
    return r;
#endif
}
GS_HD double dmul(double a, double b)
{
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    double r = a * b;
    GS_ROUND_HERE(r);
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
// Same with magic = 1.5*2^52 - c0: also valid when x < c0 (after the 1023-chip wrap), the
// result is then floor(x) - c0 as a two's complement 32-bit number.
GS_HD uint32_t chips_since_signed(double x, double magic15)
{
#ifdef __CUDA_ARCH__
    return (uint32_t)__double2loint(__dadd_rd(x, magic15));
#else
    int c0 = (int)(6755399441055744.0 - magic15);
    return (uint32_t)((int)x - c0);
#endif
}
// floor(x) (0 <= x < 2^31) as an integer and the magic of chips_since() for it, 2^52 - floor(x), with two
// FP64 adds and no conversion instruction: 2^52 + floor(x) = RD(x + 2^52), and 2^53 minus that is exact.
GS_HD double floor_magic(double x, int &c0)
{
#ifdef __CUDA_ARCH__
    const double m1 = __dadd_rd(x, 4503599627370496.0);
    c0 = __double2loint(m1);
    return __dadd_rn(9007199254740992.0, -m1);
#else
    c0 = (int)x;
    return 4503599627370496.0 - (double)c0;
#endif
}
// (hi:lo << (sh & 31)) >> 32: the hardware funnel shift takes the count modulo 32 by itself
GS_HD uint32_t funnel_l_wrap(uint32_t lo, uint32_t hi, uint32_t sh)
{
#ifdef __CUDA_ARCH__
    uint32_t r;
    asm("shf.l.wrap.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(lo), "r"(hi), "r"(sh));
    return r;
#else
    sh &= 31u;
    return sh ? (hi << sh) | (lo >> (32 - sh)) : hi;
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
// The first 16 bytes are all the tuned kernel reads per run of samples; the second half is
// needed once per chunk, in the wrap path and by the generic kernel.  The carrier phase is
// kept shifted left by 7 (only bits 16..24 of carr_phase are ever used, gpssim.c:2202), so
// the table index is the top 9 bits.
struct alignas(16) DevRow {
    double d;          // RN(f_code*delt): the addend of gpssim.c:2212
    int32_t steps;     // carr_phasestep << 7 (gpssim.c:2176)
    uint16_t cthr_prn; // bits 0..9: a run starting at floor(code_phase) >= cthr may reach the 1023-chip wrap; bits 10..15: prn
    uint16_t woff;     // byte offset of this PRN's chips in the kernel's 64-bit chip-window table (prn * kCaWinBytes)
    uint32_t ph0s;     // carr_phase at epoch start, << 7
    int32_t gain;      // gain[i], gpssim.c:2186
    uint32_t nav_bits; // next 32 data bits, MSB first
    uint16_t icode0;   // chan[i].icode at epoch start
    uint16_t flags;    // bit0: outside the tuned kernel's ranges
};
static_assert(sizeof(DevRow) == 32, "DevRow layout");

// floor(code_phase) from which a run of up to 32 samples might wrap: 32 adds of d plus a whole
// chip of margin over the rounding of the chain
GS_HD uint16_t wrap_threshold(double d)
{
    const double t = 1022.0 - 33.0 * d;
    return t <= 0.0 ? (uint16_t)0 : (uint16_t)(int)t;
}
GS_HD uint16_t pack_cthr_prn(double d, int prn) { return (uint16_t)(wrap_threshold(d) | ((uint32_t)prn << 10)); }
GS_HD int row_cthr(const DevRow &r) { return (int)(r.cthr_prn & kCthrMask); }
GS_HD int row_prn(const DevRow &r) { return (int)(r.cthr_prn >> 10); }

constexpr uint32_t kRowNeedsGeneric = 1u;
constexpr uint32_t kRowTieInLinRange = 2u; // f_code*delt is an exact half-ulp tie in one of the binades [2^5, 2^10) (synth_lin)
constexpr uint32_t kRowNavRef = 4u;        // nav_bits still holds a reference into the device-built data words (k0_navbits resolves it)
// reference to a position in a device-built frame: frame << 11 | iword << 5 | ibit
constexpr uint32_t kNavRefMaxFrames = 1u << 21;
constexpr int kTunedMaxGain = 255; // 16 ch * 250 * 255 + 64 < 2^20: the accumulator fields below hold it

// ---- accumulator policies of the tuned kernel -----------------------------------------
// Both keep the I and the Q sum of one sample in ONE 64-bit register pair and add a
// channel with ONE instruction; both expose the sums as "biased" integers
//     i_biased = ((i_acc + 64) >> 7) + 8192,   q_biased likewise   (gpssim.c:2258-2259)
// which the packers below turn into SC16 / SC08 / SC01 bytes.  The carrier table is
// replicated in shared memory so that a lookup never has a bank conflict: entries are
// 128 bytes apart and the replica is chosen by the lane, byte offset
//     ((e >> 16) & 0xff80) | lane_off            e = phase with the chip sign folded in.

// (1) 64-bit integer multiply-add on a packed 32-bit table word.
//     T = cos*2^21 + sin, g = dataBit*gain*2^4:  T*g = (cos*dataBit*gain)<<25 + (sin*dataBit*gain)<<4.
//     Both fields carry a bias of 2^20+64: never negative (no borrow between fields), and the
//     +64 of the rounding is already inside.  32 replicas x 4 bytes.
struct AccWide {
    typedef int64_t acc_t;
    typedef int32_t tab_t;
    typedef int32_t gain_t;
    static constexpr int kLaneMask = 31, kLaneShift = 2;
    static constexpr bool kSignInGain = false; // an integer gain cannot be negated by flipping one bit
    static constexpr int kShiftQ = 4, kShiftI = 25;
    static constexpr int64_t kFieldBias = (1 << 20) + 64;
    static GS_HD acc_t init() { return (kFieldBias << kShiftI) + (kFieldBias << kShiftQ); }
    static GS_HD tab_t table_entry(int cosv, int sinv) { return cosv * (1 << 21) + sinv; }
    static GS_HD gain_t make_gain(int signed_gain) { return signed_gain * (1 << kShiftQ); }
    static GS_HD void mad(acc_t &acc, tab_t t, gain_t g) { acc = mad_wide(t, g, acc); }
    static GS_HD void mad_s(acc_t &acc, tab_t t, uint32_t g) { acc = mad_wide(t, (gain_t)g, acc); } // unused (kSignInGain false)
    static GS_HD uint32_t gain_bits(int signed_gain) { return (uint32_t)make_gain(signed_gain); }
    static GS_HD int32_t i_biased(acc_t acc) { return (int32_t)(acc >> 32); }
    static GS_HD int32_t q_biased(acc_t acc) { return (int32_t)(((uint32_t)acc >> 11) & 0x3fffu); }
    static constexpr bool kFloorFma = false;
    template <int K> static GS_HD uint64_t floor_shift(acc_t acc) { return (uint64_t)acc; } // unused
    static GS_HD uint64_t positive_signs(acc_t acc) { return (uint64_t)acc; }               // unused
};

// (2) packed fp32x2 FMA (Blackwell FFMA2) on a float2 table entry (cos, sin).  All values are
//     integers below 2^24, so fp32 arithmetic is exact.  The accumulator starts at
//     1.5*2^23 + 64: its mantissa bits then hold (sum + 64) in two's complement, no conversion
//     instruction needed.  16 replicas x 8 bytes (a 64-bit LDS is served per half warp).
struct AccF32x2 {
    typedef uint64_t acc_t;
    typedef uint64_t tab_t;
    typedef uint64_t gain_t;
    static constexpr int kLaneMask = 15, kLaneShift = 3;
    static constexpr bool kSignInGain = true; // the chip sign can be the sign bit of the fp32 gain
    static constexpr uint32_t kMagicBits = 0x4b400000u; // 12582912.0f
    static GS_HD uint32_t fbits(float f)
    {
#ifdef __CUDA_ARCH__
        return __float_as_uint(f);
#else
        uint32_t u;
        memcpy(&u, &f, 4);
        return u;
#endif
    }
    static GS_HD float ffrom(uint32_t u)
    {
#ifdef __CUDA_ARCH__
        return __uint_as_float(u);
#else
        float f;
        memcpy(&f, &u, 4);
        return f;
#endif
    }
    static GS_HD uint64_t pack(float lo, float hi) { return (uint64_t)fbits(lo) | ((uint64_t)fbits(hi) << 32); }
    static GS_HD acc_t init() { return pack(12582976.0f, 12582976.0f); }
    static GS_HD tab_t table_entry(int cosv, int sinv) { return pack((float)cosv, (float)sinv); }
    static GS_HD gain_t make_gain(int signed_gain) { return pack((float)signed_gain, (float)signed_gain); }
    static GS_HD void mad(acc_t &acc, tab_t t, gain_t g)
    {
#ifdef __CUDA_ARCH__
        asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(t), "l"(g));
#else
        const float lo = ffrom((uint32_t)t) * ffrom((uint32_t)g) + ffrom((uint32_t)acc);
        const float hi = ffrom((uint32_t)(t >> 32)) * ffrom((uint32_t)(g >> 32)) + ffrom((uint32_t)(acc >> 32));
        acc = pack(lo, hi);
#endif
    }
    // same with the gain as ONE fp32 value (bits gbits) for I and Q: the instruction's broadcast form
    static GS_HD void mad_s(acc_t &acc, tab_t t, uint32_t gbits)
    {
#ifdef __CUDA_ARCH__
        asm("{\n\t.reg .b64 gg;\n\tmov.b64 gg, {%2, %2};\n\tfma.rn.f32x2 %0, %1, gg, %0;\n\t}" : "+l"(acc) : "l"(t), "r"(gbits));
#else
        mad(acc, t, (uint64_t)gbits | ((uint64_t)gbits << 32));
#endif
    }
    static GS_HD uint32_t gain_bits(int signed_gain) { return fbits((float)signed_gain); }
    // mantissa bits = sum + 64 (two's complement around the magic); +2^20 keeps the shift unsigned
    static GS_HD int32_t i_biased(acc_t acc) { return (int32_t)(((uint32_t)acc - kMagicBits + (1u << 20)) >> 7); }
    static GS_HD int32_t q_biased(acc_t acc) { return (int32_t)(((uint32_t)(acc >> 32) - kMagicBits + (1u << 20)) >> 7); }
    // floor((sum + 64) / 2^K) for I (low half) and Q (high half), each as a two's complement number in the
    // low mantissa bits of 1.5*2^23: acc * 2^-K + (1.5*2^23 - 1.5*2^23 * 2^-K), rounded down - the product
    // is exact inside the FMA, and one unit of the result binade is 1.  K = 7 or 11.
    static constexpr bool kFloorFma = GS_PACK_FMA != 0;
    template <int K>
    static GS_HD uint64_t floor_shift(acc_t acc)
    {
#ifdef __CUDA_ARCH__
        const uint32_t sc = (uint32_t)(127 - K) << 23;                                   // 2^-K
        const uint32_t m2 = __float_as_uint(12582912.0f - 12582912.0f / (float)(1 << K)); // exact: K <= 11
        uint64_t r;
        asm("{\n\t.reg .b64 ss, mm;\n\tmov.b64 ss, {%2, %2};\n\tmov.b64 mm, {%3, %3};\n\tfma.rm.f32x2 %0, %1, ss, mm;\n\t}"
            : "=l"(r) : "l"(acc), "r"(sc), "r"(m2));
        return r;
#else
        const uint32_t lo = kMagicBits + (uint32_t)((i_biased(acc) >> (K - 7)) - (8192 >> (K - 7)));
        const uint32_t hi = kMagicBits + (uint32_t)((q_biased(acc) >> (K - 7)) - (8192 >> (K - 7)));
        return (uint64_t)lo | ((uint64_t)hi << 32);
#endif
    }
    // bit 31 of each half set iff the 16-bit sample (sum + 64) >> 7 is > 0, i.e. sum + 64 >= 128:
    // (1.5*2^23 + 127) - acc is negative exactly then (an exact difference of integers; zero is +0).
    static GS_HD uint64_t positive_signs(acc_t acc)
    {
#ifdef __CUDA_ARCH__
        uint64_t r;
        asm("{\n\t.reg .b64 ss, mm;\n\tmov.b64 ss, {%2, %2};\n\tmov.b64 mm, {%3, %3};\n\tfma.rn.f32x2 %0, %1, ss, mm;\n\t}"
            : "=l"(r) : "l"(acc), "r"(0xbf800000u), "r"(__float_as_uint(12583039.0f)));
        return r;
#else
        return (i_biased(acc) > 8192 ? 0x80000000ull : 0ull) | (q_biased(acc) > 8192 ? 0x80000000ull << 32 : 0ull);
#endif
    }
};

// =====================================================================================
// K1 - code-phase chain.  The reference advances code_phase by N rounded double adds per
// epoch and channel (gpssim.c:2212-2218); the samples of an epoch can only be generated
// in parallel if the exact value is known at the start of every thread's chunk.  This
// walks the chain without doing N adds: inside one binade [2^e, 2^(e+1)) every "x += d"
// moves x by the same multiple of ulp(x), delta = RN_ulp(x)(d), so k steps are ONE fused
// multiply-add (x + k*delta is representable, hence exact); only the step that crosses a
// binade edge or the wrap is executed as a real IEEE addition.  ~10 jumps per code period
// instead of ~2600 adds (2.6 MS/s); bit-exact by construction, and checked against the
// plain replay in tests/test_chain.py.
//
// emit(j, x, wraps) is called for sample indices j*every, j = 0 .. ceil(n_total/every)-1.
// =====================================================================================
GS_HD double dfma(double a, double b, double c) // one rounding of the exact a*b+c
{
#ifdef __CUDA_ARCH__
    return __fma_rn(a, b, c);
#else
    return __builtin_fma(a, b, c);
#endif
}
GS_HD double dadd_down(double a, double b) // a + b rounded towards -infinity (exact sums are not affected)
{
#ifdef __CUDA_ARCH__
    return __dadd_rd(a, b);
#else
    // host: only ever called with a in [0, 2^31] and b = 2^52, where the rounded-down sum is 2^52 + floor(a)
    return b + __builtin_floor(a);
#endif
}
GS_HD int ctz64(uint64_t v) // v != 0
{
#ifdef __CUDA_ARCH__
    return __ffsll((long long)v) - 1;
#else
    return __builtin_ctzll(v);
#endif
}
// unbiased exponent e of the one binade [2^e, 2^(e+1)) in which the step d > 0 is an exact half-ulp tie
// (x + d halfway between two doubles for every x of that binade), or a large negative number if there is none
GS_HD int chain_tie_binade(double d)
{
    const uint64_t db = dbits(d);
    const uint64_t dm = db & 0xfffffffffffffull;
    return dm ? (int)(db >> 52) + ctz64(dm) + 1 - 1023 : -100000;
}
GS_HD bool warp_any(unsigned mask, bool pred) // kept for callers that vote; on the host a "warp" is one thread
{
#ifdef __CUDA_ARCH__
    return __any_sync(mask, pred);
#else
    (void)mask;
    return pred;
#endif
}
#if defined(__GNUC__) || defined(__CUDACC__)
#define GS_UNLIKELY(c) __builtin_expect(!!(c), 0)
#else
#define GS_UNLIKELY(c) (c)
#endif

// Exact walk of the recurrence
//     x += d;  if (x >= M) x -= M;  else if (x < 0) x += M;          (|d| < M, either sign)
// for n_end steps; emit(j, x, wraps) at step indices j*every (j >= 0, j*every <= n_end); returns
// the value after n_end steps.  Two users:
//   code phase    M = 1023, d = RN(f_code*delt) > 0                    gpssim.c:2212-2218
//   carrier phase M = 512,  d = 512*RN(f_carr*delt), x = 512*carr_phase (FLOAT_CARR_PHASE hosts;
//                 the power-of-two scaling commutes with every rounding)   gpssim.c:2245-2250
// kSign: +1 the caller knows d >= 0 (the code phase), -1 d < 0, 0 decided at run time.
//
// One trip of the loop = at most one jump inside the current binade followed by ONE real step (which
// is the one that crosses the binade edge or wraps).  Why the jump is exact, with u = ulp(x),
// x in [c, 2c), c = 2^e:
//   * every sum x + d whose real value stays inside [c, 2c) is rounded on the grid of u, to
//     x + delta with delta = RN_u(d); ties go to the even significand, and delta computed as
//     (c + d) - c (rising; (1.5c + d) - 1.5c falling: an anchor with an even significand that keeps
//     anchor + d inside the binade) rounds them exactly like a sum from an even significand does.  The
//     one case that rounds the other way - an exact tie from an ODD significand - takes a real step
//     first; after it the significand is even for good (tie_s: the only binade distance at which d
//     is a tie is 1 + the number of trailing zero bits of its significand).
//   * k steps are then x + k*delta, ONE fma: exact because the result is a multiple of u inside the
//     binade.  The landing must stay STRICTLY inside: at most pred(min(2c, M)) rising, at least
//     succ(c) falling (on the edge c itself the grid below is twice as fine and the reference's sum
//     may round to c - u/2).
//   * k only has to be a LOWER bound of the number of steps that fit: floor(room * r) with
//     r = (1/|d|) * (1 - 2^(s-50)), s = e - exponent(d).  |delta| differs from |d| by at most u/2, a
//     relative 2^(s-53), and the roundings of 1/|d| and of the product by 2^-52 each, so room * r <
//     room / |delta| and the floor never overshoots: no verification, no fix-up.  When it falls
//     short by a step (probability ~2^(2s-50) per jump) that step is simply taken as a real addition.
//     The floor is the low word of RD(room * r + 2^52): no conversion instructions on the chain.
// The constants of one binade [2^e, 2^(e+1)) for a given step d: everything in the jump that does not
// depend on x.  The walk visits the binades in order (a rising chain leaves binade e into e+1), so the
// constants of the NEXT binade are computed while the current trip's dependent chain (subtract, multiply,
// floor, fused multiply-add, real add: ~6 FP64 latencies) is in flight - the walk is latency bound, its
// issue slots are idle.
struct BinadeConsts {
    int bex;      // biased exponent these constants belong to
    bool ok;      // a jump is possible in this binade (|d| < c/2, k < 2^31, d != 0)
    double delta; // RN_ulp(d): what one step adds (0 when !ok)
    double lim;   // furthest landing: pred(min(2c, M)) rising, succ(c) falling
    double r;     // (1/|d|) * (1 - 2^(s-50))
};
template <int kSign>
GS_HD BinadeConsts binade_consts(const int bex, const double d, const int bex_d, const uint64_t mbits, const double rd)
{
    BinadeConsts k;
    const int shift = bex - bex_d;
    k.bex = bex;
    k.ok = shift >= 2 && shift <= 30 && rd != 0.0;
    k.delta = 0.0;
    k.lim = 0.0;
    k.r = 0.0;
    if (k.ok) {
        const uint64_t cb = (uint64_t)bex << 52; // c = 2^e
        if (kSign > 0) {
            const double c = dfrombits(cb);
            k.delta = dadd(dadd(c, d), -c);
            const uint64_t c2b = cb + (1ull << 52);                // 2c
            k.lim = dfrombits((c2b > mbits ? mbits : c2b) - 1);   // pred(min(2c, M))
        } else {
            const double c15 = dfrombits(cb | (1ull << 51)); // 1.5c
            k.delta = dadd(dadd(c15, d), -c15);              // negative
            k.lim = dfrombits(cb + 1);                       // succ(c)
        }
        k.r = dfma(-rd, dfrombits((uint64_t)(1023 - 50 + shift) << 52), rd);
    }
    return k;
}

template <int kSign = 0, class Emit>
GS_HD double phase_chain(double x, const double d, const double M, const int n_end, const int every,
                         Emit emit, const unsigned mask = 0xffffffffu)
{
    (void)mask;
    if (kSign == 0)
        return d < 0.0 ? phase_chain<-1>(x, d, M, n_end, every, emit) : phase_chain<1>(x, d, M, n_end, every, emit);
    constexpr bool neg = kSign < 0;
    constexpr int kS = kSign == 0 ? 1 : kSign;
    const uint64_t db = dbits(neg ? -d : d);
    const int bex_d = (int)(db >> 52); // biased exponent of |d| (d is finite and normal or zero)
    const uint64_t dm = db & 0xfffffffffffffull;
    const int tie_bex = dm ? bex_d + ctz64(dm) + 1 : -1; // the one binade in which d is an exact half-ulp tie
    const double rd = db ? 1.0 / (neg ? -d : d) : 0.0;   // d == 0: no jumps, every step is a real "x += 0"
    const uint64_t mbits = dbits(M);
    int n = 0, next = every, j = 1, wraps = 0;

    emit(0, x, 0);
    BinadeConsts cur = binade_consts<kS>((int)(dbits(x) >> 52), d, bex_d, mbits, rd);
    while (n < n_end) {
        const uint64_t xb = dbits(x);
        const int bex = (int)(xb >> 52); // x is in [0, M): sign bit clear
        if (GS_UNLIKELY(bex != cur.bex)) // after the wrap, in the binades too small to jump in, or when a jump fell short
            cur = binade_consts<kS>(bex, d, bex_d, mbits, rd);
        // the binade the real step of this trip will (almost always) land in
        const BinadeConsts nxt = binade_consts<kS>(neg ? bex - 1 : bex + 1, d, bex_d, mbits, rd);
        // jump inside this binade; an exact tie from an odd significand takes its real step first
        const bool go = cur.ok && !(bex == tie_bex && (xb & 1u));
        const double room = neg ? dadd(x, -cur.lim) : dadd(cur.lim, -x); // falling: -ulp when x == c
        const double m1 = dadd_down(dmul(room, cur.r), 4503599627370496.0); // 2^52 + floor(room * r)
        int k = (int)(uint32_t)dbits(m1);
        double kf = dadd(m1, -4503599627370496.0);
        if (!go || k < 0) {
            k = 0;
            kf = 0.0;
        }
        if (GS_UNLIKELY(k > n_end - n)) {
            k = n_end - n;
            kf = (double)k;
        }
        // checkpoints that fall inside the jump
        while (GS_UNLIKELY(k > 0 && next - n <= k)) {
            emit(j++, dfma((double)(next - n), cur.delta, x), wraps);
            next += every;
        }
        x = dfma(kf, cur.delta, x); // k == 0: x + 0*delta = x
        n += k;
        if (GS_UNLIKELY(n >= n_end))
            break;
        // one genuine step: crosses the binade edge, wraps, or walks the binades too small to jump in
        x = dadd(x, d);
        if (!neg) { // a rising chain leaves [0, M) at the top only, a falling one at the bottom
            if (x >= M) {
                x = dadd(x, -M);
                wraps++;
            }
        } else if (x < 0.0) {
            x = dadd(x, M);
            wraps++;
        }
        n++;
        if (GS_UNLIKELY(n == next)) {
            emit(j++, x, wraps);
            next += every;
        }
        cur = nxt;
    }
    return x;
}

// ---- the same walk with the binade constants tabulated up front ------------------------------------
// phase_chain() derives the constants of a binade whenever the chain enters it - on a GPU that is most of
// the instructions of a trip (ncu, round 2: ~130 SASS instructions per trip, 6 cycles each on a lone
// in-order warp).  (On the host's out-of-order cores phase_chain() stays the faster walk - the constants of
// the next binade are computed in the shadow of the current jump, where this one has a table lookup on its
// serial path: 31 against 44 us per epoch of a 1.5 kHz carrier - so gpusim_advance_carrier_f64 keeps it.)  The constants depend on d only, and a chain below M <= 1024 only ever visits the binades
// [2^-4, 2^10): phase_chain_tab() tabulates them once per chain (14 entries x {delta, lim, r}, shared
// memory on the device, the stack on the host) and a trip is then: three loads, lim - x, one round-down
// fused multiply-add (the floor of the jump count), the jump itself (one FMA) and the real step - the
// arithmetic, and with it every value, is the one phase_chain() documents.  Entry 0 stands for everything
// below 2^-3 and never jumps (r = 0 gives k = 0): those few steps are real additions.
constexpr int kChainTabLo = 1023 - 4; // biased exponent of table entry 0
constexpr int kChainTabN = 14;        // entries 1..13: binades 2^-3 .. 2^9
struct ChainTabHost {                 // host / emulation storage; the device uses shared memory (gpusim_kernels.cu)
    double v[3][kChainTabN];
    GS_HD void set(int b, double delta, double lim, double r)
    {
        v[0][b] = delta;
        v[1][b] = lim;
        v[2][b] = r;
    }
    GS_HD double delta(int b) const { return v[0][b]; }
    GS_HD double lim(int b) const { return v[1][b]; }
    GS_HD double r(int b) const { return v[2][b]; }
};
GS_HD double dfma_down(double a, double b, double c) // 2^52 + floor(a * b), for c = 2^52 and 0 <= |a * b| < 2^31
{
#ifdef __CUDA_ARCH__
    return __fma_rd(a, b, c); // the exact product, one rounding (towards -infinity)
#else
    // host: the product rounded to nearest first.  Both are lower bounds of the number of steps that fit (the
    // margin in r covers a rounding of the product, see phase_chain), and any count up to the bound gives the
    // same chain - the host and the device may split a binade into different jumps, never into different values.
    return c + __builtin_floor(a * b);
#endif
}

GS_HD double int_as_double_exact(uint32_t k) // (double)k for 0 <= k < 2^32 without a conversion instruction
{
    return dadd(dfrombits(0x4330000000000000ull | (uint64_t)k), -4503599627370496.0);
}

template <int kSign, class Tab, class Emit>
GS_HD double phase_chain_tab(double x, const double d, const double M, const int n_end, const int every, Tab &tab,
                             Emit emit)
{
    static_assert(kSign == 1 || kSign == -1, "the caller resolves the sign of d");
    constexpr bool neg = kSign < 0;
    const uint64_t db = dbits(neg ? -d : d);
    const int bex_d = (int)(db >> 52);
    const uint64_t dm = db & 0xfffffffffffffull;
    const int tie_bi = dm ? bex_d + ctz64(dm) + 1 - kChainTabLo : -1; // table index of the one binade where d is a tie
    const double rd = db ? 1.0 / (neg ? -d : d) : 0.0;
    const uint64_t mbits = dbits(M);
    tab.set(0, 0.0, 0.0, 0.0);
    for (int b = 1; b < kChainTabN; b++) {
        const BinadeConsts k = binade_consts<kSign>(kChainTabLo + b, d, bex_d, mbits, rd);
        tab.set(b, k.delta, k.lim, k.ok ? k.r : 0.0);
    }
    // The walk is driven by the checkpoints, not the other way round.  State: x = the value after n steps and
    // the SEGMENT that starts there - k more steps of delta each that stay inside x's binade.  A checkpoint
    // inside the segment is x + (target - n) * delta (any count up to k is as exact as k itself); only a target
    // beyond it makes the chain take the segment, one genuine step (the one that crosses the binade edge, wraps,
    // or walks the untabulated binades) and open the next segment.  On the device every lane of a warp thus
    // emits checkpoint j at the same point of the program - one coalesced store per warp - whatever the number
    // of trips each lane needed to get there, and no trip is spent on stopping at a checkpoint.  The trip is
    // branch-free on purpose: the binades too small to jump in go through the same instructions with r = 0 (a
    // shortcut for them - no lookup, no floor - makes the lanes of a warp take different paths: K1 0.21 -> 0.24 ms,
    // 0.66 -> 0.80 ms with carrier chains, measured).
    int n = 0, wraps = 0, k = 0;
    double delta = 0.0, kf = 0.0;
    auto open_segment = [&]() {
        const uint64_t xb = dbits(x);
        int bi = (int)(xb >> 52) - kChainTabLo; // x in [0, M), M <= 1024: bi <= 13
        bi = bi < 0 ? 0 : bi;
        // no jump (k = 0) is "r = 0" or "room = 0" BEFORE the multiply, so that k and (double)k both come straight
        // out of the one round-down multiply-add: the untabulated binades have r = 0 in the table; an exact tie
        // from an odd significand takes its real step first; a falling chain sitting on the binade's lower edge
        // has room = -ulp
        double r = tab.r(bi);
        if (bi == tie_bi && (xb & 1u))
            r = 0.0;
        double room = neg ? dadd(x, -tab.lim(bi)) : dadd(tab.lim(bi), -x);
        if (neg && room < 0.0)
            room = 0.0;
        const double m1 = dfma_down(room, r, 4503599627370496.0); // 2^52 + floor(room * r)
        k = (int)(uint32_t)dbits(m1);
        kf = dadd(m1, -4503599627370496.0);
        delta = tab.delta(bi);
    };
    auto value_at = [&](int target) { // target - n in [0, k]
        return dfma(int_as_double_exact((uint32_t)(target - n)), delta, x);
    };
    open_segment();
    int j = 0;
    for (int target = 0;;) {
        while (target - n > k) {
            x = dfma(kf, delta, x);
            x = dadd(x, d);
            if (!neg) {
                if (x >= M) {
                    x = dadd(x, -M);
                    wraps++;
                }
            } else if (x < 0.0) {
                x = dadd(x, M);
                wraps++;
            }
            n += k + 1;
            open_segment();
        }
        if (target == n_end && j > 0 && (n_end % every) != 0)
            break; // the final advance to n_end is not a checkpoint
        emit(j++, value_at(target), wraps);
        if (target == n_end)
            break;
        target = (n_end - target < every) ? n_end : target + every;
    }
    return value_at(n_end);
}

// K1's use: code-phase checkpoints at sample indices j*every, j = 0 .. ceil(n_total/every)-1.
template <class Tab, class Emit>
GS_HD void code_chain(double x, const double d, const int n_total, const int every, Tab &tab, Emit emit)
{
    const int last = ((n_total - 1) / every) * every; // sample index of the last checkpoint
    phase_chain_tab<1>(x, d, (double)kCaLen, last, every, tab, emit); // f_code*delt > 0
}
template <class Emit>
GS_HD void code_chain(double x, const double d, const int n_total, const int every, Emit emit)
{
    ChainTabHost tab;
    code_chain(x, d, n_total, every, tab, emit);
}
// carrier chain of a FLOAT_CARR_PHASE host (512 * carr_phase, either sign of the step): checkpoints as above,
// returns the phase after n_end steps
template <class Tab, class Emit>
GS_HD double carrier_chain(double x, const double dc, const int n_end, const int every, Tab &tab, Emit emit)
{
    return dc < 0.0 ? phase_chain_tab<-1>(x, dc, kCarrMod, n_end, every, tab, emit)
                    : phase_chain_tab<1>(x, dc, kCarrMod, n_end, every, tab, emit);
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
//   win    chip_window(): this PRN's 32 chips from floor(x) on, INVERTED (bit set = chip 0 =
//          codeCA -1), current chip in bit 31
//   lut    the replicated carrier table in shared memory, lane_off this lane's replica
//
// The chip sign is folded into the carrier phase: negating (cos,sin) equals adding half
// a cycle, i.e. flipping bit 31 of phs (sin[i^256] == -sin[i] for this table).  A sample
// of a channel then costs 2 FP64 adds (advance x; floor(x)-c0 through a round-down magic
// add), a shift of the 32-chip window, one LOP3 (xor), one shift, one LOP3 (mask|lane),
// one conflict-free LDS, one multiply-add for I and Q together, one add for the carrier.
// =====================================================================================
template <class A>
GS_HD typename A::tab_t lut_at(const typename A::tab_t *lut, uint32_t e, uint32_t lane_off)
{
#if defined(__CUDA_ARCH__) && GS_ADDR_IMAD
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(lut) + lane_off; // loop invariant
    uint32_t addr;
    asm("mad.lo.u32 %0, %1, 128, %2;" : "=r"(addr) : "r"(e >> 23), "r"(base));
    typename A::tab_t v;
    if (sizeof(typename A::tab_t) == 8)
        asm volatile("ld.shared.b64 %0, [%1];" : "=l"(*reinterpret_cast<uint64_t *>(&v)) : "r"(addr));
    else
        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(*reinterpret_cast<uint32_t *>(&v)) : "r"(addr));
    return v;
#else
    const uint32_t off = ((e >> 16) & 0xff80u) | lane_off; // (table index << 7) | replica, bytes
    return *reinterpret_cast<const typename A::tab_t *>(reinterpret_cast<const char *>(lut) + off);
#endif
}

// the 32 (inverted) chips from chip c0 on, chip c0 in bit 31; the table continues past chip 1022
GS_HD uint32_t chip_window(const uint32_t *negw, int c0)
{
    return funnel_l(negw[(c0 >> 5) + 1], negw[c0 >> 5], (uint32_t)c0 & 31u);
}

struct ChanState {
    double x;
    uint32_t phs;
    int32_t icode; // 0..19
    int32_t bitk;  // data bits consumed since the row
};
// icode | bitk << 8 | (dataBit * gain) << 16: the per-thread, per-channel word kept next to x and phs
GS_HD uint32_t pack_meta(int icode, int bitk, int signed_gain)
{
    return (uint32_t)icode | ((uint32_t)bitk << 8) | ((uint32_t)signed_gain << 16);
}
GS_HD int meta_icode(uint32_t m) { return (int)(m & 0xffu); }
GS_HD int meta_bitk(uint32_t m) { return (int)((m >> 8) & 0xffu); }
GS_HD int meta_sgain(uint32_t m) { return (int)(int32_t)m >> 16; }

template <class A, int S>
GS_HD void synth_fast(typename A::acc_t (&acc)[S], ChanState &st, const double d, const uint32_t steps,
                      const int signed_gain, const uint32_t win, const typename A::tab_t *lut,
                      const uint32_t lane_off)
{
    double x = st.x;
    uint32_t phs = st.phs;
    const int c0 = (int)x;
    const double magic = 4503599627370496.0 - (double)c0;
    const typename A::gain_t g = A::make_gain(signed_gain);
    const uint32_t gb = A::gain_bits(signed_gain);
#pragma unroll
    for (int j = 0; j < S; j++) {
        const uint32_t adv = chips_since(x, magic);
        const uint32_t chip = (win << adv) & 0x80000000u;
        if (A::kSignInGain && GS_CHIP_GAIN_INT)
            A::mad_s(acc[j], lut_at<A>(lut, phs, lane_off), gb ^ chip);
        else
            A::mad(acc[j], lut_at<A>(lut, phs ^ chip, lane_off), g);
        x = dadd(x, d);
        phs += steps;
    }
    st.x = x;
    st.phs = phs;
}

// The integer-carrier kernel's form of synth_fast: the signed gain arrives as fp32 bits (the broadcast
// operand of FFMA2, kept in the thread's state so that no conversion is needed per run), the window
// magic is computed by the caller together with floor(x) (floor_magic), sample 0 needs no chip advance,
// and the low 7 bits of phs may carry the caller's bookkeeping (only bits 23..31 are ever looked at and
// steps has its low 7 bits clear, gpssim.c:2202).
template <class A, int S>
GS_HD void synth_fast_g(typename A::acc_t (&acc)[S], double &x, uint32_t &phs, const double d, const uint32_t steps,
                        const uint32_t gbits, const uint32_t win, const double magic, const typename A::tab_t *lut,
                        const uint32_t lane_off)
{
#pragma unroll
    for (int j = 0; j < S; j++) {
        const uint32_t t = j == 0 ? win : win << chips_since(x, magic);
        A::mad_s(acc[j], lut_at<A>(lut, phs ^ (t & 0x80000000u), lane_off), gbits);
        x = dadd(x, d);
        phs += steps;
    }
}

// ---- low chip rates: the chips of a whole run from ONE exact linear model --------------------------------
// At >= ~8 samples per chip (config 5: 20 MS/s, 19.5 samples per chip) a run of S samples sees at most NB
// chip boundaries, yet synth_fast_g pays two FP64 adds, a shift and a LOP3 per sample to find the chip.
// Inside one binade [2^e, 2^(e+1)) the reference's chain is exactly linear, x_j = x_0 + j*delta with
// delta = RN_ulp(x)(d) (see phase_chain: every sum is a multiple of ulp(x), hence exact - except in the one
// binade where d is an exact half-ulp tie, which the caller excludes), so the sample at which chip c0+b
// begins is n_b = min{ j : x_0 + j*delta >= c0 + b }: a quotient estimate from a job-wide 1/d (good to 1e-3
// samples) corrected by two exact FMA tests.  The NB boundaries give a 32-bit sign mask for the run; a
// sample then costs a select of +-gain (predicates straight from the mask: R2P), the carrier phase add, the
// table address (shift + multiply-add), one LDS and one FFMA2 - 6.25 instead of 9 instructions, and the new
// code phase is x_0 + S*delta, one more FMA.
//
// Preconditions (lin_ok(), voted per warp by the caller; otherwise synth_fast_g / synth_wrap run):
//   * no 1023-chip wrap in the run (the caller's threshold test), c0 = floor(x_0) >= kLinMinChip,
//   * no power of two in (c0, c0 + NB + 1]: x_0 .. x_0 + S*delta stays inside x_0's binade,
//   * S * d < NB, and the row's tie binade is not one of [2^5, 2^10) (upload routes such rows to synth_wrap).
constexpr int kLinMinChip = 32;
GS_HD bool lin_ok(int c0, int nb)
{
    return c0 >= kLinMinChip && (uint32_t)(c0 ^ (c0 + nb + 1)) <= (uint32_t)c0;
}
GS_HD uint32_t shr_clamp(uint32_t v, uint32_t n) // v >> n, 0 for n >= 32
{
#ifdef __CUDA_ARCH__
    uint32_t r;
    asm("shr.b32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(n));
    return r;
#else
    return n >= 32u ? 0u : v >> n;
#endif
}
template <class A, int S, int NB>
GS_HD void synth_lin(typename A::acc_t (&acc)[S], double &x, uint32_t &phs, const double d, const uint32_t steps,
                     const uint32_t gbits, const uint32_t win, const int c0, const double rinv0,
                     const typename A::tab_t *lut, const uint32_t lane_off)
{
    const double c = dfrombits(dbits(x) & 0x7ff0000000000000ull); // 2^e: the binade of x
    const double delta = dadd(dadd(c, d), -c);                    // RN_ulp(x)(d): what one step adds
    double t = dadd(int_as_double_exact((uint32_t)c0 + 1u), -x);  // distance to the next chip, exact, in (0, 1]
    uint32_t m = (uint32_t)((int32_t)win >> 31);                  // bit 31-j: sample j has chip sign "negative"
    const uint32_t tw = win ^ (win << 1);                         // bit 31-b: chip c0+b+1 differs from chip c0+b
#pragma unroll
    for (int b = 0; b < NB; b++) {
        const double q = dfma_down(t, rinv0, 4503599627370496.0); // 2^52 + floor(t / d_nominal)
        const double nf = dadd(q, -4503599627370496.0);
        const double r0 = dfma(nf, delta, -t);                     // x_nf - (c0+b+1), exact
        const double r1 = dadd(r0, delta);                         // x_(nf+1) - (c0+b+1), exact
        // first sample at or past the boundary: the estimate is within one of it, on either side
        const uint32_t n = (uint32_t)dbits(q) + (uint32_t)(dbits(r0) >> 63) + (uint32_t)(dbits(r1) >> 63);
        const uint32_t flip = (uint32_t)((int32_t)(tw << b) >> 31);
        m ^= flip & shr_clamp(0xffffffffu, n); // samples n.. change sign if the chips differ
        t = dadd(t, 1.0);
    }
    // +-gain by a select on a bit of the mask (predicates straight from the mask's bytes, R2P + SEL on the ALU
    // pipe).  The negated gain is made opaque: the compiler would otherwise rewrite the select as
    // gbits ^ ((m << j) & 2^31), a shift on the FMA pipe - the busiest one of this loop - plus a LOP3.
#ifdef __CUDA_ARCH__
    uint32_t gneg;
    asm("xor.b32 %0, %1, 0x80000000;" : "=r"(gneg) : "r"(gbits));
#else
    const uint32_t gneg = gbits ^ 0x80000000u;
#endif
#pragma unroll
    for (int j = 0; j < S; j++) {
        const uint32_t gs = (m & (0x80000000u >> j)) ? gneg : gbits;
        A::mad_s(acc[j], lut_at<A>(lut, phs, lane_off), gs);
        phs += steps;
    }
    x = dfma((double)S, delta, x);
}

// Per-thread, per-channel state of the integer-carrier kernel, 16 bytes in shared memory:
//   x (f64) | carr_phase << 7, low 7 bits = icode0 + wraps so far in this epoch (<= 19 + 101) | fp32 bits of dataBit*gain
GS_HD uint32_t lean_phase_word(uint32_t phs, int ic) { return (phs & ~127u) | (uint32_t)ic; }
GS_HD int lean_ic(uint32_t phase_word) { return (int)(phase_word & 127u); }

// Same samples, but the 1023-chip wrap (and with it the icode / data-bit walk of
// gpssim.c:2214-2238) may happen inside the run - at most once, a run is far shorter than a
// code period.  Branch-free: the chip window comes from a code table that continues past chip
// 1022 with chips 0,1,.. again, so after the wrap the advance is simply counted 1023 chips
// further; the data bit (and with it the signed gain) switches to a value prepared up front.
template <class A, int S>
GS_HD void synth_wrap(typename A::acc_t (&acc)[S], ChanState &st, const double d, const uint32_t steps,
                      const int32_t gain, const uint32_t nav_bits, const uint32_t win,
                      const typename A::tab_t *lut, const uint32_t lane_off)
{
    double x = st.x;
    uint32_t phs = st.phs;
    const int c0 = (int)x;
    const double magic = 6755399441055744.0 - (double)c0; // 1.5 * 2^52 - c0
    const int bit_after = st.bitk + (st.icode == 19 ? 1 : 0);
    typename A::gain_t g = A::make_gain(data_sign(nav_bits, st.bitk) * gain);
    const typename A::gain_t g_after = A::make_gain(data_sign(nav_bits, bit_after) * gain);
    uint32_t wrap_off = 0;
#pragma unroll
    for (int j = 0; j < S; j++) {
        const uint32_t adv = chips_since_signed(x, magic) + wrap_off;
        const uint32_t e = phs ^ ((win << adv) & 0x80000000u);
        A::mad(acc[j], lut_at<A>(lut, e, lane_off), g);
        x = dadd(x, d);
        phs += steps;
        const bool wrapped = x >= (double)kCaLen;
        x = wrapped ? dadd(x, -(double)kCaLen) : x;
        wrap_off = wrapped ? (uint32_t)kCaLen : wrap_off;
        g = wrapped ? g_after : g;
    }
    if (wrap_off) {
        st.icode = st.icode == 19 ? 0 : st.icode + 1;
        st.bitk = bit_after;
    }
    st.x = x;
    st.phs = phs;
}

// =====================================================================================
// FLOAT_CARR_PHASE hosts (gpssim.h:4 as shipped): the carrier phase is a double in [0,1) that is
// advanced by RN(f_carr*delt) per sample and wrapped both ways (gpssim.c:2245-2250), the table index
// is floor(carr_phase*512) (gpssim.c:2200).  The device keeps cph = 512*carr_phase (an exact
// power-of-two rescaling: every rounding commutes with it), so the index is floor(cph), again taken
// with one round-down magic add (carrier_index).
// =====================================================================================

GS_HD double carrier_step(double cph, const double dc) // gpssim.c:2245-2250, scaled by 512
{
    cph = dadd(cph, dc);
    if (cph >= kCarrMod)
        cph = dadd(cph, -kCarrMod);
    else if (cph < 0.0)
        cph = dadd(cph, kCarrMod);
    return cph;
}
// The same update when the sign of dc is known: from [0,512) a rising phase can only leave at the
// top, a falling one only at the bottom, so one of the two tests of the reference can never fire.
// One add, one compare, one predicated add - no branch.
template <bool kFalling>
GS_HD double carrier_step_signed(double cph, const double dc)
{
    cph = dadd(cph, dc);
#ifdef __CUDA_ARCH__
#if GS_CARRIER_WRAP_PRED
    // The wrap as ONE predicated add behind an INTEGER test of the upper word: ISETP + @P DADD, two
    // instructions and a 21-cycle recurrence.  (Round 1 wrote it as "add -512 or -0.0", the addend's upper
    // word selected with shift + LOP3: the compiler then re-creates the addend's zero lower word for every
    // sample - four instructions, 26 cycles; plain C makes it DADD + DSETP + two FSEL.)  Rising: cph in
    // [0,1024), so cph >= 512 <=> upper word >= 0x40800000.  Falling: cph in (-512,512) and never -0.0 (an
    // exact zero sum is +0.0 in round-to-nearest), so cph < 0 <=> sign bit.
    if (kFalling)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ge.s32 p, %1, 0;\n\t@p bra SKIP;\n\tadd.rn.f64 %0, %0, 0d4080000000000000;\nSKIP:\n\t}"
            : "+d"(cph) : "r"(__double2hiint(cph)));
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.lt.s32 p, %1, 0x40800000;\n\t@p bra SKIP;\n\tadd.rn.f64 %0, %0, 0dC080000000000000;\nSKIP:\n\t}"
            : "+d"(cph) : "r"(__double2hiint(cph)));
#else
    // The wrap as "add -512 or -0.0" chosen by an INTEGER test of the upper word: one compare and one
    // select on the ALU pipe and one DADD, instead of the compiler's DADD + DSETP + two FSEL (it
    // computes cph -+ 512 unconditionally and selects both halves).  Rising: cph in [0,1024), so
    // cph >= 512 <=> upper word >= 0x40800000.  Falling: cph in (-512,512) and never -0.0 (an exact
    // zero sum is +0.0 in round-to-nearest), so cph < 0 <=> sign bit.  x + (-0.0) == x for every x.
    const int hi = __double2hiint(cph);
    if (kFalling)
        cph = __dadd_rn(cph, __hiloint2double(hi < 0 ? 0x40800000 : 0, 0));
    else
        cph = __dadd_rn(cph, __hiloint2double(hi >= 0x40800000 ? (int)0xC0800000u : (int)0x80000000u, 0));
#endif
#else
    if (kFalling) {
        if (cph < 0.0)
            cph = dadd(cph, kCarrMod);
    } else {
        if (cph >= kCarrMod)
            cph = dadd(cph, -kCarrMod);
    }
#endif
    return cph;
}

struct ChanStateF {
    double x;      // code phase
    double cph;    // 512 * carr_phase
    int32_t icode; // 0..19
    int32_t bitk;  // data bits consumed since the row
};

// floor(cph) for cph in [0, 512): the table index of gpssim.c:2200, again one round-down magic add
GS_HD uint32_t carrier_index(double cph)
{
#ifdef __CUDA_ARCH__
    return (uint32_t)__double2loint(__dadd_rd(cph, 4503599627370496.0)); // 2^52
#else
    return (uint32_t)(int64_t)__builtin_floor(cph);
#endif
}

// Table entry for index idx (0..511) with the chip sign applied: the chip sign is half a cycle, i.e.
// bit 8 of the index; t8 carries the current (inverted) chip in bit 8 (other bits are junk).
// Device: the byte address (index*128 + this lane's replica) is ONE integer multiply-add on the FMA
// pipe - the ALU pipe, the busiest one, only does the xor - and the load goes through a 32-bit
// shared-window address.
template <class A>
GS_HD typename A::tab_t lut_at_f(const typename A::tab_t *lut, uint32_t idx, uint32_t t8, uint32_t lane_off)
{
    const uint32_t p = idx ^ (t8 & 0x100u);
#ifdef __CUDA_ARCH__
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(lut) + lane_off; // loop invariant
    uint32_t addr;
    asm("mad.lo.u32 %0, %1, 128, %2;" : "=r"(addr) : "r"(p), "r"(base));
    typename A::tab_t v;
    if (sizeof(typename A::tab_t) == 8)
        asm volatile("ld.shared.b64 %0, [%1];" : "=l"(*reinterpret_cast<uint64_t *>(&v)) : "r"(addr));
    else
        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(*reinterpret_cast<uint32_t *>(&v)) : "r"(addr));
    return v;
#else
    return *reinterpret_cast<const typename A::tab_t *>(reinterpret_cast<const char *>(lut) + ((p << 7) | lane_off));
#endif
}

// gbits: fp32 bits of dataBit*gain (kept in the thread's state: no conversion per run); c0 / magic: floor(x) and
// 2^52 - floor(x) from floor_magic() (two FP64 adds instead of a conversion each way).  The chip sign goes onto
// the gain: the table lookup depends on the carrier phase only.
template <class A, int S, bool kFalling>
GS_HD void synth_fast_f(typename A::acc_t (&acc)[S], ChanStateF &st, const double d, const double dc,
                        const uint32_t gbits, const uint32_t win, const double magic, const typename A::tab_t *lut,
                        const uint32_t lane_off)
{
    // (instantiated, never run, for the 64-bit integer accumulator: only AccF32x2 has a sign bit in its gain)
    double x = st.x, cph = st.cph;
#pragma unroll
    for (int j = 0; j < S; j++) {
        const uint32_t t = j == 0 ? win : win << chips_since(x, magic);
        A::mad_s(acc[j], lut_at_f<A>(lut, carrier_index(cph), 0u, lane_off), gbits ^ (t & 0x80000000u));
        x = dadd(x, d);
        cph = carrier_step_signed<kFalling>(cph, dc);
    }
    st.x = x;
    st.cph = cph;
}

template <class A, int S, bool kFalling>
GS_HD void synth_wrap_f(typename A::acc_t (&acc)[S], ChanStateF &st, const double d, const double dc,
                        const int32_t gain, const uint32_t nav_bits, const uint32_t win,
                        const typename A::tab_t *lut, const uint32_t lane_off)
{
    double x = st.x, cph = st.cph;
    const int c0 = (int)x;
    const double magic = 6755399441055744.0 - (double)c0; // 1.5 * 2^52 - c0
    const int bit_after = st.bitk + (st.icode == 19 ? 1 : 0);
    typename A::gain_t g = A::make_gain(data_sign(nav_bits, st.bitk) * gain);
    const typename A::gain_t g_after = A::make_gain(data_sign(nav_bits, bit_after) * gain);
    uint32_t wrap_off = 0;
    const uint32_t wlo = win << 9, whi = win >> 23;
#pragma unroll
    for (int j = 0; j < S; j++) {
        const uint32_t adv = chips_since_signed(x, magic) + wrap_off;
        A::mad(acc[j], lut_at_f<A>(lut, carrier_index(cph), funnel_l(wlo, whi, adv), lane_off), g);
        x = dadd(x, d);
        cph = carrier_step_signed<kFalling>(cph, dc);
        const bool wrapped = x >= (double)kCaLen;
        x = wrapped ? dadd(x, -(double)kCaLen) : x;
        wrap_off = wrapped ? (uint32_t)kCaLen : wrap_off;
        g = wrapped ? g_after : g;
    }
    if (wrap_off) {
        st.icode = st.icode == 19 ? 0 : st.icode + 1;
        st.bitk = bit_after;
    }
    st.x = x;
    st.cph = cph;
}

// ---- output packing (gpssim.c:2258-2288) from the biased sums -------------------------
GS_HD uint32_t pack_sc16(int32_t ib, int32_t qb) // little-endian short I, short Q
{
    return ((uint32_t)(ib - 8192) & 0xffffu) | ((uint32_t)(qb - 8192) << 16);
}
GS_HD uint32_t pack_sc08(int32_t ib, int32_t qb) // (signed char)(short>>4) for I then Q, in the low 16 bits
{
    return ((uint32_t)((ib >> 4) - 512) & 0xffu) | (((uint32_t)((qb >> 4) - 512) & 0xffu) << 8);
}
GS_HD uint32_t pack_sc01(int32_t ib, int32_t qb) // 2 bits: (I>0)<<1 | (Q>0)
{
    return ((ib > 8192) ? 2u : 0u) | ((qb > 8192) ? 1u : 0u);
}

// S consecutive samples of one thread -> their bytes at dst (16-byte aligned for SC16 / SC08,
// 2-byte aligned for SC01).  S is a multiple of 8.
GS_HD void store16(uint8_t *dst, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
#ifdef __CUDA_ARCH__
    *reinterpret_cast<uint4 *>(dst) = make_uint4(a, b, c, d);
#else
    const uint32_t w[4] = {a, b, c, d};
    memcpy(dst, w, 16);
#endif
}

GS_HD uint32_t byte_perm(uint32_t a, uint32_t b, uint32_t sel) // PRMT: result byte i = byte (sel nibble i) of b:a
{
#ifdef __CUDA_ARCH__
    return __byte_perm(a, b, sel);
#else
    const uint64_t v = (uint64_t)a | ((uint64_t)b << 32);
    uint32_t r = 0;
    for (int i = 0; i < 4; i++)
        r |= (uint32_t)((v >> (8 * ((sel >> (4 * i)) & 7u))) & 0xffu) << (8 * i);
    return r;
#endif
}
// one 16-bit sample (short I, short Q) / two 8-bit samples (I0 Q0 I1 Q1) from round-down FFMA2 results
template <class A>
GS_HD uint32_t word_sc16(typename A::acc_t acc)
{
    const uint64_t r = A::template floor_shift<7>(acc);
    return byte_perm((uint32_t)r, (uint32_t)(r >> 32), 0x5410u);
}
template <class A>
GS_HD uint32_t word_sc08(typename A::acc_t a0, typename A::acc_t a1)
{
    const uint64_t r0 = A::template floor_shift<11>(a0), r1 = A::template floor_shift<11>(a1);
    return byte_perm(byte_perm((uint32_t)r0, (uint32_t)(r0 >> 32), 0x0040u),
                     byte_perm((uint32_t)r1, (uint32_t)(r1 >> 32), 0x0040u), 0x5410u);
}

template <class A, int FMT, int S>
GS_HD void store_run(uint8_t *dst, const typename A::acc_t (&acc)[S])
{
    if (A::kFloorFma && FMT == 16) {
#pragma unroll
        for (int q = 0; q < S / 4; q++)
            store16(dst + 16 * q, word_sc16<A>(acc[4 * q]), word_sc16<A>(acc[4 * q + 1]), word_sc16<A>(acc[4 * q + 2]),
                    word_sc16<A>(acc[4 * q + 3]));
        return;
    }
    if (A::kFloorFma && FMT == 8) {
#pragma unroll
        for (int q = 0; q < S / 8; q++)
            store16(dst + 16 * q, word_sc08<A>(acc[8 * q], acc[8 * q + 1]), word_sc08<A>(acc[8 * q + 2], acc[8 * q + 3]),
                    word_sc08<A>(acc[8 * q + 4], acc[8 * q + 5]), word_sc08<A>(acc[8 * q + 6], acc[8 * q + 7]));
        return;
    }
    if (A::kFloorFma && FMT == 1) {
        // byte b = samples 4b..4b+3 as I0 Q0 I1 Q1 I2 Q2 I3 Q3, MSB first (gpssim.c:2268-2274): one FFMA2 puts
        // "I > 0" and "Q > 0" into two sign bits, one funnel shift each appends them to the byte
#pragma unroll
        for (int q = 0; q < S / 8; q++) {
            uint32_t b[2] = {0u, 0u};
#pragma unroll
            for (int h = 0; h < 2; h++)
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const uint64_t r = A::positive_signs(acc[8 * q + 4 * h + j]);
                    b[h] = funnel_l((uint32_t)r, b[h], 1u);
                    b[h] = funnel_l((uint32_t)(r >> 32), b[h], 1u);
                }
            const uint16_t w = (uint16_t)byte_perm(b[0], b[1], 0x0040u);
#ifdef __CUDA_ARCH__
            *reinterpret_cast<uint16_t *>(dst + 2 * q) = w;
#else
            memcpy(dst + 2 * q, &w, 2);
#endif
        }
        return;
    }
#define GS_P16(j) pack_sc16(A::i_biased(acc[j]), A::q_biased(acc[j]))
#define GS_P08(j) pack_sc08(A::i_biased(acc[j]), A::q_biased(acc[j]))
#define GS_P01(j) pack_sc01(A::i_biased(acc[j]), A::q_biased(acc[j]))
    if (FMT == 16) {
#pragma unroll
        for (int q = 0; q < S / 4; q++)
            store16(dst + 16 * q, GS_P16(4 * q), GS_P16(4 * q + 1), GS_P16(4 * q + 2), GS_P16(4 * q + 3));
    } else if (FMT == 8) {
#pragma unroll
        for (int q = 0; q < S / 8; q++) {
            uint32_t w[4];
#pragma unroll
            for (int h = 0; h < 4; h++)
                w[h] = GS_P08(8 * q + 2 * h) | (GS_P08(8 * q + 2 * h + 1) << 16);
            store16(dst + 16 * q, w[0], w[1], w[2], w[3]);
        }
    } else {
        // byte b holds samples 4b..4b+3 as I0 Q0 I1 Q1 I2 Q2 I3 Q3, MSB first (gpssim.c:2268-2274);
        // two bytes (8 samples) per store: a chunk only has to start on a multiple of 8 samples
#pragma unroll
        for (int q = 0; q < S / 8; q++) {
            uint32_t w = 0;
#pragma unroll
            for (int b = 0; b < 2; b++) {
                const int s0 = 8 * q + 4 * b;
                const uint32_t byte = (GS_P01(s0) << 6) | (GS_P01(s0 + 1) << 4) | (GS_P01(s0 + 2) << 2) | GS_P01(s0 + 3);
                w |= byte << (8 * b);
            }
#ifdef __CUDA_ARCH__
            *reinterpret_cast<uint16_t *>(dst + 2 * q) = (uint16_t)w;
#else
            const uint16_t h = (uint16_t)w;
            memcpy(dst + 2 * q, &h, 2);
#endif
        }
    }
#undef GS_P16
#undef GS_P08
#undef GS_P01
}

// =====================================================================================
// Generic exact path: one sample, one channel at a time, any samples_per_epoch, any
// gain, any f_code*delt.  Slow (scalar stores, table reads from global memory); used when
// a table violates the tuned kernel's ranges and as the on-device cross-check of it.
// =====================================================================================
struct GenericChan {
    double x, d;
    double cph, dc;    // FLOAT hosts: 512*carr_phase and its per-sample step
    uint32_t phs;      // INT hosts: carr_phase << 7
    int32_t steps, gain, icode, bitk;
    uint32_t nav_bits;
    const uint32_t *negw;
};

template <bool kFloatCarrier>
GS_HD void generic_sample(GenericChan *ch, int nc, const int16_t *sin512, const int16_t *cos512,
                          int &i16, int &q16)
{
    int i_acc = 0, q_acc = 0;
    for (int k = 0; k < nc; k++) {
        GenericChan &c = ch[k];
        const int chip = (int)c.x;
        const int neg = (int)((c.negw[chip >> 5] >> (31 - (chip & 31))) & 1u);
        const int sgn = (neg ? -1 : 1) * data_sign(c.nav_bits, c.bitk);
        // (carr_phase >> 16) & 0x1ff (gpssim.c:2202) / (int)floor(carr_phase*512) (gpssim.c:2200)
        const int it = kFloatCarrier ? ((int)c.cph & 0x1ff) : (int)(c.phs >> 23);
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
        if (kFloatCarrier)
            c.cph = carrier_step(c.cph, c.dc);
        else
            c.phs += (uint32_t)c.steps;
    }
    i16 = (int)(short)((i_acc + 64) >> 7);
    q16 = (int)(short)((q_acc + 64) >> 7);
}

// =====================================================================================
// K0 - navigation data words (replaces generateNavMsg, gpssim.c:1467-1547, and its parity
// routine computeChecksum, gpssim.c:693-756, for hosts that hand the library subframes instead of data bits).
//
// A frame is what ONE generateNavMsg() call leaves in chan->dwrd[60]: the last subframe of the previous
// frame (words 0..9) followed by subframes 1..5 (words 10..59) of the 30 s that start at g0.  Every word is
// the 24 source bits from eph2sbf() plus, in the hand-over word (word 2 of a subframe), the TOW count of the
// NEXT subframe and, in word 3 of subframe 1, the week number - then the six parity bits of IS-GPS-200
// 20.3.5.2, chained through the two last bits of the word before.  Words 2 and 10 of a subframe solve their
// bits 23/24 so that their own two last bits come out zero, which is why every subframe starts from a
// previous word of 0 and the six subframes of a frame can be built independently of each other (the
// reference's refresh copies words 50..59 of the previous frame, which end in such a word).
// =====================================================================================
constexpr int kNavWordsPerSubframe = 10; // N_DWRD_SBF, gpssim.h:30
constexpr int kNavSubframes = 6;         // N_SBF + 1,  gpssim.h:27,:33
constexpr int kNavWords = kNavWordsPerSubframe * kNavSubframes; // N_DWRD

GS_HD uint32_t nav_popc(uint32_t v)
{
#ifdef __CUDA_ARCH__
    return (uint32_t)__popc(v);
#else
    return (uint32_t)__builtin_popcount(v);
#endif
}

// One transmitted word from its source word.  src: bits 29..6 = d1..d24, bits 31/30 = D29*/D30* of the word
// sent before.  Returns D1..D30 in bits 29..0 (data bits inverted when D30* is set, parity in bits 5..0).
// solve_tail: word 2 / word 10 - bits d23, d24 are not information; choose them so that D29 = D30 = 0.
GS_HD uint32_t nav_word(uint32_t src, bool solve_tail)
{
    // rows of the parity matrix over d1..d24 (bit 29 = d1), IS-GPS-200 table 20-XIV; `star` = which of
    // D29* (1) / D30* (0) enters the row
    const uint32_t rows[6] = {0x3B1F3480u, 0x1D8F9A40u, 0x2EC7CD00u, 0x1763E680u, 0x2BB1F340u, 0x0B7A89C0u};
    const uint32_t star[6] = {1u, 0u, 1u, 0u, 0u, 1u};
    const uint32_t prev[2] = {(src >> 30) & 1u, (src >> 31) & 1u}; // [0] = D30*, [1] = D29*
    uint32_t d = src & 0x3FFFFFC0u;
    if (solve_tail) {
        // d24 (bit 6) is in row D29 but d23 (bit 7) is not; both are in row D30: fix D29 with d24 first
        if ((prev[star[4]] + nav_popc(rows[4] & d)) & 1u)
            d ^= 1u << 6;
        if ((prev[star[5]] + nav_popc(rows[5] & d)) & 1u)
            d ^= 1u << 7;
    }
    uint32_t w = prev[0] ? d ^ 0x3FFFFFC0u : d;
#pragma unroll
    for (int r = 0; r < 6; r++)
        w |= ((prev[star[r]] + nav_popc(rows[r] & d)) & 1u) << (5 - r);
    return w & 0x3FFFFFFFu;
}

// Ten words of one subframe.  src10: the subframe's source words (chan->sbf[isbf], gpssim.h:174); tow_next:
// TOW count written into the hand-over word; week10: < 0, or the 10-bit week number for word 3 (subframe 1).
GS_HD void nav_subframe(const uint32_t *src10, uint32_t tow_next, int week10, uint32_t *out10)
{
    uint32_t before = 0u;
#pragma unroll 1
    for (int i = 0; i < kNavWordsPerSubframe; i++) {
        uint32_t sw = src10[i];
        if (i == 2 && week10 >= 0)
            sw |= ((uint32_t)week10 & 0x3FFu) << 20; // gpssim.c:1527
        if (i == 1)
            sw |= (tow_next & 0x1FFFFu) << 13;        // gpssim.c:1531, :1493
        sw |= before << 30;                            // D29*, D30* (upper bits of `before` fall off)
        before = nav_word(sw, i == 1 || i == 9);
        out10[i] = before;
    }
}

// The 32 data bits a row starts with: bit 31 = (dwrd[iword] >> (29 - ibit)) & 1, then the bits the sample
// loop walks to (gpssim.c:2223-2236); past word 59 they read 0.  Same contract as gpusim_pack_nav_bits().
GS_HD uint32_t nav_row_bits(const uint32_t *dwrd60, int iword, int ibit)
{
    uint32_t out = 0u;
    int w = iword, b = ibit;
    for (int k = 0; k < 32; k++) {
        if (w >= 0 && w < kNavWords)
            out |= ((dwrd60[w] >> (29 - b)) & 1u) << (31 - k);
        if (++b == 30) {
            b = 0;
            w++;
        }
    }
    return out;
}

// Device layout of one frame request (what crosses the ABI as gpusim_nav_frame): 64 words.
struct NavFrame {
    uint32_t sbf[5][kNavWordsPerSubframe]; // source words of subframes 1..5
    uint32_t first[kNavWordsPerSubframe];  // source words of the frame's words 0..9 (a subframe 5)
    uint32_t tow_first;                    // TOW count for those
    uint32_t tow;                          // TOW count of the frame start; subframe s (1..5) carries tow + s
    uint32_t week10;                       // transmission week number mod 1024
    uint32_t reserved;
};
static_assert(sizeof(NavFrame) == 256, "NavFrame layout");

// subframe s of frame f (s = 0: the leading subframe 5, s = 1..5: subframes 1..5) -> out[s*10 .. s*10+9]
GS_HD void nav_build_subframe(const NavFrame &f, int s, uint32_t *dwrd60)
{
    if (s == 0)
        nav_subframe(f.first, f.tow_first, -1, dwrd60);
    else
        nav_subframe(f.sbf[s - 1], f.tow + (uint32_t)s, s == 1 ? (int)(f.week10 & 0x3FFu) : -1, dwrd60 + s * kNavWordsPerSubframe);
}

// ---- subframes from an ephemeris (replaces eph2sbf, gpssim.c:490-665) ---------------------------------------
// The 24 source bits of every word of subframes 1..3 (clock and orbit of the satellite itself), of page 18 of
// subframe 4 (ionosphere / UTC; page 25 when the host has no such data) and of page 25 of subframe 5, IS-GPS-200
// 20.3.3.3 - 20.3.3.5, from the broadcast values the host read from the RINEX file.  Every field is the
// reference's quotient "value / LSB" (its divisors are the decimal constants of gpssim.h:44-55 - written out here
// with the same digits, they are not exact powers of two - and pi = 3.1415926535898), truncated towards zero by the
// cast to long (rounded half away from zero for the ionosphere / UTC terms), two's complement, cut to the field width.
struct NavEph {   // what crosses the ABI as gpusim_nav_eph: the fields of ephem_t that eph2sbf() reads (gpssim.h:101-135)
    double toe_sec, toc_sec;
    double deltan, cuc, cus, cic, cis, crc, crs, ecc, sqrta, m0, omg0, inc0, aop, omgdot, idot, af0, af1, af2, tgd;
    int32_t toe_week, iodc, iode, svhlth, codeL2, reserved;
};
static_assert(sizeof(NavEph) == 192, "NavEph layout");
struct NavIono {  // gpusim_nav_iono: ionoutc_t (gpssim.h:137-146)
    double alpha0, alpha1, alpha2, alpha3, beta0, beta1, beta2, beta3, A0, A1;
    int32_t vflg, dtls, tot, wnt;
};
static_assert(sizeof(NavIono) == 96, "NavIono layout");
struct NavFrameRef { // gpusim_nav_frame_ref: a frame whose subframes come from device-built ephemeris subframes
    int32_t eph, eph_first;
    uint32_t tow_first, tow, week10, reserved;
};
static_assert(sizeof(NavFrameRef) == 24, "NavFrameRef layout");
constexpr int kNavSbfWords = 5 * kNavWordsPerSubframe;

GS_HD int64_t nav_trunc(double q) { return (int64_t)q; }   // (long)(x): towards zero
GS_HD int64_t nav_round(double q)                          // (signed long)round(x): half away from zero
{
#ifdef __CUDA_ARCH__
    return (int64_t)round(q);
#else
    return (int64_t)__builtin_round(q);
#endif
}
// `width` low bits of a two's complement value, placed with their LSB at bit `at` of the word
GS_HD uint32_t nav_field(int64_t v, int width, int at) { return (uint32_t)(((uint64_t)v & ((1ull << width) - 1ull)) << at); }

GS_HD void nav_eph_subframes(const NavEph &e, const NavIono &io, uint32_t *sbf /* [5][10] */)
{
    const double pi = 3.1415926535898;                      // gpssim.h:60
    const double m5 = 0.03125, m19 = 1.907348632812500e-6, m29 = 1.862645149230957e-9, m31 = 4.656612873077393e-10,
                 m33 = 1.164153218269348e-10, m43 = 1.136868377216160e-13, m55 = 2.775557561562891e-17,
                 m50 = 8.881784197001252e-016, m30 = 9.313225746154785e-010, m27 = 7.450580596923828e-009,
                 m24 = 5.960464477539063e-008;              // gpssim.h:44-55, digit for digit
    const int64_t toe = nav_trunc(e.toe_sec / 16.0), toc = nav_trunc(e.toc_sec / 16.0);
    const int64_t iode = e.iode, iodc = e.iodc;
    const int64_t deltan = nav_trunc(e.deltan / m43 / pi), omgdot = nav_trunc(e.omgdot / m43 / pi), idot = nav_trunc(e.idot / m43 / pi);
    const int64_t cuc = nav_trunc(e.cuc / m29), cus = nav_trunc(e.cus / m29), cic = nav_trunc(e.cic / m29), cis = nav_trunc(e.cis / m29);
    const int64_t crc = nav_trunc(e.crc / m5), crs = nav_trunc(e.crs / m5);
    const int64_t ecc = nav_trunc(e.ecc / m33), sqrta = nav_trunc(e.sqrta / m19);
    const int64_t m0 = nav_trunc(e.m0 / m31 / pi), omg0 = nav_trunc(e.omg0 / m31 / pi), inc0 = nav_trunc(e.inc0 / m31 / pi),
                  aop = nav_trunc(e.aop / m31 / pi);
    const int64_t af0 = nav_trunc(e.af0 / m31), af1 = nav_trunc(e.af1 / m43), af2 = nav_trunc(e.af2 / m55), tgd = nav_trunc(e.tgd / m31);
    const uint32_t tlm = 0x8Bu << 22;                       // preamble 10001011 in d1..d8, the rest of the TLM word zero
    for (int i = 0; i < kNavSbfWords; i++)
        sbf[i] = 0u;
    for (int sub = 0; sub < 5; sub++) {
        sbf[sub * 10 + 0] = tlm;
        sbf[sub * 10 + 1] = (uint32_t)(sub + 1) << 8;       // subframe ID of the hand-over word; TOW count is added per frame
    }
    uint32_t *s1 = sbf, *s2 = sbf + 10, *s3 = sbf + 20, *s4 = sbf + 30, *s5 = sbf + 40;
    // subframe 1: week number (added per frame) | L2 code | URA 0 | health | IODC msbs ; Tgd ; IODC lsbs | toc ; af2 | af1 ; af0
    s1[2] = nav_field(e.codeL2, 2, 18) | nav_field(e.svhlth, 6, 8) | nav_field(iodc >> 8, 2, 6);
    s1[6] = nav_field(tgd, 8, 6);
    s1[7] = nav_field(iodc, 8, 22) | nav_field(toc, 16, 6);
    s1[8] = nav_field(af2, 8, 22) | nav_field(af1, 16, 6);
    s1[9] = nav_field(af0, 22, 8);
    // subframe 2: IODE | Crs ; delta n | M0 msbs ; M0 lsbs ; Cuc | e msbs ; e lsbs ; Cus | sqrt(A) msbs ; sqrt(A) lsbs ; toe
    s2[2] = nav_field(iode, 8, 22) | nav_field(crs, 16, 6);
    s2[3] = nav_field(deltan, 16, 14) | nav_field(m0 >> 24, 8, 6);
    s2[4] = nav_field(m0, 24, 6);
    s2[5] = nav_field(cuc, 16, 14) | nav_field(ecc >> 24, 8, 6);
    s2[6] = nav_field(ecc, 24, 6);
    s2[7] = nav_field(cus, 16, 14) | nav_field(sqrta >> 24, 8, 6);
    s2[8] = nav_field(sqrta, 24, 6);
    s2[9] = nav_field(toe, 16, 14);
    // subframe 3: Cic | Omega0 msbs ; lsbs ; Cis | i0 msbs ; lsbs ; Crc | omega msbs ; lsbs ; Omega dot ; IODE | IDOT
    s3[2] = nav_field(cic, 16, 14) | nav_field(omg0 >> 24, 8, 6);
    s3[3] = nav_field(omg0, 24, 6);
    s3[4] = nav_field(cis, 16, 14) | nav_field(inc0 >> 24, 8, 6);
    s3[5] = nav_field(inc0, 24, 6);
    s3[6] = nav_field(crc, 16, 14) | nav_field(aop >> 24, 8, 6);
    s3[7] = nav_field(aop, 24, 6);
    s3[8] = nav_field(omgdot, 24, 6);
    s3[9] = nav_field(iode, 8, 22) | nav_field(idot, 14, 8);
    // subframe 4: data ID 1, page 18 (SV ID 56: ionosphere and UTC) when the host has the parameters, else page 25 (SV ID 63)
    if (io.vflg == 1) {
        const int64_t a0 = nav_round(io.alpha0 / m30), a1 = nav_round(io.alpha1 / m27), a2 = nav_round(io.alpha2 / m24),
                      a3 = nav_round(io.alpha3 / m24);
        const int64_t b0 = nav_round(io.beta0 / 2048.0), b1 = nav_round(io.beta1 / 16384.0), b2 = nav_round(io.beta2 / 65536.0),
                      b3 = nav_round(io.beta3 / 65536.0);
        const int64_t A0 = nav_round(io.A0 / m30), A1 = nav_round(io.A1 / m50);
        const int64_t tot = io.tot / 4096, wnt = io.wnt % 256;
        const int64_t wnlsf = 1929 % 256, dn = 7, dtlsf = 18; // the reference's fixed leap-second schedule (gpssim.c:585-587)
        s4[2] = nav_field(1, 2, 28) | nav_field(56, 6, 22) | nav_field(a0, 8, 14) | nav_field(a1, 8, 6);
        s4[3] = nav_field(a2, 8, 22) | nav_field(a3, 8, 14) | nav_field(b0, 8, 6);
        s4[4] = nav_field(b1, 8, 22) | nav_field(b2, 8, 14) | nav_field(b3, 8, 6);
        s4[5] = nav_field(A1, 24, 6);
        s4[6] = nav_field(A0 >> 8, 24, 6);
        s4[7] = nav_field(A0, 8, 22) | nav_field(tot, 8, 14) | nav_field(wnt, 8, 6);
        s4[8] = nav_field(io.dtls, 8, 22) | nav_field(wnlsf, 8, 14) | nav_field(dn, 8, 6);
        s4[9] = nav_field(dtlsf, 8, 22);
    } else {
        s4[2] = nav_field(1, 2, 28) | nav_field(63, 6, 22);
    }
    // subframe 5: data ID 1, page 25 (SV ID 51): almanac reference time and week
    s5[2] = nav_field(1, 2, 28) | nav_field(51, 6, 22) | nav_field(nav_trunc(e.toe_sec / 4096.0), 8, 14) | nav_field(e.toe_week % 256, 8, 6);
}

// subframe s of a frame whose source words are device-built ephemeris subframes
GS_HD void nav_build_subframe_ref(const NavFrameRef &f, const uint32_t *sbf_all, int s, uint32_t *dwrd60)
{
    if (s == 0)
        nav_subframe(sbf_all + (size_t)f.eph_first * kNavSbfWords + 4 * kNavWordsPerSubframe, f.tow_first, -1, dwrd60);
    else
        nav_subframe(sbf_all + (size_t)f.eph * kNavSbfWords + (s - 1) * kNavWordsPerSubframe, f.tow + (uint32_t)s,
                     s == 1 ? (int)(f.week10 & 0x3FFu) : -1, dwrd60 + s * kNavWordsPerSubframe);
}

} // namespace gpusim
#endif
