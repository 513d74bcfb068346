// gpusim_tables.cpp - constant tables of the sample-synthesis path, host side.
//
// These are the three pieces of read-only data the reference's sample loop
// indexes (gpssim.c:2204-2205 sin/cosTable512, gpssim.c:2241 chan[i].ca[],
// gpssim.c:2236 chan[i].dwrd[]) in the form the device kernels consume.
// Nothing here is copied from the reference: the carrier table is regenerated
// from its closed form and guarded by a checksum, the C/A code comes from the
// ICD-GPS-200 G1/G2 shift registers.  tests/test_tables.py checks both against
// the reference's own arrays / codegen() whenever oracle/_ref is available.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "gpusim.h"
#include "gpusim_tables.h"
#include "gpusim_core.h"

namespace gpusim {

// sinTable512 (gpssim.c:15-48) is a 512-point sine of amplitude 250 sampled at
// the centre of each phase bin; cosTable512 (gpssim.c:50-83) is the same table
// advanced by a quarter cycle.  With the nominal amplitude 250 one entry
// (250*sin(2*pi*35.5/512) = 105.50007) sits 7e-5 from a rounding boundary and
// the reference holds 105, so the closed form below uses 249.9925 (any value in
// [249.984, 249.995] reproduces the table): every entry is then >= 0.0029 away
// from a rounding boundary - twelve orders of magnitude more than libm's sin
// error - and the rounded values equal the reference table exactly.
void carrier_lut(int32_t *sin512, int32_t *cos512)
{
    const double two_pi = 6.283185307179586476925286766559;
    int32_t quarter[128];
    for (int i = 0; i < 128; i++)
        quarter[i] = (int32_t)std::lrint(249.9925 * std::sin(two_pi * (i + 0.5) / 512.0));
    for (int i = 0; i < 512; i++) {
        int k = i & 255;
        int32_t v = (k < 128) ? quarter[k] : quarter[255 - k];
        sin512[i] = (i < 256) ? v : -v;
    }
    for (int i = 0; i < 512; i++)
        cos512[i] = sin512[(i + 128) & 511];

    // self-check: position-weighted checksum of the 128 defining values
    uint32_t h = 2166136261u;
    for (int i = 0; i < 128; i++)
        h = (h ^ (uint32_t)quarter[i]) * 16777619u;
    if (h != kCarrierLutFnv1a) {
        std::fprintf(stderr, "gpusim: carrier table self-check failed (0x%08x)\n", h);
        std::abort();
    }
}

// G2 code phase delay in chips for PRN 1..32 (ICD-GPS-200, Table 3-I).
static const uint16_t kG2Delay[32] = {5,   6,   7,   8,   17,  18,  139, 140, 141, 251, 252,
                                      254, 255, 256, 257, 258, 469, 470, 471, 472, 473, 474,
                                      509, 512, 513, 514, 515, 516, 859, 860, 861, 862};

static inline unsigned parity10(unsigned v)
{
    v ^= v >> 8;
    v ^= v >> 4;
    v ^= v >> 2;
    v ^= v >> 1;
    return v & 1u;
}

// chips[i] in {0,1}; equals codegen(ca, prn) of the reference (gpssim.c:132-171).
int ca_code(int prn, uint8_t *chips)
{
    if (prn < 1 || prn > 32)
        return -1;
    uint8_t g1[GPUSIM_CA_SEQ_LEN], g2[GPUSIM_CA_SEQ_LEN];
    unsigned r1 = 0x3ff, r2 = 0x3ff; // bit k = stage k+1, all ones at start
    for (int i = 0; i < GPUSIM_CA_SEQ_LEN; i++) {
        g1[i] = (r1 >> 9) & 1u; // stage 10
        g2[i] = (r2 >> 9) & 1u;
        unsigned f1 = parity10(r1 & 0x204u); // G1 = 1 + x^3 + x^10
        unsigned f2 = parity10(r2 & 0x3a6u); // G2 = 1 + x^2 + x^3 + x^6 + x^8 + x^9 + x^10
        r1 = ((r1 << 1) | f1) & 0x3ffu;
        r2 = ((r2 << 1) | f2) & 0x3ffu;
    }
    const int delay = kG2Delay[prn - 1];
    for (int i = 0; i < GPUSIM_CA_SEQ_LEN; i++)
        chips[i] = g1[i] ^ g2[(i + GPUSIM_CA_SEQ_LEN - delay) % GPUSIM_CA_SEQ_LEN];
    return 0;
}

// 35 words per PRN: chip k (k < 1023+64, taken modulo 1023) at word k>>5, bit 31-(k&31).
void ca_words(int prn, uint32_t *words35)
{
    uint8_t chips[GPUSIM_CA_SEQ_LEN];
    for (int w = 0; w < kCaWordsPerPrn; w++)
        words35[w] = 0;
    if (ca_code(prn, chips) != 0)
        return;
    for (int k = 0; k < GPUSIM_CA_SEQ_LEN + 64; k++)
        if (chips[k % GPUSIM_CA_SEQ_LEN])
            words35[k >> 5] |= 0x80000000u >> (k & 31);
}

} // namespace gpusim

extern "C" {

void gpusim_carrier_lut(int32_t *sin512, int32_t *cos512) { gpusim::carrier_lut(sin512, cos512); }

int gpusim_ca_code(int32_t prn, int32_t *ca1023)
{
    uint8_t chips[GPUSIM_CA_SEQ_LEN];
    if (gpusim::ca_code(prn, chips) != 0)
        return GPUSIM_ERR_ARG;
    for (int i = 0; i < GPUSIM_CA_SEQ_LEN; i++)
        ca1023[i] = chips[i];
    return GPUSIM_OK;
}

// dataBit at (iword, ibit) is (dwrd[iword] >> (29-ibit)) & 1 (gpssim.c:1345, :2236); the sample
// loop then walks ibit -> 30 -> next iword (gpssim.c:2223-2228).  Bits past the buffer read 0.
uint32_t gpusim_pack_nav_bits(const unsigned long *dwrd, int32_t n_dwrd, int32_t iword, int32_t ibit)
{
    uint32_t out = 0;
    long flat = (long)iword * 30 + ibit;
    for (int k = 0; k < 32; k++, flat++) {
        long w = flat / 30;
        int b = (int)(flat % 30);
        if (flat < 0 || w >= n_dwrd)
            continue;
        out |= (uint32_t)((dwrd[w] >> (29 - b)) & 1UL) << (31 - k);
    }
    return out;
}

// N executions of the FLOAT_CARR_PHASE carrier update (gpssim.c:2245-2250) in O(carrier cycles)
// instead of O(N): the same exact binade walk the device uses, on the host.
} // extern "C"

namespace {
inline double advance_carrier(double carr_phase, double f_carr, double delt, int32_t n_samples)
{
    const volatile double d = f_carr * delt;
    if (n_samples <= 0 || d == 0.0)
        return carr_phase;
    auto nothing = [](int, double, int) {};
    const double x0 = carr_phase * 512.0, d512 = (double)d * 512.0;
    return (d512 < 0.0 ? gpusim::phase_chain<-1>(x0, d512, 512.0, n_samples, 1 << 30, nothing)
                       : gpusim::phase_chain<1>(x0, d512, 512.0, n_samples, 1 << 30, nothing)) / 512.0;
}
#if defined(__x86_64__) && defined(__GNUC__)
// The walk is a chain of fused multiply-adds.  Built for plain x86-64 (the reference's flags) every one of them
// is a call into libm's software fma; with the FMA3 instructions of every CPU a B200 host can have it is one
// instruction - same single rounding, same value.  (Explicit fma() calls only: -ffp-contract=off still keeps the
// compiler from fusing any a*b+c on its own.)
__attribute__((target("fma"))) double advance_carrier_fma3(double carr_phase, double f_carr, double delt, int32_t n_samples)
{
    return advance_carrier(carr_phase, f_carr, delt, n_samples);
}
#endif
} // namespace

extern "C" {
double gpusim_advance_carrier_f64(double carr_phase, double f_carr, double delt, int32_t n_samples)
{
#if defined(__x86_64__) && defined(__GNUC__)
    static const bool has_fma = __builtin_cpu_supports("fma");
    if (has_fma)
        return advance_carrier_fma3(carr_phase, f_carr, delt, n_samples);
#endif
    return advance_carrier(carr_phase, f_carr, delt, n_samples);
}

} // extern "C"
