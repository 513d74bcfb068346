// emu.cpp - runs the device algorithms of gps_sdr_sim_b200/csrc/gpusim_core.h thread by
// thread on the CPU.  TEST INFRASTRUCTURE ONLY: it lets the CPU test-suite compare the
// exact code the kernels execute (chain walk, packed accumulator, wrap handling, packing)
// with the oracle before any GPU time is spent.  The product never loads this.
#include <cstdint>
#include <cstring>
#include <vector>

#include "gpusim.h"
#include "gpusim_core.h"
#include "gpusim_tables.h"

using namespace gpusim;

namespace {

struct Tables {
    std::vector<int32_t> lut;   // [512][32] replicated like the kernel's shared memory
    std::vector<uint64_t> lut2; // [512][16]
    std::vector<int16_t> s16, c16;
    std::vector<uint32_t> negw; // [33][33]
    Tables() : lut(512 * 32), lut2(512 * 16), s16(512), c16(512), negw(kCaPrns * kCaWords, 0xffffffffu)
    {
        int32_t s[512], c[512];
        carrier_lut(s, c);
        for (int i = 0; i < 512; i++) {
            for (int l = 0; l < 32; l++)
                lut[i * 32 + l] = AccWide::table_entry(c[i], s[i]);
            for (int l = 0; l < 16; l++)
                lut2[i * 16 + l] = AccF32x2::table_entry(c[i], s[i]);
            s16[i] = (int16_t)s[i];
            c16[i] = (int16_t)c[i];
        }
        for (int prn = 1; prn <= 32; prn++) {
            uint32_t w[kCaWords];
            ca_words(prn, w);
            for (int i = 0; i < kCaWords; i++)
                negw[prn * kCaWords + i] = ~w[i];
        }
    }
};

template <class A> const typename A::tab_t *table_of(const Tables &T);
template <> const int32_t *table_of<AccWide>(const Tables &T) { return T.lut.data(); }
template <> const uint64_t *table_of<AccF32x2>(const Tables &T) { return T.lut2.data(); }

// one run of SR samples, all channels (mirrors synth_run in gpusim_kernels.cu, one lane at a time)
template <class A, int FMT, int SR>
void emu_run(const Tables &T, const DevRow *rows, int nc, ChanState *st, uint32_t *meta, int force_wrap,
             uint32_t lane_off, uint8_t *dst)
{
    const typename A::tab_t *lut = table_of<A>(T);
    typename A::acc_t acc[SR];
    for (int j = 0; j < SR; j++)
        acc[j] = A::init();
    for (int k = 0; k < nc; k++) {
        const DevRow &r = rows[k];
        const bool wrap = (int)st[k].x >= row_cthr(r) || force_wrap;
        const uint32_t *nw = T.negw.data() + (size_t)row_prn(r) * kCaWords;
        if (!wrap) {
            synth_fast<A, SR>(acc, st[k], r.d, (uint32_t)r.steps, meta_sgain(meta[k]), chip_window(nw, (int)st[k].x), lut, lane_off);
        } else {
            st[k].icode = meta_icode(meta[k]);
            st[k].bitk = meta_bitk(meta[k]);
            synth_wrap<A, SR>(acc, st[k], r.d, (uint32_t)r.steps, r.gain, r.nav_bits, chip_window(nw, (int)st[k].x), lut, lane_off);
            meta[k] = pack_meta(st[k].icode, st[k].bitk, data_sign(r.nav_bits, st[k].bitk) * r.gain);
        }
    }
    store_run<A, FMT, SR>(dst, acc);
}

template <class A, int FMT, int S>
void tuned_chunk(const Tables &T, const DevRow *rows, int nc, const double *ckx, const uint16_t *ckw, int kc,
                 int jc, int chunk, int N, int force_wrap, int lane, uint8_t *epoch_out)
{
    const int n0 = jc * chunk;
    const int nrun = std::min(chunk, N - n0);
    ChanState st[kMaxChan];
    uint32_t meta[kMaxChan];
    for (int k = 0; k < nc; k++) {
        const int ic = rows[k].icode0 + ckw[k * kc + jc];
        const int bitk = ic / 20;
        st[k].x = ckx[k * kc + jc];
        st[k].phs = rows[k].ph0s + (uint32_t)n0 * (uint32_t)rows[k].steps;
        meta[k] = pack_meta(ic - bitk * 20, bitk, data_sign(rows[k].nav_bits, bitk) * rows[k].gain);
    }
    constexpr int kBytesPer8 = (FMT == 16) ? 32 : (FMT == 8) ? 16 : 2;
    uint8_t *outp = epoch_out + (size_t)(n0 / 8) * kBytesPer8;
    const uint32_t lane_off = (uint32_t)(lane & A::kLaneMask) << A::kLaneShift;
    const int full = nrun / S, tail8 = (nrun - full * S) / 8;
    for (int i = 0; i < full; i++)
        emu_run<A, FMT, S>(T, rows, nc, st, meta, force_wrap, lane_off, outp + (size_t)i * (S / 8) * kBytesPer8);
    for (int i = 0; i < tail8; i++)
        emu_run<A, FMT, 8>(T, rows, nc, st, meta, force_wrap, lane_off, outp + ((size_t)full * (S / 8) + i) * kBytesPer8);
}

// ---- integer carrier, k2_lean (mirrors lean_run / k2_lean in gpusim_kernels.cu, one lane at a time) ----
struct LeanSlot {
    double x;
    uint32_t phase_word, gain_bits;
};
// low chip rate: chip boundaries per run the linear model (synth_lin) is instantiated for (0 = off) and the job-wide
// estimate of 1 / (f_code*delt); set per table by emu_generate exactly as gpusim_upload_table / plan_job do
static int g_lin_nb = 0;
static double g_lin_rinv = 0.0;
static long g_lin_runs = 0, g_fast_runs = 0; // (run, channel) pairs that took synth_lin / synth_fast_g since the last emu_generate

template <int FMT, int SR>
void emu_run_lean(const Tables &T, const std::vector<uint32_t> &win64, const DevRow *rows, int nc, LeanSlot *st,
                  int force_wrap, uint32_t lane_off, uint8_t *dst)
{
    typedef AccF32x2 A;
    const A::tab_t *lut = table_of<A>(T);
    A::acc_t acc[SR];
    for (int j = 0; j < SR; j++)
        acc[j] = A::init();
    for (int k = 0; k < nc; k++) {
        const DevRow &r = rows[k];
        double x = st[k].x;
        int c0;
        const double magic = floor_magic(x, c0);
        const bool wrap = c0 >= row_cthr(r) || force_wrap;
        const uint32_t *ww = win64.data() + (r.woff + (((uint32_t)c0 >> 5) << 3)) / 4; // {word i+1, word i}
        const uint32_t win = funnel_l_wrap(ww[0], ww[1], (uint32_t)c0);
        uint32_t phs = st[k].phase_word;
        if (!wrap) {
            // the device votes per warp; every path computes the same samples, so one lane may decide for itself
            ((g_lin_nb > 0 && lin_ok(c0, g_lin_nb)) ? g_lin_runs : g_fast_runs)++;
            if (g_lin_nb == 2 && lin_ok(c0, 2))
                synth_lin<A, SR, 2>(acc, x, phs, r.d, (uint32_t)r.steps, st[k].gain_bits, win, c0, g_lin_rinv, lut, lane_off);
            else if (g_lin_nb == 4 && lin_ok(c0, 4))
                synth_lin<A, SR, 4>(acc, x, phs, r.d, (uint32_t)r.steps, st[k].gain_bits, win, c0, g_lin_rinv, lut, lane_off);
            else
                synth_fast_g<A, SR>(acc, x, phs, r.d, (uint32_t)r.steps, st[k].gain_bits, win, magic, lut, lane_off);
            st[k].x = x;
            st[k].phase_word = phs;
        } else {
            ChanState cs;
            cs.x = x;
            cs.phs = phs;
            const int ic = lean_ic(phs);
            cs.bitk = ic / 20;
            cs.icode = ic - cs.bitk * 20;
            synth_wrap<A, SR>(acc, cs, r.d, (uint32_t)r.steps, r.gain, r.nav_bits, win, lut, lane_off);
            st[k].x = cs.x;
            st[k].phase_word = lean_phase_word(cs.phs, cs.bitk * 20 + cs.icode);
            st[k].gain_bits = A::gain_bits(data_sign(r.nav_bits, cs.bitk) * r.gain);
        }
    }
    store_run<A, FMT, SR>(dst, acc);
}

template <int FMT, int S>
void tuned_chunk_lean(const Tables &T, const DevRow *rows, int nc, const double *ckx, const uint16_t *ckw, int kc,
                      int jc, int chunk, int N, int force_wrap, int lane, uint8_t *epoch_out)
{
    typedef AccF32x2 A;
    static std::vector<uint32_t> win64;
    if (win64.empty()) {
        win64.resize((size_t)kCaPrns * kCaWin64 * 2);
        for (int i = 0; i < kCaPrns * kCaWin64; i++) {
            const int prn = i / kCaWin64, w = i - prn * kCaWin64;
            win64[2 * i] = T.negw[prn * kCaWords + w + 1];
            win64[2 * i + 1] = T.negw[prn * kCaWords + w];
        }
    }
    const int n0 = jc * chunk;
    const int nrun = std::min(chunk, N - n0);
    LeanSlot st[kMaxChan];
    for (int k = 0; k < nc; k++) {
        const int ic = rows[k].icode0 + ckw[k * kc + jc];
        st[k].x = ckx[k * kc + jc];
        st[k].phase_word = lean_phase_word(rows[k].ph0s + (uint32_t)n0 * (uint32_t)rows[k].steps, ic);
        st[k].gain_bits = A::gain_bits(data_sign(rows[k].nav_bits, ic / 20) * rows[k].gain);
    }
    constexpr int kBytesPer8 = (FMT == 16) ? 32 : (FMT == 8) ? 16 : 2;
    uint8_t *outp = epoch_out + (size_t)(n0 / 8) * kBytesPer8;
    const uint32_t lane_off = (uint32_t)(lane & A::kLaneMask) << A::kLaneShift;
    const int full = nrun / S, tail8 = (nrun - full * S) / 8;
    for (int i = 0; i < full; i++)
        emu_run_lean<FMT, S>(T, win64, rows, nc, st, force_wrap, lane_off, outp + (size_t)i * (S / 8) * kBytesPer8);
    for (int i = 0; i < tail8; i++)
        emu_run_lean<FMT, 8>(T, win64, rows, nc, st, force_wrap, lane_off, outp + ((size_t)full * (S / 8) + i) * kBytesPer8);
}

// ---- FLOAT_CARR_PHASE (double carrier phase) variants ---------------------------------------------
template <class A, int FMT, int SR>
void emu_run_f(const Tables &T, const DevRow *rows, const double *dcs, int nc, ChanStateF *st, uint32_t *meta,
               int force_wrap, uint32_t lane_off, uint8_t *dst)
{
    const typename A::tab_t *lut = table_of<A>(T);
    typename A::acc_t acc[SR];
    for (int j = 0; j < SR; j++)
        acc[j] = A::init();
    for (int k = 0; k < nc; k++) {
        const DevRow &r = rows[k];
        int c0;
        const double magic = floor_magic(st[k].x, c0);
        const bool wrap = c0 >= row_cthr(r) || force_wrap;
        const uint32_t *nw = T.negw.data() + (size_t)row_prn(r) * kCaWords;
        if (!wrap) {
            const uint32_t gbits = A::gain_bits(meta_sgain(meta[k])); // the device keeps these bits next to meta
            if (dcs[k] < 0.0)
                synth_fast_f<A, SR, true>(acc, st[k], r.d, dcs[k], gbits, chip_window(nw, c0), magic, lut, lane_off);
            else
                synth_fast_f<A, SR, false>(acc, st[k], r.d, dcs[k], gbits, chip_window(nw, c0), magic, lut, lane_off);
        } else {
            st[k].icode = meta_icode(meta[k]);
            st[k].bitk = meta_bitk(meta[k]);
            if (dcs[k] < 0.0)
                synth_wrap_f<A, SR, true>(acc, st[k], r.d, dcs[k], r.gain, r.nav_bits, chip_window(nw, (int)st[k].x), lut, lane_off);
            else
                synth_wrap_f<A, SR, false>(acc, st[k], r.d, dcs[k], r.gain, r.nav_bits, chip_window(nw, (int)st[k].x), lut, lane_off);
            meta[k] = pack_meta(st[k].icode, st[k].bitk, data_sign(r.nav_bits, st[k].bitk) * r.gain);
        }
    }
    store_run<A, FMT, SR>(dst, acc);
}

template <int FMT, int S>
void tuned_chunk_f(const Tables &T, const DevRow *rows, const double *dcs, int nc, const double *ckx,
                   const uint16_t *ckw, const double *ckc, int kc, int jc, int chunk, int N, int force_wrap,
                   int lane, uint8_t *epoch_out)
{
    typedef AccF32x2 A;
    const int n0 = jc * chunk;
    const int nrun = std::min(chunk, N - n0);
    ChanStateF st[kMaxChan];
    uint32_t meta[kMaxChan];
    for (int k = 0; k < nc; k++) {
        const int ic = rows[k].icode0 + ckw[k * kc + jc];
        const int bitk = ic / 20;
        st[k].x = ckx[k * kc + jc];
        st[k].cph = ckc[k * kc + jc];
        meta[k] = pack_meta(ic - bitk * 20, bitk, data_sign(rows[k].nav_bits, bitk) * rows[k].gain);
    }
    constexpr int kBytesPer8 = (FMT == 16) ? 32 : (FMT == 8) ? 16 : 2;
    uint8_t *outp = epoch_out + (size_t)(n0 / 8) * kBytesPer8;
    const uint32_t lane_off = (uint32_t)(lane & A::kLaneMask) << A::kLaneShift;
    const int full = nrun / S, tail8 = (nrun - full * S) / 8;
    for (int i = 0; i < full; i++)
        emu_run_f<A, FMT, S>(T, rows, dcs, nc, st, meta, force_wrap, lane_off, outp + (size_t)i * (S / 8) * kBytesPer8);
    for (int i = 0; i < tail8; i++)
        emu_run_f<A, FMT, 8>(T, rows, dcs, nc, st, meta, force_wrap, lane_off,
                             outp + ((size_t)full * (S / 8) + i) * kBytesPer8);
}

template <int FMT, bool CF>
void generic_chunk(const Tables &T, const DevRow *rows, const double *dcs, int nc, const double *ckx,
                   const uint16_t *ckw, const double *ckc, int kc, int jc, int chunk, int N, uint8_t *base)
{
    const int n0 = jc * chunk;
    const int nrun = std::min(chunk, N - n0);
    GenericChan ch[kMaxChan];
    for (int k = 0; k < nc; k++) {
        const DevRow &r = rows[k];
        const int ic = r.icode0 + ckw[k * kc + jc];
        ch[k].x = ckx[k * kc + jc];
        ch[k].d = r.d;
        ch[k].phs = r.ph0s + (uint32_t)n0 * (uint32_t)r.steps;
        ch[k].steps = r.steps;
        ch[k].cph = CF ? ckc[k * kc + jc] : 0.0;
        ch[k].dc = CF ? dcs[k] : 0.0;
        ch[k].gain = r.gain;
        ch[k].icode = ic % 20;
        ch[k].bitk = ic / 20;
        ch[k].nav_bits = r.nav_bits;
        ch[k].negw = T.negw.data() + (size_t)row_prn(r) * kCaWords;
    }
    uint32_t byte = 0;
    for (int n = 0; n < nrun; n++) {
        int i16, q16;
        generic_sample<CF>(ch, nc, T.s16.data(), T.c16.data(), i16, q16);
        const int s = n0 + n;
        if (FMT == 16) {
            const uint32_t w = ((uint32_t)i16 & 0xffffu) | ((uint32_t)q16 << 16);
            memcpy(base + 4 * (size_t)s, &w, 4);
        } else if (FMT == 8) {
            const uint16_t w = (uint16_t)(((uint32_t)(i16 >> 4) & 0xffu) | (((uint32_t)(q16 >> 4) & 0xffu) << 8));
            memcpy(base + 2 * (size_t)s, &w, 2);
        } else {
            byte = (byte << 2) | (i16 > 0 ? 2u : 0u) | (q16 > 0 ? 1u : 0u);
            if ((s & 3) == 3) {
                if ((s >> 2) < N / 4)
                    base[s >> 2] = (uint8_t)byte;
                byte = 0;
            }
        }
    }
}

} // namespace

extern "C" {

// kernel: 0 = tuned S=32, 1 = tuned S=16, 2 = generic.  Returns 0, or -1 if the table is outside
// the selected kernel's documented ranges.
int emu_generate(const gpusim_epoch_table *t, int N, double delt, int fmt, int chunk, int kernel,
                 int force_wrap, int chain_replay, int accum, int carrier_float, uint8_t *out)
{
    static const Tables T;
    // low chip rate: the rule of gpusim_upload_table / plan_job (accum == 1 is the lean kernel; accum == 5 is
    // the lean kernel with the linear path switched off, option lowrate=0)
    g_lin_nb = 0;
    g_lin_runs = g_fast_runs = 0;
    double dmin = 1e300, dmax = 0.0;
    for (size_t r = 0; r < (size_t)t->n_epochs * kMaxChan; r++)
        if (t->prn[r] > 0) {
            const double d = dmul(t->f_code[r], delt);
            dmin = std::min(dmin, d);
            dmax = std::max(dmax, d);
        }
    if (!carrier_float && accum == 1 && kernel == 0 && dmax > 0.0 && 32.0 * dmax < 4.0 && 5.0 * (dmax - dmin) < 0.25 * dmin * dmin) {
        g_lin_nb = 32.0 * dmax < 2.0 ? 2 : 4;
        g_lin_rinv = 2.0 / (dmin + dmax);
    }
    if (accum == 5)
        accum = 1;
    const size_t eb = fmt == 1 ? (size_t)(N / 4) : fmt == 8 ? (size_t)2 * N : (size_t)4 * N;
    const int kc = (N + chunk - 1) / chunk;
    // tuned kernels: chunks and epochs of a multiple of 8 samples (runs of 32/16 plus 8-sample tails)
    if ((kernel == 2 ? chunk % 32 != 0 : chunk % 8 != 0) || (kernel != 2 && N % 8 != 0))
        return -1;
    std::vector<double> ckx((size_t)kMaxChan * kc);
    std::vector<uint16_t> ckw((size_t)kMaxChan * kc);
    std::vector<double> ckc((size_t)kMaxChan * kc);
    for (int e = 0; e < t->n_epochs; e++) {
        DevRow rows[kMaxChan];
        double x0[kMaxChan];
        double dcs[kMaxChan] = {0}, cph0[kMaxChan] = {0};
        int nc = 0;
        for (int i = 0; i < kMaxChan; i++) {
            const size_t r = (size_t)e * kMaxChan + i;
            if (t->prn[r] <= 0)
                continue;
            DevRow &o = rows[nc];
            o.d = dmul(t->f_code[r], delt);
            o.steps = carrier_float ? 0 : (int32_t)((uint32_t)t->carr_phasestep[r] << 7);
            o.cthr_prn = pack_cthr_prn(o.d, t->prn[r]);
            if (g_lin_nb > 0 && chain_tie_binade(o.d) >= 5 && chain_tie_binade(o.d) <= 9)
                o.cthr_prn &= (uint16_t)~kCthrMask; // exact tie in a binade of the linear model: always the per-sample loop
            o.woff = (uint16_t)(t->prn[r] * kCaWinBytes);
            o.gain = t->gain[r];
            o.ph0s = carrier_float ? 0u : t->carr_phase[r] << 7;
            if (carrier_float) {
                dcs[nc] = dmul(t->f_carr[r], delt) * 512.0;
                cph0[nc] = t->carr_phase_f[r] * 512.0;
            }
            o.nav_bits = t->nav_bits[r];
            o.icode0 = (uint16_t)t->icode[r];
            o.flags = 0;
            if (kernel != 2 && (o.gain < 0 || o.gain > kTunedMaxGain || o.d > (kernel == 0 ? 0.9999 : 2.0)))
                return -1;
            x0[nc++] = t->code_phase[r];
        }
        for (int k = 0; k < nc; k++) {
            double *cx = ckx.data() + (size_t)k * kc;
            uint16_t *cw = ckw.data() + (size_t)k * kc;
            auto emit = [&](int j, double x, int wraps) {
                cx[j] = x;
                cw[j] = (uint16_t)wraps;
            };
            if (chain_replay)
                code_chain_replay(x0[k], rows[k].d, N, chunk, emit);
            else
                code_chain(x0[k], rows[k].d, N, chunk, emit);
            if (carrier_float) {
                double *cc = ckc.data() + (size_t)k * kc;
                auto emit_c = [&](int j, double x, int) { cc[j] = x; };
                const int last = ((N - 1) / chunk) * chunk;
                if (chain_replay) {
                    double x = cph0[k];
                    for (int n = 0; n <= last; n++) {
                        if (n % chunk == 0)
                            cc[n / chunk] = x;
                        x = carrier_step(x, dcs[k]);
                    }
                } else {
                    { ChainTabHost tab; carrier_chain(cph0[k], dcs[k], last, chunk, tab, emit_c); }
                }
            }
        }
        uint8_t *eo = out + (size_t)e * eb;
        for (int jc = 0; jc < kc; jc++) {
            const int lane = jc & 31;
#define EMU_TUNED(F, S)                                                                                          \
    do {                                                                                                         \
        if (carrier_float) tuned_chunk_f<F, S>(T, rows, dcs, nc, ckx.data(), ckw.data(), ckc.data(), kc, jc, chunk, N, force_wrap, lane, eo); \
        else if (accum == 1) tuned_chunk_lean<F, S>(T, rows, nc, ckx.data(), ckw.data(), kc, jc, chunk, N, force_wrap, lane, eo); \
        else if (accum == 3) tuned_chunk<AccF32x2, F, S>(T, rows, nc, ckx.data(), ckw.data(), kc, jc, chunk, N, force_wrap, lane, eo); \
        else tuned_chunk<AccWide, F, S>(T, rows, nc, ckx.data(), ckw.data(), kc, jc, chunk, N, force_wrap, lane, eo);            \
    } while (0)
#define EMU_GENERIC(F)                                                                                           \
    do {                                                                                                         \
        if (carrier_float) generic_chunk<F, true>(T, rows, dcs, nc, ckx.data(), ckw.data(), ckc.data(), kc, jc, chunk, N, eo); \
        else generic_chunk<F, false>(T, rows, dcs, nc, ckx.data(), ckw.data(), ckc.data(), kc, jc, chunk, N, eo); \
    } while (0)
            if (kernel == 0) {
                if (fmt == 16) EMU_TUNED(16, 32); else if (fmt == 8) EMU_TUNED(8, 32); else EMU_TUNED(1, 32);
            } else if (kernel == 1) {
                if (fmt == 16) EMU_TUNED(16, 16); else if (fmt == 8) EMU_TUNED(8, 16); else EMU_TUNED(1, 16);
            } else {
                if (fmt == 16) EMU_GENERIC(16); else if (fmt == 8) EMU_GENERIC(8); else EMU_GENERIC(1);
            }
#undef EMU_GENERIC
#undef EMU_TUNED
        }
    }
    return 0;
}

// One run of 32 samples of one channel through synth_lin<NB> and through synth_fast_g from the same state:
// returns 1 when the accumulators, the code phase and the carrier phase after the run agree bit for bit.
// x_end[0] / x_end[1]: the code phase after the run (linear model / per-sample chain).
int emu_lin_matches_fast(double x0, double d, uint32_t phs0, int32_t steps, int prn, int nb, double rinv, double *x_end)
{
    static const Tables T;
    typedef AccF32x2 A;
    const A::tab_t *lut = table_of<A>(T);
    int c0;
    const double magic = floor_magic(x0, c0);
    const uint32_t win = chip_window(T.negw.data() + (size_t)prn * kCaWords, c0);
    const uint32_t gb = A::gain_bits(77);
    A::acc_t a[32], b[32];
    for (int j = 0; j < 32; j++)
        a[j] = b[j] = A::init();
    double xa = x0, xb = x0;
    uint32_t pa = phs0, pb = phs0;
    if (nb == 2)
        synth_lin<A, 32, 2>(a, xa, pa, d, (uint32_t)steps, gb, win, c0, rinv, lut, 8u);
    else
        synth_lin<A, 32, 4>(a, xa, pa, d, (uint32_t)steps, gb, win, c0, rinv, lut, 8u);
    synth_fast_g<A, 32>(b, xb, pb, d, (uint32_t)steps, gb, win, magic, lut, 8u);
    x_end[0] = xa;
    x_end[1] = xb;
    return memcmp(a, b, sizeof(a)) == 0 && dbits(xa) == dbits(xb) && pa == pb;
}

// how many (run, channel) pairs of the last emu_generate call took the linear low-chip-rate path / the per-sample loop
long emu_lin_runs(void) { return g_lin_runs; }
long emu_fast_runs(void) { return g_fast_runs; }

// the chain walk alone: checkpoints every `every` samples
void emu_code_chain(double x0, double d, int n_total, int every, int replay, double *x_out, int *w_out)
{
    auto emit = [&](int j, double x, int wraps) {
        x_out[j] = x;
        w_out[j] = wraps;
    };
    if (replay)
        code_chain_replay(x0, d, n_total, every, emit);
    else
        code_chain(x0, d, n_total, every, emit);
}

// the generic walk:  x += d; wrap into [0, M)  for n_end steps; checkpoints every `every` steps;
// returns the final value
double emu_phase_chain(double x0, double d, double M, int n_end, int every, double *x_out, int *w_out)
{
    auto emit = [&](int j, double x, int wraps) {
        x_out[j] = x;
        w_out[j] = wraps;
    };
    return phase_chain(x0, d, M, n_end, every, emit);
}

// the same walk through the compile-time-sign instantiations the device kernel and the host's
// gpusim_advance_carrier_f64 use (phase_chain<1> for d >= 0, phase_chain<-1> for d < 0)
double emu_phase_chain_signed(double x0, double d, double M, int n_end, int every, double *x_out, int *w_out)
{
    auto emit = [&](int j, double x, int wraps) {
        x_out[j] = x;
        w_out[j] = wraps;
    };
    return d < 0.0 ? phase_chain<-1>(x0, d, M, n_end, every, emit) : phase_chain<1>(x0, d, M, n_end, every, emit);
}

// the tabulated walk (phase_chain_tab): what k1_chain runs on the device and gpusim_advance_carrier_f64 on the host; M <= 1024
double emu_phase_chain_tab(double x0, double d, double M, int n_end, int every, double *x_out, int *w_out)
{
    auto emit = [&](int j, double x, int wraps) {
        x_out[j] = x;
        w_out[j] = wraps;
    };
    ChainTabHost tab;
    return d < 0.0 ? phase_chain_tab<-1>(x0, d, M, n_end, every, tab, emit) : phase_chain_tab<1>(x0, d, M, n_end, every, tab, emit);
}

// K0 on the host: the 60 data words of n frames (k0_navmsg) and the 32 data bits of a row (k0_navbits)
void emu_nav_build(const gpusim_nav_frame *frames, int n, uint32_t *dwrd)
{
    static_assert(sizeof(gpusim_nav_frame) == sizeof(NavFrame), "layout");
    for (int f = 0; f < n; f++) {
        NavFrame nf;
        memcpy(&nf, &frames[f], sizeof(nf));
        for (int sub = 0; sub < kNavSubframes; sub++)
            nav_build_subframe(nf, sub, dwrd + (size_t)f * kNavWords);
    }
}
uint32_t emu_nav_row_bits(const uint32_t *dwrd60, int iword, int ibit) { return nav_row_bits(dwrd60, iword, ibit); }
uint32_t emu_nav_word(uint32_t src, int solve_tail) { return nav_word(src, solve_tail != 0); }

// k0_eph2sbf on the host: the 50 source words of n ephemerides
void emu_eph2sbf(const gpusim_nav_eph *eph, int n, const gpusim_nav_iono *iono, uint32_t *sbf)
{
    static_assert(sizeof(gpusim_nav_eph) == sizeof(NavEph) && sizeof(gpusim_nav_iono) == sizeof(NavIono), "layout");
    NavIono io;
    memcpy(&io, iono, sizeof(io));
    for (int i = 0; i < n; i++) {
        NavEph e;
        memcpy(&e, &eph[i], sizeof(e));
        nav_eph_subframes(e, io, sbf + (size_t)i * kNavSbfWords);
    }
}

} // extern "C"
