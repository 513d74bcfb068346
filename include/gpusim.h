/*
 * gpusim.h - C ABI of the B200-native GPS L1 C/A baseband generator.
 *
 * This is the drop-in boundary for ONE path of gps-sdr-sim: the per-sample IQ
 * synthesis loop and output quantise/pack of the reference
 * (gpssim.c:2190-2264 and gpssim.c:2266-2288).  Everything above it - RINEX
 * parsing, orbit propagation, computeRange/computeCodePhase, eph2sbf,
 * allocateChannel, the CLI - stays the reference's own host C code
 * (generateNavMsg too, unless the host opts for gpusim_nav_build below).  The host
 * records, for every 0.1 s epoch and channel slot, the state the reference's
 * sample loop would have started from (one row of gpusim_epoch_table) and
 * hands batches of epochs to gpusim_generate_epochs*(), which returns exactly
 * the bytes the reference's fwrite calls (gpssim.c:2276 / :2283 / :2287) would
 * have produced for those epochs, in epoch order.
 *
 * The reference has no plugin / FFI interface for this path (the loop is inlined
 * in main()), so every entry point below cites the reference lines it replaces.
 * INTEGRATION.md shows the few lines a maintainer adds to gpssim.c to bind it.
 *
 * Plain C, plain pointers and sizes.  No CPU fallback exists: every compute
 * entry point fails with GPUSIM_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef GPUSIM_H
#define GPUSIM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPUSIM_ABI_VERSION 2 /* 2: gpusim_epoch_table grew nav_frame / iword / ibit; gpusim_nav_build */

/* gpssim.h:16  MAX_CHAN - channel slots per epoch row */
#define GPUSIM_MAX_CHAN 16
/* gpssim.h:36  CA_SEQ_LEN */
#define GPUSIM_CA_SEQ_LEN 1023

/* gpssim.h:77-79  SC01 / SC08 / SC16 */
#define GPUSIM_SC01 1
#define GPUSIM_SC08 8
#define GPUSIM_SC16 16

/* gpssim.h:4  which carrier-phase branch the host was compiled with */
#define GPUSIM_CARRIER_INT 0   /* unsigned carr_phase + int carr_phasestep (gpssim.c:2202,:2252) */
#define GPUSIM_CARRIER_FLOAT 1 /* double carr_phase, FLOAT_CARR_PHASE (gpssim.c:2200,:2245-2250) */

/* status codes (0 = ok).  The reference's convention is "ERROR: ..." on stderr
 * and exit(1) (e.g. gpssim.c:2076-2108); the host shim does that on non-zero. */
#define GPUSIM_OK 0
#define GPUSIM_ERR_ARG 1      /* bad argument / table violates a documented range */
#define GPUSIM_ERR_CUDA 2     /* CUDA runtime error, or no usable device */
#define GPUSIM_ERR_CAPACITY 3 /* batch larger than the context was created for */
#define GPUSIM_ERR_SINK 4     /* the caller's sink returned non-zero */
#define GPUSIM_ERR_UNSUPPORTED 5

typedef struct gpusim_ctx gpusim_ctx;

/* Scalars the reference fixes before its epoch loop. */
typedef struct gpusim_config {
    int32_t abi_version;       /* GPUSIM_ABI_VERSION */
    int32_t device;            /* CUDA device ordinal */
    int32_t samples_per_epoch; /* iq_buff_size = floor(fs/10)      gpssim.c:1877-1878 */
    int32_t data_format;       /* GPUSIM_SC01/08/16 (-b)           gpssim.c:1797 */
    int32_t carrier_mode;      /* GPUSIM_CARRIER_INT / _FLOAT      gpssim.h:4 */
    int32_t max_batch_epochs;  /* capacity of one generate call */
    double delt;               /* 1/(10*iq_buff_size)              gpssim.c:1881 */
} gpusim_config;

/*
 * One batch of epochs, structure-of-arrays, host memory, row index
 * e*GPUSIM_MAX_CHAN + slot.  A row is the value of the reference's per-channel
 * state at the top of the sample loop (gpssim.c:2190) for that epoch, i.e.
 * after the refresh at gpssim.c:2156-2188:
 *
 *   prn            chan[i].prn; 0 = slot not allocated (gpssim.c:2197 skips it)
 *   f_code         chan[i].f_code                       (gpssim.c:1328)
 *   code_phase     chan[i].code_phase, chips in [0,1023) (gpssim.c:1334)
 *   icode          chan[i].icode, 0..19                  (gpssim.c:1342)
 *   nav_bits       the next 32 navigation data bits starting at
 *                  (chan[i].iword, chan[i].ibit), MSB first:
 *                  bit 31 = (dwrd[iword]>>(29-ibit))&1   (gpssim.c:1345),
 *                  bit 30 = the following bit, ... (an epoch consumes <= 7)
 *   gain           gain[i]                               (gpssim.c:2186)
 *   carr_phasestep chan[i].carr_phasestep                (gpssim.c:2176)   INT mode
 *   carr_phase     chan[i].carr_phase at epoch start (the host advances it by
 *                  samples_per_epoch*carr_phasestep mod 2^32 per epoch - the
 *                  sample loop used to do that at gpssim.c:2252)           INT mode
 *   f_carr         chan[i].f_carr                        (gpssim.c:1327)   FLOAT mode
 *   carr_phase_f   chan[i].carr_phase (double) at epoch start              FLOAT mode
 *
 *   nav_frame, iword, ibit   (optional) instead of nav_bits: the row's
 *                  data bits are taken on the device from frame nav_frame of the last
 *                  gpusim_nav_build() call, starting at (chan[i].iword, chan[i].ibit)
 *                  (gpssim.c:1343-1345).  Used when nav_bits is NULL.
 *
 * Arrays of the other carrier mode may be NULL.  The library copies what it
 * needs before returning; no pointer is retained.
 */
typedef struct gpusim_epoch_table {
    int32_t n_epochs;
    const int32_t *prn;
    const double *f_code;
    const double *code_phase;
    const int32_t *icode;
    const uint32_t *nav_bits;
    const int32_t *gain;
    const int32_t *carr_phasestep;
    const uint32_t *carr_phase;
    const double *f_carr;
    const double *carr_phase_f;
    const int32_t *nav_frame;
    const int32_t *iword;
    const int32_t *ibit;
} gpusim_epoch_table;

/*
 * SURVEY 8 f4 - navigation data words built on the device.
 *
 * One gpusim_nav_frame is the input of ONE generateNavMsg() call of the reference
 * (gpssim.c:1467-1547): the satellite's five subframes as eph2sbf() left them in
 * chan->sbf (gpssim.c:490-665: 24 source bits per word in bits 29..6, no TOW count,
 * no week number, no parity), and the time the frame starts at.  The device adds the
 * TOW counts and the week number and computes the parity of all 60 words
 * (computeChecksum, gpssim.c:693-756), i.e. what the reference keeps in chan->dwrd[60]
 * (gpssim.h:175):
 *
 *   words  0.. 9  subframe 5 of the frame before: built from `first` with TOW count
 *                 tow_first.  generateNavMsg(init=1) builds them from chan->sbf[4] and tow
 *                 (gpssim.c:1484-1503): first = sbf[4], tow_first = tow.  On a refresh
 *                 (init=0, gpssim.c:1504-1511) the reference copies words 50..59 of the
 *                 previous frame: first = the PREVIOUS call's sbf[4] (it differs from the
 *                 current one after an ephemeris-set switch, gpssim.c:2318-2330),
 *                 tow_first = previous tow + 5.
 *   words 10..59  subframes 1..5 from sbf[0..4] with TOW counts tow+1 .. tow+5, week number
 *                 wn in word 3 of subframe 1 (gpssim.c:1517-1543).
 *
 *   tow = ((unsigned long)g0.sec)/6, wn = g0.week%1024 with g0 as at gpssim.c:1476-1481.
 */
typedef struct gpusim_nav_frame {
    uint32_t sbf[5][10];
    uint32_t first[10];
    uint32_t tow_first;
    uint32_t tow;
    uint32_t wn;
    uint32_t reserved; /* 0 */
} gpusim_nav_frame;

/* Ordered output sink: called with consecutive byte ranges of the output file
 * (what fwrite(..., fp) received at gpssim.c:2276/:2283/:2287).  Non-zero aborts. */
typedef int (*gpusim_sink_fn)(void *user, const void *bytes, size_t n_bytes);

/* Device timing of the last generate call (CUDA events on the library's streams). */
typedef struct gpusim_timing {
    float chain_ms;      /* code-phase checkpoint kernel(s) */
    float synth_ms;      /* generate + quantise + pack kernel(s) */
    float total_ms;      /* first kernel start to last kernel end */
    int32_t launches;    /* kernels launched by the call */
    int32_t fast_path;   /* 1 = tuned kernel, 0 = generic kernel was needed */
    int32_t chain_overlapped; /* 1 = the chain kernel ran on the library's own stream beside the previous call's synthesis kernel */
} gpusim_timing;

int gpusim_abi_version(void);
const char *gpusim_strerror(int status);
/* text of the last failure on this context (or of gpusim_create when ctx==NULL) */
const char *gpusim_last_error(const gpusim_ctx *ctx);

int gpusim_create(const gpusim_config *cfg, gpusim_ctx **out_ctx);
void gpusim_destroy(gpusim_ctx *ctx);

/* bytes written per epoch: 4N (SC16), 2N (SC08), N/4 (SC01)   gpssim.c:2276,:2283,:2287 */
size_t gpusim_epoch_bytes(const gpusim_ctx *ctx);

/*
 * Replaces gpssim.c:2190-2288 for table->n_epochs consecutive epochs.
 * Host table in, host bytes out (out may be pageable); synchronous.
 */
int gpusim_generate_epochs(gpusim_ctx *ctx, const gpusim_epoch_table *table,
                           void *out, size_t out_capacity);

/*
 * Same, but the bytes are delivered in order to `sink` from pinned staging
 * buffers while later epochs are still being generated / copied
 * (double-buffered cudaMemcpyAsync overlapped with the caller's fwrite).
 */
int gpusim_generate_epochs_to_sink(gpusim_ctx *ctx, const gpusim_epoch_table *table,
                                   gpusim_sink_fn sink, void *user);

/*
 * Device-resident variant used for kernel timing and by callers that keep the
 * samples on the GPU: upload once, then generate any epoch sub-range
 * [first_epoch, first_epoch+n_epochs) of the uploaded table into device memory.
 * `stream` is a cudaStream_t (NULL = the library's own stream); the call is
 * asynchronous with respect to the host when a stream is given: the output of
 * call i is complete when `stream` reaches the point after call i.  (Option
 * "pipeline" >= 1 moves the code-phase chain kernel of a call to a library
 * stream, where it may start while the previous call's synthesis kernel drains.)
 */
int gpusim_upload_table(gpusim_ctx *ctx, const gpusim_epoch_table *table);
int gpusim_generate_device(gpusim_ctx *ctx, int32_t first_epoch, int32_t n_epochs,
                           void *out_device, size_t out_capacity, void *stream);

int gpusim_get_timing(const gpusim_ctx *ctx, gpusim_timing *out);

/*
 * Replaces generateNavMsg() + computeChecksum() (gpssim.c:1467-1547, :693-756) for n_frames
 * frames: their 60 data words each are built on the device and stay there; tables uploaded or
 * generated afterwards may reference them by index (nav_frame / iword / ibit).  A later call
 * replaces the whole set.  gpusim_nav_read() copies words of frames [first, first+n) back
 * (60 uint32 per frame) - for tests and for hosts that want to cross-check chan->dwrd.
 */
int gpusim_nav_build(gpusim_ctx *ctx, const gpusim_nav_frame *frames, int32_t n_frames);
int gpusim_nav_read(gpusim_ctx *ctx, int32_t first_frame, int32_t n_frames, uint32_t *dwrd);

/*
 * The same with eph2sbf() (gpssim.c:490-665) on the device as well: the host hands over the broadcast
 * ephemerides themselves - the fields of ephem_t / ionoutc_t (gpssim.h:101-146) that eph2sbf() reads, as
 * readRinexNavAll() filled them - and frames name their subframes by ephemeris index:
 *   eph        index of the ephemeris whose subframes chan->sbf held when generateNavMsg() ran
 *   eph_first  index of the ephemeris behind the frame's first ten words (see gpusim_nav_frame.first)
 * gpusim_nav_read_sbf() copies the 50 source words (chan->sbf) of ephemerides [first, first+n) back.
 */
typedef struct gpusim_nav_eph {
    double toe_sec, toc_sec; /* eph.toe.sec, eph.toc.sec */
    double deltan, cuc, cus, cic, cis, crc, crs, ecc, sqrta, m0, omg0, inc0, aop, omgdot, idot, af0, af1, af2, tgd;
    int32_t toe_week, iodc, iode, svhlth, codeL2, reserved;
} gpusim_nav_eph;
typedef struct gpusim_nav_iono {
    double alpha0, alpha1, alpha2, alpha3, beta0, beta1, beta2, beta3, A0, A1;
    int32_t vflg, dtls, tot, wnt;
} gpusim_nav_iono;
typedef struct gpusim_nav_frame_ref {
    int32_t eph, eph_first;
    uint32_t tow_first, tow, wn, reserved;
} gpusim_nav_frame_ref;
int gpusim_nav_build_eph(gpusim_ctx *ctx, const gpusim_nav_eph *eph, int32_t n_eph, const gpusim_nav_iono *iono,
                         const gpusim_nav_frame_ref *frames, int32_t n_frames);
int gpusim_nav_read_sbf(gpusim_ctx *ctx, int32_t first_eph, int32_t n_eph, uint32_t *sbf);

/* Test facility.  A context created with GPUSIM_GUARD=1 in the environment places every device buffer it
 * owns (rows, code-phase checkpoints, work counters, its own output buffer) between two 4 KiB poisoned guard
 * bands.  Returns the number of guard bytes that no longer hold the poison after all work of the context has
 * finished (0 = no kernel wrote outside its buffers), -1 if the context was created without guards, -2 on a
 * CUDA error.  The caller's own output buffer (gpusim_generate_device) is the caller's to guard. */
int64_t gpusim_debug_guard_violations(gpusim_ctx *ctx);

/* Tuning / test hooks (all optional).  key/value pairs documented in DESIGN.md:
 *   "chunk"  samples per thread chunk (multiple of 32), 0 = auto
 *   "force_generic" 1 = always use the generic exact kernel
 *   "force_slow" 1 = always take the wrap-checking inner loop
 *   "direct_first_mb", "direct_mb" sub-batch sizes (MiB) of gpusim_generate_epochs (default 16, 64)
 *   "pipeline" 0 = (default) chain kernel in front of its synthesis kernel on the caller's stream; 1 = on the
 *              library's chain stream when the previous call is still running; 2 = always on that stream
 *   "lean"     integer carrier: 1 = (default) k2_lean, 0 = the round-1 kernel k2_synth (cross-check)
 *   "lowrate"  1 = (default) tables with >= 8 samples per chip take the linear-model path for chip boundaries, 0 = never
 *   "accum", "layout", "chain_replay"  cross-check variants, see DESIGN.md 5
 *   "float_geom" FLOAT hosts: 0 = (default) 512-thread kernel when <= 13 channels are active, else the
 *              384-thread one; 1 = always the 384-thread kernel */
int gpusim_set_option(gpusim_ctx *ctx, const char *key, int64_t value);

/*
 * Constant tables of the path, as the device uses them, for parity tests:
 *   sin512/cos512  -> sinTable512 / cosTable512  (gpssim.c:15-83), 512 ints each
 *   ca             -> codegen(ca, prn)           (gpssim.c:132-171), 1023 ints in {0,1}
 * Host-side, no GPU needed.
 */
void gpusim_carrier_lut(int32_t *sin512, int32_t *cos512);
int gpusim_ca_code(int32_t prn, int32_t *ca1023);

/* Pack the 32 data bits starting at (iword, ibit) of a dwrd[60] word buffer
 * (gpssim.h:175, unsigned long on the host) into the nav_bits row format. */
uint32_t gpusim_pack_nav_bits(const unsigned long *dwrd, int32_t n_dwrd, int32_t iword, int32_t ibit);

/* Page-locked host memory for the `out` buffer of gpusim_generate_epochs (device->host copies into
 * it run at full PCIe rate and overlap generation).  gpusim_host_free(NULL) is a no-op. */
void *gpusim_host_alloc(size_t n_bytes);
void gpusim_host_free(void *p);

/*
 * FLOAT_CARR_PHASE hosts only: the value of chan[i].carr_phase after n_samples executions of
 * "carr_phase += f_carr*delt; wrap into [0,1)" (gpssim.c:2245-2250) - what the removed sample loop
 * left behind for the next epoch - computed exactly in O(carrier cycles).  Host-side, no GPU needed.
 */
double gpusim_advance_carrier_f64(double carr_phase, double f_carr, double delt, int32_t n_samples);

#ifdef __cplusplus
}
#endif
#endif /* GPUSIM_H */
