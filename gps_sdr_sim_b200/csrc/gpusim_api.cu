// gpusim_api.cu - the C ABI of include/gpusim.h: context, table upload, batching,
// double-buffered device->host streaming.  The kernels are in gpusim_kernels.cu.
//
// There is deliberately no CPU path: every compute entry point needs a CUDA device.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "gpusim.h"
#include "gpusim_kernels.h"
#include "gpusim_tables.h"

using namespace gpusim;

namespace {

thread_local std::string g_create_error;

constexpr size_t kStageBytes = 64u << 20;      // one pinned staging buffer / sub-batch of output
constexpr size_t kCheckpointBudget = 1u << 29; // device bytes for ONE set of K1 checkpoints (there are two)

} // namespace

struct gpusim_ctx {
    gpusim_config cfg{};
    size_t epoch_bytes = 0;
    std::string err;

    cudaStream_t s_compute = nullptr, s_copy = nullptr, s_chain = nullptr;
    // Two checkpoint sets ("slots") alternate between consecutive generate calls, so that the chain
    // kernel of call i+1 (on s_chain) can run while the synthesis kernel of call i is still busy.
    // Per slot: chain start/end, synthesis start/end (the last one also guards the slot's reuse).
    cudaEvent_t ev_c0[2] = {nullptr, nullptr}, ev_c1[2] = {nullptr, nullptr};
    cudaEvent_t ev_s0[2] = {nullptr, nullptr}, ev_s1[2] = {nullptr, nullptr};
    unsigned seq = 0;
    int last_slot = 0;
    cudaEvent_t ev_done[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};

    // constant tables
    int32_t *d_lut = nullptr;
    uint64_t *d_lut_f32 = nullptr;
    int16_t *d_sin16 = nullptr, *d_cos16 = nullptr;
    uint32_t *d_negw = nullptr;

    // uploaded table (capacity cfg.max_batch_epochs)
    DevRow *d_rows = nullptr, *h_rows = nullptr;
    uint8_t *d_nch = nullptr, *h_nch = nullptr;
    double *d_x0 = nullptr, *h_x0 = nullptr;
    int n_uploaded = 0;
    bool needs_generic = false;
    double d_max = 0.0;
    double d_min = 0.0;   // smallest f_code*delt of the uploaded table (low-chip-rate path, synth_lin)
    int lin_nb = 0;        // > 0: the uploaded table qualifies for synth_lin with this many chip boundaries per run

    // checkpoints, sized for min_chunk
    double *d_ck_x[2] = {nullptr, nullptr};
    uint16_t *d_ck_w[2] = {nullptr, nullptr};
    unsigned int *d_work = nullptr;
    // FLOAT_CARR_PHASE hosts: 512*RN(f_carr*delt), 512*carr_phase at epoch start, carrier checkpoints
    double *d_dc = nullptr, *h_dc = nullptr, *d_cph0 = nullptr, *h_cph0 = nullptr, *d_ck_c[2] = {nullptr, nullptr};
    int sm_count = 148;
    int min_chunk = 128;

    // navigation frames built on the device (gpusim_nav_build): requests and their 60 words each
    NavFrame *d_nav_req = nullptr;
    uint32_t *d_nav_words = nullptr;
    int nav_capacity = 0, nav_frames = 0;
    NavEph *d_nav_eph = nullptr;    // gpusim_nav_build_eph: ephemerides and the 50 subframe words k0_eph2sbf makes of each
    uint32_t *d_nav_sbf = nullptr;
    int nav_eph_capacity = 0, nav_ephs = 0;

    // output
    uint8_t *d_out = nullptr; // lazily: max_batch_epochs * epoch_bytes
    uint8_t *h_stage[2] = {nullptr, nullptr};

    // options
    int opt_chunk = 0, opt_force_generic = 0, opt_force_slow = 0, opt_chain_replay = 0, opt_accum = 1, opt_layout = 0, opt_pipeline = 0, opt_float_geom = 0, opt_lean = 1, opt_lowrate = 1;
    int opt_direct_first_mb = 16, opt_direct_mb = 64; // sub-batch sizes when copying straight into the caller's buffer

    gpusim_timing timing{};

    // every device allocation of the context; with GPUSIM_GUARD=1 in the environment each one sits between
    // two poisoned guard bands that gpusim_debug_guard_violations() checks (memory-safety evidence for the
    // kernels' writes where compute-sanitizer is not available)
    struct DevAlloc { unsigned char *base; unsigned char *user; size_t bytes; };
    std::vector<DevAlloc> allocs;
    bool guard = false;
};

namespace {

int fail(gpusim_ctx *ctx, int status, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx)
        ctx->err = buf;
    else
        g_create_error = buf;
    return status;
}

constexpr size_t kGuardBytes = 4096;
constexpr int kGuardPoison = 0xA5;

template <class T>
cudaError_t dev_alloc(gpusim_ctx *ctx, T **out, size_t bytes)
{
    bytes = std::max<size_t>(bytes, 16);
    const size_t pad = ctx->guard ? kGuardBytes : 0;
    unsigned char *base = nullptr;
    cudaError_t e = cudaMalloc(&base, bytes + 2 * pad);
    if (e != cudaSuccess)
        return e;
    if (pad) {
        e = cudaMemset(base, kGuardPoison, pad);
        if (e == cudaSuccess)
            e = cudaMemset(base + pad + bytes, kGuardPoison, pad);
        if (e != cudaSuccess) {
            cudaFree(base);
            return e;
        }
    }
    ctx->allocs.push_back({base, base + pad, bytes});
    *out = reinterpret_cast<T *>(base + pad);
    return cudaSuccess;
}

#define GS_CUDA(ctx, call)                                                                     \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess)                                                                 \
            return fail(ctx, GPUSIM_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

int kc_for(int n_samples, int chunk) { return (n_samples + chunk - 1) / chunk; }

// samples per thread chunk.  K2 runs one persistent 512-thread block per SM whose warps claim units
// of 32 chunks; pick the largest chunk that still leaves every warp >= 16 units to balance with.
int pick_chunk(const gpusim_ctx *ctx, int n_epochs)
{
    if (ctx->opt_chunk > 0)
        return std::max(ctx->min_chunk, (ctx->opt_chunk + 31) / 32 * 32);
    const long long want = 16LL * 32 * ctx->sm_count * (synth_threads() / 32);
    int chunk = 1024;
    while (chunk > ctx->min_chunk && chunk > 128 && (long long)n_epochs * kc_for(ctx->cfg.samples_per_epoch, chunk) < want)
        chunk /= 2;
    return std::max(chunk, ctx->min_chunk);
}

int ensure_out(gpusim_ctx *ctx)
{
    if (ctx->d_out == nullptr)
        GS_CUDA(ctx, dev_alloc(ctx, &ctx->d_out, (size_t)ctx->cfg.max_batch_epochs * ctx->epoch_bytes));
    return GPUSIM_OK;
}

int ensure_stage(gpusim_ctx *ctx)
{
    for (int i = 0; i < 2; i++)
        if (ctx->h_stage[i] == nullptr)
            GS_CUDA(ctx, cudaMallocHost(&ctx->h_stage[i], std::max(kStageBytes, ctx->epoch_bytes)));
    return GPUSIM_OK;
}

// describe the work for uploaded epochs [first, first+n): layout, chunking, kernel choice
SynthKernel plan_job(gpusim_ctx *ctx, int first, int n, uint8_t *out_dev, DeviceJob &job, int slot)
{
    job = DeviceJob{};
    job.rows = ctx->d_rows + (size_t)first * kMaxChan;
    job.nch = ctx->d_nch + first;
    job.x0 = ctx->d_x0 + (size_t)first * kMaxChan;
    job.ck_x = ctx->d_ck_x[slot];
    job.ck_w = ctx->d_ck_w[slot];
    job.carrier_float = ctx->cfg.carrier_mode == GPUSIM_CARRIER_FLOAT ? 1 : 0;
    job.dc = ctx->d_dc ? ctx->d_dc + (size_t)first * kMaxChan : nullptr;
    job.cph0 = ctx->d_cph0 ? ctx->d_cph0 + (size_t)first * kMaxChan : nullptr;
    job.ck_c = ctx->d_ck_c[slot];
    job.lut_wide = ctx->d_lut;
    job.lut_f32 = ctx->d_lut_f32;
    job.accum = ctx->opt_accum;
    job.sin16 = ctx->d_sin16;
    job.cos16 = ctx->d_cos16;
    job.negw = ctx->d_negw;
    job.out = out_dev;
    job.work_counter = ctx->d_work + 8 * slot;
    job.sm_count = ctx->sm_count;
    job.n_epochs = n;
    job.n_samples = ctx->cfg.samples_per_epoch;

    SynthKernel which = SynthKernel::Tuned32;
    if (ctx->opt_force_generic || ctx->needs_generic || (job.n_samples % 8) != 0 || ctx->d_max >= 2.0)
        which = SynthKernel::Generic;
    else if (ctx->d_max > 0.9999)
        which = SynthKernel::Tuned16;

    // Work layout of the tuned kernel.  Aligned: an epoch is 100 nominal C/A code periods; a chunk is
    // 1/q of a period and the 32 lanes of a unit take the same slot of 32 consecutive periods, so
    // the 1023-chip wrap reaches all lanes of a warp in the same run (see k2_synth).  Needs the
    // period to be a whole number of samples and the chunk a multiple of 8 samples; otherwise
    // (or with option layout=1) chunks are a power of two and units are 32 consecutive chunks.
    job.ppe = 0;
    job.q = 0;
    job.chunk = pick_chunk(ctx, n);
    const long long want_units = 8LL * ctx->sm_count * (synth_threads() / 32);
    if (which != SynthKernel::Generic && ctx->opt_layout != 1 && job.n_samples % 100 == 0) {
        const int period = job.n_samples / 100;
        const int floor_chunk = std::max(ctx->min_chunk, 64);
        int best_q = 0;
        for (int q = 1; q <= period; q++) {
            if (period % q != 0 || (period / q) % 8 != 0 || period / q < floor_chunk || period / q > 4096)
                continue;
            best_q = q; // smallest q (largest chunk) with enough units; else the largest admissible q
            if (((long long)n * 100 + 31) / 32 * q >= want_units)
                break;
        }
        if (best_q > 0) {
            job.ppe = 100;
            job.q = best_q;
            job.chunk = period / best_q;
        }
    }
    job.kc = kc_for(job.n_samples, job.chunk);
    job.n_units = job.ppe > 0 ? (int32_t)((((long long)n * job.ppe + 31) / 32) * job.q)
                              : (int32_t)(((long long)n * job.kc + 31) / 32);
    job.fmt = ctx->cfg.data_format;
    job.epoch_bytes = (int32_t)ctx->epoch_bytes;
    job.max_active = 1;
    for (int e = first; e < first + n; e++)
        job.max_active = std::max<int>(job.max_active, ctx->h_nch[e]);
    job.force_wrap_path = ctx->opt_force_slow;
    job.float_narrow = ctx->opt_float_geom == 1 ? 1 : 0;
    job.lean = ctx->opt_lean;
    job.lin_nb = 0;
    job.lin_rinv = 0.0;
    if (which == SynthKernel::Tuned32 && !job.carrier_float && job.accum == 1 && job.lean && ctx->opt_lowrate && ctx->lin_nb > 0) {
        job.lin_nb = ctx->lin_nb;
        job.lin_rinv = 2.0 / (ctx->d_min + ctx->d_max);
    }

    return which;
}

// a sub-range [first, first+n) of a planned job (same layout and checkpoints): K2 only
DeviceJob sub_job(const gpusim_ctx *ctx, const DeviceJob &whole, int first, int n, uint8_t *out_dev)
{
    DeviceJob job = whole;
    job.rows = whole.rows + (size_t)first * kMaxChan;
    job.nch = whole.nch + first;
    job.x0 = whole.x0 + (size_t)first * kMaxChan;
    job.ck_e0 = whole.ck_e0 + first;
    if (whole.carrier_float) {
        job.dc = whole.dc + (size_t)first * kMaxChan;
        job.cph0 = whole.cph0 + (size_t)first * kMaxChan;
    }
    job.out = out_dev;
    job.n_epochs = n;
    job.n_units = job.ppe > 0 ? (int32_t)((((long long)n * job.ppe + 31) / 32) * job.q)
                              : (int32_t)(((long long)n * job.kc + 31) / 32);
    (void)ctx;
    return job;
}

// Launch K1 + K2 for uploaded epochs [first, first+n) into out_dev.  K2 runs on `stream`.  When calls are
// issued back to back (a stream of batches) and the previous call's K2 is still in flight, K1 of this call
// goes to the library's chain stream instead: it then starts on the SMs the previous K2 has already left (a
// persistent kernel drains over its last work unit) instead of waiting for its last block.  The two sets of
// checkpoints and work counters the calls alternate between make that safe.  Option pipeline: 0 never
// (default: with the chain kernel at 0.2 ms the two orders time the same, 4.27 against 4.28 ms per step),
// 1 when the previous K2 is still running, 2 always.
int launch_range(gpusim_ctx *ctx, int first, int n, uint8_t *out_dev, cudaStream_t stream)
{
    if (n <= 0)
        return GPUSIM_OK;
    const int slot = (int)(ctx->seq++ & 1u);
    DeviceJob job;
    const SynthKernel which = plan_job(ctx, first, n, out_dev, job, slot);
    bool overlap = ctx->opt_pipeline != 0;
    if (overlap && ctx->opt_pipeline != 2) {
        overlap = cudaEventQuery(ctx->ev_s1[slot ^ 1]) == cudaErrorNotReady;
        (void)cudaGetLastError(); // "not ready" is an answer, not an error
    }
    cudaStream_t cs = overlap ? ctx->s_chain : stream;
    GS_CUDA(ctx, cudaStreamWaitEvent(cs, ctx->ev_s1[slot], 0)); // K2 of two calls ago read this slot
    GS_CUDA(ctx, cudaEventRecord(ctx->ev_c0[slot], cs));
    GS_CUDA(ctx, launch_chain(job, ctx->opt_chain_replay ? ChainAlgo::Replay : ChainAlgo::Jump, cs));
    GS_CUDA(ctx, cudaEventRecord(ctx->ev_c1[slot], cs));
    GS_CUDA(ctx, cudaStreamWaitEvent(stream, ctx->ev_c1[slot], 0));
    GS_CUDA(ctx, cudaEventRecord(ctx->ev_s0[slot], stream));
    GS_CUDA(ctx, launch_synth(job, which, stream));
    GS_CUDA(ctx, cudaEventRecord(ctx->ev_s1[slot], stream));
    ctx->last_slot = slot;
    ctx->timing.launches += 2;
    ctx->timing.fast_path = (which != SynthKernel::Generic) ? 1 : 0;
    ctx->timing.chain_overlapped = overlap ? 1 : 0;
    return GPUSIM_OK;
}

int collect_timing(gpusim_ctx *ctx)
{
    float a = 0.f, b = 0.f;
    const int slot = ctx->last_slot;
    GS_CUDA(ctx, cudaEventSynchronize(ctx->ev_s1[slot]));
    GS_CUDA(ctx, cudaEventElapsedTime(&a, ctx->ev_c0[slot], ctx->ev_c1[slot]));
    GS_CUDA(ctx, cudaEventElapsedTime(&b, ctx->ev_s0[slot], ctx->ev_s1[slot]));
    ctx->timing.chain_ms += a;
    ctx->timing.synth_ms += b;
    ctx->timing.total_ms += a + b;
    return GPUSIM_OK;
}

// no kernel of an earlier call may still be reading the uploaded rows or a checkpoint slot
int drain(gpusim_ctx *ctx)
{
    GS_CUDA(ctx, cudaStreamSynchronize(ctx->s_chain));
    GS_CUDA(ctx, cudaStreamSynchronize(ctx->s_compute));
    GS_CUDA(ctx, cudaStreamSynchronize(ctx->s_copy));
    for (int i = 0; i < 2; i++)
        GS_CUDA(ctx, cudaEventSynchronize(ctx->ev_s1[i]));
    return GPUSIM_OK;
}

} // namespace

extern "C" {

int gpusim_abi_version(void) { return GPUSIM_ABI_VERSION; }

void *gpusim_host_alloc(size_t n_bytes)
{
    void *p = nullptr;
    if (cudaMallocHost(&p, n_bytes ? n_bytes : 1) != cudaSuccess)
        return nullptr;
    return p;
}

void gpusim_host_free(void *p)
{
    if (p)
        cudaFreeHost(p);
}

const char *gpusim_strerror(int status)
{
    switch (status) {
    case GPUSIM_OK: return "ok";
    case GPUSIM_ERR_ARG: return "invalid argument or table value out of range";
    case GPUSIM_ERR_CUDA: return "CUDA error / no usable device";
    case GPUSIM_ERR_CAPACITY: return "batch exceeds the context's capacity";
    case GPUSIM_ERR_SINK: return "output sink failed";
    case GPUSIM_ERR_UNSUPPORTED: return "not supported by this build";
    default: return "unknown status";
    }
}

const char *gpusim_last_error(const gpusim_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

size_t gpusim_epoch_bytes(const gpusim_ctx *ctx) { return ctx ? ctx->epoch_bytes : 0; }

void gpusim_destroy(gpusim_ctx *ctx)
{
    if (!ctx)
        return;
    cudaSetDevice(ctx->cfg.device);
    if (ctx->s_chain) cudaStreamSynchronize(ctx->s_chain);
    if (ctx->s_compute) cudaStreamSynchronize(ctx->s_compute);
    if (ctx->s_copy) cudaStreamSynchronize(ctx->s_copy);
    for (const gpusim_ctx::DevAlloc &a : ctx->allocs)
        cudaFree(a.base);
    cudaFreeHost(ctx->h_dc); cudaFreeHost(ctx->h_cph0);
    cudaFreeHost(ctx->h_rows); cudaFreeHost(ctx->h_nch); cudaFreeHost(ctx->h_x0);
    cudaFreeHost(ctx->h_stage[0]); cudaFreeHost(ctx->h_stage[1]);
    for (cudaEvent_t ev : {ctx->ev_c0[0], ctx->ev_c0[1], ctx->ev_c1[0], ctx->ev_c1[1], ctx->ev_s0[0], ctx->ev_s0[1],
                           ctx->ev_s1[0], ctx->ev_s1[1], ctx->ev_done[0], ctx->ev_done[1],
                           ctx->ev_copied[0], ctx->ev_copied[1]})
        if (ev) cudaEventDestroy(ev);
    if (ctx->s_chain) cudaStreamDestroy(ctx->s_chain);
    if (ctx->s_compute) cudaStreamDestroy(ctx->s_compute);
    if (ctx->s_copy) cudaStreamDestroy(ctx->s_copy);
    delete ctx;
}

int gpusim_create(const gpusim_config *cfg, gpusim_ctx **out_ctx)
{
    if (!cfg || !out_ctx)
        return fail(nullptr, GPUSIM_ERR_ARG, "gpusim_create: null argument");
    *out_ctx = nullptr;
    if (cfg->abi_version != GPUSIM_ABI_VERSION)
        return fail(nullptr, GPUSIM_ERR_ARG, "ABI version mismatch: host %d, library %d", cfg->abi_version, GPUSIM_ABI_VERSION);
    if (cfg->samples_per_epoch < 1 || cfg->max_batch_epochs < 1 || !(cfg->delt > 0.0))
        return fail(nullptr, GPUSIM_ERR_ARG, "samples_per_epoch, max_batch_epochs and delt must be positive");
    if (cfg->data_format != GPUSIM_SC01 && cfg->data_format != GPUSIM_SC08 && cfg->data_format != GPUSIM_SC16)
        return fail(nullptr, GPUSIM_ERR_ARG, "data_format must be 1, 8 or 16");
    if (cfg->carrier_mode != GPUSIM_CARRIER_INT && cfg->carrier_mode != GPUSIM_CARRIER_FLOAT)
        return fail(nullptr, GPUSIM_ERR_ARG, "unknown carrier_mode");

    // GPUSIM_VERBOSE=2: where the start-up time of a context goes (stderr)
    const char *vb = getenv("GPUSIM_VERBOSE");
    const bool verbose = vb != nullptr && atoi(vb) >= 2;
    auto t_last = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (!verbose)
            return;
        cudaDeviceSynchronize();
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "gpusim_create: %-34s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
        t_last = now;
    };
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
        return fail(nullptr, GPUSIM_ERR_CUDA, "no CUDA device: %s (this library has no CPU path)",
                    ce != cudaSuccess ? cudaGetErrorString(ce) : "device count is 0");
    if (cfg->device < 0 || cfg->device >= ndev)
        return fail(nullptr, GPUSIM_ERR_ARG, "device %d out of range (0..%d)", cfg->device, ndev - 1);

    gpusim_ctx *ctx = new gpusim_ctx();
    ctx->cfg = *cfg;
    {
        const char *g = getenv("GPUSIM_GUARD");
        ctx->guard = g != nullptr && atoi(g) != 0;
    }
    const int N = cfg->samples_per_epoch;
    ctx->epoch_bytes = cfg->data_format == GPUSIM_SC01 ? (size_t)(N / 4) : cfg->data_format == GPUSIM_SC08 ? (size_t)2 * N : (size_t)4 * N;

#define GS_CREATE(call)                                                                            \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            fail(nullptr, GPUSIM_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_));         \
            gpusim_destroy(ctx);                                                                   \
            return GPUSIM_ERR_CUDA;                                                                \
        }                                                                                          \
    } while (0)

    lap("driver init (cudaGetDeviceCount)");
    GS_CREATE(cudaSetDevice(cfg->device));
    GS_CREATE(cudaFree(nullptr));
    lap("context (cudaSetDevice + cudaFree(0))");
    GS_CREATE(cudaStreamCreateWithFlags(&ctx->s_compute, cudaStreamNonBlocking));
    GS_CREATE(cudaStreamCreateWithFlags(&ctx->s_chain, cudaStreamNonBlocking));
    GS_CREATE(cudaStreamCreateWithFlags(&ctx->s_copy, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) {
        GS_CREATE(cudaEventCreate(&ctx->ev_c0[i]));
        GS_CREATE(cudaEventCreate(&ctx->ev_c1[i]));
        GS_CREATE(cudaEventCreate(&ctx->ev_s0[i]));
        GS_CREATE(cudaEventCreate(&ctx->ev_s1[i]));
    }
    for (int i = 0; i < 2; i++) {
        GS_CREATE(cudaEventCreateWithFlags(&ctx->ev_done[i], cudaEventDisableTiming));
        GS_CREATE(cudaEventCreateWithFlags(&ctx->ev_copied[i], cudaEventDisableTiming));
    }

    // constant tables
    {
        int32_t s[512], c[512], lut[512];
        uint64_t lut2[512];
        int16_t s16[512], c16[512];
        carrier_lut(s, c);
        for (int i = 0; i < 512; i++) {
            lut[i] = AccWide::table_entry(c[i], s[i]);
            lut2[i] = AccF32x2::table_entry(c[i], s[i]);
            s16[i] = (int16_t)s[i];
            c16[i] = (int16_t)c[i];
        }
        std::vector<uint32_t> negw((size_t)kCaPrns * kCaWords, 0xffffffffu);
        for (int prn = 1; prn <= 32; prn++) {
            uint32_t w[kCaWords];
            ca_words(prn, w);
            for (int i = 0; i < kCaWords; i++)
                negw[(size_t)prn * kCaWords + i] = ~w[i]; // bit set = chip 0 = codeCA -1 (gpssim.c:2241)
        }
        GS_CREATE(dev_alloc(ctx, &ctx->d_lut, sizeof(lut)));
        GS_CREATE(dev_alloc(ctx, &ctx->d_lut_f32, sizeof(lut2)));
        GS_CREATE(cudaMemcpy(ctx->d_lut_f32, lut2, sizeof(lut2), cudaMemcpyHostToDevice));
        GS_CREATE(dev_alloc(ctx, &ctx->d_sin16, sizeof(s16)));
        GS_CREATE(dev_alloc(ctx, &ctx->d_cos16, sizeof(c16)));
        GS_CREATE(dev_alloc(ctx, &ctx->d_negw, negw.size() * sizeof(uint32_t)));
        GS_CREATE(cudaMemcpy(ctx->d_lut, lut, sizeof(lut), cudaMemcpyHostToDevice));
        GS_CREATE(cudaMemcpy(ctx->d_sin16, s16, sizeof(s16), cudaMemcpyHostToDevice));
        GS_CREATE(cudaMemcpy(ctx->d_cos16, c16, sizeof(c16), cudaMemcpyHostToDevice));
        GS_CREATE(cudaMemcpy(ctx->d_negw, negw.data(), negw.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    }

    lap("streams, events, constant tables");
    const size_t rows = (size_t)cfg->max_batch_epochs * kMaxChan;
    GS_CREATE(dev_alloc(ctx, &ctx->d_rows, rows * sizeof(DevRow)));
    GS_CREATE(dev_alloc(ctx, &ctx->d_nch, (size_t)cfg->max_batch_epochs));
    GS_CREATE(dev_alloc(ctx, &ctx->d_x0, rows * sizeof(double)));
    lap("device row buffers");
    GS_CREATE(cudaMallocHost(&ctx->h_rows, rows * sizeof(DevRow)));
    GS_CREATE(cudaMallocHost(&ctx->h_nch, (size_t)cfg->max_batch_epochs));
    GS_CREATE(cudaMallocHost(&ctx->h_x0, rows * sizeof(double)));

    lap("page-locked row buffers");
    // checkpoints: 10 (18 with a double carrier) bytes per (row, chunk); raise the minimum chunk until
    // they fit the budget
    const bool cf = cfg->carrier_mode == GPUSIM_CARRIER_FLOAT;
    ctx->min_chunk = 128;
    while (rows * (size_t)kc_for(N, ctx->min_chunk) * (cf ? 18 : 10) > kCheckpointBudget && ctx->min_chunk < (1 << 20))
        ctx->min_chunk *= 2;
    const size_t cks = ck_elems(cfg->max_batch_epochs, kc_for(N, ctx->min_chunk));
    for (int i = 0; i < 2; i++) {
        GS_CREATE(dev_alloc(ctx, &ctx->d_ck_x[i], cks * sizeof(double)));
        GS_CREATE(dev_alloc(ctx, &ctx->d_ck_w[i], cks * sizeof(uint16_t)));
        if (cf)
            GS_CREATE(dev_alloc(ctx, &ctx->d_ck_c[i], cks * sizeof(double)));
    }
    GS_CREATE(dev_alloc(ctx, &ctx->d_work, 64));
    if (cf) {
        GS_CREATE(dev_alloc(ctx, &ctx->d_dc, rows * sizeof(double)));
        GS_CREATE(dev_alloc(ctx, &ctx->d_cph0, rows * sizeof(double)));
        GS_CREATE(cudaMallocHost(&ctx->h_dc, rows * sizeof(double)));
        GS_CREATE(cudaMallocHost(&ctx->h_cph0, rows * sizeof(double)));
    }
    GS_CREATE(cudaMemset(ctx->d_work, 0, 64));
    GS_CREATE(cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, cfg->device));
    lap("checkpoint buffers");
#undef GS_CREATE

    *out_ctx = ctx;
    return GPUSIM_OK;
}

int gpusim_set_option(gpusim_ctx *ctx, const char *key, int64_t value)
{
    if (!ctx || !key)
        return GPUSIM_ERR_ARG;
    if (!strcmp(key, "chunk")) ctx->opt_chunk = (int)value;
    else if (!strcmp(key, "force_generic")) ctx->opt_force_generic = (int)value;
    else if (!strcmp(key, "force_slow")) ctx->opt_force_slow = (int)value;
    else if (!strcmp(key, "chain_replay")) ctx->opt_chain_replay = (int)value;
    else if (!strcmp(key, "accum")) ctx->opt_accum = (int)value;
    else if (!strcmp(key, "layout")) ctx->opt_layout = (int)value;
    else if (!strcmp(key, "pipeline")) ctx->opt_pipeline = (int)value;
    else if (!strcmp(key, "float_geom")) ctx->opt_float_geom = (int)value;
    else if (!strcmp(key, "lean")) ctx->opt_lean = (int)value;
    else if (!strcmp(key, "lowrate")) ctx->opt_lowrate = (int)value;
    else if (!strcmp(key, "direct_first_mb")) ctx->opt_direct_first_mb = (int)std::max<int64_t>(1, value);
    else if (!strcmp(key, "direct_mb")) ctx->opt_direct_mb = (int)std::max<int64_t>(1, value);
    else return fail(ctx, GPUSIM_ERR_ARG, "unknown option '%s'", key);
    return GPUSIM_OK;
}

int gpusim_upload_table(gpusim_ctx *ctx, const gpusim_epoch_table *t)
{
    if (!ctx || !t)
        return GPUSIM_ERR_ARG;
    if (t->n_epochs < 0 || t->n_epochs > ctx->cfg.max_batch_epochs)
        return fail(ctx, GPUSIM_ERR_CAPACITY, "table has %d epochs, context capacity is %d", t->n_epochs, ctx->cfg.max_batch_epochs);
    const bool cf = ctx->cfg.carrier_mode == GPUSIM_CARRIER_FLOAT;
    if (!t->prn || !t->f_code || !t->code_phase || !t->icode || !t->gain)
        return fail(ctx, GPUSIM_ERR_ARG, "tables need prn, f_code, code_phase, icode, gain");
    const bool nav_ref = t->nav_bits == nullptr; // data bits come from device-built frames (gpusim_nav_build)
    if (nav_ref && (!t->nav_frame || !t->iword || !t->ibit))
        return fail(ctx, GPUSIM_ERR_ARG, "tables need nav_bits, or nav_frame + iword + ibit");
    if (!cf && (!t->carr_phasestep || !t->carr_phase))
        return fail(ctx, GPUSIM_ERR_ARG, "integer-carrier tables need carr_phasestep and carr_phase");
    if (cf && (!t->f_carr || !t->carr_phase_f))
        return fail(ctx, GPUSIM_ERR_ARG, "FLOAT_CARR_PHASE tables need f_carr and carr_phase_f");
    GS_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    // the previous upload may still be read by kernels in flight
    {
        int rc_drain = drain(ctx);
        if (rc_drain != GPUSIM_OK)
            return rc_drain;
    }

    const double delt = ctx->cfg.delt;
    ctx->needs_generic = false;
    ctx->d_max = 0.0;
    ctx->d_min = (double)kCaLen;
    ctx->lin_nb = 0;
    for (int e = 0; e < t->n_epochs; e++) {
        int nc = 0;
        for (int i = 0; i < kMaxChan; i++) {
            const size_t r = (size_t)e * kMaxChan + i;
            if (t->prn[r] <= 0) // gpssim.c:2197: slot skipped
                continue;
            if (t->prn[r] > 32)
                return fail(ctx, GPUSIM_ERR_ARG, "epoch %d slot %d: prn %d (1..32 have C/A codes, gpssim.c:145)", e, i, t->prn[r]);
            const double x0 = t->code_phase[r];
            const volatile double d = t->f_code[r] * delt; // the reference's rounded product, gpssim.c:2212
            if (!(x0 >= 0.0 && x0 < (double)kCaLen) || !(d > 0.0 && d < (double)kCaLen) || t->icode[r] < 0 || t->icode[r] > 19)
                return fail(ctx, GPUSIM_ERR_ARG, "epoch %d slot %d: code_phase %.17g, f_code*delt %.17g or icode %d outside the reference's invariants", e, i, x0, (double)d, t->icode[r]);
            DevRow &o = ctx->h_rows[(size_t)e * kMaxChan + nc];
            o.d = d;
            o.steps = cf ? 0 : (int32_t)((uint32_t)t->carr_phasestep[r] << 7);
            o.cthr_prn = pack_cthr_prn(d, t->prn[r]);
            o.woff = (uint16_t)(t->prn[r] * kCaWinBytes);
            o.ph0s = cf ? 0u : t->carr_phase[r] << 7;
            if (cf) {
                // gpssim.c:2245: carr_phase += f_carr*delt (rounded product), phase in [0,1); kept x512
                const volatile double dcar = t->f_carr[r] * delt;
                const double cp = t->carr_phase_f[r];
                if (!(cp >= 0.0 && cp < 1.0) || !(dcar > -1.0 && dcar < 1.0))
                    return fail(ctx, GPUSIM_ERR_ARG, "epoch %d slot %d: carr_phase %.17g or f_carr*delt %.17g outside [0,1) / (-1,1)", e, i, cp, (double)dcar);
                ctx->h_dc[(size_t)e * kMaxChan + nc] = dcar * 512.0;
                ctx->h_cph0[(size_t)e * kMaxChan + nc] = cp * 512.0;
            }
            o.gain = t->gain[r];
            o.icode0 = (uint16_t)t->icode[r];
            o.flags = 0;
            if (nav_ref) {
                // (chan[i].iword, chan[i].ibit) of gpssim.c:1343-1344 inside frame nav_frame; resolved by k0_navbits
                if (t->nav_frame[r] < 0 || t->nav_frame[r] >= ctx->nav_frames || t->iword[r] < 0 || t->iword[r] >= kNavWords ||
                    t->ibit[r] < 0 || t->ibit[r] >= 30)
                    return fail(ctx, GPUSIM_ERR_ARG, "epoch %d slot %d: nav_frame %d (have %d), iword %d, ibit %d out of range", e, i,
                                t->nav_frame[r], ctx->nav_frames, t->iword[r], t->ibit[r]);
                o.nav_bits = ((uint32_t)t->nav_frame[r] << 11) | ((uint32_t)t->iword[r] << 5) | (uint32_t)t->ibit[r];
                o.flags |= kRowNavRef;
            } else {
                o.nav_bits = t->nav_bits[r];
            }
            if (o.gain < 0 || o.gain > kTunedMaxGain) {
                o.flags |= kRowNeedsGeneric;
                ctx->needs_generic = true;
            }
            ctx->d_max = std::max(ctx->d_max, (double)d);
            ctx->d_min = std::min(ctx->d_min, (double)d);
            if (chain_tie_binade(d) >= 5 && chain_tie_binade(d) <= 9)
                o.flags |= kRowTieInLinRange;
            ctx->h_x0[(size_t)e * kMaxChan + nc] = x0;
            nc++;
        }
        for (int k = nc; k < kMaxChan; k++) {
            memset(&ctx->h_rows[(size_t)e * kMaxChan + k], 0, sizeof(DevRow));
            ctx->h_x0[(size_t)e * kMaxChan + k] = 0.0;
            if (cf) {
                ctx->h_dc[(size_t)e * kMaxChan + k] = 0.0;
                ctx->h_cph0[(size_t)e * kMaxChan + k] = 0.0;
            }
        }
        ctx->h_nch[e] = (uint8_t)nc;
    }
    // Low chip rates (synth_lin): at most 2 or 4 chip boundaries per run of 32 samples, and one job-wide 1/d
    // good to 1e-3 samples over a whole run.  Rows whose step is an exact half-ulp tie in one of the binades
    // the linear model is used in get a wrap threshold of 0: they always take the exact per-sample loop.
    if (!cf && t->n_epochs > 0 && ctx->d_max > 0.0 && 32.0 * ctx->d_max < 4.0 &&
        5.0 * (ctx->d_max - ctx->d_min) < 0.25 * ctx->d_min * ctx->d_min) { // quotient estimate off by < 0.25 samples
        ctx->lin_nb = 32.0 * ctx->d_max < 2.0 ? 2 : 4;
        for (size_t r = 0; r < (size_t)t->n_epochs * kMaxChan; r++)
            if (ctx->h_rows[r].flags & kRowTieInLinRange)
                ctx->h_rows[r].cthr_prn &= (uint16_t)~kCthrMask;
    }
    const size_t rows = (size_t)t->n_epochs * kMaxChan;
    if (rows) {
        GS_CUDA(ctx, cudaMemcpyAsync(ctx->d_rows, ctx->h_rows, rows * sizeof(DevRow), cudaMemcpyHostToDevice, ctx->s_compute));
        GS_CUDA(ctx, cudaMemcpyAsync(ctx->d_x0, ctx->h_x0, rows * sizeof(double), cudaMemcpyHostToDevice, ctx->s_compute));
        GS_CUDA(ctx, cudaMemcpyAsync(ctx->d_nch, ctx->h_nch, (size_t)t->n_epochs, cudaMemcpyHostToDevice, ctx->s_compute));
        if (cf) {
            GS_CUDA(ctx, cudaMemcpyAsync(ctx->d_dc, ctx->h_dc, rows * sizeof(double), cudaMemcpyHostToDevice, ctx->s_compute));
            GS_CUDA(ctx, cudaMemcpyAsync(ctx->d_cph0, ctx->h_cph0, rows * sizeof(double), cudaMemcpyHostToDevice, ctx->s_compute));
        }
        if (nav_ref)
            GS_CUDA(ctx, launch_navbits(ctx->d_rows, (int)rows, ctx->d_nav_words, ctx->s_compute));
        GS_CUDA(ctx, cudaStreamSynchronize(ctx->s_compute));
    }
    ctx->n_uploaded = t->n_epochs;
    return GPUSIM_OK;
}

int gpusim_nav_build(gpusim_ctx *ctx, const gpusim_nav_frame *frames, int32_t n_frames)
{
    static_assert(sizeof(gpusim_nav_frame) == sizeof(NavFrame), "gpusim_nav_frame is the device layout");
    if (!ctx || n_frames < 0 || (n_frames > 0 && !frames))
        return GPUSIM_ERR_ARG;
    if ((uint32_t)n_frames > kNavRefMaxFrames)
        return fail(ctx, GPUSIM_ERR_CAPACITY, "%d navigation frames, at most %u per call", n_frames, kNavRefMaxFrames);
    GS_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    {
        // rows of an earlier upload may still be resolved against the old words
        const int rc_drain = drain(ctx);
        if (rc_drain != GPUSIM_OK)
            return rc_drain;
    }
    if (n_frames > ctx->nav_capacity) {
        // grows geometrically; the old buffers stay in the context's allocation list until gpusim_destroy
        const int cap = std::max(n_frames, std::max(64, 2 * ctx->nav_capacity));
        GS_CUDA(ctx, dev_alloc(ctx, &ctx->d_nav_req, (size_t)cap * sizeof(NavFrame)));
        GS_CUDA(ctx, dev_alloc(ctx, &ctx->d_nav_words, (size_t)cap * kNavWords * sizeof(uint32_t)));
        ctx->nav_capacity = cap;
    }
    ctx->nav_frames = ctx->nav_ephs = 0;
    if (n_frames > 0) {
        GS_CUDA(ctx, cudaMemcpyAsync(ctx->d_nav_req, frames, (size_t)n_frames * sizeof(NavFrame), cudaMemcpyHostToDevice, ctx->s_compute));
        GS_CUDA(ctx, launch_navmsg(ctx->d_nav_req, n_frames, ctx->d_nav_words, ctx->s_compute));
        GS_CUDA(ctx, cudaStreamSynchronize(ctx->s_compute)); // `frames` is the caller's again
    }
    ctx->nav_frames = n_frames;
    return GPUSIM_OK;
}

int gpusim_nav_build_eph(gpusim_ctx *ctx, const gpusim_nav_eph *eph, int32_t n_eph, const gpusim_nav_iono *iono,
                         const gpusim_nav_frame_ref *frames, int32_t n_frames)
{
    static_assert(sizeof(gpusim_nav_eph) == sizeof(NavEph) && sizeof(gpusim_nav_iono) == sizeof(NavIono) &&
                      sizeof(gpusim_nav_frame_ref) == sizeof(NavFrameRef) && sizeof(NavFrameRef) <= sizeof(NavFrame),
                  "ABI structs are the device layouts");
    if (!ctx || n_eph < 0 || n_frames < 0 || (n_eph > 0 && !eph) || (n_frames > 0 && !frames) || !iono)
        return GPUSIM_ERR_ARG;
    if ((uint32_t)n_frames > kNavRefMaxFrames)
        return fail(ctx, GPUSIM_ERR_CAPACITY, "%d navigation frames, at most %u per call", n_frames, kNavRefMaxFrames);
    for (int f = 0; f < n_frames; f++)
        if (frames[f].eph < 0 || frames[f].eph >= n_eph || frames[f].eph_first < 0 || frames[f].eph_first >= n_eph)
            return fail(ctx, GPUSIM_ERR_ARG, "frame %d names ephemerides %d / %d, have %d", f, frames[f].eph, frames[f].eph_first, n_eph);
    GS_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    {
        const int rc_drain = drain(ctx);
        if (rc_drain != GPUSIM_OK)
            return rc_drain;
    }
    if (n_frames > ctx->nav_capacity) {
        const int cap = std::max(n_frames, std::max(64, 2 * ctx->nav_capacity));
        GS_CUDA(ctx, dev_alloc(ctx, &ctx->d_nav_req, (size_t)cap * sizeof(NavFrame)));
        GS_CUDA(ctx, dev_alloc(ctx, &ctx->d_nav_words, (size_t)cap * kNavWords * sizeof(uint32_t)));
        ctx->nav_capacity = cap;
    }
    if (n_eph > ctx->nav_eph_capacity) {
        const int cap = std::max(n_eph, std::max(64, 2 * ctx->nav_eph_capacity));
        GS_CUDA(ctx, dev_alloc(ctx, &ctx->d_nav_eph, (size_t)cap * sizeof(NavEph)));
        GS_CUDA(ctx, dev_alloc(ctx, &ctx->d_nav_sbf, (size_t)cap * kNavSbfWords * sizeof(uint32_t)));
        ctx->nav_eph_capacity = cap;
    }
    ctx->nav_frames = ctx->nav_ephs = 0;
    NavIono io;
    memcpy(&io, iono, sizeof(io));
    if (n_eph > 0) {
        GS_CUDA(ctx, cudaMemcpyAsync(ctx->d_nav_eph, eph, (size_t)n_eph * sizeof(NavEph), cudaMemcpyHostToDevice, ctx->s_compute));
        GS_CUDA(ctx, launch_eph2sbf(ctx->d_nav_eph, n_eph, io, ctx->d_nav_sbf, ctx->s_compute));
    }
    if (n_frames > 0) {
        // the frame requests share the buffer of gpusim_nav_build's (larger) requests
        NavFrameRef *d_ref = reinterpret_cast<NavFrameRef *>(ctx->d_nav_req);
        GS_CUDA(ctx, cudaMemcpyAsync(d_ref, frames, (size_t)n_frames * sizeof(NavFrameRef), cudaMemcpyHostToDevice, ctx->s_compute));
        GS_CUDA(ctx, launch_navmsg_ref(d_ref, n_frames, ctx->d_nav_sbf, ctx->d_nav_words, ctx->s_compute));
    }
    GS_CUDA(ctx, cudaStreamSynchronize(ctx->s_compute));
    ctx->nav_frames = n_frames;
    ctx->nav_ephs = n_eph;
    return GPUSIM_OK;
}

int gpusim_nav_read_sbf(gpusim_ctx *ctx, int32_t first, int32_t n, uint32_t *sbf)
{
    if (!ctx || (n > 0 && !sbf))
        return GPUSIM_ERR_ARG;
    if (first < 0 || n < 0 || first > ctx->nav_ephs || n > ctx->nav_ephs - first)
        return fail(ctx, GPUSIM_ERR_ARG, "ephemerides first=%d n=%d outside the %d built", first, n, ctx->nav_ephs);
    if (n == 0)
        return GPUSIM_OK;
    GS_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    GS_CUDA(ctx, cudaMemcpyAsync(sbf, ctx->d_nav_sbf + (size_t)first * kNavSbfWords, (size_t)n * kNavSbfWords * sizeof(uint32_t),
                                 cudaMemcpyDeviceToHost, ctx->s_compute));
    GS_CUDA(ctx, cudaStreamSynchronize(ctx->s_compute));
    return GPUSIM_OK;
}

int gpusim_nav_read(gpusim_ctx *ctx, int32_t first, int32_t n, uint32_t *dwrd)
{
    if (!ctx || (n > 0 && !dwrd))
        return GPUSIM_ERR_ARG;
    if (first < 0 || n < 0 || first > ctx->nav_frames || n > ctx->nav_frames - first)
        return fail(ctx, GPUSIM_ERR_ARG, "frames first=%d n=%d outside the %d built", first, n, ctx->nav_frames);
    if (n == 0)
        return GPUSIM_OK;
    GS_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    GS_CUDA(ctx, cudaMemcpyAsync(dwrd, ctx->d_nav_words + (size_t)first * kNavWords, (size_t)n * kNavWords * sizeof(uint32_t),
                                 cudaMemcpyDeviceToHost, ctx->s_compute));
    GS_CUDA(ctx, cudaStreamSynchronize(ctx->s_compute));
    return GPUSIM_OK;
}

int gpusim_generate_device(gpusim_ctx *ctx, int32_t first, int32_t n, void *out_device, size_t cap, void *stream)
{
    if (!ctx)
        return GPUSIM_ERR_ARG;
    if (first < 0 || n < 0 || first > ctx->n_uploaded || n > ctx->n_uploaded - first)
        return fail(ctx, GPUSIM_ERR_ARG, "epoch range first=%d n=%d outside the uploaded table (%d epochs)", first, n, ctx->n_uploaded);
    if ((size_t)n * ctx->epoch_bytes > cap)
        return fail(ctx, GPUSIM_ERR_CAPACITY, "output needs %zu bytes, capacity is %zu", (size_t)n * ctx->epoch_bytes, cap);
    if (!out_device || ((uintptr_t)out_device & 15))
        return fail(ctx, GPUSIM_ERR_ARG, "out_device must be a 16-byte aligned device pointer");
    GS_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    cudaStream_t s = stream ? (cudaStream_t)stream : ctx->s_compute;
    ctx->timing = gpusim_timing{};
    int rc = launch_range(ctx, first, n, (uint8_t *)out_device, s);
    if (rc != GPUSIM_OK)
        return rc;
    if (!stream)
        GS_CUDA(ctx, cudaStreamSynchronize(s));
    return GPUSIM_OK;
}

int64_t gpusim_debug_guard_violations(gpusim_ctx *ctx)
{
    if (!ctx || !ctx->guard)
        return -1;
    if (cudaSetDevice(ctx->cfg.device) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess)
        return -2;
    std::vector<unsigned char> band(2 * kGuardBytes);
    int64_t bad = 0;
    for (const gpusim_ctx::DevAlloc &a : ctx->allocs) {
        if (cudaMemcpy(band.data(), a.base, kGuardBytes, cudaMemcpyDeviceToHost) != cudaSuccess ||
            cudaMemcpy(band.data() + kGuardBytes, a.user + a.bytes, kGuardBytes, cudaMemcpyDeviceToHost) != cudaSuccess)
            return -2;
        for (unsigned char b : band)
            bad += b != (unsigned char)kGuardPoison;
    }
    return bad;
}

int gpusim_get_timing(const gpusim_ctx *cctx, gpusim_timing *out)
{
    gpusim_ctx *ctx = const_cast<gpusim_ctx *>(cctx);
    if (!ctx || !out)
        return GPUSIM_ERR_ARG;
    if (ctx->timing.launches > 0 && ctx->timing.total_ms == 0.f) {
        int rc = collect_timing(ctx);
        if (rc != GPUSIM_OK)
            return rc;
    }
    *out = ctx->timing;
    return GPUSIM_OK;
}

// common driver of the two host-output entry points: sub-batches of <= kStageBytes are generated
// on s_compute and copied back on s_copy while the next sub-batch is being generated
static int generate_to_host_body(gpusim_ctx *ctx, const gpusim_epoch_table *t, uint8_t *out, gpusim_sink_fn sink, void *user);

// Whatever way the body fails (CUDA error, sink failure), nothing of this call may still be running when
// the caller gets the error back: kernels write the output buffer, copies write the staging buffers or
// the caller's own memory.  Drain both streams; the slot's "synthesis done" event is recorded so that
// later calls waiting on it see this call's work as finished.
static int generate_to_host(gpusim_ctx *ctx, const gpusim_epoch_table *t, uint8_t *out, gpusim_sink_fn sink, void *user)
{
    const int rc = generate_to_host_body(ctx, t, out, sink, user);
    if (rc != GPUSIM_OK && ctx->s_compute) {
        const std::string keep = ctx->err;
        cudaStreamSynchronize(ctx->s_compute);
        cudaStreamSynchronize(ctx->s_copy);
        cudaEventRecord(ctx->ev_s1[ctx->last_slot], ctx->s_compute);
        cudaEventSynchronize(ctx->ev_s1[ctx->last_slot]);
        (void)cudaGetLastError();
        ctx->err = keep;
    }
    return rc;
}

static int generate_to_host_body(gpusim_ctx *ctx, const gpusim_epoch_table *t, uint8_t *out, gpusim_sink_fn sink, void *user)
{
    int rc = gpusim_upload_table(ctx, t);
    if (rc != GPUSIM_OK)
        return rc;
    if ((rc = ensure_out(ctx)) != GPUSIM_OK)
        return rc;
    if (sink && (rc = ensure_stage(ctx)) != GPUSIM_OK)
        return rc;
    ctx->timing = gpusim_timing{};
    const int n = t->n_epochs;
    const size_t eb = ctx->epoch_bytes;
    if (n == 0 || eb == 0)
        return GPUSIM_OK;
    // sub-batch sizes in epochs: staging buffers bound them in sink mode; straight into the caller's
    // buffer the first one is small (the copy engine starts early) and the rest large (fewer copies)
    const size_t sub_bytes = sink ? kStageBytes : ((size_t)ctx->opt_direct_mb << 20);
    const size_t first_bytes = sink ? kStageBytes : ((size_t)ctx->opt_direct_first_mb << 20);
    const int sub = (int)std::max<size_t>(1, std::min<size_t>((size_t)n, sub_bytes / eb));
    const int sub_first = (int)std::max<size_t>(1, std::min<size_t>((size_t)sub, first_bytes / eb));

    // The chain kernel is latency bound (its duration is one chain, whatever the batch): run it once
    // for the whole table, then generate and copy back sub-batch by sub-batch.
    // (upload_table drained every earlier call, so either checkpoint slot is free.)
    const int slot = (int)(ctx->seq++ & 1u);
    ctx->last_slot = slot;
    DeviceJob whole;
    const SynthKernel which = plan_job(ctx, 0, n, ctx->d_out, whole, slot);
    GS_CUDA(ctx, cudaEventRecord(ctx->ev_c0[slot], ctx->s_compute));
    GS_CUDA(ctx, launch_chain(whole, ctx->opt_chain_replay ? ChainAlgo::Replay : ChainAlgo::Jump, ctx->s_compute));
    GS_CUDA(ctx, cudaEventRecord(ctx->ev_c1[slot], ctx->s_compute));
    GS_CUDA(ctx, cudaEventRecord(ctx->ev_s0[slot], ctx->s_compute));
    ctx->timing.launches += 1;
    ctx->timing.fast_path = (which != SynthKernel::Generic) ? 1 : 0;

    int pending = -1; // sub-batch whose copy has been issued but not yet delivered
    int pending_first = 0, pending_n = 0;
    int b = 0;
    for (int first = 0, cnt = 0; first < n; first += cnt, b++) {
        cnt = std::min(b == 0 ? sub_first : sub, n - first);
        uint8_t *dst_dev = ctx->d_out + (size_t)first * eb;
        const DeviceJob job = sub_job(ctx, whole, first, cnt, dst_dev);
        if (b > 0)
            GS_CUDA(ctx, cudaMemsetAsync(whole.work_counter, 0, sizeof(unsigned int), ctx->s_compute));
        GS_CUDA(ctx, launch_synth(job, which, ctx->s_compute));
        ctx->timing.launches += 1;
        GS_CUDA(ctx, cudaEventRecord(ctx->ev_done[b & 1], ctx->s_compute));
        GS_CUDA(ctx, cudaStreamWaitEvent(ctx->s_copy, ctx->ev_done[b & 1], 0));
        uint8_t *dst_host = sink ? ctx->h_stage[b & 1] : out + (size_t)first * eb;
        GS_CUDA(ctx, cudaMemcpyAsync(dst_host, dst_dev, (size_t)cnt * eb, cudaMemcpyDeviceToHost, ctx->s_copy));
        GS_CUDA(ctx, cudaEventRecord(ctx->ev_copied[b & 1], ctx->s_copy));
        // deliver the previous sub-batch while this one is generated and copied
        if (pending >= 0 && sink) {
            GS_CUDA(ctx, cudaEventSynchronize(ctx->ev_copied[pending & 1]));
            if (sink(user, ctx->h_stage[pending & 1], (size_t)pending_n * eb) != 0)
                return fail(ctx, GPUSIM_ERR_SINK, "sink failed at epoch %d", pending_first);
        }
        pending = b;
        pending_first = first;
        pending_n = cnt;
    }
    GS_CUDA(ctx, cudaEventRecord(ctx->ev_s1[slot], ctx->s_compute));
    GS_CUDA(ctx, cudaStreamSynchronize(ctx->s_copy));
    if (pending >= 0 && sink)
        if (sink(user, ctx->h_stage[pending & 1], (size_t)pending_n * eb) != 0)
            return fail(ctx, GPUSIM_ERR_SINK, "sink failed at epoch %d", pending_first);
    return collect_timing(ctx); // chain = K1; synth = all K2 sub-batches including copy waits between them
}

int gpusim_generate_epochs(gpusim_ctx *ctx, const gpusim_epoch_table *t, void *out, size_t cap)
{
    if (!ctx || !t || !out)
        return GPUSIM_ERR_ARG;
    if ((size_t)t->n_epochs * ctx->epoch_bytes > cap)
        return fail(ctx, GPUSIM_ERR_CAPACITY, "output needs %zu bytes, capacity is %zu", (size_t)t->n_epochs * ctx->epoch_bytes, cap);
    return generate_to_host(ctx, t, (uint8_t *)out, nullptr, nullptr);
}

int gpusim_generate_epochs_to_sink(gpusim_ctx *ctx, const gpusim_epoch_table *t, gpusim_sink_fn sink, void *user)
{
    if (!ctx || !t || !sink)
        return GPUSIM_ERR_ARG;
    return generate_to_host(ctx, t, nullptr, sink, user);
}

} // extern "C"
