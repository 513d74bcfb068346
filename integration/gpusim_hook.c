/*
 * gpusim_hook.c - host shim between the reference's epoch scheduler and libgpusim.
 *
 * Compiled with the reference's gpssim.h on the include path.  Per epoch it
 * copies the per-channel state the reference's sample loop (gpssim.c:2190-2264)
 * would have started from into a structure-of-arrays batch, advances the one
 * quantity that loop carried from epoch to epoch (the carrier phase,
 * gpssim.c:2243-2253) arithmetically, and every GPUSIM_BATCH_EPOCHS epochs hands
 * the batch to gpusim_generate_epochs_to_sink(), whose sink is the reference's
 * own fwrite on the reference's own FILE* (gpssim.c:2276/:2283/:2287).
 *
 * Environment (all optional):
 *   GPUSIM_BATCH_EPOCHS  epochs per library call, fixed     (default: 256, doubling per call up to a
 *                        memory-derived cap of at most 4096)
 *   GPUSIM_DEVICE        CUDA device ordinal                (default 0)
 *   GPUSIM_DEVICES       number of GPUs (devices 0..n-1) to time-shard batches over (default 1):
 *                        whichever GPU is free takes the next batch; one worker thread per GPU generates
 *                        into page-locked buffers, one writer thread fwrites the batches in order
 *   GPUSIM_DEVICE_LIST   same, with explicit ordinals, e.g. "0,2,5" (or "0,0" to run two workers
 *                        on one GPU)
 *   GPUSIM_DUMP          path: also write every table row to this file
 *                        (format below) - how tests/golden/ fixtures are made
 *   GPUSIM_DRYRUN        1: record (and dump) rows, generate nothing, write nothing
 *   GPUSIM_HOST_THREADS  host threads for the row pre-pass (default: online CPUs, at most 16; 1 = the
 *                        reference's serial order of calls, nothing precomputed)
 *   GPUSIM_NAV_DEVICE    1: navigation data words on the device.  Rows reference frames
 *                        (one per generateNavMsg() call, rebuilt from chan[i].sbf by gpusim_nav_build)
 *                        instead of carrying 32 data bits taken from chan[i].dwrd.
 *                        2: eph2sbf() on the device as well (gpusim_nav_build_eph): the shim hands over the
 *                        broadcast ephemerides (the eph[ieph] the epoch loop works with, seen through
 *                        GPUSIM_HOOK_RANGE) and frames name their subframes by ephemeris index.
 *   GPUSIM_NAV_CHECK     1 (with GPUSIM_NAV_DEVICE): read the device-built words back after every
 *                        build and compare them with the host's chan[i].dwrd; a difference is fatal
 *   GPUSIM_NAV_DUMP      path (with GPUSIM_NAV_DEVICE): every frame request (struct gpusim_nav_frame)
 *                        followed by the 60 words the host's generateNavMsg() holds for it
 *
 * Host pre-pass.  Producing the rows is the reference's own code and stays
 * bit-identical; what changes is when and where it runs:
 *   - the main thread only fills rows; generation + fwrite run on worker threads (one per GPU), so
 *     the pre-pass of batch b+1 overlaps the GPU work of batch b;
 *   - GPUSIM_HOOK_RANGE() replaces the computeRange() call of the epoch loop (gpssim.c:2165-2168)
 *     by a look-up into a window of ranges computed ahead of time, in parallel over epochs, by the
 *     reference's own computeRange() - a pure function of (ephemeris, iono, time, position); a
 *     window never crosses a 30 s channel/ephemeris refresh (gpssim.c:2296-2345) and every look-up
 *     re-checks time, PRN and ephemeris set, falling back to a direct call;
 *   - FLOAT_CARR_PHASE hosts: the double carrier phase chains through every sample of the run
 *     (gpssim.c:2245-2250); its per-epoch advance is an exact O(carrier cycles) walk that costs
 *     ~40 us per channel and epoch.  The chains of different channels are independent, so they are
 *     filled in per batch, one thread per channel slot, instead of inside the epoch loop.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include "gpssim.h"
#include "gpusim_hook.h"
#include "gpusim.h"

#if MAX_CHAN != GPUSIM_MAX_CHAN
#error "gpssim.h MAX_CHAN and gpusim.h GPUSIM_MAX_CHAN disagree"
#endif

/* one growable set of SoA columns */
typedef struct
{
	int cap; /* epochs */
	int n;   /* epochs filled */
	int32_t *prn;
	double *f_code;
	double *code_phase;
	int32_t *icode;
	uint32_t *nav_bits;
	int32_t *gain;
	int32_t *carr_phasestep;
	uint32_t *carr_phase;
	double *f_carr;
	double *carr_phase_f;
	int32_t *iword; /* dump; GPUSIM_NAV_DEVICE: position of the row's first data bit in its frame */
	int32_t *ibit;
	double *carr_init; /* FLOAT hosts: carrier phase of a freshly allocated channel, or HOOK_CONTINUE */
	/* GPUSIM_NAV_DEVICE: the navigation frames this batch's rows reference (index = nav_frame[row]) */
	int32_t *nav_frame;
	gpusim_nav_frame *frames;
	uint32_t *frames_expect; /* GPUSIM_NAV_CHECK: chan[i].dwrd[60] as the host built it, per frame */
	int n_frames, cap_frames;
	/* GPUSIM_NAV_DEVICE=2: the ephemerides the frames' subframes come from (index = frame_refs[].eph / .eph_first) */
	gpusim_nav_frame_ref *frame_refs; /* parallel to frames[] */
	gpusim_nav_eph *ephs;
	const ephem_t **eph_keys;         /* the host's ephemeris each entry was made from (dedupe) */
	int n_ephs, cap_ephs;
	gpusim_nav_iono iono;
} cols_t;

/* GPUSIM_NAV_DEVICE: what the shim remembers about the frame a channel slot is transmitting */
typedef struct
{
	int active;              /* the slot held this PRN in the previous epoch */
	int prn;
	gpstime_t g0;            /* chan[i].g0 of the frame (gpssim.c:1478) */
	uint32_t seen_sbf[5][N_DWRD_SBF]; /* chan[i].sbf as of the previous epoch: what the NEXT refresh's
	                                     generateNavMsg(init=0) will read - the reference rebuilds the
	                                     subframes of a new ephemeris set AFTER that call (gpssim.c:2300-2330) */
	gpusim_nav_frame frame;  /* the current frame's request */
	int batch_index;         /* its index in the current batch's frame list, -1 = not registered yet */
	/* GPUSIM_NAV_DEVICE=2: the host ephemeris behind seen_sbf, behind the current frame's subframes and behind its
	 * first ten words (eph[ieph][prn-1] of the epoch loop; the arrays live in main() for the whole run) */
	const ephem_t *seen_eph, *frame_eph, *first_eph;
} nav_slot_t;

/* the reference's own functions (gpssim.c:789, :1253; external linkage, not declared in gpssim.h) */
extern void computeRange(range_t *rho, ephem_t eph, ionoutc_t *ionoutc, gpstime_t g, double xyz[]);
extern gpstime_t incGpsTime(gpstime_t g0, double dt);

/* FLOAT hosts: value left in chan[i].carr_phase after an epoch was recorded.  The reference only ever
 * stores values in [0,1) there (gpssim.c:1622, :2245-2250), so anything else found at the next epoch
 * means allocateChannel() has (re)started the channel with a fresh phase. */
#define HOOK_CONTINUE (-1.0)

/* ---- fork/join pool for the host pre-pass ------------------------------------------------ */
typedef void (*pool_fn)(void *arg, int item);
typedef struct
{
	int n; /* helper threads (the caller works too) */
	pthread_t *threads;
	pthread_mutex_t mu;
	pthread_cond_t cv_job, cv_done;
	pool_fn fn;
	void *arg;
	int n_items;
	int next;    /* next unclaimed item (atomic) */
	int running; /* helpers still inside the current job */
	long gen;    /* job generation */
	int stop;
} pool_t;

/* look-ahead window of computeRange() results */
#define RA_WIN 512
typedef struct
{
	int enabled;
	int first, count;        /* epochs [first, first+count) of the host's iumd counter */
	const ephem_t *eph_set;  /* eph[ieph] the window was computed with */
	ionoutc_t *ionoutc;
	double (*xyz)[3];        /* user motion, NULL in static mode */
	double *xyz0;
	int prn[MAX_CHAN];
	gpstime_t g[RA_WIN];
	range_t *rho;            /* [RA_WIN][MAX_CHAN] */
	long hits, direct, windows;
	int short_windows;       /* consecutive windows that served a single epoch */
} lookahead_t;

/* GPU pipeline: a ring of batch slots; slot states advance FREE -> QUEUED -> DONE -> FREE */
#define HOOK_MAX_DEV 16
enum { SLOT_FREE = 0, SLOT_QUEUED = 1, SLOT_DONE = 2 };
typedef struct
{
	cols_t rows;
	unsigned char *out; /* page-locked, batch capacity */
	int state;
} slot_t;

struct gpusim_hook;
typedef struct
{
	struct gpusim_hook *h;
	gpusim_ctx *ctx;
	int index;
	int device;
	long batches; /* batches this worker generated (GPUSIM_VERBOSE) */
	pthread_t thread;
} worker_t;

struct gpusim_hook
{
	int N;
	double delt;
	int fmt;
	FILE *fp;
	int carrier_mode;
	int dryrun;
	const char *dump_path;
	gpusim_ctx *ctx;
	cols_t batch;
	cols_t dump;
	long epochs_done;
	double t_open, t_wait, t_chains, t_ranges; /* wall clock [s]: start; main thread waiting for a free slot, walking carrier chains, computing range windows */
	int flush_at, cap_max; /* epochs per batch: current size (doubles per batch) and its cap */
	int host_threads;
	pool_t pool;
	lookahead_t ra;
	double cph[MAX_CHAN]; /* FLOAT hosts: carrier phase of every slot at the next unfilled epoch */
	/* GPU pipeline: ndev workers (sink mode when ndev == 1: the worker fwrites, no writer thread) */
	int ndev, nslots;
	slot_t *slots;
	worker_t workers[HOOK_MAX_DEV];
	pthread_t writer;
	pthread_mutex_t mu;
	pthread_cond_t cv;
	long seq_filled; /* batches handed to the workers so far */
	long seq_taken;  /* batches claimed by a worker so far: whichever GPU is free takes the next one */
	int finishing;   /* no more batches will be queued */
	int failed;      /* a worker or the writer hit an error: everybody stops, the MAIN thread reports and exits */
	char fail_msg[600];
	long epochs_written; /* epochs delivered to the output file so far (what a failure report states) */
	gpusim_config cfg;   /* what every worker creates its context with */
	size_t epoch_bytes;  /* gpssim.c:2276/:2283/:2287 */
	int nav_device, nav_check; /* GPUSIM_NAV_DEVICE (1: subframes from the host, 2: ephemerides from the host), GPUSIM_NAV_CHECK */
	const ephem_t *cur_eph_set; /* eph[ieph] and ionoutc of the epoch being recorded (from GPUSIM_HOOK_RANGE) */
	const ionoutc_t *cur_iono;
	FILE *nav_dump;            /* GPUSIM_NAV_DUMP */
	nav_slot_t nav[MAX_CHAN];
	long nav_frames_built;
	int slots_ready;     /* page-locked ring buffers allocated (several GPUs; done by worker 0 once CUDA is up) */
	double t_ctx;        /* wall clock [s] the slowest worker spent creating its context */
};

static double now_s(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static void die(const char *what, const char *detail)
{
	fprintf(stderr, "\nERROR: %s%s%s\n", what, detail ? ": " : "", detail ? detail : "");
	exit(1);
}

/* A failure on a worker or writer thread: record it and wake everybody.  Nothing exits from those
 * threads - another thread may be in the middle of an fwrite.  The main thread notices at its next
 * hand-over (or at close), joins the pipeline, flushes what was written and exits like the reference
 * does (message on stderr, status 1). */
static void pipeline_fail(struct gpusim_hook *h, const char *what, const char *detail);
static void pipeline_check(struct gpusim_hook *h);

static void *xrealloc(void *p, size_t n)
{
	void *q = realloc(p, n ? n : 1);
	if (q == NULL)
		die("gpusim hook out of memory", NULL);
	return q;
}

static void cols_reserve(cols_t *c, int epochs)
{
	size_t rows;
	if (epochs <= c->cap)
		return;
	if (epochs < 2 * c->cap)
		epochs = 2 * c->cap;
	rows = (size_t)epochs * MAX_CHAN;
	c->prn = xrealloc(c->prn, rows * sizeof(int32_t));
	c->f_code = xrealloc(c->f_code, rows * sizeof(double));
	c->code_phase = xrealloc(c->code_phase, rows * sizeof(double));
	c->icode = xrealloc(c->icode, rows * sizeof(int32_t));
	c->nav_bits = xrealloc(c->nav_bits, rows * sizeof(uint32_t));
	c->gain = xrealloc(c->gain, rows * sizeof(int32_t));
	c->carr_phasestep = xrealloc(c->carr_phasestep, rows * sizeof(int32_t));
	c->carr_phase = xrealloc(c->carr_phase, rows * sizeof(uint32_t));
	c->f_carr = xrealloc(c->f_carr, rows * sizeof(double));
	c->carr_phase_f = xrealloc(c->carr_phase_f, rows * sizeof(double));
	c->iword = xrealloc(c->iword, rows * sizeof(int32_t));
	c->ibit = xrealloc(c->ibit, rows * sizeof(int32_t));
	c->carr_init = xrealloc(c->carr_init, rows * sizeof(double));
	c->nav_frame = xrealloc(c->nav_frame, rows * sizeof(int32_t));
	c->cap = epochs;
}

static void cols_free(cols_t *c)
{
	free(c->prn); free(c->f_code); free(c->code_phase); free(c->icode);
	free(c->nav_bits); free(c->gain); free(c->carr_phasestep); free(c->carr_phase);
	free(c->f_carr); free(c->carr_phase_f); free(c->iword); free(c->ibit); free(c->carr_init);
	free(c->nav_frame); free(c->frames); free(c->frames_expect); free(c->frame_refs); free(c->ephs); free(c->eph_keys);
	memset(c, 0, sizeof(*c));
}

/* copy epoch `e` of src to the end of dst */
static void cols_append(cols_t *dst, const cols_t *src, int e)
{
	size_t o = (size_t)e * MAX_CHAN, d;
	cols_reserve(dst, dst->n + 1);
	d = (size_t)dst->n * MAX_CHAN;
#define CP(f) memcpy(dst->f + d, src->f + o, MAX_CHAN * sizeof(*dst->f))
	CP(prn); CP(f_code); CP(code_phase); CP(icode); CP(nav_bits); CP(gain);
	CP(carr_phasestep); CP(carr_phase); CP(f_carr); CP(carr_phase_f); CP(iword); CP(ibit); CP(carr_init);
#undef CP
	dst->n++;
}

static int sink_fwrite(void *user, const void *bytes, size_t n)
{
	gpusim_hook *h = (gpusim_hook *)user;
	/* the reference does not check fwrite either (gpssim.c:2276) - but a short
	 * write to a full disk should not go unnoticed at GB/s */
	return fwrite(bytes, 1, n, h->fp) == n ? 0 : 1;
}

static void table_of(const cols_t *c, gpusim_epoch_table *t, int nav_device)
{
	memset(t, 0, sizeof(*t));
	t->n_epochs = c->n;
	if (nav_device)
	{
		t->nav_frame = c->nav_frame;
		t->iword = c->iword;
		t->ibit = c->ibit;
	}
	t->prn = c->prn;
	t->f_code = c->f_code;
	t->code_phase = c->code_phase;
	t->icode = c->icode;
	t->nav_bits = nav_device ? NULL : c->nav_bits;
	t->gain = c->gain;
	t->carr_phasestep = c->carr_phasestep;
	t->carr_phase = c->carr_phase;
	t->f_carr = c->f_carr;
	t->carr_phase_f = c->carr_phase_f;
}

/* ---- fork/join pool ---------------------------------------------------------------------- */
static void pool_work(pool_t *p)
{
	for (;;)
	{
		int item = __atomic_fetch_add(&p->next, 1, __ATOMIC_RELAXED);
		if (item >= p->n_items)
			return;
		p->fn(p->arg, item);
	}
}

static void *pool_main(void *arg)
{
	pool_t *p = (pool_t *)arg;
	long seen = 0;
	pthread_mutex_lock(&p->mu);
	for (;;)
	{
		while (!p->stop && p->gen == seen)
			pthread_cond_wait(&p->cv_job, &p->mu);
		if (p->stop)
			break;
		seen = p->gen;
		pthread_mutex_unlock(&p->mu);
		pool_work(p);
		pthread_mutex_lock(&p->mu);
		if (--p->running == 0)
			pthread_cond_signal(&p->cv_done);
	}
	pthread_mutex_unlock(&p->mu);
	return NULL;
}

static void pool_start(pool_t *p, int helpers)
{
	int i;
	memset(p, 0, sizeof(*p));
	pthread_mutex_init(&p->mu, NULL);
	pthread_cond_init(&p->cv_job, NULL);
	pthread_cond_init(&p->cv_done, NULL);
	if (helpers <= 0)
		return;
	p->threads = xrealloc(NULL, (size_t)helpers * sizeof(pthread_t));
	for (i = 0; i < helpers; i++)
	{
		if (pthread_create(&p->threads[i], NULL, pool_main, p) != 0)
			break;
		p->n++;
	}
}

/* run fn(arg, 0..n_items-1) on the helpers and the calling thread; returns when all are done */
static void pool_run(pool_t *p, pool_fn fn, void *arg, int n_items)
{
	int i;
	if (p->n == 0 || n_items <= 1)
	{
		for (i = 0; i < n_items; i++)
			fn(arg, i);
		return;
	}
	pthread_mutex_lock(&p->mu);
	p->fn = fn;
	p->arg = arg;
	p->n_items = n_items;
	__atomic_store_n(&p->next, 0, __ATOMIC_RELAXED);
	p->running = p->n;
	p->gen++;
	pthread_cond_broadcast(&p->cv_job);
	pthread_mutex_unlock(&p->mu);
	pool_work(p);
	pthread_mutex_lock(&p->mu);
	while (p->running > 0)
		pthread_cond_wait(&p->cv_done, &p->mu);
	pthread_mutex_unlock(&p->mu);
}

static void pool_stop(pool_t *p)
{
	int i;
	pthread_mutex_lock(&p->mu);
	p->stop = 1;
	pthread_cond_broadcast(&p->cv_job);
	pthread_mutex_unlock(&p->mu);
	for (i = 0; i < p->n; i++)
		pthread_join(p->threads[i], NULL);
	free(p->threads);
	p->threads = NULL;
	p->n = 0;
}

/* ---- FLOAT hosts: the carrier-phase chains of one batch, one channel slot per work item ---- */
#ifdef FLOAT_CARR_PHASE
static void carrier_slot(void *arg, int slot)
{
	gpusim_hook *h = (gpusim_hook *)arg;
	cols_t *b = &h->batch;
	double cph = h->cph[slot];
	int e;
	for (e = 0; e < b->n; e++)
	{
		size_t o = (size_t)e * MAX_CHAN + slot;
		if (b->prn[o] <= 0)
			continue;
		if (b->carr_init[o] != HOOK_CONTINUE)
			cph = b->carr_init[o]; /* allocateChannel() started this channel here, gpssim.c:1622 */
		b->carr_phase_f[o] = cph;
		/* the exact result of the N updates of gpssim.c:2245-2250 */
		cph = gpusim_advance_carrier_f64(cph, b->f_carr[o], h->delt, h->N);
	}
	h->cph[slot] = cph;
}
#endif

static void fill_carrier_chains(gpusim_hook *h)
{
#ifdef FLOAT_CARR_PHASE
	pool_run(&h->pool, carrier_slot, h, MAX_CHAN);
#else
	(void)h;
#endif
}

/* ---- computeRange() look-ahead ------------------------------------------------------------ */
#define RA_GROUP 4 /* epochs per work item */
static void lookahead_item(void *arg, int item)
{
	lookahead_t *ra = &((gpusim_hook *)arg)->ra;
	int k, i;
	for (k = item * RA_GROUP; k < (item + 1) * RA_GROUP && k < ra->count; k++)
		for (i = 0; i < MAX_CHAN; i++)
			if (ra->prn[i] > 0)
				computeRange(&ra->rho[(size_t)k * MAX_CHAN + i], ra->eph_set[ra->prn[i] - 1], ra->ionoutc, ra->g[k],
				             ra->xyz != NULL ? ra->xyz[ra->first + k] : ra->xyz0);
}

static void lookahead_build(gpusim_hook *h, int iumd, int numd, const channel_t *chan, const ephem_t *eph_set,
                            ionoutc_t *ionoutc, gpstime_t grx, double (*xyz)[3], double *xyz0)
{
	lookahead_t *ra = &h->ra;
	gpstime_t g = grx;
	int i, k;

	if (ra->rho == NULL)
		ra->rho = xrealloc(NULL, (size_t)RA_WIN * MAX_CHAN * sizeof(range_t));
	/* a window that served one epoch only means the assumptions below do not hold for this host */
	if (ra->windows > 0 && ra->count > 1 && iumd == ra->first + 1)
	{
		if (++ra->short_windows >= 8)
		{
			ra->enabled = 0;
			ra->count = 0;
			return;
		}
	}
	else
		ra->short_windows = 0;

	ra->first = iumd;
	ra->eph_set = eph_set;
	ra->ionoutc = ionoutc;
	ra->xyz = xyz;
	ra->xyz0 = xyz0;
	for (i = 0; i < MAX_CHAN; i++)
		ra->prn[i] = chan[i].prn;
	/* epochs up to and including the one whose end refreshes channels / ephemerides
	 * (gpssim.c:2294-2296); receiver time advances exactly as the host will advance it (:2348) */
	ra->count = 0;
	for (k = 0; k < RA_WIN && iumd + k < numd; k++)
	{
		ra->g[k] = g;
		ra->count = k + 1;
		if ((int)(g.sec * 10.0 + 0.5) % 300 == 0)
			break;
		g = incGpsTime(g, 0.1);
	}
	ra->windows++;
	{
		const double t0 = now_s();
		pool_run(&h->pool, lookahead_item, h, (ra->count + RA_GROUP - 1) / RA_GROUP);
		h->t_ranges += now_s() - t0;
	}
}

void gpusim_hook_range(gpusim_hook *h, range_t *rho, int slot, int iumd, int numd, const channel_t *chan,
                       const ephem_t *eph_set, ionoutc_t *ionoutc, gpstime_t grx, double (*xyz)[3], double *xyz0)
{
	lookahead_t *ra = &h->ra;
	int attempt;
	h->cur_eph_set = eph_set;
	h->cur_iono = ionoutc;
	for (attempt = 0; ra->enabled && attempt < 2; attempt++)
	{
		int k = iumd - ra->first;
		if (k >= 0 && k < ra->count && ra->eph_set == eph_set && ra->ionoutc == ionoutc && ra->xyz == xyz &&
		    ra->xyz0 == xyz0 && ra->prn[slot] == chan[slot].prn && ra->g[k].week == grx.week && ra->g[k].sec == grx.sec)
		{
			*rho = ra->rho[(size_t)k * MAX_CHAN + slot];
			ra->hits++;
			return;
		}
		if (attempt == 0)
			lookahead_build(h, iumd, numd, chan, eph_set, ionoutc, grx, xyz, xyz0);
	}
	/* the reference's own call, gpssim.c:2165-2168 */
	ra->direct++;
	computeRange(rho, eph_set[chan[slot].prn - 1], ionoutc, grx, xyz != NULL ? xyz[iumd] : xyz0);
}

/* ---- GPU pipeline: workers generate, batches are written strictly in order --------------------- */
static void pipeline_fail(gpusim_hook *h, const char *what, const char *detail)
{
	pthread_mutex_lock(&h->mu);
	if (!h->failed)
	{
		h->failed = 1;
		snprintf(h->fail_msg, sizeof(h->fail_msg), "%s%s%s", what, detail ? ": " : "", detail ? detail : "");
	}
	pthread_cond_broadcast(&h->cv);
	pthread_mutex_unlock(&h->mu);
}

static void pipeline_join(gpusim_hook *h)
{
	int d;
	pthread_mutex_lock(&h->mu);
	h->finishing = 1;
	pthread_cond_broadcast(&h->cv);
	pthread_mutex_unlock(&h->mu);
	for (d = 0; d < h->ndev; d++)
		pthread_join(h->workers[d].thread, NULL);
	if (h->ndev > 1)
		pthread_join(h->writer, NULL);
}

/* main thread only */
static void pipeline_check(gpusim_hook *h)
{
	int failed;
	if (h->dryrun)
		return;
	pthread_mutex_lock(&h->mu);
	failed = h->failed;
	pthread_mutex_unlock(&h->mu);
	if (!failed)
		return;
	pipeline_join(h);
	fflush(h->fp);
	fprintf(stderr, "\nERROR: %s\n       (%ld of the epochs handed over so far were written to the output before the failure)\n",
	        h->fail_msg, h->epochs_written);
	exit(1);
}

static void *worker_main(void *arg)
{
	worker_t *w = (worker_t *)arg;
	gpusim_hook *h = w->h;
	/* CUDA start-up (driver initialisation + context: several hundred ms) happens HERE, on the worker, while
	 * the main thread is already reading the RINEX file and filling rows. */
	{
		const double t0 = now_s();
		gpusim_config cfg = h->cfg;
		int rc;
		cfg.device = w->device;
		rc = gpusim_create(&cfg, &w->ctx);
		if (rc != GPUSIM_OK)
		{
			pipeline_fail(h, "Failed to initialise the GPU sample generator", gpusim_last_error(NULL));
			return NULL;
		}
		if (h->ndev > 1 && w->index == 0)
		{
			int d;
			for (d = 0; d < h->nslots; d++)
			{
				h->slots[d].out = gpusim_host_alloc((size_t)h->cap_max * h->epoch_bytes);
				if (h->slots[d].out == NULL)
				{
					pipeline_fail(h, "Failed to allocate page-locked output buffers", NULL);
					return NULL;
				}
			}
		}
		pthread_mutex_lock(&h->mu);
		if (h->ndev == 1 || w->index == 0)
			h->slots_ready = 1;
		if (now_s() - t0 > h->t_ctx)
			h->t_ctx = now_s() - t0;
		pthread_cond_broadcast(&h->cv);
		while (!h->slots_ready && !h->failed)
			pthread_cond_wait(&h->cv, &h->mu);
		pthread_mutex_unlock(&h->mu);
	}
	for (;;)
	{
		slot_t *s;
		gpusim_epoch_table t;
		long seq;
		int rc;
		/* Whichever GPU is free takes the next batch (not batch b -> GPU b mod n): the GPUs of a box do
		 * not share the host link evenly, and an even split would let the slowest link set the pace. */
		pthread_mutex_lock(&h->mu);
		while (!h->failed && h->seq_taken >= h->seq_filled && !h->finishing)
			pthread_cond_wait(&h->cv, &h->mu);
		if (h->failed || h->seq_taken >= h->seq_filled)
		{
			pthread_mutex_unlock(&h->mu);
			return NULL;
		}
		seq = h->seq_taken++;
		s = &h->slots[seq % h->nslots];
		pthread_mutex_unlock(&h->mu);

		table_of(&s->rows, &t, h->nav_device);
		if (h->nav_device)
		{
			/* generateNavMsg + computeChecksum for the frames of this batch, on this worker's GPU */
			rc = h->nav_device == 2
			         ? gpusim_nav_build_eph(w->ctx, s->rows.ephs, s->rows.n_ephs, &s->rows.iono, s->rows.frame_refs, s->rows.n_frames)
			         : gpusim_nav_build(w->ctx, s->rows.frames, s->rows.n_frames);
			if (rc == GPUSIM_OK && h->nav_check)
			{
				uint32_t *got = malloc((size_t)(s->rows.n_frames > 0 ? s->rows.n_frames : 1) * N_DWRD * sizeof(uint32_t));
				rc = got == NULL ? GPUSIM_ERR_ARG : gpusim_nav_read(w->ctx, 0, s->rows.n_frames, got);
				if (rc == GPUSIM_OK && memcmp(got, s->rows.frames_expect, (size_t)s->rows.n_frames * N_DWRD * sizeof(uint32_t)) != 0)
				{
					free(got);
					pipeline_fail(h, "GPUSIM_NAV_CHECK: device-built navigation words differ from generateNavMsg()", NULL);
					return NULL;
				}
				free(got);
			}
			if (rc != GPUSIM_OK)
			{
				pipeline_fail(h, "building the navigation frames on the GPU failed", gpusim_last_error(w->ctx));
				return NULL;
			}
		}
		if (h->ndev == 1) /* one GPU: this thread is also the (ordered) writer, from the library's staging buffers */
			rc = gpusim_generate_epochs_to_sink(w->ctx, &t, sink_fwrite, h);
		else
			rc = gpusim_generate_epochs(w->ctx, &t, s->out, (size_t)h->cap_max * h->epoch_bytes);
		if (rc != GPUSIM_OK)
		{
			pipeline_fail(h, rc == GPUSIM_ERR_SINK ? "Failed to write the output file" : "GPU sample generation failed",
			              gpusim_last_error(w->ctx));
			return NULL;
		}
		w->batches++;

		pthread_mutex_lock(&h->mu);
		if (h->ndev == 1)
			h->epochs_written += s->rows.n;
		s->state = h->ndev == 1 ? SLOT_FREE : SLOT_DONE;
		pthread_cond_broadcast(&h->cv);
		pthread_mutex_unlock(&h->mu);
	}
}

static void *writer_main(void *arg)
{
	gpusim_hook *h = (gpusim_hook *)arg;
	const size_t eb = h->epoch_bytes;
	long seq;
	for (seq = 0;; seq++)
	{
		slot_t *s = &h->slots[seq % h->nslots];
		pthread_mutex_lock(&h->mu);
		while (!h->failed && !(s->state == SLOT_DONE && seq < h->seq_filled) && !(h->finishing && seq >= h->seq_filled))
			pthread_cond_wait(&h->cv, &h->mu);
		if (h->failed || seq >= h->seq_filled)
		{
			pthread_mutex_unlock(&h->mu);
			return NULL;
		}
		pthread_mutex_unlock(&h->mu);

		/* strictly in batch order: this is the reference's output file */
		if (fwrite(s->out, 1, (size_t)s->rows.n * eb, h->fp) != (size_t)s->rows.n * eb)
		{
			pipeline_fail(h, "Failed to write the output file", NULL);
			return NULL;
		}

		pthread_mutex_lock(&h->mu);
		h->epochs_written += s->rows.n;
		s->state = SLOT_FREE;
		pthread_cond_broadcast(&h->cv);
		pthread_mutex_unlock(&h->mu);
	}
}

/* ---- GPUSIM_NAV_DEVICE: frames instead of data bits ----------------------------- */
/* the batch that is being filled starts without frames: every slot registers its frame again */
static void nav_new_batch(gpusim_hook *h)
{
	int i;
	h->batch.n_frames = 0;
	h->batch.n_ephs = 0;
	for (i = 0; i < MAX_CHAN; i++)
		h->nav[i].batch_index = -1;
}

static void nav_copy_sbf(uint32_t dst[5][N_DWRD_SBF], unsigned long src[5][N_DWRD_SBF])
{
	int a, b;
	for (a = 0; a < 5; a++)
		for (b = 0; b < N_DWRD_SBF; b++)
			dst[a][b] = (uint32_t)src[a][b]; /* generateNavMsg reads them into an `unsigned` (gpssim.c:1473) */
}

/* what eph2sbf() reads of ionoutc_t (gpssim.c:560-579) */
static void nav_iono_of(const ionoutc_t *io, gpusim_nav_iono *o)
{
	memset(o, 0, sizeof(*o));
	o->alpha0 = io->alpha0; o->alpha1 = io->alpha1; o->alpha2 = io->alpha2; o->alpha3 = io->alpha3;
	o->beta0 = io->beta0; o->beta1 = io->beta1; o->beta2 = io->beta2; o->beta3 = io->beta3;
	o->A0 = io->A0; o->A1 = io->A1;
	o->vflg = io->vflg; o->dtls = io->dtls; o->tot = io->tot; o->wnt = io->wnt;
}

/* GPUSIM_NAV_DEVICE=2: index of the host ephemeris `e` in the batch's list (what eph2sbf() reads of it, gpssim.c:490-665) */
static int nav_eph_index(gpusim_hook *h, const ephem_t *e)
{
	cols_t *b = &h->batch;
	gpusim_nav_eph *o;
	int k;
	if (e == NULL)
		die("GPUSIM_NAV_DEVICE=2 needs the GPUSIM_HOOK_RANGE() edit (it is how the shim sees eph[ieph])", NULL);
	for (k = 0; k < b->n_ephs; k++)
		if (b->eph_keys[k] == e)
			return k;
	if (b->n_ephs >= b->cap_ephs)
	{
		b->cap_ephs = b->cap_ephs ? 2 * b->cap_ephs : 32;
		b->ephs = xrealloc(b->ephs, (size_t)b->cap_ephs * sizeof(gpusim_nav_eph));
		b->eph_keys = xrealloc(b->eph_keys, (size_t)b->cap_ephs * sizeof(*b->eph_keys));
	}
	o = &b->ephs[b->n_ephs];
	memset(o, 0, sizeof(*o));
	o->toe_sec = e->toe.sec; o->toc_sec = e->toc.sec; o->toe_week = e->toe.week;
	o->deltan = e->deltan; o->cuc = e->cuc; o->cus = e->cus; o->cic = e->cic; o->cis = e->cis; o->crc = e->crc; o->crs = e->crs;
	o->ecc = e->ecc; o->sqrta = e->sqrta; o->m0 = e->m0; o->omg0 = e->omg0; o->inc0 = e->inc0; o->aop = e->aop;
	o->omgdot = e->omgdot; o->idot = e->idot; o->af0 = e->af0; o->af1 = e->af1; o->af2 = e->af2; o->tgd = e->tgd;
	o->iodc = e->iodc; o->iode = e->iode; o->svhlth = e->svhlth; o->codeL2 = e->codeL2;
	b->eph_keys[b->n_ephs] = e;
	return b->n_ephs++;
}

/* Frame index (in the batch being filled) of what slot i transmits in this epoch.  A frame is one
 * generateNavMsg() call: chan[i].g0 or the PRN changed since the previous epoch. */
static int nav_frame_of(gpusim_hook *h, int i, channel_t *c)
{
	nav_slot_t *t = &h->nav[i];
	cols_t *b = &h->batch;
	const int same_sv = t->active && t->prn == c->prn;
	if (!same_sv || t->g0.week != c->g0.week || t->g0.sec != c->g0.sec)
	{
		gpusim_nav_frame f;
		memset(&f, 0, sizeof(f));
		f.wn = (uint32_t)((unsigned long)(c->g0.week % 1024));   /* gpssim.c:1480 */
		f.tow = (uint32_t)(((unsigned long)c->g0.sec) / 6UL);     /* gpssim.c:1481 */
		if (same_sv)
		{
			/* refresh, generateNavMsg(grx, &chan[i], 0) (gpssim.c:2300-2304): subframes as they were before
			 * the refresh; words 0..9 are the previous frame's words 50..59 (gpssim.c:1504-1511) */
			memcpy(f.sbf, t->seen_sbf, sizeof(f.sbf));
			memcpy(f.first, t->frame.sbf[4], sizeof(f.first));
			f.tow_first = t->frame.tow + 5u;
			t->first_eph = t->frame_eph;
			t->frame_eph = t->seen_eph;
		}
		else
		{
			/* fresh channel, generateNavMsg(grx, &chan[i], 1) right after eph2sbf (gpssim.c:1604-1608) */
			nav_copy_sbf(f.sbf, c->sbf);
			memcpy(f.first, f.sbf[4], sizeof(f.first));
			f.tow_first = f.tow;
			t->frame_eph = t->first_eph = h->cur_eph_set != NULL ? &h->cur_eph_set[c->prn - 1] : NULL;
		}
		t->frame = f;
		t->g0 = c->g0;
		t->prn = c->prn;
		t->batch_index = -1;
	}
	if (t->batch_index < 0)
	{
		if (b->n_frames >= b->cap_frames)
		{
			b->cap_frames = b->cap_frames ? 2 * b->cap_frames : 64;
			b->frames = xrealloc(b->frames, (size_t)b->cap_frames * sizeof(gpusim_nav_frame));
			b->frame_refs = xrealloc(b->frame_refs, (size_t)b->cap_frames * sizeof(gpusim_nav_frame_ref));
			b->frames_expect = xrealloc(b->frames_expect, (size_t)b->cap_frames * N_DWRD * sizeof(uint32_t));
		}
		b->frames[b->n_frames] = t->frame;
		if (h->nav_device == 2)
		{
			gpusim_nav_frame_ref *r = &b->frame_refs[b->n_frames];
			memset(r, 0, sizeof(*r));
			r->eph = nav_eph_index(h, t->frame_eph);
			r->eph_first = nav_eph_index(h, t->first_eph);
			r->tow_first = t->frame.tow_first;
			r->tow = t->frame.tow;
			r->wn = t->frame.wn;
		}
		if (h->nav_check || h->nav_dump != NULL)
		{
			uint32_t *x = b->frames_expect + (size_t)b->n_frames * N_DWRD;
			int k;
			for (k = 0; k < N_DWRD; k++)
				x[k] = (uint32_t)c->dwrd[k];
			if (h->nav_dump != NULL) /* request + what generateNavMsg() made of it: tests/test_navmsg.py */
			{
				fwrite(&t->frame, sizeof(t->frame), 1, h->nav_dump);
				fwrite(x, sizeof(uint32_t), N_DWRD, h->nav_dump);
				if (h->nav_device == 2) /* + the two ephemerides and the ionosphere / UTC block the frame refers to */
				{
					gpusim_nav_iono io;
					nav_iono_of(h->cur_iono, &io);
					fwrite(&b->ephs[b->frame_refs[b->n_frames].eph], sizeof(gpusim_nav_eph), 1, h->nav_dump);
					fwrite(&b->ephs[b->frame_refs[b->n_frames].eph_first], sizeof(gpusim_nav_eph), 1, h->nav_dump);
					fwrite(&io, sizeof(io), 1, h->nav_dump);
				}
			}
		}
		t->batch_index = b->n_frames++;
		h->nav_frames_built++;
	}
	t->active = 1;
	nav_copy_sbf(t->seen_sbf, c->sbf);
	t->seen_eph = h->cur_eph_set != NULL ? &h->cur_eph_set[c->prn - 1] : NULL;
	return t->batch_index;
}

/* hand the filled batch to the next GPU: swap it into a free ring slot */
static void queue_batch(gpusim_hook *h)
{
	slot_t *s = &h->slots[h->seq_filled % h->nslots];
	cols_t tmp;
	const double t0 = now_s();
	pthread_mutex_lock(&h->mu);
	while (s->state != SLOT_FREE && !h->failed)
		pthread_cond_wait(&h->cv, &h->mu);
	if (h->failed)
	{
		pthread_mutex_unlock(&h->mu);
		pipeline_check(h); /* does not return */
	}
	h->t_wait += now_s() - t0;
	tmp = s->rows;
	s->rows = h->batch;
	h->batch = tmp;
	h->batch.n = 0;
	nav_new_batch(h);
	s->state = SLOT_QUEUED;
	h->seq_filled++;
	pthread_cond_broadcast(&h->cv);
	pthread_mutex_unlock(&h->mu);
}

static void flush_batch(gpusim_hook *h)
{
	int e;

	if (h->batch.n == 0)
		return;

	{
		const double t0 = now_s();
		fill_carrier_chains(h);
		h->t_chains += now_s() - t0;
	}
	if (h->dump_path != NULL)
		for (e = 0; e < h->batch.n; e++)
			cols_append(&h->dump, &h->batch, e);

	h->epochs_done += h->batch.n;
	if (h->nav_device == 2 && h->cur_iono != NULL)
		nav_iono_of(h->cur_iono, &h->batch.iono);
	if (!h->dryrun)
		queue_batch(h); /* generation and fwrite happen on the worker / writer threads */
	else
	{
		h->batch.n = 0;
		nav_new_batch(h);
	}
}

static int default_host_threads(void)
{
	long n = sysconf(_SC_NPROCESSORS_ONLN);
	cpu_set_t set;
	if (sched_getaffinity(0, sizeof(set), &set) == 0 && CPU_COUNT(&set) > 0 && CPU_COUNT(&set) < n)
		n = CPU_COUNT(&set);
	if (n < 1)
		n = 1;
	return n > 16 ? 16 : (int)n;
}

gpusim_hook *gpusim_hook_open(int iq_buff_size, double delt, int data_format, FILE *fp)
{
	gpusim_hook *h = calloc(1, sizeof(*h));
	const char *s;
	int batch = 256;

	if (h == NULL)
		die("gpusim hook out of memory", NULL);

	h->N = iq_buff_size;
	h->delt = delt;
	h->fmt = data_format;
	h->fp = fp;
#ifdef FLOAT_CARR_PHASE
	h->carrier_mode = GPUSIM_CARRIER_FLOAT;
#else
	h->carrier_mode = GPUSIM_CARRIER_INT;
#endif
	h->dryrun = ((s = getenv("GPUSIM_DRYRUN")) != NULL && atoi(s) != 0);
	h->nav_device = (s = getenv("GPUSIM_NAV_DEVICE")) != NULL ? (atoi(s) >= 2 ? 2 : atoi(s) != 0) : 0;
	h->nav_check = h->nav_device && ((s = getenv("GPUSIM_NAV_CHECK")) != NULL && atoi(s) != 0);
	if (h->nav_device && (s = getenv("GPUSIM_NAV_DUMP")) != NULL && *s && (h->nav_dump = fopen(s, "wb")) == NULL)
		die("cannot open GPUSIM_NAV_DUMP", s);
	h->dump_path = getenv("GPUSIM_DUMP");

	h->ndev = 1;
	if ((s = getenv("GPUSIM_DEVICE_LIST")) != NULL && *s)
	{
		h->ndev = 0;
		while (*s && h->ndev < HOOK_MAX_DEV)
		{
			h->workers[h->ndev++].device = atoi(s);
			while (*s && *s != ',')
				s++;
			if (*s == ',')
				s++;
		}
	}
	else if ((s = getenv("GPUSIM_DEVICES")) != NULL && atoi(s) > 1)
	{
		int d;
		h->ndev = atoi(s) > HOOK_MAX_DEV ? HOOK_MAX_DEV : atoi(s);
		for (d = 0; d < h->ndev; d++)
			h->workers[d].device = d;
	}

	/* Epochs per library call.  Every call costs one code-phase chain kernel, whose duration (~0.5 ms)
	 * is the latency of one chain whatever the batch, plus a few synchronisations: long runs want
	 * large batches, the first bytes want a small one.  So batches start at 256 epochs and double up
	 * to a cap set by memory (device output per call; page-locked ring buffers with several GPUs). */
	{
		const double eb = data_format == 16 ? 4.0 * iq_buff_size : data_format == 8 ? 2.0 * iq_buff_size : iq_buff_size / 4.0;
		const double budget = h->ndev > 1 ? 256.0 * 1048576.0 : 2048.0 * 1048576.0;
		double cap = budget / (eb > 1.0 ? eb : 1.0);
		batch = cap > 4096.0 ? 4096 : cap < 32.0 ? 32 : (int)cap;
		h->flush_at = batch < 256 ? batch : 256;
		if ((s = getenv("GPUSIM_BATCH_EPOCHS")) != NULL && atoi(s) > 0)
			h->flush_at = batch = atoi(s); /* fixed size */
		h->cap_max = batch;
	}
	cols_reserve(&h->batch, batch);
	h->batch.cap = batch;
	nav_new_batch(h);

	h->host_threads = default_host_threads();
	if ((s = getenv("GPUSIM_HOST_THREADS")) != NULL && atoi(s) > 0)
		h->host_threads = atoi(s) > 64 ? 64 : atoi(s);
	pool_start(&h->pool, h->host_threads - 1);
	h->ra.enabled = h->pool.n > 0;

	if (!h->dryrun)
	{
		gpusim_config cfg;
		int d;
		memset(&cfg, 0, sizeof(cfg));
		cfg.abi_version = GPUSIM_ABI_VERSION;
		cfg.samples_per_epoch = iq_buff_size;
		cfg.data_format = data_format;
		cfg.carrier_mode = h->carrier_mode;
		cfg.max_batch_epochs = batch;
		cfg.delt = delt;
		if (h->ndev == 1 && getenv("GPUSIM_DEVICE_LIST") == NULL)
			h->workers[0].device = (s = getenv("GPUSIM_DEVICE")) != NULL ? atoi(s) : 0;

		/* time-sharding over GPUs: epochs are independent given their rows, no exchange needed */
		pthread_mutex_init(&h->mu, NULL);
		pthread_cond_init(&h->cv, NULL);
		h->nslots = 2 * h->ndev;
		h->slots = calloc((size_t)h->nslots, sizeof(slot_t));
		if (h->slots == NULL)
			die("gpusim hook out of memory", NULL);
		h->cfg = cfg;
		h->epoch_bytes = data_format == 16 ? (size_t)4 * iq_buff_size : data_format == 8 ? (size_t)2 * iq_buff_size : (size_t)(iq_buff_size / 4);
		for (d = 0; d < h->ndev; d++)
		{
			h->workers[d].h = h;
			h->workers[d].index = d;
		}
		for (d = 0; d < h->nslots; d++)
		{
			cols_reserve(&h->slots[d].rows, batch);
			h->slots[d].rows.cap = batch;
		}
		for (d = 0; d < h->ndev; d++)
			pthread_create(&h->workers[d].thread, NULL, worker_main, &h->workers[d]);
		if (h->ndev > 1)
			pthread_create(&h->writer, NULL, writer_main, h);
	}
	h->t_open = now_s();
	return h;
}

void gpusim_hook_epoch(gpusim_hook *h, channel_t *chan, const int *gain)
{
	int i;
	size_t o = (size_t)h->batch.n * MAX_CHAN;
	cols_t *b = &h->batch;

	for (i = 0; i < MAX_CHAN; i++, o++)
	{
		if (chan[i].prn <= 0)
		{
			b->prn[o] = 0;
			b->f_code[o] = 0.0; b->code_phase[o] = 0.0; b->icode[o] = 0;
			b->nav_bits[o] = 0; b->gain[o] = 0; b->carr_phasestep[o] = 0;
			b->carr_phase[o] = 0; b->f_carr[o] = 0.0; b->carr_phase_f[o] = 0.0;
			b->iword[o] = 0; b->ibit[o] = 0; b->carr_init[o] = 0.0;
			b->nav_frame[o] = 0;
			h->nav[i].active = 0;
			continue;
		}

		b->prn[o] = chan[i].prn;
		b->f_code[o] = chan[i].f_code;
		b->code_phase[o] = chan[i].code_phase;
		b->icode[o] = chan[i].icode;
		b->nav_bits[o] = gpusim_pack_nav_bits(chan[i].dwrd, N_DWRD, chan[i].iword, chan[i].ibit);
		b->nav_frame[o] = h->nav_device ? nav_frame_of(h, i, &chan[i]) : 0;
		b->gain[o] = gain[i];
		b->iword[o] = chan[i].iword;
		b->ibit[o] = chan[i].ibit;
		b->f_carr[o] = chan[i].f_carr;

#ifdef FLOAT_CARR_PHASE
		b->carr_phasestep[o] = 0;
		b->carr_phase[o] = 0;
		/* The double carrier phase chains through every sample of every epoch (gpssim.c:2245-2250).
		 * Only note where a chain (re)starts; fill_carrier_chains() walks the chains of the whole
		 * batch, channel slots in parallel, before the batch is handed to the GPU. */
		b->carr_phase_f[o] = 0.0;
		b->carr_init[o] = chan[i].carr_phase;
		chan[i].carr_phase = HOOK_CONTINUE;
#else
		b->carr_phasestep[o] = chan[i].carr_phasestep;
		b->carr_phase[o] = chan[i].carr_phase;
		b->carr_phase_f[o] = 0.0;
		b->carr_init[o] = 0.0;
		/* N times "carr_phase += carr_phasestep" on an unsigned int (gpssim.c:2252) */
		chan[i].carr_phase += (unsigned int)h->N * (unsigned int)chan[i].carr_phasestep;
#endif
	}

	b->n++;
	if (b->n >= h->flush_at)
	{
		pipeline_check(h);
		flush_batch(h);
		h->flush_at = 2 * h->flush_at < h->cap_max ? 2 * h->flush_at : h->cap_max;
	}
}

/*
 * Dump format (little endian):
 *   char  magic[8] = "GPSTAB01"
 *   int32 n_epochs, samples_per_epoch, data_format, carrier_mode, max_chan, reserved
 *   double delt
 *   then n_epochs*max_chan entries of each column, in this order:
 *   prn i32, f_code f64, code_phase f64, icode i32, nav_bits u32, gain i32,
 *   carr_phasestep i32, carr_phase u32, f_carr f64, carr_phase_f f64, iword i32, ibit i32
 */
static void write_dump(const gpusim_hook *h)
{
	FILE *f = fopen(h->dump_path, "wb");
	int32_t hdr[6];
	size_t rows = (size_t)h->dump.n * MAX_CHAN;
	const cols_t *c = &h->dump;

	if (f == NULL)
		die("Failed to open table dump file", h->dump_path);
	hdr[0] = h->dump.n; hdr[1] = h->N; hdr[2] = h->fmt; hdr[3] = h->carrier_mode;
	hdr[4] = MAX_CHAN; hdr[5] = 0;
	fwrite("GPSTAB01", 1, 8, f);
	fwrite(hdr, sizeof(int32_t), 6, f);
	fwrite(&h->delt, sizeof(double), 1, f);
	fwrite(c->prn, sizeof(int32_t), rows, f);
	fwrite(c->f_code, sizeof(double), rows, f);
	fwrite(c->code_phase, sizeof(double), rows, f);
	fwrite(c->icode, sizeof(int32_t), rows, f);
	fwrite(c->nav_bits, sizeof(uint32_t), rows, f);
	fwrite(c->gain, sizeof(int32_t), rows, f);
	fwrite(c->carr_phasestep, sizeof(int32_t), rows, f);
	fwrite(c->carr_phase, sizeof(uint32_t), rows, f);
	fwrite(c->f_carr, sizeof(double), rows, f);
	fwrite(c->carr_phase_f, sizeof(double), rows, f);
	fwrite(c->iword, sizeof(int32_t), rows, f);
	fwrite(c->ibit, sizeof(int32_t), rows, f);
	fclose(f);
}

void gpusim_hook_close(gpusim_hook *h)
{
	const char *v;
	flush_batch(h);
	if (!h->dryrun)
	{
		int d;
		pipeline_join(h);
		pipeline_check(h); /* reports and exits if anything failed */
		for (d = 0; d < h->ndev; d++)
			if (h->workers[d].ctx != NULL)
				gpusim_destroy(h->workers[d].ctx);
		for (d = 0; d < h->nslots; d++)
		{
			cols_free(&h->slots[d].rows);
			if (h->slots[d].out != NULL)
				gpusim_host_free(h->slots[d].out);
		}
		free(h->slots);
	}
	pool_stop(&h->pool);
	if ((v = getenv("GPUSIM_VERBOSE")) != NULL && atoi(v) != 0)
		fprintf(stderr, "\ngpusim hook: %ld epochs, %d host threads, range look-ahead: %ld windows, %ld hits, %ld direct calls\n"
		                "gpusim hook: %.3f s from open to close (context creation on the worker: %.3f s, overlapped with the host's row pre-pass); main thread: %.3f s range windows, %.3f s carrier chains, %.3f s waiting for the GPU pipeline; batches up to %d epochs\n",
		        h->epochs_done, h->host_threads, h->ra.windows, h->ra.hits, h->ra.direct,
		        now_s() - h->t_open, h->t_ctx, h->t_ranges, h->t_chains, h->t_wait, h->cap_max);
	if (h->dump_path != NULL)
		write_dump(h);
	if (h->nav_dump != NULL)
		fclose(h->nav_dump);
	free(h->ra.rho);
	cols_free(&h->batch);
	cols_free(&h->dump);
	free(h);
}
