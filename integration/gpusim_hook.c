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
 *   GPUSIM_BATCH_EPOCHS  epochs per library call            (default 256)
 *   GPUSIM_DEVICE        CUDA device ordinal                (default 0)
 *   GPUSIM_DEVICES       number of GPUs (devices 0..n-1) to time-shard batches over (default 1):
 *                        batch b goes to device b mod n; one worker thread per GPU generates into
 *                        page-locked buffers, one writer thread fwrites the batches in order
 *   GPUSIM_DEVICE_LIST   same, with explicit ordinals, e.g. "0,2,5" (or "0,0" to run two workers
 *                        on one GPU)
 *   GPUSIM_DUMP          path: also write every table row to this file
 *                        (format below) - how tests/golden/ fixtures are made
 *   GPUSIM_DRYRUN        1: record (and dump) rows, generate nothing, write nothing
 */
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

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
	int32_t *iword; /* diagnostics only (dump) */
	int32_t *ibit;
} cols_t;

/* multi-GPU pipeline: a ring of batch slots; slot states advance FREE -> QUEUED -> DONE -> FREE */
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
	int index; /* worker d serves batches d, d+n, d+2n, ... */
	int device;
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
	/* multi-GPU (ndev > 1) */
	int ndev, nslots;
	slot_t *slots;
	worker_t workers[HOOK_MAX_DEV];
	pthread_t writer;
	pthread_mutex_t mu;
	pthread_cond_t cv;
	long seq_filled; /* batches handed to the workers so far */
	int finishing;   /* no more batches will be queued */
};

static void die(const char *what, const char *detail)
{
	fprintf(stderr, "\nERROR: %s%s%s\n", what, detail ? ": " : "", detail ? detail : "");
	exit(1);
}

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
	c->cap = epochs;
}

static void cols_free(cols_t *c)
{
	free(c->prn); free(c->f_code); free(c->code_phase); free(c->icode);
	free(c->nav_bits); free(c->gain); free(c->carr_phasestep); free(c->carr_phase);
	free(c->f_carr); free(c->carr_phase_f); free(c->iword); free(c->ibit);
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
	CP(carr_phasestep); CP(carr_phase); CP(f_carr); CP(carr_phase_f); CP(iword); CP(ibit);
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

static void table_of(const cols_t *c, gpusim_epoch_table *t)
{
	memset(t, 0, sizeof(*t));
	t->n_epochs = c->n;
	t->prn = c->prn;
	t->f_code = c->f_code;
	t->code_phase = c->code_phase;
	t->icode = c->icode;
	t->nav_bits = c->nav_bits;
	t->gain = c->gain;
	t->carr_phasestep = c->carr_phasestep;
	t->carr_phase = c->carr_phase;
	t->f_carr = c->f_carr;
	t->carr_phase_f = c->carr_phase_f;
}

/* ---- multi-GPU pipeline ------------------------------------------------------------------ */
static void *worker_main(void *arg)
{
	worker_t *w = (worker_t *)arg;
	gpusim_hook *h = w->h;
	long seq;
	for (seq = w->index;; seq += h->ndev)
	{
		slot_t *s = &h->slots[seq % h->nslots];
		gpusim_epoch_table t;
		int rc;
		pthread_mutex_lock(&h->mu);
		while (!(s->state == SLOT_QUEUED && seq < h->seq_filled) && !(h->finishing && seq >= h->seq_filled))
			pthread_cond_wait(&h->cv, &h->mu);
		if (seq >= h->seq_filled)
		{
			pthread_mutex_unlock(&h->mu);
			return NULL;
		}
		pthread_mutex_unlock(&h->mu);

		table_of(&s->rows, &t);
		rc = gpusim_generate_epochs(w->ctx, &t, s->out, (size_t)h->batch.cap * gpusim_epoch_bytes(w->ctx));
		if (rc != GPUSIM_OK)
			die("GPU sample generation failed", gpusim_last_error(w->ctx));

		pthread_mutex_lock(&h->mu);
		s->state = SLOT_DONE;
		pthread_cond_broadcast(&h->cv);
		pthread_mutex_unlock(&h->mu);
	}
}

static void *writer_main(void *arg)
{
	gpusim_hook *h = (gpusim_hook *)arg;
	const size_t eb = gpusim_epoch_bytes(h->workers[0].ctx);
	long seq;
	for (seq = 0;; seq++)
	{
		slot_t *s = &h->slots[seq % h->nslots];
		pthread_mutex_lock(&h->mu);
		while (!(s->state == SLOT_DONE && seq < h->seq_filled) && !(h->finishing && seq >= h->seq_filled))
			pthread_cond_wait(&h->cv, &h->mu);
		if (seq >= h->seq_filled)
		{
			pthread_mutex_unlock(&h->mu);
			return NULL;
		}
		pthread_mutex_unlock(&h->mu);

		/* strictly in batch order: this is the reference's output file */
		if (fwrite(s->out, 1, (size_t)s->rows.n * eb, h->fp) != (size_t)s->rows.n * eb)
			die("Failed to write the output file", NULL);

		pthread_mutex_lock(&h->mu);
		s->state = SLOT_FREE;
		pthread_cond_broadcast(&h->cv);
		pthread_mutex_unlock(&h->mu);
	}
}

/* hand the filled batch to the next GPU: swap it into a free ring slot */
static void queue_batch(gpusim_hook *h)
{
	slot_t *s = &h->slots[h->seq_filled % h->nslots];
	cols_t tmp;
	pthread_mutex_lock(&h->mu);
	while (s->state != SLOT_FREE)
		pthread_cond_wait(&h->cv, &h->mu);
	tmp = s->rows;
	s->rows = h->batch;
	h->batch = tmp;
	h->batch.n = 0;
	s->state = SLOT_QUEUED;
	h->seq_filled++;
	pthread_cond_broadcast(&h->cv);
	pthread_mutex_unlock(&h->mu);
}

static void flush_batch(gpusim_hook *h)
{
	gpusim_epoch_table t;
	int rc;

	if (h->batch.n == 0)
		return;

	if (!h->dryrun && h->ndev > 1)
	{
		h->epochs_done += h->batch.n;
		queue_batch(h);
		return;
	}

	if (!h->dryrun)
	{
		table_of(&h->batch, &t);

		rc = gpusim_generate_epochs_to_sink(h->ctx, &t, sink_fwrite, h);
		if (rc != GPUSIM_OK)
			die("GPU sample generation failed", gpusim_last_error(h->ctx));
	}

	h->epochs_done += h->batch.n;
	h->batch.n = 0;
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
	if ((s = getenv("GPUSIM_BATCH_EPOCHS")) != NULL && atoi(s) > 0)
		batch = atoi(s);
	h->dryrun = ((s = getenv("GPUSIM_DRYRUN")) != NULL && atoi(s) != 0);
	h->dump_path = getenv("GPUSIM_DUMP");

	cols_reserve(&h->batch, batch);
	h->batch.cap = batch; /* fixed: flush when full */

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

	if (!h->dryrun)
	{
		gpusim_config cfg;
		int rc, d;
		memset(&cfg, 0, sizeof(cfg));
		cfg.abi_version = GPUSIM_ABI_VERSION;
		cfg.device = (s = getenv("GPUSIM_DEVICE")) != NULL ? atoi(s) : 0;
		cfg.samples_per_epoch = iq_buff_size;
		cfg.data_format = data_format;
		cfg.carrier_mode = h->carrier_mode;
		cfg.max_batch_epochs = batch;
		cfg.delt = delt;
		if (h->ndev == 1)
		{
			rc = gpusim_create(&cfg, &h->ctx);
			if (rc != GPUSIM_OK)
				die("Failed to initialise the GPU sample generator", gpusim_last_error(NULL));
		}
		else
		{
			/* time-sharding over GPUs: epochs are independent given their rows, no exchange needed */
			pthread_mutex_init(&h->mu, NULL);
			pthread_cond_init(&h->cv, NULL);
			h->nslots = 2 * h->ndev;
			h->slots = calloc((size_t)h->nslots, sizeof(slot_t));
			if (h->slots == NULL)
				die("gpusim hook out of memory", NULL);
			for (d = 0; d < h->ndev; d++)
			{
				cfg.device = h->workers[d].device;
				rc = gpusim_create(&cfg, &h->workers[d].ctx);
				if (rc != GPUSIM_OK)
					die("Failed to initialise the GPU sample generator", gpusim_last_error(NULL));
				h->workers[d].h = h;
				h->workers[d].index = d;
			}
			for (d = 0; d < h->nslots; d++)
			{
				cols_reserve(&h->slots[d].rows, batch);
				h->slots[d].rows.cap = batch;
				h->slots[d].out = gpusim_host_alloc((size_t)batch * gpusim_epoch_bytes(h->workers[0].ctx));
				if (h->slots[d].out == NULL)
					die("Failed to allocate page-locked output buffers", NULL);
			}
			for (d = 0; d < h->ndev; d++)
				pthread_create(&h->workers[d].thread, NULL, worker_main, &h->workers[d]);
			pthread_create(&h->writer, NULL, writer_main, h);
		}
	}
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
			b->iword[o] = 0; b->ibit[o] = 0;
			continue;
		}

		b->prn[o] = chan[i].prn;
		b->f_code[o] = chan[i].f_code;
		b->code_phase[o] = chan[i].code_phase;
		b->icode[o] = chan[i].icode;
		b->nav_bits[o] = gpusim_pack_nav_bits(chan[i].dwrd, N_DWRD, chan[i].iword, chan[i].ibit);
		b->gain[o] = gain[i];
		b->iword[o] = chan[i].iword;
		b->ibit[o] = chan[i].ibit;
		b->f_carr[o] = chan[i].f_carr;

#ifdef FLOAT_CARR_PHASE
		b->carr_phasestep[o] = 0;
		b->carr_phase[o] = 0;
		b->carr_phase_f[o] = chan[i].carr_phase;
		/* The double carrier phase chains through every sample of every epoch (gpssim.c:2245-2250);
		 * keep the host's copy exact: the library walks the N updates in O(carrier cycles). */
		chan[i].carr_phase = gpusim_advance_carrier_f64(chan[i].carr_phase, chan[i].f_carr, h->delt, h->N);
#else
		b->carr_phasestep[o] = chan[i].carr_phasestep;
		b->carr_phase[o] = chan[i].carr_phase;
		b->carr_phase_f[o] = 0.0;
		/* N times "carr_phase += carr_phasestep" on an unsigned int (gpssim.c:2252) */
		chan[i].carr_phase += (unsigned int)h->N * (unsigned int)chan[i].carr_phasestep;
#endif
	}

	if (h->dump_path != NULL)
		cols_append(&h->dump, b, b->n);

	b->n++;
	if (b->n >= b->cap)
		flush_batch(h);
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
	flush_batch(h);
	if (!h->dryrun && h->ndev > 1)
	{
		int d;
		pthread_mutex_lock(&h->mu);
		h->finishing = 1;
		pthread_cond_broadcast(&h->cv);
		pthread_mutex_unlock(&h->mu);
		for (d = 0; d < h->ndev; d++)
			pthread_join(h->workers[d].thread, NULL);
		pthread_join(h->writer, NULL);
		for (d = 0; d < h->ndev; d++)
			gpusim_destroy(h->workers[d].ctx);
		for (d = 0; d < h->nslots; d++)
		{
			cols_free(&h->slots[d].rows);
			gpusim_host_free(h->slots[d].out);
		}
		free(h->slots);
	}
	if (h->dump_path != NULL)
		write_dump(h);
	if (h->ctx != NULL)
		gpusim_destroy(h->ctx);
	cols_free(&h->batch);
	cols_free(&h->dump);
	free(h);
}
