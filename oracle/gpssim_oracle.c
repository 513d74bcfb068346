/*
 * gpssim_oracle.c - CPU restatement of the reference's sample-synthesis path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may build, load or call
 * this file.  The product (gps_sdr_sim_b200/, include/, integration/) never does.
 *
 * What it restates: gps-sdr-sim's per-sample loop and output formatter,
 *   /root/reference/gpssim.c:2190-2264  (accumulate over channels, advance code
 *                                        phase / chip / data bit / carrier)
 *   /root/reference/gpssim.c:2266-2288  (SC01 / SC08 / SC16 conversion)
 * driven by the per-epoch rows defined in include/gpusim.h (the values the
 * reference's channel_t fields hold at gpssim.c:2190) instead of live channel_t
 * structs.  Both carrier branches of gpssim.h:4 are restated.
 *
 * Parity pinning: the reference ships no tests or golden vectors for this path
 * (SURVEY.md section 4), so this oracle is pinned against OUTPUTS OF THE
 * REFERENCE ITSELF: tests/test_oracle_vs_reference.py runs oracle/_ref/gps-sdr-sim-{int,float}
 * (the unmodified gpssim.c, built by oracle/build_ref.sh) and compares bytes, and
 * tests/golden/ holds reference-generated digests for machines without oracle/_ref.
 *
 * Deliberately written as the plainest possible scalar code - one sample, one
 * channel at a time, in the reference's order of operations - and compiled
 * without FMA contraction so that "code_phase += f_code*delt" is a rounded
 * multiply followed by a rounded add exactly as in the reference build
 * (Makefile:8: -O3, no -march).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "gpusim.h" /* only for the gpusim_epoch_table row layout */

#define CA_LEN 1023

/* ---- constant tables ------------------------------------------------------ */

/* sinTable512 / cosTable512 (gpssim.c:15-83): amplitude-250 sine at bin centres,
 * cos = sin advanced by 128 entries.  249.9925 instead of 250 keeps the one
 * borderline entry (index 35 and its mirrors) on the reference's side; see
 * tests/test_tables.py which compares all 1024 values with the reference's arrays. */
void oracle_carrier_lut(int *sin512, int *cos512)
{
	int i;
	for (i = 0; i < 512; i++)
		sin512[i] = (int)lround(249.9925 * sin(2.0 * M_PI * ((double)i + 0.5) / 512.0));
	for (i = 0; i < 512; i++)
		cos512[i] = sin512[(i + 128) % 512];
}

/* codegen() (gpssim.c:132-171) restated through the output recurrences of the
 * two ICD-GPS-200 generators: G1[n] = G1[n-3]^G1[n-10], G2[n] = G2[n-2]^G2[n-3]^
 * G2[n-6]^G2[n-8]^G2[n-9]^G2[n-10], both starting with ten ones, G2 delayed. */
int oracle_ca_code(int prn, int *ca)
{
	static const int delay[32] = {5, 6, 7, 8, 17, 18, 139, 140, 141, 251, 252, 254, 255, 256, 257,
	                              258, 469, 470, 471, 472, 473, 474, 509, 512, 513, 514, 515, 516,
	                              859, 860, 861, 862};
	int g1[CA_LEN], g2[CA_LEN], n;
	if (prn < 1 || prn > 32)
		return -1;
	for (n = 0; n < CA_LEN; n++)
	{
		if (n < 10)
			g1[n] = g2[n] = 1;
		else
		{
			g1[n] = g1[n - 3] ^ g1[n - 10];
			g2[n] = g2[n - 2] ^ g2[n - 3] ^ g2[n - 6] ^ g2[n - 8] ^ g2[n - 9] ^ g2[n - 10];
		}
	}
	for (n = 0; n < CA_LEN; n++)
		ca[n] = g1[n] ^ g2[(n + CA_LEN - delay[prn - 1]) % CA_LEN];
	return 0;
}

/* ---- one epoch ------------------------------------------------------------ */

typedef struct
{
	int active;
	int ca[CA_LEN];
	double f_code, code_phase;
	int icode, ibitpos; /* ibitpos: how many data bits consumed since the row */
	uint32_t nav_bits;
	int dataBit, codeCA, gain;
	unsigned int carr_phase;
	int carr_phasestep;
	double f_carr, carr_phase_f;
} och_t;

static int nav_bit(uint32_t nav_bits, int k) /* k-th bit from the row's start, as +-1 */
{
	return k < 32 ? (int)((nav_bits >> (31 - k)) & 1u) * 2 - 1 : -1;
}

static void epoch_samples(och_t *ch, int nch, int N, double delt, int mode, const int *sinT,
                          const int *cosT, short *iq)
{
	int isamp, i;
	for (isamp = 0; isamp < N; isamp++)
	{
		int i_acc = 0, q_acc = 0;
		for (i = 0; i < nch; i++)
		{
			och_t *c = &ch[i];
			int iTable;
			if (!c->active)
				continue;
			/* gpssim.c:2199-2203 */
			if (mode == GPUSIM_CARRIER_FLOAT)
				iTable = (int)floor(c->carr_phase_f * 512.0);
			else
				iTable = (c->carr_phase >> 16) & 0x1ff;
			/* gpssim.c:2204-2209 */
			i_acc += c->dataBit * c->codeCA * cosT[iTable] * c->gain;
			q_acc += c->dataBit * c->codeCA * sinT[iTable] * c->gain;
			/* gpssim.c:2212-2238 */
			{
				volatile double inc = c->f_code * delt; /* rounded product, then rounded sum */
				c->code_phase += inc;
			}
			if (c->code_phase >= CA_LEN)
			{
				c->code_phase -= CA_LEN;
				c->icode++;
				if (c->icode >= 20)
				{
					c->icode = 0;
					c->ibitpos++;
					c->dataBit = nav_bit(c->nav_bits, c->ibitpos);
				}
			}
			/* gpssim.c:2241 */
			c->codeCA = c->ca[(int)c->code_phase] * 2 - 1;
			/* gpssim.c:2244-2253 */
			if (mode == GPUSIM_CARRIER_FLOAT)
			{
				volatile double inc = c->f_carr * delt;
				c->carr_phase_f += inc;
				if (c->carr_phase_f >= 1.0)
					c->carr_phase_f -= 1.0;
				else if (c->carr_phase_f < 0.0)
					c->carr_phase_f += 1.0;
			}
			else
				c->carr_phase += (unsigned int)c->carr_phasestep;
		}
		/* gpssim.c:2258-2263 */
		i_acc = (i_acc + 64) >> 7;
		q_acc = (q_acc + 64) >> 7;
		iq[isamp * 2] = (short)i_acc;
		iq[isamp * 2 + 1] = (short)q_acc;
	}
}

size_t oracle_epoch_bytes(int N, int fmt)
{
	if (fmt == GPUSIM_SC01)
		return (size_t)(N / 4);
	if (fmt == GPUSIM_SC08)
		return (size_t)2 * N;
	return (size_t)4 * N;
}

/* gpssim.c:2266-2288 */
static void epoch_format(const short *iq, int N, int fmt, unsigned char *out)
{
	int isamp;
	if (fmt == GPUSIM_SC01)
	{
		/* the reference ORs into iq8_buff[isamp/8] for isamp < 2N but writes N/4 bytes */
		int nbytes = N / 4;
		memset(out, 0, (size_t)nbytes);
		for (isamp = 0; isamp < 2 * N; isamp++)
			if (isamp / 8 < nbytes)
				out[isamp / 8] |= (unsigned char)((iq[isamp] > 0 ? 0x01 : 0x00) << (7 - isamp % 8));
	}
	else if (fmt == GPUSIM_SC08)
	{
		for (isamp = 0; isamp < 2 * N; isamp++)
			((signed char *)out)[isamp] = (signed char)(iq[isamp] >> 4);
	}
	else
		memcpy(out, iq, (size_t)4 * N);
}

/* one epoch: rows -> bytes */
static int one_epoch(const gpusim_epoch_table *t, int e, int N, double delt, int fmt, int mode,
                     const int *sinT, const int *cosT, unsigned char *out)
{
	och_t *ch = (och_t *)calloc(GPUSIM_MAX_CHAN, sizeof(och_t));
	short *iq = (short *)malloc((size_t)4 * N);
	int i;
	if (ch == NULL || iq == NULL)
	{
		free(ch);
		free(iq);
		return -1;
	}
	for (i = 0; i < GPUSIM_MAX_CHAN; i++)
	{
		size_t r = (size_t)e * GPUSIM_MAX_CHAN + i;
		och_t *c = &ch[i];
		if (t->prn[r] <= 0 || oracle_ca_code(t->prn[r], c->ca) != 0)
			continue;
		c->active = 1;
		c->f_code = t->f_code[r];
		c->code_phase = t->code_phase[r];
		c->icode = t->icode[r];
		c->ibitpos = 0;
		c->nav_bits = t->nav_bits[r];
		c->gain = t->gain[r];
		/* initial chip and data bit as computeCodePhase leaves them (gpssim.c:1344-1345) */
		c->codeCA = c->ca[(int)c->code_phase] * 2 - 1;
		c->dataBit = nav_bit(c->nav_bits, 0);
		if (mode == GPUSIM_CARRIER_FLOAT)
		{
			c->f_carr = t->f_carr[r];
			c->carr_phase_f = t->carr_phase_f[r];
		}
		else
		{
			c->carr_phase = t->carr_phase[r];
			c->carr_phasestep = t->carr_phasestep[r];
		}
	}
	epoch_samples(ch, GPUSIM_MAX_CHAN, N, delt, mode, sinT, cosT, iq);
	epoch_format(iq, N, fmt, out);
	free(ch);
	free(iq);
	return 0;
}

typedef struct
{
	const gpusim_epoch_table *t;
	int first, count, N, fmt, mode, tid, nthreads, rc;
	double delt;
	const int *sinT, *cosT;
	unsigned char *out;
} job_t;

static void *worker(void *arg)
{
	job_t *j = (job_t *)arg;
	size_t eb = oracle_epoch_bytes(j->N, j->fmt);
	int k;
	for (k = j->tid; k < j->count; k += j->nthreads) /* epochs are independent given their rows */
		if (one_epoch(j->t, j->first + k, j->N, j->delt, j->fmt, j->mode, j->sinT, j->cosT,
		              j->out + (size_t)k * eb) != 0)
			j->rc = -1;
	return NULL;
}

/*
 * Generate epochs [first, first+count) of the table into out (count*epoch_bytes).
 * nthreads <= 1: the calling thread does everything (the reference is single-threaded);
 * otherwise epochs are dealt round-robin to nthreads pthreads.  Returns 0 or -1.
 */
int oracle_generate_epochs(const gpusim_epoch_table *t, int first, int count, int N, double delt,
                           int fmt, int mode, int nthreads, unsigned char *out)
{
	int sinT[512], cosT[512];
	job_t jobs[64];
	pthread_t th[64];
	int i, rc = 0;
	if (t == NULL || first < 0 || count < 0 || first + count > t->n_epochs || N <= 0)
		return -1;
	oracle_carrier_lut(sinT, cosT);
	if (nthreads < 1)
		nthreads = 1;
	if (nthreads > 64)
		nthreads = 64;
	for (i = 0; i < nthreads; i++)
	{
		job_t j = {t, first, count, N, fmt, mode, i, nthreads, 0, delt, sinT, cosT, out};
		jobs[i] = j;
	}
	if (nthreads == 1)
		worker(&jobs[0]);
	else
	{
		for (i = 0; i < nthreads; i++)
			pthread_create(&th[i], NULL, worker, &jobs[i]);
		for (i = 0; i < nthreads; i++)
			pthread_join(th[i], NULL);
	}
	for (i = 0; i < nthreads; i++)
		if (jobs[i].rc != 0)
			rc = -1;
	return rc;
}

/*
 * The sequential code-phase recurrence alone (gpssim.c:2212-2218), for checking
 * the device's checkpoint kernel: writes code_phase and the number of 1023-chip
 * wraps seen so far at every sample index that is a multiple of `every`.
 * n_out = ceil(N/every) entries.
 */
void oracle_code_phase_checkpoints(double code_phase, double f_code, double delt, int N, int every,
                                   double *x_out, int *wraps_out)
{
	int n, wraps = 0;
	for (n = 0; n < N; n++)
	{
		if (n % every == 0)
		{
			x_out[n / every] = code_phase;
			wraps_out[n / every] = wraps;
		}
		{
			volatile double inc = f_code * delt;
			code_phase += inc;
		}
		if (code_phase >= CA_LEN)
		{
			code_phase -= CA_LEN;
			wraps++;
		}
	}
}

/*
 * The double carrier-phase recurrence alone (FLOAT_CARR_PHASE branch, gpssim.c:2244-2250), for
 * checking the device's carrier checkpoint chain: writes carr_phase at every sample index that is a
 * multiple of `every` and returns the value after N samples (what the next epoch starts from).
 */
double oracle_carrier_phase_checkpoints(double carr_phase, double f_carr, double delt, int N, int every,
                                        double *x_out)
{
	int n;
	for (n = 0; n < N; n++)
	{
		if (n % every == 0)
			x_out[n / every] = carr_phase;
		{
			volatile double inc = f_carr * delt;
			carr_phase += inc;
		}
		if (carr_phase >= 1.0)
			carr_phase -= 1.0;
		else if (carr_phase < 0.0)
			carr_phase += 1.0;
	}
	return carr_phase;
}

/*
 * Navigation data words: what generateNavMsg() (gpssim.c:1467-1547) leaves in
 * chan->dwrd[60], restated from IS-GPS-200 20.3.5.2 instead of from the reference's mask arithmetic
 * (computeChecksum, gpssim.c:693-756): every parity bit is the XOR of the listed data bits d1..d24
 * (Table 20-XIV) and of D29* or D30* of the word before; D30* inverts the data bits; in the words whose
 * bits 23/24 carry no information (words 2 and 10 of a subframe) those two are chosen so that D29 = D30 = 0.
 * Pinned against the reference's own generateNavMsg() through oracle/_ref/libgpssim_ref_int.so
 * (tests/test_navmsg.py).
 *
 *   sbf50      the five subframes from eph2sbf() (chan->sbf, gpssim.h:174), 24 source bits in bits 29..6
 *   first10    source words of dwrd[0..9]: chan->sbf[4] of the call that built the previous frame
 *              (init = 1: of this call, gpssim.c:1484-1503)
 *   tow_first  TOW count in their hand-over word; tow: ((unsigned long)g0.sec)/6; wn: g0.week%1024
 */
static const signed char nav_parity_taps[6][16] = {
	{1, 2, 3, 5, 6, 10, 11, 12, 13, 14, 17, 18, 20, 23, 0},          /* D25, with D29* */
	{2, 3, 4, 6, 7, 11, 12, 13, 14, 15, 18, 19, 21, 24, 0},          /* D26, with D30* */
	{1, 3, 4, 5, 7, 8, 12, 13, 14, 15, 16, 19, 20, 22, 0},           /* D27, with D29* */
	{2, 4, 5, 6, 8, 9, 13, 14, 15, 16, 17, 20, 21, 23, 0},           /* D28, with D30* */
	{1, 3, 5, 6, 7, 9, 10, 14, 15, 16, 17, 18, 21, 22, 24, 0},       /* D29, with D30* */
	{3, 5, 6, 8, 9, 10, 11, 13, 15, 19, 22, 23, 24, 0},              /* D30, with D29* */
};
static const int nav_parity_uses_d29[6] = {1, 0, 1, 0, 0, 1};

static int nav_parity_bit(int row, const int *d /* d[1..24] */, int d29s, int d30s)
{
	int p = nav_parity_uses_d29[row] ? d29s : d30s, k;
	for (k = 0; nav_parity_taps[row][k]; k++)
		p ^= d[(int)nav_parity_taps[row][k]];
	return p;
}

static uint32_t nav_transmitted_word(uint32_t source24 /* d1 in bit 23 */, int d29s, int d30s, int free_tail)
{
	int d[25], k;
	uint32_t w = 0;
	for (k = 1; k <= 24; k++)
		d[k] = (int)((source24 >> (24 - k)) & 1u);
	if (free_tail)
	{
		/* d24 enters D29 and D30, d23 only D30: settle D29 with d24, then D30 with d23 */
		if (nav_parity_bit(4, d, d29s, d30s))
			d[24] ^= 1;
		if (nav_parity_bit(5, d, d29s, d30s))
			d[23] ^= 1;
	}
	for (k = 1; k <= 24; k++)
		w = (w << 1) | (uint32_t)(d[k] ^ d30s);
	for (k = 0; k < 6; k++)
		w = (w << 1) | (uint32_t)nav_parity_bit(k, d, d29s, d30s);
	return w;
}

void oracle_nav_frame(const uint32_t *sbf50, const uint32_t *first10, uint32_t tow_first, uint32_t tow, uint32_t wn,
                      uint32_t *dwrd60)
{
	int sub, i;
	uint32_t before = 0; /* the reference starts a frame from prevwrd = 0 or from a word ending in two zero bits */
	for (sub = 0; sub < 6; sub++)
	{
		const uint32_t *src = sub == 0 ? first10 : sbf50 + 10 * (sub - 1);
		const uint32_t count = sub == 0 ? tow_first : tow + (uint32_t)sub;
		for (i = 0; i < 10; i++)
		{
			uint32_t s = src[i];
			if (sub == 1 && i == 2)
				s |= (wn & 0x3ffu) << 20;
			if (i == 1)
				s |= (count & 0x1ffffu) << 13;
			before = nav_transmitted_word((s >> 6) & 0xffffffu, (int)(((before >> 1) | (s >> 31)) & 1u),
			                              (int)((before | (s >> 30)) & 1u), i == 1 || i == 9);
			dwrd60[10 * sub + i] = before;
		}
	}
}
