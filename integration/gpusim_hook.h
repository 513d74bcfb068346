/*
 * gpusim_hook.h - the reference-side binding for libgpusim.
 *
 * This header and gpusim_hook.c are what a gps-sdr-sim maintainer adds next to
 * gpssim.c.  It is compiled TOGETHER WITH the reference's own gpssim.h (it uses
 * channel_t, MAX_CHAN, N_DWRD, FLOAT_CARR_PHASE from there), so it follows
 * whichever carrier-phase branch the host is built with (gpssim.h:4).
 *
 * The three macros are the complete edit to main():
 *
 *   GPUSIM_HOOK_OPEN()   before the epoch loop        (before gpssim.c:2149)
 *   GPUSIM_HOOK_EPOCH()  instead of the sample loop + format + fwrite
 *                                                     (replaces gpssim.c:2190-2288)
 *   GPUSIM_HOOK_CLOSE()  after the epoch loop         (before gpssim.c:2355)
 *
 * They use main()'s own locals: chan, gain, delt, iq_buff_size, data_format, fp.
 *
 * Optional fourth edit, for long runs whose row pre-pass would otherwise limit the GPU:
 *
 *   GPUSIM_HOOK_RANGE()  instead of the if/else pair of computeRange() calls of the epoch loop
 *                                                     (replaces gpssim.c:2165-2168)
 *
 * looks the range up in a window computed ahead of time by the reference's own computeRange(),
 * in parallel over epochs (uses main()'s i, iumd, numd, eph, ieph, ionoutc, grx, xyz,
 * staticLocationMode and the local `rho`).
 */
#ifndef GPUSIM_HOOK_H
#define GPUSIM_HOOK_H

#include <stdio.h>

typedef struct gpusim_hook gpusim_hook;

gpusim_hook *gpusim_hook_open(int iq_buff_size, double delt, int data_format, FILE *fp);
void gpusim_hook_epoch(gpusim_hook *h, channel_t *chan, const int *gain);
void gpusim_hook_close(gpusim_hook *h);
void gpusim_hook_range(gpusim_hook *h, range_t *rho, int slot, int iumd, int numd, const channel_t *chan,
                       const ephem_t *eph_set, ionoutc_t *ionoutc, gpstime_t grx, double (*xyz)[3], double *xyz0);

#define GPUSIM_HOOK_OPEN() \
    gpusim_hook *gpusim_h = gpusim_hook_open(iq_buff_size, delt, data_format, fp)
#define GPUSIM_HOOK_EPOCH() gpusim_hook_epoch(gpusim_h, chan, gain)
#define GPUSIM_HOOK_CLOSE() gpusim_hook_close(gpusim_h)
#define GPUSIM_HOOK_RANGE() \
    gpusim_hook_range(gpusim_h, &rho, i, iumd, numd, chan, eph[ieph], &ionoutc, grx, \
                      staticLocationMode ? (double (*)[3])0 : xyz, xyz[0])

#endif
