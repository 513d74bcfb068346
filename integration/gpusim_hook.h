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
 */
#ifndef GPUSIM_HOOK_H
#define GPUSIM_HOOK_H

#include <stdio.h>

typedef struct gpusim_hook gpusim_hook;

gpusim_hook *gpusim_hook_open(int iq_buff_size, double delt, int data_format, FILE *fp);
void gpusim_hook_epoch(gpusim_hook *h, channel_t *chan, const int *gain);
void gpusim_hook_close(gpusim_hook *h);

#define GPUSIM_HOOK_OPEN() \
    gpusim_hook *gpusim_h = gpusim_hook_open(iq_buff_size, delt, data_format, fp)
#define GPUSIM_HOOK_EPOCH() gpusim_hook_epoch(gpusim_h, chan, gain)
#define GPUSIM_HOOK_CLOSE() gpusim_hook_close(gpusim_h)

#endif
