/*
 * gpu_ledger.h -- who has put what on which GPU, across processes and across schedulers.
 *
 * The reference is a one-GPU system (device 0 hard-coded, src/client.c:386; one daemon per
 * GPU, README.md:97), so "how much of this GPU is spoken for" is a per-process sum there
 * (`sum_allocated`, src/hook.c:77-78, 662).  With the peer-HBM backing tier a
 * client of GPU 0's scheduler occupies HBM of GPU 1, which may have a scheduler and clients of
 * its own: that memory has to be counted where it lives -- "peers' HBM used as backing must be
 * accounted in those GPUs' caps" (Scheduling).
 *
 * One small file per user and machine (not per scheduler socket: the lender and the borrower
 * talk to different daemons) holds a table of claims {pid, GPU, kind, bytes} under a robust
 * process-shared mutex.  GPUs are identified by UUID, not by ordinal (CUDA_VISIBLE_DEVICES
 * renumbers them per process).  Two kinds of claim:
 *   LENT  backing arenas this process has created on a GPU it does not compute on
 *   OWN   swappable memory this process has allocated on the GPU it computes on, i.e. what it
 *         needs resident there whenever it holds that GPU's lock
 * and two rules, checked under the one mutex:
 *   lend(d, n):   lent(d) + n  <=  total(d) - reserve - max over live processes of own(d)
 *   cap on d:     what cuMemGetInfo / the cuMemAlloc cap check of a client computing on d see
 *                 is total(d) - reserve - lent(d)            (hook.c)
 * Claims of processes that died are dropped when somebody is refused, and about once a second
 * by whoever asks for a sum.  Everything here is advisory bookkeeping on host memory: no call
 * ever blocks on another process beyond the mutex, and a ledger that cannot be used (foreign
 * file, other pid namespace, NVSHARE_GPU_LEDGER=off) switches the accounting off, loudly,
 * never the data path.
 */
#ifndef NVS_GPU_LEDGER_H
#define NVS_GPU_LEDGER_H

#include <stdint.h>

enum { NVS_GL_LENT = 1, NVS_GL_OWN = 2 };

#pragma GCC visibility push(hidden) /* internal to libnvshare.so / libnvs_engine.so */

/* Registers a GPU (by UUID) and returns its index in the ledger, or -1 when there is no usable
 * ledger (every other call then accepts -1 and does nothing / allows everything). */
int nvs_gl_device(const uint8_t uuid[16], uint64_t total_bytes);

/* Claim `bytes` of backing on GPU `dev`.  0 = granted, -1 = that GPU has no room to lend. */
int nvs_gl_lend(int dev, uint64_t bytes, uint64_t reserve_bytes);
void nvs_gl_return(int dev, uint64_t bytes);

/* This process now holds `delta` more (or fewer) bytes of swappable memory on its own GPU. */
void nvs_gl_own(int dev, int64_t delta);

/* Bytes of `dev` that processes computing elsewhere use as backing (all live lenders). */
uint64_t nvs_gl_lent(int dev);
/* Largest OWN claim on `dev` among live processes. */
uint64_t nvs_gl_max_own(int dev);
/* This process's own claims, for tests and the stats line. */
uint64_t nvs_gl_mine(int dev, int kind);

#pragma GCC visibility pop

#endif
