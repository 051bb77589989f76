/*
 * client.h -- seam between the interposer (hook.c) and the client runtime
 * (client.c).  Mirrors the reference's src/client.h:20-21
 * (initialize_client, continue_with_lock) and adds the two callbacks through
 * which the runtime drives the swap engine at the hand-off points SURVEY 3.4
 * names: fetch between receipt of LOCK_OK and own_lock = 1
 * (reference src/client.c:301-304), evict after cuda_sync_context() on
 * DROP_LOCK / early release (reference src/client.c:313-315, 472-476).
 */
#ifndef NVS_CLIENT_H
#define NVS_CLIENT_H

#include <stdint.h>
#include "cuda_min.h"

/* Driver entry points the runtime needs; filled in by hook.c before nvs_client_start(). */
#pragma GCC visibility push(hidden)
struct nvs_client_driver {
	CUresult (*cuInit)(unsigned);
	CUresult (*cuCtxGetCurrent)(CUcontext *);
	CUresult (*cuCtxSetCurrent)(CUcontext);
	CUresult (*cuCtxSynchronize)(void);
	/* NVML, all three or none (reference src/hook.c:111-143) */
	nvmlReturn_t (*nvmlInit)(void);
	nvmlReturn_t (*nvmlDeviceGetHandleByIndex)(unsigned, nvmlDevice_t *);
	nvmlReturn_t (*nvmlDeviceGetUtilizationRates)(nvmlDevice_t, nvmlUtilization_t *);
	/* Optional, all three or the reference's device 0: which physical GPU the application's context is
	 * on.  NVML numbers the physical GPUs and does not know CUDA_VISIBLE_DEVICES, so "index 0"
	 * (reference src/client.c:386) is somebody else's GPU for every client that computes elsewhere. */
	nvmlReturn_t (*nvmlDeviceGetHandleByUUID)(const char *, nvmlDevice_t *);
	CUresult (*cuCtxGetDevice)(CUdevice *);
	CUresult (*cuDeviceGetUuid)(uint8_t uuid[16], CUdevice);
};

/* Data-path callbacks; either may be NULL (pure UVM mode). Return 0 on success. */
struct nvs_client_datapath {
	int (*fetch_all)(void);                 /* make every allocation resident            */
	int (*evict)(uint64_t min_bytes);       /* release HBM (0 = everything)              */
	uint64_t (*nonresident_mib)(void);      /* for the REQ_LOCK "n<MiB>" hint            */
	void (*lock_state)(int holds_lock);     /* told whenever own_lock changes            */
	uint64_t (*free_hbm_mib)(void);         /* what the driver reports free right now    */
	uint64_t (*total_hbm_mib)(void);
	int (*evict_best_effort)(uint64_t min_bytes); /* same, but never waits for backing space */
	void (*evict_announce)(void);           /* an eviction for the next holder is about to start */
};

void nvs_client_start(const struct nvs_client_driver *drv, const struct nvs_client_datapath *dp);

/* Tell the scheduler that this process cannot map `mib` more MiB of HBM (other
 * clients still hold it).  Safe from any thread; rate-limited by the caller. */
void nvs_client_pressure(uint64_t mib);

/* Returns only when this process holds the GPU lock (or the scheduler is off). */
void continue_with_lock(void);
/* Must follow every continue_with_lock() once the gated driver call has returned. */
void nvs_gate_leave(void);

/* Called by the launch hooks: resets the adaptive sync window (reference
 * src/client.c:62 touches pending_kernel_window directly). */
extern void (*nvs_client_on_context_sync)(void);

#pragma GCC visibility pop

#endif /* NVS_CLIENT_H */
