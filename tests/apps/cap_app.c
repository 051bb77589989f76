/*
 * cap_app.c -- what a hooked application is told about its GPU, on request: run under
 * LD_PRELOAD=<libnvshare.so> with oracle/fake_cuda.c as libcuda (FAKE_CUDA_DEVICE selects the GPU the
 * context is on).  Commands on stdin, one answer line each:
 *   info          -> "free_mib=<n> total_mib=<n>"       (hooked cuMemGetInfo)
 *   alloc <MiB>   -> "alloc rc=<CUresult>"               (hooked cuMemAlloc; freed again at once)
 *   hold <MiB>    -> "hold rc=<CUresult>"                (hooked cuMemAlloc; kept)
 *   quit
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef int CUresult;
typedef unsigned long long CUdeviceptr;
typedef void *CUcontext;
extern CUresult cuInit(unsigned);
extern CUresult cuCtxGetDevice(int *);
extern CUresult cuDevicePrimaryCtxRetain(CUcontext *, int);
extern CUresult cuCtxSetCurrent(CUcontext);
extern CUresult cuMemAlloc_v2(CUdeviceptr *, size_t);
extern CUresult cuMemFree_v2(CUdeviceptr);
extern CUresult cuMemGetInfo_v2(size_t *, size_t *);

int main(void)
{
	CUcontext ctx;
	CUdeviceptr first = 0;
	int dev = 0;
	if (cuInit(0) != 0 || cuCtxGetDevice(&dev) != 0 || cuDevicePrimaryCtxRetain(&ctx, dev) != 0 || cuCtxSetCurrent(ctx) != 0)
		return 2;
	/* the swap engine (and with it this process's view of the ledger) comes with the first allocation */
	if (cuMemAlloc_v2(&first, 2u << 20) != 0)
		return 3;
	printf("ready device=%d\n", dev);
	fflush(stdout);
	char line[128];
	while (fgets(line, sizeof(line), stdin)) {
		size_t mib = 0;
		if (!strncmp(line, "info", 4)) {
			size_t fr = 0, tot = 0;
			CUresult rc = cuMemGetInfo_v2(&fr, &tot);
			printf("free_mib=%zu total_mib=%zu rc=%d\n", fr >> 20, tot >> 20, rc);
		} else if (sscanf(line, "alloc %zu", &mib) == 1) {
			CUdeviceptr p = 0;
			CUresult rc = cuMemAlloc_v2(&p, mib << 20);
			printf("alloc rc=%d\n", rc);
			if (rc == 0)
				cuMemFree_v2(p);
		} else if (sscanf(line, "hold %zu", &mib) == 1) {
			CUdeviceptr p = 0;
			printf("hold rc=%d\n", cuMemAlloc_v2(&p, mib << 20));
		} else if (!strncmp(line, "quit", 4)) {
			break;
		}
		fflush(stdout);
	}
	return 0;
}
