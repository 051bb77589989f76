/*
 * load_app.c -- an application in its loading phase: it allocates device
 * buffers, uploads data, reads some of it back, and never launches a kernel.
 * Used to show that such a client does not need the GPU lock at all
 * (host<->device copies to memory that is not on the GPU
 * are served from the pinned-host backing copy).
 *
 * usage: load_app <MiB per buffer> <n_buffers> <seed>
 * Prints "RESULT PASS|FAIL seconds=<t> mismatches=<m>".
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef int CUresult;
typedef unsigned long long CUdeviceptr;
typedef void *CUcontext;
extern CUresult cuInit(unsigned);
extern CUresult cuDevicePrimaryCtxRetain(CUcontext *, int);
extern CUresult cuCtxSetCurrent(CUcontext);
extern CUresult cuMemAlloc_v2(CUdeviceptr *, size_t);
extern CUresult cuMemFree_v2(CUdeviceptr);
extern CUresult cuMemcpyHtoD_v2(CUdeviceptr, const void *, size_t);
extern CUresult cuMemcpyHtoDAsync_v2(CUdeviceptr, const void *, size_t, void *);
extern CUresult cuMemcpyDtoH_v2(void *, CUdeviceptr, size_t);
extern CUresult cuMemcpyDtoHAsync_v2(void *, CUdeviceptr, size_t, void *);

static uint64_t mix(uint64_t i, uint64_t seed)
{
	uint64_t z = i * 0x9E3779B97F4A7C15ull + seed;
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	return z ^ (z >> 27);
}

#define CK(x) do { CUresult r_ = (x); if (r_ != 0) { printf("RESULT FAIL %s -> %d\n", #x, r_); return 2; } } while (0)

int main(int argc, char **argv)
{
	size_t mib = argc > 1 ? strtoull(argv[1], NULL, 0) : 16;
	int nbuf = argc > 2 ? atoi(argv[2]) : 2;
	uint64_t seed = argc > 3 ? strtoull(argv[3], NULL, 0) : 1;
	size_t bytes = mib << 20, words = bytes / 8;
	if (nbuf < 1 || nbuf > 64)
		return 2;
	struct timespec t0, t1;
	clock_gettime(CLOCK_MONOTONIC, &t0);
	CUcontext ctx;
	CK(cuInit(0));
	CK(cuDevicePrimaryCtxRetain(&ctx, 0));
	CK(cuCtxSetCurrent(ctx));
	CUdeviceptr buf[64];
	uint64_t *h = malloc(bytes), *back = malloc(bytes);
	unsigned long bad = 0;
	for (int b = 0; b < nbuf; ++b) {
		CK(cuMemAlloc_v2(&buf[b], bytes));
		for (size_t i = 0; i < words; ++i)
			h[i] = mix(i, seed + (uint64_t)b);
		/* upload in two pieces at an odd boundary, the second one "asynchronously" */
		size_t cut = bytes / 3 + 24;
		CK(cuMemcpyHtoD_v2(buf[b], h, cut));
		CK(cuMemcpyHtoDAsync_v2(buf[b] + cut, (const char *)h + cut, bytes - cut, NULL));
	}
	for (int b = nbuf - 1; b >= 0; --b) {
		memset(back, 0xEE, bytes);
		CK(cuMemcpyDtoH_v2(back, buf[b], bytes / 2));
		CK(cuMemcpyDtoHAsync_v2((char *)back + bytes / 2, buf[b] + bytes / 2, bytes - bytes / 2, NULL));
		for (size_t i = 0; i < words; ++i)
			bad += back[i] != mix(i, seed + (uint64_t)b);
	}
	for (int b = 0; b < nbuf; ++b)
		CK(cuMemFree_v2(buf[b]));
	clock_gettime(CLOCK_MONOTONIC, &t1);
	printf("RESULT %s seconds=%.3f mismatches=%lu\n", bad ? "FAIL" : "PASS",
	       (t1.tv_sec - t0.tv_sec) + (t1.tv_nsec - t0.tv_nsec) * 1e-9, bad);
	return bad ? 1 : 0;
}
