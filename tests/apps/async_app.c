/*
 * async_app.c -- an "unmodified CUDA application" that obtains its device memory the other
 * ways: stream-ordered (cuMemAllocAsync / cuMemAllocFromPoolAsync / cuMemFreeAsync) and pitched
 * (cuMemAllocPitch).  CPU test-suite (fake driver) and real GPU alike; knows nothing about nvshare.
 *
 * usage: async_app <MiB per buffer> <seconds> <seed> [cap-probe GiB]
 * Prints "CAP rc=<r>" for an allocation of <cap-probe GiB> (0 = skip), "WRAP rc=<r>" for one of 2^64 - 1 MiB, and
 * "RESULT PASS|FAIL iters=<n> mismatches=<m>".
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

typedef int CUresult;
typedef unsigned long long CUdeviceptr;
typedef void *CUcontext, *CUstream;
extern CUresult cuInit(unsigned);
extern CUresult cuDevicePrimaryCtxRetain(CUcontext *, int);
extern CUresult cuCtxSetCurrent(CUcontext);
extern CUresult cuCtxSynchronize(void);
extern CUresult cuStreamCreate(CUstream *, unsigned);
extern CUresult cuStreamSynchronize(CUstream);
extern CUresult cuMemAllocAsync(CUdeviceptr *, size_t, CUstream);
extern CUresult cuMemAllocFromPoolAsync(CUdeviceptr *, size_t, void *, CUstream);
extern CUresult cuMemFreeAsync(CUdeviceptr, CUstream);
extern CUresult cuMemAllocPitch_v2(CUdeviceptr *, size_t *, size_t, size_t, unsigned);
extern CUresult cuMemFree_v2(CUdeviceptr);
extern CUresult cuMemcpyHtoD_v2(CUdeviceptr, const void *, size_t);
extern CUresult cuMemcpyDtoH_v2(void *, CUdeviceptr, size_t);
extern CUresult cuMemcpyDtoD_v2(CUdeviceptr, CUdeviceptr, size_t);
extern CUresult cuLaunchKernel(void *, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned,
			       void *, void **, void **);

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
	double seconds = argc > 2 ? atof(argv[2]) : 2.0;
	uint64_t seed = argc > 3 ? strtoull(argv[3], NULL, 0) : 1;
	size_t cap_gib = argc > 4 ? strtoull(argv[4], NULL, 0) : 0;
	size_t bytes = mib << 20, words = bytes / 8;

	CUcontext ctx;
	CUstream st;
	CK(cuInit(0));
	CK(cuDevicePrimaryCtxRetain(&ctx, 0));
	CK(cuCtxSetCurrent(ctx));
	CK(cuStreamCreate(&st, 0));
	if (cap_gib) { /* the per-process cap applies to stream-ordered allocations too */
		CUdeviceptr big = 0;
		CUresult r = cuMemAllocAsync(&big, cap_gib << 30, st);
		printf("CAP rc=%d\n", r);
		if (r == 0)
			cuMemFreeAsync(big, st);
		/* a request so large that "allocated so far + request" wraps around must not slip under the cap */
		r = cuMemAllocAsync(&big, (size_t)0 - ((size_t)1 << 20), st);
		printf("WRAP rc=%d\n", r);
		if (r == 0)
			cuMemFreeAsync(big, st);
	}
	CUdeviceptr buf[3];
	size_t pitch = 0;
	CK(cuMemAllocAsync(&buf[0], bytes, st));
	CK(cuMemAllocFromPoolAsync(&buf[1], bytes, NULL, st));
	CK(cuMemAllocPitch_v2(&buf[2], &pitch, 4096 - 16, bytes / 4096, 16)); /* rows of 4080 bytes padded to 4096 */
	if (pitch != 4096) {
		printf("RESULT FAIL pitch=%zu\n", pitch);
		return 2;
	}
	uint64_t *h = malloc(bytes), *back = malloc(bytes);
	for (size_t i = 0; i < words; ++i)
		h[i] = mix(i, seed);
	CK(cuMemcpyHtoD_v2(buf[0], h, bytes));
	struct timespec t0, t;
	clock_gettime(CLOCK_MONOTONIC, &t0);
	unsigned long iters = 0;
	for (;;) {
		int src = (int)(iters % 3), dst = (int)((iters + 1) % 3);
		CK(cuMemcpyDtoD_v2(buf[dst], buf[src], bytes));
		CK(cuLaunchKernel((void *)0x1234, 1, 1, 1, 32, 1, 1, 0, NULL, NULL, NULL));
		iters++;
		clock_gettime(CLOCK_MONOTONIC, &t);
		if ((t.tv_sec - t0.tv_sec) + (t.tv_nsec - t0.tv_nsec) * 1e-9 >= seconds)
			break;
		usleep(2000);
	}
	CK(cuCtxSynchronize());
	CK(cuMemcpyDtoH_v2(back, buf[iters % 3], bytes));
	unsigned long bad = 0;
	for (size_t i = 0; i < words; ++i)
		bad += back[i] != h[i];
	CK(cuMemFreeAsync(buf[0], st));
	CK(cuMemFreeAsync(buf[1], st));
	CK(cuMemFree_v2(buf[2]));
	CK(cuStreamSynchronize(st));
	printf("RESULT %s iters=%lu mismatches=%lu\n", bad ? "FAIL" : "PASS", iters, bad);
	return bad ? 1 : 0;
}
