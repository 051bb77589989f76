/*
 * driver_app.c -- a tiny "unmodified CUDA application" written against the
 * driver API, used by the CPU test-suite (with oracle/fake_cuda.c standing in
 * for libcuda) and runnable on a real GPU as well.  It knows nothing about
 * nvshare: the library under test is injected with LD_PRELOAD.
 *
 * usage: driver_app <MiB per buffer> <seconds> <seed> [n_buffers] [idle seconds before verifying]
 * Prints "RESULT PASS|FAIL iters=<n> mismatches=<m>".
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

typedef int CUresult;
typedef unsigned long long CUdeviceptr;
typedef void *CUcontext;
extern CUresult cuInit(unsigned);
extern CUresult cuDevicePrimaryCtxRetain(CUcontext *, int);
extern CUresult cuCtxSetCurrent(CUcontext);
extern CUresult cuCtxSynchronize(void);
extern CUresult cuMemAlloc_v2(CUdeviceptr *, size_t);
extern CUresult cuMemFree_v2(CUdeviceptr);
extern CUresult cuMemGetInfo_v2(size_t *, size_t *);
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
	int nbuf = argc > 4 ? atoi(argv[4]) : 2;
	double idle = argc > 5 ? atof(argv[5]) : 0.0;
	size_t bytes = mib << 20, words = bytes / 8;
	if (nbuf < 2 || nbuf > 64)
		return 2;

	CUcontext ctx;
	CK(cuInit(0));
	CK(cuDevicePrimaryCtxRetain(&ctx, 0));
	CK(cuCtxSetCurrent(ctx));
	size_t fr, tot;
	CK(cuMemGetInfo_v2(&fr, &tot));
	printf("meminfo free=%zu total=%zu\n", fr, tot);

	CUdeviceptr buf[64];
	for (int i = 0; i < nbuf; ++i)
		CK(cuMemAlloc_v2(&buf[i], bytes));
	uint64_t *h = malloc(bytes), *back = malloc(bytes);
	for (size_t i = 0; i < words; ++i)
		h[i] = mix(i, seed);
	CK(cuMemcpyHtoD_v2(buf[0], h, bytes));

	const int no_launch = getenv("DRIVER_APP_NO_LAUNCH") != NULL; /* copies only: runs on the real driver too */
	struct timespec t0, t;
	clock_gettime(CLOCK_MONOTONIC, &t0);
	unsigned long iters = 0;
	for (;;) {
		/* rotate the payload through every buffer, "compute" in between */
		int src = (int)(iters % (unsigned)nbuf), dst = (int)((iters + 1) % (unsigned)nbuf);
		CK(cuMemcpyDtoD_v2(buf[dst], buf[src], bytes));
		if (!no_launch) /* (the handle only means something to the fake driver) */
			CK(cuLaunchKernel((void *)0x1234, 1, 1, 1, 32, 1, 1, 0, NULL, NULL, NULL));
		iters++;
		clock_gettime(CLOCK_MONOTONIC, &t);
		if ((t.tv_sec - t0.tv_sec) + (t.tv_nsec - t0.tv_nsec) * 1e-9 >= seconds)
			break;
		usleep(2000);
	}
	CK(cuCtxSynchronize());
	if (idle > 0) { /* an interactive application between two bursts: holds memory, submits nothing */
		printf("idle start\n");
		fflush(stdout);
		usleep((useconds_t)(idle * 1e6));
	}
	CK(cuMemcpyDtoH_v2(back, buf[iters % (unsigned)nbuf], bytes));
	unsigned long bad = 0;
	for (size_t i = 0; i < words; ++i)
		bad += back[i] != h[i];
	for (int i = 0; i < nbuf; ++i)
		CK(cuMemFree_v2(buf[i]));
	printf("RESULT %s iters=%lu mismatches=%lu\n", bad ? "FAIL" : "PASS", iters, bad);
	return bad ? 1 : 0;
}
