/*
 * mt_app.c -- a multi-threaded driver-API application: T threads, each with its
 * own buffers, copy and launch concurrently.  Stresses the part of the library
 * the reference gets away without: closing the gate must also wait for calls
 * that already passed it (close_gate_and_drain in client.c), otherwise a thread
 * touches slabs that are being unmapped.
 *
 * usage: mt_app <MiB per buffer> <seconds> <seed> <threads> [io]
 */
#include <pthread.h>
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
extern CUresult cuMemAlloc_v2(CUdeviceptr *, size_t);
extern CUresult cuMemFree_v2(CUdeviceptr);
extern CUresult cuMemcpyHtoD_v2(CUdeviceptr, const void *, size_t);
extern CUresult cuMemcpyDtoH_v2(void *, CUdeviceptr, size_t);
extern CUresult cuMemcpyDtoD_v2(CUdeviceptr, CUdeviceptr, size_t);
extern CUresult cuLaunchKernel(void *, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned,
			       void *, void **, void **);

static CUcontext ctx;
static size_t bytes;
static double seconds;
static uint64_t seed;
static int io_mode; /* also round-trip a third buffer through DtoH / HtoD every iteration */
static unsigned long total_bad, total_iters;
static pthread_mutex_t sum_mu = PTHREAD_MUTEX_INITIALIZER;

static uint64_t mix(uint64_t i, uint64_t s)
{
	uint64_t z = i * 0x9E3779B97F4A7C15ull + s;
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	return z ^ (z >> 27);
}

static void *worker(void *arg)
{
	long id = (long)arg;
	size_t words = bytes / 8;
	CUdeviceptr a, b;
	cuCtxSetCurrent(ctx);
	if (cuMemAlloc_v2(&a, bytes) || cuMemAlloc_v2(&b, bytes)) {
		printf("alloc failed in thread %ld\n", id);
		exit(2);
	}
	uint64_t *h = malloc(bytes), *back = malloc(bytes);
	for (size_t i = 0; i < words; ++i)
		h[i] = mix(i, seed * 1000 + (uint64_t)id);
	if (cuMemcpyHtoD_v2(a, h, bytes))
		exit(2);
	/* io mode: a small third buffer whose contents change every iteration; whether a copy
	 * is served by the "GPU" or from the backing copy depends on where the lock is right now */
	const size_t io_bytes = 3u << 20, io_words = io_bytes / 8, io_off = 4096 + 8 * (size_t)id;
	CUdeviceptr c = 0;
	uint64_t *io = NULL, *io_back = NULL;
	unsigned long io_bad = 0;
	if (io_mode) {
		if (cuMemAlloc_v2(&c, io_bytes + 2 * io_off))
			exit(2);
		io = malloc(io_bytes);
		io_back = malloc(io_bytes);
		for (size_t i = 0; i < io_words; ++i)
			io[i] = mix(i, 77 + (uint64_t)id);
		if (cuMemcpyHtoD_v2(c + io_off, io, io_bytes))
			exit(2);
	}
	struct timespec t0, t;
	clock_gettime(CLOCK_MONOTONIC, &t0);
	unsigned long iters = 0;
	for (;;) {
		if (io_mode) {
			if (cuMemcpyDtoH_v2(io_back, c + io_off, io_bytes))
				exit(2);
			for (size_t i = 0; i < io_words; ++i)
				io_bad += io_back[i] != io[i];
			for (size_t i = 0; i < io_words; i += 61)
				io[i] = mix(i, iters * 131 + (uint64_t)id);
			if (cuMemcpyHtoD_v2(c + io_off, io, io_bytes))
				exit(2);
		}
		if (cuMemcpyDtoD_v2(iters & 1 ? a : b, iters & 1 ? b : a, bytes))
			exit(2);
		if (cuLaunchKernel((void *)0x1234, 1, 1, 1, 32, 1, 1, 0, NULL, NULL, NULL))
			exit(2);
		iters++;
		clock_gettime(CLOCK_MONOTONIC, &t);
		if ((t.tv_sec - t0.tv_sec) + (t.tv_nsec - t0.tv_nsec) * 1e-9 >= seconds)
			break;
		usleep(500);
	}
	if (cuMemcpyDtoH_v2(back, iters & 1 ? b : a, bytes))
		exit(2);
	unsigned long bad = 0;
	for (size_t i = 0; i < words; ++i)
		bad += back[i] != h[i];
	cuMemFree_v2(a);
	cuMemFree_v2(b);
	if (io_mode)
		cuMemFree_v2(c);
	pthread_mutex_lock(&sum_mu);
	total_bad += bad + io_bad;
	total_iters += iters;
	pthread_mutex_unlock(&sum_mu);
	return NULL;
}

int main(int argc, char **argv)
{
	size_t mib = argc > 1 ? strtoull(argv[1], NULL, 0) : 8;
	seconds = argc > 2 ? atof(argv[2]) : 2.0;
	seed = argc > 3 ? strtoull(argv[3], NULL, 0) : 1;
	long threads = argc > 4 ? atol(argv[4]) : 4;
	io_mode = argc > 5 && atoi(argv[5]) != 0;
	bytes = mib << 20;
	if (cuInit(0) || cuDevicePrimaryCtxRetain(&ctx, 0) || cuCtxSetCurrent(ctx))
		return 2;
	pthread_t th[32];
	for (long i = 0; i < threads; ++i)
		pthread_create(&th[i], NULL, worker, (void *)i);
	for (long i = 0; i < threads; ++i)
		pthread_join(th[i], NULL);
	printf("RESULT %s iters=%lu mismatches=%lu\n", total_bad ? "FAIL" : "PASS", total_iters, total_bad);
	return total_bad ? 1 : 0;
}
