/*
 * coop_app.cu -- an "unmodified CUDA application" (runtime API) whose work is a COOPERATIVE launch
 * (grid-wide synchronisation) on memory from cudaMalloc.  The reference's hook does not gate
 * cuLaunchCooperativeKernel (src/hook.c:545-577); with explicitly mapped memory an ungated launch
 * that touches an evicted slab is a fatal fault, so ours does.  Knows nothing
 * about nvshare: the library under test is injected with LD_PRELOAD.
 *
 * usage: coop_app <MiB> <seconds> <seed>     prints "RESULT PASS|FAIL iters=<n>"
 * Each iteration: phase 1 every block adds `it` to its share of the buffer and writes a partial
 * sum; grid.sync(); phase 2 block 0 folds the partials into out[it & 1023].  The host checks the
 * final buffer and the last folded sum exactly (64-bit integers).
 */
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>
namespace cg = cooperative_groups;

__global__ void step(unsigned long long *buf, size_t n, unsigned long long it, unsigned long long *partial,
		     unsigned long long *out)
{
	cg::grid_group grid = cg::this_grid();
	unsigned long long acc = 0;
	for (size_t i = grid.thread_rank(); i < n; i += grid.size()) {
		buf[i] += it;
		acc += buf[i];
	}
	__shared__ unsigned long long sh[256];
	sh[threadIdx.x] = acc;
	__syncthreads();
	for (int o = 128; o > 0; o >>= 1) {
		if ((int)threadIdx.x < o)
			sh[threadIdx.x] += sh[threadIdx.x + o];
		__syncthreads();
	}
	if (threadIdx.x == 0)
		partial[blockIdx.x] = sh[0];
	grid.sync();
	if (blockIdx.x == 0 && threadIdx.x == 0) {
		unsigned long long s = 0;
		for (unsigned b = 0; b < gridDim.x; ++b)
			s += partial[b];
		out[it & 1023] = s;
	}
}

#define CK(x) do { cudaError_t r_ = (x); if (r_ != cudaSuccess) { printf("RESULT FAIL %s -> %s\n", #x, cudaGetErrorName(r_)); return 2; } } while (0)

int main(int argc, char **argv)
{
	size_t mib = argc > 1 ? strtoull(argv[1], NULL, 0) : 64;
	double seconds = argc > 2 ? atof(argv[2]) : 3.0;
	unsigned long long seed = argc > 3 ? strtoull(argv[3], NULL, 0) : 1;
	size_t n = (mib << 20) / 8;
	int dev = 0, sms = 0, per_sm = 0, coop = 0;
	CK(cudaSetDevice(dev));
	CK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
	CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
	CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, step, 256, 0));
	if (!coop || per_sm < 1) {
		printf("RESULT FAIL no cooperative launch\n");
		return 2;
	}
	unsigned grid = (unsigned)sms; /* one block per SM: co-resident even while another client's kernels run */
	unsigned long long *buf, *partial, *out;
	CK(cudaMalloc(&buf, n * 8));
	CK(cudaMalloc(&partial, 4096 * 8));
	CK(cudaMalloc(&out, 1024 * 8));
	unsigned long long *h = (unsigned long long *)malloc(n * 8);
	for (size_t i = 0; i < n; ++i)
		h[i] = seed * 1000003ull + i;
	CK(cudaMemcpy(buf, h, n * 8, cudaMemcpyHostToDevice));
	struct timespec t0, t;
	clock_gettime(CLOCK_MONOTONIC, &t0);
	unsigned long long it = 0, sum_it = 0;
	for (;;) {
		++it;
		sum_it += it;
		void *args[] = {&buf, &n, &it, &partial, &out};
		CK(cudaLaunchCooperativeKernel((void *)step, dim3(grid), dim3(256), args, 0, 0));
		if ((it & 15) == 0)
			CK(cudaDeviceSynchronize());
		clock_gettime(CLOCK_MONOTONIC, &t);
		if ((t.tv_sec - t0.tv_sec) + (t.tv_nsec - t0.tv_nsec) * 1e-9 >= seconds)
			break;
	}
	CK(cudaDeviceSynchronize());
	unsigned long long *back = (unsigned long long *)malloc(n * 8), last = 0, want = 0;
	CK(cudaMemcpy(back, buf, n * 8, cudaMemcpyDeviceToHost));
	CK(cudaMemcpy(&last, out + (it & 1023), 8, cudaMemcpyDeviceToHost));
	unsigned long long bad = 0;
	for (size_t i = 0; i < n; ++i) {
		bad += back[i] != h[i] + sum_it;
		want += h[i] + sum_it;
	}
	bad += last != want;
	printf("RESULT %s iters=%llu mismatches=%llu\n", bad ? "FAIL" : "PASS", it, bad);
	return bad ? 1 : 0;
}
