/*
 * trace_app.c -- drives the hooked driver entry points in a fixed order and
 * prints what the application sees; run under LD_PRELOAD=<libnvshare.so> with
 * oracle/fake_cuda.c as libcuda.  Used to record the reference's behaviour
 * (tests/golden/make_golden.py) and to compare ours against it.
 *
 * usage: trace_app <alloc MiB> [launches]
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

typedef int CUresult;
typedef unsigned long long CUdeviceptr;
typedef void *CUcontext;
extern CUresult cuInit(unsigned);
extern CUresult cuDevicePrimaryCtxRetain(CUcontext *, int);
extern CUresult cuCtxSetCurrent(CUcontext);
extern CUresult cuMemAlloc_v2(CUdeviceptr *, size_t);
extern CUresult cuMemFree_v2(CUdeviceptr);
extern CUresult cuMemGetInfo_v2(size_t *, size_t *);
extern CUresult cuMemcpyHtoD_v2(CUdeviceptr, const void *, size_t);
extern CUresult cuMemcpyDtoH_v2(void *, CUdeviceptr, size_t);
extern CUresult cuLaunchKernel(void *, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned,
			       void *, void **, void **);

int main(int argc, char **argv)
{
	size_t mib = argc > 1 ? strtoull(argv[1], NULL, 0) : 1024;
	int launches = argc > 2 ? atoi(argv[2]) : 3;
	CUcontext ctx;
	CUdeviceptr a = 0, b = 0, small = 0;
	size_t fr = 0, tot = 0;
	char host[64] = "nvshare", back[64] = {0};

	/* before initialisation the hooks must refuse (src/hook.c:654) */
	printf("prealloc rc=%d\n", cuMemAlloc_v2(&a, 4096));
	printf("cuInit rc=%d\n", cuInit(0));
	printf("ctx rc=%d\n", cuDevicePrimaryCtxRetain(&ctx, 0) | cuCtxSetCurrent(ctx));
	int rc = cuMemGetInfo_v2(&fr, &tot);
	printf("meminfo rc=%d reserve_mib=%zu\n", rc, (tot - fr) >> 20);
	printf("alloc1 rc=%d\n", cuMemAlloc_v2(&a, mib << 20));
	printf("alloc2 rc=%d\n", cuMemAlloc_v2(&b, mib << 20));
	printf("alloc_small rc=%d\n", cuMemAlloc_v2(&small, 4096));
	for (int i = 0; i < launches; ++i)
		printf("launch%d rc=%d\n", i + 1, cuLaunchKernel((void *)0x1234, 1, 1, 1, 32, 1, 1, 0, NULL, NULL, NULL));
	printf("h2d rc=%d\n", cuMemcpyHtoD_v2(small, host, sizeof host));
	rc = cuMemcpyDtoH_v2(back, small, sizeof back);
	printf("d2h rc=%d same=%d\n", rc, __builtin_memcmp(host, back, 64) == 0);
	printf("free1 rc=%d\n", cuMemFree_v2(a));
	printf("free_small rc=%d\n", cuMemFree_v2(small));
	printf("free_bogus rc=%d\n", cuMemFree_v2(0x1000));
	return 0;
}
