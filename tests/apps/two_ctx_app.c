/*
 * two_ctx_app.c -- a process that uses a second CUDA context (a second GPU): run under
 * LD_PRELOAD=<libnvshare.so> with oracle/fake_cuda.c as libcuda and at least two "GPUs".
 * Allocates in the first context, creates a context on device 1, allocates there, moves a pattern
 * through both allocations, frees the second context's memory from the first context.
 * Prints "RESULT PASS" when every byte came back.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef int CUresult;
typedef unsigned long long CUdeviceptr;
typedef void *CUcontext;
extern CUresult cuInit(unsigned);
extern CUresult cuDevicePrimaryCtxRetain(CUcontext *, int);
extern CUresult cuCtxCreate_v2(CUcontext *, unsigned, int);
extern CUresult cuCtxSetCurrent(CUcontext);
extern CUresult cuMemAlloc_v2(CUdeviceptr *, size_t);
extern CUresult cuMemFree_v2(CUdeviceptr);
extern CUresult cuMemcpyHtoD_v2(CUdeviceptr, const void *, size_t);
extern CUresult cuMemcpyDtoH_v2(void *, CUdeviceptr, size_t);

#define CK(x) do { CUresult r_ = (x); if (r_ != 0) { printf("%s -> %d\n", #x, r_); return 1; } } while (0)

int main(void)
{
	const size_t n = 4u << 20;
	CUcontext first, second;
	CUdeviceptr a = 0, b = 0;
	unsigned char *src = malloc(n), *dst = malloc(n);
	for (size_t i = 0; i < n; ++i)
		src[i] = (unsigned char)(i * 7 + (i >> 11));
	CK(cuInit(0));
	CK(cuDevicePrimaryCtxRetain(&first, 0));
	CK(cuCtxSetCurrent(first));
	CK(cuMemAlloc_v2(&a, n));                 /* the swap engine comes up in this context */
	CK(cuCtxCreate_v2(&second, 0, 1));        /* current from here on */
	CK(cuMemAlloc_v2(&b, n));                 /* must not come out of the first context's engine */
	CK(cuMemcpyHtoD_v2(b, src, n));
	memset(dst, 0, n);
	CK(cuMemcpyDtoH_v2(dst, b, n));
	int bad = memcmp(src, dst, n) != 0;
	CK(cuCtxSetCurrent(first));
	CK(cuMemcpyHtoD_v2(a, src, n));
	memset(dst, 0, n);
	CK(cuMemcpyDtoH_v2(dst, a, n));
	bad |= memcmp(src, dst, n) != 0;
	CK(cuMemFree_v2(b));                      /* whoever frees it, it goes back to where it came from */
	CK(cuMemFree_v2(a));
	CUdeviceptr c = 0;
	CK(cuMemAlloc_v2(&c, n));                 /* and the first context still gets engine memory */
	CK(cuMemFree_v2(c));
	printf(bad ? "RESULT FAIL\n" : "RESULT PASS\n");
	return bad;
}
