/*
 * capture_app.c -- an "unmodified CUDA application" that records a CUDA graph by stream capture:
 * more launches into the capturing stream than any sync window, and a stream-ordered free inside
 * the capture.  A synchronisation while a capture is active is an illegal call that invalidates the
 * capture, so an interposer that synchronises every N launches breaks this application.
 * CPU test-suite (fake driver); knows nothing about nvshare.
 *
 * usage: capture_app <launches>
 * Prints "FREE_IN_CAPTURE rc=<r>", "END_CAPTURE rc=<r>", "FREE_AFTER rc=<r>" and "RESULT PASS|FAIL".
 */
#include <stdio.h>
#include <stdlib.h>

typedef int CUresult;
typedef unsigned long long CUdeviceptr;
typedef void *CUcontext, *CUstream;
extern CUresult cuInit(unsigned);
extern CUresult cuDevicePrimaryCtxRetain(CUcontext *, int);
extern CUresult cuCtxSetCurrent(CUcontext);
extern CUresult cuCtxSynchronize(void);
extern CUresult cuStreamCreate(CUstream *, unsigned);
extern CUresult cuStreamBeginCapture_v2(CUstream, int);
extern CUresult cuStreamEndCapture(CUstream, void **);
extern CUresult cuMemAllocAsync(CUdeviceptr *, size_t, CUstream);
extern CUresult cuMemFreeAsync(CUdeviceptr, CUstream);
extern CUresult cuLaunchKernel(void *, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned,
			       void *, void **, void **);

#define CK(x) do { CUresult r_ = (x); if (r_ != 0) { printf("RESULT FAIL %s -> %d\n", #x, r_); return 2; } } while (0)

int main(int argc, char **argv)
{
	long launches = argc > 1 ? atol(argv[1]) : 5000;
	CUcontext ctx;
	CUstream st;
	CUdeviceptr buf = 0;
	void *graph = NULL;
	CK(cuInit(0));
	CK(cuDevicePrimaryCtxRetain(&ctx, 0));
	CK(cuCtxSetCurrent(ctx));
	CK(cuStreamCreate(&st, 0));
	CK(cuMemAllocAsync(&buf, 8u << 20, st));
	CK(cuLaunchKernel((void *)0x1234, 1, 1, 1, 32, 1, 1, 0, st, NULL, NULL)); /* takes the GPU lock before the capture */
	CK(cuCtxSynchronize());
	CK(cuStreamBeginCapture_v2(st, 0));
	for (long i = 0; i < launches; ++i)
		CK(cuLaunchKernel((void *)0x1234, 1, 1, 1, 32, 1, 1, 0, st, NULL, NULL));
	CUresult rf = cuMemFreeAsync(buf, st);
	printf("FREE_IN_CAPTURE rc=%d\n", rf);
	CUresult re = cuStreamEndCapture(st, &graph);
	printf("END_CAPTURE rc=%d\n", re);
	CUresult ra = rf == 0 ? 0 : cuMemFreeAsync(buf, st);
	printf("FREE_AFTER rc=%d\n", ra);
	printf("RESULT %s\n", re == 0 && graph && ra == 0 ? "PASS" : "FAIL");
	return re == 0 && ra == 0 ? 0 : 1;
}
