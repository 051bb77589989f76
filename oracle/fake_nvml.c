/*
 * fake_nvml.c -- TEST INFRASTRUCTURE: a stand-in for libnvidia-ml.so.1 next to the fake driver
 * (fake_cuda.c), for the one thing the client runtime asks NVML: is the GPU busy
 * (reference src/client.c:386-445: nvmlDeviceGetHandleByIndex(0) + nvmlDeviceGetUtilizationRates).
 * NVML numbers the PHYSICAL GPUs and ignores CUDA_VISIBLE_DEVICES, which is the whole point of the
 * test that uses this: a client that computes on physical GPU 1 must ask about GPU 1.
 *   FAKE_NVML_UTIL="100,0"   utilisation of physical GPU 0, 1, ... (default 0)
 *   FAKE_NVML_TRACE=<file>   one line per utilisation query: "util <physical gpu>"
 * Only on the library path of tests that ask for it (oracle/_ref/fakenvml/).
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct { unsigned gpu, memory; } util_t;

static void trace(const char *what, int gpu)
{
	const char *p = getenv("FAKE_NVML_TRACE");
	if (!p || !*p)
		return;
	FILE *f = fopen(p, "a");
	if (f) {
		fprintf(f, "%s %d\n", what, gpu);
		fclose(f);
	}
}

int nvmlInit_v2(void) { return 0; }

int nvmlDeviceGetHandleByIndex_v2(unsigned idx, void **dev)
{
	*dev = (void *)(uintptr_t)(idx + 1);
	trace("by_index", (int)idx);
	return 0;
}

/* "GPU-xxxxxxxx-xxxx-xxxx-xxxx-xxxxxxxxxxxx": the fake driver puts the physical GPU in the last byte */
int nvmlDeviceGetHandleByUUID(const char *uuid, void **dev)
{
	size_t n = uuid ? strlen(uuid) : 0;
	if (n != 40 || strncmp(uuid, "GPU-", 4) != 0)
		return 2; /* NVML_ERROR_INVALID_ARGUMENT */
	if (getenv("FAKE_NVML_NO_UUID"))
		return 6; /* NVML_ERROR_NOT_FOUND */
	unsigned gpu = (unsigned)strtoul(uuid + n - 2, NULL, 16);
	*dev = (void *)(uintptr_t)(gpu + 1);
	trace("by_uuid", (int)gpu);
	return 0;
}

int nvmlDeviceGetUtilizationRates(void *dev, util_t *u)
{
	int gpu = (int)(uintptr_t)dev - 1, k = 0;
	unsigned val = 0;
	char buf[128], *save = NULL;
	snprintf(buf, sizeof(buf), "%s", getenv("FAKE_NVML_UTIL") ? getenv("FAKE_NVML_UTIL") : "");
	for (char *t = strtok_r(buf, ",", &save); t; t = strtok_r(NULL, ",", &save), ++k)
		if (k == gpu)
			val = (unsigned)atoi(t);
	u->gpu = val;
	u->memory = 0;
	trace("util", gpu);
	return 0;
}
