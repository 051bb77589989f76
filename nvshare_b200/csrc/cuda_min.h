/*
 * cuda_min.h -- the subset of the CUDA driver ABI (cuda.h, CUDA 12.x) that
 * libnvshare.so touches, declared by hand so that the library builds with gcc
 * alone and NEEDs only libc at run time -- the same property the reference
 * gets from its src/cuda_defs.h:16-148.  Layouts were checked against
 * /usr/local/cuda/include/cuda.h (12.9): CUmemAllocationProp is 32 bytes,
 * CUmemAccessDesc 12 bytes, CUmemLocation 8 bytes.
 *
 * Compared with the reference's header this one also covers the virtual-memory
 * management API (cuMemCreate/cuMemMap/...), streams, events and module
 * loading, because the data path is ours instead of the UVM driver's.
 */
#ifndef NVS_CUDA_MIN_H
#define NVS_CUDA_MIN_H

#include <stddef.h>
#include <stdint.h>

typedef int CUresult; /* enum in cuda.h; int-sized */
enum {
	CUDA_SUCCESS = 0,
	CUDA_ERROR_INVALID_VALUE = 1,
	CUDA_ERROR_OUT_OF_MEMORY = 2,
	CUDA_ERROR_NOT_INITIALIZED = 3,
	CUDA_ERROR_INVALID_CONTEXT = 201,
	CUDA_ERROR_NOT_FOUND = 500,
	CUDA_ERROR_NOT_READY = 600,
	CUDA_ERROR_NOT_SUPPORTED = 801,
	CUDA_ERROR_UNKNOWN = 999,
};

typedef unsigned long long CUdeviceptr;
typedef int CUdevice;
typedef uint64_t cuuint64_t;
typedef struct CUctx_st *CUcontext;
typedef struct CUstream_st *CUstream;
typedef struct CUfunc_st *CUfunction;
typedef struct CUmod_st *CUmodule;
typedef struct CUevent_st *CUevent;
typedef struct CUgraphExec_st *CUgraphExec;
typedef struct CUmemPoolHandle_st *CUmemoryPool;
typedef unsigned long long CUmemGenericAllocationHandle;

#define CU_MEM_ATTACH_GLOBAL 0x1u
#define CU_MEMHOSTALLOC_PORTABLE 0x01u
#define CU_MEMHOSTALLOC_DEVICEMAP 0x02u
#define CU_STREAM_NON_BLOCKING 0x1u
#define CU_EVENT_DEFAULT 0x0u
#define CU_EVENT_DISABLE_TIMING 0x2u
#define CU_DEVICE_ATTRIBUTE_MULTIPROCESSOR_COUNT 16
#define CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES 8

typedef enum { CU_MEM_LOCATION_TYPE_DEVICE = 1, CU_MEM_LOCATION_TYPE_HOST = 2 } CUmemLocationType;
typedef enum { CU_MEM_ALLOCATION_TYPE_PINNED = 1 } CUmemAllocationType;
typedef enum { CU_MEM_HANDLE_TYPE_NONE = 0, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR = 1 } CUmemAllocationHandleType;
typedef enum { CU_MEM_ACCESS_FLAGS_PROT_NONE = 0, CU_MEM_ACCESS_FLAGS_PROT_READ = 1,
	       CU_MEM_ACCESS_FLAGS_PROT_READWRITE = 3 } CUmemAccess_flags;
typedef enum { CU_MEM_ALLOC_GRANULARITY_MINIMUM = 0, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED = 1 }
	CUmemAllocationGranularity_flags;

typedef struct {
	CUmemLocationType type;
	int id;
} CUmemLocation;

typedef struct {
	CUmemAllocationType type;
	CUmemAllocationHandleType requestedHandleTypes;
	CUmemLocation location;
	void *win32HandleMetaData;
	struct {
		unsigned char compressionType;
		unsigned char gpuDirectRDMACapable;
		unsigned short usage;
		unsigned char reserved[4];
	} allocFlags;
} CUmemAllocationProp;

typedef struct {
	CUmemLocation location;
	CUmemAccess_flags flags;
} CUmemAccessDesc;

_Static_assert(sizeof(CUmemLocation) == 8, "CUmemLocation ABI");
_Static_assert(sizeof(CUmemAllocationProp) == 32, "CUmemAllocationProp ABI");
_Static_assert(sizeof(CUmemAccessDesc) == 12, "CUmemAccessDesc ABI");

typedef enum {
	CU_GET_PROC_ADDRESS_SUCCESS = 0,
	CU_GET_PROC_ADDRESS_SYMBOL_NOT_FOUND = 1,
	CU_GET_PROC_ADDRESS_VERSION_NOT_SUFFICIENT = 2
} CUdriverProcAddressQueryResult;

/* NVML bits used by the idle detector (reference src/cuda_defs.h:82-104) */
typedef int nvmlReturn_t;
#define NVML_SUCCESS 0
typedef struct nvmlDevice_st *nvmlDevice_t;
typedef struct {
	unsigned int gpu;
	unsigned int memory;
} nvmlUtilization_t;

#endif /* NVS_CUDA_MIN_H */
