/*
 * slab_copy.cu -- the hot kernel of the B200 swap engine (sm_100a only).
 *
 * What it replaces: in the reference, the bytes of an oversubscribed
 * allocation move between HBM and host DRAM inside NVIDIA's UVM driver, one
 * GPU page fault batch at a time, after real_cuMemAllocManaged()
 * (reference src/hook.c:673) and the first touch following LOCK_OK
 * (reference src/client.c:298-307).  Here the move is explicit: the host engine
 * (engine.c) hands this kernel a list of nvs_copy_desc (<= 2 MiB each) and a
 * few resident CTAs stream them.
 *
 * Three entry points, all `extern "C"` so the engine can cuModuleGetFunction
 * them from the embedded cubin:
 *
 *   nvs_slab_copy_tma   one elected lane per warp drives a ring of S smem
 *                       stages with cp.async.bulk (TMA, non-tensor):
 *                         global --(bulk load, mbarrier complete_tx)--> smem
 *                         smem   --(bulk store, bulk_group)-----------> global
 *                       No register staging, no generic-proxy access to the
 *                       payload at all; S-1 loads are always in flight per warp.
 *   nvs_slab_copy_ldg   all lanes, 16-byte ld.global.cs / st.global.cs, 8 loads
 *                       in flight per thread before the first store.
 *   nvs_slab_verify     parity helper: counts 64-bit words that differ from the
 *                       position-dependent pattern (used by tests / bench; it
 *                       can read pinned host and peer memory as well).
 *   nvs_slab_fill       writes that pattern.
 *
 * Bound: link bandwidth (PCIe Gen5 x16 or NVLink 5), not HBM and not tensor
 * cores -- see DESIGN.md.  Work is distributed dynamically (one atomicAdd per
 * descriptor) so CTAs that share an SM with a client's kernels simply take
 * fewer slabs.
 */
#include <stdint.h>
#include "../../include/nvs_copy_desc.h"

#define NVS_MAX_WARPS  8
#define NVS_MAX_STAGES 8

/* ------------------------------------------------------------------ PTX -- */

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
	return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
		     : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
	asm volatile(
		"{\n\t"
		".reg .pred p;\n\t"
		"NVS_WAIT_%=:\n\t"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
		"@p bra NVS_DONE_%=;\n\t"
		"bra NVS_WAIT_%=;\n\t"
		"NVS_DONE_%=:\n\t"
		"}\n" ::"r"(bar),
		"r"(parity)
		: "memory");
}

/* global -> shared::cta, completion counted in bytes on an mbarrier (SASS: UBLKCP) */
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, uint64_t src_gmem, uint32_t bytes,
					 uint32_t bar)
{
	asm volatile(
		"cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
		::"r"(dst_smem), "l"(src_gmem), "r"(bytes), "r"(bar)
		: "memory");
}

/* shared::cta -> global, tracked by the per-thread bulk async-group */
__device__ __forceinline__ void bulk_s2g(uint64_t dst_gmem, uint32_t src_smem, uint32_t bytes)
{
	asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem),
		     "r"(src_smem), "r"(bytes)
		     : "memory");
}

__device__ __forceinline__ void bulk_commit(void)
{
	asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

/* all but the most recent 1 group have finished READING their smem source */
__device__ __forceinline__ void bulk_wait_read_1(void)
{
	asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
}

/* every group fully complete (writes performed) */
__device__ __forceinline__ void bulk_wait_all(void)
{
	asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

/* ------------------------------------------------- TMA (bulk) variant ---- */

/*
 * Launch: grid = any, block = 32 * n_warps (n_warps <= NVS_MAX_WARPS),
 * dynamic smem = n_warps * n_stages * tile_bytes (+ alignment slack of 128).
 * tile_bytes must be a multiple of 16; 2 <= n_stages <= NVS_MAX_STAGES.
 * `counter` points to one zeroed uint32 in device memory per launch.
 */
extern "C" __global__ void __launch_bounds__(32 * NVS_MAX_WARPS, 1)
nvs_slab_copy_tma(const nvs_copy_desc *__restrict__ descs, uint32_t n_descs, uint32_t *counter,
		  uint32_t tile_bytes, uint32_t n_stages)
{
	extern __shared__ __align__(128) uint8_t smem_raw[];
	__shared__ __align__(8) uint64_t full_bar[NVS_MAX_WARPS][NVS_MAX_STAGES];
	__shared__ uint64_t tile_dst[NVS_MAX_WARPS][NVS_MAX_STAGES];
	__shared__ uint32_t tile_len[NVS_MAX_WARPS][NVS_MAX_STAGES];

	const uint32_t warp = threadIdx.x >> 5;
	if ((threadIdx.x & 31u) != 0)
		return; /* one driving lane per warp; TMA does the moving */

	const uint32_t S = n_stages;
	const uint32_t T = tile_bytes;
	const uint32_t ring = smem_u32(smem_raw) + warp * S * T;

	for (uint32_t s = 0; s < S; ++s)
		mbar_init(smem_u32(&full_bar[warp][s]), 1);
	/* make the inits visible to the async proxy before the first bulk copy */
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");

	uint64_t cur_src = 0, cur_dst = 0, rem = 0;
	uint32_t issued = 0, consumed = 0;
	bool drained = false; /* work queue exhausted */

	/* issue the load of the next tile of the descriptor stream; false when none left */
	auto produce = [&]() -> bool {
		while (rem == 0) {
			if (drained)
				return false;
			const uint32_t idx = atomicAdd(counter, 1u);
			if (idx >= n_descs) {
				drained = true;
				return false;
			}
			cur_src = descs[idx].src;
			cur_dst = descs[idx].dst;
			rem = descs[idx].bytes; /* 0-byte descriptors are skipped */
		}
		const uint32_t b = rem < (uint64_t)T ? (uint32_t)rem : T;
		const uint32_t s = issued % S;
		const uint32_t bar = smem_u32(&full_bar[warp][s]);
		tile_dst[warp][s] = cur_dst;
		tile_len[warp][s] = b;
		mbar_expect_tx(bar, b);
		bulk_g2s(ring + s * T, cur_src, b, bar);
		cur_src += b;
		cur_dst += b;
		rem -= b;
		++issued;
		return true;
	};

	/* prologue: S-1 loads in flight */
	for (uint32_t i = 0; i + 1 < S; ++i)
		if (!produce())
			break;

	while (consumed < issued) {
		const uint32_t s = consumed % S;
		mbar_wait(smem_u32(&full_bar[warp][s]), (consumed / S) & 1u);
		bulk_s2g(tile_dst[warp][s], ring + s * T, tile_len[warp][s]);
		bulk_commit();
		++consumed;
		/*
		 * Refill.  Lookahead is S-1, so the stage the next load lands in
		 * (issued % S) was last read by the store of tile consumed - 2:
		 * every bulk group except the one just committed must have finished
		 * reading shared memory.
		 */
		if (!(drained && rem == 0)) {
			bulk_wait_read_1();
			produce();
		}
	}
	bulk_wait_all();
}

/* ------------------------------------------------- LDG/STG variant ------- */

#define NVS_LDG_UNROLL 8

/*
 * Launch: grid = any, block = any multiple of 32 (256 recommended), no dynamic
 * smem.  One CTA takes one descriptor at a time.
 */
extern "C" __global__ void __launch_bounds__(1024)
nvs_slab_copy_ldg(const nvs_copy_desc *__restrict__ descs, uint32_t n_descs, uint32_t *counter)
{
	__shared__ uint32_t s_idx;
	for (;;) {
		__syncthreads();
		if (threadIdx.x == 0)
			s_idx = atomicAdd(counter, 1u);
		__syncthreads();
		const uint32_t idx = s_idx;
		if (idx >= n_descs)
			return;
		const int4 *__restrict__ src = reinterpret_cast<const int4 *>(descs[idx].src);
		int4 *__restrict__ dst = reinterpret_cast<int4 *>(descs[idx].dst);
		const uint64_t n16 = descs[idx].bytes >> 4;
		const uint64_t step = (uint64_t)blockDim.x * NVS_LDG_UNROLL;
		uint64_t base = 0;
		for (; base + step <= n16; base += step) {
			int4 v[NVS_LDG_UNROLL];
#pragma unroll
			for (int u = 0; u < NVS_LDG_UNROLL; ++u)
				v[u] = __ldcs(src + base + (uint64_t)u * blockDim.x + threadIdx.x);
#pragma unroll
			for (int u = 0; u < NVS_LDG_UNROLL; ++u)
				__stcs(dst + base + (uint64_t)u * blockDim.x + threadIdx.x, v[u]);
		}
		for (uint64_t i = base + threadIdx.x; i < n16; i += blockDim.x)
			__stcs(dst + i, __ldcs(src + i));
		/* sub-16-byte tail (never produced by the engine; kept for the C-ABI) */
		const uint64_t tail = descs[idx].bytes & 15ull;
		if (threadIdx.x < tail) {
			const uint8_t *s8 = reinterpret_cast<const uint8_t *>(descs[idx].src) + (n16 << 4);
			uint8_t *d8 = reinterpret_cast<uint8_t *>(descs[idx].dst) + (n16 << 4);
			d8[threadIdx.x] = s8[threadIdx.x];
		}
	}
}

/* ------------------------------------------- same-filled slab elision ----- */

/*
 * A slab whose 64-bit words are all equal (zero-initialised buffers, ones(),
 * padding, freshly memset workspaces -- the reference's own test tensors are
 * torch.ones) does not have to cross the link at all: 8 bytes describe it.
 * HBM is ~120x faster than PCIe Gen5 x16 on this part, so looking at every
 * byte before moving it costs ~1 % of the move it may save.  The copy engines
 * cannot do this; it is what the SMs are for on this path.
 *
 *   nvs_slab_scan    one CTA per slab: out[i] = {word 0, all words equal?},
 *                    leaving the slab at the first 32 KiB tile that differs
 *   nvs_slab_splat   the inverse: fill dst with the 64-bit value held in src
 */
struct nvs_scan_result {
	unsigned long long value;
	unsigned long long is_const;
};

#define NVS_SCAN_UNROLL 8

extern "C" __global__ void __launch_bounds__(256)
nvs_slab_scan(const nvs_copy_desc *__restrict__ descs, uint32_t n_descs, uint32_t *counter,
	      nvs_scan_result *__restrict__ out)
{
	__shared__ uint32_t s_idx;
	for (;;) {
		__syncthreads();
		if (threadIdx.x == 0)
			s_idx = atomicAdd(counter, 1u);
		__syncthreads();
		const uint32_t idx = s_idx;
		if (idx >= n_descs)
			return;
		const ulonglong2 *__restrict__ p = reinterpret_cast<const ulonglong2 *>(descs[idx].src);
		const uint64_t n16 = descs[idx].bytes >> 4;
		const unsigned long long v0 = reinterpret_cast<const unsigned long long *>(descs[idx].src)[0];
		const uint64_t step = (uint64_t)blockDim.x * NVS_SCAN_UNROLL;
		int differs = 0;
		for (uint64_t base = 0; base < n16; base += step) {
			ulonglong2 v[NVS_SCAN_UNROLL];
#pragma unroll
			for (int u = 0; u < NVS_SCAN_UNROLL; ++u) {
				const uint64_t i = base + (uint64_t)u * blockDim.x + threadIdx.x;
				v[u] = i < n16 ? __ldcs(p + i) : make_ulonglong2(v0, v0);
			}
			int d = 0;
#pragma unroll
			for (int u = 0; u < NVS_SCAN_UNROLL; ++u)
				d |= (v[u].x != v0) | (v[u].y != v0);
			if (__syncthreads_or(d)) {
				differs = 1;
				break;
			}
		}
		if (threadIdx.x == 0) {
			out[idx].value = v0;
			out[idx].is_const = !differs && (descs[idx].bytes & 15ull) == 0;
		}
	}
}

extern "C" __global__ void __launch_bounds__(256)
nvs_slab_splat(const nvs_copy_desc *__restrict__ descs, uint32_t n_descs, uint32_t *counter)
{
	__shared__ uint32_t s_idx;
	for (;;) {
		__syncthreads();
		if (threadIdx.x == 0)
			s_idx = atomicAdd(counter, 1u);
		__syncthreads();
		const uint32_t idx = s_idx;
		if (idx >= n_descs)
			return;
		const unsigned long long v = descs[idx].src; /* the value, not an address */
		ulonglong2 *__restrict__ q = reinterpret_cast<ulonglong2 *>(descs[idx].dst);
		const uint64_t n16 = descs[idx].bytes >> 4;
		const ulonglong2 vv = make_ulonglong2(v, v);
		for (uint64_t i = threadIdx.x; i < n16; i += blockDim.x)
			__stcs(q + i, vv);
	}
}

/* ------------------------------------------------- pattern helpers ------- */

/* position-dependent 64-bit pattern: distinguishes every word of every slab */
__host__ __device__ __forceinline__ uint64_t nvs_pattern(uint64_t i, uint64_t seed)
{
	uint64_t z = i + seed * 0x9E3779B97F4A7C15ull + 0x9E3779B97F4A7C15ull;
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}

extern "C" __global__ void nvs_slab_fill(uint64_t *p, uint64_t n_words, uint64_t first_index,
					 uint64_t seed)
{
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += stride)
		p[i] = nvs_pattern(first_index + i, seed);
}

extern "C" __global__ void nvs_slab_verify(const uint64_t *p, uint64_t n_words,
					   uint64_t first_index, uint64_t seed,
					   unsigned long long *mismatches)
{
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	unsigned long long bad = 0;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += stride)
		bad += (p[i] != nvs_pattern(first_index + i, seed));
	if (bad)
		atomicAdd(mismatches, bad);
}
