// neibs.cu — neighbour engine: cell hash, radix sort, reorder + cellStart/End, neighbour list.
//
// Behavioural specification: GPUSPH src/cuda/buildneibs.cu + buildneibs_kernel.cu
// (cited per function). Implementation is new: packed-key CUB radix sort instead of
// a comparison sort, constants as __grid_constant__ parameters instead of
// __constant__ symbols + texture references, no host synchronisation.
#include "common.cuh"
#include <cub/device/device_radix_sort.cuh>

// ---------------------------------------------------------------------------
// calcHash — reference calcHashDevice, src/cuda/buildneibs_kernel.cu:664-776
// INTEGER RESULTS MUST BE BIT-EXACT: every float op below is written with explicit
// rounding intrinsics in exactly the form nvcc gives the reference expression
// (IEEE division, separate add, FMA for pos - offset*cellSize).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(BLOCK_STREAM)
calc_hash_kernel(const __grid_constant__ DevParams P, float4 *__restrict__ posArray, uint *__restrict__ particleHash,
	uint *__restrict__ particleIndex, const ushort4 *__restrict__ particleInfo,
	const uint *__restrict__ compactDeviceMap, const uint numParticles)
{
	const uint index = blockIdx.x * blockDim.x + threadIdx.x;
	if (index >= numParticles) return;

	const ushort4 info = particleInfo[index];
	uint gridHash = particleHash[index] & CELLTYPE_BITMASK;

	const bool fluid = ptype_of(info) == PT_FLUID;
	// FLUID || MOVING || (SURFACE && !FLUID), :679
	if (fluid || (info.x & B200SPH_FG_MOVING_BOUNDARY) || (info.x & B200SPH_FG_SURFACE)) {
		float4 pos = posArray[index];
		const int3 gp = grid_pos(P, gridHash);
		float pa[3] = { pos.x, pos.y, pos.z };
		const int gpa[3] = { gp.x, gp.y, gp.z };
		int ng[3];
		bool toofar = false;
#pragma unroll
		for (int a = 0; a < 3; ++a) {
			const float half = pa[a] < 0 ? 0.5f : 0.49999997f;                       // :721-724
			int off = (int)floorf(__fadd_rn(__fdiv_rn(pa[a], P.cellSize[a]), half)); // :725
			int n = gpa[a] + off;
			if (P.periodic & (1u << a)) {        // clampGridPos, :248-256
				if (n < 0) n += P.gridSize[a];
				if (n >= P.gridSize[a]) n -= P.gridSize[a];
			} else {                             // :257-262 / :288-299
				n = min(max(0, n), P.gridSize[a] - 1);
				if (abs(off) > 1 && n == gpa[a]) toofar = true;
				off = n - gpa[a];
			}
			ng[a] = n;
			pa[a] = __fmaf_rn(-(float)off, P.cellSize[a], pa[a]);                    // :750
		}
		gridHash = grid_hash(P, ng[0], ng[1], ng[2]);
		pos.x = pa[0]; pos.y = pa[1]; pos.z = pa[2];
		if (toofar) pos.w = __int_as_float(0x7fffffff);   // disable_particle: mass = NaN (:754)
		if (inactive_w(pos.w)) gridHash = CELL_HASH_MAX;  // :759
		posArray[index] = pos;
	}
	if (compactDeviceMap && gridHash != CELL_HASH_MAX) gridHash |= compactDeviceMap[gridHash];  // :768
	particleHash[index] = gridHash;
	particleIndex[index] = index;
}

extern "C" int b200sph_calc_hash(b200sph_ctx *ctx, void *pos, uint32_t *hash, uint32_t *part_index,
	const void *info, const uint32_t *cdm, uint32_t n)
{
	CHECK_CTX(ctx);
	if (n == 0) return B200SPH_OK;
	if (!pos || !hash || !part_index || !info) { b200_set_error("calcHash: null buffer"); return B200SPH_EINVAL; }
	calc_hash_kernel<<<div_up(n, BLOCK_STREAM), BLOCK_STREAM, 0, ctx->stream>>>(ctx->dp, (float4 *)pos, hash, part_index,
		(const ushort4 *)info, cdm, n);
	KERNEL_TRY();
	return B200SPH_OK;
}

// fixHash — reference fixHashDevice, src/cuda/buildneibs_kernel.cu:790-814
__global__ void __launch_bounds__(BLOCK_STREAM)
fix_hash_kernel(uint *__restrict__ particleHash, uint *__restrict__ particleIndex,
	const uint *__restrict__ compactDeviceMap, const uint numParticles)
{
	const uint index = blockIdx.x * blockDim.x + threadIdx.x;
	if (index >= numParticles) return;
	if (particleHash && compactDeviceMap) {
		const uint h = particleHash[index];
		particleHash[index] = h | compactDeviceMap[h & CELLTYPE_BITMASK];
	}
	particleIndex[index] = index;
}

extern "C" int b200sph_fix_hash(b200sph_ctx *ctx, uint32_t *hash, uint32_t *part_index,
	const void *info, const uint32_t *cdm, uint32_t n)
{
	CHECK_CTX(ctx);
	(void)info;
	if (n == 0) return B200SPH_OK;
	if (!part_index) { b200_set_error("fixHash: null buffer"); return B200SPH_EINVAL; }
	fix_hash_kernel<<<div_up(n, BLOCK_STREAM), BLOCK_STREAM, 0, ctx->stream>>>(hash, part_index, cdm, n);
	KERNEL_TRY();
	return B200SPH_OK;
}

// ---------------------------------------------------------------------------
// sort — reference thrust::sort_by_key with ptype_hash_compare, src/cuda/buildneibs.cu:358-415.
// The comparator is a TOTAL order: (hash incl. high bits, particle type, id). We pack it into one
// 64-bit radix key  [ hash : 32 | ptype : 2 | id : 30 ]  and run CUB's onesweep radix sort on
// (key, partIndex) pairs over all 64 key bits (trimming the range to the live bits would save one of
// eight passes, ~0.07 ms per rebuild at 8 M particles: not done). If some id needs more than
// 30 bits the packed key cannot hold it; then two stable passes (by id, then by hash|ptype) give
// the same order. The result is the unique sorted permutation, hence bit-identical to the
// reference's merge sort.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(BLOCK_STREAM)
make_keys_kernel(const uint *__restrict__ hash, const ushort4 *__restrict__ info, const uint *__restrict__ part_index,
	uint64_t *__restrict__ keys, uint *__restrict__ slots, ushort4 *__restrict__ info_copy, uint *__restrict__ hash_copy,
	uint *__restrict__ pidx_copy, int *__restrict__ wide_id_flag, const uint n)
{
	const uint i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const ushort4 inf = info[i];
	const uint h = hash[i];
	const uint id = id_of(inf);
	if (id >> 30) *wide_id_flag = 1;
	keys[i] = ((uint64_t)h << 32) | ((uint64_t)(ptype_of(inf) & 3) << 30) | (uint64_t)(id & 0x3FFFFFFFu);
	slots[i] = i;
	info_copy[i] = inf;
	hash_copy[i] = h;
	pidx_copy[i] = part_index[i];
}

// keys for the two-pass fallback (ids wider than 30 bits): pass 0 = id, pass 1 = hash|ptype of the
// record currently at position i of the id-ordered sequence
__global__ void __launch_bounds__(BLOCK_STREAM)
make_keys_wide_kernel(const uint *__restrict__ hash_copy, const ushort4 *__restrict__ info_copy,
	const uint *__restrict__ order, uint64_t *__restrict__ keys, const uint n, const int pass)
{
	const uint i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint src = pass == 0 ? i : order[i];
	const ushort4 inf = info_copy[src];
	keys[i] = pass == 0 ? (uint64_t)id_of(inf) : (((uint64_t)hash_copy[src] << 2) | (uint64_t)(ptype_of(inf) & 3));
}

// after the sort: hash, info and partIndex are rewritten in sorted order (the reference sorts them in place)
__global__ void __launch_bounds__(BLOCK_STREAM)
apply_sort_kernel(const uint *__restrict__ sorted_slots, const ushort4 *__restrict__ info_copy,
	const uint *__restrict__ hash_copy, const uint *__restrict__ pidx_copy,
	uint *__restrict__ hash, ushort4 *__restrict__ info, uint *__restrict__ part_index, const uint n)
{
	const uint i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint src = sorted_slots[i];
	hash[i] = __ldg(hash_copy + src);
	info[i] = __ldg(info_copy + src);
	part_index[i] = __ldg(pidx_copy + src);
}

static int ensure_sort_scratch(b200sph_ctx *ctx, uint n)
{
	if (ctx->sort_cap < n) {
		cudaFree(ctx->keys_in); cudaFree(ctx->keys_out); cudaFree(ctx->vals_out); cudaFree(ctx->info_tmp);
		ctx->keys_in = ctx->keys_out = NULL; ctx->vals_out = NULL; ctx->info_tmp = NULL; ctx->sort_cap = 0;
		const size_t cap = (size_t)n + (n >> 3) + 1024;
		CUDA_TRY(cudaMalloc(&ctx->keys_in, cap * sizeof(uint64_t)));
		CUDA_TRY(cudaMalloc(&ctx->keys_out, cap * sizeof(uint64_t)));
		CUDA_TRY(cudaMalloc(&ctx->vals_out, cap * 2 * sizeof(uint32_t)));   // slots in / out
		CUDA_TRY(cudaMalloc(&ctx->info_tmp, cap * (sizeof(ushort4) + 2 * sizeof(uint32_t))));   // info, hash, partIndex copies
		ctx->sort_cap = cap;
	}
	size_t need = 0;
	CUDA_TRY(cub::DeviceRadixSort::SortPairs(NULL, need, (const uint64_t *)NULL, (uint64_t *)NULL,
		(const uint32_t *)NULL, (uint32_t *)NULL, (int)ctx->sort_cap, 0, 64, ctx->stream));
	if (ctx->sort_tmp_bytes < need) {
		cudaFree(ctx->sort_tmp); ctx->sort_tmp = NULL; ctx->sort_tmp_bytes = 0;
		CUDA_TRY(cudaMalloc(&ctx->sort_tmp, need));
		ctx->sort_tmp_bytes = need;
	}
	return B200SPH_OK;
}

extern "C" int b200sph_sort(b200sph_ctx *ctx, uint32_t *hash, void *info, uint32_t *part_index, uint32_t n)
{
	CHECK_CTX(ctx);
	if (n == 0) return B200SPH_OK;
	if (!hash || !info || !part_index) { b200_set_error("sort: null buffer"); return B200SPH_EINVAL; }
	int rc = ensure_sort_scratch(ctx, n);
	if (rc) return rc;
	cudaStream_t s = ctx->stream;
	const uint nb = div_up(n, BLOCK_STREAM);
	ushort4 *info_copy = (ushort4 *)ctx->info_tmp;
	uint32_t *hash_copy = (uint32_t *)(info_copy + ctx->sort_cap);
	uint32_t *pidx_copy = hash_copy + ctx->sort_cap;
	uint32_t *slots_a = ctx->vals_out, *slots_b = ctx->vals_out + ctx->sort_cap;

	CUDA_TRY(cudaMemsetAsync(ctx->d_flag, 0, sizeof(int), s));
	make_keys_kernel<<<nb, BLOCK_STREAM, 0, s>>>(hash, (const ushort4 *)info, part_index, ctx->keys_in, slots_a,
		info_copy, hash_copy, pidx_copy, ctx->d_flag, n);
	KERNEL_TRY();
	size_t tmp = ctx->sort_tmp_bytes;
	CUDA_TRY(cub::DeviceRadixSort::SortPairs(ctx->sort_tmp, tmp, ctx->keys_in, ctx->keys_out,
		(const uint32_t *)slots_a, slots_b, (int)n, 0, 64, s));
	// did any id need more than 30 bits?  (one 4-byte readback; the reference's sort is host-synchronous too)
	CUDA_TRY(cudaMemcpyAsync(ctx->h_flag, ctx->d_flag, sizeof(int), cudaMemcpyDeviceToHost, s));
	CUDA_TRY(cudaStreamSynchronize(s));
	if (*ctx->h_flag) {
		// two stable LSD passes: by id, then by (hash, ptype)
		make_keys_wide_kernel<<<nb, BLOCK_STREAM, 0, s>>>(hash_copy, info_copy, NULL, ctx->keys_in, n, 0);
		KERNEL_TRY();
		tmp = ctx->sort_tmp_bytes;
		CUDA_TRY(cub::DeviceRadixSort::SortPairs(ctx->sort_tmp, tmp, ctx->keys_in, ctx->keys_out,
			(const uint32_t *)slots_a, slots_b, (int)n, 0, 32, s));
		make_keys_wide_kernel<<<nb, BLOCK_STREAM, 0, s>>>(hash_copy, info_copy, slots_b, ctx->keys_in, n, 1);
		KERNEL_TRY();
		tmp = ctx->sort_tmp_bytes;
		CUDA_TRY(cub::DeviceRadixSort::SortPairs(ctx->sort_tmp, tmp, ctx->keys_in, ctx->keys_out,
			(const uint32_t *)slots_b, slots_a, (int)n, 0, 34, s));
		apply_sort_kernel<<<nb, BLOCK_STREAM, 0, s>>>(slots_a, info_copy, hash_copy, pidx_copy, hash, (ushort4 *)info, part_index, n);
		KERNEL_TRY();
		return B200SPH_OK;
	}
	apply_sort_kernel<<<nb, BLOCK_STREAM, 0, s>>>(slots_b, info_copy, hash_copy, pidx_copy, hash, (ushort4 *)info, part_index, n);
	KERNEL_TRY();
	return B200SPH_OK;
}

// ---------------------------------------------------------------------------
// reorder + cellStart/End — reference reorderDataAndFindCellStartDevice,
// src/cuda/buildneibs_kernel.cu:840-992. One thread per sorted slot; the previous slot's
// hash comes straight from global memory (L1-resident neighbour word) instead of a
// shared-memory staging array.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(BLOCK_STREAM)
reorder_kernel(uint *__restrict__ cellStart, uint *__restrict__ cellEnd, uint *__restrict__ segmentStart,
	float4 *__restrict__ sortedPos, float4 *__restrict__ sortedVel,
	const float4 *__restrict__ unsortedPos, const float4 *__restrict__ unsortedVel,
	const uint *__restrict__ particleHash, const uint *__restrict__ particleIndex,
	const uint numParticles, uint *__restrict__ newNumParticles)
{
	const uint index = blockIdx.x * blockDim.x + threadIdx.x;
	if (index >= numParticles) return;
	const uint cellHash = particleHash[index];
	const uint prevHash = index > 0 ? particleHash[index - 1] : 0u;

	if (index == 0 || cellHash != prevHash) {                       // :903-915
		if (cellHash != CELL_HASH_MAX) cellStart[cellHash & CELLTYPE_BITMASK] = index;
		else *newNumParticles = index;
		if (index > 0) cellEnd[prevHash & CELLTYPE_BITMASK] = index;
	}
	if (cellHash == CELL_HASH_MAX) return;                          // :918
	if (index == numParticles - 1) {                                // :921-925
		cellEnd[cellHash & CELLTYPE_BITMASK] = index + 1;
		*newNumParticles = numParticles;
	}
	if (segmentStart) {                                             // :927-933
		const uint ct = cellHash >> 30, pt = prevHash >> 30;
		if (index == 0 || ct != pt) segmentStart[ct] = index;
	}
	const uint src = particleIndex[index];
	sortedPos[index] = __ldg(unsortedPos + src);
	sortedVel[index] = __ldg(unsortedVel + src);
}

template<typename T>
__global__ void __launch_bounds__(BLOCK_STREAM)
gather_extra_kernel(T *__restrict__ dst, const T *__restrict__ src, const uint *__restrict__ particleHash,
	const uint *__restrict__ particleIndex, const uint n)
{
	const uint i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n || particleHash[i] == CELL_HASH_MAX) return;
	dst[i] = __ldg(src + particleIndex[i]);
}

__global__ void init_segments_kernel(uint *segmentStart) { if (threadIdx.x < 4) segmentStart[threadIdx.x] = 0xFFFFFFFFu; }

extern "C" int b200sph_reorder(b200sph_ctx *ctx, uint32_t *cell_start, uint32_t *cell_end, uint32_t *segment_start,
	void *sorted_pos, void *sorted_vel, const void *unsorted_pos, const void *unsorted_vel,
	const b200sph_reorder_extra *extras, uint32_t num_extras,
	const void *sorted_info, const uint32_t *sorted_hash, const uint32_t *part_index,
	uint32_t n, uint32_t *new_num_particles)
{
	CHECK_CTX(ctx);
	(void)sorted_info;
	if (!cell_start || !cell_end || !sorted_pos || !sorted_vel || !unsorted_pos || !unsorted_vel ||
		!sorted_hash || !part_index || !new_num_particles) {
		b200_set_error("reorder: null buffer (the reference throws for a missing mandatory buffer, src/cuda/buildneibs.cu:254-262)");
		return B200SPH_EINVAL;
	}
	if (sorted_pos == unsorted_pos || sorted_vel == unsorted_vel) { b200_set_error("reorder: sorted and unsorted buffers must differ"); return B200SPH_EINVAL; }
	cudaStream_t s = ctx->stream;
	if (segment_start) { init_segments_kernel<<<1, 32, 0, s>>>(segment_start); KERNEL_TRY(); }   // :872
	if (n == 0) return B200SPH_OK;
	const uint nb = div_up(n, BLOCK_STREAM);
	reorder_kernel<<<nb, BLOCK_STREAM, 0, s>>>(cell_start, cell_end, segment_start, (float4 *)sorted_pos, (float4 *)sorted_vel,
		(const float4 *)unsorted_pos, (const float4 *)unsorted_vel, sorted_hash, part_index, n, new_num_particles);
	KERNEL_TRY();
	for (uint32_t e = 0; e < num_extras; ++e) {
		const b200sph_reorder_extra &x = extras[e];
		if (!x.sorted || !x.unsorted) continue;
		switch (x.elem_size) {
		case 4: gather_extra_kernel<uint><<<nb, BLOCK_STREAM, 0, s>>>((uint *)x.sorted, (const uint *)x.unsorted, sorted_hash, part_index, n); break;
		case 8: gather_extra_kernel<uint2><<<nb, BLOCK_STREAM, 0, s>>>((uint2 *)x.sorted, (const uint2 *)x.unsorted, sorted_hash, part_index, n); break;
		case 16: gather_extra_kernel<uint4><<<nb, BLOCK_STREAM, 0, s>>>((uint4 *)x.sorted, (const uint4 *)x.unsorted, sorted_hash, part_index, n); break;
		default: b200_set_error("reorder: unsupported extra element size %u", x.elem_size); return B200SPH_EINVAL;
		}
		KERNEL_TRY();
	}
	return B200SPH_OK;
}

// ---------------------------------------------------------------------------
// neighbour list — reference buildNeibsListDevice + neibsInCell,
// src/cuda/buildneibs_kernel.cu:1029-1185, 538-643. List encoding, slot order and the
// distance predicate (bit-exact float evaluation) follow the reference exactly.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint neib_list_offset(const DevParams &P, uint num, int type)   // :466-478
{
	return type == PT_FLUID ? num : (type == PT_BOUNDARY ? P.neibboundpos - num : num + P.neibboundpos + 1);
}
__device__ __forceinline__ bool too_many_neibs(const DevParams &P, uint nf, uint nb, uint nv, int type)   // :491-515
{
	switch (type) {
	case PT_FLUID: return !(nf < P.neibboundpos);
	case PT_BOUNDARY: return !(nf + nb < P.neibboundpos);
	case PT_VERTEX: return !(nv < P.neiblistsize - P.neibboundpos - 1);
	default: return true;
	}
}


// L2 policy of the list builder. Every list entry is written once and not read again by this kernel, while the
// candidate positions of a cell are re-read by the particles of its 26 neighbours — two of them one whole plane of
// cells later. At 8 M particles a plane's list section (180 MB) streams through the 126 MB L2 between those uses
// and evicts the positions; B200_NL_STORE_CS stores the entries with the streaming (evict-first) policy and
// B200_NL_LOAD_EL loads the candidates evict-last.
#ifndef B200_NL_STORE_CS
#define B200_NL_STORE_CS 0
#endif
#ifndef B200_NL_LOAD_EL
#define B200_NL_LOAD_EL 0
#endif
// B200_NL_GROUP4: distance tests four candidates at a time in the uniform-cell fast path (written at the end of round 1,
// not yet measured on a GPU: off)
#ifndef B200_NL_GROUP4
#define B200_NL_GROUP4 0
#endif
__device__ __forceinline__ void st_list(ushort *p, const ushort v)
{
#if B200_NL_STORE_CS
	asm volatile("st.global.cs.u16 [%0], %1;" :: "l"(p), "h"(v) : "memory");
#else
	*p = v;
#endif
}
__device__ __forceinline__ float4 ld_cand(const float4 *p)
{
#if B200_NL_LOAD_EL
	float4 v;
	unsigned long long pol;
	asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));       // loop-invariant: hoisted by ptxas
	asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
	return v;
#else
	return __ldg(p);
#endif
}

__global__ void __launch_bounds__(BLOCK_STREAM)
build_neibs_kernel(const __grid_constant__ DevParams P, const float4 *__restrict__ posArray,
	const ushort4 *__restrict__ infoArray, const uint *__restrict__ particleHash,
	const uint *__restrict__ cellStart, const uint *__restrict__ cellEnd,
	ushort *__restrict__ neibsList, const uint numParticles, NeibsCounters *__restrict__ counters)
{
	const uint index = blockIdx.x * blockDim.x + threadIdx.x;
	uint nf = 0, nb = 0, nv = 0;
	uint row_step;
	ushort *const column = neibsList + list_column(index, P.stride, P.neiblistsize, P.listblock, row_step);   // rows of THIS particle
	// the search radius is used once per candidate: passing it through a shuffle pins it in a register (ptxas would
	// otherwise re-load it from the constant bank inside the candidate loop)
	const float R2_pinned = __shfl_sync(0xffffffffu, P.nlSqInflRad, 0);

	if (index < numParticles) {
		const ushort4 info = infoArray[index];
		const int mytype = ptype_of(info);
		bool build_nl = mytype == PT_FLUID || mytype == PT_TESTPOINT ||
			(info.x & (B200SPH_FG_MOVING_BOUNDARY | B200SPH_FG_COMPUTE_FORCE));
		if (P.boundarytype == B200SPH_DYN_BOUNDARY) build_nl = true;                    // :1069-1070
		const float4 pos = posArray[index];
		if (build_nl && !inactive_w(pos.w)) {
			const int3 gp = grid_pos(P, particleHash[index] & CELLTYPE_BITMASK);
			const bool boundary = mytype == PT_BOUNDARY;
			const float R2 = R2_pinned;
			for (int z = -1; z <= 1; ++z) for (int y = -1; y <= 1; ++y) for (int x = -1; x <= 1; ++x) {
				int gx = gp.x + x, gy = gp.y + y, gz = gp.z + z;
				// calcNeibCell :317-384
				bool inside = true;
				if (gx < 0) { if (P.periodic & 1) gx = P.gridSize[0] - 1; else inside = false; }
				else if (gx >= P.gridSize[0]) { if (P.periodic & 1) gx = 0; else inside = false; }
				if (gy < 0) { if (P.periodic & 2) gy = P.gridSize[1] - 1; else inside = false; }
				else if (gy >= P.gridSize[1]) { if (P.periodic & 2) gy = 0; else inside = false; }
				if (gz < 0) { if (P.periodic & 4) gz = P.gridSize[2] - 1; else inside = false; }
				else if (gz >= P.gridSize[2]) { if (P.periodic & 4) gz = 0; else inside = false; }
				if (!inside) continue;
				const uint gh = grid_hash(P, gx, gy, gz);
				const uint bucketStart = __ldg(cellStart + gh);
				if (bucketStart == CELL_EMPTY) continue;
				const uint bucketEnd = __ldg(cellEnd + gh);
				const uint cell = (uint)((x + 1) + (y + 1) * 3 + (z + 1) * 9);
				// pos -= gridOffset*d_cellSize (:569), FMA-contracted in the reference build
				const float px = __fmaf_rn(-(float)x, P.cellSize[0], pos.x);
				const float py = __fmaf_rn(-(float)y, P.cellSize[1], pos.y);
				const float pz = __fmaf_rn(-(float)z, P.cellSize[2], pos.z);
				bool encode_cell = true;
				int neib_type = PT_FLUID;
				// Particles of a cell are sorted by type (sort key: cell, type, id), so first == last type means the
				// whole cell has one type: no per-candidate info load, and cells that cannot contribute at all
				// (test points; boundary cells seen from a DYN/LJ boundary particle) are skipped without touching
				// their particles. The list produced is the reference's, entry for entry.
				const int t_first = ptype_of(__ldg(infoArray + bucketStart));
				const int t_last = ptype_of(__ldg(infoArray + bucketEnd - 1));
				const bool uniform = t_first == t_last;
				const bool skip_bb = boundary && (P.boundarytype == B200SPH_DYN_BOUNDARY || P.boundarytype == B200SPH_LJ_BOUNDARY);
				if (uniform && (t_first == PT_TESTPOINT || (skip_bb && t_first == PT_BOUNDARY))) continue;
				// appends candidate j (already known to be inside the search radius) exactly like neibsInCell :618-634
				auto append = [&](const uint j, const int nt) {
					const uint cnt = nt == PT_FLUID ? nf : (nt == PT_BOUNDARY ? nb : nv);
					const uint offset = neib_list_offset(P, cnt, nt);
					if (nt == PT_FLUID) ++nf; else if (nt == PT_BOUNDARY) ++nb; else ++nv;
					if (!too_many_neibs(P, nf, nb, nv, nt)) {                             // :626-634
						const uint enc = encode_cell ? ((cell + 1) << CELLNUM_SHIFT) : 0u;
						st_list(column + offset * row_step, (ushort)((j - bucketStart) + enc));
						encode_cell = false;
					}
				};
				if (uniform) {
					// one particle type in the whole cell (the common case): nothing but the distance test per candidate.
					// The empty asm keeps the cell's base pointer in registers (otherwise it is re-derived from the
					// constant bank for every candidate).
					const float4 *cand = posArray + bucketStart;
					asm volatile("" : "+l"(cand));
					const uint count = bucketEnd - bucketStart;
					const uint self = index - bucketStart;                              // >= count when in another cell
					uint k = 0;
#if B200_NL_GROUP4
					// four candidates per trip: four loads in flight, one branch for the (52 % likely) case that none of
					// them is inside the search radius; accepted candidates are appended in index order like below
					auto sq = [&](const float4 c) {
						const float rx = __fsub_rn(px, c.x), ry = __fsub_rn(py, c.y), rz = __fsub_rn(pz, c.z);
						return __fmaf_rn(rz, rz, __fmaf_rn(ry, ry, __fmul_rn(rx, rx)));
					};
					for (; k + 4 <= count; k += 4) {
						const float4 c0 = ld_cand(cand + k), c1 = ld_cand(cand + k + 1), c2 = ld_cand(cand + k + 2), c3 = ld_cand(cand + k + 3);
						const bool a0 = sq(c0) < R2, a1 = sq(c1) < R2, a2 = sq(c2) < R2, a3 = sq(c3) < R2;
						if (!(a0 | a1 | a2 | a3)) continue;
						if (a0 && k != self && !inactive_w(c0.w)) append(bucketStart + k, t_first);
						if (a1 && k + 1 != self && !inactive_w(c1.w)) append(bucketStart + k + 1, t_first);
						if (a2 && k + 2 != self && !inactive_w(c2.w)) append(bucketStart + k + 2, t_first);
						if (a3 && k + 3 != self && !inactive_w(c3.w)) append(bucketStart + k + 3, t_first);
					}
#endif
#pragma unroll 2
					for (; k < count; ++k) {
						const float4 np = ld_cand(cand + k);
						const float rx = __fsub_rn(px, np.x), ry = __fsub_rn(py, np.y), rz = __fsub_rn(pz, np.z);
						// sqlength(relPos) = x*x + y*y + z*z as nvcc contracts it
						const float r2 = __fmaf_rn(rz, rz, __fmaf_rn(ry, ry, __fmul_rn(rx, rx)));
						if (r2 < R2) {                                                  // :386-392 (85 % of the candidates fail)
							// :553 self, :612 an inactive candidate has a non-finite w
							if (k != self && !inactive_w(np.w)) append(bucketStart + k, t_first);
						}
					}
					continue;
				}
				for (uint j = bucketStart; j < bucketEnd; ++j) {
					if (j == index) continue;
					const int nt = ptype_of(__ldg(infoArray + j));
					if (nt == PT_TESTPOINT) continue;                                     // :584
					if (!encode_cell && neib_type != nt) encode_cell = true;             // :588
					neib_type = nt;
					if (skip_bb && nt == PT_BOUNDARY) continue;                           // :591-602
					const float4 np = ld_cand(posArray + j);
					if (inactive_w(np.w)) continue;                                       // :612
					const float rx = __fsub_rn(px, np.x), ry = __fsub_rn(py, np.y), rz = __fsub_rn(pz, np.z);
					// sqlength(relPos) = x*x + y*y + z*z as nvcc contracts it
					const float r2 = __fmaf_rn(rz, rz, __fmaf_rn(ry, ry, __fmul_rn(rx, rx)));
					if (r2 < R2) append(j, nt);                                           // :386-392
				}
			}
		}
		// end markers / overflow, :1108-1137
		bool overflow = too_many_neibs(P, nf, nb, nv, PT_FLUID);
		const uint marker = overflow ? P.neibboundpos : nf;
		column[marker * row_step] = NEIBS_END;
		overflow |= too_many_neibs(P, nf, nb, nv, PT_BOUNDARY);
		if (!overflow) column[neib_list_offset(P, nb, PT_BOUNDARY) * row_step] = NEIBS_END;
		if (overflow) {
			const int myid = (int)id_of(info);
			atomicCAS(&counters->hasTooManyNeibs, -1, myid);
			if (counters->hasTooManyNeibs == myid) {
				counters->hasMaxNeibs[0] = nf; counters->hasMaxNeibs[1] = nb; counters->hasMaxNeibs[2] = nv;
			}
		}
	}

	// counters, :1140-1185 — warp-shuffle reduction, one atomic per block
	uint total = nf + nb + nv;
	uint mx = nf + nb;
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		total += __shfl_xor_sync(0xffffffffu, total, o);
		mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
	}
	__shared__ uint s_total[BLOCK_STREAM / 32], s_max[BLOCK_STREAM / 32];
	if ((threadIdx.x & 31) == 0) { s_total[threadIdx.x >> 5] = total; s_max[threadIdx.x >> 5] = mx; }
	__syncthreads();
	if (threadIdx.x == 0) {
		uint t = 0, m = 0;
#pragma unroll
		for (int w = 0; w < BLOCK_STREAM / 32; ++w) { t += s_total[w]; m = max(m, s_max[w]); }
		if (t) atomicAdd(&counters->numInteractions, (int)t);
		if (m) atomicMax(&counters->maxFluidBoundaryNeibs, (int)m);
	}
}

__global__ void reset_counters_kernel(NeibsCounters *c)
{	// src/cuda/buildneibs.cu:118-129
	c->numInteractions = 0; c->maxFluidBoundaryNeibs = 0; c->maxVertexNeibs = 0;
	c->hasMaxNeibs[0] = c->hasMaxNeibs[1] = c->hasMaxNeibs[2] = 0;
	c->hasTooManyNeibs = -1;
}

extern "C" int b200sph_neibs_resetinfo(b200sph_ctx *ctx)
{
	CHECK_CTX(ctx);
	reset_counters_kernel<<<1, 1, 0, ctx->stream>>>(ctx->d_counters);
	KERNEL_TRY();
	return B200SPH_OK;
}

extern "C" int b200sph_neibs_getinfo(b200sph_ctx *ctx, b200sph_neibs_info *out)
{	// src/cuda/buildneibs.cu:138-146
	CHECK_CTX(ctx);
	if (!out) { b200_set_error("getinfo: null out"); return B200SPH_EINVAL; }
	NeibsCounters c;
	CUDA_TRY(cudaMemcpyAsync(&c, ctx->d_counters, sizeof(c), cudaMemcpyDeviceToHost, ctx->stream));
	CUDA_TRY(cudaStreamSynchronize(ctx->stream));
	out->num_interactions = c.numInteractions;
	out->max_fluid_boundary_neibs = c.maxFluidBoundaryNeibs;
	out->max_vertex_neibs = c.maxVertexNeibs;
	out->has_too_many_neibs = c.hasTooManyNeibs;
	for (int t = 0; t < 3; ++t) out->has_max_neibs[t] = c.hasMaxNeibs[t];
	return B200SPH_OK;
}

extern "C" int b200sph_build_neibs(b200sph_ctx *ctx, const void *pos, const void *info, const uint32_t *hash,
	const uint32_t *cell_start, const uint32_t *cell_end, uint16_t *neibs_list,
	uint32_t num_particles, uint32_t particle_range_end)
{
	CHECK_CTX(ctx);
	(void)num_particles;
	if (particle_range_end == 0) return B200SPH_OK;
	if (!pos || !info || !hash || !cell_start || !cell_end || !neibs_list) { b200_set_error("buildNeibsList: null buffer"); return B200SPH_EINVAL; }
	if (particle_range_end > ctx->dp.stride) { b200_set_error("buildNeibsList: range end %u exceeds neighbour list stride %u", particle_range_end, ctx->dp.stride); return B200SPH_EINVAL; }
	build_neibs_kernel<<<div_up(particle_range_end, BLOCK_STREAM), BLOCK_STREAM, 0, ctx->stream>>>(ctx->dp,
		(const float4 *)pos, (const ushort4 *)info, hash, cell_start, cell_end, neibs_list, particle_range_end, ctx->d_counters);
	KERNEL_TRY();
	return B200SPH_OK;
}
