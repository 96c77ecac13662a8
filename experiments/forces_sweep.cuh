// forces_sweep.cuh — locality-scheduled variant of the pair kernel (included by forces.cu after the shared pieces).
//
// What limits the plain gather kernel (ncu, round 2, dambreak2m): every CTA takes the next 128 particles of the sorted
// order, and the hardware hands consecutive CTAs to different SMs. The 8 CTAs resident on an SM therefore work in 8
// unrelated places of the current cell plane, each with its own 27-cell neighbourhood (~47 KB of records): together far
// more than the ~100 KB of L1 next to their 8 x 15 KB of shared memory. L1 hit rate 59 %, 20 records pulled through L2
// per particle served, 0.7 eligible warps per cycle: the kernel waits for L2.
//
// Here the SAME per-particle code (walk_section / particle_forces / integrate_epilogue of forces.cu, hence the same
// results bit for bit) is scheduled for the L1:
//  * the work is cut into CHUNKS of up to 32 consecutive particles that lie in one ROW-WINDOW: W1 consecutive cells of
//    one cell row along COORD1 (a row is contiguous in the sorted arrays). Chunks are ordered window by window, inside a
//    window plane by plane, inside a plane row by row;
//  * persistent CTAs; every warp claims the next chunk of ITS SM's queue (a counter per SM, %smid): all warps of an SM
//    sweep the rows of one window together. Consecutive rows share 6 of their 9 neighbour rows, so what the SM's warps
//    gather at any moment is one slab of ~3 planes x ~11 rows x (W1+2) cells: ~150 KB of records;
//  * that fits because the per-thread table of 27 cell bases (13.8 KB per CTA in the gather kernel) becomes a per-WARP
//    table of the few cells a chunk spans (<= W1 cells x 27 bases: 3.4 KB per CTA): ~185 KB of the SM stay L1;
//  * an SM whose queue is empty takes chunks from the other queues, so the tail is a few chunks long.
// The chunk table is made by b200sph_build_neibs (the only call that sees cellEnd) and used by the force evaluations on
// the same cell arrays whose particle range is most of the particles (forces.cu forces_impl).
#pragma once

#ifndef SWEEP_W1
#define SWEEP_W1 6                 // cells per row-window (~113 particles at 18.8 per cell)
#endif
#define SWEEP_WARPS (BLOCK_FORCES / 32)
#ifndef SWEEP_CTAS_PER_SM
#define SWEEP_CTAS_PER_SM 8
#endif

struct SweepGrid { int G1, G2, G3, nw; };
__host__ __device__ __forceinline__ SweepGrid sweep_grid(const DevParams &P)
{
	SweepGrid g;
	g.G1 = P.gridSize[P.coord[0]]; g.G2 = P.gridSize[P.coord[1]]; g.G3 = P.gridSize[P.coord[2]];
	g.nw = (g.G1 + SWEEP_W1 - 1) / SWEEP_W1;
	return g;
}

// particle range of row-window rw = (w * G3 + c3) * G2 + c2
__device__ __forceinline__ void row_window_range(const SweepGrid &g, const uint rw, const uint *__restrict__ cellStart,
	const uint *__restrict__ cellEnd, uint &start, uint &count, uint &hash0)
{
	const int c2 = (int)(rw % (uint)g.G2), c3 = (int)((rw / (uint)g.G2) % (uint)g.G3), w = (int)(rw / ((uint)g.G2 * (uint)g.G3));
	const size_t row = ((size_t)c3 * g.G2 + c2) * g.G1;
	const int c1a = w * SWEEP_W1, c1b = min(c1a + SWEEP_W1, g.G1);
	uint first = CELL_EMPTY, last_end = 0;
	for (int c1 = c1a; c1 < c1b; ++c1) {
		const uint cs = __ldg(cellStart + row + c1);
		if (cs == CELL_EMPTY) continue;
		if (first == CELL_EMPTY) first = cs;
		last_end = __ldg(cellEnd + row + c1);
	}
	hash0 = (uint)(row + c1a);
	if (first == CELL_EMPTY) { start = 0; count = 0; } else { start = first; count = last_end - first; }
}

// pass 1: chunks per row-window;  pass 2 (after an exclusive scan): the chunk table {first particle, count | first cell}
__global__ void __launch_bounds__(BLOCK_STREAM)
sweep_count_kernel(const __grid_constant__ DevParams P, const uint *__restrict__ cellStart, const uint *__restrict__ cellEnd,
	uint *__restrict__ counts, const uint nrw)
{
	const uint rw = blockIdx.x * blockDim.x + threadIdx.x;
	if (rw >= nrw) return;
	uint start, count, hash0;
	row_window_range(sweep_grid(P), rw, cellStart, cellEnd, start, count, hash0);
	counts[rw] = (count + 31u) / 32u;
}

struct SweepChunk { uint start; uint count; uint hash0; uint pad; };    // count <= 32; hash0 = first cell of the row-window

__global__ void __launch_bounds__(BLOCK_STREAM)
sweep_fill_kernel(const __grid_constant__ DevParams P, const uint *__restrict__ cellStart, const uint *__restrict__ cellEnd,
	const uint *__restrict__ offsets, SweepChunk *__restrict__ chunks, const uint nrw)
{
	const uint rw = blockIdx.x * blockDim.x + threadIdx.x;
	if (rw >= nrw) return;
	uint start, count, hash0;
	row_window_range(sweep_grid(P), rw, cellStart, cellEnd, start, count, hash0);
	uint o = offsets[rw];
	for (uint s = 0; s < count; s += 32u, ++o) {
		SweepChunk c;
		c.start = start + s; c.count = min(32u, count - s); c.hash0 = hash0; c.pad = 0;
		chunks[o] = c;
	}
}

__device__ __forceinline__ uint sm_id() { uint v; asm volatile("mov.u32 %0, %%smid;" : "=r"(v)); return v; }

template<int RHODIFF, bool ARTVISC, bool LAMINAR, bool MULTIFLUID, bool WIDE>
__global__ void __launch_bounds__(BLOCK_FORCES, (LAMINAR || MULTIFLUID || WIDE || RHODIFF == RHODIFF_RUNTIME) ? 7 : B200_MIN_BLOCKS)
forces_sweep_kernel(const __grid_constant__ DevParams P, const PosVel *__restrict__ pvArray,
	const ushort4 *__restrict__ infoArray, const uint *__restrict__ particleHash,
	const uint *__restrict__ cellStart, const ushort *__restrict__ neibsList,
	float4 *__restrict__ forces, float *__restrict__ cfl, const BodyOut bo,
	const uint fromParticle, const uint toParticle, const uint cflOffset,
	const SweepChunk *__restrict__ chunks, const uint numChunks, uint *__restrict__ queues, const uint numQueues)
{
	constexpr bool GEN = RHODIFF == RHODIFF_RUNTIME;
	// per warp: the 27 neighbour-cell bases of each of the (at most W1) cells its chunk spans
	__shared__ uint s_cellbase[SWEEP_WARPS][SWEEP_W1][28];
	__shared__ float4 s_celloff[27];
	__shared__ Pinned s_pin;
	const uint lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	if (threadIdx.x < 27) s_celloff[threadIdx.x] = cell_offset(P, threadIdx.x);
	PairConsts k = make_pair_consts<RHODIFF, MULTIFLUID>(P);
	EosConsts E;
	E.gamma = P.gammacoeff[0]; E.sspow = P.sspowercoeff[0]; E.b = P.bcoeff[0]; E.ss = P.sscoeff[0]; E.rho0 = P.rho0[0];
	uint a_off = smem_u32(s_celloff);
	const PosVel *pv = pvArray;
	ListGeom L;
	L.list = neibsList; L.stride = P.stride; L.rows = P.neiblistsize; L.boundpos = P.neibboundpos;
	if (threadIdx.x == 0) {
		float *v = s_pin.v;
		v[0] = k.inv_h; v[1] = k.fc; v[2] = k.R2; v[3] = k.h_alpha; v[4] = k.eps; v[5] = k.g0; v[6] = k.g1; v[7] = k.g2;
		v[8] = k.diff; v[9] = k.grav_scale; v[10] = E.gamma; v[11] = E.sspow; v[12] = E.b; v[13] = E.ss; v[14] = E.rho0;
		v[15] = __uint_as_float(a_off); v[16] = k.h; v[17] = __uint_as_float(L.stride);
		v[18] = __uint_as_float((uint)(uintptr_t)neibsList); v[19] = __uint_as_float((uint)((uintptr_t)neibsList >> 32));
		v[20] = __uint_as_float((uint)(uintptr_t)pvArray); v[21] = __uint_as_float((uint)((uintptr_t)pvArray >> 32));
	}
	__syncthreads();
	{
		const uint a = smem_u32(s_pin.v);
		auto ld = [&](int i) { float x; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(a + 4u * i)); return x; };
		k.inv_h = ld(0); k.fc = ld(1); k.R2 = ld(2); k.h_alpha = ld(3); k.eps = ld(4); k.g0 = ld(5); k.g1 = ld(6); k.g2 = ld(7);
		k.diff = ld(8); k.grav_scale = ld(9); E.gamma = ld(10); E.sspow = ld(11); E.b = ld(12); E.ss = ld(13); E.rho0 = ld(14);
		a_off = __float_as_uint(ld(15)); k.h = ld(16); L.stride = __float_as_uint(ld(17));
		L.list = (const ushort *)((uintptr_t)__float_as_uint(ld(18)) | ((uintptr_t)__float_as_uint(ld(19)) << 32));
		pv = (const PosVel *)((uintptr_t)__float_as_uint(ld(20)) | ((uintptr_t)__float_as_uint(ld(21)) << 32));
	}
	// the queue of this SM: chunks [q * numChunks / numQueues, (q + 1) * numChunks / numQueues)
	uint q = sm_id() % numQueues;
	uint q_lo = (uint)((unsigned long long)q * numChunks / numQueues), q_hi = (uint)((unsigned long long)(q + 1) * numChunks / numQueues);
	uint tried = 0;
	const uint a_base = smem_u32(&s_cellbase[warp][0][0]);

	while (true) {
		uint ci = 0;
		if (lane == 0) ci = q_lo + atomicAdd(queues + q, 1u);
		ci = __shfl_sync(0xffffffffu, ci, 0);
		if (ci >= q_hi) {
			// own queue empty: help the next one (locality no longer matters in the tail)
			if (++tried >= numQueues) break;
			q = (q + 1) % numQueues;
			q_lo = (uint)((unsigned long long)q * numChunks / numQueues); q_hi = (uint)((unsigned long long)(q + 1) * numChunks / numQueues);
			continue;
		}
		const SweepChunk ch = chunks[ci];
		const uint index = ch.start + lane;
		const bool mine = lane < ch.count && index >= fromParticle && index < toParticle;
		// cells this chunk spans, relative to the first cell of its row-window (cells of a row have consecutive hashes)
		uint myslot = 0, cellHash = 0;
		if (lane < ch.count) { cellHash = particleHash[index] & CELLTYPE_BITMASK; myslot = cellHash - ch.hash0; }
		const uint slot_lo = __shfl_sync(0xffffffffu, myslot, 0);
		const uint slot_hi = __shfl_sync(0xffffffffu, myslot, (int)ch.count - 1);
		__syncwarp();
		// the 27 neighbour-cell bases of every spanned cell (lanes 0..26: one neighbour cell each)
		for (uint s = slot_lo; s <= slot_hi; ++s) {
			if (lane < 27) {
				const int h0 = (int)(ch.hash0 + s);
				const int3 gp = grid_pos(P, (uint)h0);
				const int sx = P.hstride[0], sy = P.hstride[1], sz = P.hstride[2];
				const int Gx = P.gridSize[0], Gy = P.gridSize[1], Gz = P.gridSize[2];
				const int cx = (int)lane % 3, cy = ((int)lane / 3) % 3, cz = (int)lane / 9;
				const int dx = cx == 0 ? (gp.x == 0 ? (Gx - 1) * sx : -sx) : (cx == 2 ? (gp.x == Gx - 1 ? -(Gx - 1) * sx : sx) : 0);
				const int dy = cy == 0 ? (gp.y == 0 ? (Gy - 1) * sy : -sy) : (cy == 2 ? (gp.y == Gy - 1 ? -(Gy - 1) * sy : sy) : 0);
				const int dz = cz == 0 ? (gp.z == 0 ? (Gz - 1) * sz : -sz) : (cz == 2 ? (gp.z == Gz - 1 ? -(Gz - 1) * sz : sz) : 0);
				s_cellbase[warp][s][lane] = __ldg(cellStart + (h0 + dx + dy + dz));
			}
		}
		__syncwarp();
		if (mine) {
			const ushort4 info = infoArray[index];
			const int type = ptype_of(info);
			float4 pos, vel;
			ld_posvel(pv + index, pos, vel);
			float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
			bool have_acc = false;
			if ((type == PT_FLUID || type == PT_BOUNDARY) && fabsf(pos.w) < __int_as_float(0x7f800000)) {
				have_acc = true;
				const uint a_mine = a_base + myslot * (28u * 4u);
				auto lut = [=](const uint cell, uint &base, float &ox, float &oy, float &oz) {
					base = lds_u32(a_mine + cell * 4u);
					const float4 o = lds_f4(a_off + cell * 16u);
					ox = o.x; oy = o.y; oz = o.z;
				};
				auto fetch = [&](const uint j, float4 &np, float4 &nv) { ld_posvel(pv + j, np, nv); };
				auto eos = [&](const uint j, const float4 nv) {
					return MULTIFLUID ? eos_from_density(P, nv.w, fluid_num_of(__ldg(infoArray + j))) : eos_from_density(E, nv.w);
				};
				const float cfl_term = particle_forces<RHODIFF, ARTVISC, LAMINAR, MULTIFLUID, GATHER_PF, WIDE, B200_GATHER_AHEAD != 0>(P, k, index, info, type,
					pos, vel, eos_from_density(P, vel.w, MULTIFLUID ? fluid_num_of(info) : 0), cellHash, bo, lut, L, fetch, eos, forces,
					GEN ? bo.xsph : NULL, &acc);
				// one CFL slot per 128 particles like the reference's per-block maxima (slots zeroed by the launcher;
				// non-negative floats order like their bit patterns)
				if (cfl && cfl_term > 0.0f)
					atomicMax(reinterpret_cast<unsigned int *>(cfl + cflOffset + (index - fromParticle) / BLOCK_FORCES), __float_as_uint(cfl_term));
			}
			integrate_epilogue(P, bo, index, info, pos, vel, acc, have_acc, particleHash, forces);
		}
		__syncwarp();
	}
}
