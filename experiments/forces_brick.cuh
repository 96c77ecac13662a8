// forces_brick.cuh — the staged variant of the pair kernel (included by forces.cu after the shared pieces).
//
// The gather kernel's limit is the neighbourhood working set: the 8 CTAs resident on an SM cover ~1000 consecutive
// particles whose 27-cell neighbourhoods (~300 KB of 32-byte records) do not fit the ~100 KB of L1 left next to their
// shared memory, so 40 % of the gathered sectors come from L2 and the kernel waits on them (ncu, round 2: L1 hit 59 %,
// 0.7 eligible warps per cycle, long-scoreboard stalls 10.9 per issue). Here the neighbourhood is made to fit, by shape:
//
//  * the domain is cut into BRICKS of B1 x B2 x B3 cells (4 x 4 x 2: ~560 particles). A brick's neighbourhood is the
//    (B1+2)(B2+2)(B3+2) = 144 cells around it, ~2500 records = 80 KB: 3.9 records staged per particle served, instead of
//    the ~20 sectors per particle the gathers pull through L2 (and 11 staged per particle for a 128-particle tile of
//    one cell row, which is what sank the staged kernel of round 1);
//  * cells of one row along COORD1 are consecutive in the sorted particle arrays, so the neighbourhood is 24 rows of
//    6 cells: each row is staged with up to three TMA bulk copies (cp.async.bulk global -> shared, completion on an
//    mbarrier; three because the first and last cell of a row may wrap around a periodic COORD1);
//  * ONE persistent CTA of 32 warps per SM holds TWO such neighbourhoods (2 x 100 KB of the 227 KB): while the warps
//    work through the particles of one brick in batches of 32 claimed from a shared-memory counter, the next brick's
//    records land in the other buffer. The warp that leaves a buffer last becomes the producer for it (fetches the next
//    brick from a global counter, builds its cell -> slot table, issues the copies); nobody waits for a CTA-wide
//    barrier, there is no pass quantisation and no tail until the bricks run out;
//  * inside a batch a thread walks its particle's neighbour-list column exactly like the gather kernel (same
//    walk_section, same physics, same summation order); only `fetch` differs: two LDS.128 from the staged records.
//  * a brick whose neighbourhood does not fit the buffer (cells can hold 27 particles where the lattice is compressed)
//    is processed from global memory by the same code path with the other `fetch`: nothing is dropped.
//
// The brick list (ids of the bricks with at least one particle, in grid order) is made by b200sph_build_neibs, the only
// call that sees cellEnd; forces calls whose cell_start is the one of that build use it (forces.cu forces_impl).
#pragma once

#ifndef BRK_B1
#define BRK_B1 4
#endif
#ifndef BRK_B2
#define BRK_B2 4
#endif
#ifndef BRK_B3
#define BRK_B3 2
#endif
#define BRK_H1 (BRK_B1 + 2)
#define BRK_H2 (BRK_B2 + 2)
#define BRK_H3 (BRK_B3 + 2)
#define BRK_HC (BRK_H1 * BRK_H2 * BRK_H3)      // cells of a neighbourhood
#define BRK_HR (BRK_H2 * BRK_H3)               // rows of a neighbourhood
#define BRK_NP (3 * BRK_HR)                    // staged pieces: first cell, middle run, last cell of every row
#define BRK_CR (BRK_B2 * BRK_B3)               // rows of the brick itself
#define BRK_WARPS 32
#define BRK_THREADS (32 * BRK_WARPS)
#ifndef BRK_CAP
#define BRK_CAP 3200                           // records per buffer (100 KB)
#endif
#define BRK_NONE 0xFFFFFFFFu

struct BrickBuf {
	uint slotbase[BRK_HC];        // first record of each neighbourhood cell: shared-memory slot (staged) or particle index
	uint cen_start[BRK_CR];       // the brick's own rows: first particle,
	uint cen_off[BRK_CR + 1];     //   prefix of their particle counts (thread -> particle mapping),
	uint cen_hash0[BRK_CR];       //   hash of the row's first brick cell,
	uint cen_h0[BRK_CR];          //   neighbourhood index of that cell
	uint ncen, nbatches, staged, brick;
	uint batch_ctr, done_ctr;
	unsigned long long full_bar;
};

struct BrickSmem {
	PosVel rec[2][BRK_CAP];
	BrickBuf buf[2];
	float4 celloff[27];
	int hdelta[27];
	Pinned pin;
};

// ---- mbarrier / TMA bulk copy (SASS: SYNCS.*, UBLKCP) ----
__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint count)
{ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint bytes)
{ asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
{ asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(unsigned long long *bar, uint parity)
{
	uint ok;
	asm volatile("{\n\t.reg .pred P_OUT;\n\tmbarrier.try_wait.parity.acquire.cta.shared::cta.b64 P_OUT, [%1], %2;\n\tselp.b32 %0, 1, 0, P_OUT;\n\t}"
		: "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
	return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint bytes, unsigned long long *bar)
{
	asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
		:: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// grid geometry in COORD order (COORD1 fastest hash digit)
struct BrickGrid { int G1, G2, G3; int nb1, nb2, nb3; uint per1, per2, per3; };

__host__ __device__ __forceinline__ BrickGrid brick_grid(const DevParams &P)
{
	BrickGrid g;
	g.G1 = P.gridSize[P.coord[0]]; g.G2 = P.gridSize[P.coord[1]]; g.G3 = P.gridSize[P.coord[2]];
	g.nb1 = (g.G1 + BRK_B1 - 1) / BRK_B1; g.nb2 = (g.G2 + BRK_B2 - 1) / BRK_B2; g.nb3 = (g.G3 + BRK_B3 - 1) / BRK_B3;
	g.per1 = (P.periodic >> P.coord[0]) & 1u; g.per2 = (P.periodic >> P.coord[1]) & 1u; g.per3 = (P.periodic >> P.coord[2]) & 1u;
	return g;
}

// brick list: flag of every brick (any particle in its cells?), compacted in grid order by CUB on the host side
__global__ void __launch_bounds__(BLOCK_STREAM)
brick_flags_kernel(const __grid_constant__ DevParams P, const uint *__restrict__ cellStart, unsigned char *__restrict__ flags, const uint nbricks)
{
	const uint b = blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= nbricks) return;
	const BrickGrid g = brick_grid(P);
	const int q1 = (int)(b % (uint)g.nb1), q2 = (int)((b / (uint)g.nb1) % (uint)g.nb2), q3 = (int)(b / ((uint)g.nb1 * (uint)g.nb2));
	bool any = false;
	for (int k = 0; k < BRK_B3 && !any; ++k) for (int j = 0; j < BRK_B2 && !any; ++j) {
		const int c2 = q2 * BRK_B2 + j, c3 = q3 * BRK_B3 + k;
		if (c2 >= g.G2 || c3 >= g.G3) continue;
		for (int i = 0; i < BRK_B1; ++i) {
			const int c1 = q1 * BRK_B1 + i;
			if (c1 < g.G1 && __ldg(cellStart + ((size_t)c3 * g.G2 + c2) * g.G1 + c1) != CELL_EMPTY) { any = true; break; }
		}
	}
	flags[b] = any ? 1 : 0;
}

// One warp prepares buffer `b` for the next brick: claims it, builds the tables, starts the copies.
__device__ __forceinline__ void
brick_produce(const DevParams &P, const BrickGrid &g, BrickSmem &S, const int b, const uint *__restrict__ brickList, const uint numBricks,
	uint *__restrict__ nextBrick, const uint *__restrict__ cellStart, const uint *__restrict__ cellEnd, const PosVel *__restrict__ pv)
{
	const uint lane = threadIdx.x & 31u;
	BrickBuf &B = S.buf[b];
	uint slot = 0;
	if (lane == 0) slot = atomicAdd(nextBrick, 1u);
	slot = __shfl_sync(0xffffffffu, slot, 0);
	if (slot >= numBricks) {
		if (lane == 0) { B.brick = BRK_NONE; B.ncen = 0; B.nbatches = 0; B.batch_ctr = 0; B.done_ctr = 0; mbar_arrive(&B.full_bar); }
		return;
	}
	const uint id = __ldg(brickList + slot);
	const int q1 = (int)(id % (uint)g.nb1), q2 = (int)((id / (uint)g.nb1) % (uint)g.nb2), q3 = (int)(id / ((uint)g.nb1 * (uint)g.nb2));
	const int o1 = q1 * BRK_B1 - 1, o2 = q2 * BRK_B2 - 1, o3 = q3 * BRK_B3 - 1;       // cell of neighbourhood index 0
	// global cell of neighbourhood coordinates (i, j, k), or -1 outside a non-periodic domain / beyond the grid
	auto cell_of = [&](const int i, const int j, const int k) -> long long {
		int c1 = o1 + i, c2 = o2 + j, c3 = o3 + k;
		// a brick may stick out of the grid (G not a multiple of B): those cells do not exist. Wrap only the true halo.
		if (c1 < 0) { if (g.per1) c1 += g.G1; else return -1; } else if (c1 >= g.G1) { if (g.per1 && c1 == g.G1) c1 = 0; else return -1; }
		if (c2 < 0) { if (g.per2) c2 += g.G2; else return -1; } else if (c2 >= g.G2) { if (g.per2 && c2 == g.G2) c2 = 0; else return -1; }
		if (c3 < 0) { if (g.per3) c3 += g.G3; else return -1; } else if (c3 >= g.G3) { if (g.per3 && c3 == g.G3) c3 = 0; else return -1; }
		return ((long long)c3 * g.G2 + c2) * g.G1 + c1;
	};
	if (lane < BRK_CR) { B.cen_start[lane] = 0; B.cen_off[lane + 1] = 0; B.cen_hash0[lane] = 0; B.cen_h0[lane] = 0; }
	__syncwarp();
	// pieces p = lane, lane + 32, lane + 64: row r = p / 3 -> (j, k); part 0: cell i = 0, part 1: cells 1 .. H1-2, part 2: cell H1-1
	uint pstart[3], pcnt[3];
#pragma unroll
	for (int u = 0; u < 3; ++u) {
		const uint p = lane + 32u * u;
		pstart[u] = 0; pcnt[u] = 0;
		if (p < BRK_NP) {
			const int r = (int)(p / 3u), part = (int)(p % 3u), j = r % BRK_H2, k = r / BRK_H2;
			const int i0 = part == 0 ? 0 : (part == 1 ? 1 : BRK_H1 - 1), i1 = part == 1 ? BRK_H1 - 2 : i0;
			uint first = CELL_EMPTY, last_end = 0;
			for (int i = i0; i <= i1; ++i) {
				const long long c = cell_of(i, j, k);
				if (c < 0) continue;
				const uint cs = __ldg(cellStart + c);
				if (cs == CELL_EMPTY) continue;
				if (first == CELL_EMPTY) first = cs;
				last_end = __ldg(cellEnd + c);
			}
			if (first != CELL_EMPTY) { pstart[u] = first; pcnt[u] = last_end - first; }
		}
	}
	// slots: lane-major order of the pieces (any order will do, the table below is the only reader)
	const uint mine = pcnt[0] + pcnt[1] + pcnt[2];
	uint incl = mine;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) { const uint v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (uint)o) incl += v; }
	const uint total = __shfl_sync(0xffffffffu, incl, 31);
	const bool staged = total <= BRK_CAP;
	uint poff[3];
	poff[0] = incl - mine; poff[1] = poff[0] + pcnt[0]; poff[2] = poff[1] + pcnt[1];
	if (lane == 0 && staged) mbar_expect_tx(&B.full_bar, total * (uint)sizeof(PosVel));
	__syncwarp();
#pragma unroll
	for (int u = 0; u < 3; ++u) {
		const uint p = lane + 32u * u;
		if (p >= BRK_NP) continue;
		const int r = (int)(p / 3u), part = (int)(p % 3u), j = r % BRK_H2, k = r / BRK_H2;
		const int i0 = part == 0 ? 0 : (part == 1 ? 1 : BRK_H1 - 1), i1 = part == 1 ? BRK_H1 - 2 : i0;
		if (staged && pcnt[u]) bulk_g2s(&S.rec[b][poff[u]], pv + pstart[u], pcnt[u] * (uint)sizeof(PosVel), &B.full_bar);
		for (int i = i0; i <= i1; ++i) {
			const long long c = cell_of(i, j, k);
			const uint cs = c < 0 ? CELL_EMPTY : __ldg(cellStart + c);
			// an empty cell is never referenced by a list entry; keep its table entry harmless
			B.slotbase[i + BRK_H1 * (j + BRK_H2 * k)] = cs == CELL_EMPTY ? 0u : (staged ? poff[u] + (cs - pstart[u]) : cs);
		}
		// the middle piece of an inner row IS one row of the brick itself (rows of a brick that sticks out of the grid do
		// not exist: with a periodic axis their wrapped images belong to the first brick of that axis)
		if (part == 1 && j >= 1 && j <= BRK_B2 && k >= 1 && k <= BRK_B3 && o2 + j < g.G2 && o3 + k < g.G3) {
			const int cr = (j - 1) + BRK_B2 * (k - 1);
			B.cen_start[cr] = pstart[u];
			B.cen_off[cr + 1] = pcnt[u];                  // counts first, prefix below
			const long long c = cell_of(1, j, k);         // may be -1 when the brick sticks out of the grid (then count = 0)
			B.cen_hash0[cr] = c < 0 ? 0u : (uint)c;
			B.cen_h0[cr] = 1 + BRK_H1 * (j + BRK_H2 * k);
		}
	}
	__syncwarp();
	if (lane == 0) {
		uint acc = 0;
		B.cen_off[0] = 0;
#pragma unroll
		for (int cr = 0; cr < BRK_CR; ++cr) { acc += B.cen_off[cr + 1]; B.cen_off[cr + 1] = acc; }
		B.ncen = acc; B.nbatches = (acc + 31u) / 32u; B.staged = staged ? 1u : 0u; B.brick = id;
		B.batch_ctr = 0; B.done_ctr = 0;
		mbar_arrive(&B.full_bar);                         // releases the table; the copies complete the phase
	}
}

template<int RHODIFF, bool ARTVISC, bool LAMINAR, bool MULTIFLUID, bool WIDE>
__global__ void __launch_bounds__(BRK_THREADS, 1)
forces_brick_kernel(const __grid_constant__ DevParams P, const PosVel *__restrict__ pvArray,
	const ushort4 *__restrict__ infoArray, const uint *__restrict__ particleHash,
	const uint *__restrict__ cellStart, const uint *__restrict__ cellEnd, const ushort *__restrict__ neibsList,
	float4 *__restrict__ forces, float *__restrict__ cfl, const BodyOut bo,
	const uint fromParticle, const uint toParticle, const uint cflOffset,
	const uint *__restrict__ brickList, const uint numBricks, uint *__restrict__ nextBrick)
{
	constexpr bool GEN = RHODIFF == RHODIFF_RUNTIME;
	extern __shared__ __align__(128) unsigned char brick_smem_raw[];
	BrickSmem &S = *reinterpret_cast<BrickSmem *>(brick_smem_raw);
	const uint tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
	const BrickGrid g = brick_grid(P);

	if (tid < 27) {
		S.celloff[tid] = cell_offset(P, tid);
		// cell code = (x+1) + 3(y+1) + 9(z+1): its step in neighbourhood coordinates (COORD order)
		const int d[3] = { (int)tid % 3 - 1, ((int)tid / 3) % 3 - 1, (int)tid / 9 - 1 };
		S.hdelta[tid] = d[P.coord[0]] + BRK_H1 * (d[P.coord[1]] + BRK_H2 * d[P.coord[2]]);
	}
	if (tid == 0) {
		mbar_init(&S.buf[0].full_bar, 1); mbar_init(&S.buf[1].full_bar, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	PairConsts k = make_pair_consts<RHODIFF, MULTIFLUID>(P);
	EosConsts E;
	E.gamma = P.gammacoeff[0]; E.sspow = P.sspowercoeff[0]; E.b = P.bcoeff[0]; E.ss = P.sscoeff[0]; E.rho0 = P.rho0[0];
	uint a_off = smem_u32(S.celloff);
	const PosVel *pv = pvArray;
	ListGeom L;
	L.list = neibsList; L.stride = P.stride; L.rows = P.neiblistsize; L.boundpos = P.neibboundpos;
	// loop constants pinned in registers (see the gather kernel)
	if (tid == 0) {
		float *v = S.pin.v;
		v[0] = k.inv_h; v[1] = k.fc; v[2] = k.R2; v[3] = k.h_alpha; v[4] = k.eps; v[5] = k.g0; v[6] = k.g1; v[7] = k.g2;
		v[8] = k.diff; v[9] = k.grav_scale; v[10] = E.gamma; v[11] = E.sspow; v[12] = E.b; v[13] = E.ss; v[14] = E.rho0;
		v[15] = __uint_as_float(a_off); v[16] = k.h; v[17] = __uint_as_float(L.stride);
		v[18] = __uint_as_float((uint)(uintptr_t)neibsList); v[19] = __uint_as_float((uint)((uintptr_t)neibsList >> 32));
		v[20] = __uint_as_float((uint)(uintptr_t)pvArray); v[21] = __uint_as_float((uint)((uintptr_t)pvArray >> 32));
	}
	__syncthreads();
	{
		const uint a = smem_u32(S.pin.v);
		auto ld = [&](int i) { float x; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(a + 4u * i)); return x; };
		k.inv_h = ld(0); k.fc = ld(1); k.R2 = ld(2); k.h_alpha = ld(3); k.eps = ld(4); k.g0 = ld(5); k.g1 = ld(6); k.g2 = ld(7);
		k.diff = ld(8); k.grav_scale = ld(9); E.gamma = ld(10); E.sspow = ld(11); E.b = ld(12); E.ss = ld(13); E.rho0 = ld(14);
		a_off = __float_as_uint(ld(15)); k.h = ld(16); L.stride = __float_as_uint(ld(17));
		L.list = (const ushort *)((uintptr_t)__float_as_uint(ld(18)) | ((uintptr_t)__float_as_uint(ld(19)) << 32));
		pv = (const PosVel *)((uintptr_t)__float_as_uint(ld(20)) | ((uintptr_t)__float_as_uint(ld(21)) << 32));
	}
	// the first two bricks of this CTA
	if (warp < 2) brick_produce(P, g, S, (int)warp, brickList, numBricks, nextBrick, cellStart, cellEnd, pv);

	for (uint it = 0;; ++it) {
		const int b = (int)(it & 1u);
		BrickBuf &B = S.buf[b];
		while (!mbar_try_wait(&B.full_bar, (it >> 1) & 1u)) { }
		if (B.brick == BRK_NONE) break;
		const uint ncen = B.ncen, nbatches = B.nbatches;
		const bool staged = B.staged != 0;
		const uint a_slot = smem_u32(B.slotbase), a_rec = smem_u32(&S.rec[b][0]), a_hd = smem_u32(S.hdelta);
		while (true) {
			uint batch = 0;
			if (lane == 0) batch = atomicAdd(&B.batch_ctr, 1u);
			batch = __shfl_sync(0xffffffffu, batch, 0);
			if (batch >= nbatches) break;
			const uint t = batch * 32u + lane;
			if (t < ncen) {
				int cr = 0;
#pragma unroll
				for (int r = 1; r < BRK_CR; ++r) cr += (t >= B.cen_off[r]) ? 1 : 0;
				const uint index = B.cen_start[cr] + (t - B.cen_off[cr]);
				if (index >= fromParticle && index < toParticle) {
					const ushort4 info = infoArray[index];
					const int type = ptype_of(info);
					float4 pos, vel;
					ld_posvel(pv + index, pos, vel);
					float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
					bool have_acc = false;
					if ((type == PT_FLUID || type == PT_BOUNDARY) && fabsf(pos.w) < __int_as_float(0x7f800000)) {
						have_acc = true;
						const uint cellHash = particleHash[index] & CELLTYPE_BITMASK;
						// neighbourhood index of the particle's own cell: cells of a row have consecutive hashes
						const uint h0 = B.cen_h0[cr] + (cellHash - B.cen_hash0[cr]);
						auto lut = [=](const uint cell, uint &base, float &ox, float &oy, float &oz) {
							const int hd = (int)lds_u32(a_hd + cell * 4u);
							base = lds_u32(a_slot + (uint)((int)h0 + hd) * 4u);
							const float4 o = lds_f4(a_off + cell * 16u);
							ox = o.x; oy = o.y; oz = o.z;
						};
						auto eos = [&](const uint j, const float4 nv) {
							return MULTIFLUID ? eos_from_density(P, nv.w, fluid_num_of(__ldg(infoArray + j))) : eos_from_density(E, nv.w);
						};
						const float4 e = eos_from_density(P, vel.w, MULTIFLUID ? fluid_num_of(info) : 0);
						float cfl_term;
						if (staged) {
							auto fetch = [=](const uint j, float4 &np, float4 &nv) { np = lds_f4(a_rec + j * 32u); nv = lds_f4(a_rec + j * 32u + 16u); };
							cfl_term = particle_forces<RHODIFF, ARTVISC, LAMINAR, MULTIFLUID, GATHER_PF, WIDE, false>(P, k, index, info, type, pos, vel, e,
								cellHash, bo, lut, L, fetch, eos, forces, GEN ? bo.xsph : NULL, &acc);
						} else {
							auto fetch = [&](const uint j, float4 &np, float4 &nv) { ld_posvel(pv + j, np, nv); };
							cfl_term = particle_forces<RHODIFF, ARTVISC, LAMINAR, MULTIFLUID, GATHER_PF, WIDE, true>(P, k, index, info, type, pos, vel, e,
								cellHash, bo, lut, L, fetch, eos, forces, GEN ? bo.xsph : NULL, &acc);
						}
						// one CFL slot per 128 particles like the reference's per-block maxima (slots zeroed by the launcher;
						// non-negative floats order like their bit patterns)
						if (cfl && cfl_term > 0.0f)
							atomicMax(reinterpret_cast<unsigned int *>(cfl + cflOffset + (index - fromParticle) / BLOCK_FORCES), __float_as_uint(cfl_term));
					}
					integrate_epilogue(P, bo, index, info, pos, vel, acc, have_acc, particleHash, forces);
				}
			}
		}
		// this warp is done with buffer b; the last one to leave refills it
		__syncwarp();
		uint prev = 0;
		if (lane == 0) prev = atomicAdd(&B.done_ctr, 1u);
		prev = __shfl_sync(0xffffffffu, prev, 0);
		if (prev == BRK_WARPS - 1) brick_produce(P, g, S, b, brickList, numBricks, nextBrick, cellStart, cellEnd, pv);
	}
}
