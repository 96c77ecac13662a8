// forces.cu — forces engine: the fused pair-interaction kernel, CFL reduction, dt.
//
// Behavioural specification: GPUSPH src/cuda/forces.cu + forces_kernel.def (cited inline and in pair_physics.cuh).
// The reference evaluates one half-step with FOUR launches (forcesDevice<fluid,fluid>, <fluid,boundary>,
// <boundary,fluid>, finalizeforcesDevice), each re-reading the particle and read-modify-writing forces[] in
// global memory, and gathers every neighbour's position, velocity and info through three texture fetches. Here:
//  * ONE launch (`forces_gather_kernel`) walks both neighbour-list sections of a particle with the accumulator in
//    registers, applies the finalize step, reduces the CFL term and (b200sph_forces_euler) integrates the particle in
//    its epilogue — forces[] is written exactly once and never read back;
//  * neighbours are gathered as ONE 32-byte record {pos.xyz, mass, vel.xyz, rho~} with one 256-bit load
//    (ld.global.nc.v8.f32 -> LDG.E.256, sm_100): one sector in one line per neighbour instead of two sectors in two
//    lines. The records are an interleaved copy of the reference's pos / vel buffers, made by a streaming pre-pass
//    (`pack_state_kernel`, 64 B per particle) or handed from launch to launch by the fused integration epilogue;
//  * the neighbour's P/rho^2 and sound speed are re-evaluated per pair from rho~ with the reference's __powf
//    expressions sharing one logarithm (12 ALU/MUFU instructions instead of a third gather);
//  * physics options are template parameters, per-pair divisions/roots are MUFU approximations (the bit-exact
//    distance test belongs to the list builder); the lean single-fluid variants work in density ratios and apply the
//    factor diff / (rho_i/rho0) shared by all density-diffusion terms of a particle once, to their sum
//    (`RatioSpace`, pair_physics.cuh);
//  * the loop waits for its gather, not for issue slots: the record of the next list entry is requested before the
//    current one is evaluated, and the lean variants run 8 CTAs per SM (64 registers).
// Summation order inside each list section is the reference's (list order); the fluid and boundary partial sums
// are combined as (0 + sum_fluid) + sum_boundary like the reference's RMW sequence (the ratio-space variants keep the
// density-diffusion terms in a sum of their own, added at the end).
// The neighbour list is read in the blocked layout of include/b200sph.h (B200SPH_NEIBLIST_BLOCK).
#include "pair_physics.cuh"
#include "euler_update.cuh"
#include <stdlib.h>
#include <mutex>
#include <cub/device/device_scan.cuh>


// the per-kernel caches below are process-wide and the engines are shared by one host thread per device
// (SURVEY.md section 8b "Threading")
static std::mutex g_kernel_cache_lock;

// list rows kept in flight by the gather kernel. Round 1 (two 128-bit gathers per neighbour, no record look-ahead),
// dambreak2m / lattice2m: 1 row 0.585 / 0.769 ms, 2 rows 0.535 / 0.805 ms, 4 rows 0.651 / 0.808 ms. Round 2 (one
// 256-bit gather, record look-ahead, 8 CTAs per SM), dambreak2m: 1 row 0.471 ms, 2 rows 0.480 ms, 3 rows 0.70 ms
// (spills): one row - a register is worth more than the second row in flight.
#ifndef GATHER_PF
#define GATHER_PF 1
#endif
// B200_GATHER_AHEAD=1: the neighbour record of the next list entry is requested before the current one is evaluated
#ifndef B200_GATHER_AHEAD
#define B200_GATHER_AHEAD 1
#endif

// ---------------------------------------------------------------------------
// per-particle pre-passes
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(BLOCK_STREAM)
eos_probe_kernel(const __grid_constant__ DevParams P, const float4 *__restrict__ vel, const ushort4 *__restrict__ info,
	float2 *__restrict__ out, const uint n)
{
	const uint i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const float4 e = eos_from_density(P, vel[i].w, fluid_num_of(info[i]));
	out[i] = make_float2(e.x, e.y);
}

extern "C" int b200sph_eos_probe(b200sph_ctx *ctx, const void *vel, const void *info, void *out, uint32_t n)
{
	CHECK_CTX(ctx);
	if (n == 0) return B200SPH_OK;
	if (!vel || !info || !out) { b200_set_error("eos_probe: null buffer"); return B200SPH_EINVAL; }
	eos_probe_kernel<<<div_up(n, BLOCK_STREAM), BLOCK_STREAM, 0, ctx->stream>>>(ctx->dp, (const float4 *)vel, (const ushort4 *)info, (float2 *)out, n);
	KERNEL_TRY();
	return B200SPH_OK;
}

// interleave pos / vel into the 32-byte records the pair kernel gathers (streaming: R 32 + W 32 per particle)
__global__ void __launch_bounds__(BLOCK_STREAM)
pack_state_kernel(const float4 *__restrict__ pos, const float4 *__restrict__ vel, PosVel *__restrict__ pv, const uint from, const uint to)
{
	const uint i = from + blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= to) return;
	st_posvel(pv + i, pos[i], vel[i]);
}

extern "C" int b200sph_pack_state(b200sph_ctx *ctx, const void *pos, const void *vel, void *packed, uint32_t from, uint32_t to)
{
	CHECK_CTX(ctx);
	if (to <= from) return B200SPH_OK;
	if (!pos || !vel || !packed) { b200_set_error("pack_state: null buffer"); return B200SPH_EINVAL; }
	if ((uintptr_t)packed & 31u) { b200_set_error("pack_state: the record buffer must be 32-byte aligned"); return B200SPH_EINVAL; }
	pack_state_kernel<<<div_up(to - from, BLOCK_STREAM), BLOCK_STREAM, 0, ctx->stream>>>((const float4 *)pos, (const float4 *)vel,
		(PosVel *)packed, from, to);
	KERNEL_TRY();
	return B200SPH_OK;
}

// context-owned record buffers (which = 0, 1): the pre-pass target of reference-style calls, and the two state copies of
// the host-buffer step (hoststep.cu)
int b200_packed_scratch(b200sph_ctx *ctx, int which, uint32_t n, PosVel **out)
{
	if (ctx->pv_cap[which] < n) {
		// the buffer may still be read by work in flight on another lane of this context
		CUDA_TRY(cudaDeviceSynchronize());
		cudaFree(ctx->pv[which]); ctx->pv[which] = NULL; ctx->pv_cap[which] = 0;
		const size_t cap = (size_t)n + (n >> 3) + 1024;
		CUDA_TRY(cudaMalloc(&ctx->pv[which], cap * sizeof(PosVel)));
		ctx->pv_cap[which] = cap;
	}
	*out = ctx->pv[which];
	return B200SPH_OK;
}
// ---------------------------------------------------------------------------
// pieces of the pair kernel
// ---------------------------------------------------------------------------
// cellStart of the 27 neighbouring cells of a particle in cell h0 -> out[cell * stride_out] (27 independent loads).
// calcGridHashPeriodic, cellgrid.cuh:174-185 (cells outside a non-periodic domain are never listed, the wrapped
// index only keeps the load in bounds).
__device__ __forceinline__ void load_cell_starts(const DevParams &P, const int h0, const uint *__restrict__ cellStart,
	uint *out, const int stride_out)
{
	const int3 gp = grid_pos(P, (uint)h0);
	const int sx = P.hstride[0], sy = P.hstride[1], sz = P.hstride[2];
	const int Gx = P.gridSize[0], Gy = P.gridSize[1], Gz = P.gridSize[2];
	const int dx[3] = { gp.x == 0 ? (Gx - 1) * sx : -sx, 0, gp.x == Gx - 1 ? -(Gx - 1) * sx : sx };
	const int dy[3] = { gp.y == 0 ? (Gy - 1) * sy : -sy, 0, gp.y == Gy - 1 ? -(Gy - 1) * sy : sy };
	const int dz[3] = { gp.z == 0 ? (Gz - 1) * sz : -sz, 0, gp.z == Gz - 1 ? -(Gz - 1) * sz : sz };
#pragma unroll
	for (int cell = 0; cell < 27; ++cell)
		out[cell * stride_out] = __ldg(cellStart + (h0 + dx[cell % 3] + dy[(cell / 3) % 3] + dz[cell / 9]));
}

// neighbour-cell offset times cell size, for pos_corr = pos - offset*cellSize (cellgrid.cuh:215)
__device__ __forceinline__ float4 cell_offset(const DevParams &P, const int c)
{
	return make_float4((float)(c % 3 - 1) * P.cellSize[0], (float)((c / 3) % 3 - 1) * P.cellSize[1],
		(float)(c / 9 - 1) * P.cellSize[2], 0.f);
}

// Walk one section of a particle's neighbour-list column. fetch(j, np, nv) loads the record of neighbour j =
// base-of-its-cell + offset-in-cell; eos(j, nv) gives its {P/rho^2, sound speed, density, fluid# | pressure};
// lut(cell, base, ox, oy, oz) returns the first particle of neighbour cell `cell` and its offset times the cell size. PF list rows are read ahead of use; the read-ahead offset is clamped to the list.
// WIDE: 64-bit list offsets (lists of 2^31 entries or more), else 32-bit (one VIADDMNMX per row).
// Accumulation order = list order, as in the reference (neibs_iteration.cuh:56-200).
struct ListGeom { const ushort *list; uint stride, rows, boundpos, block; };      // neighbour list + its shape (DevParams copies)
template<bool WIDE> struct ListOff { typedef uint type; typedef int stype; };
template<> struct ListOff<true> { typedef unsigned long long type; typedef long long stype; };

template<bool NFLUID, int RHODIFF, bool ARTVISC, bool LAMINAR, bool MULTIFLUID, int PF, bool WIDE, bool AHEAD, typename Lut, typename Fetch, typename Eos>
__device__ __forceinline__ void
walk_section(const DevParams &P, const PairConsts &k, const Central &c, const uint index, Lut lut,
	const ListGeom &L, Fetch fetch, Eos eos, float4 &acc, float3 &xs)
{
	typedef typename ListOff<WIDE>::type off_t;
	typedef typename ListOff<WIDE>::stype soff_t;
	const ushort *__restrict__ neibsList = L.list;
	// the particle's column in the blocked list layout (include/b200sph.h): first row, distance between rows
	uint row_step;
	const off_t lo = (off_t)list_column(index, L.stride, L.rows, L.block, row_step);
	const off_t stride = (off_t)row_step;
	// element offset of the next row to read ahead: the fluid section grows up from row 0, the boundary section
	// down from neibboundpos; the last (first) row of the column is re-read instead of running off it (the read-ahead
	// must not depend on the entry just read: predicating it on "not the end marker" chains the list loads and
	// was measured 4 % slower)
	const off_t hi = lo + (off_t)(L.rows - 1) * stride;
	off_t off = NFLUID ? lo : lo + (off_t)L.boundpos * stride;
	auto advance = [&](off_t o) -> off_t {
		return NFLUID ? min(o + stride, hi) : (off_t)max((soff_t)(o - stride), (soff_t)lo);
	};
	uint base = 0;
	float pcx = 0.f, pcy = 0.f, pcz = 0.f;
	// PF list rows are kept in flight: the column is a stride-N walk through HBM/L2 (one row = one sector per warp)
	uint q[PF];
#pragma unroll
	for (int i = 0; i < PF; ++i) { q[i] = ld_neib(neibsList + off); off = advance(off); }
	auto next_entry = [&]() -> uint {              // pops the oldest row in flight, reads one more
		const uint nd = q[0];
#pragma unroll
		for (int i = 0; i + 1 < PF; ++i) q[i] = q[i + 1];
		q[PF - 1] = ld_neib(neibsList + off);        // rows past the end marker are never used
		off = advance(off);
		return nd;
	};
	auto decode = [&](uint nd) -> uint {           // list entry -> neighbour index (getNeibIndex, cellgrid.cuh:198-226)
		if (nd >= CELLNUM_ENCODED) {
			const uint cell = (nd >> CELLNUM_SHIFT) - 1;
			nd &= NEIBINDEX_MASK;
			float ox, oy, oz;
			lut(cell, base, ox, oy, oz);
			pcx = c.pos.x - ox; pcy = c.pos.y - oy; pcz = c.pos.z - oz;
		}
		return base + nd;
	};
	if (AHEAD) {
	// software pipeline: the record of entry i+1 is requested before entry i is evaluated, so that a warp has one
	// gather in flight while it computes (the gather kernel is latency-bound on exactly that load)
	uint nd = next_entry();
	if (nd == NEIBS_END) return;
	uint j = decode(nd);
	float4 np, nv;
	fetch(j, np, nv);
	float rx = pcx - np.x, ry = pcy - np.y, rz = pcz - np.z;
	while (true) {
		const uint nd2 = next_entry();
		const bool more = nd2 != NEIBS_END;
		float4 np2 = np, nv2 = nv;
		uint j2 = j;
		if (more) { j2 = decode(nd2); fetch(j2, np2, nv2); }
		const float r2 = fmaf(rz, rz, fmaf(ry, ry, rx * rx));
		// skip inactive neighbours and pairs beyond the kernel support (forces_kernel.def:3987-3999)
		if ((r2 < k.R2) && (fabsf(np.w) < __int_as_float(0x7f800000)))
			pair_interaction_x<RHODIFF, ARTVISC, LAMINAR, MULTIFLUID>(P, k, c, rx, ry, rz, r2, np.w,
				make_float4(c.vel.x - nv.x, c.vel.y - nv.y, c.vel.z - nv.z, RatioSpace<RHODIFF, LAMINAR, MULTIFLUID>::value ? nv.w + 1.0f : nv.w),
				eos(j, make_float4(0.f, 0.f, 0.f, RatioSpace<RHODIFF, LAMINAR, MULTIFLUID>::value ? nv.w + 1.0f : nv.w)), NFLUID, acc, xs);
		if (!more) break;
		np = np2; nv = nv2; j = j2;
		rx = pcx - np.x; ry = pcy - np.y; rz = pcz - np.z;
	}
	} else {
	while (true) {
		const uint nd = next_entry();
		if (nd == NEIBS_END) break;
		const uint j = decode(nd);
		float4 np, nv;
		fetch(j, np, nv);
		const float rx = pcx - np.x, ry = pcy - np.y, rz = pcz - np.z;
		const float r2 = fmaf(rz, rz, fmaf(ry, ry, rx * rx));
		// skip inactive neighbours and pairs beyond the kernel support (forces_kernel.def:3987-3999)
		if (!(r2 < k.R2) || !(fabsf(np.w) < __int_as_float(0x7f800000))) continue;
		pair_interaction_x<RHODIFF, ARTVISC, LAMINAR, MULTIFLUID>(P, k, c, rx, ry, rz, r2, np.w, make_float4(c.vel.x - nv.x, c.vel.y - nv.y, c.vel.z - nv.z, nv.w), eos(j, nv), NFLUID, acc, xs);
	}
	}
}

// all sections of one particle (forces.cu:759,782,792 in the reference) + finalize; returns the CFL term
template<int RHODIFF, bool ARTVISC, bool LAMINAR, bool MULTIFLUID, int PF, bool WIDE, bool AHEAD, typename Lut, typename Fetch, typename Eos>
__device__ __forceinline__ float
particle_forces(const DevParams &P, const PairConsts &k, const uint index, const ushort4 info, const int type,
	const float4 pos, const float4 vel, const float4 e, const uint cellHash, const BodyOut &bo, Lut lut,
	const ListGeom &L, Fetch fetch, Eos eos, float4 *__restrict__ forces, float4 *__restrict__ xsph = NULL, float4 *acc_out = NULL)
{
	Central c;
	c.pos = pos; c.vel = vel;
	c.p_precalc = e.x; c.sspeed = e.y; c.rho = e.z;
	c.ratio = vel.w + 1.0f;
	c.fnum = MULTIFLUID ? __float_as_int(e.w) : 0;
	// the Molteni-Colagrossi switch compares raw pressures (forces_kernel.def:1925-1928)
	c.press = (RHODIFF == B200SPH_RHODIFF_COLAGROSSI || RHODIFF == RHODIFF_RUNTIME) ? eos_pressure(P, vel.w, c.fnum) : 0.0f;
	c.xsph = false;
	float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
	float3 xs = make_float3(0.f, 0.f, 0.f);
	if (type == PT_FLUID) {
		// fluid<-fluid then fluid<-boundary; DYN boundary neighbours interact like fluid ones (forces_kernel.def:3717-3726)
		c.momentum = true;
		c.xsph = RHODIFF == RHODIFF_RUNTIME && xsph != NULL;
		walk_section<true, RHODIFF, ARTVISC, LAMINAR, MULTIFLUID, PF, WIDE, AHEAD>(P, k, c, index, lut, L, fetch, eos, acc, xs);
		walk_section<false, RHODIFF, ARTVISC, LAMINAR, MULTIFLUID, PF, WIDE, AHEAD>(P, k, c, index, lut, L, fetch, eos, acc, xs);
		// write_xsph :3366-3368
		if (c.xsph) xsph[index] = make_float4(2.0f * xs.x, 2.0f * xs.y, 2.0f * xs.z, 0.0f);
	} else {
		// boundary<-fluid: density always, momentum only with force feedback (forces_kernel.def:3634-3667)
		c.momentum = (info.x & B200SPH_FG_COMPUTE_FORCE) != 0;
		walk_section<true, RHODIFF, ARTVISC, LAMINAR, MULTIFLUID, PF, WIDE, AHEAD>(P, k, c, index, lut, L, fetch, eos, acc, xs);
	}
	// ratio-space variants: the density-diffusion sum of the particle, scaled once (pair_physics.cuh)
	if (RatioSpace<RHODIFF, LAMINAR, MULTIFLUID>::value && RHODIFF != B200SPH_RHODIFF_NONE)
		acc.w = fmaf(k.diff / c.ratio, xs.x, acc.w);
	const float cfl_term = finalize_particle(P, type, c.fnum, c.sspeed, info, pos, vel, c.rho, cellHash, bo, acc);
	forces[index] = acc;
	if (acc_out) *acc_out = acc;
	return cfl_term;
}

// Fused integration epilogue: exactly the stand-alone euler kernel's update of this particle (euler_update.cuh), with
// the forces still in registers. A particle the pair loop skips integrates with its FORCES entry as is.
__device__ __forceinline__ void
integrate_epilogue(const DevParams &P, const BodyOut &bo, const uint index, const ushort4 info, const float4 pos, const float4 vel,
	float4 acc, const bool have_acc, const uint *__restrict__ particleHash, const float4 *__restrict__ forces)
{
	if (!bo.eul_step) return;
	if (!have_acc) acc = forces[index];
	float4 p = pos, v = vel;
	if (bo.eul_old_pos) { p = bo.eul_old_pos[index]; v = bo.eul_old_vel[index]; }
	const float4 none = make_float4(0.f, 0.f, 0.f, 0.f);
	if (bo.eul_step == 1)
		euler_update<1>(P, p, v, acc, info, particleHash, index, euler_dt<1>(bo.eul_state, bo.eul_dt), bo.eul_bodies, false, none);
	else
		euler_update<2>(P, p, v, acc, info, particleHash, index, euler_dt<2>(bo.eul_state, bo.eul_dt), bo.eul_bodies, false, none);
	bo.eul_new_pos[index] = p;
	bo.eul_new_vel[index] = v;
	// the integrated state as the record the NEXT force evaluation gathers (no pack pre-pass between launches)
	if (bo.eul_new_packed) st_posvel(bo.eul_new_packed + index, p, v);
}

// ---------------------------------------------------------------------------
// gather kernel: neighbours through L1/L2
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint smem_u32(const void *p) { return (uint)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint lds_u32(uint a) { uint v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ float4 lds_f4(uint a)
{ float4 v; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a)); return v; }
// Loop-invariant kernel parameters live in the constant bank; on sm_100 every use inside the pair loop costs an
// LDCU issue slot (ALU instructions no longer take constant-bank operands). B200_HOIST=1 pins the ones the pair
// loop needs in vector registers by passing them through a warp shuffle, which ptxas will not rematerialise.
#ifndef B200_HOIST
#define B200_HOIST 1
#endif
#ifndef B200_PIN_CELLBASE
#define B200_PIN_CELLBASE 1
#endif
#ifndef B200_MIN_BLOCKS
#define B200_MIN_BLOCKS 8
#endif
// (the detour through shared memory is what stops ptxas from re-deriving the value from the constant bank)
struct Pinned { float v[24]; };

// CTAs per SM of the lean variants. The loop waits on its gather (one record in flight per warp; long-scoreboard stalls
// 2.5-4.4 per issued instruction, profiles/r02_forces_gather_final_ncu.txt), so warps per SM count for more than
// instructions per pair: trimming the loop from 114 to 96 instructions (ratio-space physics, pinned shared-memory
// address) at 7 CTAs / 72 registers bought 1.5 %; the same loop at 8 CTAs / 64 registers 8 % (dambreak2m, forces launch:
// 6 CTAs 0.560 ms, 7 0.507, 8 0.471, 9 0.478 with the constants left in the constant bank (0.74 pinned: spills in the
// loop), 10 0.528). The variants with more live state keep 7 (Colagrossi + artificial viscosity would spill inside the loop).
template<int RHODIFF, bool ARTVISC, bool LAMINAR, bool MULTIFLUID, bool WIDE>
__global__ void __launch_bounds__(BLOCK_FORCES, (LAMINAR || MULTIFLUID || WIDE || RHODIFF == RHODIFF_RUNTIME ||
	(RHODIFF == B200SPH_RHODIFF_COLAGROSSI && ARTVISC)) ? 7 : B200_MIN_BLOCKS)
forces_gather_kernel(const __grid_constant__ DevParams P, const PosVel *__restrict__ pvArray,
	const ushort4 *__restrict__ infoArray, const uint *__restrict__ particleHash,
	const uint *__restrict__ cellStart, const ushort *__restrict__ neibsList,
	float4 *__restrict__ forces, float *__restrict__ cfl, const BodyOut bo,
	const uint fromParticle, const uint toParticle, const uint cflOffset)
{
	constexpr bool GEN = RHODIFF == RHODIFF_RUNTIME;
	__shared__ uint s_cellbase[27 * BLOCK_FORCES];
	__shared__ float4 s_celloff[27];
	const uint index = blockIdx.x * blockDim.x + threadIdx.x + fromParticle;
	float cfl_term = 0.0f;
	if (threadIdx.x < 27) s_celloff[threadIdx.x] = cell_offset(P, threadIdx.x);
	constexpr bool FAST = RatioSpace<RHODIFF, LAMINAR, MULTIFLUID>::value;
	PairConsts k = make_pair_consts<RHODIFF, MULTIFLUID, LAMINAR>(P);
	EosConsts E;
	E.gamma = P.gammacoeff[0]; E.sspow = P.sspowercoeff[0]; E.b = P.bcoeff[0]; E.ss = P.sscoeff[0]; E.rho0 = P.rho0[0];
	if (FAST) { E.b = P.bcoeff[0] / (P.rho0[0] * P.rho0[0]); E.rho0 = P.bcoeff[0]; }     // eos_ratio_from_density
	uint a_off = smem_u32(s_celloff);
	uint a_cellbase = smem_u32(s_cellbase);
	const PosVel *pv = pvArray;       // pinned below: the record base is used by every gather
	ListGeom L;
	L.list = neibsList; L.stride = P.stride; L.rows = P.neiblistsize; L.boundpos = P.neibboundpos; L.block = P.listblock;
#if B200_HOIST
	__shared__ Pinned s_pin;
	if (threadIdx.x == 0) {
		float *v = s_pin.v;
		v[0] = k.inv_h; v[1] = k.fc; v[2] = k.R2; v[3] = k.h_alpha; v[4] = k.eps; v[5] = k.g0; v[6] = k.g1; v[7] = k.g2;
		v[8] = k.diff; v[9] = k.grav_scale; v[10] = E.gamma; v[11] = E.sspow; v[12] = E.b; v[13] = E.ss; v[14] = E.rho0;
		v[15] = __uint_as_float(a_off); v[16] = k.h; v[17] = __uint_as_float(L.stride);
		v[18] = __uint_as_float((uint)(uintptr_t)neibsList); v[19] = __uint_as_float((uint)((uintptr_t)neibsList >> 32));
		v[20] = __uint_as_float((uint)(uintptr_t)pvArray); v[21] = __uint_as_float((uint)((uintptr_t)pvArray >> 32));
		v[22] = __uint_as_float(smem_u32(s_cellbase));
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
		a_cellbase = __float_as_uint(ld(22));
	}
#else
	__syncthreads();
#endif

	if (index < toParticle) {
		const ushort4 info = infoArray[index];
		const int type = ptype_of(info);
		float4 pos, vel;
		ld_posvel(pv + index, pos, vel);
		float4 acc;
		bool have_acc = false;
		if ((type == PT_FLUID || type == PT_BOUNDARY) && fabsf(pos.w) < __int_as_float(0x7f800000)) {
			have_acc = true;
			uint *my_base = s_cellbase + threadIdx.x;
			const uint cellHash = particleHash[index] & CELLTYPE_BITMASK;
			load_cell_starts(P, (int)cellHash, cellStart, my_base, BLOCK_FORCES);
			// 32-bit shared-memory addresses computed once (the generic-pointer form is re-derived per use)
#if B200_PIN_CELLBASE
			const uint a_base = a_cellbase + 4u * threadIdx.x;
#else
			const uint a_base = smem_u32(my_base);
#endif
			auto lut = [=](const uint cell, uint &base, float &ox, float &oy, float &oz) {
				base = lds_u32(a_base + cell * (BLOCK_FORCES * 4u));
				const float4 o = lds_f4(a_off + cell * 16u);
				ox = o.x; oy = o.y; oz = o.z;
			};
			// one 256-bit gather per neighbour: its {pos, mass, vel, rho~} record; EOS terms from rho~ on the fly
			auto fetch = [&](const uint j, float4 &np, float4 &nv) { ld_posvel(pv + j, np, nv); };
			auto eos = [&](const uint j, const float4 nv) {
				if (FAST) return eos_ratio_from_density(E, nv.w, B200_GATHER_AHEAD != 0);
				return MULTIFLUID ? eos_from_density(P, nv.w, fluid_num_of(__ldg(infoArray + j))) : eos_from_density(E, nv.w);
			};
			cfl_term = particle_forces<RHODIFF, ARTVISC, LAMINAR, MULTIFLUID, GATHER_PF, WIDE, B200_GATHER_AHEAD != 0>(P, k, index, info, type, pos,
				vel, eos_from_density(P, vel.w, MULTIFLUID ? fluid_num_of(info) : 0), cellHash, bo, lut, L, fetch, eos, forces,
				GEN ? bo.xsph : NULL, &acc);
		}
		// (the lean variants do not keep rho~ of the particle across the pair loop: one more register for the loop)
		if (FAST && bo.eul_step && have_acc) vel.w = __ldg(&pv[index].vel.w);
		integrate_epilogue(P, bo, index, info, pos, vel, acc, have_acc, particleHash, forces);
	}

	// block max (maxBlockReduce, device_core.cu:40-59) with warp shuffles
	if (cfl) {
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) cfl_term = fmaxf(cfl_term, __shfl_xor_sync(0xffffffffu, cfl_term, o));
		__shared__ float s_max[BLOCK_FORCES / 32];
		if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = cfl_term;
		__syncthreads();
		if (threadIdx.x == 0) {
			float m = s_max[0];
#pragma unroll
			for (int w = 1; w < BLOCK_FORCES / 32; ++w) m = fmaxf(m, s_max[w]);
			cfl[cflOffset + blockIdx.x] = m;
		}
	}
}


// ---------------------------------------------------------------------------
// launcher
// ---------------------------------------------------------------------------
typedef void (*gather_kernel_t)(const DevParams, const PosVel *, const ushort4 *, const uint *,
	const uint *, const ushort *, float4 *, float *, const BodyOut, const uint, const uint, const uint);
template<int RHODIFF>
static void pick_kernels(bool artvisc, bool laminar, bool multi, gather_kernel_t *g /* [wide] */)
{
#define PICK(A, L, M) do { g[0] = forces_gather_kernel<RHODIFF, A, L, M, false>; g[1] = forces_gather_kernel<RHODIFF, A, L, M, true>; } while (0)
	if (multi) {
		if (artvisc) { if (laminar) PICK(true, true, true); else PICK(true, false, true); }
		else { if (laminar) PICK(false, true, true); else PICK(false, false, true); }
	} else {
		if (artvisc) { if (laminar) PICK(true, true, false); else PICK(true, false, false); }
		else { if (laminar) PICK(false, true, false); else PICK(false, false, false); }
	}
#undef PICK
}

// Shared-memory carve-out of the gather kernel = what the CTAs its register count allows per SM actually need (the
// rest of the 228 KB stays L1 for the neighbour gathers). Said explicitly because a host application may have set
// a device-wide cache preference (GPUSPH sets cudaFuncCachePreferL1, src/cuda/cudautil.cc:71-79), which would
// otherwise shrink the carve-out and cut the occupancy of this kernel several times over.
static int gather_carveout(const void *kernel)      // caller holds g_kernel_cache_lock
{
	static const void *known[128]; static int pct[128]; static int nknown = 0;
	for (int i = 0; i < nknown; ++i) if (known[i] == kernel) return pct[i];
	int result = 66;
	cudaFuncAttributes fa;
	if (cudaFuncGetAttributes(&fa, kernel) == cudaSuccess && fa.numRegs > 0) {
		const int regs = (fa.numRegs + 7) / 8 * 8;
		int blocks = 65536 / (regs * BLOCK_FORCES);
		if (blocks > 2048 / BLOCK_FORCES) blocks = 2048 / BLOCK_FORCES;
		const size_t per_block = fa.sharedSizeBytes + 1024;           // + the per-CTA reservation
		while (blocks > 1 && blocks * per_block > 227 * 1024) --blocks;
		if (const char *e = getenv("B200SPH_FORCES_BLOCKS")) { const int b = atoi(e); if (b > 0 && b < blocks) blocks = b; }
		result = (int)((blocks * per_block * 100 + 228 * 1024 - 1) / (228 * 1024));
		if (result > 100) result = 100;
	}
	if (nknown < 128) { known[nknown] = kernel; pct[nknown] = result; ++nknown; }
	return result;
}

// the carve-out preference is a per-device function attribute: set it once per (kernel, device), not per launch
// (cudaFuncSetAttribute costs as much host time as the launch itself, and the striped host-buffer step launches 16
// grids per time step)
static int set_carveout(const b200sph_ctx *ctx, const void *kernel)
{
	static const void *known[256]; static int dev[256]; static int nknown = 0;
	std::lock_guard<std::mutex> guard(g_kernel_cache_lock);
	for (int i = 0; i < nknown; ++i) if (known[i] == kernel && dev[i] == ctx->device) return B200SPH_OK;
	CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, gather_carveout(kernel)));
	if (nknown < 256) { known[nknown] = kernel; dev[nknown] = ctx->device; ++nknown; }
	return B200SPH_OK;
}

extern "C" int b200sph_forces(b200sph_ctx *ctx, const void *pos, const void *vel, const void *info,
	const uint32_t *hash, const uint32_t *cell_start, const uint16_t *neibs_list,
	void *forces, float *cfl, uint32_t num_particles, uint32_t from, uint32_t to,
	uint32_t cfl_offset, uint32_t *num_cfl_blocks)
{
	return b200sph_forces_bodies(ctx, pos, vel, info, hash, cell_start, neibs_list, forces, cfl, NULL, NULL,
		num_particles, from, to, cfl_offset, num_cfl_blocks);
}

extern "C" int b200sph_forces_bodies(b200sph_ctx *ctx, const void *pos, const void *vel, const void *info,
	const uint32_t *hash, const uint32_t *cell_start, const uint16_t *neibs_list,
	void *forces, float *cfl, void *rb_forces, void *rb_torques, uint32_t num_particles, uint32_t from, uint32_t to,
	uint32_t cfl_offset, uint32_t *num_cfl_blocks)
{
	b200sph_forces_args a;
	memset(&a, 0, sizeof(a));
	a.pos = pos; a.vel = vel; a.info = info; a.hash = hash; a.cell_start = cell_start; a.neibs_list = neibs_list;
	a.forces = forces; a.cfl = cfl; a.rb_forces = rb_forces; a.rb_torques = rb_torques;
	a.num_particles = num_particles; a.from_particle = from; a.to_particle = to; a.cfl_offset = cfl_offset;
	return b200sph_forces_ex(ctx, &a, num_cfl_blocks);
}

static int forces_impl(b200sph_ctx *ctx, const b200sph_forces_args *args, const b200sph_fused_euler_args *eul, uint32_t *num_cfl_blocks);

extern "C" int b200sph_forces_ex(b200sph_ctx *ctx, const b200sph_forces_args *args, uint32_t *num_cfl_blocks)
{
	return forces_impl(ctx, args, NULL, num_cfl_blocks);
}

// forces of [from, to) followed by the integration of the same particles (b200sph_forces_ex + b200sph_euler_ex): one
// launch (fused epilogue) unless ENABLE_XSPH is on, whose mean velocity the integration reads from the buffer
extern "C" int b200sph_forces_euler(b200sph_ctx *ctx, const b200sph_forces_args *args, const b200sph_fused_euler_args *eul,
	uint32_t *num_cfl_blocks)
{
	if (!eul) { b200_set_error("forces_euler: null integration arguments"); return B200SPH_EINVAL; }
	return forces_impl(ctx, args, eul, num_cfl_blocks);
}

static int forces_impl(b200sph_ctx *ctx, const b200sph_forces_args *args, const b200sph_fused_euler_args *eul, uint32_t *num_cfl_blocks)
{
	CHECK_CTX(ctx);
	if (!args) { b200_set_error("forces: null argument block"); return B200SPH_EINVAL; }
	const void *pos = args->pos, *vel = args->vel, *info = args->info;
	const uint32_t *hash = args->hash, *cell_start = args->cell_start;
	const uint16_t *neibs_list = args->neibs_list;
	void *forces = args->forces; float *cfl = args->cfl;
	void *rb_forces = args->rb_forces, *rb_torques = args->rb_torques;
	const uint32_t num_particles = args->num_particles, from = args->from_particle, to = args->to_particle, cfl_offset = args->cfl_offset;
	BodyOut bo;
	memset(&bo, 0, sizeof(bo));
	bo.rb_forces = (float4 *)rb_forces; bo.rb_torques = (float4 *)rb_torques;
	if (eul) {
		if (eul->step != 1 && eul->step != 2) { b200_set_error("unsupported predcorr timestep %d", eul->step); return B200SPH_EINVAL; }
		if (args->to_particle > args->from_particle) {
			if (!eul->old_pos || !eul->old_vel || !eul->new_pos || !eul->new_vel) { b200_set_error("forces_euler: null buffer"); return B200SPH_EINVAL; }
			// [new + from, new + to) against [gathered, gathered + num_particles): any overlap, not just equal bases
			auto overlaps = [&](const void *w, size_t wsz, const void *r, size_t rsz) {
				if (!w || !r) return false;
				const char *w0 = (const char *)w + (size_t)args->from_particle * wsz, *w1 = (const char *)w + (size_t)args->to_particle * wsz;
				const char *r0 = (const char *)r, *r1 = r0 + (size_t)args->num_particles * rsz;
				return w0 < r1 && r0 < w1;
			};
			// with caller-provided records the pair loop gathers from THEM, and pos / vel may be integrated in place
			const bool own_records = args->packed != NULL;
			if ((!own_records && (overlaps(eul->new_pos, 16, args->pos, 16) || overlaps(eul->new_pos, 16, args->vel, 16) ||
					overlaps(eul->new_vel, 16, args->pos, 16) || overlaps(eul->new_vel, 16, args->vel, 16))) ||
				overlaps(eul->new_pos, 16, args->packed, 32) || overlaps(eul->new_vel, 16, args->packed, 32) ||
				overlaps(eul->new_packed, 32, args->packed, 32) || overlaps(eul->new_packed, 32, args->pos, 16) ||
				overlaps(eul->new_packed, 32, args->vel, 16)) {
				b200_set_error("forces_euler: the integrated state must not overwrite the state the pair loop reads"); return B200SPH_EINVAL; }
			if ((uintptr_t)eul->new_packed & 31u) { b200_set_error("forces_euler: new_packed must be 32-byte aligned"); return B200SPH_EINVAL; }
		}
	}
	if (rb_forces) {
		if (!rb_torques) { b200_set_error("forces: rb_forces without rb_torques"); return B200SPH_EINVAL; }
		if ((ctx->bodies_set & (BODY_SET_CG_FORCES | BODY_SET_START)) != (BODY_SET_CG_FORCES | BODY_SET_START)) {
			b200_set_error("forces: body output requested before setrbcg/setrbstart"); return B200SPH_EINVAL; }
		bo.bodies = ctx->d_bodies;
	}
	const bool xsph = (ctx->hp.simflags & B200SPH_ENABLE_XSPH) != 0;
	if (xsph && !args->xsph && to > from) { b200_set_error("forces: ENABLE_XSPH needs the xsph buffer"); return B200SPH_EINVAL; }
	if (xsph) bo.xsph = (float4 *)args->xsph;
	const bool brezzi = ctx->dp.densitydiffusiontype == B200SPH_RHODIFF_BREZZI;
	if (brezzi && !args->dt_from_device && !(args->dt > 0) && to > from) { b200_set_error("forces: BREZZI density diffusion needs the command's dt"); return B200SPH_EINVAL; }
	if (brezzi && args->dt_from_device && args->step != 1 && args->step != 2) { b200_set_error("forces: bad integrator step %d", args->step); return B200SPH_EINVAL; }
	if (num_cfl_blocks) *num_cfl_blocks = 0;
	if (to <= from) return B200SPH_OK;
	if (!pos || !vel || !info || !hash || !cell_start || !neibs_list || !forces) { b200_set_error("forces: null buffer"); return B200SPH_EINVAL; }
	if (to > num_particles) { b200_set_error("forces: range end beyond numParticles"); return B200SPH_EINVAL; }
	if ((uintptr_t)args->packed & 31u) { b200_set_error("forces: packed records must be 32-byte aligned"); return B200SPH_EINVAL; }
	// CFL blocks: grid of the reference rounded to a multiple of 4 (forces.cu:741-744) so that the array can be
	// reduced as float4
	uint nblocks = div_up(to - from, BLOCK_FORCES);
	nblocks = (nblocks + 3) / 4 * 4;
	const DevParams &d = ctx->dp;
	const bool artvisc = d.turbmodel == B200SPH_TURB_ARTIFICIAL, laminar = !d.inviscid, multi = d.numFluids > 1;
	// options served by the general (run-time switched) variant of the kernel only
	const bool general = brezzi || xsph || (laminar && d.viscmodel != B200SPH_VISCMODEL_MORRIS);
	DevParams dp_launch = ctx->dp;
	dp_launch.cmd_dt = args->dt; dp_launch.cmd_step = args->step;
	dp_launch.dev_state = args->dt_from_device ? ctx->d_step : NULL;
	// neighbour records: the caller's, or an interleaved copy of pos / vel made now (every neighbour of [from, to) can
	// be anywhere in [0, num_particles))
	const PosVel *pv = (const PosVel *)args->packed;
	if (!pv) {
		PosVel *scratch = NULL;
		{ const int rc = b200_packed_scratch(ctx, 0, num_particles, &scratch); if (rc) return rc; }
		pack_state_kernel<<<div_up(num_particles, BLOCK_STREAM), BLOCK_STREAM, 0, ctx->stream>>>((const float4 *)pos, (const float4 *)vel,
			scratch, 0, num_particles);
		KERNEL_TRY();
		pv = scratch;
	}
	// integration: fused into the kernel's epilogue, or a second launch with XSPH (whose mean velocity the integration
	// reads from the buffer)
	const bool fuse = eul && !xsph;
	if (fuse) {
		bo.eul_step = eul->step; bo.eul_dt = eul->dt; bo.eul_state = eul->dt_from_device ? ctx->d_step : NULL;
		const bool old_is_current = eul->old_pos == args->pos && eul->old_vel == args->vel;
		bo.eul_old_pos = old_is_current ? NULL : (const float4 *)eul->old_pos;
		bo.eul_old_vel = old_is_current ? NULL : (const float4 *)eul->old_vel;
		bo.eul_new_pos = (float4 *)eul->new_pos; bo.eul_new_vel = (float4 *)eul->new_vel;
		bo.eul_new_packed = (PosVel *)eul->new_packed;
		{ const int rc = b200_euler_bodies(ctx, hash, &bo.eul_bodies); if (rc) return rc; }
	}
	// 32-bit list offsets unless the list has 2^31 entries or more
	const bool wide = ((unsigned long long)d.neiblistsize + 1) * d.stride + num_particles >= 0x7fffffffull;
	gather_kernel_t gks[2];
	if (general) {
		gks[0] = forces_gather_kernel<RHODIFF_RUNTIME, true, true, true, false>;
		gks[1] = forces_gather_kernel<RHODIFF_RUNTIME, true, true, true, true>;
	} else switch (d.densitydiffusiontype) {
	case B200SPH_RHODIFF_FERRARI: pick_kernels<B200SPH_RHODIFF_FERRARI>(artvisc, laminar, multi, gks); break;
	case B200SPH_RHODIFF_COLAGROSSI: pick_kernels<B200SPH_RHODIFF_COLAGROSSI>(artvisc, laminar, multi, gks); break;
	default: pick_kernels<B200SPH_RHODIFF_NONE>(artvisc, laminar, multi, gks); break;
	}
	gather_kernel_t gk = gks[wide ? 1 : 0];
	{ const int rc = set_carveout(ctx, (const void *)gk); if (rc) return rc; }
	gk<<<nblocks, BLOCK_FORCES, 0, ctx->stream>>>(general ? dp_launch : ctx->dp, pv, (const ushort4 *)info, hash, cell_start, neibs_list,
		(float4 *)forces, cfl, bo, from, to, cfl_offset);
	KERNEL_TRY();
	if (num_cfl_blocks) *num_cfl_blocks = nblocks;
	if (eul && !fuse) {
		const size_t o = (size_t)from * 16;
		const int rc = b200sph_euler_ex(ctx, (const char *)eul->old_pos + o, (const char *)eul->old_vel + o, (const char *)info + (size_t)from * 8,
			hash + from, (const char *)forces + o, args->xsph ? (const char *)args->xsph + o : NULL, (char *)eul->new_pos + o,
			(char *)eul->new_vel + o, to - from, to - from, eul->dt, eul->step, eul->dt_from_device);
		if (rc) return rc;
		if (eul->new_packed) return b200sph_pack_state(ctx, eul->new_pos, eul->new_vel, eul->new_packed, from, to);
	}
	return B200SPH_OK;
}

// ---------------------------------------------------------------------------
// reduceRbForces — reference src/cuda/forces.cu:967-1003 (thrust::inclusive_scan_by_key in place, then the last
// element of every body's segment is read back)
// ---------------------------------------------------------------------------
struct Float4Plus { __host__ __device__ float4 operator()(const float4 &a, const float4 &b) const
	{ return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); } };

extern "C" int b200sph_reduce_rb_forces(b200sph_ctx *ctx, void *rb_forces, void *rb_torques, const uint32_t *rb_keys,
	const uint32_t *lastindex, float *total_force, float *total_torque, uint32_t nb, uint32_t np)
{
	CHECK_CTX(ctx);
	if (nb == 0 || np == 0) return B200SPH_OK;
	if (!rb_forces || !rb_torques || !rb_keys || !lastindex || !total_force || !total_torque) { b200_set_error("reduceRbForces: null buffer"); return B200SPH_EINVAL; }
	if (nb > B200SPH_MAX_BODIES) { b200_set_error("reduceRbForces: too many bodies"); return B200SPH_EINVAL; }
	float4 *arrays[2] = { (float4 *)rb_forces, (float4 *)rb_torques };
	for (int a = 0; a < 2; ++a) {
		size_t tmp = 0;
		CUDA_TRY(cub::DeviceScan::InclusiveScanByKey(NULL, tmp, rb_keys, arrays[a], arrays[a], Float4Plus(), (int)np, cub::Equality(), ctx->stream));
		if (ctx->sort_tmp_bytes < tmp) {
			cudaFree(ctx->sort_tmp); ctx->sort_tmp = NULL; ctx->sort_tmp_bytes = 0;
			CUDA_TRY(cudaMalloc(&ctx->sort_tmp, tmp));
			ctx->sort_tmp_bytes = tmp;
		}
		tmp = ctx->sort_tmp_bytes;
		CUDA_TRY(cub::DeviceScan::InclusiveScanByKey(ctx->sort_tmp, tmp, rb_keys, arrays[a], arrays[a], Float4Plus(), (int)np, cub::Equality(), ctx->stream));
	}
	float4 host[2 * B200SPH_MAX_BODIES];
	for (uint32_t b = 0; b < nb; ++b) {
		CUDA_TRY(cudaMemcpyAsync(&host[2 * b], arrays[0] + lastindex[b], sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
		CUDA_TRY(cudaMemcpyAsync(&host[2 * b + 1], arrays[1] + lastindex[b], sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
	}
	CUDA_TRY(cudaStreamSynchronize(ctx->stream));
	for (uint32_t b = 0; b < nb; ++b) {
		total_force[3 * b] = host[2 * b].x; total_force[3 * b + 1] = host[2 * b].y; total_force[3 * b + 2] = host[2 * b].z;
		total_torque[3 * b] = host[2 * b + 1].x; total_torque[3 * b + 1] = host[2 * b + 1].y; total_torque[3 * b + 2] = host[2 * b + 1].z;
	}
	return B200SPH_OK;
}

// ---------------------------------------------------------------------------
// dtreduce — reference src/cuda/forces.cu:557-607 (+ cflmax :153-177, fmaxDevice
// forces_kernel.cu:729-795). Single launch, result through pinned memory.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
max_reduce_kernel(const float *__restrict__ in, const uint n, float *__restrict__ out)
{
	float m = 0.0f;   // CFL terms are non-negative
	for (uint i = threadIdx.x; i < n; i += blockDim.x) m = fmaxf(m, in[i]);
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
	__shared__ float s[32];
	if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = m;
	__syncthreads();
	if (threadIdx.x < 32) {
		m = threadIdx.x < (blockDim.x >> 5) ? s[threadIdx.x] : 0.0f;
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
		if (threadIdx.x == 0) *out = m;
	}
}

static float dt_from_cfl(const b200sph_params &hp, const float maxcfl)
{
	float dt = hp.dtadaptfactor * fminf(sqrtf(hp.slength / maxcfl), hp.slength / hp.max_sound_speed_cfl);
	if (hp.rheologytype != B200SPH_RHEOLOGY_INVISCID || hp.turbmodel > B200SPH_TURB_ARTIFICIAL) {
		float dt_visc = hp.slength * hp.slength / hp.max_kinvisc;
		dt_visc *= 0.125f;
		if (dt_visc < dt) dt = dt_visc;
	}
	return dt;
}

extern "C" int b200sph_cflmax(b200sph_ctx *ctx, const float *cfl, uint32_t num_blocks, float *max_out)
{
	CHECK_CTX(ctx);
	if (!cfl || !max_out) { b200_set_error("cflmax: null buffer"); return B200SPH_EINVAL; }
	max_reduce_kernel<<<1, 1024, 0, ctx->stream>>>(cfl, num_blocks, max_out);
	KERNEL_TRY();
	return B200SPH_OK;
}

extern "C" int b200sph_dt_from_cfl(const b200sph_ctx *ctx, float max_cfl, float *dt_out)
{
	if (!ctx || !dt_out) { b200_set_error("dt_from_cfl: null argument"); return B200SPH_EINVAL; }
	*dt_out = dt_from_cfl(ctx->hp, max_cfl);
	return B200SPH_OK;
}

static int dtreduce_impl(b200sph_ctx *ctx, const float *cfl, uint32_t num_blocks, const b200sph_params &hp, float *dt_out)
{
	if (!cfl || !dt_out) { b200_set_error("dtreduce: null buffer"); return B200SPH_EINVAL; }
	max_reduce_kernel<<<1, 1024, 0, ctx->stream>>>(cfl, num_blocks, ctx->d_scalar);
	KERNEL_TRY();
	CUDA_TRY(cudaMemcpyAsync(ctx->h_scalar, ctx->d_scalar, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
	CUDA_TRY(cudaStreamSynchronize(ctx->stream));
	*dt_out = dt_from_cfl(hp, ctx->h_scalar[0]);
	return B200SPH_OK;
}

extern "C" int b200sph_dtreduce(b200sph_ctx *ctx, const float *cfl, float *temp_cfl, uint32_t num_blocks, float *dt_out)
{
	CHECK_CTX(ctx);
	(void)temp_cfl;
	return dtreduce_impl(ctx, cfl, num_blocks, ctx->hp, dt_out);
}

extern "C" int b200sph_dtreduce_ex(b200sph_ctx *ctx, const float *cfl, float *temp_cfl, uint32_t num_blocks,
	float slength, float dtadaptfactor, float sspeed_cfl, float max_kinematic, float *dt_out)
{
	CHECK_CTX(ctx);
	(void)temp_cfl;
	b200sph_params hp = ctx->hp;
	hp.slength = slength; hp.dtadaptfactor = dtadaptfactor; hp.max_sound_speed_cfl = sspeed_cfl; hp.max_kinvisc = max_kinematic;
	return dtreduce_impl(ctx, cfl, num_blocks, hp, dt_out);
}

// ---------------------------------------------------------------------------
// device-resident dt (b200sph_step_* in include/b200sph.h)
// ---------------------------------------------------------------------------
struct DtConsts { float dtadaptfactor, slength, max_ss_cfl, max_kinvisc; int viscous; };

__device__ __forceinline__ float dt_from_cfl_dev(const DtConsts c, const float maxcfl)
{	// same arithmetic as dt_from_cfl() above (src/cuda/forces.cu:571-600)
	float dt = c.dtadaptfactor * fminf(sqrtf(c.slength / maxcfl), c.slength / c.max_ss_cfl);
	if (c.viscous) {
		float dt_visc = c.slength * c.slength / c.max_kinvisc;
		dt_visc *= 0.125f;
		if (dt_visc < dt) dt = dt_visc;
	}
	return dt;
}

__global__ void __launch_bounds__(1024)
dtreduce_async_kernel(const float *__restrict__ in, const uint n, const DtConsts c, StepState *__restrict__ st, const int which)
{
	float m = 0.0f;
	for (uint i = threadIdx.x; i < n; i += blockDim.x) m = fmaxf(m, in[i]);
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
	__shared__ float s[32];
	if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = m;
	__syncthreads();
	if (threadIdx.x < 32) {
		m = threadIdx.x < (blockDim.x >> 5) ? s[threadIdx.x] : 0.0f;
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
		if (threadIdx.x == 0) { const float dt = dt_from_cfl_dev(c, m); if (which == 1) st->dt1 = dt; else st->dt2 = dt; }
	}
}

__global__ void step_end_kernel(StepState *st)
{
	st->t += (double)st->dt;
	st->iterations += 1;
	st->dt = fminf(st->dt1, st->dt2);
}

__global__ void step_set_dt_kernel(StepState *st, const float dt) { st->dt = dt; st->dt1 = dt; st->dt2 = dt; }

extern "C" int b200sph_step_set_dt(b200sph_ctx *ctx, float dt)
{
	CHECK_CTX(ctx);
	step_set_dt_kernel<<<1, 1, 0, ctx->stream>>>(ctx->d_step, dt);
	KERNEL_TRY();
	return B200SPH_OK;
}

extern "C" int b200sph_dtreduce_async(b200sph_ctx *ctx, const float *cfl, uint32_t num_blocks, int which)
{
	CHECK_CTX(ctx);
	if (!cfl || (which != 1 && which != 2)) { b200_set_error("dtreduce_async: bad argument"); return B200SPH_EINVAL; }
	const b200sph_params &hp = ctx->hp;
	DtConsts c;
	c.dtadaptfactor = hp.dtadaptfactor; c.slength = hp.slength; c.max_ss_cfl = hp.max_sound_speed_cfl; c.max_kinvisc = hp.max_kinvisc;
	c.viscous = (hp.rheologytype != B200SPH_RHEOLOGY_INVISCID || hp.turbmodel > B200SPH_TURB_ARTIFICIAL) ? 1 : 0;
	dtreduce_async_kernel<<<1, 1024, 0, ctx->stream>>>(cfl, num_blocks, c, ctx->d_step, which);
	KERNEL_TRY();
	return B200SPH_OK;
}

extern "C" int b200sph_step_end(b200sph_ctx *ctx)
{
	CHECK_CTX(ctx);
	step_end_kernel<<<1, 1, 0, ctx->stream>>>(ctx->d_step);
	KERNEL_TRY();
	return B200SPH_OK;
}

extern "C" int b200sph_step_query(b200sph_ctx *ctx, double *t, float *dt, uint64_t *iterations)
{
	CHECK_CTX(ctx);
	CUDA_TRY(cudaMemcpyAsync(ctx->h_step, ctx->d_step, sizeof(StepState), cudaMemcpyDeviceToHost, ctx->stream));
	CUDA_TRY(cudaStreamSynchronize(ctx->stream));
	if (t) *t = ctx->h_step->t;
	if (dt) *dt = ctx->h_step->dt;
	if (iterations) *iterations = ctx->h_step->iterations;
	return B200SPH_OK;
}
