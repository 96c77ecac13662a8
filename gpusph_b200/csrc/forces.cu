// forces.cu — forces engine: fused pair-interaction kernel, CFL reduction, dt.
//
// Behavioural specification: GPUSPH src/cuda/forces.cu + forces_kernel.def (cited inline).
// The reference evaluates one half-step with FOUR launches (forcesDevice<fluid,fluid>,
// <fluid,boundary>, <boundary,fluid>, finalizeforcesDevice), each re-reading the particle and
// read-modify-writing forces[] in global memory, and re-evaluates the equation of state
// (two __powf = 4 MUFU + an IEEE division) for BOTH particles of every pair.
// Here:
//  * one pre-pass evaluates P/rho^2 and the sound speed ONCE per particle (same __powf
//    expressions, so the values are the ones the reference recomputes per pair);
//  * one launch walks both neighbour-list sections of a particle, keeps the accumulator in
//    registers, applies the finalize step (1/rho0, gravity) and reduces the CFL term with
//    warp shuffles — forces[] is written exactly once.
// Summation order inside each list section is the reference's (list order); the fluid and
// boundary partial sums are combined as  (0 + sum_fluid) + sum_boundary  like the reference's
// RMW sequence, so results differ from the reference only through FMA contraction choices.
#include "common.cuh"

// ---- equation of state, reference src/cuda/phys_core.cu:99-151 ----
__device__ __forceinline__ float eos_pressure(const DevParams &P, float rho_tilde, int f)
{
	const float rho_ratio = rho_tilde + 1.0f;
	return P.bcoeff[f] * (__powf(rho_ratio, P.gammacoeff[f]) - 1.0f);
}
__device__ __forceinline__ float eos_sound_speed(const DevParams &P, float rho_tilde, int f)
{
	const float rho_ratio = rho_tilde + 1.0f;
	return P.sscoeff[f] * __powf(rho_ratio, P.sspowercoeff[f]);
}
__device__ __forceinline__ float phys_density(const DevParams &P, float rho_tilde, int f)
{
	return (rho_tilde + 1.0f) * P.rho0[f];
}

// per particle: x = P/rho^2 (precalc_pressure<SPH_F1>, forces_kernel.def:419-429), y = sound speed
__global__ void __launch_bounds__(BLOCK_STREAM)
eos_kernel(const __grid_constant__ DevParams P, const float4 *__restrict__ vel, const ushort4 *__restrict__ info,
	float2 *__restrict__ eos, const uint n)
{
	const uint i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const float rho_tilde = vel[i].w;
	const int f = fluid_num_of(info[i]);
	const float rho = phys_density(P, rho_tilde, f);
	float2 e;
	e.x = eos_pressure(P, rho_tilde, f) / (rho * rho);
	e.y = eos_sound_speed(P, rho_tilde, f);
	eos[i] = e;
}

// Packed neighbour record, 48 bytes = 3 x float4, written once per force evaluation:
//   [0] pos.xyz (cell-local), mass          [1] vel.xyz, rho~
//   [2] P/rho^2, sound speed, physical density, fluid number (as int bits)
// One base address + three 128-bit loads per neighbour instead of three separately addressed gathers,
// and the neighbour's EOS terms / density come precomputed (the reference re-evaluates __powf twice per pair).
__global__ void __launch_bounds__(BLOCK_STREAM)
pack_kernel(const __grid_constant__ DevParams P, const float4 *__restrict__ pos, const float4 *__restrict__ vel,
	const ushort4 *__restrict__ info, float4 *__restrict__ rec, const uint n)
{
	const uint i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const float4 v = vel[i];
	const int f = fluid_num_of(info[i]);
	const float rho = phys_density(P, v.w, f);
	float4 e;
	e.x = eos_pressure(P, v.w, f) / (rho * rho);
	e.y = eos_sound_speed(P, v.w, f);
	e.z = rho;
	e.w = __int_as_float(f);
	rec[3 * (size_t)i + 0] = pos[i];
	rec[3 * (size_t)i + 1] = v;
	rec[3 * (size_t)i + 2] = e;
}

int b200_eos_precompute(b200sph_ctx *ctx, const float4 *pos, const float4 *vel, const ushort4 *info, uint n)
{
	if (ctx->eos_cap < n) {
		cudaFree(ctx->eos); ctx->eos = NULL; ctx->eos_cap = 0;
		const size_t cap = (size_t)n + (n >> 3) + 1024;
		CUDA_TRY(cudaMalloc(&ctx->eos, cap * 3 * sizeof(float4)));
		ctx->eos_cap = cap;
	}
	pack_kernel<<<div_up(n, BLOCK_STREAM), BLOCK_STREAM, 0, ctx->stream>>>(ctx->dp, pos, vel, info, (float4 *)ctx->eos, n);
	KERNEL_TRY();
	return B200SPH_OK;
}

extern "C" int b200sph_eos_probe(b200sph_ctx *ctx, const void *vel, const void *info, void *out, uint32_t n)
{
	CHECK_CTX(ctx);
	if (n == 0) return B200SPH_OK;
	if (!vel || !info || !out) { b200_set_error("eos_probe: null buffer"); return B200SPH_EINVAL; }
	eos_kernel<<<div_up(n, BLOCK_STREAM), BLOCK_STREAM, 0, ctx->stream>>>(ctx->dp, (const float4 *)vel, (const ushort4 *)info, (float2 *)out, n);
	KERNEL_TRY();
	return B200SPH_OK;
}

// density-only viscous averaging, reference src/cuda/visc_avg.cu
__device__ __forceinline__ float visc_avg_density(const DevParams &P, float rho, float nrho, float nmass)
{
	switch (P.viscavgop) {
	case B200SPH_AVG_ARITHMETIC: return nmass * (rho + nrho) / (rho * nrho);
	case B200SPH_AVG_HARMONIC: return 4 * nmass / (rho + nrho);
	default: return 2 * nmass * rsqrtf(rho * nrho);
	}
}
__device__ __forceinline__ float visc_avg_dyn(const DevParams &P, float v, float nv, float rho, float nrho, float nmass)
{
	switch (P.viscavgop) {
	case B200SPH_AVG_ARITHMETIC: return nmass * (v + nv) / (rho * nrho);
	case B200SPH_AVG_HARMONIC: return 4 * nmass * (v * nv) / (v + nv) / (rho * nrho);
	default: return 2 * nmass * sqrtf(v * nv) / (rho * nrho);
	}
}

// ---------------------------------------------------------------------------
// The pair kernel. It is instruction-issue bound (ncu: profiles/forces_r01_*.txt), so everything here
// is about instructions per pair:
//  * physics options are TEMPLATE parameters (the reference does the same through its SFINAE
//    specialisations, forces_kernel.def:1560-2770): no run-time branches, no loads of unused constants;
//  * per-thread constants live in registers; the only per-pair divisions/roots are MUFU approximations
//    (rsqrt, rcp) — the neighbour-list builder, not this kernel, owns the bit-exact distance test;
//  * the 27 neighbour-cell base indices are staged in shared memory once per particle (the reference's
//    getNeibIndex re-reads cellStart from global memory at every cell change, cellgrid.cuh:198-226);
//  * the list column is read one row ahead (software prefetch), the three gathers of a pair are issued
//    together.
// Accumulation order is the list order, exactly as in the reference (neibs_iteration.cuh:56-200).
// ---------------------------------------------------------------------------
struct PairConsts {
	float inv_h, fc, R2;       // 1/h, Wendland gradient coefficient, squared influence radius
	float h_alpha, eps;        // artificial viscosity: h*alpha, eps
	float g0, g1, g2;          // gravity
	float diff;                // density diffusion coefficient (Colagrossi single-fluid: xi*2h*c0)
	float grav_scale;          // Ferrari: rho0/c0^2
	float h;
};

__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rsqrt_approx(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// 16-bit neighbour-list entry, zero-extended, through the read-only path
__device__ __forceinline__ uint ld_neib(const ushort *p) { uint v; asm volatile("ld.global.nc.u16 %0, [%1];" : "=r"(v) : "l"(p)); return v; }

template<bool NFLUID, int RHODIFF, bool ARTVISC, bool LAMINAR, bool MULTIFLUID>
__device__ __forceinline__ void
walk_section(const DevParams &P, const PairConsts &k, const uint index, const float4 pos, const float4 vel,
	const float rho, const float p_precalc, const float sspeed, const int fnum, const bool momentum,
	const uint *s_base /* [27][BLOCK_FORCES] + tid */, const float4 *s_off /* [27] cell offset * cell size */,
	const float4 *__restrict__ rec, const ushort *__restrict__ neibsList, float4 &acc)
{
	const size_t stride = P.stride;
	// list column of this particle; fluid section grows up from row 0, boundary section down from neibboundpos
	const ushort *row = neibsList + index + (NFLUID ? (size_t)0 : (size_t)P.neibboundpos * stride);
	const ptrdiff_t rstep = NFLUID ? (ptrdiff_t)stride : -(ptrdiff_t)stride;
	uint base = 0;
	float pcx = 0.f, pcy = 0.f, pcz = 0.f;
	uint nd = ld_neib(row);
	while (nd != NEIBS_END) {
		// prefetch the next row: always in bounds, the section is terminated by NEIBS_END before the list ends
		// (buildneibs_kernel.cu:1108-1137)
		row += rstep;
		const uint nd_next = ld_neib(row);
		if (nd >= CELLNUM_ENCODED) {                                    // getNeibIndex, cellgrid.cuh:198-226
			const uint cell = (nd >> CELLNUM_SHIFT) - 1;
			nd &= NEIBINDEX_MASK;
			base = s_base[cell * BLOCK_FORCES];
			const float4 o = s_off[cell];                               // pos_corr = pos - offset*cellSize (:215)
			pcx = pos.x - o.x; pcy = pos.y - o.y; pcz = pos.z - o.z;
		}
		const float4 *nr = rec + 3 * (size_t)(base + nd);
		const float4 np = __ldg(nr);
		const float4 nv = __ldg(nr + 1);
		const float4 ne = __ldg(nr + 2);
		nd = nd_next;
		const float rx = pcx - np.x, ry = pcy - np.y, rz = pcz - np.z;
		const float nmass = np.w;
		const float r2 = fmaf(rz, rz, fmaf(ry, ry, rx * rx));
		// skip inactive neighbours and pairs beyond the kernel support (forces_kernel.def:3987-3999)
		if (!(r2 < k.R2) || !(fabsf(nmass) < __int_as_float(0x7f800000))) continue;
		const int nfnum = MULTIFLUID ? __float_as_int(ne.w) : 0;
		const float r = r2 * rsqrt_approx(r2 + 1e-30f);
		// common_neib_data :1099-1130
		const float rvx = vel.x - nv.x, rvy = vel.y - nv.y, rvz = vel.z - nv.z;
		const float vel_dot_pos = fmaf(rvz, rz, fmaf(rvy, ry, rvx * rx));
		const float qm2 = fmaf(r, k.inv_h, -2.0f);                       // F<WENDLAND>, sph_core.cu:168-174
		const float f = qm2 * qm2 * qm2 * k.fc;
		const float mf = nmass * f;
		const float np_precalc = ne.x, nsspeed = ne.y, nrho = ne.z;

		// --- continuity: mass_continuity_div_vel_term :2140-2150 ---
		float DrDt = mf * vel_dot_pos;
		if (NFLUID) {   // no density diffusion from DYN boundary neighbours, :1594-1606
			if (RHODIFF == B200SPH_RHODIFF_FERRARI) {                   // :1614-1636
				const float gdot = fmaf(k.g2, rz, fmaf(k.g1, ry, k.g0 * rx));
				const float grav_corr = -gdot * (MULTIFLUID ? P.rho0[fnum] / P.sqC0[fnum] : k.grav_scale);
				// ferraricor . relPos = max(c) (rho - rho_j + corr)/rho / r * r^2   (zero for r <= 1e-4 h)
				const float s = (r > 1e-4f * k.h) ? fmaxf(sspeed, nsspeed) * (rho - nrho + grav_corr) * rcp_approx(rho) * r : 0.0f;
				DrDt = fmaf(k.diff * mf, s, DrDt);
			} else if (RHODIFF == B200SPH_RHODIFF_COLAGROSSI) {         // :1916-1951
				if (!MULTIFLUID || fnum == nfnum) {
					const float Pi = p_precalc * (rho * rho), Pj = np_precalc * (nrho * nrho);
					const float gdot = fmaf(k.g2, rz, fmaf(k.g1, ry, k.g0 * rx));
					if (!(fabsf(Pi - Pj) < fabsf(gdot * rho)))
						DrDt -= k.diff * (MULTIFLUID ? P.sscoeff[fnum] : 1.0f) * (nrho * rcp_approx(rho) - 1.0f) * mf;
				}
			}
		}
		acc.w += DrDt;                                                  // :2189

		if (momentum) {
			// compute_pressure_contrib, general formulation :2450-2466:  -(P_i/rho_i^2 + P_j/rho_j^2) m_j F r_ij
			float coef = -(p_precalc + np_precalc) * mf;
			// artificial viscosity :2744-2764, artvisc visc_kernel.cu:75-85
			if (ARTVISC) {
				const float visc = vel_dot_pos * k.h_alpha * (sspeed + nsspeed) * rcp_approx((r2 + k.eps) * (rho + nrho));
				coef = (vel_dot_pos < 0.0f) ? fmaf(visc, mf, coef) : coef;
			}
			float dvx = coef * rx, dvy = coef * ry, dvz = coef * rz;
			// laminar (Morris) :2605-2625
			if (LAMINAR) {
				const float vc = P.visccoeff[fnum], nvc = P.visccoeff[nfnum];
				float visc;
				if (P.compvisc == B200SPH_COMPVISC_KINEMATIC)
					visc = P.is_const_visc ? vc * visc_avg_density(P, rho, nrho, nmass) : visc_avg_dyn(P, vc * rho, nvc * nrho, rho, nrho, nmass);
				else
					visc = P.is_const_visc ? 2 * nmass * vc / (rho * nrho) : visc_avg_dyn(P, vc, nvc, rho, nrho, nmass);
				const float s = visc * f;
				dvx = fmaf(s, rvx, dvx); dvy = fmaf(s, rvy, dvy); dvz = fmaf(s, rvz, dvz);
			}
			acc.x += dvx; acc.y += dvy; acc.z += dvz;                   // :3590
		}
	}
}

template<int RHODIFF, bool ARTVISC, bool LAMINAR, bool MULTIFLUID>
__global__ void __launch_bounds__(BLOCK_FORCES)
forces_kernel(const __grid_constant__ DevParams P, const ushort4 *__restrict__ infoArray,
	const uint *__restrict__ particleHash, const float4 *__restrict__ rec,
	const uint *__restrict__ cellStart, const ushort *__restrict__ neibsList,
	float4 *__restrict__ forces, float *__restrict__ cfl,
	const uint fromParticle, const uint toParticle, const uint cflOffset)
{
	__shared__ uint s_cellbase[27 * BLOCK_FORCES];
	__shared__ float4 s_celloff[27];
	const uint index = blockIdx.x * blockDim.x + threadIdx.x + fromParticle;
	float cfl_term = 0.0f;
	if (threadIdx.x < 27) {
		const int c = threadIdx.x;
		s_celloff[c] = make_float4((float)(c % 3 - 1) * P.cellSize[0], (float)((c / 3) % 3 - 1) * P.cellSize[1],
			(float)(c / 9 - 1) * P.cellSize[2], 0.f);
	}
	__syncthreads();

	if (index < toParticle) {
		const ushort4 info = infoArray[index];
		const int type = ptype_of(info);
		const float4 pos = rec[3 * (size_t)index];
		if ((type == PT_FLUID || type == PT_BOUNDARY) && fabsf(pos.w) < __int_as_float(0x7f800000)) {
			const float4 vel = rec[3 * (size_t)index + 1];
			const float4 e = rec[3 * (size_t)index + 2];
			const int fnum = MULTIFLUID ? __float_as_int(e.w) : 0;
			const float rho = e.z;
			PairConsts k;
			k.h = P.slength; k.inv_h = 1.0f / P.slength; k.fc = P.fcoeff_wendland;
			k.R2 = P.influenceradius * P.influenceradius;
			k.h_alpha = P.slength * P.artvisccoeff; k.eps = P.epsartvisc;
			k.g0 = P.gravity[0]; k.g1 = P.gravity[1]; k.g2 = P.gravity[2];
			k.diff = RHODIFF == B200SPH_RHODIFF_COLAGROSSI && !MULTIFLUID ? P.densityDiffCoeff * P.sscoeff[0] : P.densityDiffCoeff;
			k.grav_scale = P.rho0[0] / P.sqC0[0];
			// first particle of each of the 27 neighbouring cells -> shared memory (27 independent loads);
			// calcGridHashPeriodic, cellgrid.cuh:174-185 (cells outside a non-periodic domain are never listed)
			uint *my_base = s_cellbase + threadIdx.x;
			{
				const int h0 = (int)(particleHash[index] & CELLTYPE_BITMASK);
				const int3 gp = grid_pos(P, (uint)h0);
				const int sx = P.hstride[0], sy = P.hstride[1], sz = P.hstride[2];
				const int Gx = P.gridSize[0], Gy = P.gridSize[1], Gz = P.gridSize[2];
				// hash delta of a step of -1 / 0 / +1 cells along each axis, wrapped at the domain faces
				const int dx[3] = { gp.x == 0 ? (Gx - 1) * sx : -sx, 0, gp.x == Gx - 1 ? -(Gx - 1) * sx : sx };
				const int dy[3] = { gp.y == 0 ? (Gy - 1) * sy : -sy, 0, gp.y == Gy - 1 ? -(Gy - 1) * sy : sy };
				const int dz[3] = { gp.z == 0 ? (Gz - 1) * sz : -sz, 0, gp.z == Gz - 1 ? -(Gz - 1) * sz : sz };
#pragma unroll
				for (int cell = 0; cell < 27; ++cell)
					my_base[cell * BLOCK_FORCES] = __ldg(cellStart + (h0 + dx[cell % 3] + dy[(cell / 3) % 3] + dz[cell / 9]));
			}
			float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
			if (type == PT_FLUID) {
				// forcesDevice<fluid,fluid> then <fluid,boundary> (forces.cu:759,782); DYN boundary neighbours
				// interact like fluid ones (forces_kernel.def:3717-3726)
				walk_section<true, RHODIFF, ARTVISC, LAMINAR, MULTIFLUID>(P, k, index, pos, vel, rho, e.x, e.y, fnum, true,
					my_base, s_celloff, rec, neibsList, acc);
				walk_section<false, RHODIFF, ARTVISC, LAMINAR, MULTIFLUID>(P, k, index, pos, vel, rho, e.x, e.y, fnum, true,
					my_base, s_celloff, rec, neibsList, acc);
			} else {
				// forcesDevice<boundary,fluid> (forces.cu:792): density always, momentum only with force feedback
				// (forces_kernel.def:3634-3667)
				walk_section<true, RHODIFF, ARTVISC, LAMINAR, MULTIFLUID>(P, k, index, pos, vel, rho, e.x, e.y, fnum,
					(info.x & B200SPH_FG_COMPUTE_FORCE) != 0, my_base, s_celloff, rec, neibsList, acc);
			}
			// finalizeforcesDevice :4037-4153
			acc.w /= P.rho0[fnum];                                          // forces_fixup :3212-3219
			if (type == PT_FLUID) {
				acc.x += P.gravity[0]; acc.y += P.gravity[1]; acc.z += P.gravity[2];   // :4091
				// dyndt_forces_shared_data::store :3436-3456
				cfl_term = fmaxf(sqrtf(acc.x * acc.x + acc.y * acc.y + acc.z * acc.z), e.y * e.y / P.slength);
			}
			forces[index] = acc;
		}
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

typedef void (*forces_kernel_t)(const DevParams, const ushort4 *, const uint *, const float4 *,
	const uint *, const ushort *, float4 *, float *, const uint, const uint, const uint);

template<int RHODIFF>
static forces_kernel_t pick_forces_kernel(bool artvisc, bool laminar, bool multifluid)
{
	if (multifluid) {
		if (artvisc) return laminar ? forces_kernel<RHODIFF, true, true, true> : forces_kernel<RHODIFF, true, false, true>;
		return laminar ? forces_kernel<RHODIFF, false, true, true> : forces_kernel<RHODIFF, false, false, true>;
	}
	if (artvisc) return laminar ? forces_kernel<RHODIFF, true, true, false> : forces_kernel<RHODIFF, true, false, false>;
	return laminar ? forces_kernel<RHODIFF, false, true, false> : forces_kernel<RHODIFF, false, false, false>;
}

extern "C" int b200sph_forces(b200sph_ctx *ctx, const void *pos, const void *vel, const void *info,
	const uint32_t *hash, const uint32_t *cell_start, const uint16_t *neibs_list,
	void *forces, float *cfl, uint32_t num_particles, uint32_t from, uint32_t to,
	uint32_t cfl_offset, uint32_t *num_cfl_blocks)
{
	CHECK_CTX(ctx);
	if (num_cfl_blocks) *num_cfl_blocks = 0;
	if (to <= from) return B200SPH_OK;
	if (!pos || !vel || !info || !hash || !cell_start || !neibs_list || !forces) { b200_set_error("forces: null buffer"); return B200SPH_EINVAL; }
	if (to > num_particles) { b200_set_error("forces: range end beyond numParticles"); return B200SPH_EINVAL; }
	int rc = b200_eos_precompute(ctx, (const float4 *)pos, (const float4 *)vel, (const ushort4 *)info, num_particles);
	if (rc) return rc;
	// grid rounded to a multiple of 4 blocks like the reference (forces.cu:741-744) so that the CFL
	// array can be reduced as float4
	uint nblocks = div_up(to - from, BLOCK_FORCES);
	nblocks = (nblocks + 3) / 4 * 4;
	const DevParams &d = ctx->dp;
	const bool artvisc = d.turbmodel == B200SPH_TURB_ARTIFICIAL, laminar = !d.inviscid, multi = d.numFluids > 1;
	forces_kernel_t kern;
	switch (d.densitydiffusiontype) {
	case B200SPH_RHODIFF_FERRARI: kern = pick_forces_kernel<B200SPH_RHODIFF_FERRARI>(artvisc, laminar, multi); break;
	case B200SPH_RHODIFF_COLAGROSSI: kern = pick_forces_kernel<B200SPH_RHODIFF_COLAGROSSI>(artvisc, laminar, multi); break;
	default: kern = pick_forces_kernel<B200SPH_RHODIFF_NONE>(artvisc, laminar, multi); break;
	}
	kern<<<nblocks, BLOCK_FORCES, 0, ctx->stream>>>(ctx->dp, (const ushort4 *)info, hash, (const float4 *)ctx->eos,
		cell_start, neibs_list, (float4 *)forces, cfl, from, to, cfl_offset);
	KERNEL_TRY();
	if (num_cfl_blocks) *num_cfl_blocks = nblocks;
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

extern "C" int b200sph_dtreduce(b200sph_ctx *ctx, const float *cfl, float *temp_cfl, uint32_t num_blocks, float *dt_out)
{
	CHECK_CTX(ctx);
	(void)temp_cfl;
	if (!cfl || !dt_out) { b200_set_error("dtreduce: null buffer"); return B200SPH_EINVAL; }
	const b200sph_params &hp = ctx->hp;
	max_reduce_kernel<<<1, 1024, 0, ctx->stream>>>(cfl, num_blocks, ctx->d_scalar);
	KERNEL_TRY();
	CUDA_TRY(cudaMemcpyAsync(ctx->h_scalar, ctx->d_scalar, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
	CUDA_TRY(cudaStreamSynchronize(ctx->stream));
	const float maxcfl = ctx->h_scalar[0];
	float dt = hp.dtadaptfactor * fminf(sqrtf(hp.slength / maxcfl), hp.slength / hp.max_sound_speed_cfl);
	if (hp.rheologytype != B200SPH_RHEOLOGY_INVISCID || hp.turbmodel > B200SPH_TURB_ARTIFICIAL) {
		float dt_visc = hp.slength * hp.slength / hp.max_kinvisc;
		dt_visc *= 0.125f;
		if (dt_visc < dt) dt = dt_visc;
	}
	*dt_out = dt;
	return B200SPH_OK;
}
