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

int b200_eos_precompute(b200sph_ctx *ctx, const float4 *vel, const ushort4 *info, uint n)
{
	if (ctx->eos_cap < n) {
		cudaFree(ctx->eos); ctx->eos = NULL; ctx->eos_cap = 0;
		const size_t cap = (size_t)n + (n >> 3) + 1024;
		CUDA_TRY(cudaMalloc(&ctx->eos, cap * sizeof(float2)));
		ctx->eos_cap = cap;
	}
	eos_kernel<<<div_up(n, BLOCK_STREAM), BLOCK_STREAM, 0, ctx->stream>>>(ctx->dp, vel, info, ctx->eos, n);
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

// One neighbour-list section of one particle.
//  NFLUID : neighbours are fluid (section grows upwards from row 0) or boundary (downwards from neibboundpos)
//  MOMENTUM: accumulate the momentum equation (false for DYN boundary particles without force feedback)
template<bool NFLUID>
__device__ __forceinline__ void
walk_section(const DevParams &P, const uint index, const float4 pos, const float4 vel, const int3 gp,
	const int fnum, const float rho, const float p_precalc, const float sspeed, const bool momentum,
	const float4 *__restrict__ posArray, const float4 *__restrict__ velArray, const ushort4 *__restrict__ infoArray,
	const float2 *__restrict__ eos, const uint *__restrict__ cellStart, const ushort *__restrict__ neibsList,
	float4 &acc)
{
	const size_t stride = P.stride;
	float pcx = 0.f, pcy = 0.f, pcz = 0.f;
	uint base = 0;
	// neighbour-list iteration, reference src/cuda/neibs_iteration.cuh:56-200 + getNeibIndex src/cuda/cellgrid.cuh:198-226
	long long slot = NFLUID ? 0 : (long long)P.neibboundpos;
	const long long step = NFLUID ? 1 : -1;
	const float h = P.slength;
	for (;; slot += step) {
		uint nd = neibsList[(size_t)slot * stride + index];
		if (nd == NEIBS_END) break;
		if (nd >= CELLNUM_ENCODED) {
			const int cell = (int)(nd >> CELLNUM_SHIFT) - 1;
			nd &= NEIBINDEX_MASK;
			const int ox = cell % 3 - 1, oy = (cell / 3) % 3 - 1, oz = cell / 9 - 1;
			pcx = pos.x - (float)ox * P.cellSize[0];
			pcy = pos.y - (float)oy * P.cellSize[1];
			pcz = pos.z - (float)oz * P.cellSize[2];
			int gx = gp.x + ox, gy = gp.y + oy, gz = gp.z + oz;
			// calcGridHashPeriodic, cellgrid.cuh:174-185
			if (gx < 0) gx = P.gridSize[0] - 1; if (gx >= P.gridSize[0]) gx = 0;
			if (gy < 0) gy = P.gridSize[1] - 1; if (gy >= P.gridSize[1]) gy = 0;
			if (gz < 0) gz = P.gridSize[2] - 1; if (gz >= P.gridSize[2]) gz = 0;
			base = __ldg(cellStart + grid_hash(P, gx, gy, gz));
		}
		const uint j = base + nd;
		const float4 np = __ldg(posArray + j);
		const float rx = pcx - np.x, ry = pcy - np.y, rz = pcz - np.z;
		const float nmass = np.w;
		if (inactive_w(nmass)) continue;                                   // forces_kernel.def:3987
		const float r = sqrtf(rx * rx + ry * ry + rz * rz);
		if (r >= P.influenceradius) continue;                              // :3999
		const float4 nv = __ldg(velArray + j);
		const float2 ne = __ldg(eos + j);
		// common_neib_data :1099-1130
		const float rvx = vel.x - nv.x, rvy = vel.y - nv.y, rvz = vel.z - nv.z;
		const float vel_dot_pos = rvx * rx + rvy * ry + rvz * rz;
		const float qm2 = r / h - 2.0f;                                     // F<WENDLAND>, sph_core.cu:168-174
		const float f = qm2 * qm2 * qm2 * P.fcoeff_wendland;
		const float nsspeed = ne.y, np_precalc = ne.x;
		int nfnum = 0;
		float nrho;
		if (P.numFluids > 1) {
			nfnum = fluid_num_of(__ldg(infoArray + j));
			nrho = phys_density(P, nv.w, nfnum);
		} else
			nrho = phys_density(P, nv.w, 0);

		// --- continuity: mass_continuity_div_vel_term :2140-2150 ---
		float DrDt = nmass * vel_dot_pos * f;
		if (NFLUID) {   // no density diffusion from DYN boundary neighbours, :1594-1606
			if (P.densitydiffusiontype == B200SPH_RHODIFF_FERRARI) {        // :1614-1636
				const float grav_corr = -(P.gravity[0] * rx + P.gravity[1] * ry + P.gravity[2] * rz) * P.rho0[fnum] / P.sqC0[fnum];
				float fx = 0.f, fy = 0.f, fz = 0.f;
				if (r > 1e-4f * h) {
					const float s = fmaxf(sspeed, nsspeed) * (rho - nrho + grav_corr) / rho / r;
					fx = s * rx; fy = s * ry; fz = s * rz;
				}
				DrDt += P.densityDiffCoeff * nmass * (fx * rx + fy * ry + fz * rz) * f;
			} else if (P.densitydiffusiontype == B200SPH_RHODIFF_COLAGROSSI) {   // :1916-1951
				if (fnum == nfnum) {
					// P(rho) recovered from the precomputed P/rho^2
					const float Pi = p_precalc * (rho * rho), Pj = np_precalc * (nrho * nrho);
					const float gdot = P.gravity[0] * rx + P.gravity[1] * ry + P.gravity[2] * rz;
					if (!(fabsf(Pi - Pj) < fabsf(gdot * rho)))
						DrDt -= P.densityDiffCoeff * P.sscoeff[fnum] * (nrho / rho - 1) * f * nmass;
				}
			}
		}
		acc.w += DrDt;                                                      // :2189

		if (momentum) {
			float dvx, dvy, dvz;
			// compute_pressure_contrib, general formulation :2450-2466
			const float pg = (p_precalc + np_precalc) * nmass * f;
			dvx = -pg * rx; dvy = -pg * ry; dvz = -pg * rz;
			// artificial viscosity :2744-2764, artvisc visc_kernel.cu:75-85
			if (P.turbmodel == B200SPH_TURB_ARTIFICIAL && vel_dot_pos < 0.0f) {
				const float visc = vel_dot_pos * h * P.artvisccoeff * (sspeed + nsspeed) / ((r * r + P.epsartvisc) * (rho + nrho));
				const float s = visc * nmass * f;
				dvx += s * rx; dvy += s * ry; dvz += s * rz;
			}
			// laminar (Morris) :2605-2625
			if (!P.inviscid) {
				const float vc = P.visccoeff[fnum], nvc = P.visccoeff[nfnum];
				float visc;
				if (P.compvisc == B200SPH_COMPVISC_KINEMATIC)
					visc = P.is_const_visc ? vc * visc_avg_density(P, rho, nrho, nmass) : visc_avg_dyn(P, vc * rho, nvc * nrho, rho, nrho, nmass);
				else
					visc = P.is_const_visc ? 2 * nmass * vc / (rho * nrho) : visc_avg_dyn(P, vc, nvc, rho, nrho, nmass);
				const float s = visc * f;
				dvx += s * rvx; dvy += s * rvy; dvz += s * rvz;
			}
			acc.x += dvx; acc.y += dvy; acc.z += dvz;                       // :3590
		}
	}
}

__global__ void __launch_bounds__(BLOCK_FORCES)
forces_kernel(const __grid_constant__ DevParams P, const float4 *__restrict__ posArray, const float4 *__restrict__ velArray,
	const ushort4 *__restrict__ infoArray, const uint *__restrict__ particleHash, const float2 *__restrict__ eos,
	const uint *__restrict__ cellStart, const ushort *__restrict__ neibsList,
	float4 *__restrict__ forces, float *__restrict__ cfl,
	const uint fromParticle, const uint toParticle, const uint cflOffset)
{
	const uint index = blockIdx.x * blockDim.x + threadIdx.x + fromParticle;
	float cfl_term = 0.0f;

	if (index < toParticle) {
		const ushort4 info = infoArray[index];
		const int type = ptype_of(info);
		const float4 pos = posArray[index];
		if ((type == PT_FLUID || type == PT_BOUNDARY) && !inactive_w(pos.w)) {
			const float4 vel = velArray[index];
			const float2 e = eos[index];
			const int fnum = fluid_num_of(info);
			const float rho = phys_density(P, vel.w, fnum);
			const int3 gp = grid_pos(P, particleHash[index] & CELLTYPE_BITMASK);
			float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
			if (type == PT_FLUID) {
				// forcesDevice<fluid,fluid> then <fluid,boundary> (forces.cu:759,782); DYN boundary neighbours
				// interact like fluid ones (forces_kernel.def:3717-3726)
				walk_section<true>(P, index, pos, vel, gp, fnum, rho, e.x, e.y, true, posArray, velArray, infoArray, eos, cellStart, neibsList, acc);
				walk_section<false>(P, index, pos, vel, gp, fnum, rho, e.x, e.y, true, posArray, velArray, infoArray, eos, cellStart, neibsList, acc);
			} else {
				// forcesDevice<boundary,fluid> (forces.cu:792): density always, momentum only with force feedback
				// (forces_kernel.def:3634-3667)
				walk_section<true>(P, index, pos, vel, gp, fnum, rho, e.x, e.y, (info.x & B200SPH_FG_COMPUTE_FORCE) != 0,
					posArray, velArray, infoArray, eos, cellStart, neibsList, acc);
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
	int rc = b200_eos_precompute(ctx, (const float4 *)vel, (const ushort4 *)info, num_particles);
	if (rc) return rc;
	// grid rounded to a multiple of 4 blocks like the reference (forces.cu:741-744) so that the CFL
	// array can be reduced as float4
	uint nblocks = div_up(to - from, BLOCK_FORCES);
	nblocks = (nblocks + 3) / 4 * 4;
	forces_kernel<<<nblocks, BLOCK_FORCES, 0, ctx->stream>>>(ctx->dp, (const float4 *)pos, (const float4 *)vel,
		(const ushort4 *)info, hash, ctx->eos, cell_start, neibs_list, (float4 *)forces, cfl, from, to, cfl_offset);
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
