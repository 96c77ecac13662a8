// pair_physics.cuh — per-pair WCSPH physics of the forces kernel.
// Behavioural specification: GPUSPH src/cuda/forces_kernel.def (lines cited inline), sph_core.cu, phys_core.cu,
// visc_kernel.cu, visc_avg.cu.
#pragma once
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


#define RHODIFF_RUNTIME_VALUE (-1)      // = RHODIFF_RUNTIME, defined with its explanation below
struct PairConsts {
	float inv_h, fc, R2;       // 1/h, Wendland gradient coefficient, squared influence radius
	float h_alpha, eps;        // artificial viscosity: h*alpha, eps
	float g0, g1, g2;          // gravity
	float diff;                // density diffusion coefficient (Colagrossi single-fluid: xi*2h*c0)
	float grav_scale;          // Ferrari: rho0/c0^2
	float h;
};

// The lean single-fluid variants (one fluid, no laminar viscosity, physics fixed at compile time) work with density
// RATIOS rho/rho0 = 1 + rho~ instead of densities: rho0 folds into the constants (h_alpha -> h alpha / rho0,
// grav_scale -> 1/c0^2, the EOS prefactor -> B/rho0^2), and the factor diff / (rho_i/rho0) every density-diffusion
// term of a particle shares is applied once per particle, to the sum, instead of once per pair. Same algebra as the
// reference's expressions (cited below), fewer instructions in a loop that is bound by instruction issue.
#ifndef B200_RATIO_SPACE
#define B200_RATIO_SPACE 1
#endif
template<int RHODIFF, bool LAMINAR, bool MULTIFLUID>
struct RatioSpace { static constexpr bool value = B200_RATIO_SPACE && RHODIFF != RHODIFF_RUNTIME_VALUE && !LAMINAR && !MULTIFLUID; };

template<int RHODIFF, bool MULTIFLUID, bool LAMINAR = true>
__device__ __forceinline__ PairConsts make_pair_consts(const DevParams &P)
{
	constexpr bool FAST = RatioSpace<RHODIFF, LAMINAR, MULTIFLUID>::value;
	PairConsts k;
	k.h = P.slength; k.inv_h = 1.0f / P.slength; k.fc = P.fcoeff_wendland;
	k.R2 = P.influenceradius * P.influenceradius;
	k.h_alpha = P.slength * P.artvisccoeff; k.eps = P.epsartvisc;
	if (FAST) k.h_alpha /= P.rho0[0];
	k.g0 = P.gravity[0]; k.g1 = P.gravity[1]; k.g2 = P.gravity[2];
	k.diff = RHODIFF == B200SPH_RHODIFF_COLAGROSSI && !MULTIFLUID ? P.densityDiffCoeff * P.sscoeff[0] : P.densityDiffCoeff;
	k.grav_scale = FAST ? 1.0f / P.sqC0[0] : P.rho0[0] / P.sqC0[0];
	// Ferrari, ratio space: gravity only enters as g.r / c0^2
	if (FAST && RHODIFF == B200SPH_RHODIFF_FERRARI) { k.g0 *= k.grav_scale; k.g1 *= k.grav_scale; k.g2 *= k.grav_scale; }
	return k;
}

__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rsqrt_approx(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// 16-bit neighbour-list entry, zero-extended, through the read-only path
// B200_LIST_CACHE: L1 policy of the list stream (every entry is used exactly once by exactly one thread, so it only
// displaces the neighbour records the gathers want to find in L1): 0 default, 1 no_allocate, 2 evict_first
#ifndef B200_LIST_CACHE
#define B200_LIST_CACHE 0
#endif
__device__ __forceinline__ uint ld_neib(const ushort *p)
{
	uint v;
#if B200_LIST_CACHE == 1
	asm volatile("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=r"(v) : "l"(p));
#elif B200_LIST_CACHE == 2
	asm volatile("ld.global.nc.L1::evict_first.u16 %0, [%1];" : "=r"(v) : "l"(p));
#else
	asm volatile("ld.global.nc.u16 %0, [%1];" : "=r"(v) : "l"(p));
#endif
	return v;
}
__device__ __forceinline__ float lg2_approx(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// {P/rho^2, sound speed, density, fluid#} of a particle from its relative density, evaluated per pair like the
// reference does (P() and soundSpeed() with __powf = ex2(y*lg2(x)), phys_core.cu:99-136; precalc_pressure
// forces_kernel.def:419-429) but sharing the logarithm between the two powers. Trading ~12 ALU/MUFU instructions
// for one scattered 128-bit gather is a win because the pair kernel is L1-wavefront bound, not issue bound.
// e.w: the fluid number (as int bits) from the DevParams overload, the raw pressure from the EosConsts (single-fluid) one
struct EosConsts { float gamma, sspow, b, ss, rho0; };      // of fluid 0, kept in registers by the single-fluid kernels
__device__ __forceinline__ float4 eos_from_density(const EosConsts &E, const float rho_tilde)
{
	const float ratio = rho_tilde + 1.0f;
	const float lg = lg2_approx(ratio);
	const float pw = ex2_approx(E.gamma * lg);
	const float rho = ratio * E.rho0;
	float4 e;
	e.x = E.b * (pw - 1.0f) * rcp_approx(rho * rho);
	e.y = E.ss * ex2_approx(E.sspow * lg);
	e.z = rho;
	e.w = E.b * (pw - 1.0f);        // single fluid: the raw pressure P() rides in the slot the fluid number takes otherwise
	return e;
}
// ratio-space flavour (RatioSpace variants): e.z = rho/rho0, and the constants arrive folded: E.b = B/rho0^2,
// E.rho0 = B (the slot of the density scale, which this flavour does not need)
// (is_ratio: the caller already added the 1)
__device__ __forceinline__ float4 eos_ratio_from_density(const EosConsts &E, const float rho_tilde, const bool is_ratio = false)
{
	const float ratio = is_ratio ? rho_tilde : rho_tilde + 1.0f;
	const float lg = lg2_approx(ratio);
	const float pw1 = ex2_approx(E.gamma * lg) - 1.0f;
	float4 e;
	e.x = E.b * pw1 * rcp_approx(ratio * ratio);
	e.y = E.ss * ex2_approx(E.sspow * lg);
	e.z = ratio;
	e.w = E.rho0 * pw1;             // raw pressure (Molteni-Colagrossi switch; dead code elsewhere)
	return e;
}
__device__ __forceinline__ float4 eos_from_density(const DevParams &P, const float rho_tilde, const int f)
{
	EosConsts E;
	E.gamma = P.gammacoeff[f]; E.sspow = P.sspowercoeff[f]; E.b = P.bcoeff[f]; E.ss = P.sscoeff[f]; E.rho0 = P.rho0[f];
	float4 e = eos_from_density(E, rho_tilde);
	e.w = __int_as_float(f);
	return e;
}

// the central particle of a pair
struct Central {
	float4 pos, vel;
	float rho, p_precalc, sspeed;
	float ratio;      // rho/rho0 (RatioSpace variants)
	float press;      // raw pressure P(rho~), Molteni-Colagrossi switch only
	int fnum;
	bool momentum;    // accumulate the momentum equation (false for DYN boundary particles without force feedback)
	bool xsph;        // accumulate the XSPH mean velocity (fluid particle and ENABLE_XSPH; general variant only)
};

// RHODIFF template value of the GENERAL kernel variant: every physics option is read from DevParams at run time
// (uniform branches). It serves the options that are not worth a specialised instantiation each — BREZZI diffusion,
// the MONAGHAN / ESPANOL_REVENGA viscous models, XSPH — while the specialised variants stay lean for the
// benchmarked configurations.
#define RHODIFF_RUNTIME (-1)

// average<avgop>, src/average.h:75-100
__device__ __forceinline__ float average_op(const uint op, const float a, const float b)
{
	switch (op) {
	case B200SPH_AVG_ARITHMETIC: return (a + b) * 0.5f;
	case B200SPH_AVG_HARMONIC: return 2.0f * a * b / (a + b);
	default: return sqrtf(a * b);
	}
}

// One pair interaction. (rx,ry,rz) = relPos, r2 its squared length (already known to be inside the support),
// nmass = neighbour mass, rv = {relative velocity v_i - v_j, neighbour rho~} (the caller subtracts: the pair loop frees the
// registers of the neighbour record as early as it can), ne = neighbour {P/rho^2, sound speed, density, fluid#}.
// nfluid: the neighbour is a fluid particle (density diffusion applies), else a DYN boundary particle
// (forces_kernel.def:1594-1606, 3717-3726). xs accumulates the XSPH mean velocity (general variant only).
template<int RHODIFF, bool ARTVISC, bool LAMINAR, bool MULTIFLUID>
__device__ __forceinline__ void
pair_interaction_x(const DevParams &P, const PairConsts &k, const Central &c, const float rx, const float ry, const float rz,
	const float r2, const float nmass, const float4 rv, const float4 ne, const bool nfluid, float4 &acc, float3 &xs)
{
	constexpr bool GEN = RHODIFF == RHODIFF_RUNTIME;
	// FAST: ne.z and `rho` are density ratios, xs.x collects the density-diffusion sum still to be scaled by diff/ratio_i
	constexpr bool FAST = RatioSpace<RHODIFF, LAMINAR, MULTIFLUID>::value;
	const int rhodiff = GEN ? (int)P.densitydiffusiontype : RHODIFF;
	const bool artvisc = GEN ? P.turbmodel == B200SPH_TURB_ARTIFICIAL : ARTVISC;
	const bool laminar = GEN ? !P.inviscid : LAMINAR;
	const int nfnum = MULTIFLUID ? __float_as_int(ne.w) : 0;
	const float r = r2 * rsqrt_approx(r2 + 1e-30f);
	// common_neib_data :1099-1130
	const float rvx = rv.x, rvy = rv.y, rvz = rv.z;
	const float vel_dot_pos = fmaf(rvz, rz, fmaf(rvy, ry, rvx * rx));
	const float qm2 = fmaf(r, k.inv_h, -2.0f);                       // F<WENDLAND>, sph_core.cu:168-174
	const float f = qm2 * qm2 * qm2 * k.fc;
	const float mf = nmass * f;
	const float np_precalc = ne.x, nsspeed = ne.y, nrho = ne.z;
	const float rho = FAST ? c.ratio : c.rho;

	// --- continuity: mass_continuity_div_vel_term :2140-2150 ---
	float DrDt = mf * vel_dot_pos;
	if (nfluid) {
		if (rhodiff == B200SPH_RHODIFF_FERRARI) {                   // :1614-1636
			const float gdot = fmaf(k.g2, rz, fmaf(k.g1, ry, k.g0 * rx));
			if (FAST) {
				// (rho - rho_j + corr) / rho = ((a - b) - g.r / c0^2) / a with a, b the density ratios; diff / a: caller
				// (k.g* arrive divided by c0^2 in this variant)
				const float s = fmaxf(c.sspeed, nsspeed) * ((rho - nrho) - gdot) * r;
				xs.x = fmaf(mf, (r > 1e-4f * k.h) ? s : 0.0f, xs.x);
			} else {
			const float grav_corr = -gdot * (MULTIFLUID ? P.rho0[c.fnum] / P.sqC0[c.fnum] : k.grav_scale);
			// ferraricor . relPos = max(c) (rho - rho_j + corr)/rho / r * r^2   (zero for r <= 1e-4 h)
			const float s = (r > 1e-4f * k.h) ? fmaxf(c.sspeed, nsspeed) * (rho - nrho + grav_corr) * rcp_approx(rho) * r : 0.0f;
			DrDt = fmaf(k.diff * mf, s, DrDt);
			}
		} else if (rhodiff == B200SPH_RHODIFF_COLAGROSSI) {         // :1916-1951
			if (!MULTIFLUID || c.fnum == nfnum) {
				// P() of both particles as the reference compares them (:1925-1928), not rebuilt from P/rho^2: the test is
				// a discontinuous switch, a borderline pair must fall on the reference's side
				const float Pi = c.press, Pj = MULTIFLUID ? eos_pressure(P, rv.w, nfnum) : ne.w;
				const float gdot = fmaf(k.g2, rz, fmaf(k.g1, ry, k.g0 * rx));
				if (FAST) {
					// (rho_j / rho - 1) = (b - a) / a; diff / a: caller
					if (!(fabsf(Pi - Pj) < fabsf(gdot * c.rho))) xs.x = fmaf(rho - nrho, mf, xs.x);
				} else
				if (!(fabsf(Pi - Pj) < fabsf(gdot * rho)))
					DrDt -= k.diff * (MULTIFLUID ? P.sscoeff[c.fnum] : 1.0f) * (nrho * rcp_approx(rho) - 1.0f) * mf;
			}
		} else if (GEN && rhodiff == B200SPH_RHODIFF_BREZZI) {      // :1765-1782
			const float dt = P.dev_state ? (P.cmd_step == 1 ? P.dev_state->dt / 2 : P.dev_state->dt) : P.cmd_dt;
			const float Pi = eos_pressure(P, c.vel.w, c.fnum), Pj = eos_pressure(P, rv.w, nfnum);
			const float gdot = fmaf(k.g2, rz, fmaf(k.g1, ry, k.g0 * rx));
			DrDt += k.diff * ((2.0f / (rho + nrho)) * (Pi - Pj) - gdot) * nmass / nrho * f * dt * 2.0f * rho;
		}
	}
	acc.w += DrDt;                                                  // :2189

	if (c.momentum) {
		// compute_pressure_contrib, general formulation :2450-2466:  -(P_i/rho_i^2 + P_j/rho_j^2) m_j F r_ij
		float coef = -(c.p_precalc + np_precalc) * mf;
		// artificial viscosity :2744-2764, artvisc visc_kernel.cu:75-85
		if (artvisc) {
			const float visc = vel_dot_pos * k.h_alpha * (c.sspeed + nsspeed) * rcp_approx((r2 + k.eps) * (rho + nrho));
			coef = (vel_dot_pos < 0.0f) ? fmaf(visc, mf, coef) : coef;
		}
		float dvx = coef * rx, dvy = coef * ry, dvz = coef * rz;
		// laminar viscosity :2605-2625 (Morris / Monaghan), :2651-2678 (Espanol & Revenga)
		if (laminar) {
			const float vc = P.visccoeff[c.fnum], nvc = P.visccoeff[nfnum];
			if (GEN && P.viscmodel == B200SPH_VISCMODEL_ESPANOL_REVENGA) {
				// dynamic viscosities (get_dynamic_visc :276-289), bulk viscosities d_visc2coeff
				const float pvisc = P.compvisc == B200SPH_COMPVISC_KINEMATIC ? vc * rho : vc;
				const float nvisc = P.compvisc == B200SPH_COMPVISC_KINEMATIC ? nvc * nrho : nvc;
				const float visc_thirds = average_op(P.viscavgop, pvisc, nvisc) / 3;
				const float bulk = average_op(P.viscavgop, P.visc2coeff[c.fnum], P.visc2coeff[nfnum]);
				const float cf = nmass / (rho * nrho) * f;            // viscous_volume_coefficient :2575-2580
				const float pos_den = r2 + k.eps;
				const float a = 5 * visc_thirds - bulk, b = 5 * (visc_thirds + bulk) * vel_dot_pos / pos_den;
				dvx += cf * (a * rvx + b * rx); dvy += cf * (a * rvy + b * ry); dvz += cf * (a * rvz + b * rz);
			} else {
				float visc;
				if (P.compvisc == B200SPH_COMPVISC_KINEMATIC)
					visc = P.is_const_visc ? vc * visc_avg_density(P, rho, nrho, nmass) : visc_avg_dyn(P, vc * rho, nvc * nrho, rho, nrho, nmass);
				else
					visc = P.is_const_visc ? 2 * nmass * vc / (rho * nrho) : visc_avg_dyn(P, vc, nvc, rho, nrho, nmass);
				const float s = visc * f;
				if (GEN && P.viscmodel == B200SPH_VISCMODEL_MONAGHAN) {
					// viscous_vector_component<MONAGHAN> :2534-2559: along relPos, only for approaching particles
					const float m = vel_dot_pos < 0 ? P.monaghanViscCoeff * vel_dot_pos / (r2 + k.eps) : 0.0f;
					dvx = fmaf(s, m * rx, dvx); dvy = fmaf(s, m * ry, dvy); dvz = fmaf(s, m * rz, dvz);
				} else {
					dvx = fmaf(s, rvx, dvx); dvy = fmaf(s, rvy, dvy); dvz = fmaf(s, rvz, dvz);
				}
			}
		}
		// compute_mean_vel :2986-2992 (fluid central, fluid neighbour, ENABLE_XSPH)
		if (GEN && nfluid && c.xsph) {
			float w = fmaf(-0.5f * r, k.inv_h, 1.0f);                 // W<WENDLAND>, sph_core.cu:104-117
			w *= w; w *= w;
			w *= fmaf(2.0f * r, k.inv_h, 1.0f);
			const float s = nmass * w * P.wcoeff_wendland / (rho + nrho);
			xs.x -= s * rvx; xs.y -= s * rvy; xs.z -= s * rvz;
		}
		acc.x += dvx; acc.y += dvy; acc.z += dvz;                   // :3590
	}
}

template<int RHODIFF, bool ARTVISC, bool LAMINAR, bool MULTIFLUID>
__device__ __forceinline__ void
pair_interaction(const DevParams &P, const PairConsts &k, const Central &c, const float rx, const float ry, const float rz,
	const float r2, const float nmass, const float4 rv, const float4 ne, const bool nfluid, float4 &acc)
{
	float3 xs = make_float3(0.f, 0.f, 0.f);
	pair_interaction_x<RHODIFF, ARTVISC, LAMINAR, MULTIFLUID>(P, k, c, rx, ry, rz, r2, nmass, rv, ne, nfluid, acc, xs);
}

// Lennard-Jones repulsion and wall friction of the geometric planes on a fluid particle: GeometryForce / PlaneForce /
// LJForce, src/cuda/forces_kernel.cu:94-204; PlaneDistance src/cuda/geom_core.cu:65-85.
__device__ __forceinline__ void plane_forces(const DevParams &P, const int3 gp, const float4 pos, const float4 vel,
	const float dynvisc, float4 &acc)
{
	for (uint i = 0; i < P.numplanes; ++i) {
		const float nx = P.planeNormal[i][0], ny = P.planeNormal[i][1], nz = P.planeNormal[i][2];
		// globalDistance(gridPos, pos, plane.gridPos, plane.pos), cellgrid.cuh:152-160
		const float dx = (float)(gp.x - P.planeGridPos[i][0]) * P.cellSize[0] + (pos.x - P.planePos[i][0]);
		const float dy = (float)(gp.y - P.planeGridPos[i][1]) * P.cellSize[1] + (pos.y - P.planePos[i][1]);
		const float dz = (float)(gp.z - P.planeGridPos[i][2]) * P.cellSize[2] + (pos.z - P.planePos[i][2]);
		const float r = fabsf(dx * nx + dy * ny + dz * nz);
		if (r < P.r0) {
			const float q = P.r0 / r;
			const float DvDt = P.dcoeff * (__powf(q, P.p1coeff) - __powf(q, P.p2coeff)) / (r * r);   // r <= r0 always here
			const float px = nx * r, py = ny * r, pz = nz * r;         // relPos = normal * r
			acc.x += DvDt * px; acc.y += DvDt * py; acc.z += DvDt * pz;
			// tangential velocity v_t = vel - dot(vel, relPos)/r * relPos/r, friction -mu A / (m r)
			const float vn = (vel.x * px + vel.y * py + vel.z * pz) / r;
			const float coeff = -dynvisc * P.partsurf / (pos.w * r);
			acc.x += coeff * (vel.x - vn * px / r); acc.y += coeff * (vel.y - vn * py / r); acc.z += coeff * (vel.z - vn * pz / r);
		}
	}
}

// finalizeforcesDevice :4037-4153: 1/rho0 on the continuity term (forces_fixup :3212-3219), gravity on fluid
// particles (:4091), CFL term max(|a|, c^2/h) (dyndt_forces_shared_data::store :3436-3456); for particles of a
// force-feedback body (FG_COMPUTE_FORCE) the acceleration is turned into a force (x mass) and, with its torque
// about the body's centre of gravity, scattered to the body buffers (:4116-4141).
struct BodyOut {
	const BodyData *bodies;     // NULL: no body output requested
	float4 *rb_forces, *rb_torques;
	float4 *xsph;               // XSPH mean-velocity output (general variant, ENABLE_XSPH), else NULL
	// fused integration epilogue (gather kernel only; b200sph_forces_euler): eul_step 1 / 2 integrates the particle
	// right after its forces are known — state n from eul_old_*, result to eul_new_* (0: forces only)
	int eul_step;
	float eul_dt;                          // dt of the sub-step unless eul_state is given
	const StepState *eul_state;            // device-resident dt record
	const float4 *eul_old_pos, *eul_old_vel;   // NULL: state n is the state the pair loop reads (predictor)
	float4 *eul_new_pos, *eul_new_vel;
	PosVel *eul_new_packed;                // the same state as neighbour records for the next force evaluation (NULL: none)
	const BodyData *eul_bodies;            // rigid motion of moving bodies (NULL: none)
};

__device__ __forceinline__ float finalize_particle(const DevParams &P, const int type, const int fnum, const float sspeed,
	const ushort4 info, const float4 pos, const float4 vel, const float rho, const uint cellHash, const BodyOut &bo, float4 &acc)
{
	acc.w /= P.rho0[fnum];
	float cfl_term = 0.0f;
	if (type == PT_FLUID) {
		acc.x += P.gravity[0]; acc.y += P.gravity[1]; acc.z += P.gravity[2];
		if (P.numplanes) {
			// viscous_plane_coefficient :3103-3113: free slip when inviscid, else the laminar dynamic viscosity
			const float dynvisc = P.inviscid ? 0.0f :
				(P.compvisc == B200SPH_COMPVISC_KINEMATIC ? P.visccoeff[fnum] * rho : P.visccoeff[fnum]);
			plane_forces(P, grid_pos(P, cellHash), pos, vel, dynvisc, acc);          // :4105-4110
		}
		cfl_term = fmaxf(sqrtf(acc.x * acc.x + acc.y * acc.y + acc.z * acc.z), sspeed * sspeed / P.slength);
	}
	if (bo.bodies && (info.x & B200SPH_FG_COMPUTE_FORCE) && type != PT_VERTEX) {
		const int obj = object_of_y(info.y);
		acc.x *= pos.w; acc.y *= pos.w; acc.z *= pos.w;               // :4130-4131 (the stored force is scaled too)
		const uint rbindex = id_of(info) + (uint)bo.bodies->startIndex[obj];   // rb_particle_data :520-528
		bo.rb_forces[rbindex] = acc;
		// arm = globalDistance(gridPos, pos, cg cell, cg in-cell pos), cellgrid.cuh:152-160
		const int3 gp = grid_pos(P, cellHash);
		const float ax = (float)(gp.x - bo.bodies->cgGridPos[obj][0]) * P.cellSize[0] + (pos.x - bo.bodies->cgPos[obj][0]);
		const float ay = (float)(gp.y - bo.bodies->cgGridPos[obj][1]) * P.cellSize[1] + (pos.y - bo.bodies->cgPos[obj][1]);
		const float az = (float)(gp.z - bo.bodies->cgGridPos[obj][2]) * P.cellSize[2] + (pos.z - bo.bodies->cgPos[obj][2]);
		bo.rb_torques[rbindex] = make_float4(ay * acc.z - az * acc.y, az * acc.x - ax * acc.z, ax * acc.y - ay * acc.x, 0.0f);
	}
	return cfl_term;
}
