// examples/lattice_steps.cu — the C ABI driven from plain C++ (no Python, no GPUSPH headers): a cubic lattice of
// fluid particles stepped with the whole-step entry points — neighbour rebuild every 10
// steps, forces with the integration in the kernel's epilogue (b200sph_forces_euler), dt on the device.
//
//   nvcc -O2 -std=c++17 -Iinclude -o build/lattice_steps examples/lattice_steps.cu -Lgpusph_b200 -lb200sph \
//        -Xlinker -rpath -Xlinker '$ORIGIN/../gpusph_b200'
//   build/lattice_steps [n = 64] [steps = 20]
//
// The host-side setup restates what ProblemCore does before the first step (src/ProblemCore.cc:1433-1496
// set_grid_params, :1554-1583 calc_localpos_and_hash; defaults src/simparams.h:280-300, src/physparams.h:385-400),
// like gpusph_b200/problems.py does for the tests; the call order is GPUWorker's (src/GPUWorker.cc:1779-2269).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "b200sph.h"

#define CK(call) do { int rc__ = (call); if (rc__) { fprintf(stderr, "%s failed (%d): %s\n", #call, rc__, b200sph_last_error()); return 1; } } while (0)
#define CU(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e__)); return 1; } } while (0)

int main(int argc, char **argv)
{
	const int n = argc > 1 ? atoi(argv[1]) : 64, steps = argc > 2 ? atoi(argv[2]) : 20;
	const uint32_t N = (uint32_t)n * n * n;
	const double dp = 0.01, rho0 = 1000.0, c0 = 20.0, gamma = 7.0, pad = 2.6 * dp, L = n * dp;

	// ---- parameters (the subset of SimParams / PhysParams the engines read) ----
	b200sph_params p;
	memset(&p, 0, sizeof(p));
	p.abi_version = B200SPH_ABI_VERSION;
	const float slength = (float)(1.3 * dp), influence = slength * 2.0f;
	for (int a = 0; a < 3; ++a) {
		const double size = L + 2 * pad;
		p.grid_size[a] = (uint32_t)floor(size / influence);              // ProblemCore.cc:1475-1477
		p.cell_size[a] = (float)(size / p.grid_size[a]);                  // :1491-1493
		p.world_origin[a] = (float)-pad;
	}
	p.coord[0] = 1; p.coord[1] = 2; p.coord[2] = 0;                       // reference default linearisation yzx
	p.neiblistsize = 128; p.neibboundpos = 127; p.neiblist_stride = N;
	p.nl_sq_influence_radius = influence * influence;
	p.kerneltype = B200SPH_KERNEL_WENDLAND; p.sph_formulation = B200SPH_SPH_F1;
	p.densitydiffusiontype = B200SPH_RHODIFF_NONE; p.boundarytype = B200SPH_DYN_BOUNDARY;
	p.rheologytype = B200SPH_RHEOLOGY_INVISCID; p.turbmodel = B200SPH_TURB_ARTIFICIAL;
	p.slength = slength; p.influenceradius = influence; p.deltap = (float)dp;
	p.dtadaptfactor = 0.3f; p.num_fluids = 1;
	p.rho0[0] = (float)rho0; p.bcoeff[0] = (float)(rho0 * c0 * c0 / gamma); p.gammacoeff[0] = (float)gamma;
	p.sscoeff[0] = (float)c0; p.sspowercoeff[0] = (float)((gamma - 1.0) / 2.0);
	p.gravity[2] = -9.81f;
	p.artvisccoeff = 0.3f; p.epsartvisc = (float)(0.01 * (double)slength * (double)slength);
	p.max_sound_speed_cfl = (float)((double)(float)c0 * 1.1); p.dtadapt = 1;       // float *= 1.1 (double), GPUWorker.cc:3011 p.simflags = B200SPH_ENABLE_DTADAPT;
	p.epsxsph = 0.5f; p.monaghan_visc_coeff = 10.0f;
	p.r0 = (float)dp; p.dcoeff = (float)(5.0 * 9.81); p.p1coeff = 12.0f; p.p2coeff = 6.0f;
	CK(b200sph_validate(&p));

	// ---- particles: (i + 1/2) dp, v = 0.1 c0 (sin, cos, sin)(2 pi x / L), cell-local positions + cell hash ----
	std::vector<float> pos(4 * (size_t)N), vel(4 * (size_t)N, 0.0f);
	std::vector<uint16_t> info(4 * (size_t)N, 0);
	std::vector<uint32_t> hash(N);
	const uint32_t G[3] = { p.grid_size[0], p.grid_size[1], p.grid_size[2] };
	uint32_t i = 0;
	for (int ix = 0; ix < n; ++ix) for (int iy = 0; iy < n; ++iy) for (int iz = 0; iz < n; ++iz, ++i) {
		const double g[3] = { (ix + 0.5) * dp, (iy + 0.5) * dp, (iz + 0.5) * dp };
		uint32_t c[3];
		for (int a = 0; a < 3; ++a) {
			// relative to the FLOAT world origin the engines were given (ProblemCore.cc:1554-1583)
			const double rel = g[a] - (double)p.world_origin[a];
			long q = (long)floor(rel / (double)p.cell_size[a]);
			c[a] = (uint32_t)(q < 0 ? 0 : (q >= (long)G[a] ? G[a] - 1 : q));
			pos[4 * (size_t)i + a] = (float)(rel - (c[a] + 0.5) * (double)p.cell_size[a]);
			vel[4 * (size_t)i + a] = (float)(0.1 * c0 * (a == 1 ? cos(2 * M_PI * g[a] / L) : sin(2 * M_PI * g[a] / L)));
		}
		pos[4 * (size_t)i + 3] = (float)(rho0 * dp * dp * dp);
		hash[i] = c[p.coord[2]] * G[p.coord[1]] * G[p.coord[0]] + c[p.coord[1]] * G[p.coord[0]] + c[p.coord[0]];
		info[4 * (size_t)i + 2] = (uint16_t)(i & 0xFFFF); info[4 * (size_t)i + 3] = (uint16_t)(i >> 16);   // PT_FLUID, id = i
	}

	// ---- device buffers in the reference's layouts ----
	const uint32_t ncells = G[0] * G[1] * G[2], ncfl = 2 * (b200sph_fmax_elements(N) + 8);
	void *d_pos[2], *d_vel[2], *d_info, *d_forces;
	uint32_t *d_hash, *d_pidx, *d_cs, *d_ce, *d_newn;
	uint16_t *d_nl;
	float *d_cfl;
	for (int s = 0; s < 2; ++s) { CU(cudaMalloc(&d_pos[s], 16 * (size_t)N)); CU(cudaMalloc(&d_vel[s], 16 * (size_t)N)); }
	CU(cudaMalloc(&d_info, 8 * (size_t)N)); CU(cudaMalloc(&d_forces, 16 * (size_t)N));
	CU(cudaMalloc(&d_hash, 4 * (size_t)N)); CU(cudaMalloc(&d_pidx, 4 * (size_t)N));
	CU(cudaMalloc(&d_cs, 4 * (size_t)ncells)); CU(cudaMalloc(&d_ce, 4 * (size_t)ncells)); CU(cudaMalloc(&d_newn, 4));
	CU(cudaMalloc(&d_nl, 2 * (size_t)p.neiblistsize * N)); CU(cudaMalloc(&d_cfl, 4 * (size_t)ncfl));
	CU(cudaMemcpy(d_pos[0], pos.data(), 16 * (size_t)N, cudaMemcpyHostToDevice));
	CU(cudaMemcpy(d_vel[0], vel.data(), 16 * (size_t)N, cudaMemcpyHostToDevice));
	CU(cudaMemcpy(d_info, info.data(), 8 * (size_t)N, cudaMemcpyHostToDevice));
	CU(cudaMemcpy(d_hash, hash.data(), 4 * (size_t)N, cudaMemcpyHostToDevice));
	CU(cudaMemset(d_forces, 0, 16 * (size_t)N));

	b200sph_ctx *ctx;
	CK(b200sph_create(&p, &ctx));
	// first dt: ProblemCore::check_dt (src/ProblemCore.cc:748-803) = dtadaptfactor * min(h / c0, sqrt(h / |g|))
	const float dt_ss = slength / (float)c0 * p.dtadaptfactor;
	const float dt_g = (float)sqrt((double)slength / sqrt((double)p.gravity[2] * (double)p.gravity[2])) * p.dtadaptfactor;
	CK(b200sph_step_set_dt(ctx, fminf(dt_ss, dt_g)));

	int cur = 0;
	uint32_t np = N;
	b200sph_neibs_info ni;
	memset(&ni, 0, sizeof(ni));
	long long interactions = 0;
	cudaEvent_t e0, e1;
	CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
	CU(cudaEventRecord(e0));
	for (int it = 0; it < steps; ++it) {
		if (it % 10 == 0) {                                          // NEIBS_LIST phase, src/Integrator.cc:93-249
			if (it == 0) CK(b200sph_fix_hash(ctx, d_hash, d_pidx, d_info, NULL, np));
			else CK(b200sph_calc_hash(ctx, d_pos[cur], d_hash, d_pidx, d_info, NULL, np));
			CK(b200sph_sort(ctx, d_hash, d_info, d_pidx, np));
			CU(cudaMemset(d_cs, 0xFF, 4 * (size_t)ncells));              // the caller clobbers CELLSTART, src/GPUWorker.cc:1846
			CK(b200sph_reorder(ctx, d_cs, d_ce, NULL, d_pos[1 - cur], d_vel[1 - cur], d_pos[cur], d_vel[cur], NULL, 0,
				d_info, d_hash, d_pidx, np, d_newn));
			cur = 1 - cur;
			CU(cudaMemcpy(&np, d_newn, 4, cudaMemcpyDeviceToHost));
			CK(b200sph_neibs_resetinfo(ctx));
			CU(cudaMemset(d_nl, 0xFF, 2 * (size_t)p.neiblistsize * N));  // and NEIBSLIST, :1883
			CK(b200sph_build_neibs(ctx, d_pos[cur], d_info, d_hash, d_cs, d_ce, d_nl, np, np));
			CK(b200sph_neibs_getinfo(ctx, &ni));
		}
		// predictor: forces(n) + euler dt/2 -> n*;  corrector: forces(n*) + euler dt IN PLACE -> n+1 (stays in `cur`)
		for (int step = 1; step <= 2; ++step) {
			b200sph_forces_args f;
			memset(&f, 0, sizeof(f));
			f.pos = d_pos[step == 1 ? cur : 1 - cur]; f.vel = d_vel[step == 1 ? cur : 1 - cur];
			f.info = d_info; f.hash = d_hash; f.cell_start = d_cs; f.neibs_list = d_nl; f.forces = d_forces; f.cfl = d_cfl;
			f.num_particles = np; f.from_particle = 0; f.to_particle = np; f.cfl_offset = 0; f.step = step; f.dt_from_device = 1;
			b200sph_fused_euler_args e;
			memset(&e, 0, sizeof(e));
			e.old_pos = d_pos[cur]; e.old_vel = d_vel[cur];
			e.new_pos = d_pos[step == 1 ? 1 - cur : cur]; e.new_vel = d_vel[step == 1 ? 1 - cur : cur];
			e.dt = 0.0f; e.step = step; e.dt_from_device = 1;
			uint32_t nblocks = 0;
			CK(b200sph_forces_euler(ctx, &f, &e, &nblocks));
			CK(b200sph_dtreduce_async(ctx, d_cfl, nblocks, step));
		}
		CK(b200sph_step_end(ctx));
		interactions += 2LL * ni.num_interactions;
	}
	CU(cudaEventRecord(e1));
	double t; float dt; uint64_t iters;
	CK(b200sph_step_query(ctx, &t, &dt, &iters));                    // synchronises
	float ms = 0.0f;
	CU(cudaEventElapsedTime(&ms, e0, e1));
	CU(cudaMemcpy(pos.data(), d_pos[cur], 16 * (size_t)np, cudaMemcpyDeviceToHost));
	CU(cudaMemcpy(vel.data(), d_vel[cur], 16 * (size_t)np, cudaMemcpyDeviceToHost));
	double sv = 0.0, sr = 0.0;
	for (uint32_t k = 0; k < np; ++k) { sv += fabs((double)vel[4 * (size_t)k]) + fabs((double)vel[4 * (size_t)k + 1]) + fabs((double)vel[4 * (size_t)k + 2]); sr += vel[4 * (size_t)k + 3]; }
	printf("{\"particles\": %u, \"steps\": %d, \"iterations\": %llu, \"t\": %.9e, \"dt\": %.9e, \"neibs_per_particle\": %.3f, "
		"\"sum_abs_vel\": %.9e, \"sum_rho_tilde\": %.9e, \"ms_per_step\": %.4f, \"M_interactions_per_s\": %.1f}\n",
		np, steps, (unsigned long long)iters, t, (double)dt, (double)ni.num_interactions / np, sv, sr, ms / steps,
		interactions / (ms * 1e-3) / 1e6);
	b200sph_destroy(ctx);
	return 0;
}
