// api.cu — context management and error plumbing of the C ABI (include/b200sph.h).
#include "common.cuh"
#include <stdarg.h>
#include <stdlib.h>
#include <math.h>

static thread_local char g_err[512] = "";

void b200_set_error(const char *fmt, ...)
{
	va_list ap; va_start(ap, fmt);
	vsnprintf(g_err, sizeof(g_err), fmt, ap);
	va_end(ap);
}

extern "C" const char *b200sph_last_error(void) { return g_err; }
extern "C" int b200sph_abi_version(void) { return B200SPH_ABI_VERSION; }

extern "C" int b200sph_device_count(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
	return n;
}

// Which option combinations are implemented. Everything else fails loudly
// (never silently fall back).
extern "C" int b200sph_validate(const b200sph_params *p)
{
	if (!p) { b200_set_error("null params"); return B200SPH_EINVAL; }
	if (p->abi_version != B200SPH_ABI_VERSION) {
		b200_set_error("ABI version mismatch: caller %u, library %d", p->abi_version, B200SPH_ABI_VERSION);
		return B200SPH_EINVAL;
	}
	for (int a = 0; a < 3; ++a) {
		if (p->grid_size[a] == 0 || !(p->cell_size[a] > 0)) { b200_set_error("empty grid / non-positive cell size"); return B200SPH_EINVAL; }
		if (p->coord[a] > 2) { b200_set_error("bad linearisation"); return B200SPH_EINVAL; }
	}
	if (p->coord[0] == p->coord[1] || p->coord[0] == p->coord[2] || p->coord[1] == p->coord[2]) {
		b200_set_error("linearisation must be a permutation of xyz"); return B200SPH_EINVAL;
	}
	const uint64_t ncells = (uint64_t)p->grid_size[0] * p->grid_size[1] * p->grid_size[2];
	if (ncells > (0xFFFFFFFFu >> 2)) { b200_set_error("too many cells (MAX_CELLS, src/multi_gpu_defines.h:56)"); return B200SPH_EINVAL; }
	if (p->neiblistsize < 4 || p->neibboundpos >= p->neiblistsize) { b200_set_error("bad neiblistsize/neibboundpos"); return B200SPH_EINVAL; }
	if (p->neiblist_block && (p->neiblist_block < 32 || (p->neiblist_block & (p->neiblist_block - 1)))) {
		b200_set_error("neiblist_block %u: a power of two >= 32 (or 0 for the default) expected", p->neiblist_block); return B200SPH_EINVAL;
	}
	if (p->num_fluids < 1 || p->num_fluids > B200SPH_MAX_FLUIDS) { b200_set_error("num_fluids out of range"); return B200SPH_EINVAL; }
	if (p->kerneltype != B200SPH_KERNEL_WENDLAND) { b200_set_error("unsupported SPH kernel %u: only WENDLAND is implemented", p->kerneltype); return B200SPH_EUNSUP; }
	if (p->sph_formulation != B200SPH_SPH_F1) { b200_set_error("unsupported SPH formulation %u: only SPH_F1 is implemented", p->sph_formulation); return B200SPH_EUNSUP; }
	if (p->boundarytype != B200SPH_DYN_BOUNDARY) { b200_set_error("unsupported boundary type %u: only DYN_BOUNDARY is implemented", p->boundarytype); return B200SPH_EUNSUP; }
	if (p->densitydiffusiontype > B200SPH_RHODIFF_BREZZI) { b200_set_error("unknown density diffusion %u", p->densitydiffusiontype); return B200SPH_EINVAL; }
	if (p->rheologytype > B200SPH_RHEOLOGY_NEWTONIAN) { b200_set_error("unsupported rheology %u", p->rheologytype); return B200SPH_EUNSUP; }
	if (p->turbmodel > B200SPH_TURB_ARTIFICIAL) { b200_set_error("unsupported turbulence model %u", p->turbmodel); return B200SPH_EUNSUP; }
	if (p->rheologytype == B200SPH_RHEOLOGY_NEWTONIAN && p->viscmodel > B200SPH_VISCMODEL_ESPANOL_REVENGA) { b200_set_error("unknown viscous model %u", p->viscmodel); return B200SPH_EINVAL; }
	if (p->simflags & ~B200SPH_SUPPORTED_SIMFLAGS) {
		// the whole of SimParams::simflags arrives here: anything outside the allow-list is refused, never ignored
		static const char *names[] = { "ENABLE_DTADAPT", "ENABLE_XSPH", "ENABLE_PLANES", "ENABLE_DEM", "ENABLE_MOVING_BODIES",
			"ENABLE_INLET_OUTLET", "ENABLE_WATER_DEPTH", "ENABLE_DENSITY_SUM", "ENABLE_GAMMA_QUADRATURE", "ENABLE_REPACKING",
			"ENABLE_INTERNAL_ENERGY", "ENABLE_MULTIFLUID" };
		const uint32_t bad = p->simflags & ~B200SPH_SUPPORTED_SIMFLAGS;
		int bit = 0; while (!((bad >> bit) & 1u)) ++bit;
		b200_set_error("unsupported simulation flag %s (0x%x; out of scope, SURVEY.md section 8)", bit < 12 ? names[bit] : "(unknown)", bad);
		return B200SPH_EUNSUP;
	}
	if ((p->simflags & B200SPH_ENABLE_MULTIFLUID) == 0 && p->num_fluids > 1) { b200_set_error("%u fluids without ENABLE_MULTIFLUID", p->num_fluids); return B200SPH_EINVAL; }
	if ((p->simflags & B200SPH_ENABLE_PLANES) && !(p->r0 > 0)) { b200_set_error("ENABLE_PLANES needs the Lennard-Jones radius r0 > 0"); return B200SPH_EINVAL; }
	if (p->viscavgop > B200SPH_AVG_GEOMETRIC || p->compvisc > B200SPH_COMPVISC_DYNAMIC) { b200_set_error("bad viscous averaging / computational viscosity"); return B200SPH_EINVAL; }
	if (!(p->slength > 0) || !(p->influenceradius > 0)) { b200_set_error("non-positive smoothing length"); return B200SPH_EINVAL; }
	return B200SPH_OK;
}

static void fill_devparams(const b200sph_params *p, DevParams *d)
{
	memset(d, 0, sizeof(*d));
	for (int a = 0; a < 3; ++a) {
		d->cellSize[a] = p->cell_size[a];
		d->gridSize[a] = (int)p->grid_size[a];
		d->coord[a] = (int)p->coord[a];
		d->gravity[a] = p->gravity[a];
	}
	d->periodic = p->periodic;
	// calcGridHash (src/cuda/cellgrid.cuh:101-106): COORD1 has stride 1, COORD2 stride G.C1, COORD3 stride G.C1*G.C2
	d->hstride[p->coord[0]] = 1;
	d->hstride[p->coord[1]] = (int)p->grid_size[p->coord[0]];
	d->hstride[p->coord[2]] = (int)(p->grid_size[p->coord[0]] * p->grid_size[p->coord[1]]);
	d->neiblistsize = p->neiblistsize; d->neibboundpos = p->neibboundpos; d->stride = p->neiblist_stride;
	d->listblock = p->neiblist_block ? p->neiblist_block : (uint32_t)B200SPH_NEIBLIST_BLOCK;
	d->nlSqInflRad = p->nl_sq_influence_radius;
	d->kerneltype = p->kerneltype; d->densitydiffusiontype = p->densitydiffusiontype; d->boundarytype = p->boundarytype;
	d->inviscid = (p->rheologytype == B200SPH_RHEOLOGY_INVISCID);
	d->turbmodel = p->turbmodel; d->compvisc = p->compvisc; d->viscavgop = p->viscavgop; d->is_const_visc = p->is_const_visc;
	d->slength = p->slength; d->influenceradius = p->influenceradius; d->deltap = p->deltap;
	// same expression and evaluation types as the reference (float h powers, double M_PI): src/cuda/forces.cu:274-291
	const float h = p->slength; const float h2 = h * h; const float h4 = h2 * h2; const float h5 = h4 * h;
	d->fcoeff_wendland = (float)(105.0f / (128.0f * M_PI * h5));
	d->densityDiffCoeff = p->density_diff_coeff; d->artvisccoeff = p->artvisccoeff; d->epsartvisc = p->epsartvisc;
	d->numFluids = p->num_fluids;
	for (uint f = 0; f < B200SPH_MAX_FLUIDS; ++f) {
		d->rho0[f] = p->rho0[f]; d->bcoeff[f] = p->bcoeff[f]; d->gammacoeff[f] = p->gammacoeff[f];
		d->sscoeff[f] = p->sscoeff[f]; d->sspowercoeff[f] = p->sspowercoeff[f]; d->visccoeff[f] = p->visccoeff[f];
		d->sqC0[f] = p->sscoeff[f] * p->sscoeff[f];   // src/cuda/forces.cu:318-323
		d->visc2coeff[f] = p->visc2coeff[f];
	}
	const float h3 = h2 * h;
	d->wcoeff_wendland = (float)(21.0f / (16.0f * M_PI * h3));      // src/cuda/forces.cu:283
	d->viscmodel = p->viscmodel; d->simflags = p->simflags;
	d->epsxsph = p->epsxsph; d->monaghanViscCoeff = p->monaghan_visc_coeff;
	d->r0 = p->r0; d->dcoeff = p->dcoeff; d->p1coeff = p->p1coeff; d->p2coeff = p->p2coeff;
	d->partsurf = p->partsurf == 0.0f ? p->r0 * p->r0 : p->partsurf;   // src/cuda/forces.cu:364-368
	d->numplanes = 0;
	d->cmd_dt = 0.0f; d->cmd_step = 0; d->dev_state = NULL;
}

extern "C" int b200sph_create(const b200sph_params *p, b200sph_ctx **out)
{
	if (!out) { b200_set_error("null out pointer"); return B200SPH_EINVAL; }
	*out = NULL;
	int rc = b200sph_validate(p);
	if (rc) return rc;
	if (b200sph_device_count() <= 0) {
		b200_set_error("no CUDA device: this library has no CPU fallback");
		return B200SPH_ENODEV;
	}
	b200sph_ctx *ctx = (b200sph_ctx *)calloc(1, sizeof(b200sph_ctx));
	if (!ctx) { b200_set_error("out of host memory"); return B200SPH_ENOMEM; }
	ctx->hp = *p;
	fill_devparams(p, &ctx->dp);
	CUDA_TRY(cudaGetDevice(&ctx->device));
	ctx->stream = 0;
	CUDA_TRY(cudaMalloc(&ctx->d_counters, sizeof(NeibsCounters)));
	CUDA_TRY(cudaMalloc(&ctx->d_scalar, 4 * sizeof(float)));
	CUDA_TRY(cudaMalloc(&ctx->d_flag, sizeof(int)));
	CUDA_TRY(cudaMallocHost(&ctx->h_scalar, 4 * sizeof(float)));
	CUDA_TRY(cudaMallocHost(&ctx->h_flag, sizeof(int)));
	CUDA_TRY(cudaMemset(ctx->d_counters, 0, sizeof(NeibsCounters)));
	CUDA_TRY(cudaMalloc(&ctx->d_bodies, sizeof(BodyData)));
	CUDA_TRY(cudaMemset(ctx->d_bodies, 0, sizeof(BodyData)));
	CUDA_TRY(cudaMallocHost(&ctx->h_bodies, sizeof(BodyData)));
	memset(ctx->h_bodies, 0, sizeof(BodyData));
	CUDA_TRY(cudaMalloc(&ctx->d_step, sizeof(StepState)));
	CUDA_TRY(cudaMemset(ctx->d_step, 0, sizeof(StepState)));
	CUDA_TRY(cudaMallocHost(&ctx->h_step, sizeof(StepState)));
	*out = ctx;
	return B200SPH_OK;
}

extern "C" int b200sph_destroy(b200sph_ctx *ctx)
{
	if (!ctx) return B200SPH_OK;
	cudaSetDevice(ctx->device);
	b200_hoststep_destroy(ctx);
	cudaFree(ctx->sort_tmp); cudaFree(ctx->keys_in); cudaFree(ctx->keys_out); cudaFree(ctx->vals_out);
	cudaFree(ctx->info_tmp); cudaFree(ctx->pv[0]); cudaFree(ctx->pv[1]);

	cudaFree(ctx->d_step); cudaFreeHost(ctx->h_step); cudaFree(ctx->d_bodies); cudaFreeHost(ctx->h_bodies);
	cudaFree(ctx->d_counters); cudaFree(ctx->d_scalar); cudaFree(ctx->d_flag);
	cudaFreeHost(ctx->h_scalar); cudaFreeHost(ctx->h_flag);
	free(ctx);
	return B200SPH_OK;
}

extern "C" int b200sph_set_stream(b200sph_ctx *ctx, void *s)
{
	if (!ctx) { b200_set_error("null context"); return B200SPH_EINVAL; }
	ctx->stream = (cudaStream_t)s;
	return B200SPH_OK;
}

extern "C" int b200sph_set_gravity(b200sph_ctx *ctx, const float g[3])
{
	if (!ctx || !g) { b200_set_error("null argument"); return B200SPH_EINVAL; }
	for (int a = 0; a < 3; ++a) { ctx->hp.gravity[a] = g[a]; ctx->dp.gravity[a] = g[a]; }
	return B200SPH_OK;
}

// AbstractForcesEngine::setplanes, src/cuda/forces.cu:443-447
extern "C" int b200sph_set_planes(b200sph_ctx *ctx, const float *normals, const int *grid_pos, const float *pos, int n)
{
	if (!ctx) { b200_set_error("null context"); return B200SPH_EINVAL; }
	if (n < 0 || n > B200SPH_MAX_PLANES) { b200_set_error("too many planes (%d > %d)", n, B200SPH_MAX_PLANES); return B200SPH_EINVAL; }
	if (n && (!normals || !grid_pos || !pos)) { b200_set_error("null plane array"); return B200SPH_EINVAL; }
	if (n && !(ctx->hp.simflags & B200SPH_ENABLE_PLANES)) { b200_set_error("planes given but ENABLE_PLANES is not set in simflags"); return B200SPH_EINVAL; }
	for (int i = 0; i < n; ++i) for (int a = 0; a < 3; ++a) {
		ctx->dp.planeNormal[i][a] = normals[3 * i + a];
		ctx->dp.planeGridPos[i][a] = grid_pos[3 * i + a];
		ctx->dp.planePos[i][a] = pos[3 * i + a];
	}
	ctx->dp.numplanes = (uint)n;
	return B200SPH_OK;
}

extern "C" int b200sph_get_neibboundpos(const b200sph_ctx *ctx, uint32_t *v)
{
	if (!ctx || !v) { b200_set_error("null argument"); return B200SPH_EINVAL; }
	*v = ctx->hp.neibboundpos;
	return B200SPH_OK;
}

// reference: src/cuda/forces.cu:540-554, 961-965
extern "C" uint32_t b200sph_fmax_elements(uint32_t n) { return (div_up(n, BLOCK_FORCES) + 3) / 4 * 4; }
// reducefmax(NULL, NULL, n): src/cuda/forces.cu:106-141 (n = number of CFL elements, a multiple of 4)
extern "C" uint32_t b200sph_fmax_temp_elements(uint32_t n)
{
	uint32_t nb = div_up(div_up(n, 4u), 256u);
	if (nb > 1) { nb = (nb + 3) / 4 * 4; if (nb > 1024u) nb = 1024u; }
	return nb;
}
extern "C" uint32_t b200sph_round_particles(uint32_t n) { return (n / BLOCK_FORCES) * BLOCK_FORCES; }

// ---- moving bodies: host copies are kept in pinned memory and pushed to the device record on the context's stream ----
static int push_bodies(b200sph_ctx *ctx)
{
	// the copy must be stream-ordered with the kernels that read it, but the pinned source is reused by the next
	// set call: synchronise (these are per-step host calls in the reference too, each a blocking symbol upload)
	CUDA_TRY(cudaMemcpyAsync(ctx->d_bodies, ctx->h_bodies, sizeof(BodyData), cudaMemcpyHostToDevice, ctx->stream));
	CUDA_TRY(cudaStreamSynchronize(ctx->stream));
	ctx->have_bodies = 1;
	return B200SPH_OK;
}
#define CHECK_BODIES(n, p) do { CHECK_CTX(ctx); if ((n) < 0 || (n) > B200SPH_MAX_BODIES) { b200_set_error("too many bodies (%d > %d)", (n), B200SPH_MAX_BODIES); return B200SPH_EINVAL; } \
	if ((n) && !(p)) { b200_set_error("null body array"); return B200SPH_EINVAL; } if ((n) == 0) return B200SPH_OK; } while (0)

// the FORCES engine's copy of the centres of gravity (torque arm, finalize_particle)
extern "C" int b200sph_set_rbcg(b200sph_ctx *ctx, const int *g, const float *c, int n)
{
	CHECK_BODIES(n, g);
	if (!c) { b200_set_error("null body array"); return B200SPH_EINVAL; }
	for (int b = 0; b < n; ++b) for (int a = 0; a < 3; ++a) { ctx->h_bodies->cgGridPos[b][a] = g[3 * b + a]; ctx->h_bodies->cgPos[b][a] = c[3 * b + a]; }
	ctx->bodies_set |= BODY_SET_CG_FORCES;
	return push_bodies(ctx);
}
// the INTEGRATION engine's copy (rigid motion, euler_update): stays at cg(n) while the forces copy moves on
extern "C" int b200sph_set_rbcg_euler(b200sph_ctx *ctx, const int *g, const float *c, int n)
{
	CHECK_BODIES(n, g);
	if (!c) { b200_set_error("null body array"); return B200SPH_EINVAL; }
	for (int b = 0; b < n; ++b) for (int a = 0; a < 3; ++a) { ctx->h_bodies->eulCgGridPos[b][a] = g[3 * b + a]; ctx->h_bodies->eulCgPos[b][a] = c[3 * b + a]; }
	ctx->bodies_set |= BODY_SET_CG_EULER;
	return push_bodies(ctx);
}
extern "C" int b200sph_set_rbstart(b200sph_ctx *ctx, const int *first, int n)
{
	CHECK_BODIES(n, first);
	for (int b = 0; b < n; ++b) ctx->h_bodies->startIndex[b] = first[b];
	ctx->bodies_set |= BODY_SET_START;
	return push_bodies(ctx);
}
extern "C" int b200sph_set_rbtrans(b200sph_ctx *ctx, const float *t, int n)
{
	CHECK_BODIES(n, t);
	for (int b = 0; b < n; ++b) for (int a = 0; a < 3; ++a) ctx->h_bodies->trans[b][a] = t[3 * b + a];
	ctx->bodies_set |= BODY_SET_TRANS;
	return push_bodies(ctx);
}
extern "C" int b200sph_set_rbsteprot(b200sph_ctx *ctx, const float *r, int n)
{
	CHECK_BODIES(n, r);
	for (int b = 0; b < n; ++b) for (int a = 0; a < 9; ++a) ctx->h_bodies->steprot[b][a] = r[9 * b + a];
	ctx->bodies_set |= BODY_SET_STEPROT;
	return push_bodies(ctx);
}
extern "C" int b200sph_set_rblinearvel(b200sph_ctx *ctx, const float *v, int n)
{
	CHECK_BODIES(n, v);
	for (int b = 0; b < n; ++b) for (int a = 0; a < 3; ++a) ctx->h_bodies->linearvel[b][a] = v[3 * b + a];
	ctx->bodies_set |= BODY_SET_LINVEL;
	return push_bodies(ctx);
}
extern "C" int b200sph_set_rbangularvel(b200sph_ctx *ctx, const float *v, int n)
{
	CHECK_BODIES(n, v);
	for (int b = 0; b < n; ++b) for (int a = 0; a < 3; ++a) ctx->h_bodies->angularvel[b][a] = v[3 * b + a];
	ctx->bodies_set |= BODY_SET_ANGVEL;
	return push_bodies(ctx);
}

// The integration moves FG_MOVING_BOUNDARY particles with the body data; it must never skip that silently. Called by
// both integration paths (stand-alone kernel, fused epilogue). Returns the device record, or NULL when no body was
// ever described (then the caller's particles must not contain moving ones: that is the contract the reference's
// GPUWorker fulfils by uploading all five arrays whenever numbodies > 0, src/GPUWorker.cc:1749, 3128-3160).
int b200_euler_bodies(b200sph_ctx *ctx, const uint32_t *hash, const BodyData **out)
{
	*out = NULL;
	const unsigned eul = ctx->bodies_set & BODY_SET_EULER_ALL;
	if (!eul) {
		if (ctx->hp.simflags & B200SPH_ENABLE_MOVING_BODIES) {
			b200_set_error("euler: ENABLE_MOVING_BODIES is set but no setrbcg/setrbtrans/setrbsteprot/setrblinearvel/setrbangularvel call was made");
			return B200SPH_EINVAL;
		}
		return B200SPH_OK;
	}
	if (eul != BODY_SET_EULER_ALL) {
		b200_set_error("euler: incomplete body description (setrb* calls made: mask 0x%x, needed 0x%x)", eul, BODY_SET_EULER_ALL);
		return B200SPH_EINVAL;
	}
	if (!hash) { b200_set_error("euler: moving bodies need the particle hash (cell of each particle)"); return B200SPH_EINVAL; }
	*out = ctx->d_bodies;
	return B200SPH_OK;
}
