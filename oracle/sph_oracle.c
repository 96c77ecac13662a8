/*
 * sph_oracle.c — CPU restatement of GPUSPH's WCSPH per-timestep hot path.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing in gpusph_b200/ may import, link or call
 * this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and only as the checker / the CPU baseline.
 *
 * Every function restates one reference device function in plain scalar C and
 * cites the file:line it follows (paths relative to the GPUSPH tree). The
 * restatement keeps the reference's control flow, data layouts AND summation
 * order (three separate passes fluid<-fluid, fluid<-boundary, boundary<-fluid
 * accumulating into the force array in neighbour-list order).
 *
 * Parity status: integer/index outputs (hash, sort permutation, cellStart/End,
 * neighbour list) are pinned against the shim-built reference binary
 * (oracle/_ref, see oracle/build_ref.sh) on the GPU box — fixtures under
 * tests/golden/ — because the reference ships no golden vectors of its own
 * (SURVEY.md §8c). Float outputs are compared to a stated tolerance: the
 * reference evaluates the equation of state with the GPU's approximate
 * __powf (src/cuda/phys_core.cu:99-136), which a CPU cannot reproduce
 * bit-for-bit; callers may inject device-evaluated P/rho^2 and sound speed
 * (arrays eos_p, eos_c) to remove that source of difference.
 *
 * Floating point: compiled with -ffp-contract=off; the few places where nvcc's
 * default FMA contraction of the reference expression decides an INTEGER
 * result (cell migration, neighbour distance test) use explicit fmaf() in the
 * contraction order nvcc applies (a*b + c*d + e*f -> fma(e,f, fma(c,d, a*b)),
 * x - i*s -> fma(-i, s, x)).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include "../include/b200sph.h"

typedef struct { float x, y, z, w; } f4;
typedef struct { uint16_t x, y, z, w; } us4;
typedef struct { int x, y, z; } i3;

#define CELLTYPE_BITMASK (~(3U << 30))      /* src/multi_gpu_defines.h:79 */
#define CELL_HASH_MAX 0xFFFFFFFFu           /* src/hashkey.h:44 */
#define CELL_EMPTY 0xFFFFFFFFu              /* src/common_types.h:72 */
#define NEIBS_END 0xFFFFu                   /* src/common_types.h:71 */
#define CELLNUM_SHIFT 11                    /* src/common_types.h:63-68 */
#define CELLNUM_ENCODED (1U << CELLNUM_SHIFT)
#define NEIBINDEX_MASK (CELLNUM_ENCODED - 1)

/* src/particleinfo.h:135-330 */
static inline int ptype(us4 i) { return i.x & 7; }
static inline uint32_t pid(us4 i) { return (uint32_t)i.z | ((uint32_t)i.w << 16); }
static inline int fluid_num(us4 i) { return i.y >> 12; }
static inline int is_fluid(us4 i) { return ptype(i) == B200SPH_PT_FLUID; }
static inline int is_boundary(us4 i) { return ptype(i) == B200SPH_PT_BOUNDARY; }
static inline int is_testpoint(us4 i) { return ptype(i) == B200SPH_PT_TESTPOINT; }
static inline int is_moving(us4 i) { return i.x & B200SPH_FG_MOVING_BOUNDARY; }
static inline int is_surface(us4 i) { return i.x & B200SPH_FG_SURFACE; }
static inline int compute_force(us4 i) { return i.x & B200SPH_FG_COMPUTE_FORCE; }
static inline int inactive(f4 p) { return !isfinite(p.w); }

/* src/cuda/cellgrid.cuh:101-106 */
static inline uint32_t calc_grid_hash(const b200sph_params *P, i3 g)
{
	const int gp[3] = { g.x, g.y, g.z };
	const uint32_t *G = P->grid_size, *c = P->coord;
	return (uint32_t)(gp[c[2]] * (int)G[c[1]] * (int)G[c[0]] + gp[c[1]] * (int)G[c[0]] + gp[c[0]]);
}
/* src/cuda/cellgrid.cuh:117-128 */
static inline i3 grid_pos_from_hash(const b200sph_params *P, uint32_t cellHash)
{
	const uint32_t *G = P->grid_size, *c = P->coord;
	int gp[3];
	int temp = (int)(G[c[1]] * G[c[0]]);
	gp[c[2]] = (int)cellHash / temp;
	temp = (int)cellHash - gp[c[2]] * temp;
	gp[c[1]] = temp / (int)G[c[0]];
	gp[c[0]] = temp - gp[c[1]] * (int)G[c[0]];
	i3 r = { gp[0], gp[1], gp[2] };
	return r;
}
/* src/cuda/cellgrid.cuh:174-185 */
static inline uint32_t calc_grid_hash_periodic(const b200sph_params *P, i3 g)
{
	const int Gx = (int)P->grid_size[0], Gy = (int)P->grid_size[1], Gz = (int)P->grid_size[2];
	if (g.x < 0) g.x = Gx - 1;
	if (g.x >= Gx) g.x = 0;
	if (g.y < 0) g.y = Gy - 1;
	if (g.y >= Gy) g.y = 0;
	if (g.z < 0) g.z = Gz - 1;
	if (g.z >= Gz) g.z = 0;
	return calc_grid_hash(P, g);
}

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

/* ------------------------------------------------------------------------ */
/* calcHash: src/cuda/buildneibs_kernel.cu:664-776 (+ clampGridPos :231-305)  */
/* ------------------------------------------------------------------------ */
void oracle_calc_hash(const b200sph_params *P, f4 *pos, uint32_t *hash, uint32_t *part_index,
	const us4 *info, const uint32_t *cdm, uint32_t n)
{
	const int G[3] = { (int)P->grid_size[0], (int)P->grid_size[1], (int)P->grid_size[2] };
	for (uint32_t i = 0; i < n; ++i) {
		const us4 inf = info[i];
		uint32_t gridHash = hash[i] & CELLTYPE_BITMASK;
		if (is_fluid(inf) || is_moving(inf) || (is_surface(inf) && !is_fluid(inf))) {
			f4 p = pos[i];
			const i3 gp = grid_pos_from_hash(P, gridHash);
			const int gpa[3] = { gp.x, gp.y, gp.z };
			float pa[3] = { p.x, p.y, p.z };
			int off[3], ng[3];
			int toofar = 0;
			for (int a = 0; a < 3; ++a) {
				/* :721-725 different half constants for negative / positive pos */
				const float half = pa[a] < 0 ? 0.5f : 0.49999997f;
				off[a] = (int)floorf(pa[a] / P->cell_size[a] + half);
				ng[a] = gpa[a] + off[a];
				if (P->periodic & (1u << a)) {          /* :248-256 */
					if (ng[a] < 0) ng[a] += G[a];
					if (ng[a] >= G[a]) ng[a] -= G[a];
				} else {                                  /* :257-262, :288-299 */
					ng[a] = imin(imax(0, ng[a]), G[a] - 1);
					if (abs(off[a]) > 1 && ng[a] == gpa[a]) toofar = 1;
					off[a] = ng[a] - gpa[a];
				}
				/* :750 as_float3(pos) -= gridOffset*d_cellSize  (nvcc contracts to FMA) */
				pa[a] = fmaf(-(float)off[a], P->cell_size[a], pa[a]);
			}
			i3 ngp = { ng[0], ng[1], ng[2] };
			gridHash = calc_grid_hash(P, ngp);
			p.x = pa[0]; p.y = pa[1]; p.z = pa[2];
			if (toofar) p.w = NAN;                         /* :754-755 disable_particle */
			if (inactive(p)) gridHash = CELL_HASH_MAX;     /* :759-760 */
			pos[i] = p;
		}
		if (cdm && gridHash != CELL_HASH_MAX) gridHash |= cdm[gridHash]; /* :768-769 */
		hash[i] = gridHash;
		part_index[i] = i;
	}
}

/* fixHash: src/cuda/buildneibs_kernel.cu:790-814 */
void oracle_fix_hash(const b200sph_params *P, uint32_t *hash, uint32_t *part_index,
	const us4 *info, const uint32_t *cdm, uint32_t n)
{
	(void)P; (void)info;
	for (uint32_t i = 0; i < n; ++i) {
		if (hash) {
			const uint32_t gridHash = hash[i] & CELLTYPE_BITMASK;
			if (cdm) hash[i] = hash[i] | cdm[gridHash];
		}
		part_index[i] = i;
	}
}

/* ------------------------------------------------------------------------ */
/* sort: src/cuda/buildneibs.cu:358-415 (ptype_hash_compare)                  */
/* ------------------------------------------------------------------------ */
typedef struct { uint32_t hash; us4 info; uint32_t idx; } sort_rec;
static int sort_cmp(const void *pa, const void *pb)
{
	const sort_rec *a = (const sort_rec *)pa, *b = (const sort_rec *)pb;
	if (a->hash != b->hash) return a->hash < b->hash ? -1 : 1;
	const int ta = ptype(a->info), tb = ptype(b->info);
	if (ta != tb) return ta < tb ? -1 : 1;
	const uint32_t ia = pid(a->info), ib = pid(b->info);
	if (ia != ib) return ia < ib ? -1 : 1;
	return 0;
}
void oracle_sort(uint32_t *hash, us4 *info, uint32_t *part_index, uint32_t n)
{
	sort_rec *r = (sort_rec *)malloc(sizeof(sort_rec) * (n ? n : 1));
	for (uint32_t i = 0; i < n; ++i) { r[i].hash = hash[i]; r[i].info = info[i]; r[i].idx = part_index[i]; }
	qsort(r, n, sizeof(sort_rec), sort_cmp);
	for (uint32_t i = 0; i < n; ++i) { hash[i] = r[i].hash; info[i] = r[i].info; part_index[i] = r[i].idx; }
	free(r);
}

/* ------------------------------------------------------------------------ */
/* reorderDataAndFindCellStart: src/cuda/buildneibs_kernel.cu:840-992         */
/* cell_start / cell_end must be pre-filled with 0xFF by the caller.          */
/* ------------------------------------------------------------------------ */
void oracle_reorder(uint32_t *cell_start, uint32_t *cell_end, uint32_t *segment_start,
	f4 *sorted_pos, f4 *sorted_vel, const f4 *unsorted_pos, const f4 *unsorted_vel,
	const us4 *sorted_info, const uint32_t *sorted_hash, const uint32_t *part_index,
	uint32_t n, uint32_t *new_num_particles)
{
	(void)sorted_info;
	if (segment_start) for (int s = 0; s < 4; ++s) segment_start[s] = 0xFFFFFFFFu;
	for (uint32_t i = 0; i < n; ++i) {
		const uint32_t cellHash = sorted_hash[i];
		const uint32_t prev = i ? sorted_hash[i - 1] : 0;
		if (i == 0 || cellHash != prev) {
			if (cellHash != CELL_HASH_MAX) cell_start[cellHash & CELLTYPE_BITMASK] = i;
			else *new_num_particles = i;
			if (i > 0) cell_end[prev & CELLTYPE_BITMASK] = i;
		}
		if (cellHash == CELL_HASH_MAX) continue;
		if (i == n - 1) {
			cell_end[cellHash & CELLTYPE_BITMASK] = i + 1;
			*new_num_particles = n;
		}
		if (segment_start) {
			const unsigned ct = cellHash >> 30, pt = prev >> 30;
			if (i == 0 || ct != pt) segment_start[ct] = i;
		}
		const uint32_t s = part_index[i];
		sorted_pos[i] = unsorted_pos[s];
		sorted_vel[i] = unsorted_vel[s];
	}
}

/* ------------------------------------------------------------------------ */
/* buildNeibsList: src/cuda/buildneibs_kernel.cu:1029-1185, neibsInCell       */
/* :538-643, calcNeibCell :317-384, neibListOffset :466-478, too_many_neibs   */
/* :491-515. neibs_list must be pre-filled with 0xFF by the caller.           */
/* ------------------------------------------------------------------------ */
typedef struct {
	int32_t num_interactions, max_fluid_boundary_neibs, max_vertex_neibs, has_too_many_neibs;
	int32_t has_max_neibs[3];
} oracle_neibs_info;

static inline uint32_t neib_list_offset(const b200sph_params *P, uint32_t num, int type)
{
	return type == B200SPH_PT_FLUID ? num :
		type == B200SPH_PT_BOUNDARY ? P->neibboundpos - num : num + P->neibboundpos + 1;
}
static inline int too_many_neibs(const b200sph_params *P, const uint32_t *nn, int type)
{
	switch (type) {
	case B200SPH_PT_FLUID: return !(nn[0] < P->neibboundpos);
	case B200SPH_PT_BOUNDARY: return !(nn[0] + nn[1] < P->neibboundpos);
	case B200SPH_PT_VERTEX: return !(nn[2] < P->neiblistsize - P->neibboundpos - 1);
	default: return 1;
	}
}

void oracle_build_neibs(const b200sph_params *P, const f4 *pos, const us4 *info, const uint32_t *hash,
	const uint32_t *cell_start, const uint32_t *cell_end, uint16_t *neibs_list,
	uint32_t num_particles, uint32_t range_end, oracle_neibs_info *out)
{
	const size_t stride = P->neiblist_stride;
	const int G[3] = { (int)P->grid_size[0], (int)P->grid_size[1], (int)P->grid_size[2] };
	long long total = 0; int maxfb = 0;
	int has_too_many = -1; int has_max[3] = { 0, 0, 0 };
	(void)num_particles;
#pragma omp parallel for schedule(dynamic, 1024) reduction(+:total) reduction(max:maxfb)
	for (uint32_t index = 0; index < range_end; ++index) {
		uint32_t nn[3] = { 0, 0, 0 };
		const us4 inf = info[index];
		const f4 p = pos[index];
		/* :1060-1074 which particles get a list (DYN: all; else fluid/testpoint/floating/compute-force) */
		int build_nl = is_fluid(inf) || is_testpoint(inf) ||
			(inf.x & (B200SPH_FG_MOVING_BOUNDARY | B200SPH_FG_COMPUTE_FORCE));
		if (P->boundarytype == B200SPH_DYN_BOUNDARY) build_nl = 1;
		if (build_nl && !inactive(p)) {
			const i3 gp = grid_pos_from_hash(P, hash[index] & CELLTYPE_BITMASK);
			const int boundary = is_boundary(inf);
			for (int z = -1; z <= 1; ++z) for (int y = -1; y <= 1; ++y) for (int x = -1; x <= 1; ++x) {
				const int cell = (x + 1) + (y + 1) * 3 + (z + 1) * 9;
				int g[3] = { gp.x + x, gp.y + y, gp.z + z };
				/* calcNeibCell :317-384 */
				int inside = 1;
				for (int a = 0; a < 3; ++a) {
					if (g[a] < 0) { if (P->periodic & (1u << a)) g[a] = G[a] - 1; else inside = 0; }
					else if (g[a] >= G[a]) { if (P->periodic & (1u << a)) g[a] = 0; else inside = 0; }
				}
				if (!inside) continue;
				i3 ng = { g[0], g[1], g[2] };
				const uint32_t gh = calc_grid_hash(P, ng);
				const uint32_t bs = cell_start[gh], be = cell_end[gh];
				if (bs == CELL_EMPTY) continue;
				/* :569 pos -= gridOffset*d_cellSize (FMA-contracted by nvcc) */
				const float px = fmaf(-(float)x, P->cell_size[0], p.x);
				const float py = fmaf(-(float)y, P->cell_size[1], p.y);
				const float pz = fmaf(-(float)z, P->cell_size[2], p.z);
				int encode_cell = 1;
				int neib_type = B200SPH_PT_FLUID;
				for (uint32_t j = bs; j < be; ++j) {
					if (j == index) continue;
					const us4 ninf = info[j];
					if (is_testpoint(ninf)) continue;
					if (!encode_cell && neib_type != ptype(ninf)) encode_cell = 1;
					neib_type = ptype(ninf);
					if (P->boundarytype == B200SPH_LJ_BOUNDARY && boundary && is_boundary(ninf)) continue;
					if (P->boundarytype == B200SPH_DYN_BOUNDARY && boundary && is_boundary(ninf)) continue;
					const f4 np = pos[j];
					if (inactive(np)) continue;
					const float rx = px - np.x, ry = py - np.y, rz = pz - np.z;
					/* sqlength = x*x + y*y + z*z, contracted (src/vector_math.h:560-575) */
					const float r2 = fmaf(rz, rz, fmaf(ry, ry, rx * rx));
					if (r2 < P->nl_sq_influence_radius) {
						const uint32_t offset = neib_list_offset(P, nn[neib_type], neib_type);
						nn[neib_type]++;
						if (!too_many_neibs(P, nn, neib_type)) {
							const int enc = encode_cell ? ((cell + 1) << CELLNUM_SHIFT) : 0;
							neibs_list[offset * stride + index] = (uint16_t)((j - bs) + enc);
							encode_cell = 0;
						}
					}
				}
			}
		}
		/* :1108-1137 end markers and overflow */
		int overflow = too_many_neibs(P, nn, B200SPH_PT_FLUID);
		const uint32_t marker = overflow ? P->neibboundpos : nn[0];
		neibs_list[marker * stride + index] = NEIBS_END;
		overflow |= too_many_neibs(P, nn, B200SPH_PT_BOUNDARY);
		if (!overflow)
			neibs_list[neib_list_offset(P, nn[1], B200SPH_PT_BOUNDARY) * stride + index] = NEIBS_END;
		if (overflow) {
#pragma omp critical
			{
				if (has_too_many == -1 || (int)pid(inf) < has_too_many) {
					has_too_many = (int)pid(inf);
					has_max[0] = nn[0]; has_max[1] = nn[1]; has_max[2] = nn[2];
				}
			}
		}
		/* :1140-1185 counters */
		const int fb = (int)(nn[0] + nn[1]);
		if (fb > maxfb) maxfb = fb;
		total += fb + nn[2];
	}
	if (out) {
		out->num_interactions = (int32_t)total;
		out->max_fluid_boundary_neibs = maxfb;
		out->max_vertex_neibs = 0;
		out->has_too_many_neibs = has_too_many;
		memcpy(out->has_max_neibs, has_max, sizeof(has_max));
	}
}

/* ------------------------------------------------------------------------ */
/* physics: src/cuda/phys_core.cu:99-151, src/cuda/sph_core.cu:106-181,       */
/* src/cuda/visc_kernel.cu:75-85, src/cuda/visc_avg.cu:40-161                 */
/* ------------------------------------------------------------------------ */
static inline float eos_P(const b200sph_params *P, float rho_tilde, int f)
{ return P->bcoeff[f] * (powf(rho_tilde + 1.0f, P->gammacoeff[f]) - 1.0f); }
static inline float eos_c(const b200sph_params *P, float rho_tilde, int f)
{ return P->sscoeff[f] * powf(rho_tilde + 1.0f, P->sspowercoeff[f]); }
static inline float phys_rho(const b200sph_params *P, float rho_tilde, int f)
{ return (rho_tilde + 1.0f) * P->rho0[f]; }

static float kernel_fcoeff(const b200sph_params *P)
{	/* src/cuda/forces.cu:276-291 */
	const float h = P->slength; const float h2 = h * h; const float h4 = h2 * h2; const float h5 = h4 * h;
	return (float)(105.0f / (128.0f * M_PI * h5));
}

static inline float visc_avg_density(const b200sph_params *P, float rho, float nrho, float nmass)
{	/* src/cuda/visc_avg.cu density-only operators */
	switch (P->viscavgop) {
	case B200SPH_AVG_ARITHMETIC: return nmass * (rho + nrho) / (rho * nrho);
	case B200SPH_AVG_HARMONIC: return 4 * nmass / (rho + nrho);
	default: return 2 * nmass * (1.0f / sqrtf(rho * nrho));
	}
}
static inline float visc_avg_dyn(const b200sph_params *P, float v, float nv, float rho, float nrho, float nmass)
{	/* src/cuda/visc_avg.cu:47-112 non-constant dynamic viscosity */
	switch (P->viscavgop) {
	case B200SPH_AVG_ARITHMETIC: return nmass * (v + nv) / (rho * nrho);
	case B200SPH_AVG_HARMONIC: return 4 * nmass * (v * nv) / (v + nv) / (rho * nrho);
	default: return 2 * nmass * sqrtf(v * nv) / (rho * nrho);
	}
}

/* ------------------------------------------------------------------------ */
/* forces: forcesDevice src/cuda/forces_kernel.def:3923-4027 run three times  */
/* (src/cuda/forces.cu:759,782,792) + finalizeforcesDevice :4037-4153.        */
/* Optional eos_p / eos_c: per-particle P/rho^2 and sound speed evaluated     */
/* elsewhere (e.g. on the device with __powf); NULL = evaluate here (powf).   */
/* Optional abssum: per particle sum of |pair contribution| (xyz: max over     */
/* components, w) — the natural scale for comparing float sums.               */
/* ------------------------------------------------------------------------ */
typedef struct { float x, y, z; } f3;

/* arguments only some option combinations read (b200sph_forces_args / b200sph_set_planes on the product side) */
typedef struct {
	float dt;                 /* the command's dt: BREZZI diffusion (forces_kernel.def:1765-1782) */
	f4 *xsph;                 /* ENABLE_XSPH output, zeroed by the caller (forces_kernel.def:3366-3368) */
	int numplanes;            /* geometric planes, plane_t layout (src/planes.h:42-46) */
	float plane_normal[B200SPH_MAX_PLANES][3];
	int plane_gridpos[B200SPH_MAX_PLANES][3];
	float plane_pos[B200SPH_MAX_PLANES][3];
} oracle_forces_opts;

/* average<avgop>, src/average.h:75-100 */
static inline float average_op(unsigned op, float a, float b)
{
	switch (op) {
	case B200SPH_AVG_ARITHMETIC: return (a + b) * 0.5f;
	case B200SPH_AVG_HARMONIC: return 2.0f * a * b / (a + b);
	default: return sqrtf(a * b);
	}
}

/* W<WENDLAND>, src/cuda/sph_core.cu:104-117; coefficient src/cuda/forces.cu:274-284 */
static inline float kernel_wcoeff(const b200sph_params *P)
{
	const float h = P->slength; const float h2 = h * h; const float h3 = h2 * h;
	return (float)(21.0f / (16.0f * M_PI * h3));
}
static inline float wendland_W(float r, float h, float wcoeff)
{
	const float R = r / h;
	float val = 1.0f - 0.5f * R;
	val *= val;
	val *= val;
	val *= 1.0f + 2.0f * R;
	val *= wcoeff;
	return val;
}

static void forces_pass(const b200sph_params *P, const oracle_forces_opts *opts, int cptype, int nptype,
	const f4 *pos, const f4 *vel, const us4 *info, const uint32_t *hash,
	const uint32_t *cell_start, const uint16_t *neibs_list,
	const float *pprec, const float *ssp,
	f4 *forces, f4 *abssum, uint32_t from, uint32_t to)
{
	const size_t stride = P->neiblist_stride;
	const float fcoeff = kernel_fcoeff(P);
	const float wcoeff = kernel_wcoeff(P);
	const float h = P->slength;
	const int inviscid = P->rheologytype == B200SPH_RHEOLOGY_INVISCID;
	/* computes_xsph :176-189: ENABLE_XSPH, fluid central, fluid neighbour */
	const int xsph = (P->simflags & B200SPH_ENABLE_XSPH) && cptype == B200SPH_PT_FLUID && nptype == B200SPH_PT_FLUID && opts && opts->xsph;
#pragma omp parallel for schedule(dynamic, 512)
	for (uint32_t index = from; index < to; ++index) {
		const us4 inf = info[index];
		if (ptype(inf) != cptype) continue;
		const f4 p = pos[index];
		if (inactive(p)) continue;
		const f4 v = vel[index];
		const int fnum = fluid_num(inf);
		const i3 gp = grid_pos_from_hash(P, hash[index] & CELLTYPE_BITMASK);
		const float rho = phys_rho(P, v.w, fnum);
		const float p_precalc = pprec[index];
		const float sspeed = ssp[index];
		f4 force = forces[index];             /* common_particle_output: RMW, :886-897 */
		f4 asum = abssum ? abssum[index] : (f4){ 0, 0, 0, 0 };
		f3 mean_vel = { 0, 0, 0 };            /* xsph_particle_output :962-969 */

		/* neighbour-list traversal: src/cuda/neibs_iteration.cuh:56-200, src/cuda/cellgrid.cuh:198-226 */
		float pcx = 0, pcy = 0, pcz = 0;
		uint32_t base = 0;
		long long slot = (nptype == B200SPH_PT_FLUID) ? 0 : (long long)P->neibboundpos;
		const long long step = (nptype == B200SPH_PT_BOUNDARY) ? -1 : 1;
		for (;; slot += step) {
			uint32_t nd = neibs_list[(size_t)slot * stride + index];
			if (nd == NEIBS_END) break;
			if (nd >= CELLNUM_ENCODED) {
				const int cell = (int)(nd >> CELLNUM_SHIFT) - 1;
				nd &= NEIBINDEX_MASK;
				const int ox = cell % 3 - 1, oy = (cell / 3) % 3 - 1, oz = cell / 9 - 1;
				pcx = p.x - (float)ox * P->cell_size[0];
				pcy = p.y - (float)oy * P->cell_size[1];
				pcz = p.z - (float)oz * P->cell_size[2];
				i3 ng = { gp.x + ox, gp.y + oy, gp.z + oz };
				base = cell_start[calc_grid_hash_periodic(P, ng)];
			}
			const uint32_t j = base + nd;
			const f4 np = pos[j];
			const float rx = pcx - np.x, ry = pcy - np.y, rz = pcz - np.z;
			const float nmass = np.w;
			if (!isfinite(nmass)) continue;
			const float r = sqrtf(rx * rx + ry * ry + rz * rz);
			if (r >= P->influenceradius) continue;
			const us4 ninf = info[j];
			const f4 nv = vel[j];
			const int nfnum = fluid_num(ninf);
			/* common_neib_data :1099-1130 */
			const float rvx = v.x - nv.x, rvy = v.y - nv.y, rvz = v.z - nv.z;
			const float nrho_t = nv.w;
			const float vel_dot_pos = rvx * rx + rvy * ry + rvz * rz;
			const float qm2 = r / h - 2.0f;
			const float f = qm2 * qm2 * qm2 * fcoeff;           /* F<WENDLAND> sph_core.cu:168-174 */
			const float nsspeed = ssp[j];
			const float np_precalc = pprec[j];
			const float nrho = phys_rho(P, nrho_t, nfnum);

			float DrDt = 0, dvx = 0, dvy = 0, dvz = 0;
			float a_w = 0, a_v = 0;
			const int all = (cptype == B200SPH_PT_FLUID);       /* FF and F<-B(DYN): compute_all_pp_interaction :3568-3598,3717-3726 */

			/* compute_density_derivative :2178-2190 */
			DrDt = nmass * vel_dot_pos * f;                      /* mass_continuity_div_vel_term :2140-2150 */
			a_w += fabsf(DrDt);
			if (nptype == B200SPH_PT_FLUID) {                   /* no diffusion from DYN boundary neighbours :1594-1606 */
				if (P->densitydiffusiontype == B200SPH_RHODIFF_FERRARI) {          /* :1614-1636 */
					const float grav_corr = -(P->gravity[0] * rx + P->gravity[1] * ry + P->gravity[2] * rz) *
						P->rho0[fnum] / (P->sscoeff[fnum] * P->sscoeff[fnum]);
					float fx = 0, fy = 0, fz = 0;
					if (r > 1e-4f * h) {
						const float s = fmaxf(sspeed, nsspeed) * (rho - nrho + grav_corr) / rho / r;
						fx = s * rx; fy = s * ry; fz = s * rz;
					}
					const float t = P->density_diff_coeff * nmass * (fx * rx + fy * ry + fz * rz) * f;
					DrDt += t; a_w += fabsf(t);
				} else if (P->densitydiffusiontype == B200SPH_RHODIFF_COLAGROSSI) { /* :1916-1951 */
					if (fnum == nfnum) {
						/* the switch compares the pressures P() themselves (:1925-1928), not P/rho^2 scaled back */
						const float Pi = eos_P(P, v.w, fnum), Pj = eos_P(P, nrho_t, nfnum);
						const float gdot = P->gravity[0] * rx + P->gravity[1] * ry + P->gravity[2] * rz;
						if (!(fabsf(Pi - Pj) < fabsf(gdot * rho))) {
							const float t = P->density_diff_coeff * P->sscoeff[fnum] * (nrho / rho - 1) * f * nmass;
							DrDt -= t; a_w += fabsf(t);
						}
					}
				} else if (P->densitydiffusiontype == B200SPH_RHODIFF_BREZZI) {     /* :1765-1782 */
					const float Pi = eos_P(P, v.w, fnum), Pj = eos_P(P, nrho_t, nfnum);
					const float gdot = P->gravity[0] * rx + P->gravity[1] * ry + P->gravity[2] * rz;
					const float dt = opts ? opts->dt : 0.0f;
					const float t = P->density_diff_coeff * ((2.0f / (rho + nrho)) * (Pi - Pj) - gdot) * nmass / nrho * f * dt * 2.0f * rho;
					DrDt += t; a_w += fabsf(t);
				}
			}
			force.w += DrDt;

			const int momentum = all || compute_force(inf);      /* B<-F DYN: :3634-3667 */
			if (momentum) {
				/* compute_pressure_contrib general :2450-2466 */
				const float pg = (p_precalc + np_precalc) * nmass * f;
				dvx -= pg * rx; dvy -= pg * ry; dvz -= pg * rz;
				a_v += fabsf(pg) * r;
				/* ARTIFICIAL viscosity :2744-2764, artvisc visc_kernel.cu:75-85 */
				if (P->turbmodel == B200SPH_TURB_ARTIFICIAL && vel_dot_pos < 0.0f) {
					const float visc = vel_dot_pos * h * P->artvisccoeff * (sspeed + nsspeed) /
						((r * r + P->epsartvisc) * (rho + nrho));
					const float s = visc * nmass * f;
					dvx += visc * rx * nmass * f; dvy += visc * ry * nmass * f; dvz += visc * rz * nmass * f;
					a_v += fabsf(s) * r;
				}
				/* Espanol & Revenga :2651-2678 */
				if (!inviscid && P->viscmodel == B200SPH_VISCMODEL_ESPANOL_REVENGA) {
					const float vc = P->visccoeff[fnum], nvc = P->visccoeff[nfnum];
					const float pvisc = P->compvisc == B200SPH_COMPVISC_KINEMATIC ? vc * rho : vc;      /* get_dynamic_visc :276-289 */
					const float nvisc = P->compvisc == B200SPH_COMPVISC_KINEMATIC ? nvc * nrho : nvc;
					const float visc_thirds = average_op(P->viscavgop, pvisc, nvisc) / 3;
					const float bulk = average_op(P->viscavgop, P->visc2coeff[fnum], P->visc2coeff[nfnum]);
					const float coeff = nmass / (rho * nrho) * f;                                        /* :2575-2580 */
					const float pos_den = (rx * rx + ry * ry + rz * rz) + P->epsartvisc;
					const float a = 5 * visc_thirds - bulk, b = 5 * (visc_thirds + bulk) * vel_dot_pos / pos_den;
					dvx += coeff * (a * rvx + b * rx); dvy += coeff * (a * rvy + b * ry); dvz += coeff * (a * rvz + b * rz);
					a_v += fabsf(coeff) * (fabsf(a) * sqrtf(rvx * rvx + rvy * rvy + rvz * rvz) + fabsf(b) * r);
				} else
				/* laminar Morris / Monaghan :2605-2625 */
				if (!inviscid) {
					float visc;
					const float vc = P->visccoeff[fnum], nvc = P->visccoeff[nfnum];
					if (P->compvisc == B200SPH_COMPVISC_KINEMATIC) {
						if (P->is_const_visc) visc = vc * visc_avg_density(P, rho, nrho, nmass);
						else visc = visc_avg_dyn(P, vc * rho, nvc * nrho, rho, nrho, nmass);
					} else {
						if (P->is_const_visc) visc = 2 * nmass * vc / (rho * nrho);
						else visc = visc_avg_dyn(P, vc, nvc, rho, nrho, nmass);
					}
					const float s = visc * f;
					if (P->viscmodel == B200SPH_VISCMODEL_MONAGHAN) {       /* viscous_vector_component<MONAGHAN> :2534-2559 */
						const float den = (rx * rx + ry * ry + rz * rz) + P->epsartvisc;
						const float m = vel_dot_pos < 0 ? P->monaghan_visc_coeff * vel_dot_pos / den : 0.0f;
						dvx += s * (m * rx); dvy += s * (m * ry); dvz += s * (m * rz);
						a_v += fabsf(s * m) * r;
					} else {
						dvx += s * rvx; dvy += s * rvy; dvz += s * rvz;
						a_v += fabsf(s) * sqrtf(rvx * rvx + rvy * rvy + rvz * rvz);
					}
				}
				/* compute_mean_vel :2986-2992 */
				if (xsph) {
					const float s = nmass * wendland_W(r, h, wcoeff) / (rho + nrho);
					mean_vel.x -= s * rvx; mean_vel.y -= s * rvy; mean_vel.z -= s * rvz;
				}
				force.x += dvx; force.y += dvy; force.z += dvz;
			}
			asum.w += a_w; asum.x += a_v;
		}
		forces[index] = force;
		if (abssum) abssum[index] = asum;
		if (xsph) opts->xsph[index] = (f4){ 2.0f * mean_vel.x, 2.0f * mean_vel.y, 2.0f * mean_vel.z, 0.0f };   /* write_xsph :3366-3368 */
	}
}

/* moving / force-feedback bodies: host mirror of the reference's __constant__ arrays
 * (src/cuda/forces_kernel.cu:81-83, src/cuda/euler_kernel.cu:45-50) */
typedef struct {
	int cgGridPos[B200SPH_MAX_BODIES][3];
	float cgPos[B200SPH_MAX_BODIES][3];
	int startIndex[B200SPH_MAX_BODIES];
	float trans[B200SPH_MAX_BODIES][3];
	float steprot[B200SPH_MAX_BODIES][9];
	float linearvel[B200SPH_MAX_BODIES][3];
	float angularvel[B200SPH_MAX_BODIES][3];
} oracle_bodies;

/* forces + finalize. Returns the number of CFL blocks written, like
 * CUDAForcesEngine::basicstep (src/cuda/forces.cu:901-932). forces must be zeroed by the caller. */
uint32_t oracle_forces_ex(const b200sph_params *P, const f4 *pos, const f4 *vel, const us4 *info,
	const uint32_t *hash, const uint32_t *cell_start, const uint16_t *neibs_list,
	const float *eos_p_in, const float *eos_c_in,
	f4 *forces, float *cfl, f4 *abssum,
	uint32_t num_particles, uint32_t from, uint32_t to, uint32_t cfl_offset,
	const oracle_bodies *bodies, f4 *rb_forces, f4 *rb_torques, const oracle_forces_opts *opts)
{
	float *pprec = (float *)malloc(sizeof(float) * (num_particles ? num_particles : 1));
	float *ssp = (float *)malloc(sizeof(float) * (num_particles ? num_particles : 1));
#pragma omp parallel for
	for (uint32_t i = 0; i < num_particles; ++i) {
		const int f = fluid_num(info[i]);
		const float rho = phys_rho(P, vel[i].w, f);
		/* precalc_pressure SPH_F1 :419-429 */
		pprec[i] = eos_p_in ? eos_p_in[i] : eos_P(P, vel[i].w, f) / (rho * rho);
		ssp[i] = eos_c_in ? eos_c_in[i] : eos_c(P, vel[i].w, f);
	}
	/* src/cuda/forces.cu:759,782,792: fluid<-fluid, fluid<-boundary, boundary<-fluid */
	forces_pass(P, opts, B200SPH_PT_FLUID, B200SPH_PT_FLUID, pos, vel, info, hash, cell_start, neibs_list, pprec, ssp, forces, abssum, from, to);
	forces_pass(P, opts, B200SPH_PT_FLUID, B200SPH_PT_BOUNDARY, pos, vel, info, hash, cell_start, neibs_list, pprec, ssp, forces, abssum, from, to);
	if (P->boundarytype == B200SPH_DYN_BOUNDARY)
		forces_pass(P, opts, B200SPH_PT_BOUNDARY, B200SPH_PT_FLUID, pos, vel, info, hash, cell_start, neibs_list, pprec, ssp, forces, abssum, from, to);

	/* finalizeforcesDevice :4037-4153 with forces_fixup :3212-3219, gravity :4091,
	 * dyndt_forces_shared_data :3436-3456, maxBlockReduce device_core.cu:40-59.
	 * Grid = div_up(to-from, 128) rounded up to a multiple of 4 (forces.cu:741-744). */
	const uint32_t BLOCK = 128;
	const uint32_t n = to - from;
	uint32_t nblocks = (n + BLOCK - 1) / BLOCK;
	nblocks = (nblocks + 3) / 4 * 4;
	for (uint32_t b = 0; b < nblocks; ++b) {
		float m = 0.0f;
		for (uint32_t t = 0; t < BLOCK; ++t) {
			const uint32_t index = from + b * BLOCK + t;
			if (index >= to) break;
			const f4 p = pos[index];
			if (inactive(p)) continue;
			const us4 inf = info[index];
			const int f = fluid_num(inf);
			f4 fo = forces[index];
			fo.w /= P->rho0[f];
			if (is_fluid(inf)) {
				fo.x += P->gravity[0]; fo.y += P->gravity[1]; fo.z += P->gravity[2];
				/* geometric planes :4105-4110: GeometryForce / PlaneForce / LJForce src/cuda/forces_kernel.cu:94-204,
				 * PlaneDistance src/cuda/geom_core.cu:65-85, viscous_plane_coefficient :3103-3113 */
				if ((P->simflags & B200SPH_ENABLE_PLANES) && opts && opts->numplanes) {
					const float rho = phys_rho(P, vel[index].w, f);
					const float dynvisc = P->rheologytype == B200SPH_RHEOLOGY_INVISCID ? 0.0f :
						(P->compvisc == B200SPH_COMPVISC_KINEMATIC ? P->visccoeff[f] * rho : P->visccoeff[f]);
					const float partsurf = P->partsurf == 0.0f ? P->r0 * P->r0 : P->partsurf;
					const i3 gp = grid_pos_from_hash(P, hash[index] & CELLTYPE_BITMASK);
					const int gpa[3] = { gp.x, gp.y, gp.z };
					const float pa[3] = { p.x, p.y, p.z };
					const f4 v4 = vel[index];
					for (int k = 0; k < opts->numplanes; ++k) {
						float d[3];
						for (int a = 0; a < 3; ++a)
							d[a] = (float)(gpa[a] - opts->plane_gridpos[k][a]) * P->cell_size[a] + (pa[a] - opts->plane_pos[k][a]);
						const float *nrm = opts->plane_normal[k];
						const float r = fabsf(d[0] * nrm[0] + d[1] * nrm[1] + d[2] * nrm[2]);
						if (r < P->r0) {
							const float DvDt = P->dcoeff * (powf(P->r0 / r, P->p1coeff) - powf(P->r0 / r, P->p2coeff)) / (r * r);
							const float rp[3] = { nrm[0] * r, nrm[1] * r, nrm[2] * r };
							fo.x += DvDt * rp[0]; fo.y += DvDt * rp[1]; fo.z += DvDt * rp[2];
							const float vn = (v4.x * rp[0] + v4.y * rp[1] + v4.z * rp[2]) / r;
							const float coeff = -dynvisc * partsurf / (p.w * r);
							fo.x += coeff * (v4.x - vn * rp[0] / r); fo.y += coeff * (v4.y - vn * rp[1] / r); fo.z += coeff * (v4.z - vn * rp[2] / r);
						}
					}
				}
				const float c = ssp[index];
				const float v = fmaxf(sqrtf(fo.x * fo.x + fo.y * fo.y + fo.z * fo.z), c * c / P->slength);
				if (v > m) m = v;
			}
			/* force-feedback bodies: :4116-4141 (force x mass, torque about the centre of gravity) */
			if (bodies && rb_forces && compute_force(inf) && ptype(inf) != B200SPH_PT_VERTEX) {
				const int obj = inf.y & 0xFFF;
				fo.x *= p.w; fo.y *= p.w; fo.z *= p.w;
				const uint32_t rbindex = pid(inf) + (uint32_t)bodies->startIndex[obj];
				rb_forces[rbindex] = fo;
				const i3 gp = grid_pos_from_hash(P, hash[index] & CELLTYPE_BITMASK);
				const float ax = (float)(gp.x - bodies->cgGridPos[obj][0]) * P->cell_size[0] + (p.x - bodies->cgPos[obj][0]);
				const float ay = (float)(gp.y - bodies->cgGridPos[obj][1]) * P->cell_size[1] + (p.y - bodies->cgPos[obj][1]);
				const float az = (float)(gp.z - bodies->cgGridPos[obj][2]) * P->cell_size[2] + (p.z - bodies->cgPos[obj][2]);
				f4 tq = { ay * fo.z - az * fo.y, az * fo.x - ax * fo.z, ax * fo.y - ay * fo.x, 0.0f };
				rb_torques[rbindex] = tq;
			}
			forces[index] = fo;
		}
		if (cfl) cfl[cfl_offset + b] = m;
	}
	free(pprec); free(ssp);
	return nblocks;
}

uint32_t oracle_forces(const b200sph_params *P, const f4 *pos, const f4 *vel, const us4 *info,
	const uint32_t *hash, const uint32_t *cell_start, const uint16_t *neibs_list,
	const float *eos_p_in, const float *eos_c_in,
	f4 *forces, float *cfl, f4 *abssum,
	uint32_t num_particles, uint32_t from, uint32_t to, uint32_t cfl_offset,
	const oracle_bodies *bodies, f4 *rb_forces, f4 *rb_torques)
{
	return oracle_forces_ex(P, pos, vel, info, hash, cell_start, neibs_list, eos_p_in, eos_c_in, forces, cfl, abssum,
		num_particles, from, to, cfl_offset, bodies, rb_forces, rb_torques, NULL);
}

/* per-particle EOS quantities as the oracle evaluates them (for tests) */
void oracle_eos(const b200sph_params *P, const f4 *vel, const us4 *info, float *p_precalc, float *sspeed, uint32_t n)
{
	for (uint32_t i = 0; i < n; ++i) {
		const int f = fluid_num(info[i]);
		const float rho = phys_rho(P, vel[i].w, f);
		p_precalc[i] = eos_P(P, vel[i].w, f) / (rho * rho);
		sspeed[i] = eos_c(P, vel[i].w, f);
	}
}

/* dtreduce: src/cuda/forces.cu:557-607 (cflmax = plain max) */
float oracle_dtreduce(const b200sph_params *P, const float *cfl, uint32_t num_blocks)
{
	float maxcfl = 0.0f;
	for (uint32_t i = 0; i < num_blocks; ++i) if (cfl[i] > maxcfl) maxcfl = cfl[i];
	float dt = P->dtadaptfactor * fminf(sqrtf(P->slength / maxcfl), P->slength / P->max_sound_speed_cfl);
	if (P->rheologytype != B200SPH_RHEOLOGY_INVISCID || P->turbmodel > B200SPH_TURB_ARTIFICIAL) {
		float dt_visc = P->slength * P->slength / P->max_kinvisc;
		dt_visc *= 0.125f;
		if (dt_visc < dt) dt = dt_visc;
	}
	return dt;
}

/* euler: src/cuda/euler_kernel.def:396-540, :117-134 (corrected velocity), :200-206 (continuity).
 * including the rigid motion of moving-body particles (:470-503). */
void oracle_euler(const b200sph_params *P, const f4 *old_pos, const f4 *old_vel, const us4 *info,
	const uint32_t *hash, const f4 *forces, f4 *new_pos, f4 *new_vel,
	uint32_t num_particles, uint32_t range_end, float dt, int step, const oracle_bodies *bodies)
{
	void oracle_euler_ex(const b200sph_params *, const f4 *, const f4 *, const us4 *, const uint32_t *, const f4 *, const f4 *,
		f4 *, f4 *, uint32_t, uint32_t, float, int, const oracle_bodies *);
	oracle_euler_ex(P, old_pos, old_vel, info, hash, forces, NULL, new_pos, new_vel, num_particles, range_end, dt, step, bodies);
}

/* ... with the XSPH correction (euler_kernel.def:165-180): velc += epsxsph * xsph */
void oracle_euler_ex(const b200sph_params *P, const f4 *old_pos, const f4 *old_vel, const us4 *info,
	const uint32_t *hash, const f4 *forces, const f4 *xsph, f4 *new_pos, f4 *new_vel,
	uint32_t num_particles, uint32_t range_end, float dt, int step, const oracle_bodies *bodies)
{
	(void)num_particles;
	const int integrate_boundary = (P->boundarytype == B200SPH_DYN_BOUNDARY || P->boundarytype == B200SPH_SA_BOUNDARY);
#pragma omp parallel for
	for (uint32_t i = 0; i < range_end; ++i) {
		f4 p = old_pos[i], v = old_vel[i];
		const f4 f = forces[i];
		const us4 inf = info[i];
		const int t = ptype(inf);
		if (!inactive(p) && !(t == B200SPH_PT_BOUNDARY && !integrate_boundary && !is_moving(inf))) {
			float vcx = v.x, vcy = v.y, vcz = v.z;
			if (step == 2) { const float hdt = dt / 2; vcx += f.x * hdt; vcy += f.y * hdt; vcz += f.z * hdt; }
			if (xsph && (P->simflags & B200SPH_ENABLE_XSPH)) {
				vcx += P->epsxsph * xsph[i].x; vcy += P->epsxsph * xsph[i].y; vcz += P->epsxsph * xsph[i].z;
			}
			if (t == B200SPH_PT_FLUID) {
				p.x += vcx * dt; p.y += vcy * dt; p.z += vcz * dt;
				v.w += dt * f.w;
				v.x += dt * f.x; v.y += dt * f.y; v.z += dt * f.z;
			} else if (t == B200SPH_PT_BOUNDARY || t == B200SPH_PT_VERTEX) {
				if (is_moving(inf) && bodies) {              /* :470-503, applyrot euler_kernel.cu:67-74 */
					const int obj = inf.y & 0xFFF;
					const i3 gp = grid_pos_from_hash(P, hash[i] & CELLTYPE_BITMASK);
					const float rx = (float)(gp.x - bodies->cgGridPos[obj][0]) * P->cell_size[0] + (p.x - bodies->cgPos[obj][0]);
					const float ry = (float)(gp.y - bodies->cgGridPos[obj][1]) * P->cell_size[1] + (p.y - bodies->cgPos[obj][1]);
					const float rz = (float)(gp.z - bodies->cgGridPos[obj][2]) * P->cell_size[2] + (p.z - bodies->cgPos[obj][2]);
					const float *rot = bodies->steprot[obj];
					p.x += (rot[0] - 1.0f) * rx + rot[1] * ry + rot[2] * rz;
					p.y += rot[3] * rx + (rot[4] - 1.0f) * ry + rot[5] * rz;
					p.z += rot[6] * rx + rot[7] * ry + (rot[8] - 1.0f) * rz;
					p.x += bodies->trans[obj][0]; p.y += bodies->trans[obj][1]; p.z += bodies->trans[obj][2];
					const float *w = bodies->angularvel[obj], *lv = bodies->linearvel[obj];
					v.x = lv[0] + (w[1] * rz - w[2] * ry);
					v.y = lv[1] + (w[2] * rx - w[0] * rz);
					v.z = lv[2] + (w[0] * ry - w[1] * rx);
				}
				if (P->boundarytype == B200SPH_DYN_BOUNDARY) v.w += dt * f.w;
			}
		}
		new_pos[i] = p; new_vel[i] = v;
	}
}

/* host-side particle placement: ProblemCore::calc_localpos_and_hash src/ProblemCore.cc:1554-1583 */
void oracle_localpos_and_hash(const b200sph_params *P, const double *gpos /* xyz */, float mass,
	f4 *localpos, uint32_t *hash)
{
	int g[3];
	for (int a = 0; a < 3; ++a) {
		/* calc_grid_pos :1508-1520 uses the double cell size of the problem; here the float one
		 * (callers that need exactness pass positions that are not on cell faces) */
		g[a] = (int)floor((gpos[a] - (double)P->world_origin[a]) / (double)P->cell_size[a]);
		g[a] = imin(imax(0, g[a]), (int)P->grid_size[a] - 1);
	}
	i3 gp = { g[0], g[1], g[2] };
	*hash = calc_grid_hash(P, gp);
	localpos->x = (float)(gpos[0] - (double)P->world_origin[0] - (g[0] + 0.5) * (double)P->cell_size[0]);
	localpos->y = (float)(gpos[1] - (double)P->world_origin[1] - (g[1] + 0.5) * (double)P->cell_size[1]);
	localpos->z = (float)(gpos[2] - (double)P->world_origin[2] - (g[2] + 0.5) * (double)P->cell_size[2]);
	localpos->w = mass;
}

/* ------------------------------------------------------------------------ */
/* neighbour-list traversal as an iterator: src/cuda/neibs_iteration.cuh      */
/* :56-396 (for_each_neib2), getNeibIndex src/cuda/cellgrid.cuh:198-226       */
/* ------------------------------------------------------------------------ */
typedef struct {
	const b200sph_params *P; const uint32_t *cell_start; const uint16_t *list; const f4 *pos;
	uint32_t index; f4 p; i3 gp;
	int section, nsections; long long row;
	uint32_t base; float pcx, pcy, pcz;
	uint32_t j; f4 rel;     /* current neighbour and relPos (w = neighbour mass) */
} neib_iter;

static void neib_iter_init(neib_iter *it, const b200sph_params *P, uint32_t index, const f4 *pos, const uint32_t *hash,
	const uint32_t *cell_start, const uint16_t *list, int with_boundary)
{
	it->P = P; it->cell_start = cell_start; it->list = list; it->pos = pos;
	it->index = index; it->p = pos[index];
	it->gp = grid_pos_from_hash(P, hash[index] & CELLTYPE_BITMASK);
	it->section = 0; it->nsections = with_boundary ? 2 : 1; it->row = -1;
	it->base = 0; it->pcx = it->pcy = it->pcz = 0;
}
static int neib_iter_next(neib_iter *it)
{
	const b200sph_params *P = it->P;
	for (;;) {
		if (it->section >= it->nsections) return 0;
		if (it->row < 0) it->row = it->section == 0 ? 0 : (long long)P->neibboundpos;
		else it->row += it->section == 0 ? 1 : -1;
		if (it->row < 0 || it->row >= (long long)P->neiblistsize) { it->section++; it->row = -1; continue; }
		uint32_t nd = it->list[(size_t)it->row * P->neiblist_stride + it->index];
		if (nd == NEIBS_END) { it->section++; it->row = -1; continue; }
		if (nd >= CELLNUM_ENCODED) {
			const int cell = (int)(nd >> CELLNUM_SHIFT) - 1;
			nd &= NEIBINDEX_MASK;
			const int ox = cell % 3 - 1, oy = (cell / 3) % 3 - 1, oz = cell / 9 - 1;
			it->pcx = it->p.x - (float)ox * P->cell_size[0];
			it->pcy = it->p.y - (float)oy * P->cell_size[1];
			it->pcz = it->p.z - (float)oz * P->cell_size[2];
			i3 ng = { it->gp.x + ox, it->gp.y + oy, it->gp.z + oz };
			it->base = it->cell_start[calc_grid_hash_periodic(P, ng)];
		}
		it->j = it->base + nd;
		const f4 np = it->pos[it->j];
		it->rel.x = it->pcx - np.x; it->rel.y = it->pcy - np.y; it->rel.z = it->pcz - np.z; it->rel.w = np.w;
		return 1;
	}
}

/* Shepard filter: shepardDevice, src/cuda/forces_kernel.cu:418-507 */
void oracle_shepard(const b200sph_params *P, const f4 *pos, const f4 *old_vel, f4 *new_vel, const us4 *info,
	const uint32_t *hash, const uint32_t *cell_start, const uint16_t *neibs_list, uint32_t range_end)
{
	const float wcoeff = kernel_wcoeff(P), h = P->slength;
#pragma omp parallel for schedule(dynamic, 512)
	for (uint32_t index = 0; index < range_end; ++index) {
		const us4 inf = info[index];
		const f4 p = pos[index];
		if (inactive(p)) continue;
		f4 v = old_vel[index];
		if (!is_fluid(inf)) { new_vel[index] = v; continue; }
		const int fnum = fluid_num(inf);
		float temp1 = p.w * wendland_W(0, h, wcoeff);
		float temp2 = temp1 / phys_rho(P, v.w, fnum);
		neib_iter it;
		neib_iter_init(&it, P, index, pos, hash, cell_start, neibs_list, P->boundarytype == B200SPH_DYN_BOUNDARY);
		while (neib_iter_next(&it)) {
			if (!isfinite(it.rel.w)) continue;
			const float r = sqrtf(it.rel.x * it.rel.x + it.rel.y * it.rel.y + it.rel.z * it.rel.z);
			const float neib_rho = phys_rho(P, old_vel[it.j].w, fluid_num(info[it.j]));
			if (r < P->influenceradius) {
				const float w = wendland_W(r, h, wcoeff) * it.rel.w;
				temp1 += w;
				temp2 += w / neib_rho;
			}
		}
		v.w = (temp1 / temp2) / P->rho0[fnum] - 1.0f;      /* numerical_density, phys_core.cu:145-151 */
		new_vel[index] = v;
	}
}

/* symtensor4 algebra: src/cuda/tensor.cu:64-100 (det), :240-249 (dot), :261-271 (ddot), :273-283 (adjugate_row1);
 * hypot: src/vector_math.h:1231-1240 (float4 / float multiplies by the reciprocal, :1093-1097) */
typedef struct { float xx, xy, xz, xw, yy, yz, yw, zz, zw, ww; } st4;
static float st4_det(const st4 *T)
{
	float ret = 0, M = 0;
	M += T->xx * (T->yy * T->zz - T->yz * T->yz);
	M -= T->xy * (T->xy * T->zz - T->xz * T->yz);
	M += T->xz * (T->xy * T->yz - T->xz * T->yy);
	ret += M * T->ww;
	M = 0;
	M += T->xx * (T->yy * T->zw - T->yz * T->yw);
	M -= T->xy * (T->xy * T->zw - T->xz * T->yw);
	M += T->xw * (T->xy * T->yz - T->xz * T->yy);
	ret -= M * T->zw;
	M = 0;
	M += T->xx * (T->yz * T->zw - T->zz * T->yw);
	M -= T->xz * (T->xy * T->zw - T->xz * T->yw);
	M += T->xw * (T->xy * T->zz - T->xz * T->yz);
	ret += M * T->yw;
	M = 0;
	M += T->xy * (T->yz * T->zw - T->zz * T->yw);
	M -= T->xz * (T->yy * T->zw - T->yz * T->yw);
	M += T->xw * (T->yy * T->zz - T->yz * T->yz);
	ret -= M * T->xw;
	return ret;
}
static f4 st4_dot(const st4 *T, f4 v)
{
	f4 r = { T->xx * v.x + T->xy * v.y + T->xz * v.z + T->xw * v.w,
	         T->xy * v.x + T->yy * v.y + T->yz * v.z + T->yw * v.w,
	         T->xz * v.x + T->yz * v.y + T->zz * v.z + T->zw * v.w,
	         T->xw * v.x + T->yw * v.y + T->zw * v.z + T->ww * v.w };
	return r;
}
static float st4_ddot(const st4 *T, f4 v)
{
	return T->xx * v.x * v.x + T->yy * v.y * v.y + T->zz * v.z * v.z + T->ww * v.w * v.w +
		2 * ((T->xy * v.y + T->xw * v.w) * v.x + (T->yz * v.z + T->yw * v.w) * v.y + (T->xz * v.x + T->zw * v.w) * v.z);
}
static f4 st4_adj_row1(const st4 *T)
{
	f4 r = {
		T->yy * T->zz * T->ww + T->yz * T->zw * T->yw + T->yw * T->yz * T->zw - T->yy * T->zw * T->zw - T->yz * T->yz * T->ww - T->yw * T->zz * T->yw,
		T->xy * T->zw * T->zw + T->yz * T->xz * T->ww + T->yw * T->zz * T->xw - T->xy * T->zz * T->ww - T->yz * T->zw * T->xw - T->yw * T->xz * T->zw,
		T->xy * T->yz * T->ww + T->yy * T->zw * T->xw + T->yw * T->xz * T->yw - T->xy * T->zw * T->yw - T->yy * T->xz * T->ww - T->yw * T->yz * T->xw,
		T->xy * T->zz * T->yw + T->yy * T->xz * T->zw + T->yz * T->yz * T->xw - T->xy * T->yz * T->zw - T->yy * T->zz * T->xw - T->yz * T->xz * T->yw };
	return r;
}
static float f4_hypot(f4 v)
{
	const float p = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)));
	if (!p) return 0;
	const float inv = 1.0f / p;
	const float wx = v.x * inv, wy = v.y * inv, wz = v.z * inv, ww = v.w * inv;
	return p * sqrtf(wx * wx + wy * wy + wz * wz + ww * ww);
}

/* MLS filter: MlsDevice, src/cuda/forces_kernel.cu:509-721 (MlsMatrixContrib :235-249, MlsCorrContrib :256-260) */
void oracle_mls(const b200sph_params *P, const f4 *pos, const f4 *old_vel, f4 *new_vel, const us4 *info,
	const uint32_t *hash, const uint32_t *cell_start, const uint16_t *neibs_list, uint32_t range_end)
{
	const float wcoeff = kernel_wcoeff(P), h = P->slength;
	const int dyn = P->boundarytype == B200SPH_DYN_BOUNDARY;
#pragma omp parallel for schedule(dynamic, 512)
	for (uint32_t index = 0; index < range_end; ++index) {
		const us4 inf = info[index];
		const f4 p = pos[index];
		if (inactive(p)) continue;
		f4 v = old_vel[index];
		const int fnum = fluid_num(inf);
		st4 mls; memset(&mls, 0, sizeof(mls));
		mls.xx = wendland_W(0, h, wcoeff) * p.w / phys_rho(P, v.w, fnum);
		neib_iter it;
		neib_iter_init(&it, P, index, pos, hash, cell_start, neibs_list, dyn);
		while (neib_iter_next(&it)) {
			if (!isfinite(it.rel.w)) continue;
			const float r = sqrtf(it.rel.x * it.rel.x + it.rel.y * it.rel.y + it.rel.z * it.rel.z);
			const float neib_rho = phys_rho(P, old_vel[it.j].w, fluid_num(info[it.j]));
			if (r < P->influenceradius) {
				const float w = wendland_W(r, h, wcoeff) * it.rel.w / neib_rho;
				const float inv_h = 1.0f / h;
				const float x = it.rel.x * inv_h, y = it.rel.y * inv_h, z = it.rel.z * inv_h;
				mls.xx += w;
				mls.xy += x * w; mls.xz += y * w; mls.xw += z * w;
				mls.yy += x * x * w; mls.yz += x * y * w; mls.yw += x * z * w;
				mls.zz += y * y * w; mls.zw += y * z * w;
				mls.ww += z * z * w;
			}
		}
		const float D = st4_det(&mls);
		f4 B;
		if (fabsf(D) < 1.1920929e-07f) {
			st4 m2 = mls;
			const float eps = fabsf(D) + 1.1920929e-07f;
			m2.xx += eps; m2.yy += eps; m2.zz += eps; m2.ww += eps;
			const float inv = 1.0f / st4_det(&m2);
			const f4 a = st4_adj_row1(&m2);
			B = (f4){ a.x * inv, a.y * inv, a.z * inv, a.w * inv };
		} else {
			const float inv = 1.0f / D;
			const f4 a = st4_adj_row1(&mls);
			B = (f4){ a.x * inv, a.y * inv, a.z * inv, a.w * inv };
		}
		for (unsigned steps = 0; steps < 32; ++steps) {
			const float lenB = f4_hypot(B);
			const f4 MB = st4_dot(&mls, B);
			const f4 res = { 1.0f - MB.x, 0.0f - MB.y, 0.0f - MB.z, 0.0f - MB.w };
			const float num = st4_ddot(&mls, res);
			const f4 Mp = st4_dot(&mls, res);
			const float den = Mp.x * Mp.x + Mp.y * Mp.y + Mp.z * Mp.z + Mp.w * Mp.w;
			const float s = num / den;
			const f4 corr = { s * res.x, s * res.y, s * res.z, s * res.w };
			const float lencorr = f4_hypot(corr);
			if (f4_hypot(res) < lenB * 1.1920929e-07f) break;
			if (lencorr < 2 * lenB * 1.1920929e-07f) break;
			B.x += corr.x; B.y += corr.y; B.z += corr.z; B.w += corr.w;
		}
		B.y /= h; B.z /= h; B.w /= h;
		v.w = B.x * wendland_W(0, h, wcoeff) * p.w;
		neib_iter_init(&it, P, index, pos, hash, cell_start, neibs_list, dyn);
		while (neib_iter_next(&it)) {
			if (!isfinite(it.rel.w)) continue;
			const float r = sqrtf(it.rel.x * it.rel.x + it.rel.y * it.rel.y + it.rel.z * it.rel.z);
			if (r < P->influenceradius && (dyn || is_fluid(info[it.j]))) {
				const float w = wendland_W(r, h, wcoeff) * it.rel.w;
				v.w += (B.x + B.y * it.rel.x + B.z * it.rel.y + B.w * it.rel.z) * w;
			}
		}
		v.w = v.w / P->rho0[fnum] - 1.0f;
		new_vel[index] = v;
	}
}

/* TESTPOINTS: calcTestpointsVelocityDevice, src/cuda/post_process_kernel.cu:134-240 (vel / tke / epsilon in place) */
void oracle_testpoints(const b200sph_params *P, const f4 *pos, f4 *vel, float *tke, float *epsilon, const us4 *info,
	const uint32_t *hash, const uint32_t *cell_start, const uint16_t *neibs_list, uint32_t range_end)
{
	const float wcoeff = kernel_wcoeff(P), h = P->slength;
	for (uint32_t index = 0; index < range_end; ++index) {
		if (!is_testpoint(info[index])) continue;
		f4 avg = { 0, 0, 0, 0 };
		float tkeavg = 0, epsavg = 0, alpha = 0;
		neib_iter it;
		neib_iter_init(&it, P, index, pos, hash, cell_start, neibs_list, 0);
		while (neib_iter_next(&it)) {
			const float r = sqrtf(it.rel.x * it.rel.x + it.rel.y * it.rel.y + it.rel.z * it.rel.z);
			if (r < P->influenceradius) {
				const f4 nv = vel[it.j];
				const int nf = fluid_num(info[it.j]);
				const float w = wendland_W(r, h, wcoeff) * it.rel.w / phys_rho(P, nv.w, nf);
				avg.x += w * nv.x; avg.y += w * nv.y; avg.z += w * nv.z;
				avg.w += w * eos_P(P, nv.w, nf);
				if (tke) tkeavg += w * tke[it.j];
				if (epsilon) epsavg += w * epsilon[it.j];
				alpha += w;
			}
		}
		if (alpha > 1e-5f) {
			const float inv = 1.0f / alpha;
			avg.x *= inv; avg.y *= inv; avg.z *= inv; avg.w *= inv;
			tkeavg /= alpha; epsavg /= alpha;
		} else {
			avg = (f4){ 0, 0, 0, 0 };
			tkeavg = epsavg = 0;
		}
		vel[index] = avg;
		if (tke) tke[index] = tkeavg;
		if (epsilon) epsilon[index] = epsavg;
	}
}
