// common.cuh — shared declarations of the B200 WCSPH engine (internal; the public
// boundary is include/b200sph.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/b200sph.h"

typedef unsigned int uint;
typedef unsigned short ushort;

// ---- data contract constants (reference: src/multi_gpu_defines.h:58-84,
// src/hashkey.h:44, src/common_types.h:57-74) ----
#define CELLTYPE_BITMASK (~(3U << 30))
#define CELL_HASH_MAX 0xFFFFFFFFu
#define CELL_EMPTY 0xFFFFFFFFu
#define NEIBS_END 0xFFFFu
#define CELLNUM_SHIFT 11
#define CELLNUM_ENCODED (1U << CELLNUM_SHIFT)
#define NEIBINDEX_MASK (CELLNUM_ENCODED - 1)
// blocked neighbour-list layout (include/b200sph.h, B200SPH_NEIBLIST_BLOCK): element offset of row 0 of particle `index`
// and the distance between its rows. `block` = particles per block, a power of two.
__device__ __forceinline__ size_t list_column(const uint index, const uint stride, const uint rows, const uint block, uint &row_step)
{
	const uint b0 = index & ~(block - 1u);
	row_step = min(block, stride - b0);
	return (size_t)b0 * rows + (index - b0);
}

#define PT_FLUID 0
#define PT_BOUNDARY 1
#define PT_VERTEX 2
#define PT_TESTPOINT 3

__host__ __device__ static inline int object_of_y(unsigned short y) { return y & 0xFFF; }   // src/particleinfo.h:403-410

// block sizes; forces uses 128 so that the CFL array has exactly the layout the
// reference's getFmaxElements() sizes it for (src/cuda/forces.cu:56-68,540-544)
#define BLOCK_STREAM 256
#define BLOCK_FORCES 128

// Device-side copy of everything the kernels need. Passed BY VALUE as a
// __grid_constant__ kernel parameter (constant bank, no symbol upload, no sync,
// one context per device/thread without global state).
struct DevParams {
	float cellSize[3];
	int gridSize[3];
	int coord[3];           // axis of COORD1,2,3
	int hstride[3];         // linear-hash stride of one cell step along x, y, z (derived from coord + gridSize)
	uint periodic;
	uint neiblistsize, neibboundpos, stride;
	uint listblock;         // particles per block of the neighbour-list layout
	float nlSqInflRad;
	uint kerneltype, densitydiffusiontype, boundarytype;
	uint inviscid, turbmodel, compvisc, viscavgop, is_const_visc;
	float slength, influenceradius, deltap;
	float fcoeff_wendland;  // 105/(128 pi h^5), src/cuda/forces.cu:289
	float densityDiffCoeff, artvisccoeff, epsartvisc;
	uint numFluids;
	float rho0[B200SPH_MAX_FLUIDS], bcoeff[B200SPH_MAX_FLUIDS], gammacoeff[B200SPH_MAX_FLUIDS];
	float sscoeff[B200SPH_MAX_FLUIDS], sspowercoeff[B200SPH_MAX_FLUIDS], visccoeff[B200SPH_MAX_FLUIDS];
	float sqC0[B200SPH_MAX_FLUIDS];
	float gravity[3];
	// options only the general forces kernel / the filters read
	uint viscmodel, simflags;
	float wcoeff_wendland;  // 21/(16 pi h^3), src/cuda/forces.cu:283
	float epsxsph, monaghanViscCoeff;
	float visc2coeff[B200SPH_MAX_FLUIDS];
	// geometric planes (src/planes.h:42-46, src/cuda/geom_core.cu:52-53) and their Lennard-Jones repulsion
	uint numplanes;
	float r0, dcoeff, p1coeff, p2coeff, partsurf;
	float planeNormal[B200SPH_MAX_PLANES][3];
	int planeGridPos[B200SPH_MAX_PLANES][3];
	float planePos[B200SPH_MAX_PLANES][3];
	// per launch (filled by the launcher on its by-value copy): dt of the command, read by BREZZI diffusion only
	float cmd_dt;
	int cmd_step;                        // 1 / 2
	const struct StepState *dev_state;   // non-NULL: dt from the device-resident record (dt/2 for step 1)
};

// One neighbour record of the pair kernel: the particle's pos and vel entries side by side, so that a gather is ONE
// 256-bit load touching one 32-byte sector (forces.cu)
struct __align__(32) PosVel { float4 pos, vel; };

struct NeibsCounters {     // mirrors the reference's device counters, src/cuda/buildneibs_kernel.cu:108-112
	int numInteractions;
	int maxFluidBoundaryNeibs;
	int maxVertexNeibs;
	int hasTooManyNeibs;
	int hasMaxNeibs[3];
	int pad;
};

// moving / force-feedback bodies: device copy of what the reference keeps in __constant__ arrays
// (src/cuda/forces_kernel.cu:81-83, src/cuda/euler_kernel.cu:45-50)
// The centre of gravity exists TWICE on purpose, like the reference's two __constant__ copies: the integrator uploads
// cg(n+1) to the forces engine right after MOVE_BODIES (FORCES_UPLOAD_OBJECTS_CG) while the integration keeps
// rotating about cg(n) until the end of the step (EULER_UPLOAD_OBJECTS_CG;
// src/integrators/PredictorCorrectorIntegrator.cc:332,556-587, "We always have cg = cg(n)" src/cuda/euler_kernel.def:488).
struct BodyData {
	int cgGridPos[B200SPH_MAX_BODIES][3];      // forces engine copy: torque arm in finalize
	float cgPos[B200SPH_MAX_BODIES][3];
	int eulCgGridPos[B200SPH_MAX_BODIES][3];   // integration engine copy: rigid motion in euler
	float eulCgPos[B200SPH_MAX_BODIES][3];
	int startIndex[B200SPH_MAX_BODIES];
	float trans[B200SPH_MAX_BODIES][3];
	float steprot[B200SPH_MAX_BODIES][9];
	float linearvel[B200SPH_MAX_BODIES][3];
	float angularvel[B200SPH_MAX_BODIES][3];
};

#define BODY_SET_CG_FORCES 1u
#define BODY_SET_START     2u
#define BODY_SET_CG_EULER  4u
#define BODY_SET_TRANS     8u
#define BODY_SET_STEPROT   16u
#define BODY_SET_LINVEL    32u
#define BODY_SET_ANGVEL    64u
#define BODY_SET_EULER_ALL (BODY_SET_CG_EULER | BODY_SET_TRANS | BODY_SET_STEPROT | BODY_SET_LINVEL | BODY_SET_ANGVEL)

// device-resident time-stepping record (b200sph_step_* entry points)
struct StepState {
	double t;
	unsigned long long iterations;
	float dt, dt1, dt2, pad;
};

#define B200_MAX_LANES 2
struct b200sph_ctx {
	b200sph_params hp;      // host copy
	DevParams dp;
	int device;
	cudaStream_t stream;
	// scratch (grown on demand)
	void *sort_tmp; size_t sort_tmp_bytes;
	uint64_t *keys_in, *keys_out; uint32_t *vals_out; void *info_tmp; size_t sort_cap;
	PosVel *pv[2]; size_t pv_cap[2];        // context-owned neighbour records (forces.cu b200_packed_scratch)
	NeibsCounters *d_counters;
	float *d_scalar;                        // device scalar for reductions
	float *h_scalar;                        // pinned host scalar
	int *d_flag; int *h_flag;
	StepState *d_step; StepState *h_step;
	BodyData *d_bodies; BodyData *h_bodies; int have_bodies;
	unsigned bodies_set;                    // BODY_SET_* bits: which setrb* calls were made
	// pipelined stepping of a host-resident state (hoststep.cu): copy streams, per-stripe events, and what the
	// previous call left in flight (the next call chains on it stripe by stripe)
	cudaStream_t up_stream, down_stream;
	cudaEvent_t up_ev[B200SPH_MAX_STRIPES], down_ev[B200SPH_MAX_STRIPES], comp_ev[B200SPH_MAX_STRIPES], pred_ev[B200SPH_MAX_STRIPES];
	cudaEvent_t fence_ev, up_all_ev, down_all_ev;
	cudaStream_t lane_stream[B200_MAX_LANES]; int host_lanes;        // extra compute lanes ([0] unused: the context's stream)
	cudaEvent_t fork_ev[2], join_ev[2 * B200_MAX_LANES];
	uint32_t host_bounds[B200SPH_MAX_STRIPES + 1]; uint32_t host_nstripes;
	uint32_t host_ring_next;      // periodic COORD3: the stripe the next b200sph_step_host starts its ring at (hoststep.cu)
	const void *host_pos_last, *host_vel_last, *dev_pos_last, *dev_vel_last;
	int host_pending;
	cudaEvent_t *trace_ev; int trace_resident;      // B200SPH_HOST_TRACE diagnostics
};

// ---- error plumbing ----
void b200_set_error(const char *fmt, ...);
#define CUDA_TRY(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { \
	b200_set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
	return B200SPH_ECUDA; } } while (0)
#define KERNEL_TRY() CUDA_TRY(cudaGetLastError())
#define CHECK_CTX(ctx) do { if (!(ctx)) { b200_set_error("null context"); return B200SPH_EINVAL; } \
	CUDA_TRY(cudaSetDevice((ctx)->device)); } while (0)

static inline uint div_up(uint a, uint b) { return (a + b - 1) / b; }

// ---- device helpers ----
#ifdef __CUDACC__
__device__ __forceinline__ void ld_posvel(const PosVel *p, float4 &a, float4 &b)
{
	asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
		: "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}
__device__ __forceinline__ void st_posvel(PosVel *p, const float4 a, const float4 b)
{
	asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
		:: "l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w) : "memory");
}
__device__ __forceinline__ int ptype_of(ushort4 info) { return info.x & 7; }
__device__ __forceinline__ uint id_of(ushort4 info) { return (uint)info.z | ((uint)info.w << 16); }
__device__ __forceinline__ int fluid_num_of(ushort4 info) { return info.y >> 12; }
__device__ __forceinline__ bool inactive_w(float w) { return !isfinite(w); }

// linear cell index from grid position, reference calcGridHash (src/cuda/cellgrid.cuh:101-106)
__device__ __forceinline__ uint grid_hash(const DevParams &P, int gx, int gy, int gz)
{
	// gp.C3*G.C2*G.C1 + gp.C2*G.C1 + gp.C1 with the axis strides precomputed on the host
	return (uint)(gx * P.hstride[0] + gy * P.hstride[1] + gz * P.hstride[2]);
}
// reference calcGridPosFromCellHash (src/cuda/cellgrid.cuh:117-128)
__device__ __forceinline__ int3 grid_pos(const DevParams &P, uint cellHash)
{
	auto selG = [&](int c) { return c == 0 ? P.gridSize[0] : (c == 1 ? P.gridSize[1] : P.gridSize[2]); };
	const int G1 = selG(P.coord[0]), G2 = selG(P.coord[1]);
	int temp = G2 * G1;
	const int g3 = (int)cellHash / temp;
	temp = (int)cellHash - g3 * temp;
	const int g2 = temp / G1;
	const int g1 = temp - g2 * G1;
	int3 r;
	// scatter back to x,y,z
	r.x = P.coord[0] == 0 ? g1 : (P.coord[1] == 0 ? g2 : g3);
	r.y = P.coord[0] == 1 ? g1 : (P.coord[1] == 1 ? g2 : g3);
	r.z = P.coord[0] == 2 ? g1 : (P.coord[1] == 2 ? g2 : g3);
	return r;
}
#endif

// ---- internal launchers implemented in the .cu files ----
int b200_packed_scratch(b200sph_ctx *ctx, int which, uint32_t n, PosVel **out);   // forces.cu
void b200_hoststep_destroy(b200sph_ctx *ctx);
int b200_euler_bodies(b200sph_ctx *ctx, const uint32_t *hash, const BodyData **out);   // api.cu
