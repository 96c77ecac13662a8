/*
 * b200sph.h — C ABI of the B200-native WCSPH per-timestep engine.
 *
 * This is the drop-in boundary for the hot path of GPUSPH (neighbour build ->
 * forces -> integration). Every entry point replaces one virtual of the
 * reference's three abstract engines; the reference file:line it replaces is
 * cited on each declaration (paths relative to the GPUSPH source tree).
 * The C++ adapter classes that a GPUSPH maintainer would add on the reference
 * side (thin subclasses of AbstractNeibsEngine / AbstractForcesEngine /
 * AbstractIntegrationEngine forwarding raw device pointers here) are in
 * gpusph_b200/host/b200_engines.h and described in INTEGRATION.md.
 *
 * Conventions
 *  - plain C types only; all particle arrays are DEVICE pointers in the
 *    reference's own layouts:
 *      pos    float4[N]  xyz = position relative to the cell centre, w = mass
 *                         (non-finite w  => particle inactive)
 *      vel    float4[N]  xyz = velocity, w = relative density rho/rho0 - 1
 *      info   ushort4[N] x = type(3 bits) | flags<<3, y = fluid<<12 | object,
 *                         z,w = id lo,hi              (src/particleinfo.h:66)
 *      hash   uint32[N]  bits 0-29 linear cell index, bits 30-31 cell type,
 *                         0xFFFFFFFF inactive          (src/hashkey.h:40-61)
 *      forces float4[N]  xyz = Dv/Dt, w = D(rho~)/Dt
 *      cellStart/cellEnd uint32[C], 0xFFFFFFFF = empty cell
 *      neibsList ushort[neiblistsize][stride]  (src/cuda/buildneibs_kernel.cu:1029)
 *  - every function returns 0 on success, a negative B200SPH_E* code on
 *    failure; b200sph_last_error() gives the message. The reference throws
 *    C++ exceptions (std::runtime_error / std::invalid_argument); the adapter
 *    turns the codes back into those exceptions.
 *  - work is enqueued on the context's CUDA stream (default: the legacy
 *    default stream, like the reference). Calls that return host values
 *    (b200sph_neibs_getinfo, b200sph_dtreduce, ...) synchronise that stream.
 *  - There is NO CPU fallback: without a CUDA device every compute call
 *    fails with B200SPH_ENODEV.
 */
#ifndef B200SPH_H
#define B200SPH_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200SPH_ABI_VERSION 4

/* error codes */
#define B200SPH_OK        0
#define B200SPH_EINVAL   -1  /* invalid argument (reference: std::invalid_argument) */
#define B200SPH_EUNSUP   -2  /* option combination not implemented: fails loudly, never falls back */
#define B200SPH_ECUDA    -3  /* CUDA runtime error (reference: CUDA_SAFE_CALL exception) */
#define B200SPH_ENODEV   -4  /* no CUDA device */
#define B200SPH_ENOMEM   -5

/* enumerations: numeric values are the reference's (src/particledefine.h:79-224,
 * src/visc_spec.h) so that SimParams fields can be passed through unchanged */
enum { B200SPH_KERNEL_CUBICSPLINE = 1, B200SPH_KERNEL_QUADRATIC, B200SPH_KERNEL_WENDLAND, B200SPH_KERNEL_GAUSSIAN };
enum { B200SPH_SPH_F1 = 1, B200SPH_SPH_F2, B200SPH_SPH_GRENIER, B200SPH_SPH_HA };
enum { B200SPH_RHODIFF_NONE = 0, B200SPH_RHODIFF_FERRARI, B200SPH_RHODIFF_COLAGROSSI, B200SPH_RHODIFF_BREZZI };
enum { B200SPH_LJ_BOUNDARY = 0, B200SPH_MK_BOUNDARY, B200SPH_SA_BOUNDARY, B200SPH_DYN_BOUNDARY };
enum { B200SPH_PERIODIC_X = 1, B200SPH_PERIODIC_Y = 2, B200SPH_PERIODIC_Z = 4 };
/* viscous specification (src/visc_spec.h) */
enum { B200SPH_RHEOLOGY_INVISCID = 0, B200SPH_RHEOLOGY_NEWTONIAN = 1 };
enum { B200SPH_TURB_LAMINAR = 0, B200SPH_TURB_ARTIFICIAL = 1 };
enum { B200SPH_COMPVISC_KINEMATIC = 0, B200SPH_COMPVISC_DYNAMIC = 1 };
enum { B200SPH_VISCMODEL_MORRIS = 0, B200SPH_VISCMODEL_MONAGHAN = 1, B200SPH_VISCMODEL_ESPANOL_REVENGA = 2 };
enum { B200SPH_AVG_ARITHMETIC = 0, B200SPH_AVG_HARMONIC = 1, B200SPH_AVG_GEOMETRIC = 2 };

/* simulation flags; numeric values are the reference's (src/simflags.h:62-160) so that the WHOLE of SimParams::simflags
 * is passed through. b200sph_validate refuses (B200SPH_EUNSUP) every bit outside B200SPH_SUPPORTED_SIMFLAGS: nothing is
 * silently ignored. ENABLE_REPACKING only makes the reference's `--repack` run mode available (src/main.cc:357); that
 * run mode itself is refused by the adapter (RunMode REPACK), the flag has no effect on a simulation. */
#define B200SPH_ENABLE_DTADAPT          (1u << 0)
#define B200SPH_ENABLE_XSPH             (1u << 1)
#define B200SPH_ENABLE_PLANES           (1u << 2)
#define B200SPH_ENABLE_DEM              (1u << 3)
#define B200SPH_ENABLE_MOVING_BODIES    (1u << 4)
#define B200SPH_ENABLE_INLET_OUTLET     (1u << 5)
#define B200SPH_ENABLE_WATER_DEPTH      (1u << 6)
#define B200SPH_ENABLE_DENSITY_SUM      (1u << 7)
#define B200SPH_ENABLE_GAMMA_QUADRATURE (1u << 8)
#define B200SPH_ENABLE_REPACKING        (1u << 9)
#define B200SPH_ENABLE_INTERNAL_ENERGY  (1u << 10)
#define B200SPH_ENABLE_MULTIFLUID       (1u << 11)
#define B200SPH_SUPPORTED_SIMFLAGS (B200SPH_ENABLE_DTADAPT | B200SPH_ENABLE_XSPH | B200SPH_ENABLE_PLANES | \
	B200SPH_ENABLE_MOVING_BODIES | B200SPH_ENABLE_REPACKING | B200SPH_ENABLE_MULTIFLUID)

#define B200SPH_MAX_FLUIDS 4
#define B200SPH_MAX_PLANES 8   /* src/particledefine.h:325 */

/* particle type / flag bits (src/particleinfo.h:135-165) */
#define B200SPH_PT_FLUID     0
#define B200SPH_PT_BOUNDARY  1
#define B200SPH_PT_VERTEX    2
#define B200SPH_PT_TESTPOINT 3
#define B200SPH_FG_COMPUTE_FORCE   (1u << 3)
#define B200SPH_FG_MOVING_BOUNDARY (1u << 4)
#define B200SPH_FG_SURFACE         (1u << 9)

/*
 * Everything the three reference engines receive through setconstants()
 * (src/cuda/buildneibs.cu:64-98, src/cuda/forces.cu:269-420, src/cuda/euler.cu:52-95),
 * i.e. the relevant subset of SimParams (src/simparams.h) and PhysParams
 * (src/physparams.h), flattened into one POD.
 */
typedef struct b200sph_params {
	uint32_t abi_version;          /* must be B200SPH_ABI_VERSION */
	/* cell grid (src/cuda/cellgrid.cuh:66-69) */
	float    world_origin[3];
	float    cell_size[3];
	uint32_t grid_size[3];
	uint32_t coord[3];             /* linearisation: axis index (0=x,1=y,2=z) of COORD1,2,3; reference default yzx = {1,2,0} (src/linearization.h) */
	uint32_t periodic;             /* Periodicity bitmask */
	/* neighbour list (src/simparams.h neiblistsize/neibboundpos; src/cuda/buildneibs.cu:64-98) */
	uint32_t neiblistsize;         /* rows of the list, DamBreak3D: 128 */
	uint32_t neibboundpos;         /* row of the first boundary neighbour (grows downwards), = neiblistsize-1 unless SA */
	uint32_t neiblist_stride;      /* allocated particles: the list holds neiblistsize * neiblist_stride entries (layout: see B200SPH_NEIBLIST_BLOCK) */
	float    nl_sq_influence_radius; /* squared neighbour-search radius (simparams nlSqInfluenceRadius) */
	/* SPH */
	uint32_t kerneltype, sph_formulation, densitydiffusiontype, boundarytype;
	uint32_t rheologytype, turbmodel, compvisc, viscmodel, viscavgop;
	uint32_t is_const_visc;        /* FullViscSpec::is_const_visc (src/visc_spec.h:262) */
	float    slength;              /* smoothing length h */
	float    influenceradius;      /* kernel radius * h */
	float    deltap;
	float    density_diff_coeff;   /* d_densityDiffCoeff, already scaled as the reference does (src/cuda/forces.cu:395-410) */
	float    dtadaptfactor;
	/* physics (src/physparams.h) */
	uint32_t num_fluids;
	float    rho0[B200SPH_MAX_FLUIDS];
	float    bcoeff[B200SPH_MAX_FLUIDS];      /* rho0 c0^2 / gamma */
	float    gammacoeff[B200SPH_MAX_FLUIDS];
	float    sscoeff[B200SPH_MAX_FLUIDS];     /* c0 */
	float    sspowercoeff[B200SPH_MAX_FLUIDS];/* (gamma-1)/2 */
	float    visccoeff[B200SPH_MAX_FLUIDS];   /* kinematic (or dynamic) viscosity as uploaded to d_visccoeff */
	float    gravity[3];
	float    artvisccoeff;
	float    epsartvisc;
	/* adaptive dt (src/GPUWorker.cc:3006-3011) */
	float    max_sound_speed_cfl;  /* 1.1 * max c0 */
	float    max_kinvisc;          /* max kinematic viscosity (0 if inviscid) */
	uint32_t dtadapt;              /* ENABLE_DTADAPT */
	/* ---- ABI version 2 ---- */
	uint32_t simflags;             /* the whole of SimParams::simflags (see B200SPH_SUPPORTED_SIMFLAGS) */
	float    epsxsph;              /* PhysParams::epsxsph (src/cuda/euler.cu:56) */
	float    monaghan_visc_coeff;  /* PhysParams::monaghan_visc_coeff = 2(d+2) (src/cuda/forces.cu:334) */
	float    visc2coeff[B200SPH_MAX_FLUIDS]; /* bulk viscosity, ESPANOL_REVENGA only (src/cuda/forces.cu:328) */
	/* Lennard-Jones repulsion of geometric planes (src/cuda/forces.cu:339-368; partsurf 0 => r0^2) */
	float    r0, dcoeff, p1coeff, p2coeff, partsurf;
	/* ---- ABI version 4 ---- */
	uint32_t neiblist_block;       /* particles per block of the neighbour-list layout (power of two >= 32); 0 = B200SPH_NEIBLIST_BLOCK */
} b200sph_params;

/* mirror of TimingInfo's neighbour counters (src/timing.h:42-97) */
typedef struct b200sph_neibs_info {
	int32_t num_interactions;
	int32_t max_fluid_boundary_neibs;
	int32_t max_vertex_neibs;
	int32_t has_too_many_neibs;     /* id of one offending particle, or -1 */
	int32_t has_max_neibs[3];       /* its per-type neighbour counts */
} b200sph_neibs_info;

typedef struct b200sph_ctx b200sph_ctx;

const char *b200sph_last_error(void);
int b200sph_abi_version(void);
/* number of CUDA devices visible (0 on a CPU box; never fails) */
int b200sph_device_count(void);

/* ---- context = the engines' per-device constants and scratch ------------- */

/* replaces {CUDANeibsEngine,CUDAForcesEngine,CUDAPredCorrEngine}::setconstants
 * (src/engine_neibs.h:52, src/engine_forces.h:49, src/engine_integration.h:47).
 * Fails with B200SPH_EUNSUP for option combinations that are not implemented. */
int b200sph_create(const b200sph_params *params, b200sph_ctx **out);
int b200sph_destroy(b200sph_ctx *ctx);
/* check only: would b200sph_create accept these params? (no device needed) */
int b200sph_validate(const b200sph_params *params);
/* cudaStream_t as an opaque pointer; NULL = legacy default stream */
int b200sph_set_stream(b200sph_ctx *ctx, void *cuda_stream);
/* AbstractForcesEngine::setgravity (src/engine_forces.h:58) */
int b200sph_set_gravity(b200sph_ctx *ctx, const float gravity[3]);
/* AbstractForcesEngine::setplanes (src/engine_forces.h:56; src/cuda/forces.cu:443-447). Host arrays of numplanes
 * (<= B200SPH_MAX_PLANES) entries laid out like plane_t (src/planes.h:42-46): unit normal, cell of the reference
 * point, in-cell position of the reference point. Planes act on fluid particles in the finalize stage of
 * b200sph_forces when params.simflags has B200SPH_ENABLE_PLANES (src/cuda/forces_kernel.def:4105-4110). */
int b200sph_set_planes(b200sph_ctx *ctx, const float *normals, const int *grid_pos, const float *pos, int numplanes);
/* AbstractNeibsEngine::getconstants (src/engine_neibs.h:57): returns neibboundpos */
int b200sph_get_neibboundpos(const b200sph_ctx *ctx, uint32_t *neibboundpos);

/* ---- neighbour engine ---------------------------------------------------- */

/* AbstractNeibsEngine::calcHash (src/engine_neibs.h:66; kernel src/cuda/buildneibs_kernel.cu:664-776).
 * pos, hash updated in place; part_index written. compact_dev_map may be NULL. Bit-exact. */
int b200sph_calc_hash(b200sph_ctx *ctx, void *pos, uint32_t *hash, uint32_t *part_index,
	const void *info, const uint32_t *compact_dev_map, uint32_t num_particles);

/* AbstractNeibsEngine::fixHash (src/engine_neibs.h:71; kernel src/cuda/buildneibs_kernel.cu:790-814). */
int b200sph_fix_hash(b200sph_ctx *ctx, uint32_t *hash, uint32_t *part_index,
	const void *info, const uint32_t *compact_dev_map, uint32_t num_particles);

/* AbstractNeibsEngine::sort (src/engine_neibs.h:84; src/cuda/buildneibs.cu:358-415).
 * Sorts (hash, info) in place by the reference's total order
 * (hash incl. high bits, particle type, id) and permutes part_index alike. Bit-exact. */
int b200sph_sort(b200sph_ctx *ctx, uint32_t *hash, void *info, uint32_t *part_index,
	uint32_t num_particles);

/* one extra per-particle array to be permuted by reorder (the reference gathers up to
 * 11 optional buffers, src/cuda/buildneibs_kernel.cu:840-992) */
typedef struct b200sph_reorder_extra {
	const void *unsorted;
	void       *sorted;
	uint32_t    elem_size;   /* 4, 8 or 16 bytes */
} b200sph_reorder_extra;

/* AbstractNeibsEngine::reorderDataAndFindCellStart (src/engine_neibs.h:76; kernel
 * src/cuda/buildneibs_kernel.cu:840-992). cell_start/cell_end must have been pre-filled
 * with 0xFF by the caller (as GPUWorker does, src/GPUWorker.cc:1846).
 * segment_start (4 uints, device) may be NULL. new_num_particles is a device pointer. */
int b200sph_reorder(b200sph_ctx *ctx, uint32_t *cell_start, uint32_t *cell_end,
	uint32_t *segment_start,
	void *sorted_pos, void *sorted_vel,
	const void *unsorted_pos, const void *unsorted_vel,
	const b200sph_reorder_extra *extras, uint32_t num_extras,
	const void *sorted_info, const uint32_t *sorted_hash, const uint32_t *part_index,
	uint32_t num_particles, uint32_t *new_num_particles);

/* AbstractNeibsEngine::resetinfo / getinfo (src/engine_neibs.h:60-63; src/cuda/buildneibs.cu:118-146) */
int b200sph_neibs_resetinfo(b200sph_ctx *ctx);
int b200sph_neibs_getinfo(b200sph_ctx *ctx, b200sph_neibs_info *out);

/* Layout of the neighbour list. The reference interleaves the columns of ALL particles: entry k of particle i at
 * list[k * stride + i] (src/cuda/neibs_iteration.cuh:60-75), so consecutive entries of a particle lie stride * 2
 * bytes apart - 16 MB at 8 M particles. The list builder scatters the rows of a warp over that distance and was
 * measured to lose a quarter of the neighbour rebuild to it (7.7 ms at 7.9 M particles against 5.7 ms with rows at
 * most 4 MB apart; nothing at 2 M particles, where the rows are 4 MB apart anyway), while the pair kernel wants
 * concurrently running CTAs to read neighbouring memory, i.e. the interleaved layout (2-4 % slower with blocks of
 * 128 ... 256 k particles). The list is private to the three engines (nothing else in the reference reads
 * BUFFER_NEIBSLIST on this path), so here the columns are interleaved per BLOCK of B = neiblist_block consecutive
 * particles: with b = i / B and w = min(B, stride - B b) (only the last block can be narrower),
 *     entry k of particle i  =  list[B * b * neiblistsize + k * w + (i - B b)].
 * Up to B particles this IS the reference's layout; the buffer always has the reference's size (neiblistsize *
 * stride entries). Entry VALUES (cell markers, offsets inside the cell, end markers, section placement) are the
 * reference's, bit for bit. */
#ifndef B200SPH_NEIBLIST_BLOCK
#define B200SPH_NEIBLIST_BLOCK 2097152
#endif

/* AbstractNeibsEngine::buildNeibsList (src/engine_neibs.h:89; kernel
 * src/cuda/buildneibs_kernel.cu:1029-1185). neibs_list must have been pre-filled with 0xFF
 * by the caller (src/GPUWorker.cc:1883). List contents are bit-exact with the reference
 * (in the blocked layout above). */
int b200sph_build_neibs(b200sph_ctx *ctx, const void *pos, const void *info,
	const uint32_t *hash, const uint32_t *cell_start, const uint32_t *cell_end,
	uint16_t *neibs_list, uint32_t num_particles, uint32_t particle_range_end);

/* ---- forces engine ------------------------------------------------------- */

/* AbstractForcesEngine::getFmaxElements / getFmaxTempElements / round_particles
 * (src/engine_forces.h:157-161,108; src/cuda/forces.cu:540-554,961-965) */
uint32_t b200sph_fmax_elements(uint32_t n);
uint32_t b200sph_fmax_temp_elements(uint32_t n);
uint32_t b200sph_round_particles(uint32_t n);

/* AbstractForcesEngine::basicstep (src/engine_forces.h:133; src/cuda/forces.cu:901-932,718-799):
 * the reference's forcesDevice<F,F> + <F,B> + <B,F> + finalizeforcesDevice in ONE fused
 * launch. forces must have been zeroed by the caller (src/GPUWorker.cc:1949) - the result
 * overwrites it. cfl receives one max per 128-particle block at cfl[cfl_offset + block];
 * the number of blocks written is returned through num_cfl_blocks (the reference returns it). */
int b200sph_forces(b200sph_ctx *ctx, const void *pos, const void *vel, const void *info,
	const uint32_t *hash, const uint32_t *cell_start, const uint16_t *neibs_list,
	void *forces, float *cfl,
	uint32_t num_particles, uint32_t from_particle, uint32_t to_particle,
	uint32_t cfl_offset, uint32_t *num_cfl_blocks);

/* Diagnostic (no reference counterpart): the per-particle {P/rho^2, sound speed} pairs the forces
 * kernel uses, evaluated with the same device __powf expressions as the reference's P() and
 * soundSpeed() (src/cuda/phys_core.cu:99-136). out = float2[num_particles] (device). Lets a
 * CPU checker separate the GPU's approximate-pow rounding from the summation logic. */
int b200sph_eos_probe(b200sph_ctx *ctx, const void *vel, const void *info, void *out, uint32_t num_particles);

/* AbstractForcesEngine::dtreduce (src/engine_forces.h:163; src/cuda/forces.cu:557-607).
 * Synchronises the stream and returns dt on the host, like the reference. b200sph_dtreduce uses the values given at
 * context creation; b200sph_dtreduce_ex takes the four arguments the reference passes on every call (GPUWorker updates
 * max_kinematic at run time, src/GPUWorker.cc:2642; sspeed_cfl is the product 1.1 * max c0 evaluated in double and
 * rounded once, src/GPUWorker.cc:3011). */
int b200sph_dtreduce(b200sph_ctx *ctx, const float *cfl, float *temp_cfl,
	uint32_t num_blocks, float *dt_out);
int b200sph_dtreduce_ex(b200sph_ctx *ctx, const float *cfl, float *temp_cfl, uint32_t num_blocks,
	float slength, float dtadaptfactor, float sspeed_cfl, float max_kinematic, float *dt_out);

/* The two halves of dtreduce for callers that combine CFL maxima across devices ON the device (multi-GPU: one
 * NCCL all-reduce(MAX) on `max_out` instead of a host loop over per-device dt, src/GPUSPH.cc:650-657):
 * b200sph_cflmax writes max(cfl[0..num_blocks)) to the DEVICE float *max_out without synchronising;
 * b200sph_dt_from_cfl applies the reference's formula (src/cuda/forces.cu:571-600) to a host value. */
int b200sph_cflmax(b200sph_ctx *ctx, const float *cfl, uint32_t num_blocks, float *max_out);
int b200sph_dt_from_cfl(const b200sph_ctx *ctx, float max_cfl, float *dt_out);

/* ---- moving / force-feedback bodies ------------
 * AbstractForcesEngine::setrbcg / setrbstart (src/engine_forces.h:62-66; src/cuda/forces.cu:430-447) and
 * AbstractIntegrationEngine::setrbcg / setrbtrans / setrbsteprot / setrblinearvel / setrbangularvel
 * (src/engine_integration.h:54-68; src/cuda/euler.cu:76-95). Host arrays, at most B200SPH_MAX_BODIES bodies.
 * cg_grid_pos: int[3*n] cell of each centre of gravity, cg_pos: float[3*n] in-cell coordinate. */
#define B200SPH_MAX_BODIES 16
/* The centres of gravity are kept TWICE, like the reference's two __constant__ copies (src/cuda/forces_kernel.cu:81-83,
 * src/cuda/euler_kernel.cu:45-50): b200sph_set_rbcg is AbstractForcesEngine::setrbcg (torque arm in finalize),
 * b200sph_set_rbcg_euler is AbstractIntegrationEngine::setrbcg (centre of the rigid motion). The reference's integrator
 * moves the forces copy to cg(n+1) in the middle of a step while the integration keeps cg(n)
 * (src/integrators/PredictorCorrectorIntegrator.cc:332,556-587). */
int b200sph_set_rbcg(b200sph_ctx *ctx, const int *cg_grid_pos, const float *cg_pos, int numbodies);
int b200sph_set_rbcg_euler(b200sph_ctx *ctx, const int *cg_grid_pos, const float *cg_pos, int numbodies);
int b200sph_set_rbstart(b200sph_ctx *ctx, const int *rbfirstindex, int numbodies);
int b200sph_set_rbtrans(b200sph_ctx *ctx, const float *trans, int numbodies);
int b200sph_set_rbsteprot(b200sph_ctx *ctx, const float *rot /* 9 per body */, int numbodies);
int b200sph_set_rblinearvel(b200sph_ctx *ctx, const float *linearvel, int numbodies);
int b200sph_set_rbangularvel(b200sph_ctx *ctx, const float *angularvel, int numbodies);

/* b200sph_forces with compute_object_forces = true (src/engine_forces.h:133-149): additionally scatters, for every
 * particle flagged FG_COMPUTE_FORCE, force x mass and the torque about its body's centre of gravity into
 * rb_forces / rb_torques (float4[num body particles], index = id + rbfirstindex[object]) exactly like
 * finalizeforcesDevice (src/cuda/forces_kernel.def:4116-4141). rb_forces == NULL: same as b200sph_forces. */
int b200sph_forces_bodies(b200sph_ctx *ctx, const void *pos, const void *vel, const void *info,
	const uint32_t *hash, const uint32_t *cell_start, const uint16_t *neibs_list,
	void *forces, float *cfl, void *rb_forces, void *rb_torques,
	uint32_t num_particles, uint32_t from_particle, uint32_t to_particle,
	uint32_t cfl_offset, uint32_t *num_cfl_blocks);

/* The complete AbstractForcesEngine::basicstep (src/engine_forces.h:133-149): everything b200sph_forces_bodies takes,
 * plus the arguments that only some option combinations read:
 *   xsph   float4[N], written for fluid particles when params.simflags has B200SPH_ENABLE_XSPH: 2 x the mean
 *          neighbourhood velocity (src/cuda/forces_kernel.def:2986-2992, 3366-3368); must have been zeroed by the
 *          caller (src/GPUWorker.cc:1951-1952). NULL otherwise.
 *   dt     the command's dt (dt/2 on the predictor, dt on the corrector, src/GPUWorker.cc:1932), read by the BREZZI
 *          density diffusion only (src/cuda/forces_kernel.def:1765-1782)
 *   step   1 or 2; with dt_from_device != 0, dt is taken from the context's device-resident record instead
 *          (dt/2 for step 1), see "device-resident time stepping" below. */
typedef struct b200sph_forces_args {
	const void *pos, *vel, *info;
	const uint32_t *hash, *cell_start;
	const uint16_t *neibs_list;
	void *forces;
	float *cfl;
	void *rb_forces, *rb_torques;   /* NULL: no body output */
	void *xsph;                     /* NULL unless ENABLE_XSPH */
	uint32_t num_particles, from_particle, to_particle, cfl_offset;
	float dt;
	int step;
	int dt_from_device;
	/* ABI 3. The pair kernel gathers each neighbour as one 32-byte record {pos.xyz, mass, vel.xyz, rho~} (one 256-bit
	 * load). packed == NULL (reference-style calls): the library interleaves pos / vel of [0, num_particles) into its
	 * own scratch in a streaming pre-pass before the launch. packed != NULL: device buffer of 32 bytes x num_particles,
	 * 32-byte aligned, that the CALLER vouches holds exactly the state in pos / vel (written by b200sph_pack_state or
	 * by b200sph_forces_euler's new_packed); the pre-pass is skipped. */
	const void *packed;
} b200sph_forces_args;
int b200sph_forces_ex(b200sph_ctx *ctx, const b200sph_forces_args *args, uint32_t *num_cfl_blocks);

/* b200sph_forces_ex immediately followed by b200sph_euler_ex for the same particles [from_particle, to_particle)
 * (no reference counterpart: the reference runs forcesDevice, finalizeforcesDevice and eulerDevice as separate
 * launches with FORCES round-tripping through memory, src/cuda/forces.cu:718-799, src/cuda/euler.cu:330-372).
 * With the default pair kernel the integration runs in the kernel's epilogue while the particle's forces are still in
 * registers; results are bitwise those of the two separate calls (same update code). old_* = state n (may be the
 * buffers args->pos / args->vel: predictor), new_* receives the integrated state and may alias old_* (in place) but
 * must not alias args->pos / args->vel (nor args->packed), which other particles still gather from - unless the caller
 * provides args->packed, in which case the gathers read the records and pos / vel may be integrated in place.
 * forces is written as usual. */
typedef struct b200sph_fused_euler_args {
	const void *old_pos, *old_vel;
	void *new_pos, *new_vel;
	float dt;                 /* dt of the sub-step (dt/2 for step 1), unless dt_from_device */
	int step;                 /* 1 predictor, 2 corrector */
	int dt_from_device;       /* dt from the device-resident record (dt/2 for step 1) */
	void *new_packed;         /* ABI 3, may be NULL: also receives the integrated state as 32-byte neighbour records
	                             (see b200sph_forces_args.packed) for [from_particle, to_particle) */
} b200sph_fused_euler_args;
int b200sph_forces_euler(b200sph_ctx *ctx, const b200sph_forces_args *args, const b200sph_fused_euler_args *euler,
	uint32_t *num_cfl_blocks);

/* Interleave pos / vel of [from_particle, to_particle) into the 32-byte neighbour records of
 * b200sph_forces_args.packed (no reference counterpart; streaming, 64 bytes per particle). */
int b200sph_pack_state(b200sph_ctx *ctx, const void *pos, const void *vel, void *packed,
	uint32_t from_particle, uint32_t to_particle);

/* AbstractForcesEngine::reduceRbForces (src/engine_forces.h:68-74; src/cuda/forces.cu:967-1003): in-place segmented
 * inclusive scan of rb_forces / rb_torques keyed by rb_keys, then the last element of each body's segment
 * (lastindex[b], host array) is copied to total_force / total_torque (host float[3*numbodies]). Synchronises. */
int b200sph_reduce_rb_forces(b200sph_ctx *ctx, void *rb_forces, void *rb_torques, const uint32_t *rb_keys,
	const uint32_t *lastindex, float *total_force, float *total_torque,
	uint32_t numforcesbodies, uint32_t num_forces_bodies_particles);

/* ---- integration engine -------------------------------------------------- */

/* AbstractIntegrationEngine::basicstep (src/engine_integration.h:117; src/cuda/euler.cu:330-372;
 * kernel src/cuda/euler_kernel.def:396-540). step = 1 (predictor, dt = dt/2 passed by the
 * caller as in the reference) or 2 (corrector). old_* = state n, forces = current state. */
int b200sph_euler(b200sph_ctx *ctx, const void *old_pos, const void *old_vel,
	const void *info, const uint32_t *hash, const void *forces,
	void *new_pos, void *new_vel,
	uint32_t num_particles, uint32_t particle_range_end, float dt, int step);

/* b200sph_euler with the XSPH correction (src/cuda/euler_kernel.def:165-180): velc += epsxsph * xsph[i]. xsph = the
 * buffer b200sph_forces_ex wrote (NULL: same as b200sph_euler). dt_from_device != 0: dt from the device record. */
int b200sph_euler_ex(b200sph_ctx *ctx, const void *old_pos, const void *old_vel,
	const void *info, const uint32_t *hash, const void *forces, const void *xsph,
	void *new_pos, void *new_vel,
	uint32_t num_particles, uint32_t particle_range_end, float dt, int step, int dt_from_device);

/* b200sph_euler_ex that ALSO writes the integrated particles as the pair kernel's 32-byte neighbour records
 * (b200sph_forces_args.packed) into new_packed[0 .. particle_range_end) (no reference counterpart; saves the
 * b200sph_pack_state pass before the next force evaluation). new_packed == NULL: exactly b200sph_euler_ex. */
int b200sph_euler_packed(b200sph_ctx *ctx, const void *old_pos, const void *old_vel,
	const void *info, const uint32_t *hash, const void *forces, const void *xsph,
	void *new_pos, void *new_vel, void *new_packed,
	uint32_t num_particles, uint32_t particle_range_end, float dt, int step, int dt_from_device);
/* the inverse of b200sph_pack_state: records of [from_particle, to_particle) back into pos / vel (multi-GPU: the halo
 * particles only live as records between neighbour rebuilds, gpusph_b200/multigpu.py) */
int b200sph_unpack_state(b200sph_ctx *ctx, const void *packed, void *pos, void *vel,
	uint32_t from_particle, uint32_t to_particle);

/* ---- filter engines and post-processing ---------
 * AbstractFilterEngine::process for SHEPARD_FILTER and MLS_FILTER (src/engine_filter.h:75-82;
 * src/cuda/forces.cu:1026-1122; kernels src/cuda/forces_kernel.cu:418-507, 509-721): density of every fluid
 * particle (MLS: of every particle) re-initialised from its neighbours through the neighbour list; old_vel is read,
 * new_vel written for [0, particle_range_end) (other particles' velocities are copied through). */
int b200sph_filter_shepard(b200sph_ctx *ctx, const void *pos, const void *old_vel, void *new_vel, const void *info,
	const uint32_t *hash, const uint32_t *cell_start, const uint16_t *neibs_list,
	uint32_t num_particles, uint32_t particle_range_end);
int b200sph_filter_mls(b200sph_ctx *ctx, const void *pos, const void *old_vel, void *new_vel, const void *info,
	const uint32_t *hash, const uint32_t *cell_start, const uint16_t *neibs_list,
	uint32_t num_particles, uint32_t particle_range_end);

/* AbstractPostProcessEngine::process for TESTPOINTS (src/engine_postprocess.h:85-92; src/cuda/post_process.cu:148-216;
 * kernel src/cuda/post_process_kernel.cu:134-240): every PT_TESTPOINT particle gets the Shepard-normalised velocity
 * (xyz) and pressure (w) of its fluid neighbours, IN PLACE in vel; tke / epsilon (float[N], in place) may be NULL. */
int b200sph_testpoints(b200sph_ctx *ctx, const void *pos, void *vel, float *tke, float *epsilon, const void *info,
	const uint32_t *hash, const uint32_t *cell_start, const uint16_t *neibs_list,
	uint32_t num_particles, uint32_t particle_range_end);

/* ---- device-resident time stepping (no reference counterpart) ---------------
 * The reference reads the CFL maximum back to the host after every force evaluation (two blocking 4-byte copies
 * per step, src/cuda/forces.cu:153-177) and hands dt back down as a kernel argument. These entry points keep
 * {dt, dt candidates, t, iteration} in a small device record owned by the context so that a caller can enqueue
 * whole time steps without synchronising:
 *   b200sph_step_set_dt      host -> device: dt for the next step (first step / after a checkpoint)
 *   b200sph_dtreduce_async   dtreduce, result stored as candidate `which` (1 = predictor, 2 = corrector)
 *   b200sph_euler_async      b200sph_euler with dt taken from the device record (dt/2 for step 1)
 *   b200sph_step_end         t += dt, iteration++, dt = min(candidate 1, candidate 2)
 *                            (src/GPUWorker.cc:2224-2229, src/GPUSPH.cc:636-699)
 *   b200sph_step_query       device -> host (synchronises)
 * Results are identical to the host-dt path: same kernels, same arithmetic. */
int b200sph_step_set_dt(b200sph_ctx *ctx, float dt);
int b200sph_dtreduce_async(b200sph_ctx *ctx, const float *cfl, uint32_t num_blocks, int which);
int b200sph_euler_async(b200sph_ctx *ctx, const void *old_pos, const void *old_vel,
	const void *info, const uint32_t *hash, const void *forces,
	void *new_pos, void *new_vel,
	uint32_t num_particles, uint32_t particle_range_end, int step);
int b200sph_step_end(b200sph_ctx *ctx);
int b200sph_step_query(b200sph_ctx *ctx, double *t, float *dt, uint64_t *iterations);

/* ---- stepping a state that lives in HOST memory (no reference counterpart) ----
 * The reference keeps the particle state on the device and moves it only for writes (GPUWorker::dumpBuffers,
 * src/GPUWorker.cc:1227-1300). A caller whose state lives in (pinned) host memory gets one call per time step:
 * b200sph_step_host uploads state n, runs the command sequence of one predictor-corrector step
 * (src/integrators/PredictorCorrectorIntegrator.cc:917-1068: forces, dt candidate, euler dt/2, forces, dt candidate,
 * euler dt, step end - all with the device-resident dt above) and downloads state n+1 into the same host buffers.
 * The copies are pipelined with the force evaluations in `num_stripes` particle ranges on two extra streams owned by
 * the context; a stripe must consist of whole cell layers along COORD3 (the slowest hash digit) so that every
 * neighbour of a particle of stripe s lies in stripes s-1..s+1, and COORD3 must not be periodic. Consecutive calls on
 * the same buffers are chained stripe by stripe: the upload of stripe s only waits for the previous call's download
 * of stripe s, so the host never has to synchronise between steps. Results are bitwise those of the resident path.
 * Nothing is synchronised: the host buffers hold state n+1 after b200sph_host_sync (or a device synchronisation).
 *   resident != 0   state n is already in pos / vel on the device (it was uploaded with b200sph_host_upload and
 *                   re-sorted by a neighbour rebuild): no upload, only the striped downloads
 * b200sph_host_upload: the upload half on its own, for the steps that start with a neighbour rebuild (the sort needs
 * the whole state): chained on the previous call's downloads like above; the context's stream waits for it.
 * b200sph_host_fence: the context's stream waits for the copies still in flight; call it before using the state
 * buffers through any other entry point. */
#define B200SPH_MAX_STRIPES 32
typedef struct b200sph_host_step_args {
	void *host_pos, *host_vel;          /* pinned host float4[num_particles]: state n in, state n+1 out */
	void *pos, *vel;                    /* device float4[N]: receive state n, end up holding state n+1 */
	void *pos_star, *vel_star;          /* device float4[N]: scratch for the predicted state n* */
	const void *info;
	const uint32_t *hash, *cell_start;
	const uint16_t *neibs_list;
	void *forces;                       /* device float4[N] scratch */
	float *cfl;                         /* device float[cfl_elements]: room for both force evaluations, every stripe's
	                                       blocks rounded up to 4: 2 x sum_s round4(ceil(stripe_s / 128)) */
	void *xsph;                         /* NULL unless ENABLE_XSPH */
	uint32_t cfl_elements;
	uint32_t num_particles;
	const uint32_t *stripe_bounds;      /* host array [num_stripes + 1]: 0 = b[0] < b[1] < ... < b[num_stripes] = num_particles */
	uint32_t num_stripes;               /* 1 .. B200SPH_MAX_STRIPES */
	int resident;
} b200sph_host_step_args;
int b200sph_step_host(b200sph_ctx *ctx, const b200sph_host_step_args *args);
int b200sph_host_upload(b200sph_ctx *ctx, const void *host_pos, const void *host_vel, void *pos, void *vel,
	uint32_t num_particles);
int b200sph_host_fence(b200sph_ctx *ctx);
int b200sph_host_sync(b200sph_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* B200SPH_H */
