// euler.cu — integration engine: predictor/corrector update.
// Behavioural specification: GPUSPH eulerDevice, src/cuda/euler_kernel.def:396-540
// (corrected velocity :117-134, continuity :200-206). Streaming kernel: 60 B in, 32 B out per particle.
#include "common.cuh"
#include "euler_update.cuh"

template<int STEP>
__global__ void __launch_bounds__(BLOCK_STREAM)
euler_kernel(const __grid_constant__ DevParams P, const float4 *__restrict__ oldPos, const float4 *__restrict__ oldVel,
	const ushort4 *__restrict__ infoArray, const uint *__restrict__ particleHash, const float4 *__restrict__ forces,
	float4 *__restrict__ newPos, float4 *__restrict__ newVel, const uint numParticles, const float dt_arg,
	const StepState *__restrict__ dev_state, const BodyData *__restrict__ bodies, const float4 *__restrict__ xsph,
	PosVel *__restrict__ newPacked)
{
	const uint index = blockIdx.x * blockDim.x + threadIdx.x;
	if (index >= numParticles) return;
	const float dt = euler_dt<STEP>(dev_state, dt_arg);
	float4 pos = oldPos[index];
	float4 vel = oldVel[index];
	const float4 force = forces[index];
	const ushort4 info = infoArray[index];
	euler_update<STEP>(P, pos, vel, force, info, particleHash, index, dt, bodies, xsph != NULL,
		xsph ? xsph[index] : make_float4(0.f, 0.f, 0.f, 0.f));
	newPos[index] = pos;
	newVel[index] = vel;
	// the integrated state as the pair kernel's neighbour record (forces.cu), when the caller keeps records
	if (newPacked) st_posvel(newPacked + index, pos, vel);
}

// records -> pos / vel (the inverse of pack_state_kernel, forces.cu)
__global__ void __launch_bounds__(BLOCK_STREAM)
unpack_state_kernel(const PosVel *__restrict__ pv, float4 *__restrict__ pos, float4 *__restrict__ vel, const uint from, const uint to)
{
	const uint i = from + blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= to) return;
	float4 p, v;
	ld_posvel(pv + i, p, v);
	pos[i] = p; vel[i] = v;
}

extern "C" int b200sph_unpack_state(b200sph_ctx *ctx, const void *packed, void *pos, void *vel, uint32_t from, uint32_t to)
{
	CHECK_CTX(ctx);
	if (to <= from) return B200SPH_OK;
	if (!pos || !vel || !packed) { b200_set_error("unpack_state: null buffer"); return B200SPH_EINVAL; }
	if ((uintptr_t)packed & 31u) { b200_set_error("unpack_state: the record buffer must be 32-byte aligned"); return B200SPH_EINVAL; }
	unpack_state_kernel<<<div_up(to - from, BLOCK_STREAM), BLOCK_STREAM, 0, ctx->stream>>>((const PosVel *)packed, (float4 *)pos, (float4 *)vel, from, to);
	KERNEL_TRY();
	return B200SPH_OK;
}

static int launch_euler(b200sph_ctx *ctx, const void *old_pos, const void *old_vel, const void *info,
	const uint32_t *hash, const void *forces, const void *xsph, void *new_pos, void *new_vel,
	uint32_t num_particles, uint32_t particle_range_end, float dt, int step, int dt_from_device, void *new_packed = NULL)
{
	CHECK_CTX(ctx);
	if (step != 1 && step != 2) { b200_set_error("unsupported predcorr timestep %d", step); return B200SPH_EINVAL; }   // euler.cu:361-362
	if (particle_range_end == 0) return B200SPH_OK;
	const BodyData *bodies = NULL;
	{ const int rc = b200_euler_bodies(ctx, hash, &bodies); if (rc) return rc; }
	if (!old_pos || !old_vel || !info || !forces || !new_pos || !new_vel) { b200_set_error("euler: null buffer"); return B200SPH_EINVAL; }
	if ((uintptr_t)new_packed & 31u) { b200_set_error("euler: new_packed must be 32-byte aligned"); return B200SPH_EINVAL; }
	if ((ctx->hp.simflags & B200SPH_ENABLE_XSPH) && !xsph) { b200_set_error("euler: ENABLE_XSPH needs the xsph buffer"); return B200SPH_EINVAL; }
	if (!(ctx->hp.simflags & B200SPH_ENABLE_XSPH)) xsph = NULL;
	const uint nb = div_up(particle_range_end, BLOCK_STREAM);
	// NOTE the reference passes numParticles (not the range end) as the kernel bound, euler.cu:352
	const uint bound = num_particles < particle_range_end ? num_particles : particle_range_end;
	const StepState *st = dt_from_device ? ctx->d_step : NULL;
	if (step == 1)
		euler_kernel<1><<<nb, BLOCK_STREAM, 0, ctx->stream>>>(ctx->dp, (const float4 *)old_pos, (const float4 *)old_vel,
			(const ushort4 *)info, hash, (const float4 *)forces, (float4 *)new_pos, (float4 *)new_vel, bound, dt, st, bodies, (const float4 *)xsph, (PosVel *)new_packed);
	else
		euler_kernel<2><<<nb, BLOCK_STREAM, 0, ctx->stream>>>(ctx->dp, (const float4 *)old_pos, (const float4 *)old_vel,
			(const ushort4 *)info, hash, (const float4 *)forces, (float4 *)new_pos, (float4 *)new_vel, bound, dt, st, bodies, (const float4 *)xsph, (PosVel *)new_packed);
	KERNEL_TRY();
	return B200SPH_OK;
}

extern "C" int b200sph_euler(b200sph_ctx *ctx, const void *old_pos, const void *old_vel, const void *info,
	const uint32_t *hash, const void *forces, void *new_pos, void *new_vel,
	uint32_t num_particles, uint32_t particle_range_end, float dt, int step)
{
	if (ctx && (ctx->hp.simflags & B200SPH_ENABLE_XSPH)) { b200_set_error("euler: ENABLE_XSPH needs b200sph_euler_ex with the xsph buffer"); return B200SPH_EINVAL; }
	return launch_euler(ctx, old_pos, old_vel, info, hash, forces, NULL, new_pos, new_vel, num_particles, particle_range_end, dt, step, 0);
}

extern "C" int b200sph_euler_async(b200sph_ctx *ctx, const void *old_pos, const void *old_vel, const void *info,
	const uint32_t *hash, const void *forces, void *new_pos, void *new_vel,
	uint32_t num_particles, uint32_t particle_range_end, int step)
{
	if (ctx && (ctx->hp.simflags & B200SPH_ENABLE_XSPH)) { b200_set_error("euler: ENABLE_XSPH needs b200sph_euler_ex with the xsph buffer"); return B200SPH_EINVAL; }
	return launch_euler(ctx, old_pos, old_vel, info, hash, forces, NULL, new_pos, new_vel, num_particles, particle_range_end, 0.0f, step, 1);
}

extern "C" int b200sph_euler_ex(b200sph_ctx *ctx, const void *old_pos, const void *old_vel, const void *info,
	const uint32_t *hash, const void *forces, const void *xsph, void *new_pos, void *new_vel,
	uint32_t num_particles, uint32_t particle_range_end, float dt, int step, int dt_from_device)
{
	return launch_euler(ctx, old_pos, old_vel, info, hash, forces, xsph, new_pos, new_vel, num_particles, particle_range_end, dt, step, dt_from_device);
}

// b200sph_euler_ex that also writes the integrated particles as the pair kernel's 32-byte neighbour records
extern "C" int b200sph_euler_packed(b200sph_ctx *ctx, const void *old_pos, const void *old_vel, const void *info,
	const uint32_t *hash, const void *forces, const void *xsph, void *new_pos, void *new_vel, void *new_packed,
	uint32_t num_particles, uint32_t particle_range_end, float dt, int step, int dt_from_device)
{
	return launch_euler(ctx, old_pos, old_vel, info, hash, forces, xsph, new_pos, new_vel, num_particles, particle_range_end, dt, step,
		dt_from_device, new_packed);
}
