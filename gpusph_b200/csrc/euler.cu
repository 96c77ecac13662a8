// euler.cu — integration engine: predictor/corrector update.
// Behavioural specification: GPUSPH eulerDevice, src/cuda/euler_kernel.def:396-540
// (corrected velocity :117-134, continuity :200-206). Streaming kernel: 60 B in, 32 B out per particle.
#include "common.cuh"

template<int STEP>
__global__ void __launch_bounds__(BLOCK_STREAM)
euler_kernel(const __grid_constant__ DevParams P, const float4 *__restrict__ oldPos, const float4 *__restrict__ oldVel,
	const ushort4 *__restrict__ infoArray, const uint *__restrict__ particleHash, const float4 *__restrict__ forces,
	float4 *__restrict__ newPos, float4 *__restrict__ newVel, const uint numParticles, const float dt_arg,
	const StepState *__restrict__ dev_state, const BodyData *__restrict__ bodies, const float4 *__restrict__ xsph)
{
	const uint index = blockIdx.x * blockDim.x + threadIdx.x;
	if (index >= numParticles) return;
	// dt either as an argument (reference-style call) or from the device-resident record: dt/2 on the predictor
	// (PredictorCorrector::getDtOperatorForStep), dt on the corrector
	const float dt = dev_state ? (STEP == 1 ? dev_state->dt / 2 : dev_state->dt) : dt_arg;
	float4 pos = oldPos[index];
	float4 vel = oldVel[index];
	const float4 force = forces[index];
	const ushort4 info = infoArray[index];
	const int type = ptype_of(info);
	const bool integrateBoundary = (P.boundarytype == B200SPH_DYN_BOUNDARY || P.boundarytype == B200SPH_SA_BOUNDARY);   // :424-425
	if (!inactive_w(pos.w) && !(type == PT_BOUNDARY && !integrateBoundary && !(info.x & B200SPH_FG_MOVING_BOUNDARY))) {
		// velc = vel (+ force*dt/2 on the corrector), :117-134
		float vcx = vel.x, vcy = vel.y, vcz = vel.z;
		if (STEP == 2) {
			const float hdt = dt / 2;
			vcx += force.x * hdt; vcy += force.y * hdt; vcz += force.z * hdt;
		}
		if (xsph) {                                               // XSPH correction, :165-180
			const float4 mv = xsph[index];
			vcx += P.epsxsph * mv.x; vcy += P.epsxsph * mv.y; vcz += P.epsxsph * mv.z;
		}
		if (type == PT_FLUID) {                                   // :441-462
			pos.x += vcx * dt; pos.y += vcy * dt; pos.z += vcz * dt;
			vel.w += dt * force.w;
			vel.x += dt * force.x; vel.y += dt * force.y; vel.z += dt * force.z;
		} else if (type == PT_BOUNDARY || type == PT_VERTEX) {     // :468-512
			// particles of a moving / floating body follow the rigid motion of the body (:470-503)
			if ((info.x & B200SPH_FG_MOVING_BOUNDARY) && bodies) {
				const int obj = object_of_y(info.y);
				const int3 gp = grid_pos(P, particleHash[index] & CELLTYPE_BITMASK);
				// relPos = x - x_cg (globalDistance, cellgrid.cuh:152-160)
				const float rx = (float)(gp.x - bodies->cgGridPos[obj][0]) * P.cellSize[0] + (pos.x - bodies->cgPos[obj][0]);
				const float ry = (float)(gp.y - bodies->cgGridPos[obj][1]) * P.cellSize[1] + (pos.y - bodies->cgPos[obj][1]);
				const float rz = (float)(gp.z - bodies->cgGridPos[obj][2]) * P.cellSize[2] + (pos.z - bodies->cgPos[obj][2]);
				const float *rot = bodies->steprot[obj];
				// applyrot, euler_kernel.cu:67-74
				pos.x += (rot[0] - 1.0f) * rx + rot[1] * ry + rot[2] * rz;
				pos.y += rot[3] * rx + (rot[4] - 1.0f) * ry + rot[5] * rz;
				pos.z += rot[6] * rx + rot[7] * ry + (rot[8] - 1.0f) * rz;
				pos.x += bodies->trans[obj][0]; pos.y += bodies->trans[obj][1]; pos.z += bodies->trans[obj][2];
				// V(P) = V(Cg) + omega x PCg
				const float *w = bodies->angularvel[obj], *lv = bodies->linearvel[obj];
				vel.x = lv[0] + (w[1] * rz - w[2] * ry);
				vel.y = lv[1] + (w[2] * rx - w[0] * rz);
				vel.z = lv[2] + (w[0] * ry - w[1] * rx);
			}
			if (P.boundarytype == B200SPH_DYN_BOUNDARY) vel.w += dt * force.w;
		}
	}
	newPos[index] = pos;
	newVel[index] = vel;
}

static int launch_euler(b200sph_ctx *ctx, const void *old_pos, const void *old_vel, const void *info,
	const uint32_t *hash, const void *forces, const void *xsph, void *new_pos, void *new_vel,
	uint32_t num_particles, uint32_t particle_range_end, float dt, int step, int dt_from_device)
{
	CHECK_CTX(ctx);
	if (step != 1 && step != 2) { b200_set_error("unsupported predcorr timestep %d", step); return B200SPH_EINVAL; }   // euler.cu:361-362
	const BodyData *bodies = (ctx->have_bodies && hash) ? ctx->d_bodies : NULL;
	if (particle_range_end == 0) return B200SPH_OK;
	if (!old_pos || !old_vel || !info || !forces || !new_pos || !new_vel) { b200_set_error("euler: null buffer"); return B200SPH_EINVAL; }
	if ((ctx->hp.simflags & B200SPH_ENABLE_XSPH) && !xsph) { b200_set_error("euler: ENABLE_XSPH needs the xsph buffer"); return B200SPH_EINVAL; }
	if (!(ctx->hp.simflags & B200SPH_ENABLE_XSPH)) xsph = NULL;
	const uint nb = div_up(particle_range_end, BLOCK_STREAM);
	// NOTE the reference passes numParticles (not the range end) as the kernel bound, euler.cu:352
	const uint bound = num_particles < particle_range_end ? num_particles : particle_range_end;
	const StepState *st = dt_from_device ? ctx->d_step : NULL;
	if (step == 1)
		euler_kernel<1><<<nb, BLOCK_STREAM, 0, ctx->stream>>>(ctx->dp, (const float4 *)old_pos, (const float4 *)old_vel,
			(const ushort4 *)info, hash, (const float4 *)forces, (float4 *)new_pos, (float4 *)new_vel, bound, dt, st, bodies, (const float4 *)xsph);
	else
		euler_kernel<2><<<nb, BLOCK_STREAM, 0, ctx->stream>>>(ctx->dp, (const float4 *)old_pos, (const float4 *)old_vel,
			(const ushort4 *)info, hash, (const float4 *)forces, (float4 *)new_pos, (float4 *)new_vel, bound, dt, st, bodies, (const float4 *)xsph);
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
