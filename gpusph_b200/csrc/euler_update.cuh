// euler_update.cuh — the predictor/corrector update of ONE particle, shared by the stand-alone integration kernel
// (euler.cu) and the fused epilogue of the forces kernel (forces.cu) so that both paths run the same arithmetic.
// Behavioural specification: GPUSPH eulerDevice, src/cuda/euler_kernel.def:396-540 (corrected velocity :117-134,
// continuity :200-206, moving bodies :470-503, XSPH :165-180).
#pragma once
#include "common.cuh"

// pos / vel: state n of the particle on entry, the integrated state on return. `force` = the current FORCES entry.
// mv: the XSPH mean velocity of the particle (used when use_xsph).
template<int STEP>
__device__ __forceinline__ void
euler_update(const DevParams &P, float4 &pos, float4 &vel, const float4 force, const ushort4 info, const uint *__restrict__ particleHash,
	const uint index, const float dt, const BodyData *__restrict__ bodies, const bool use_xsph, const float4 mv)
{
	const int type = ptype_of(info);
	const bool integrateBoundary = (P.boundarytype == B200SPH_DYN_BOUNDARY || P.boundarytype == B200SPH_SA_BOUNDARY);   // :424-425
	if (!inactive_w(pos.w) && !(type == PT_BOUNDARY && !integrateBoundary && !(info.x & B200SPH_FG_MOVING_BOUNDARY))) {
		// velc = vel (+ force*dt/2 on the corrector), :117-134
		float vcx = vel.x, vcy = vel.y, vcz = vel.z;
		if (STEP == 2) {
			const float hdt = dt / 2;
			vcx += force.x * hdt; vcy += force.y * hdt; vcz += force.z * hdt;
		}
		if (use_xsph) {                                           // XSPH correction, :165-180
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
				const float rx = (float)(gp.x - bodies->eulCgGridPos[obj][0]) * P.cellSize[0] + (pos.x - bodies->eulCgPos[obj][0]);
				const float ry = (float)(gp.y - bodies->eulCgGridPos[obj][1]) * P.cellSize[1] + (pos.y - bodies->eulCgPos[obj][1]);
				const float rz = (float)(gp.z - bodies->eulCgGridPos[obj][2]) * P.cellSize[2] + (pos.z - bodies->eulCgPos[obj][2]);
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
}

// dt of a sub-step: an argument (reference-style call) or the device-resident record — dt/2 on the predictor
// (PredictorCorrector::getDtOperatorForStep), dt on the corrector
template<int STEP>
__device__ __forceinline__ float euler_dt(const StepState *__restrict__ dev_state, const float dt_arg)
{
	return dev_state ? (STEP == 1 ? dev_state->dt / 2 : dev_state->dt) : dt_arg;
}
