// filters.cu — density filters (Shepard, MLS) and the TESTPOINTS post-process: the consumers of the neighbour
// list next to the forces kernel.
// Behavioural specification: GPUSPH shepardDevice / MlsDevice (src/cuda/forces_kernel.cu:418-507, 509-721),
// calcTestpointsVelocityDevice (src/cuda/post_process_kernel.cu:134-240), MlsMatrixContrib / MlsCorrContrib
// (src/cuda/forces_kernel.cu:235-261), symtensor4 algebra (src/cuda/tensor.cu:64-100, 246-285), hypot
// (src/vector_math.h:1231-1240). One thread per particle walking its list column in list order, like the reference,
// so the float sums are accumulated in the reference's order.
#include "common.cuh"
#include <float.h>

#define BLOCK_FILTER 128

__device__ __forceinline__ uint ldl(const ushort *p) { uint v; asm volatile("ld.global.nc.u16 %0, [%1];" : "=r"(v) : "l"(p)); return v; }

// for_each_neib2(PT_FLUID, boundary ? PT_BOUNDARY : PT_NONE, ...) — neibs_iteration.cuh:56-396 with getNeibIndex
// (cellgrid.cuh:198-226): the fluid section upwards from row 0, then the boundary section downwards from
// neibboundpos. body(neib_index, relPos) with relPos.w = neighbour mass.
template<typename Body>
__device__ __forceinline__ void
for_each_neib(const DevParams &P, const uint index, const float4 pos, const uint cellHash, const uint *__restrict__ cellStart,
	const ushort *__restrict__ neibsList, const float4 *__restrict__ posArray, const bool with_boundary, Body body)
{
	const int3 gp = grid_pos(P, cellHash);
	uint row_step;
	const ushort *const column = neibsList + list_column(index, P.stride, P.neiblistsize, P.listblock, row_step);
	for (int section = 0; section < (with_boundary ? 2 : 1); ++section) {
		long long row = section == 0 ? 0 : (long long)P.neibboundpos;
		const long long step = section == 0 ? 1 : -1;
		uint base = 0;
		float pcx = 0.f, pcy = 0.f, pcz = 0.f;
		for (; row >= 0 && row < (long long)P.neiblistsize; row += step) {
			uint nd = ldl(column + (size_t)row * row_step);
			if (nd == NEIBS_END) break;
			if (nd >= CELLNUM_ENCODED) {
				const int cell = (int)(nd >> CELLNUM_SHIFT) - 1;
				nd &= NEIBINDEX_MASK;
				const int ox = cell % 3 - 1, oy = (cell / 3) % 3 - 1, oz = cell / 9 - 1;
				pcx = pos.x - (float)ox * P.cellSize[0]; pcy = pos.y - (float)oy * P.cellSize[1]; pcz = pos.z - (float)oz * P.cellSize[2];
				// calcGridHashPeriodic, cellgrid.cuh:174-185
				int gx = gp.x + ox, gy = gp.y + oy, gz = gp.z + oz;
				if (gx < 0) gx = P.gridSize[0] - 1; else if (gx >= P.gridSize[0]) gx = 0;
				if (gy < 0) gy = P.gridSize[1] - 1; else if (gy >= P.gridSize[1]) gy = 0;
				if (gz < 0) gz = P.gridSize[2] - 1; else if (gz >= P.gridSize[2]) gz = 0;
				base = __ldg(cellStart + grid_hash(P, gx, gy, gz));
			}
			const uint j = base + nd;
			const float4 np = __ldg(posArray + j);
			body(j, make_float4(pcx - np.x, pcy - np.y, pcz - np.z, np.w));
		}
	}
}

__device__ __forceinline__ float wendland_W(const DevParams &P, const float r)
{	// W<WENDLAND>, src/cuda/sph_core.cu:104-117
	const float R = r / P.slength;
	float val = 1.0f - 0.5f * R;
	val *= val;
	val *= val;
	val *= 1.0f + 2.0f * R;
	val *= P.wcoeff_wendland;
	return val;
}
__device__ __forceinline__ float phys_rho(const DevParams &P, const float rho_tilde, const int f) { return (rho_tilde + 1.0f) * P.rho0[f]; }
__device__ __forceinline__ float num_rho(const DevParams &P, const float rho, const int f) { return rho / P.rho0[f] - 1.0f; }   // phys_core.cu:145-151

// ---------------------------------------------------------------------------
// Shepard filter, src/cuda/forces_kernel.cu:418-507
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(BLOCK_FILTER)
shepard_kernel(const __grid_constant__ DevParams P, const float4 *__restrict__ posArray, const float4 *__restrict__ oldVel,
	float4 *__restrict__ newVel, const ushort4 *__restrict__ infoArray, const uint *__restrict__ particleHash,
	const uint *__restrict__ cellStart, const ushort *__restrict__ neibsList, const uint numParticles)
{
	const uint index = blockIdx.x * blockDim.x + threadIdx.x;
	if (index >= numParticles) return;
	const ushort4 info = infoArray[index];
	const float4 pos = posArray[index];
	if (inactive_w(pos.w)) return;
	float4 vel = oldVel[index];
	if (ptype_of(info) != PT_FLUID) { newVel[index] = vel; return; }
	const int fnum = fluid_num_of(info);
	// self contribution
	float temp1 = pos.w * wendland_W(P, 0.0f);
	float temp2 = temp1 / phys_rho(P, vel.w, fnum);
	for_each_neib(P, index, pos, particleHash[index] & CELLTYPE_BITMASK, cellStart, neibsList, posArray,
		P.boundarytype == B200SPH_DYN_BOUNDARY, [&](const uint j, const float4 relPos) {
			if (inactive_w(relPos.w)) return;
			const float r = sqrtf(relPos.x * relPos.x + relPos.y * relPos.y + relPos.z * relPos.z);
			if (r < P.influenceradius) {
				const float neib_rho = phys_rho(P, __ldg(oldVel + j).w, fluid_num_of(__ldg(infoArray + j)));
				const float w = wendland_W(P, r) * relPos.w;
				temp1 += w;
				temp2 += w / neib_rho;
			}
		});
	vel.w = num_rho(P, temp1 / temp2, fnum);
	newVel[index] = vel;
}

// ---------------------------------------------------------------------------
// MLS filter, src/cuda/forces_kernel.cu:509-721
// ---------------------------------------------------------------------------
struct SymTensor4 { float xx, xy, xz, xw, yy, yz, yw, zz, zw, ww; };

__device__ __forceinline__ float det4(const SymTensor4 &T)
{	// src/cuda/tensor.cu:64-100
	float ret = 0;
	float M = 0;
	M += T.xx * (T.yy * T.zz - T.yz * T.yz);
	M -= T.xy * (T.xy * T.zz - T.xz * T.yz);
	M += T.xz * (T.xy * T.yz - T.xz * T.yy);
	ret += M * T.ww;
	M = 0;
	M += T.xx * (T.yy * T.zw - T.yz * T.yw);
	M -= T.xy * (T.xy * T.zw - T.xz * T.yw);
	M += T.xw * (T.xy * T.yz - T.xz * T.yy);
	ret -= M * T.zw;
	M = 0;
	M += T.xx * (T.yz * T.zw - T.zz * T.yw);
	M -= T.xz * (T.xy * T.zw - T.xz * T.yw);
	M += T.xw * (T.xy * T.zz - T.xz * T.yz);
	ret += M * T.yw;
	M = 0;
	M += T.xy * (T.yz * T.zw - T.zz * T.yw);
	M -= T.xz * (T.yy * T.zw - T.yz * T.yw);
	M += T.xw * (T.yy * T.zz - T.yz * T.yz);
	ret -= M * T.xw;
	return ret;
}
__device__ __forceinline__ float4 tdot(const SymTensor4 &T, const float4 v)
{	// src/cuda/tensor.cu:240-249
	return make_float4(
		T.xx * v.x + T.xy * v.y + T.xz * v.z + T.xw * v.w,
		T.xy * v.x + T.yy * v.y + T.yz * v.z + T.yw * v.w,
		T.xz * v.x + T.yz * v.y + T.zz * v.z + T.zw * v.w,
		T.xw * v.x + T.yw * v.y + T.zw * v.z + T.ww * v.w);
}
__device__ __forceinline__ float tddot(const SymTensor4 &T, const float4 v)
{	// src/cuda/tensor.cu:261-271
	return T.xx * v.x * v.x + T.yy * v.y * v.y + T.zz * v.z * v.z + T.ww * v.w * v.w +
		2 * ((T.xy * v.y + T.xw * v.w) * v.x + (T.yz * v.z + T.yw * v.w) * v.y + (T.xz * v.x + T.zw * v.w) * v.z);
}
__device__ __forceinline__ float4 adjugate_row1(const SymTensor4 &T)
{	// src/cuda/tensor.cu:273-283
	return make_float4(
		T.yy * T.zz * T.ww + T.yz * T.zw * T.yw + T.yw * T.yz * T.zw - T.yy * T.zw * T.zw - T.yz * T.yz * T.ww - T.yw * T.zz * T.yw,
		T.xy * T.zw * T.zw + T.yz * T.xz * T.ww + T.yw * T.zz * T.xw - T.xy * T.zz * T.ww - T.yz * T.zw * T.xw - T.yw * T.xz * T.zw,
		T.xy * T.yz * T.ww + T.yy * T.zw * T.xw + T.yw * T.xz * T.yw - T.xy * T.zw * T.yw - T.yy * T.xz * T.ww - T.yw * T.yz * T.xw,
		T.xy * T.zz * T.yw + T.yy * T.xz * T.zw + T.yz * T.yz * T.xw - T.xy * T.yz * T.zw - T.yy * T.zz * T.xw - T.yz * T.xz * T.yw);
}
__device__ __forceinline__ float hypot4(const float4 v)
{	// src/vector_math.h:1231-1240
	const float p = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)));
	if (!p) return 0;
	const float inv = 1.0f / p;
	const float wx = v.x * inv, wy = v.y * inv, wz = v.z * inv, ww = v.w * inv;
	return p * sqrtf(wx * wx + wy * wy + wz * wz + ww * ww);
}

__global__ void __launch_bounds__(BLOCK_FILTER)
mls_kernel(const __grid_constant__ DevParams P, const float4 *__restrict__ posArray, const float4 *__restrict__ oldVel,
	float4 *__restrict__ newVel, const ushort4 *__restrict__ infoArray, const uint *__restrict__ particleHash,
	const uint *__restrict__ cellStart, const ushort *__restrict__ neibsList, const uint numParticles)
{
	const uint index = blockIdx.x * blockDim.x + threadIdx.x;
	if (index >= numParticles) return;
	const ushort4 info = infoArray[index];
	const float4 pos = posArray[index];
	if (inactive_w(pos.w)) return;
	float4 vel = oldVel[index];
	const int fnum = fluid_num_of(info);
	const bool dyn = P.boundarytype == B200SPH_DYN_BOUNDARY;
	const uint cellHash = particleHash[index] & CELLTYPE_BITMASK;
	const float h = P.slength;

	SymTensor4 mls;
	mls.xx = mls.xy = mls.xz = mls.xw = mls.yy = mls.yz = mls.yw = mls.zz = mls.zw = mls.ww = 0;
	// self contribution
	mls.xx = wendland_W(P, 0.0f) * pos.w / phys_rho(P, vel.w, fnum);

	// first loop: the MLS matrix (MlsMatrixContrib :235-249 on relPos/h)
	for_each_neib(P, index, pos, cellHash, cellStart, neibsList, posArray, dyn, [&](const uint j, const float4 relPos) {
		if (inactive_w(relPos.w)) return;
		const float r = sqrtf(relPos.x * relPos.x + relPos.y * relPos.y + relPos.z * relPos.z);
		const float neib_rho = phys_rho(P, __ldg(oldVel + j).w, fluid_num_of(__ldg(infoArray + j)));
		if (r < P.influenceradius) {
			const float w = wendland_W(P, r) * relPos.w / neib_rho;      // Wij*Vj
			const float inv_h = 1.0f / h;                                  // float4 / float multiplies by the reciprocal (vector_math.h:1093-1097)
			const float x = relPos.x * inv_h, y = relPos.y * inv_h, z = relPos.z * inv_h;
			mls.xx += w;
			mls.xy += x * w; mls.xz += y * w; mls.xw += z * w;
			mls.yy += x * x * w; mls.yz += x * y * w; mls.yw += x * z * w;
			mls.zz += y * y * w; mls.zw += y * z * w;
			mls.ww += z * z * w;
		}
	});

	// B = first row of the inverse, refined by conjugate-residual iterations (:602-660)
	const float4 E = make_float4(1, 0, 0, 0);
	const float D = det4(mls);
	float4 B;
	if (fabsf(D) < FLT_EPSILON) {
		SymTensor4 m2 = mls;
		const float eps = fabsf(D) + FLT_EPSILON;
		m2.xx += eps; m2.yy += eps; m2.zz += eps; m2.ww += eps;
		const float inv = 1.0f / det4(m2);
		const float4 a = adjugate_row1(m2);
		B = make_float4(a.x * inv, a.y * inv, a.z * inv, a.w * inv);
	} else {
		const float inv = 1.0f / D;
		const float4 a = adjugate_row1(mls);
		B = make_float4(a.x * inv, a.y * inv, a.z * inv, a.w * inv);
	}
	for (uint steps = 0; steps < 32; ++steps) {
		const float lenB = hypot4(B);
		const float4 MdotB = tdot(mls, B);
		const float4 residual = make_float4(E.x - MdotB.x, E.y - MdotB.y, E.z - MdotB.z, E.w - MdotB.w);
		const float num = tddot(mls, residual);
		const float4 Mp = tdot(mls, residual);
		const float den = Mp.x * Mp.x + Mp.y * Mp.y + Mp.z * Mp.z + Mp.w * Mp.w;
		const float s = num / den;
		const float4 corr = make_float4(s * residual.x, s * residual.y, s * residual.z, s * residual.w);
		const float lencorr = hypot4(corr);
		if (hypot4(residual) < lenB * FLT_EPSILON) break;
		if (lencorr < 2 * lenB * FLT_EPSILON) break;
		B.x += corr.x; B.y += corr.y; B.z += corr.z; B.w += corr.w;
	}
	B.y /= h; B.z /= h; B.w /= h;

	// self contribution, then the second loop (MlsCorrContrib :256-260)
	vel.w = B.x * wendland_W(P, 0.0f) * pos.w;
	for_each_neib(P, index, pos, cellHash, cellStart, neibsList, posArray, dyn, [&](const uint j, const float4 relPos) {
		if (inactive_w(relPos.w)) return;
		const float r = sqrtf(relPos.x * relPos.x + relPos.y * relPos.y + relPos.z * relPos.z);
		if (r < P.influenceradius && (dyn || ptype_of(__ldg(infoArray + j)) == PT_FLUID)) {
			const float w = wendland_W(P, r) * relPos.w;                  // mj*Wij
			vel.w += (B.x + B.y * relPos.x + B.z * relPos.y + B.w * relPos.z) * w;
		}
	});
	vel.w = num_rho(P, vel.w, fnum);
	newVel[index] = vel;
}

// ---------------------------------------------------------------------------
// TESTPOINTS, src/cuda/post_process_kernel.cu:134-240 (in-place update of vel / tke / epsilon)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(BLOCK_FILTER)
testpoints_kernel(const __grid_constant__ DevParams P, const float4 *__restrict__ posArray, float4 *vel, float *tke, float *epsilon,
	const ushort4 *__restrict__ infoArray, const uint *__restrict__ particleHash,
	const uint *__restrict__ cellStart, const ushort *__restrict__ neibsList, const uint numParticles)
{
	const uint index = blockIdx.x * blockDim.x + threadIdx.x;
	if (index >= numParticles) return;
	const ushort4 info = infoArray[index];
	if (ptype_of(info) != PT_TESTPOINT) return;
	const float4 pos = posArray[index];
	float4 velavg = make_float4(0.f, 0.f, 0.f, 0.f);
	float tkeavg = 0.f, epsavg = 0.f, alpha = 0.f;
	// only fluid neighbours are read and only test points are written: the in-place update is race free
	for_each_neib(P, index, pos, particleHash[index] & CELLTYPE_BITMASK, cellStart, neibsList, posArray, false,
		[&](const uint j, const float4 relPos) {
			const float r = sqrtf(relPos.x * relPos.x + relPos.y * relPos.y + relPos.z * relPos.z);
			if (r < P.influenceradius) {
				const float4 nv = vel[j];
				const int nf = fluid_num_of(__ldg(infoArray + j));
				const float w = wendland_W(P, r) * relPos.w / phys_rho(P, nv.w, nf);
				velavg.x += w * nv.x; velavg.y += w * nv.y; velavg.z += w * nv.z;
				// P(), src/cuda/phys_core.cu:99-110
				velavg.w += w * (P.bcoeff[nf] * (__powf(nv.w + 1.0f, P.gammacoeff[nf]) - 1.0f));
				if (tke) tkeavg += w * tke[j];
				if (epsilon) epsavg += w * epsilon[j];
				alpha += w;
			}
		});
	if (alpha > 1e-5f) {
		const float inv = 1.0f / alpha;                                // float4 /= float, vector_math.h:1105-1109
		velavg.x *= inv; velavg.y *= inv; velavg.z *= inv; velavg.w *= inv;
		tkeavg /= alpha; epsavg /= alpha;
	} else {
		velavg = make_float4(0.f, 0.f, 0.f, 0.f);
		tkeavg = epsavg = 0.f;
	}
	vel[index] = velavg;
	if (tke) tke[index] = tkeavg;
	if (epsilon) epsilon[index] = epsavg;
}

// ---------------------------------------------------------------------------
// entry points
// ---------------------------------------------------------------------------
static int check_list_args(const char *what, const void *pos, const void *vel, const void *out, const void *info, const uint32_t *hash,
	const uint32_t *cell_start, const uint16_t *neibs_list, uint32_t num_particles, uint32_t range_end, uint32_t stride)
{
	if (!pos || !vel || !out || !info || !hash || !cell_start || !neibs_list) { b200_set_error("%s: null buffer", what); return B200SPH_EINVAL; }
	if (range_end > num_particles) { b200_set_error("%s: range end beyond numParticles", what); return B200SPH_EINVAL; }
	if (range_end > stride) { b200_set_error("%s: range end %u exceeds neighbour list stride %u", what, range_end, stride); return B200SPH_EINVAL; }
	return B200SPH_OK;
}

extern "C" int b200sph_filter_shepard(b200sph_ctx *ctx, const void *pos, const void *old_vel, void *new_vel, const void *info,
	const uint32_t *hash, const uint32_t *cell_start, const uint16_t *neibs_list, uint32_t num_particles, uint32_t range_end)
{
	CHECK_CTX(ctx);
	if (range_end == 0) return B200SPH_OK;
	int rc = check_list_args("shepard filter", pos, old_vel, new_vel, info, hash, cell_start, neibs_list, num_particles, range_end, ctx->dp.stride);
	if (rc) return rc;
	if (old_vel == new_vel) { b200_set_error("shepard filter: old and new velocity buffers must differ"); return B200SPH_EINVAL; }
	shepard_kernel<<<div_up(range_end, BLOCK_FILTER), BLOCK_FILTER, 0, ctx->stream>>>(ctx->dp, (const float4 *)pos, (const float4 *)old_vel,
		(float4 *)new_vel, (const ushort4 *)info, hash, cell_start, neibs_list, range_end);
	KERNEL_TRY();
	return B200SPH_OK;
}

extern "C" int b200sph_filter_mls(b200sph_ctx *ctx, const void *pos, const void *old_vel, void *new_vel, const void *info,
	const uint32_t *hash, const uint32_t *cell_start, const uint16_t *neibs_list, uint32_t num_particles, uint32_t range_end)
{
	CHECK_CTX(ctx);
	if (range_end == 0) return B200SPH_OK;
	int rc = check_list_args("MLS filter", pos, old_vel, new_vel, info, hash, cell_start, neibs_list, num_particles, range_end, ctx->dp.stride);
	if (rc) return rc;
	if (old_vel == new_vel) { b200_set_error("MLS filter: old and new velocity buffers must differ"); return B200SPH_EINVAL; }
	mls_kernel<<<div_up(range_end, BLOCK_FILTER), BLOCK_FILTER, 0, ctx->stream>>>(ctx->dp, (const float4 *)pos, (const float4 *)old_vel,
		(float4 *)new_vel, (const ushort4 *)info, hash, cell_start, neibs_list, range_end);
	KERNEL_TRY();
	return B200SPH_OK;
}

extern "C" int b200sph_testpoints(b200sph_ctx *ctx, const void *pos, void *vel, float *tke, float *epsilon, const void *info,
	const uint32_t *hash, const uint32_t *cell_start, const uint16_t *neibs_list, uint32_t num_particles, uint32_t range_end)
{
	CHECK_CTX(ctx);
	if (range_end == 0) return B200SPH_OK;
	int rc = check_list_args("testpoints", pos, vel, vel, info, hash, cell_start, neibs_list, num_particles, range_end, ctx->dp.stride);
	if (rc) return rc;
	testpoints_kernel<<<div_up(range_end, BLOCK_FILTER), BLOCK_FILTER, 0, ctx->stream>>>(ctx->dp, (const float4 *)pos, (float4 *)vel,
		tke, epsilon, (const ushort4 *)info, hash, cell_start, neibs_list, range_end);
	KERNEL_TRY();
	return B200SPH_OK;
}
