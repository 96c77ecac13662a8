// hoststep.cu — one whole predictor-corrector time step of a particle state that lives in HOST memory
// (b200sph_step_host / b200sph_host_upload / b200sph_host_fence / b200sph_host_sync in include/b200sph.h).
//
// No reference counterpart: GPUSPH keeps the state on the device and only moves it for writes
// (GPUWorker::dumpBuffers, src/GPUWorker.cc:1227-1300). This is the entry point a caller with host-resident buffers
// uses (and what bench.py's e2e leg times): the command sequence is the one of src/integrators/
// PredictorCorrectorIntegrator.cc:917-1068, but the copies are pipelined with the force evaluations in stripes of
// whole cell layers on three streams, and consecutive calls are chained stripe by stripe (the upload of stripe s of
// step n+1 only waits for the download of stripe s of step n), so PCIe runs in both directions while the pair kernel
// computes. Results are bitwise those of the resident path: same kernels, same arguments, same order per particle.
//
// Periodicity along COORD3 (the stripes' axis) closes the stripes into a RING: stripe 0 and stripe ns-1 are neighbours.
// A call then starts the ring at stripe r: uploads r-1, r, r+1, ..., r-2; predictors r, ..., r-1; correctors (and
// downloads) r+1, ..., r-2 behind the predictor wave, then r-1 and r, which need the last predictor. The next call
// starts at r+2, so its upload order (r+1, r+2, ...) is this call's download order and the chain stays stripe by stripe.
#include "common.cuh"
#include <stdlib.h>

// B200SPH_HOST_TRACE=1: timing events next to the ordering events; b200sph_host_sync prints the time line of the last
// b200sph_step_host (per stripe: upload landed, predictor forces done, corrector done, download landed) to stderr
enum { TR_BASE = 0, TR_UP = 1, TR_PRED = TR_UP + B200SPH_MAX_STRIPES, TR_CORR = TR_PRED + B200SPH_MAX_STRIPES,
	TR_DOWN = TR_CORR + B200SPH_MAX_STRIPES, TR_UPS = TR_DOWN + B200SPH_MAX_STRIPES, TR_END = TR_UPS + B200SPH_MAX_STRIPES, TR_COUNT };
#define TRACE(i, stream) do { if (ctx->trace_ev) CUDA_TRY(cudaEventRecord(ctx->trace_ev[i], stream)); } while (0)

static int host_streams(b200sph_ctx *ctx)
{
	if (ctx->up_stream) return B200SPH_OK;
	if (const char *e = getenv("B200SPH_HOST_TRACE")) if (atoi(e) > 0) {
		ctx->trace_ev = (cudaEvent_t *)calloc(TR_COUNT, sizeof(cudaEvent_t));
		for (int i = 0; ctx->trace_ev && i < TR_COUNT; ++i) CUDA_TRY(cudaEventCreate(&ctx->trace_ev[i]));
	}
	CUDA_TRY(cudaStreamCreateWithFlags(&ctx->up_stream, cudaStreamNonBlocking));
	CUDA_TRY(cudaStreamCreateWithFlags(&ctx->down_stream, cudaStreamNonBlocking));
	for (int i = 0; i < B200SPH_MAX_STRIPES; ++i) {
		CUDA_TRY(cudaEventCreateWithFlags(&ctx->up_ev[i], cudaEventDisableTiming));
		CUDA_TRY(cudaEventCreateWithFlags(&ctx->down_ev[i], cudaEventDisableTiming));
		CUDA_TRY(cudaEventCreateWithFlags(&ctx->comp_ev[i], cudaEventDisableTiming));
		CUDA_TRY(cudaEventCreateWithFlags(&ctx->pred_ev[i], cudaEventDisableTiming));
	}
	{ const char *e = getenv("B200SPH_HOST_LANES"); ctx->host_lanes = e ? atoi(e) : 2; }
	if (ctx->host_lanes < 1) ctx->host_lanes = 1;
	if (ctx->host_lanes > B200_MAX_LANES) ctx->host_lanes = B200_MAX_LANES;
	for (int l = 1; l < B200_MAX_LANES; ++l) CUDA_TRY(cudaStreamCreateWithFlags(&ctx->lane_stream[l], cudaStreamNonBlocking));
	for (int l = 0; l < 2 * B200_MAX_LANES; ++l) CUDA_TRY(cudaEventCreateWithFlags(&ctx->join_ev[l], cudaEventDisableTiming));
	for (int l = 0; l < 2; ++l) CUDA_TRY(cudaEventCreateWithFlags(&ctx->fork_ev[l], cudaEventDisableTiming));
	CUDA_TRY(cudaEventCreateWithFlags(&ctx->fence_ev, cudaEventDisableTiming));
	CUDA_TRY(cudaEventCreateWithFlags(&ctx->up_all_ev, cudaEventDisableTiming));
	CUDA_TRY(cudaEventCreateWithFlags(&ctx->down_all_ev, cudaEventDisableTiming));
	return B200SPH_OK;
}

void b200_hoststep_destroy(b200sph_ctx *ctx)
{
	if (!ctx->up_stream) return;
	cudaStreamSynchronize(ctx->up_stream); cudaStreamSynchronize(ctx->down_stream);
	for (int i = 0; i < B200SPH_MAX_STRIPES; ++i) { cudaEventDestroy(ctx->up_ev[i]); cudaEventDestroy(ctx->down_ev[i]); cudaEventDestroy(ctx->comp_ev[i]); cudaEventDestroy(ctx->pred_ev[i]); }
	cudaEventDestroy(ctx->fence_ev); cudaEventDestroy(ctx->up_all_ev); cudaEventDestroy(ctx->down_all_ev);
	cudaStreamDestroy(ctx->up_stream); cudaStreamDestroy(ctx->down_stream);
	for (int l = 1; l < B200_MAX_LANES; ++l) if (ctx->lane_stream[l]) { cudaStreamSynchronize(ctx->lane_stream[l]); cudaStreamDestroy(ctx->lane_stream[l]); }
	for (int l = 0; l < 2 * B200_MAX_LANES; ++l) if (ctx->join_ev[l]) cudaEventDestroy(ctx->join_ev[l]);
	for (int l = 0; l < 2; ++l) if (ctx->fork_ev[l]) cudaEventDestroy(ctx->fork_ev[l]);
	ctx->up_stream = ctx->down_stream = NULL;
	if (ctx->trace_ev) { for (int i = 0; i < TR_COUNT; ++i) cudaEventDestroy(ctx->trace_ev[i]); free(ctx->trace_ev); ctx->trace_ev = NULL; }
}

// Uploads of [0, n) from the host state, stripe by stripe on the upload stream. If the previous call downloaded into
// the same host buffers from the same device buffers with a stripe table covering n particles, stripe s only waits
// for that download (which itself followed the last device-side use of the stripe); otherwise the upload stream
// waits for everything enqueued so far on the compute and download streams.
// `first` != 0 (ring order first, first+1, ..., first-1): chained only on an identical stripe table, stripe by stripe.
static int enqueue_uploads(b200sph_ctx *ctx, const void *host_pos, const void *host_vel, void *pos, void *vel,
	const uint32_t *bounds, uint32_t ns, uint32_t first = 0)
{
	bool chained = ctx->host_pending && ctx->host_pos_last == host_pos && ctx->host_vel_last == host_vel &&
		ctx->dev_pos_last == pos && ctx->dev_vel_last == vel && ctx->host_nstripes > 0 &&
		ctx->host_bounds[ctx->host_nstripes] == bounds[ns];
	if (first != 0) {
		chained = chained && ctx->host_nstripes == ns;
		for (uint32_t k = 0; chained && k <= ns; ++k) chained = ctx->host_bounds[k] == bounds[k];
	}
	if (!chained) {
		CUDA_TRY(cudaEventRecord(ctx->fence_ev, ctx->stream));
		CUDA_TRY(cudaStreamWaitEvent(ctx->up_stream, ctx->fence_ev, 0));
		if (ctx->host_pending) CUDA_TRY(cudaStreamWaitEvent(ctx->up_stream, ctx->down_all_ev, 0));
	}
	uint32_t waited = 0;   // stripes [0, waited) of the previous call's table have been waited for
	for (uint32_t i = 0; i < ns; ++i) {
		const uint32_t k = (first + i) % ns;
		const uint32_t a = bounds[k], b = bounds[k + 1];
		// the previous call's downloads into [a, b): the upload stream is ordered, so with ascending stripes every
		// download event is waited for once, before the first upload that touches its range
		if (chained && first != 0)
			CUDA_TRY(cudaStreamWaitEvent(ctx->up_stream, ctx->down_ev[k], 0));
		else if (chained)
			for (; waited < ctx->host_nstripes && ctx->host_bounds[waited] < b; ++waited)
				CUDA_TRY(cudaStreamWaitEvent(ctx->up_stream, ctx->down_ev[waited], 0));
		const size_t off = (size_t)a * 16, bytes = (size_t)(b - a) * 16;
		TRACE(TR_UPS + k, ctx->up_stream);
		CUDA_TRY(cudaMemcpyAsync((char *)pos + off, (const char *)host_pos + off, bytes, cudaMemcpyHostToDevice, ctx->up_stream));
		CUDA_TRY(cudaMemcpyAsync((char *)vel + off, (const char *)host_vel + off, bytes, cudaMemcpyHostToDevice, ctx->up_stream));
		CUDA_TRY(cudaEventRecord(ctx->up_ev[k], ctx->up_stream));
		TRACE(TR_UP + k, ctx->up_stream);
	}
	CUDA_TRY(cudaEventRecord(ctx->up_all_ev, ctx->up_stream));
	return B200SPH_OK;
}

static int check_bounds(const uint32_t *bounds, uint32_t ns, uint32_t n)
{
	if (!bounds || ns < 1 || ns > B200SPH_MAX_STRIPES) { b200_set_error("host step: 1..%d stripes expected", B200SPH_MAX_STRIPES); return B200SPH_EINVAL; }
	if (bounds[0] != 0 || bounds[ns] != n) { b200_set_error("host step: stripe table must cover [0, num_particles)"); return B200SPH_EINVAL; }
	for (uint32_t k = 0; k < ns; ++k) if (bounds[k + 1] <= bounds[k]) { b200_set_error("host step: empty or unordered stripe %u", k); return B200SPH_EINVAL; }
	return B200SPH_OK;
}

extern "C" int b200sph_host_upload(b200sph_ctx *ctx, const void *host_pos, const void *host_vel, void *pos, void *vel,
	uint32_t num_particles)
{
	CHECK_CTX(ctx);
	if (num_particles == 0) return B200SPH_OK;
	if (!host_pos || !host_vel || !pos || !vel) { b200_set_error("host upload: null buffer"); return B200SPH_EINVAL; }
	int rc = host_streams(ctx);
	if (rc) return rc;
	// reuse the previous stripe table when it covers the same particles (keeps the chain), else one stripe
	uint32_t one[2] = { 0, num_particles };
	const bool reuse = ctx->host_nstripes > 0 && ctx->host_bounds[ctx->host_nstripes] == num_particles;
	rc = enqueue_uploads(ctx, host_pos, host_vel, pos, vel, reuse ? ctx->host_bounds : one, reuse ? ctx->host_nstripes : 1);
	if (rc) return rc;
	// whatever follows on the compute stream (the neighbour rebuild) sees the whole state
	CUDA_TRY(cudaStreamWaitEvent(ctx->stream, ctx->up_all_ev, 0));
	return B200SPH_OK;
}

extern "C" int b200sph_step_host(b200sph_ctx *ctx, const b200sph_host_step_args *a)
{
	CHECK_CTX(ctx);
	if (!a) { b200_set_error("host step: null argument block"); return B200SPH_EINVAL; }
	const uint32_t n = a->num_particles, ns = a->num_stripes;
	if (n == 0) return B200SPH_OK;
	if (!a->host_pos || !a->host_vel || !a->pos || !a->vel || !a->pos_star || !a->vel_star || !a->info || !a->hash ||
		!a->cell_start || !a->neibs_list || !a->forces || !a->cfl) { b200_set_error("host step: null buffer"); return B200SPH_EINVAL; }
	int rc = check_bounds(a->stripe_bounds, ns, n);
	if (rc) return rc;
	const bool xsph = (ctx->hp.simflags & B200SPH_ENABLE_XSPH) != 0;
	if (xsph && !a->xsph) { b200_set_error("host step: ENABLE_XSPH needs the xsph buffer"); return B200SPH_EINVAL; }
	uint32_t need = 0;
	for (uint32_t k = 0; k < ns; ++k) need += (div_up(a->stripe_bounds[k + 1] - a->stripe_bounds[k], BLOCK_FORCES) + 3) / 4 * 4;
	if (2 * need > a->cfl_elements) { b200_set_error("host step: cfl array too small (%u < %u)", a->cfl_elements, 2 * need); return B200SPH_EINVAL; }
	rc = host_streams(ctx);
	if (rc) return rc;
	const uint32_t *B = a->stripe_bounds;
	// periodic along the stripes' axis: a ring of stripes (2 stripes are each other's only neighbours either way)
	const bool ring = (ctx->hp.periodic & (1u << ctx->hp.coord[2])) != 0 && ns >= 3;
	const uint32_t r0 = ring ? ctx->host_ring_next % ns : 0;                 // where this call's ring starts
	auto ring_pos = [&](uint32_t k) { return (k + ns - r0) % ns; };           // place of stripe k in the predictor order
	// Two compute lanes: P (the context's stream) runs the predictor of stripe k — forces(n), euler step 1 (dt/2) -> n* —
	// and Q the corrector of stripe k-1 — forces(n*), euler step 2 IN PLACE into the state-n buffers, download — as
	// soon as n* exists for stripes k-2..k. The corrector of a stripe therefore starts long before the predictor of the
	// last stripes has run: downloads (and the next call's chained uploads) are spread over the whole step instead of
	// being packed into its second half, and the two lanes fill each other's grid tails (back to back on one stream,
	// eight 2-wave grids cost 1.5x the single launch; measured). B200SPH_HOST_LANES=1: same order
	// on one stream.
	cudaStream_t P = ctx->stream, D = ctx->down_stream;
	cudaStream_t Q = ctx->host_lanes > 1 ? ctx->lane_stream[1] : P;
	struct Restore { b200sph_ctx *c; cudaStream_t s; ~Restore() { c->stream = s; } } restore = { ctx, P };

	TRACE(TR_BASE, P);
	ctx->trace_resident = a->resident;
	if (Q != P) {                                     // Q starts after whatever precedes this call on P
		CUDA_TRY(cudaEventRecord(ctx->fork_ev[0], P));
		CUDA_TRY(cudaStreamWaitEvent(Q, ctx->fork_ev[0], 0));
	}
	if (!a->resident) {
		rc = enqueue_uploads(ctx, a->host_pos, a->host_vel, a->pos, a->vel, B, ns, ring ? (r0 + ns - 1) % ns : 0);
		if (rc) return rc;
	}

	// neighbour records of the pair kernel (forces.cu): state n in the context's record buffer 0, interleaved stripe by
	// stripe as the uploads land; n* in buffer 1, written by the predictor's epilogue
	PosVel *pv_n = NULL, *pv_star = NULL;
	rc = b200_packed_scratch(ctx, 0, n, &pv_n); if (rc) return rc;
	rc = b200_packed_scratch(ctx, 1, n, &pv_star); if (rc) return rc;
	uint32_t packed_upto = 0;                         // stripes [0, packed_upto) have their records
	auto pack_upto = [&](uint32_t stripes) -> int {   // on P, after the caller made P wait for the uploads involved
		if (stripes > ns) stripes = ns;
		if (stripes <= packed_upto) return B200SPH_OK;
		ctx->stream = P;
		const int r = b200sph_pack_state(ctx, a->pos, a->vel, pv_n, B[packed_upto], B[stripes]);
		packed_upto = stripes;
		return r;
	};
	if (a->resident) { rc = pack_upto(ns); if (rc) return rc; }
	bool have[B200SPH_MAX_STRIPES] = {};              // ring: stripes whose upload P has waited for and whose records exist
	auto need_stripe = [&](uint32_t k) -> int {
		if (a->resident || have[k]) return B200SPH_OK;
		have[k] = true;
		CUDA_TRY(cudaStreamWaitEvent(P, ctx->up_ev[k], 0));
		ctx->stream = P;
		return b200sph_pack_state(ctx, a->pos, a->vel, pv_n, B[k], B[k + 1]);
	};

	b200sph_forces_args f;
	memset(&f, 0, sizeof(f));
	f.info = a->info; f.hash = a->hash; f.cell_start = a->cell_start; f.neibs_list = a->neibs_list;
	f.forces = a->forces; f.cfl = a->cfl; f.xsph = a->xsph;
	f.num_particles = n; f.dt_from_device = 1;
	uint32_t offP = 0, offQ = 0, nb = 0;
	const uint32_t cflQ = need;                       // the corrector's CFL blocks live behind the predictor's

	// state n -> n* for stripe k (the neighbours of stripe k lie in stripes k-1..k+1: upload k+1 must have landed)
	auto predictor = [&](uint32_t k) -> int {
		const uint32_t s = B[k], e = B[k + 1];
		const size_t o16 = (size_t)s * 16;
		ctx->stream = P;
		if (ring) {
			for (uint32_t d = 0; d < 3; ++d) { const int rn = need_stripe((k + ns - 1 + d) % ns); if (rn) return rn; }
		} else {
			if (!a->resident) CUDA_TRY(cudaStreamWaitEvent(P, ctx->up_ev[k + 1 < ns ? k + 1 : ns - 1], 0));
			const int rp = pack_upto(k + 2); if (rp) return rp;
		}
		if (xsph) CUDA_TRY(cudaMemsetAsync((char *)a->xsph + o16, 0, (size_t)(e - s) * 16, P));
		f.pos = a->pos; f.vel = a->vel; f.step = 1; f.packed = pv_n;
		f.from_particle = s; f.to_particle = e; f.cfl_offset = offP;
		// forces + euler step 1 (dt/2) of the stripe in one launch (fused epilogue)
		b200sph_fused_euler_args eu;
		eu.old_pos = a->pos; eu.old_vel = a->vel; eu.new_pos = a->pos_star; eu.new_vel = a->vel_star;
		eu.dt = 0.0f; eu.step = 1; eu.dt_from_device = 1; eu.new_packed = pv_star;
		int r = b200sph_forces_euler(ctx, &f, &eu, &nb);
		if (r) return r;
		offP += nb;
		if (Q != P) CUDA_TRY(cudaEventRecord(ctx->pred_ev[k], P));
		TRACE(TR_PRED + k, P);
		return B200SPH_OK;
	};
	// state n* -> n+1 for stripe j: needs n* of stripes j-1..j+1, i.e. the predictor of stripe j+1. The forces of the
	// other stripes read n*, not the state-n buffers this integrates in place; the last readers of state n of stripe j
	// (the predictor's forces of stripes j-1..j+1) precede it.
	auto corrector = [&](uint32_t j) -> int {
		const uint32_t s = B[j], e = B[j + 1];
		const size_t o16 = (size_t)s * 16;
		ctx->stream = Q;
		uint32_t last = j + 1 < ns ? j + 1 : ns - 1;     // the predictor that completes n* of stripes j-1..j+1
		if (ring) {
			last = j;
			for (uint32_t d = 0; d < 3; d += 2) { const uint32_t k = (j + ns - 1 + d) % ns; if (ring_pos(k) > ring_pos(last)) last = k; }
		}
		if (Q != P) CUDA_TRY(cudaStreamWaitEvent(Q, ctx->pred_ev[last], 0));
		if (xsph) CUDA_TRY(cudaMemsetAsync((char *)a->xsph + o16, 0, (size_t)(e - s) * 16, Q));
		f.pos = a->pos_star; f.vel = a->vel_star; f.step = 2; f.packed = pv_star;
		f.from_particle = s; f.to_particle = e; f.cfl_offset = cflQ + offQ;
		// forces + euler step 2 of the stripe, in place into the state-n buffers, in one launch
		b200sph_fused_euler_args eu;
		eu.old_pos = a->pos; eu.old_vel = a->vel; eu.new_pos = a->pos; eu.new_vel = a->vel;
		eu.dt = 0.0f; eu.step = 2; eu.dt_from_device = 1; eu.new_packed = NULL;
		int r = b200sph_forces_euler(ctx, &f, &eu, &nb);
		if (r) return r;
		offQ += nb;
		CUDA_TRY(cudaEventRecord(ctx->comp_ev[j], Q));
		TRACE(TR_CORR + j, Q);
		CUDA_TRY(cudaStreamWaitEvent(D, ctx->comp_ev[j], 0));
		CUDA_TRY(cudaMemcpyAsync((char *)a->host_pos + o16, (char *)a->pos + o16, (size_t)(e - s) * 16, cudaMemcpyDeviceToHost, D));
		CUDA_TRY(cudaMemcpyAsync((char *)a->host_vel + o16, (char *)a->vel + o16, (size_t)(e - s) * 16, cudaMemcpyDeviceToHost, D));
		CUDA_TRY(cudaEventRecord(ctx->down_ev[j], D));
		TRACE(TR_DOWN + j, D);
		return B200SPH_OK;
	};
	if (ring) {
		for (uint32_t i = 0; i < ns; ++i) {
			rc = predictor((r0 + i) % ns);
			if (rc) return rc;
			if (i >= 2) { rc = corrector((r0 + i - 1) % ns); if (rc) return rc; }
		}
		rc = corrector((r0 + ns - 1) % ns);           // the two stripes next to where the ring was cut
		if (rc) return rc;
		rc = corrector(r0);
		if (rc) return rc;
		ctx->host_ring_next = (r0 + 2) % ns;
	} else {
		for (uint32_t k = 0; k < ns; ++k) {
			rc = predictor(k);
			if (rc) return rc;
			if (k >= 1) { rc = corrector(k - 1); if (rc) return rc; }
		}
		rc = corrector(ns - 1);
		if (rc) return rc;
	}
	CUDA_TRY(cudaEventRecord(ctx->down_all_ev, D));
	// dt candidates of the two force evaluations, then t += dt, dt = min(candidates) once both are known
	ctx->stream = P;
	rc = b200sph_dtreduce_async(ctx, a->cfl, offP, 1);
	if (rc) return rc;
	if (Q != P) {
		CUDA_TRY(cudaEventRecord(ctx->fork_ev[1], P));
		CUDA_TRY(cudaStreamWaitEvent(Q, ctx->fork_ev[1], 0));
	}
	ctx->stream = Q;
	rc = b200sph_dtreduce_async(ctx, a->cfl + cflQ, offQ, 2);
	if (rc) return rc;
	rc = b200sph_step_end(ctx);
	if (rc) return rc;
	if (Q != P) {                                     // everything of this step is ordered before what follows on P
		CUDA_TRY(cudaEventRecord(ctx->join_ev[0], Q));
		CUDA_TRY(cudaStreamWaitEvent(P, ctx->join_ev[0], 0));
	}
	ctx->stream = P;
	TRACE(TR_END, P);
	// what the next call chains on
	for (uint32_t k = 0; k <= ns; ++k) ctx->host_bounds[k] = B[k];
	ctx->host_nstripes = ns;
	ctx->host_pos_last = a->host_pos; ctx->host_vel_last = a->host_vel;
	ctx->dev_pos_last = a->pos; ctx->dev_vel_last = a->vel;
	ctx->host_pending = 1;
	return B200SPH_OK;
}

// the compute stream waits for the copies still in flight (before anything else touches the state buffers)
extern "C" int b200sph_host_fence(b200sph_ctx *ctx)
{
	CHECK_CTX(ctx);
	if (!ctx->up_stream || !ctx->host_pending) return B200SPH_OK;
	CUDA_TRY(cudaStreamWaitEvent(ctx->stream, ctx->up_all_ev, 0));
	CUDA_TRY(cudaStreamWaitEvent(ctx->stream, ctx->down_all_ev, 0));
	ctx->host_pending = 0;
	return B200SPH_OK;
}

// the host waits until the state of the last b200sph_step_host has landed in the host buffers
extern "C" int b200sph_host_sync(b200sph_ctx *ctx)
{
	CHECK_CTX(ctx);
	if (!ctx->up_stream) return B200SPH_OK;
	CUDA_TRY(cudaStreamSynchronize(ctx->down_stream));
	CUDA_TRY(cudaStreamSynchronize(ctx->stream));
	if (ctx->trace_ev && ctx->host_nstripes) {
		CUDA_TRY(cudaStreamSynchronize(ctx->up_stream));
		auto at = [&](int i) { float ms = 0.f; cudaEventElapsedTime(&ms, ctx->trace_ev[TR_BASE], ctx->trace_ev[i]); return ms; };
		fprintf(stderr, "[b200sph host trace] ms after the compute stream entered the step (%u stripes%s); step end %.3f\n",
			ctx->host_nstripes, ctx->trace_resident ? ", resident" : "", at(TR_END));
		for (uint32_t k = 0; k < ctx->host_nstripes; ++k)
			fprintf(stderr, "  stripe %2u [%8u, %8u): up %7.3f .. %7.3f  pred %7.3f  corr %7.3f  down %7.3f\n", k, ctx->host_bounds[k],
				ctx->host_bounds[k + 1], ctx->trace_resident ? 0.f : at(TR_UPS + k), ctx->trace_resident ? 0.f : at(TR_UP + k),
				at(TR_PRED + k), at(TR_CORR + k), at(TR_DOWN + k));
	}
	return B200SPH_OK;
}
