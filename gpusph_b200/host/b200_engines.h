/*
 * b200_engines.h — reference-side binding of the B200 engines.
 *
 * This header is compiled INSIDE the GPUSPH tree (it includes GPUSPH's own headers): thin subclasses of
 *   AbstractNeibsEngine        (src/engine_neibs.h:45-107)
 *   AbstractForcesEngine       (src/engine_forces.h:42-179)
 *   AbstractIntegrationEngine  (src/engine_integration.h:40-143)
 * that pull raw device pointers out of the BufferLists exactly like the reference's CUDA engines do
 * (src/cuda/buildneibs.cu:166-171, src/cuda/forces.cu:901-932, src/cuda/euler.cu:330-372) and forward them
 * to the C ABI of include/b200sph.h. Error codes are turned back into the exceptions the reference throws.
 * FilterEngine (SHEPARD_FILTER / MLS_FILTER, src/engine_filter.h:40-83) and TestpointsEngine (TESTPOINTS,
 * src/engine_postprocess.h:47-116) bind the two neighbour-list consumers DamBreak3D / Poiseuille enable next to the forces.
 * Methods that belong to out-of-scope subsystems (SA boundaries, DEM, density summation, repacking: SURVEY.md
 * section 8 rows "out of scope") throw std::runtime_error — they never fall back.
 *
 * Engines are shared by all worker threads (one SimFramework per process, SURVEY.md section 8b "Threading"), so the
 * per-device context lives in a map keyed by the CUDA device of the calling thread.
 *
 * How a maintainer wires it in: the framework seam gpusph_b200/host/cudasimframework.cu builds these engines when a
 * problem file says SETUP_FRAMEWORK(...).
 */
#ifndef B200_ENGINES_H
#define B200_ENGINES_H

#include <cuda_runtime.h>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>
#include <cfloat>

#include "engine_neibs.h"
#include "engine_forces.h"
#include "engine_integration.h"
#include "engine_filter.h"
#include "engine_postprocess.h"
#include "simframework.h"
#include "simparams.h"
#include "physparams.h"
#include "buffer.h"
#include "define_buffers.h"
#include "timing.h"
#include "linearization.h"

#include "b200sph.h"

namespace b200 {

inline void check(int rc)
{
	if (rc == B200SPH_OK) return;
	const std::string msg = b200sph_last_error();
	if (rc == B200SPH_EINVAL) throw std::invalid_argument(msg);
	throw std::runtime_error(msg);
}

//! Per-device contexts shared by the three engines
class Contexts
{
	std::mutex m_mutex;
	std::map<int, b200sph_ctx*> m_ctx;
	b200sph_params m_params;
	bool m_have_params;
public:
	Contexts() : m_have_params(false) {}
	~Contexts() { for (auto &kv : m_ctx) b200sph_destroy(kv.second); }

	//! Flatten SimParams/PhysParams the way the reference's setconstants upload them
	//! (src/cuda/buildneibs.cu:64-98, src/cuda/forces.cu:269-420, src/cuda/euler.cu:52-95)
	void configure(const SimParams *sp, const PhysParams *pp, float3 const& worldOrigin, uint3 const& gridSize,
		float3 const& cellSize, idx_t const& allocatedParticles)
	{
		std::lock_guard<std::mutex> lock(m_mutex);
		b200sph_params p;
		memset(&p, 0, sizeof(p));
		p.abi_version = B200SPH_ABI_VERSION;
		p.world_origin[0] = worldOrigin.x; p.world_origin[1] = worldOrigin.y; p.world_origin[2] = worldOrigin.z;
		p.cell_size[0] = cellSize.x; p.cell_size[1] = cellSize.y; p.cell_size[2] = cellSize.z;
		p.grid_size[0] = gridSize.x; p.grid_size[1] = gridSize.y; p.grid_size[2] = gridSize.z;
		// linearisation: COORD1..3 are macros expanding to x, y or z (src/linearization.h)
		struct { int x, y, z; } axis = { 0, 1, 2 };
		p.coord[0] = axis.COORD1; p.coord[1] = axis.COORD2; p.coord[2] = axis.COORD3;
		p.periodic = sp->periodicbound;
		p.neiblistsize = sp->neiblistsize;
		p.neibboundpos = sp->neibboundpos;
		p.neiblist_stride = (uint32_t)allocatedParticles;
		p.nl_sq_influence_radius = (float)sp->nlSqInfluenceRadius;
		p.kerneltype = sp->kerneltype;
		p.sph_formulation = sp->sph_formulation;
		p.densitydiffusiontype = sp->densitydiffusiontype;
		p.boundarytype = sp->boundarytype;
		p.rheologytype = sp->rheologytype;
		p.turbmodel = sp->turbmodel;
		p.compvisc = sp->compvisc;
		p.viscmodel = sp->viscmodel;
		p.viscavgop = sp->viscavgop;
		p.is_const_visc = sp->is_const_visc;
		p.slength = (float)sp->slength;
		p.influenceradius = (float)sp->influenceRadius;
		p.deltap = 0.0f;          // only used by SA boundaries
		p.density_diff_coeff = sp->densityDiffCoeff;
		p.dtadaptfactor = sp->dtadaptfactor;
		p.num_fluids = (uint32_t)pp->numFluids();
		if (p.num_fluids > B200SPH_MAX_FLUIDS) throw std::runtime_error("too many fluids for the B200 engines");
		float max_ss = 0, max_kin = 0;
		for (uint32_t f = 0; f < p.num_fluids; ++f) {
			p.rho0[f] = pp->rho0[f]; p.bcoeff[f] = pp->bcoeff[f]; p.gammacoeff[f] = pp->gammacoeff[f];
			p.sscoeff[f] = pp->sscoeff[f]; p.sspowercoeff[f] = pp->sspowercoeff[f]; p.visccoeff[f] = pp->visccoeff[f];
			max_ss = fmaxf(max_ss, pp->sscoeff[f]);
			max_kin = fmaxf(max_kin, pp->kinematicvisc[f]);
		}
		p.gravity[0] = pp->gravity.x; p.gravity[1] = pp->gravity.y; p.gravity[2] = pp->gravity.z;
		p.artvisccoeff = pp->artvisccoeff;
		p.epsartvisc = pp->epsartvisc;
		// `m_max_sound_speed *= 1.1` (float *= double literal, src/GPUWorker.cc:3010-3011): one rounding, in double.
		// Only the device-resident dt path reads these two; the reference-style dtreduce() forwards the caller's values.
		p.max_sound_speed_cfl = (float)((double)max_ss * 1.1);
		p.max_kinvisc = sp->rheologytype == INVISCID ? 0.0f : max_kin;
		if (sp->viscmodel == MONAGHAN) p.max_kinvisc *= pp->monaghan_visc_coeff;        // src/GPUWorker.cc:2011-2023
		else if (sp->viscmodel == ESPANOL_REVENGA) p.max_kinvisc *= 5;
		p.dtadapt = (sp->simflags & ENABLE_DTADAPT) ? 1 : 0;
		// the WHOLE flag word (same bit values, src/simflags.h:62-160): b200sph_validate refuses what is not implemented
		static_assert(ENABLE_MULTIFLUID == B200SPH_ENABLE_MULTIFLUID && ENABLE_REPACKING == B200SPH_ENABLE_REPACKING &&
			ENABLE_MOVING_BODIES == B200SPH_ENABLE_MOVING_BODIES, "simflags bit values changed");
		p.simflags = (uint32_t)sp->simflags;
		p.epsxsph = pp->epsxsph;
		p.monaghan_visc_coeff = pp->monaghan_visc_coeff;
		for (uint32_t f = 0; f < p.num_fluids; ++f) p.visc2coeff[f] = pp->visc2coeff[f];
		p.r0 = pp->r0; p.dcoeff = pp->dcoeff; p.p1coeff = pp->p1coeff; p.p2coeff = pp->p2coeff; p.partsurf = pp->partsurf;
		check(b200sph_validate(&p));      // unsupported option combinations fail here, loudly
		m_params = p;
		m_have_params = true;
	}

	//! context of the calling thread's device (created on first use, like the per-device __constant__ uploads)
	b200sph_ctx *get()
	{
		int dev = -1;
		if (cudaGetDevice(&dev) != cudaSuccess) throw std::runtime_error("cudaGetDevice failed");
		std::lock_guard<std::mutex> lock(m_mutex);
		auto it = m_ctx.find(dev);
		if (it != m_ctx.end()) return it->second;
		if (!m_have_params) throw std::runtime_error("B200 engines used before setconstants");
		b200sph_ctx *ctx = NULL;
		check(b200sph_create(&m_params, &ctx));
		m_ctx[dev] = ctx;
		return ctx;
	}
};

class NeibsEngine : public AbstractNeibsEngine
{
	std::shared_ptr<Contexts> m_c;
public:
	explicit NeibsEngine(std::shared_ptr<Contexts> c) : m_c(c) {}

	void setconstants(const SimParams *simparams, const PhysParams *physparams,
		float3 const& worldOrigin, uint3 const& gridSize, float3 const& cellSize,
		idx_t const& allocatedParticles) override
	{
		m_c->configure(simparams, physparams, worldOrigin, gridSize, cellSize, allocatedParticles); m_c->get();
	}

	void getconstants(SimParams *simparams, PhysParams *) override
	{ uint32_t v; check(b200sph_get_neibboundpos(m_c->get(), &v)); simparams->neibboundpos = v; }

	void resetinfo() override { check(b200sph_neibs_resetinfo(m_c->get())); }

	void getinfo(TimingInfo &ti) override
	{
		b200sph_neibs_info i;
		check(b200sph_neibs_getinfo(m_c->get(), &i));
		ti.numInteractions = i.num_interactions;
		ti.maxFluidBoundaryNeibs = i.max_fluid_boundary_neibs;
		ti.maxVertexNeibs = i.max_vertex_neibs;
		ti.hasTooManyNeibs = i.has_too_many_neibs;
		for (int t = 0; t < 3; ++t) ti.hasMaxNeibs[t] = i.has_max_neibs[t];
	}

	void calcHash(const BufferList& bufread, BufferList& bufwrite, const uint numParticles) override
	{
		check(b200sph_calc_hash(m_c->get(), bufwrite.getData<BUFFER_POS>(), bufwrite.getData<BUFFER_HASH>(),
			bufwrite.getData<BUFFER_PARTINDEX>(), bufread.getData<BUFFER_INFO>(),
			bufread.getData<BUFFER_COMPACT_DEV_MAP>(), numParticles));
	}

	void fixHash(const BufferList& bufread, BufferList& bufwrite, const uint numParticles) override
	{
		check(b200sph_fix_hash(m_c->get(), bufwrite.getData<BUFFER_HASH>(), bufwrite.getData<BUFFER_PARTINDEX>(),
			bufread.getData<BUFFER_INFO>(), bufread.getData<BUFFER_COMPACT_DEV_MAP>(), numParticles));
	}

	void sort(const BufferList&, BufferList& bufwrite, uint numParticles) override
	{
		check(b200sph_sort(m_c->get(), bufwrite.getData<BUFFER_HASH>(), bufwrite.getData<BUFFER_INFO>(),
			bufwrite.getData<BUFFER_PARTINDEX>(), numParticles));
	}

	void reorderDataAndFindCellStart(uint *segmentStart, BufferList& sorted_buffers,
		const BufferList& unsorted_buffers, const uint numParticles, uint *newNumParticles) override
	{
		// optional per-particle buffers the reference permutes too (src/cuda/buildneibs_kernel.cu:840-992)
		b200sph_reorder_extra extras[12];
		uint32_t ne = 0;
#define B200_EXTRA(KEY) do { auto *src = unsorted_buffers.getData<KEY>(); auto *dst = sorted_buffers.getData<KEY>(); \
		if (src && dst) { extras[ne].unsorted = src; extras[ne].sorted = dst; extras[ne].elem_size = sizeof(*src); ++ne; } } while (0)
		B200_EXTRA(BUFFER_VOLUME);
		B200_EXTRA(BUFFER_INTERNAL_ENERGY);
		B200_EXTRA(BUFFER_TKE);
		B200_EXTRA(BUFFER_EPSILON);
		B200_EXTRA(BUFFER_TURBVISC);
		B200_EXTRA(BUFFER_EFFPRES);
		B200_EXTRA(BUFFER_EULERVEL);
		B200_EXTRA(BUFFER_NEXTID);
#undef B200_EXTRA
		if (unsorted_buffers.getData<BUFFER_VERTICES>() || unsorted_buffers.getData<BUFFER_BOUNDELEMENTS>())
			throw std::runtime_error("B200 engines: SA boundary buffers are out of scope");
		check(b200sph_reorder(m_c->get(), sorted_buffers.getData<BUFFER_CELLSTART>(), sorted_buffers.getData<BUFFER_CELLEND>(),
			segmentStart, sorted_buffers.getData<BUFFER_POS>(), sorted_buffers.getData<BUFFER_VEL>(),
			unsorted_buffers.getData<BUFFER_POS>(), unsorted_buffers.getData<BUFFER_VEL>(), extras, ne,
			sorted_buffers.getData<BUFFER_INFO>(), sorted_buffers.getData<BUFFER_HASH>(),
			sorted_buffers.getData<BUFFER_PARTINDEX>(), numParticles, newNumParticles));
	}

	void buildNeibsList(const BufferList& bufread, BufferList& bufwrite, const uint numParticles,
		const uint particleRangeEnd, const uint, const float, const float) override
	{
		check(b200sph_build_neibs(m_c->get(), bufread.getData<BUFFER_POS>(), bufread.getData<BUFFER_INFO>(),
			bufread.getData<BUFFER_HASH>(), bufread.getData<BUFFER_CELLSTART>(), bufread.getData<BUFFER_CELLEND>(),
			bufwrite.getData<BUFFER_NEIBSLIST>(), numParticles, particleRangeEnd));
	}
};

class ForcesEngine : public AbstractForcesEngine
{
	std::shared_ptr<Contexts> m_c;
	static void unsupported(const char *what)
	{ throw std::runtime_error(std::string("B200 forces engine: ") + what + " is out of scope (SURVEY.md section 8)"); }
public:
	explicit ForcesEngine(std::shared_ptr<Contexts> c) : m_c(c) {}

	void setconstants(const SimParams *simparams, const PhysParams *physparams,
		float3 const& worldOrigin, uint3 const& gridSize, float3 const& cellSize,
		idx_t const& allocatedParticles) override
	{
		m_c->configure(simparams, physparams, worldOrigin, gridSize, cellSize, allocatedParticles); m_c->get();
	}

	void getconstants(PhysParams *pp) override { }

	void setplanes(PlaneList const& planes) override
	{
		// plane_t = { float3 normal; int3 gridPos; float3 pos; } (src/planes.h:42-46)
		std::vector<float> nrm, pos; std::vector<int> gp;
		for (auto const& pl : planes) {
			nrm.push_back(pl.normal.x); nrm.push_back(pl.normal.y); nrm.push_back(pl.normal.z);
			gp.push_back(pl.gridPos.x); gp.push_back(pl.gridPos.y); gp.push_back(pl.gridPos.z);
			pos.push_back(pl.pos.x); pos.push_back(pl.pos.y); pos.push_back(pl.pos.z);
		}
		check(b200sph_set_planes(m_c->get(), nrm.data(), gp.data(), pos.data(), (int)planes.size()));
	}
	void setgravity(float3 const& g) override
	{ const float v[3] = { g.x, g.y, g.z }; check(b200sph_set_gravity(m_c->get(), v)); }
	// moving / force-feedback bodies (src/cuda/forces.cu:430-447, 967-1003)
	void setrbcg(const int3* cgGridPos, const float3* cgPos, int numbodies) override
	{ check(b200sph_set_rbcg(m_c->get(), (const int*)cgGridPos, (const float*)cgPos, numbodies)); }
	void setrbstart(const int* rbfirstindex, int numbodies) override
	{ check(b200sph_set_rbstart(m_c->get(), rbfirstindex, numbodies)); }
	void reduceRbForces(BufferList& bufwrite, uint *lastindex, float3 *totalforce, float3 *totaltorque,
		uint numforcesbodies, uint numForcesBodiesParticles) override
	{
		check(b200sph_reduce_rb_forces(m_c->get(), bufwrite.getData<BUFFER_RB_FORCES>(), bufwrite.getData<BUFFER_RB_TORQUES>(),
			bufwrite.getConstData<BUFFER_RB_KEYS>(), lastindex, (float*)totalforce, (float*)totaltorque,
			numforcesbodies, numForcesBodiesParticles));
	}

	// no texture references on this architecture: neighbours are gathered through the read-only path directly
	void bind_textures(const BufferList&, uint, RunMode) override {}
	void unbind_textures(RunMode) override {}

	void setDEM(const float*, int, int) override { unsupported("DEM"); }
	void unsetDEM() override {}

	uint round_particles(uint n) override { return b200sph_round_particles(n); }

	void compute_density(const BufferList&, BufferList&, uint, float, float) override { unsupported("SPH_GRENIER compute_density"); }
	void compute_density_diffusion(const BufferList&, BufferList&, const uint, const uint, const float, const float,
		const float, const float) override { unsupported("density-sum density diffusion"); }

	uint basicstep(const BufferList& bufread, BufferList& bufwrite, uint numParticles, uint fromParticle, uint toParticle,
		float, float, float, float, const float, uint*, uint cflOffset, const RunMode run_mode, const int step, const float dt,
		const bool compute_object_forces) override
	{
		if (run_mode == REPACK) unsupported("repacking");
		uint32_t nblocks = 0;
		(void)compute_object_forces;
		b200sph_forces_args a;
		memset(&a, 0, sizeof(a));
		a.pos = bufread.getData<BUFFER_POS>(); a.vel = bufread.getData<BUFFER_VEL>(); a.info = bufread.getData<BUFFER_INFO>();
		a.hash = bufread.getData<BUFFER_HASH>(); a.cell_start = bufread.getData<BUFFER_CELLSTART>();
		a.neibs_list = bufread.getData<BUFFER_NEIBSLIST>();
		a.forces = bufwrite.getData<BUFFER_FORCES>(); a.cfl = bufwrite.getData<BUFFER_CFL>();
		// the reference's finalize kernel scatters body forces whenever the particle carries FG_COMPUTE_FORCE
		// (forces_kernel.def:4116-4141); the RB buffers exist exactly when there are force-feedback bodies
		a.rb_forces = bufwrite.getData<BUFFER_RB_FORCES>(); a.rb_torques = bufwrite.getData<BUFFER_RB_TORQUES>();
		a.xsph = bufwrite.getData<BUFFER_XSPH>();           // exists iff ENABLE_XSPH (src/GPUWorker.cc:135-136)
		a.num_particles = numParticles; a.from_particle = fromParticle; a.to_particle = toParticle; a.cfl_offset = cflOffset;
		a.dt = dt; a.step = step;                           // read by BREZZI diffusion only
		check(b200sph_forces_ex(m_c->get(), &a, &nblocks));
		return nblocks;
	}

	uint getFmaxElements(const uint n) override { return b200sph_fmax_elements(n); }
	uint getFmaxTempElements(const uint n) override { return b200sph_fmax_temp_elements(n); }

	float dtreduce(float slength, float dtadaptfactor, float sspeed_cfl, float max_kinematic, BufferList const& bufread,
		BufferList& bufwrite, uint numBlocks, uint) override
	{
		float dt = FLT_MAX;
		check(b200sph_dtreduce_ex(m_c->get(), bufread.getData<BUFFER_CFL>(), bufwrite.getData<BUFFER_CFL_TEMP>(), numBlocks,
			slength, dtadaptfactor, sspeed_cfl, max_kinematic, &dt));
		return dt;
	}
};

class IntegrationEngine : public AbstractIntegrationEngine
{
	std::shared_ptr<Contexts> m_c;
	static void unsupported(const char *what)
	{ throw std::runtime_error(std::string("B200 integration engine: ") + what + " is out of scope (SURVEY.md section 8)"); }
public:
	explicit IntegrationEngine(std::shared_ptr<Contexts> c) : m_c(c) {}

	// everything this engine needs was already flattened by the neibs/forces setconstants
	void setconstants(const PhysParams *pp, float3 const& o, uint3 const& g, float3 const& c, idx_t const& a, int const& n, float const& h) override
	{ }
	void getconstants(PhysParams *pp) override { }

	// moving bodies (src/cuda/euler.cu:76-95)
	// the integration engine's OWN copy of the centres of gravity (cg(n) for the whole step, src/cuda/euler_kernel.def:488)
	void setrbcg(const int3* g, const float3* c, int n) override
	{ check(b200sph_set_rbcg_euler(m_c->get(), (const int*)g, (const float*)c, n)); }
	void setrbtrans(const float3* t, int n) override
	{ check(b200sph_set_rbtrans(m_c->get(), (const float*)t, n)); }
	void setrbsteprot(const float* r, int n) override
	{ check(b200sph_set_rbsteprot(m_c->get(), r, n)); }
	void setrblinearvel(const float3* v, int n) override
	{ check(b200sph_set_rblinearvel(m_c->get(), (const float*)v, n)); }
	void setrbangularvel(const float3* v, int n) override
	{ check(b200sph_set_rbangularvel(m_c->get(), (const float*)v, n)); }

	void density_sum(const BufferList&, BufferList&, const uint, const uint, const float, const int, const float,
		const float, const float, const float, const float) override { unsupported("density summation (SA)"); }
	void integrate_gamma(const BufferList&, BufferList&, const uint, const uint, const float, const int, const float,
		const float, const float, const float, const RunMode) override { unsupported("gamma integration (SA)"); }
	void apply_density_diffusion(const BufferList&, BufferList&, const uint, const uint, const float) override
	{ unsupported("density-sum density diffusion"); }

	void basicstep(const BufferList& bufread, BufferList& bufwrite, const uint numParticles, const uint particleRangeEnd,
		const float dt, const int step, const float, const float, const float, const RunMode run_mode) override
	{
		if (run_mode == REPACK) unsupported("repacking");
		check(b200sph_euler_ex(m_c->get(), bufread.getData<BUFFER_POS>(), bufread.getData<BUFFER_VEL>(),
			bufread.getData<BUFFER_INFO>(), bufread.getData<BUFFER_HASH>(), bufread.getData<BUFFER_FORCES>(),
			bufread.getData<BUFFER_XSPH>(),
			bufwrite.getData<BUFFER_POS>(), bufwrite.getData<BUFFER_VEL>(), numParticles, particleRangeEnd, dt, step, 0));
	}

	void disableFreeSurfParts(float4*, const particleinfo*, const uint, const uint) override { unsupported("repacking"); }
};

//! SHEPARD_FILTER / MLS_FILTER (src/cuda/forces.cu:1026-1146)
class FilterEngine : public AbstractFilterEngine
{
	std::shared_ptr<Contexts> m_c;
	FilterType m_type;
public:
	FilterEngine(std::shared_ptr<Contexts> c, FilterType type, uint frequency) :
		AbstractFilterEngine(frequency), m_c(c), m_type(type)
	{
		if (type != SHEPARD_FILTER && type != MLS_FILTER)
			throw std::invalid_argument("B200 filter engine: unknown filter type");
	}

	void setconstants() override {}
	void getconstants() override {}

	void process(const BufferList& bufread, BufferList& bufwrite, uint numParticles, uint particleRangeEnd,
		float, float) override
	{
		auto fn = m_type == SHEPARD_FILTER ? b200sph_filter_shepard : b200sph_filter_mls;
		check(fn(m_c->get(), bufread.getData<BUFFER_POS>(), bufread.getData<BUFFER_VEL>(), bufwrite.getData<BUFFER_VEL>(),
			bufread.getData<BUFFER_INFO>(), bufread.getData<BUFFER_HASH>(), bufread.getData<BUFFER_CELLSTART>(),
			bufread.getData<BUFFER_NEIBSLIST>(), numParticles, particleRangeEnd));
	}
};

//! TESTPOINTS post-processing (src/cuda/post_process.cu:148-216): VEL / TKE / EPSILON updated in place
class TestpointsEngine : public AbstractPostProcessEngine
{
	std::shared_ptr<Contexts> m_c;
public:
	explicit TestpointsEngine(std::shared_ptr<Contexts> c, flag_t options = NO_FLAGS) :
		AbstractPostProcessEngine(options), m_c(c) {}

	void setconstants(const SimParams *sp, const PhysParams *pp, idx_t const& n) const override
	{ }
	void getconstants() override {}

	void process(const BufferList& bufread, BufferList& bufwrite, uint numParticles, uint particleRangeEnd,
		uint, const GlobalData * const) override
	{
		check(b200sph_testpoints(m_c->get(), bufread.getData<BUFFER_POS>(), bufwrite.getData<BUFFER_VEL>(),
			bufwrite.getData<BUFFER_TKE>(), bufwrite.getData<BUFFER_EPSILON>(), bufread.getData<BUFFER_INFO>(),
			bufread.getData<BUFFER_HASH>(), bufread.getData<BUFFER_CELLSTART>(), bufread.getData<BUFFER_NEIBSLIST>(),
			numParticles, particleRangeEnd));
	}

	flag_t get_written_buffers() const override { return NO_FLAGS; }
	flag_t get_updated_buffers() const override { return BUFFER_VEL | BUFFER_TKE | BUFFER_EPSILON; }

	void hostAllocate(const GlobalData * const) override {}
	void hostProcess(const GlobalData * const) override {}
	void write(WriterMap, double) override {}
};

} // namespace b200
#endif
