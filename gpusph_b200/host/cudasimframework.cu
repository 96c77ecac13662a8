/*
 * cudasimframework.cu — the B200 engines' implementation of GPUSPH's framework seam.
 *
 * GPUSPH problems pick their engines at SOURCE level: a problem file does `#include "cudasimframework.cu"` and
 *   SETUP_FRAMEWORK(viscosity<ARTVISC>, boundary<DYN_BOUNDARY>, add_flags<...>).select_options(RHODIFF, USE_PLANES, add_flags<ENABLE_PLANES>())
 * (src/problems/DamBreak3D.cu:35,53-61, src/problems/Poiseuille.inc:50,102-121; the macro is src/ProblemCore.h:117).
 * Put this directory on the include path BEFORE src/cuda and the unmodified problem files get the B200 engines: no
 * reference source is edited, none of the reference's CUDA kernels is compiled (see tools/build_dropin.sh).
 *
 * What a problem file may say (reference: src/cuda/cudasimframework.cu:355-466, 486-607) and what it means here:
 *   CUDASimFramework<named options, any order, any number up to what the problem writes>   a VALUE holding run-time options
 *   kernel<> formulation<> densitydiffusion<> rheology<> turbulence_model<> computational_visc<> visc_model<>
 *   visc_average<> viscosity<> boundary<> periodicity<> add_flags<> disable_flags<>            each edits that value
 *   .select_options(enum value | bool, named option, ...)                                   edits it at run time
 *   conversion to SimFramework*                                                             builds the framework
 * The reference instantiates one set of engine templates per option combination (the select_options chain compiles
 * every alternative: 4 diffusion models x 2 for DamBreak3D, 4 x 2 x 3 x 3 for Poiseuille). The B200 engines take their
 * options at run time (b200sph_params, include/b200sph.h), so here the option set is plain data: applying a named
 * option is an assignment, and a combination the engines do not implement is refused when the engines are configured
 * (b200sph_validate -> std::runtime_error from setconstants) instead of at compile time.
 */
#ifndef B200_CUDASIMFRAMEWORK_CU
#define B200_CUDASIMFRAMEWORK_CU

#include <mutex>
#include <stdexcept>
#include <type_traits>

#include "simframework.h"
#include "predcorr_alloc_policy.h"
#include "simflags.h"
#include "option_range.h"
#include "visc_spec.h"

#include "b200_engines.h"

// the reference's framework file says this at file scope (src/cuda/cudasimframework.cu:55) and problem files rely on it
// (src/problems/Poiseuille.inc:99,190 write `string`, `invalid_argument` unqualified)
using namespace std;

namespace b200 {

//! Everything a problem can choose about the framework; defaults as the reference's (cudasimframework.cu:346-360)
struct FrameworkOptions
{
	KernelType kerneltype = WENDLAND;
	SPHFormulation sph_formulation = SPH_F1;
	DensityDiffusionType densitydiffusiontype = DENSITY_DIFFUSION_NONE;
	RheologyType rheologytype = INVISCID;
	TurbulenceModel turbmodel = ARTIFICIAL;
	ComputationalViscosityType compvisc = KINEMATIC;
	ViscousModel viscmodel = MORRIS;
	AverageOperator viscavgop = ARITHMETIC;
	LegacyViscosityType legacyvisctype = INVALID_VISCOSITY;
	BoundaryType boundarytype = LJ_BOUNDARY;
	Periodicity periodicbound = PERIODIC_NONE;
	flag_t simflags = DEFAULT_FLAGS;

	//! constant-viscosity assumption (cudasimframework.cu:133-137)
	bool is_const_visc() const
	{
		return legacyvisctype == KINEMATICVISC ||
			(IS_SINGLEFLUID(simflags) && rheologytype == NEWTONIAN && turbmodel != KEPSILON);
	}
	//! Grenier's formulation with a legacy viscous specification keeps harmonic averaging (cudasimframework.cu:194-199)
	AverageOperator effective_viscavgop() const
	{
		return (sph_formulation == SPH_GRENIER && legacyvisctype != INVALID_VISCOSITY) ? HARMONIC : viscavgop;
	}

	// run-time overrides by option value (what selector_for<> maps to in the reference)
	void set(KernelType v) { kerneltype = v; }
	void set(SPHFormulation v) { sph_formulation = v; }
	void set(DensityDiffusionType v) { densitydiffusiontype = v; }
	void set(RheologyType v) { rheologytype = v; }
	void set(TurbulenceModel v) { turbmodel = v; }
	void set(ComputationalViscosityType v) { compvisc = v; }
	void set(ViscousModel v) { viscmodel = v; }
	void set(AverageOperator v) { viscavgop = v; }
	void set(BoundaryType v) { boundarytype = v; }
	void set(Periodicity v) { periodicbound = v; }
};

//! SimParams' only constructor reads the options as STATIC members of a framework type (src/simparams.h:261-274);
//! this is that type for an option set known at run time. Guarded by a lock: frameworks are built one at a time.
struct StaticOptions
{
	static KernelType kerneltype;
	static SPHFormulation sph_formulation;
	static DensityDiffusionType densitydiffusiontype;
	static RheologyType rheologytype;
	static TurbulenceModel turbmodel;
	static ComputationalViscosityType compvisc;
	static ViscousModel viscmodel;
	static AverageOperator viscavgop;
	static bool is_const_visc;
	static BoundaryType boundarytype;
	static Periodicity periodicbound;
	static flag_t simflags;

	static SimParams *make_simparams(FrameworkOptions const& o)
	{
		static std::mutex lock;
		std::lock_guard<std::mutex> guard(lock);
		kerneltype = o.kerneltype; sph_formulation = o.sph_formulation; densitydiffusiontype = o.densitydiffusiontype;
		rheologytype = o.rheologytype; turbmodel = o.turbmodel; compvisc = o.compvisc; viscmodel = o.viscmodel;
		viscavgop = o.effective_viscavgop(); is_const_visc = o.is_const_visc();
		boundarytype = o.boundarytype; periodicbound = o.periodicbound; simflags = o.simflags;
		return new SimParams((StaticOptions *)NULL);
	}
};
// one definition per program: the problem file is the only translation unit that includes this header
KernelType StaticOptions::kerneltype;
SPHFormulation StaticOptions::sph_formulation;
DensityDiffusionType StaticOptions::densitydiffusiontype;
RheologyType StaticOptions::rheologytype;
TurbulenceModel StaticOptions::turbmodel;
ComputationalViscosityType StaticOptions::compvisc;
ViscousModel StaticOptions::viscmodel;
AverageOperator StaticOptions::viscavgop;
bool StaticOptions::is_const_visc;
BoundaryType StaticOptions::boundarytype;
Periodicity StaticOptions::periodicbound;
flag_t StaticOptions::simflags;

//! The viscosity pre-computation (SPS stress tensor, per-particle effective viscosity, the Jacobi solver of granular
//! rheologies) is outside the hot path: GPUWorker only reaches it through CALC_VISC / JACOBI_*
//! commands, which the integrator issues for SPS, k-epsilon and non-Newtonian rheologies alone
//! (src/integrators/PredictorCorrectorIntegrator.cc). Those options are refused by b200sph_validate, so these methods
//! are never called; they throw rather than return silently.
class NoViscEngine : public AbstractViscEngine
{
	static void unsupported()
	{ throw std::runtime_error("B200 engines: viscosity pre-computation (SPS / k-epsilon / non-Newtonian) is out of scope"); }
public:
	void setconstants() override {}
	void getconstants() override {}
	float calc_visc(const BufferList&, BufferList&, const uint, const uint, const float, const float, const float) override
	{ unsupported(); return 0; }
	void enforce_jacobi_fs_boundary_conditions(const BufferList&, BufferList&, const uint, const uint, const float,
		const float, const float) override { unsupported(); }
	float enforce_jacobi_wall_boundary_conditions(const BufferList&, BufferList&, const uint, const uint, const float,
		const float, const float) override { unsupported(); return 0; }
	void build_jacobi_vectors(const BufferList&, BufferList&, const uint, const uint, const float, const float,
		const float) override { unsupported(); }
	float update_jacobi_effpres(const BufferList&, BufferList&, const uint, const uint, const float, const float,
		const float) override { unsupported(); return 0; }
};

//! The framework GPUSPH / GPUWorker see (src/simframework.h:62-133): three hot-path engines on the C ABI of
//! include/b200sph.h, the two density filters and the TESTPOINTS post-process on the same neighbour list.
class SimFrameworkB200 : public SimFramework
{
	std::shared_ptr<Contexts> m_contexts;
public:
	explicit SimFrameworkB200(FrameworkOptions const& o) : SimFramework(), m_contexts(std::make_shared<Contexts>())
	{
		m_neibsEngine = new NeibsEngine(m_contexts);
		m_integrationEngine = new IntegrationEngine(m_contexts);
		m_viscEngine = new NoViscEngine();
		m_forcesEngine = new ForcesEngine(m_contexts);
		m_bcEngine = NULL;      // the reference has one for SA boundaries only (cudasimframework.cu:82-98)
		m_allocPolicy = std::make_shared<PredCorrAllocPolicy>();
		m_simparams = StaticOptions::make_simparams(o);
	}

protected:
	AbstractFilterEngine* newFilterEngine(FilterType filtertype, int frequency) override
	{
		if (filtertype == SHEPARD_FILTER || filtertype == MLS_FILTER)
			return new FilterEngine(m_contexts, filtertype, frequency);
		throw std::runtime_error("Invalid filter type");
	}

	AbstractPostProcessEngine* newPostProcessEngine(PostProcessType pptype, flag_t options = NO_FLAGS) override
	{
		if (pptype == TESTPOINTS)
			return new TestpointsEngine(m_contexts, options);
		throw std::runtime_error("B200 engines: this post-processing engine is out of scope (only TESTPOINTS is on the path)");
	}
};

} // namespace b200

// ---- named options (the names and meanings of src/cuda/cudasimframework.cu:355-466) ----

#define B200_NAMED_OPTION(name, Type) \
template<Type value__> struct name { static void apply(b200::FrameworkOptions &o) { o.set(value__); } }

B200_NAMED_OPTION(kernel, KernelType);
B200_NAMED_OPTION(formulation, SPHFormulation);
B200_NAMED_OPTION(densitydiffusion, DensityDiffusionType);
B200_NAMED_OPTION(rheology, RheologyType);
B200_NAMED_OPTION(turbulence_model, TurbulenceModel);
B200_NAMED_OPTION(computational_visc, ComputationalViscosityType);
B200_NAMED_OPTION(visc_model, ViscousModel);
B200_NAMED_OPTION(visc_average, AverageOperator);
B200_NAMED_OPTION(boundary, BoundaryType);
B200_NAMED_OPTION(periodicity, Periodicity);
#undef B200_NAMED_OPTION

//! legacy viscous specification: sets the five viscous options at once (src/visc_spec.h:347-392)
template<LegacyViscosityType visctype>
struct viscosity
{
	static void apply(b200::FrameworkOptions &o)
	{
		typedef typename ConvertLegacyVisc<visctype>::type Spec;
		o.legacyvisctype = visctype;
		o.rheologytype = Spec::rheologytype; o.turbmodel = Spec::turbmodel; o.compvisc = Spec::compvisc;
		o.viscmodel = Spec::viscmodel; o.viscavgop = Spec::avgop;
	}
};

template<flag_t flags>
struct add_flags { static void apply(b200::FrameworkOptions &o) { o.simflags |= flags; } };

template<flag_t flags>
struct disable_flags { static void apply(b200::FrameworkOptions &o) { o.simflags = DISABLE_FLAGS(o.simflags, flags); } };

//! The factory SETUP_FRAMEWORK(...) assigns to the problem's SimFramework* (src/ProblemCore.h:117)
template<typename... Named>
class CUDASimFramework
{
	b200::FrameworkOptions m_options;

	template<typename First, typename... Rest> void apply_all() { First::apply(m_options); apply_all<Rest...>(); }
	template<typename... None> typename std::enable_if<sizeof...(None) == 0>::type apply_all() {}

public:
	CUDASimFramework() { apply_all<Named...>(); }

	//! the framework for the options selected so far
	operator SimFramework *() { return new b200::SimFrameworkB200(m_options); }

	b200::FrameworkOptions const& options() const { return m_options; }

	// Run-time selectors, same call forms as the reference's (cudasimframework.cu:551-606): a bool followed by the
	// named option it switches on, or an option value (applied as the named option of its type); any number of them.
	SimFramework *select_options() { return *this; }

	template<typename Extra, typename... Rest>
	SimFramework *select_options(bool selector, Extra, Rest... rest)
	{
		if (selector) Extra::apply(m_options);
		return select_options(rest...);
	}

	template<typename Option, typename... Rest>
	typename std::enable_if<option_range<Option>::defined, SimFramework *>::type
	select_options(Option selector, Rest... rest)
	{
		if (!is_in_range(selector)) throw std::runtime_error("invalid selector value");
		m_options.set(selector);
		return select_options(rest...);
	}
};

#endif
/* vim: set ft=cuda sw=4 ts=4 : */
