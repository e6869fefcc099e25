// Material plugins of the hot path.  Factories keep the reference's names, aliases, parameter keys and
// defaults; the objects describe themselves into prb_material (the BSDF math itself runs on the device,
// pearray_b200/csrc/materials.cuh).
#include "prh.h"

#include <sstream>

namespace PR {
namespace {
inline uint32 nodeContribFlags(const std::initializer_list<std::shared_ptr<FloatSpectralNode>>& nodes)
{ // INode::materialFlags(): SpectralVarying node -> MaterialSampleFlag::SpectralVarying (INode.h:33-43)
	for (const auto& n : nodes)
		if (n->isSpectralVarying())
			return PRB_MATF_SPECTRAL_VARYING;
	return 0;
}
inline float constScalar(const std::shared_ptr<FloatScalarNode>& n, const char* what)
{
	if (!n->isConst())
		PR_LOG(L_WARNING) << "Scalar parameter '" << what << "' is not constant; the device path evaluates it once at UV (0,0)" << std::endl;
	return n->eval(ShadingContext());
}

class LambertMaterial : public IMaterial { // lambert.cpp:14-89
public:
	LambertMaterial(const std::shared_ptr<FloatSpectralNode>& alb, bool twoSided)
		: mAlbedo(alb)
		, mTwoSided(twoSided)
	{
	}
	void describe(prb_material& out, NodeEmitter& e) const override
	{
		out.type	= PRB_MAT_DIFFUSE;
		out.flags	= mTwoSided ? PRB_MATF_TWO_SIDED : 0;
		out.node[0] = mAlbedo->emit(e);
	}
	std::string dumpInformation() const override
	{
		std::stringstream s;
		s << "  <DiffuseMaterial>:\n    Albedo:   " << mAlbedo->dumpInformation() << "\n    TwoSided: " << (mTwoSided ? "true" : "false") << "\n";
		return s.str();
	}

private:
	std::shared_ptr<FloatSpectralNode> mAlbedo;
	bool mTwoSided;
};
class LambertMaterialPlugin : public IMaterialPlugin {
public:
	std::shared_ptr<IMaterial> create(const std::string&, const SceneLoadContext& ctx) override
	{
		const auto albedo = ctx.lookupSpectralNode({ "albedo", "base", "diffuse" }, 1);
		return std::make_shared<LambertMaterial>(albedo, ctx.parameters().getBool("two_sided", true));
	}
	const std::vector<std::string>& getNames() const override
	{
		static const std::vector<std::string> names({ "diffuse", "lambert" });
		return names;
	}
	std::string specification(const std::string&) const override { return "Lambert BSDF: albedo|base|diffuse (spectral, 1), two_sided (bool, true)"; }
};

class DielectricMaterial : public IMaterial { // dielectric.cpp:18-135
public:
	DielectricMaterial(const std::shared_ptr<FloatSpectralNode>& spec, const std::shared_ptr<FloatSpectralNode>& trans,
					   const std::shared_ptr<FloatSpectralNode>& ior, bool hasTrans, bool thin)
		: mSpecularity(spec)
		, mTransmission(trans)
		, mIOR(ior)
		, mHasTrans(hasTrans)
		, mThin(thin)
	{
	}
	bool hasOnlyDeltaDistribution() const override { return true; }
	void describe(prb_material& out, NodeEmitter& e) const override
	{
		out.type	= PRB_MAT_DIELECTRIC;
		out.flags	= PRB_MATF_ONLY_DELTA | nodeContribFlags({ mIOR }) | (mHasTrans ? PRB_MATF_TRANSMISSION_COLOR : 0) | (mThin ? PRB_MATF_THIN : 0);
		out.node[0] = mSpecularity->emit(e);
		out.node[1] = mTransmission->emit(e);
		out.node[2] = mIOR->emit(e);
	}
	std::string dumpInformation() const override
	{
		std::stringstream s;
		s << "  <DielectricMaterial>:\n    Specularity:     " << mSpecularity->dumpInformation() << "\n    Transmission:    " << mTransmission->dumpInformation()
		  << "\n    IOR:             " << mIOR->dumpInformation() << "\n    IsThin:          " << (mThin ? "true" : "false") << "\n";
		return s.str();
	}

private:
	std::shared_ptr<FloatSpectralNode> mSpecularity, mTransmission, mIOR;
	bool mHasTrans, mThin;
};
class DielectricMaterialPlugin : public IMaterialPlugin {
public:
	std::shared_ptr<IMaterial> create(const std::string&, const SceneLoadContext& ctx) override
	{
		// Construct rough dielectric instead (dielectric.cpp:171-175)
		if (ctx.parameters().hasParameter("roughness") || ctx.parameters().hasParameter("roughness_x") || ctx.parameters().hasParameter("roughness_y"))
			return ctx.loadMaterial("roughdielectric", ctx.parameters());
		const bool hasTrans = ctx.parameters().hasParameter("transmission");
		const bool thin		= ctx.parameters().getBool("thin", false);
		const auto spec		= ctx.lookupSpectralNode("specularity", 1);
		return std::make_shared<DielectricMaterial>(spec, hasTrans ? ctx.lookupSpectralNode("transmission", 1) : spec,
													ctx.lookupSpectralNode("index", 1.55f), hasTrans, thin);
	}
	const std::vector<std::string>& getNames() const override
	{
		static const std::vector<std::string> names({ "glass", "dielectric" });
		return names;
	}
	std::string specification(const std::string&) const override { return "Delta Dielectric BSDF: index (1.55), specularity (1), transmission (1), thin (false)"; }
};

class ConductorMaterial : public IMaterial { // conductor.cpp:16-95
public:
	ConductorMaterial(const std::shared_ptr<FloatSpectralNode>& eta, const std::shared_ptr<FloatSpectralNode>& k, const std::shared_ptr<FloatSpectralNode>& spec)
		: mEta(eta)
		, mK(k)
		, mSpecularity(spec)
	{
	}
	bool hasOnlyDeltaDistribution() const override { return true; }
	void describe(prb_material& out, NodeEmitter& e) const override
	{
		out.type	= PRB_MAT_CONDUCTOR;
		out.flags	= PRB_MATF_ONLY_DELTA | nodeContribFlags({ mEta, mK });
		out.node[0] = mEta->emit(e);
		out.node[1] = mK->emit(e);
		out.node[2] = mSpecularity->emit(e);
	}
	std::string dumpInformation() const override
	{
		std::stringstream s;
		s << "  <ConductorMaterial>:\n    Eta:             " << mEta->dumpInformation() << "\n    K:               " << mK->dumpInformation()
		  << "\n    Specularity:     " << mSpecularity->dumpInformation() << "\n";
		return s.str();
	}

private:
	std::shared_ptr<FloatSpectralNode> mEta, mK, mSpecularity;
};
class ConductorMaterialPlugin : public IMaterialPlugin {
public:
	std::shared_ptr<IMaterial> create(const std::string&, const SceneLoadContext& ctx) override
	{
		if (ctx.parameters().hasParameter("roughness") || ctx.parameters().hasParameter("roughness_x") || ctx.parameters().hasParameter("roughness_y"))
			return ctx.loadMaterial("roughconductor", ctx.parameters()); // conductor.cpp:101-105
		return std::make_shared<ConductorMaterial>(ctx.lookupSpectralNode({ "eta", "index", "ior" }, 1.2f), ctx.lookupSpectralNode({ "k", "kappa" }, 2.605f),
												   ctx.lookupSpectralNode("specularity", 1));
	}
	const std::vector<std::string>& getNames() const override
	{
		static const std::vector<std::string> names({ "conductor", "metal" });
		return names;
	}
	std::string specification(const std::string&) const override { return "Delta Conductor BSDF: eta|index|ior (1.2), k|kappa (2.605), specularity (1)"; }
};

struct Roughness {
	float rx = 0, ry = 0;
	bool anisotropic = false;
};
static Roughness parseRoughness(const SceneLoadContext& ctx)
{ // roughconductor.cpp:148-161 / roughdielectric.cpp:285-298: anisotropic iff roughness_y is given
	Roughness r;
	std::shared_ptr<FloatScalarNode> rx, ry;
	if (ctx.parameters().hasParameter("roughness_x"))
		rx = ctx.lookupScalarNode("roughness_x", 0);
	else
		rx = ctx.lookupScalarNode("roughness", 0);
	if (ctx.parameters().hasParameter("roughness_y")) {
		ry			  = ctx.lookupScalarNode("roughness_y", 0);
		r.anisotropic = true;
	} else {
		ry = rx;
	}
	r.rx = constScalar(rx, "roughness_x");
	r.ry = constScalar(ry, "roughness_y");
	return r;
}

class RoughConductorMaterial : public IMaterial { // roughconductor.cpp:16-143
public:
	RoughConductorMaterial(const std::shared_ptr<FloatSpectralNode>& eta, const std::shared_ptr<FloatSpectralNode>& k,
						   const std::shared_ptr<FloatSpectralNode>& spec, const Roughness& r, bool vndf)
		: mEta(eta)
		, mK(k)
		, mSpecularity(spec)
		, mR(r)
		, mVNDF(vndf)
	{
	}
	void describe(prb_material& out, NodeEmitter& e) const override
	{
		out.type	= PRB_MAT_ROUGHCONDUCTOR;
		out.flags	= nodeContribFlags({ mEta, mK }) | (mVNDF ? PRB_MATF_VNDF : 0) | (mR.anisotropic ? PRB_MATF_ANISOTROPIC : 0);
		out.node[0] = mEta->emit(e);
		out.node[1] = mK->emit(e);
		out.node[2] = mSpecularity->emit(e);
		out.f[0]	= mR.rx;
		out.f[1]	= mR.ry;
	}
	std::string dumpInformation() const override
	{
		std::stringstream s;
		s << "  <RoughConductorMaterial>:\n    Eta:             " << mEta->dumpInformation() << "\n    K:               " << mK->dumpInformation()
		  << "\n    Specularity:     " << mSpecularity->dumpInformation() << "\n    RoughnessX:      " << mR.rx << "\n    RoughnessY:      " << mR.ry
		  << "\n    VNDF:            " << (mVNDF ? "true" : "false") << "\n";
		return s.str();
	}

private:
	std::shared_ptr<FloatSpectralNode> mEta, mK, mSpecularity;
	Roughness mR;
	bool mVNDF;
};
class RoughConductorMaterialPlugin : public IMaterialPlugin {
public:
	std::shared_ptr<IMaterial> create(const std::string&, const SceneLoadContext& ctx) override
	{
		const Roughness r = parseRoughness(ctx);
		return std::make_shared<RoughConductorMaterial>(ctx.lookupSpectralNode({ "eta", "index", "ior" }, 1.2f), ctx.lookupSpectralNode({ "k", "kappa" }, 2.605f),
														ctx.lookupSpectralNode("specularity", 1), r, ctx.parameters().getBool("vndf", true));
	}
	const std::vector<std::string>& getNames() const override
	{
		static const std::vector<std::string> names({ "roughconductor", "roughmirror", "roughmetal" });
		return names;
	}
	std::string specification(const std::string&) const override
	{
		return "Rough Conductor BSDF: roughness | roughness_x roughness_y (0), eta|index|ior (1.2), k|kappa (2.605), specularity (1), vndf (true)";
	}
};

class RoughDielectricMaterial : public IMaterial { // roughdielectric.cpp:42-279
public:
	RoughDielectricMaterial(const std::shared_ptr<FloatSpectralNode>& spec, const std::shared_ptr<FloatSpectralNode>& trans,
							const std::shared_ptr<FloatSpectralNode>& ior, bool hasTrans, const Roughness& r, bool vndf)
		: mSpecularity(spec)
		, mTransmission(trans)
		, mIOR(ior)
		, mHasTrans(hasTrans)
		, mR(r)
		, mVNDF(vndf)
	{
	}
	void describe(prb_material& out, NodeEmitter& e) const override
	{
		out.type  = PRB_MAT_ROUGHDIELECTRIC;
		out.flags = nodeContribFlags({ mIOR }) | (mHasTrans ? PRB_MATF_TRANSMISSION_COLOR : 0) | (mVNDF ? PRB_MATF_VNDF : 0)
					| (mR.anisotropic ? PRB_MATF_ANISOTROPIC : 0);
		out.node[0] = mSpecularity->emit(e);
		out.node[1] = mTransmission->emit(e);
		out.node[2] = mIOR->emit(e);
		out.f[0]	= mR.rx;
		out.f[1]	= mR.anisotropic ? mR.ry : mR.rx; // getClosure(): IsAnisotropic ? roughnessY : m1
	}
	std::string dumpInformation() const override
	{
		std::stringstream s;
		s << "  <RoughDielectricMaterial>:\n    Specularity:     " << mSpecularity->dumpInformation() << "\n    Transmission:    "
		  << mTransmission->dumpInformation() << "\n    IOR:             " << mIOR->dumpInformation() << "\n    RoughnessX:      " << mR.rx
		  << "\n    RoughnessY:      " << mR.ry << "\n    VNDF:            " << (mVNDF ? "true" : "false") << "\n";
		return s.str();
	}

private:
	std::shared_ptr<FloatSpectralNode> mSpecularity, mTransmission, mIOR;
	bool mHasTrans;
	Roughness mR;
	bool mVNDF;
};
class RoughDielectricMaterialPlugin : public IMaterialPlugin {
public:
	std::shared_ptr<IMaterial> create(const std::string&, const SceneLoadContext& ctx) override
	{
		const Roughness r	= parseRoughness(ctx);
		const bool hasTrans = ctx.parameters().hasParameter("transmission");
		const auto spec		= ctx.lookupSpectralNode("specularity", 1);
		return std::make_shared<RoughDielectricMaterial>(spec, hasTrans ? ctx.lookupSpectralNode("transmission", 1) : spec,
														 ctx.lookupSpectralNode({ "eta", "index", "ior" }, 1.55f), hasTrans, r,
														 ctx.parameters().getBool("vndf", true));
	}
	const std::vector<std::string>& getNames() const override
	{
		static const std::vector<std::string> names({ "roughglass", "roughdielectric", "rough_glass", "rough_dielectric" });
		return names;
	}
	std::string specification(const std::string&) const override
	{
		return "Rough Dielectric BSDF: roughness | roughness_x roughness_y (0), index|eta|ior (1.55), specularity (1), transmission (1), vndf (true)";
	}
};

class PrincipledMaterial : public IMaterial { // principled.cpp:34-631
public:
	std::shared_ptr<FloatSpectralNode> base, ior;
	float f[PRB_PR__COUNT] = {};
	bool vndf = true, thin = false, hasTransmission = false;
	void describe(prb_material& out, NodeEmitter& e) const override
	{
		out.type	= PRB_MAT_PRINCIPLED;
		out.flags	= (vndf ? PRB_MATF_VNDF : 0) | (thin ? PRB_MATF_THIN : 0) | (hasTransmission ? PRB_MATF_HAS_TRANSMISSION : 0);
		out.node[0] = base->emit(e);
		out.node[1] = ior->emit(e);
		for (int i = 0; i < PRB_PR__COUNT; ++i)
			out.f[i] = f[i];
	}
	std::string dumpInformation() const override
	{
		std::stringstream s;
		s << "  <PrincipledMaterial>:\n    BaseColor:            " << base->dumpInformation() << "\n    IOR:                  " << ior->dumpInformation()
		  << "\n    Roughness:            " << f[PRB_PR_ROUGHNESS] << "\n    Metallic:             " << f[PRB_PR_METALLIC] << "\n    Thin:                 "
		  << (thin ? "true" : "false") << "\n    VNDF:                 " << (vndf ? "true" : "false") << "\n";
		return s.str();
	}
};
class PrincipledMaterialPlugin : public IMaterialPlugin {
public:
	std::shared_ptr<IMaterial> create(const std::string&, const SceneLoadContext& ctx) override
	{
		auto m			   = std::make_shared<PrincipledMaterial>();
		const auto& p	   = ctx.parameters();
		m->vndf			   = p.getBool("vndf", true);
		m->thin			   = p.getBool("thin", false);
		m->hasTransmission = p.hasParameter("specular_transmission") || p.hasParameter("spec_trans") || p.hasParameter("diffuse_transmission")
							 || p.hasParameter("diff_trans");
		m->base				 = ctx.lookupSpectralNode({ "base_color", "base" }, 0.8f);
		m->ior				 = ctx.lookupSpectralNode({ "ior", "eta", "index" }, 1.55f);
		m->f[PRB_PR_DIFF_TRANS]		 = constScalar(ctx.lookupScalarNode({ "diffuse_transmission", "diff_trans" }, 0.0f), "diffuse_transmission");
		m->f[PRB_PR_SPEC_TRANS]		 = constScalar(ctx.lookupScalarNode({ "specular_transmission", "spec_trans" }, 0.0f), "specular_transmission");
		m->f[PRB_PR_SPEC_TINT]		 = constScalar(ctx.lookupScalarNode("specular_tint", 0.0f), "specular_tint");
		m->f[PRB_PR_ROUGHNESS]		 = constScalar(ctx.lookupScalarNode("roughness", 0.5f), "roughness");
		m->f[PRB_PR_ANISOTROPIC]	 = constScalar(ctx.lookupScalarNode("anisotropic", 0.0f), "anisotropic");
		m->f[PRB_PR_FLATNESS]		 = constScalar(ctx.lookupScalarNode({ "flatness", "subsurface" }, 0.0f), "flatness");
		m->f[PRB_PR_METALLIC]		 = constScalar(ctx.lookupScalarNode("metallic", 0.0f), "metallic");
		m->f[PRB_PR_SHEEN]			 = constScalar(ctx.lookupScalarNode("sheen", 0.0f), "sheen");
		m->f[PRB_PR_SHEEN_TINT]		 = constScalar(ctx.lookupScalarNode("sheen_tint", 0.0f), "sheen_tint");
		m->f[PRB_PR_CLEARCOAT]		 = constScalar(ctx.lookupScalarNode("clearcoat", 0.0f), "clearcoat");
		m->f[PRB_PR_CLEARCOAT_GLOSS] = constScalar(ctx.lookupScalarNode("clearcoat_gloss", 0.0f), "clearcoat_gloss");
		return m;
	}
	const std::vector<std::string>& getNames() const override
	{
		static const std::vector<std::string> names({ "principled" });
		return names;
	}
	std::string specification(const std::string&) const override
	{
		return "Principled BSDF: base_color|base (0.8), ior|eta|index (1.55), roughness (0.5), anisotropic, diffuse_transmission, specular_transmission, "
			   "specular_tint, flatness|subsurface, metallic, sheen, sheen_tint, clearcoat, clearcoat_gloss (0), vndf (true), thin (false)";
	}
};
class MirrorMaterial : public IMaterial { // mirror.cpp:14-65
public:
	explicit MirrorMaterial(const std::shared_ptr<FloatSpectralNode>& spec)
		: mSpecularity(spec)
	{
	}
	bool hasOnlyDeltaDistribution() const override { return true; }
	void describe(prb_material& out, NodeEmitter& e) const override
	{
		out.type	= PRB_MAT_MIRROR;
		out.flags	= PRB_MATF_ONLY_DELTA;
		out.node[0] = mSpecularity->emit(e);
	}
	std::string dumpInformation() const override { return "  <MirrorMaterial>:\n    Specularity: " + mSpecularity->dumpInformation() + "\n"; }

private:
	std::shared_ptr<FloatSpectralNode> mSpecularity;
};
class MirrorMaterialPlugin : public IMaterialPlugin {
public:
	std::shared_ptr<IMaterial> create(const std::string&, const SceneLoadContext& ctx) override
	{
		return std::make_shared<MirrorMaterial>(ctx.lookupSpectralNode("specularity", 1));
	}
	const std::vector<std::string>& getNames() const override
	{
		static const std::vector<std::string> names({ "mirror", "reflection" });
		return names;
	}
	std::string specification(const std::string&) const override { return "Delta Mirror BSDF: specularity (spectral, 1)"; }
};

class OrenNayarMaterial : public IMaterial { // orennayar.cpp:16-86
public:
	OrenNayarMaterial(const std::shared_ptr<FloatSpectralNode>& alb, float roughness)
		: mAlbedo(alb)
		, mRoughness(roughness)
	{
	}
	void describe(prb_material& out, NodeEmitter& e) const override
	{
		out.type	= PRB_MAT_ORENNAYAR;
		out.flags	= 0;
		out.node[0] = mAlbedo->emit(e);
		out.f[0]	= mRoughness;
	}
	std::string dumpInformation() const override
	{
		std::stringstream s;
		s << "  <OrenNayarMaterial>:\n    Albedo: " << mAlbedo->dumpInformation() << "\n    Roughness: " << mRoughness << "\n";
		return s.str();
	}

private:
	std::shared_ptr<FloatSpectralNode> mAlbedo;
	float mRoughness;
};
class OrenNayarMaterialPlugin : public IMaterialPlugin {
public:
	std::shared_ptr<IMaterial> create(const std::string&, const SceneLoadContext& ctx) override
	{
		return std::make_shared<OrenNayarMaterial>(ctx.lookupSpectralNode("albedo", 1), constScalar(ctx.lookupScalarNode("roughness", 0.5f), "roughness"));
	}
	const std::vector<std::string>& getNames() const override
	{
		static const std::vector<std::string> names({ "orennayar", "oren", "rough" });
		return names;
	}
	std::string specification(const std::string&) const override { return "OrenNayar BSDF: albedo (spectral, 1), roughness (scalar, 0.5)"; }
};
class CombineMaterial : public IMaterial { // blend.cpp:20-148 (BlendMaterial<Delta>) and add.cpp:20-122 (AddMaterial<Delta>)
public:
	CombineMaterial(bool add, const std::shared_ptr<IMaterial>& m0, const std::shared_ptr<IMaterial>& m1, float factor)
		: mAdd(add)
		, mMaterials{ m0, m1 }
		, mFactor(factor)
	{
	}
	bool hasOnlyDeltaDistribution() const override { return mMaterials[0]->hasOnlyDeltaDistribution() && mMaterials[1]->hasOnlyDeltaDistribution(); }
	// levels of combinations below and including this one
	int depth() const
	{
		int d = 0;
		for (const auto& m : mMaterials)
			if (const auto c = std::dynamic_pointer_cast<CombineMaterial>(m))
				d = std::max(d, c->depth());
		return d + 1;
	}
	void describe(prb_material& out, NodeEmitter&) const override
	{
		out.type	= mAdd ? PRB_MAT_ADD : PRB_MAT_BLEND;
		out.flags	= (mMaterials[0]->hasOnlyDeltaDistribution() ? PRB_MATF_CHILD0_DELTA : 0) | (mMaterials[1]->hasOnlyDeltaDistribution() ? PRB_MATF_CHILD1_DELTA : 0)
					| (hasOnlyDeltaDistribution() ? PRB_MATF_ONLY_DELTA : 0);
		out.node[0] = mMaterials[0]->id();
		out.node[1] = mMaterials[1]->id();
		out.f[0]	= mFactor;
	}
	std::string dumpInformation() const override
	{
		return std::string(mAdd ? "  <AddMaterial>:\n" : "  <BlendMaterial>:\n") + "    [0]: " + mMaterials[0]->dumpInformation() + "    [1]: " + mMaterials[1]->dumpInformation();
	}

private:
	bool mAdd;
	std::shared_ptr<IMaterial> mMaterials[2];
	float mFactor;
};
class CombineMaterialPlugin : public IMaterialPlugin {
public:
	explicit CombineMaterialPlugin(bool add)
		: mAdd(add)
	{
	}
	std::shared_ptr<IMaterial> create(const std::string&, const SceneLoadContext& ctx) override
	{
		const ParameterGroup& params = ctx.parameters();
		const uint32 id1 = ctx.lookupMaterialID(params.getParameter("material1")), id2 = ctx.lookupMaterialID(params.getParameter("material2"));
		const auto& db	 = ctx.environment()->sceneDatabase()->Materials;
		const auto mat1 = id1 != PR_INVALID_ID ? db.getSafe(id1) : nullptr, mat2 = id2 != PR_INVALID_ID ? db.getSafe(id2) : nullptr;
		if (!mat1 || !mat2) {
			PR_LOG(L_ERROR) << "Valid material1 or material2 parameters for blend material missing" << std::endl;
			return nullptr;
		}
		const float factor = mAdd ? 0.5f : constScalar(ctx.lookupScalarNode("factor", 0.5f), "factor");
		auto mat		   = std::make_shared<CombineMaterial>(mAdd, mat1, mat2, factor);
		if (mat->depth() > 3) { // COMBINE_MAX_DEPTH of csrc/dev_shade.cuh: the device unrolls the nesting at compile time
			PR_LOG(L_ERROR) << "blend / add materials nested more than 3 levels deep are not supported on the device path" << std::endl;
			return nullptr;
		}
		return mat;
	}
	const std::vector<std::string>& getNames() const override
	{
		static const std::vector<std::string> blend({ "blend", "mix" });
		static const std::vector<std::string> add({ "add" });
		return mAdd ? add : blend;
	}
	std::string specification(const std::string&) const override
	{
		return mAdd ? "Add BSDF: material1, material2 (material references)" : "Blend BSDF: material1, material2 (material references), factor (scalar, 0.5)";
	}

private:
	bool mAdd;
};
} // namespace

void registerMaterialPlugins(std::vector<std::shared_ptr<IPlugin>>& out)
{
	out.push_back(std::make_shared<LambertMaterialPlugin>());
	out.push_back(std::make_shared<DielectricMaterialPlugin>());
	out.push_back(std::make_shared<ConductorMaterialPlugin>());
	out.push_back(std::make_shared<RoughConductorMaterialPlugin>());
	out.push_back(std::make_shared<RoughDielectricMaterialPlugin>());
	out.push_back(std::make_shared<PrincipledMaterialPlugin>());
	out.push_back(std::make_shared<MirrorMaterialPlugin>());
	out.push_back(std::make_shared<OrenNayarMaterialPlugin>());
	out.push_back(std::make_shared<CombineMaterialPlugin>(false));
	out.push_back(std::make_shared<CombineMaterialPlugin>(true));
}
} // namespace PR
