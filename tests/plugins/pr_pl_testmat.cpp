// A minimal EXTERNAL material plugin in the reference's plugin ABI (src/loader/plugin/Plugin.h:26-66): the shared object
// exports one extern "C" data symbol `_pr_exports` {APIVersion, FileName, ClassName, PluginName, PluginVersion, InitFunction},
// InitFunction returns an IPlugin whose type() routes it to the material manager (Environment.cpp:203-241).
// Built by tests/test_plugin_abi.py with g++ against pearray_b200/host/prh.h; -DTEST_API_VERSION=<n> builds the
// wrong-version twin that PluginManager::tryLoad must reject (PluginManager.cpp:186-214).
#include "prh.h"

#ifndef TEST_API_VERSION
#define TEST_API_VERSION PR_PLUGIN_API_VERSION
#endif
#ifndef TEST_NAME
#define TEST_NAME "testmat"
#endif

namespace {
using namespace PR;
// "testmat": a grey Lambert surface whose albedo is HALF the `albedo` parameter -- something no embedded plugin does, so a
// render that shows it proves the external factory was used
class TestMaterial : public IMaterial {
public:
	explicit TestMaterial(const std::shared_ptr<FloatSpectralNode>& albedo)
		: mAlbedo(albedo)
	{
	}
	void describe(prb_material& out, NodeEmitter& e) const override
	{
		out.type	= PRB_MAT_DIFFUSE;
		out.flags	= PRB_MATF_TWO_SIDED;
		out.node[0] = mAlbedo->emit(e);
	}
	std::string dumpInformation() const override { return "  <TestMaterial>\n"; }

private:
	std::shared_ptr<FloatSpectralNode> mAlbedo;
};
class TestMaterialPlugin : public IMaterialPlugin {
public:
	std::shared_ptr<IMaterial> create(const std::string&, const SceneLoadContext& ctx) override
	{
		const float a = ctx.parameters().getNumber("albedo", 1.0f);
		return std::make_shared<TestMaterial>(makeConstSpectralNode(0.5f * a));
	}
	const std::vector<std::string>& getNames() const override
	{
		static const std::vector<std::string> names({ TEST_NAME });
		return names;
	}
	std::string specification(const std::string&) const override { return "Test material: albedo (number, 1)"; }
};
} // namespace

extern "C" {
static PR::IPlugin* testmat_init() { return new TestMaterialPlugin(); }
__attribute__((visibility("default"))) PR::PluginInterface _pr_exports = { TEST_API_VERSION, "pr_pl_" TEST_NAME, "TestMaterialPlugin", TEST_NAME, "1.0", testmat_init };
}
