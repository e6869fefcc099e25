"""The reference's plugin ABI exercised end to end: an EXTERNAL shared object `pr_pl_testmat.so` exporting
`_pr_exports` (src/loader/plugin/Plugin.h:26-66, API version 1) is found through PR_PLUGIN_PATH
(PluginManager.cpp:7,14-66: directories searched for (lib)?pr_pl_<name>.so), loaded with dlopen, version-checked
(PluginManager.cpp:186-214), routed by IPlugin::type() to the material manager and used by a scene; a twin with the wrong
API version is rejected and its type stays unknown."""
import os
import subprocess

import numpy as np
import pytest

import pearray_b200 as prb
from conftest import ROOT

SCENE = """
(scene :name 'plug' :render_width 32 :render_height 32 :camera 'Camera'
 (integrator :type 'direct' :max_ray_depth 4)
 (sampler :slot 'aa' :type 'mjitt' :sample_count 16)
 (filter :slot 'pixel' :type 'block' :radius 0)
 (camera :name 'Camera' :type 'standard' :width 1 :height 1 :local_direction [0,0,-1] :local_up [0,1,0] :local_right [1,0,0]
   :near 0.01 :far 100 :transform [1,0,0,0, 0,1,0,0, 0,0,1,3, 0,0,0,1])
 (material :name 'm' :type '%s' :albedo %s)
 (entity :name 'ball' :type 'sphere' :radius 1 :material 'm')
 (light :type 'env' :radiance 1)
)
"""


def build_plugin(out_dir, name, api_version=None):
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libpr_pl_%s.so" % name)
    cmd = ["g++", "-std=c++17", "-O1", "-fPIC", "-shared", os.path.join(ROOT, "tests", "plugins", "pr_pl_testmat.cpp"), "-I" + os.path.join(ROOT, "pearray_b200", "host"),
           "-I" + os.path.join(ROOT, "include"), '-DTEST_NAME="%s"' % name, "-o", so]
    if api_version is not None:
        cmd.append("-DTEST_API_VERSION=%d" % api_version)
    subprocess.check_call(cmd)
    return so


@pytest.fixture(scope="module")
def plugin_dir(tmp_path_factory):
    d = tmp_path_factory.mktemp("plugins")
    build_plugin(str(d), "testmat")
    build_plugin(str(d), "futuremat", api_version=2)
    build_plugin(str(d), "ancientmat", api_version=0)
    (d / "pr_pl_notaplugin.txt").write_text("ignored: not a shared object")
    (d / "libunrelated.so").write_bytes(b"")  # name does not match (lib)?pr_pl_<name>: never opened
    return str(d)


@pytest.fixture()
def plugin_path(plugin_dir, monkeypatch):
    monkeypatch.setenv("PR_PLUGIN_PATH", "/nonexistent/dir:" + plugin_dir)
    return plugin_dir


def test_external_plugin_is_loaded_and_used(plugin_path):
    scene = prb.Scene.from_string(SCENE % ("testmat", "0.8"))
    reg = {l.split(":")[0]: l.split(":")[1].split() for l in scene.plugins().strip().splitlines()}
    assert "testmat" in reg["material"]
    d = scene.desc.contents
    assert d.n_materials == 1 and d.materials[0].type == 0  # described as a Lambert surface ...
    node = d.nodes[d.materials[0].node[0]]
    assert node.type == 0 and abs(node.p[0] - 0.4) < 1e-7  # ... whose constant albedo is HALF the parameter: the plugin's doing


def test_wrong_api_versions_are_rejected(plugin_path):
    """PluginManager.cpp:186-214: older and newer API versions are refused, the object is unloaded, its types stay unknown"""
    scene = prb.Scene.from_string(SCENE % ("diffuse", "0.8"))
    reg = {l.split(":")[0]: l.split(":")[1].split() for l in scene.plugins().strip().splitlines()}
    assert "futuremat" not in reg["material"] and "ancientmat" not in reg["material"]
    # a scene that asks for the rejected type loads like any scene with an unknown material type: the loader logs an error and
    # skips the object (SceneLoader.cpp:215-226); the entity is left without a material
    bad = prb.Scene.from_string(SCENE % ("futuremat", "0.8"))
    assert bad.desc.contents.n_materials == 0


def test_without_plugin_path_the_type_is_unknown(monkeypatch):
    monkeypatch.delenv("PR_PLUGIN_PATH", raising=False)
    scene = prb.Scene.from_string(SCENE % ("testmat", "0.8"))
    assert scene.desc.contents.n_materials == 0


@pytest.mark.gpu
def test_render_with_external_plugin_matches_builtin_equivalent(plugin_path):
    """the plugged-in material renders exactly like the embedded 'diffuse' plugin with the albedo it describes"""
    from oracle_binding import OracleScene
    films = []
    for typ, alb in (("testmat", "0.8"), ("diffuse", "0.4")):
        scene = prb.Scene.from_string(SCENE % (typ, alb))
        ctx = prb.Context(0)
        ctx.upload_scene(scene)
        ctx.upload_rng(scene.rng_map())
        ctx.render_tiles([(0, 0, 32, 32)], 0, 8)
        xyz, _ = ctx.film()
        films.append(xyz)
        if typ == "testmat":
            ref = OracleScene(scene).render([(0, 0, 32, 32)], 0, 8)
            assert np.array_equal(xyz.view(np.uint32), ref["filtered"].view(np.uint32))
    assert np.array_equal(films[0].view(np.uint32), films[1].view(np.uint32))
    assert films[0][16, 16, 1] > 0
