"""Light path expression channels on the oracle: every fragment of the `direct` integrator carries the token
string of its path (reference direct.cpp:67,124-134,197,337-351,387,409; IntegratorUtils.h:19) and goes to every spectral
channel whose expression accepts it (LocalFrameOutputDevice.cpp:100-111).  Known answers that need no second renderer:
an expression that accepts every path reproduces the colour channel BIT FOR BIT (same fragments, same order), and channels
that partition the path space add up to it."""
import numpy as np

import pearray_b200 as prb
from oracle_binding import OracleScene
from scene_strings import LPE_EXPRESSIONS, LPE_ZOO


def render():
    scene = prb.Scene.from_string(LPE_ZOO)
    assert int(scene.desc.contents.n_lpe) == len(LPE_EXPRESSIONS)
    return scene, OracleScene(scene).render([(0, 0, scene.width, scene.height)], 0, 8)


def test_scene_carries_one_automaton_per_expression():
    scene = prb.Scene.from_string(LPE_ZOO)
    d = scene.desc.contents
    assert d.n_lpe == len(LPE_EXPRESSIONS) <= prb.MAX_LPE
    for k in range(d.n_lpe):
        l = d.lpe[k]
        assert 0 < l.n_states <= 255 and l.start_state == 0
        assert l.next_offset + 15 * l.n_states <= d.n_lpe_bytes and l.final_offset + l.n_states <= d.n_lpe_bytes
        nxt = np.ctypeslib.as_array(d.lpe_tables, (d.n_lpe_bytes,))[l.next_offset:l.next_offset + 15 * l.n_states]
        assert np.all((nxt == 0xFF) | (nxt < l.n_states))
        # every expression starts at the camera: the start state only leaves on Camera tokens (any event)
        assert [s for s in range(15) if nxt[s] != 0xFF] == [0, 1, 2]


def test_accept_all_expression_equals_the_colour_channel_bitwise():
    _, r = render()
    k = LPE_EXPRESSIONS.index("C.*L")
    assert np.any(r["film"] > 0)
    assert np.array_equal(r["lpe"][k].view(np.uint32), r["film"].view(np.uint32))


def test_partitions_add_up():
    _, r = render()
    ch = {e: r["lpe"][i].astype(np.float64) for i, e in enumerate(LPE_EXPRESSIONS)}
    film = r["film"].astype(np.float64)
    scale = np.abs(film).max()
    # directly seen light + everything that scattered at least once
    assert np.allclose(ch["CL"] + ch["C.+L"], film, rtol=0, atol=2e-6 * scale)
    # paths that end on the environment + paths that end on an emitter
    assert np.allclose(ch["C.*B"] + ch["C.*E"], film, rtol=0, atol=2e-6 * scale)
    for e in ("CL", "C.+L", "C.*B", "C.*E", "CDL", "C<T,S>+.*L", "C<R,S>[DS]*E"):
        assert ch[e].min() >= 0 and ch[e].max() > 0, e
        assert np.all(ch[e] <= film + 2e-6 * scale), e
    # one diffuse bounce is a subset of "scattered at least once"
    assert np.all(ch["CDL"] <= ch["C.+L"] + 2e-6 * scale)


def test_direct_light_channel_is_where_the_camera_sees_a_light():
    _, r = render()
    direct = r["lpe"][LPE_EXPRESSIONS.index("CL")].sum(axis=2) > 0
    # the environment is seen around the floor plane and the lamp is outside the view: a pixel none of whose primary rays hit
    # anything (sample count 0: the count is incremented per primary hit, pushSPFragment) is lit by 'CL' alone
    never_hit = r["count"] == 0
    assert np.any(never_hit) and np.all(direct[never_hit])
    only_cl = r["lpe"][LPE_EXPRESSIONS.index("C.+L")].sum(axis=2) == 0
    assert np.all(only_cl[never_hit])
