"""Light path expressions: the host's LPE compiler (pearray_b200/host/lpe.cpp, an own NFA -> DFA construction)
against the reference's known-answer cases, src/tests/lpe.cpp:9-139, ported one to one (expression, path, expected match), plus
the grammar corners of src/core/path/LPE_Parser.cpp."""
import numpy as np
import pytest

import pearray_b200 as prb

# ScatteringType / ScatteringEvent, src/core/path/LightPathToken.h:6-20
CAMERA, EMISSIVE, REFRACTION, REFLECTION, BACKGROUND = range(5)
DIFFUSE, SPECULAR, NONE = range(3)
C = (CAMERA, NONE)             # LightPathToken::Camera()
B = (BACKGROUND, NONE)         # LightPathToken::Background()
E = (EMISSIVE, DIFFUSE)        # the tests build emissive tokens with the Diffuse event
RD, RS = (REFLECTION, DIFFUSE), (REFLECTION, SPECULAR)
TS, TD = (REFRACTION, SPECULAR), (REFRACTION, DIFFUSE)


def match(expr, path):
    t = np.ascontiguousarray(np.array(path, np.int32).reshape(-1, 2))
    return prb.host_lib().prh_lpe_match(expr.encode(), t.ctypes.data, len(path))


# (expression, path, expected) -- src/tests/lpe.cpp
REFERENCE_CASES = [
    ("CD*L", [C, RD, E], True),                                   # :9-32
    ("CD*L", [C, E], True),
    ("CD*L", [C, RS, E], False),
    ("C(DS)+D?E", [C, RD, TS, E], True),                          # :38-73
    ("C(DS)+D?E", [C, RD, TS, RD, TS, RD, E], True),
    ("C(DS)+D?E", [C, RS, E], False),
    ("C(DS)+D?E", [C, RS, RD, E], False),
    ("C[DS]+D?B", [C, RD, TS, B], True),                          # :74-98
    ("C[DS]+D?B", [C, RD, TS, RD, RD, B], True),
    ("C[DS]+D?B", [C, (EMISSIVE, SPECULAR), B], False),
    ("C(DS+)+.*L", [C, RD, TS, B], True),                         # :99-139
    ("C(DS+)+.*L", [C, RD, TS, RD, RD, B], True),
    ("C(DS+)+.*L", [C, RD, TS, TS, RD, TS, RD, RD, B], True),
    ("C(DS+)+.*L", [C, (EMISSIVE, SPECULAR), B], False),
]


@pytest.mark.parametrize("expr,path,expected", REFERENCE_CASES)
def test_reference_known_answers(expr, path, expected):
    assert match(expr, path) == (1 if expected else 0)


def test_invalid_expressions():
    assert match("RD*L", [C]) == -1          # lpe.cpp:33-37: must start at the camera
    assert match("", [C]) == -1
    assert match("C[^S]+S?B", [C]) == -1     # negated unions are rejected, LPE_Parser.cpp:139-144
    assert match("C(DS", [C]) == -1
    assert match("CD{3,2}E", [C]) == -1      # maximum less than minimum, LPE_Parser.cpp:269-273
    assert match("C D*E", [C]) == -1         # the reference parser does not skip blanks
    assert match("CX", [C]) == -1


def test_token_classes_and_quantifiers():
    # Token::match (LPE_RegState.h:39-78): L = emissive or background, '.' = any scattering, R / T by type, <T,E> explicit
    assert match("C", [C]) == 1 and match("C", [C, E]) == 0
    assert match("CL", [C, E]) == 1 and match("CL", [C, B]) == 1 and match("CE", [C, B]) == 0 and match("CB", [C, B]) == 1
    assert match("C.E", [C, TD, E]) == 1 and match("C.E", [C, RS, E]) == 1 and match("C.E", [C, E, E]) == 0
    assert match("CRE", [C, RS, E]) == 1 and match("CRE", [C, TS, E]) == 0 and match("CTE", [C, TD, E]) == 1
    assert match("C<R,D>E", [C, RD, E]) == 1 and match("C<R,D>E", [C, RS, E]) == 0 and match("C<RD>E", [C, RD, E]) == 1
    assert match("C<T.>+E", [C, TS, TD, E]) == 1 and match("C<..>E", [C, RS, E]) == 1
    assert match('C<R,D,"floor">E', [C, RD, E]) == 0   # labelled tokens need labelled paths; `direct` never labels
    # {n}, {n,m}; {0} is repeatLast(0, 0) = '*' as the reference writes it (LPE_Parser.cpp:255-275, LPE_RegExpr.cpp:129-139)
    assert match("CD{2}E", [C, RD, RD, E]) == 1 and match("CD{2}E", [C, RD, E]) == 0 and match("CD{2}E", [C, RD, RD, RD, E]) == 0
    assert match("CD{1,3}E", [C, RD, E]) == 1 and match("CD{1,3}E", [C, RD, RD, RD, E]) == 1 and match("CD{1,3}E", [C, RD, RD, RD, RD, E]) == 0
    assert match("CD{0}E", [C, E]) == 1 and match("CD{0}E", [C, RD, RD, E]) == 1
    assert match("C(D|S)E", [C]) == -1       # '|' is not part of the grammar: unions are written [DS]
    assert match("C[D(SS)]*E", [C, RD, TS, RS, RD, E]) == 1 and match("C[D(SS)]*E", [C, RD, TS, E]) == 0


def test_paths_of_the_direct_integrator():
    """the token sequences direct.cpp produces: C, then one (type, event) per scattering, closed by E (emitter hit or NEE
    towards an area light) or B (miss or NEE towards an infinite light)"""
    direct_lighting = "C[DS]L"               # one bounce
    assert match(direct_lighting, [C, RD, E]) == 1 and match(direct_lighting, [C, RD, RD, E]) == 0
    caustics = "CDS+L"
    assert match(caustics, [C, RD, TS, TS, B]) == 1 and match(caustics, [C, RD, B]) == 0
    emission_only = "CE"
    assert match(emission_only, [C, E]) == 1 and match(emission_only, [C, RD, E]) == 0
