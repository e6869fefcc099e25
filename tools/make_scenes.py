#!/usr/bin/env python3
"""Generate the benchmark / parity scene fixtures under scenes/ from the reference's example DATA
(/root/reference/examples/*.prc), applying the overrides of SURVEY 8(d):

  c1_sphere.prc          examples/sphere.prc, integrator -> 'direct', aa sampler sobol 64 spp
  c2_cornellbox.prc      examples/cornellbox.prc (+ inlined cornellbox_mesh.prc.inc), integrator -> 'direct' depth 6
  c3_cornellbox_glassy.prc  examples/cornellbox_glassy.prc (+ inlined mesh include), as shipped
  c4_boltsandgears.prc   examples/boltsandgears.prc (+ inlined meshes), as shipped
  c4b_complex_env.prc    examples/complex.prc with the 'sky' and 'sun' lights replaced by one constant D65 environment light
                         (interim form until the sky/sun rows of 8(f) land), everything else as shipped
  c4c_complex.prc        examples/complex.prc (+ inlined meshes), as shipped: 'sky' and 'sun' infinite lights
  c0_evaluation.prc      examples/evaluation/scene.prc with the eight (embed :loader 'obj') meshes inlined as (mesh ...) blocks;
                         this is the scene of the reference's golden image examples/evaluation/cbox.exr

Scene files are input data, not code; they are re-serialised (comments dropped, includes inlined, whitespace
normalised) so that the GPU box, which has no /root/reference, can load them.  Run in the build container only.
"""
import os, re, sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EX = os.path.join(REF, "examples")


def strip_comments(src):
    out = []
    for line in src.splitlines():
        res, q = [], None
        for ch in line:
            if q:
                res.append(ch)
                if ch == q:
                    q = None
            elif ch in "'\"":
                q = ch
                res.append(ch)
            elif ch == ";":
                break
            else:
                res.append(ch)
        s = "".join(res).rstrip()
        if s.strip():
            out.append(s)
    return "\n".join(out)


def inline_includes(src, base):
    def repl(m):
        path = os.path.join(base, m.group(2))
        inner = strip_comments(open(path).read())
        return inline_includes(inner, os.path.dirname(path))
    return re.sub(r"\(include\s+(['\"])([^'\"]+)\1\s*\)", repl, src)


def replace_block(src, name, new):
    # replace the first top-level "(name ...)" block (balanced parentheses)
    i = re.search(r"\(\s*" + name + r"\b", src)
    if not i:
        return src
    depth, j = 0, i.start()
    while j < len(src):
        if src[j] == "(":
            depth += 1
        elif src[j] == ")":
            depth -= 1
            if depth == 0:
                break
        j += 1
    return src[:i.start()] + new + src[j + 1:]


def replace_all_blocks(src, name, new_first):
    """replace the first "(name ...)" block by new_first and drop the others"""
    out = replace_block(src, name, "\0")
    while re.search(r"\(\s*" + name + r"\b", out):
        out = replace_block(out, name, "")
    return out.replace("\0", new_first)


def obj_to_mesh(path, name):
    """Wavefront OBJ -> (mesh ...) block: polygons fan-triangulated, corners merged by (v, vn) -- the same construction as
    loadWavefront in pearray_b200/host/loader.cpp"""
    P, N, polys = [], [], []
    for line in open(path):
        t = line.split()
        if not t or t[0].startswith("#"):
            continue
        if t[0] == "v":
            P.append(tuple(t[1:4]))
        elif t[0] == "vn":
            N.append(tuple(t[1:4]))
        elif t[0] == "f":
            poly = []
            for tok in t[1:]:
                f = (tok.split("/") + ["", ""])[:3]
                v = int(f[0]); n = int(f[2]) if f[2] else 0
                poly.append((v - 1 if v > 0 else len(P) + v, (n - 1 if n > 0 else len(N) + n) if N else -1))
            polys.append(poly)
    merged, verts, faces = {}, [], []
    for poly in polys:
        for k in range(1, len(poly) - 1):
            tri = []
            for c in (poly[0], poly[k], poly[k + 1]):
                c = (max(0, c[0]), max(0, c[1]) if N else -1)
                if c not in merged:
                    merged[c] = len(verts)
                    verts.append(c)
                tri.append(merged[c])
            faces.append(tri)
    out = ["(mesh :name '%s'" % name, "(attribute :type 'p'"]
    out += ["[%s]" % ", ".join(P[v]) for v, _ in verts]
    out.append(")")
    if N:
        out.append("(attribute :type 'n'")
        out += ["[%s]" % ", ".join(N[n]) for _, n in verts]
        out.append(")")
    out.append("(faces")
    out += ["[%d, %d, %d]" % tuple(f) for f in faces]
    out.append("))")
    return "\n".join(out)


def inline_embeds(src, base):
    def repl(m):
        body = m.group(0)
        f = re.search(r":file\s+'([^']+)'", body).group(1)
        n = re.search(r":name\s+'([^']+)'", body).group(1)
        return obj_to_mesh(os.path.join(base, f), n)
    return re.sub(r"\(embed\b[^()]*\)", repl, src)


def normalise(src):
    src = re.sub(r"[ \t]+", " ", src)
    src = re.sub(r"\n\s*", "\n", src)
    return src.strip() + "\n"


def build(name, out, edits=()):
    path = os.path.join(EX, name)
    src = inline_includes(strip_comments(open(path).read()), os.path.dirname(path))
    for fn in edits:
        src = fn(src)
    header = "; generated by tools/make_scenes.py from the reference example '%s' (scene data, overrides per SURVEY 8(d))\n" % name
    open(os.path.join(ROOT, "scenes", out), "w").write(header + normalise(src))
    print("wrote", out, len(src))


os.makedirs(os.path.join(ROOT, "scenes"), exist_ok=True)
build("sphere.prc", "c1_sphere.prc", [
    lambda s: replace_block(s, "integrator", "(integrator :type 'direct')"),
    lambda s: replace_block(s, "sampler", "(sampler :slot 'aa' :type 'sobol' :sample_count 64)"),
])
build("cornellbox.prc", "c2_cornellbox.prc", [
    lambda s: replace_block(s, "integrator", "(integrator :type 'direct' :max_ray_depth 6)"),
])
build("cornellbox_glassy.prc", "c3_cornellbox_glassy.prc")
build("boltsandgears.prc", "c4_boltsandgears.prc")
build("complex.prc", "c4b_complex_env.prc", [
    lambda s: replace_all_blocks(s, "light", "(light :name 'env' :type 'env' :radiance (illuminant 'D65'))"),
])
build("complex.prc", "c4c_complex.prc")  # as shipped: Hosek-Wilkie 'sky' + 'sun' lights
build("evaluation/scene.prc", "c0_evaluation.prc", [
    lambda s: inline_embeds(s, os.path.join(EX, "evaluation")),
])
