#!/usr/bin/env python3
"""transpile.py -- turn the reference's GLSL fragment shaders into C++ class bodies.

*** TEST INFRASTRUCTURE (oracle/), NOT PRODUCT.

Reads  <reference>/part {3,4,5} .../source code/shaders/fshader.fsh  (and P5's pass3.fsh)  WHERE THEY LIE (read-only) and
writes  oracle/_ref/shader_<variant>.inc  (git-ignored build output, never committed): the shader's
own text with only the mechanical edits GLSL -> C++ needs.  ref_shader_host.cpp #includes each .inc
inside `struct Shader_<variant> { ... }` and runs main() per pixel against oracle/ref_shader/glsl_emul.h.

The edits (each one asserted to apply, so a changed reference fails loudly):
  1. drop the BOM and the #version line;
  2. give every unsuffixed floating literal an `f` suffix (GLSL literals are 32-bit floats);
  3. swizzles .xyz .rgb .xy .rg become accessor calls;
  4. `inout T x` -> `T& x`; the `in` parameter qualifier is dropped;
  5. `uniform T name;` -> `T name = U.name;`, `in vec3 pix;` -> `vec3 pix = pix_in;`, `out vec4 fragColor;`
     -> `vec4 fragColor;` (members initialised in declaration order, so the shader's global
     initialisers -- the RNG seeds -- see the uniforms, as in GLSL);
  6. `const uint V[...]` (the Sobol table) becomes a static constexpr member; `void main()` -> `void shader_main()`;
  7. the two statements that pass two rand() calls as arguments of one call are split into sequenced
     statements: GLSL evaluates arguments left to right, C++ leaves the order unspecified;
  8. variant p5sobol: main() calls pathTracing instead of pathTracingImportanceSampling (the call the
     reference keeps commented out next to it, P5/fsh:937-938);
  9. `maxBounce` literals of main() are replaced by the harness's max_bounce (default = the literal).
"""
import os
import re
import sys

PARTS = {
    "p3": "part 3 -- OpenGL Raytracing",
    "p4": "part 4 -- Disney Principle BRDF",
    "p5": "part 5 -- Importance Sampling & Low Discrepancy Sequence",
}
VARIANTS = {"p3": "p3", "p4": "p4", "p5sobol": "p5", "p5is": "p5"}

FLOAT_LIT = re.compile(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][+-]?\d+)?|\d+[eE][+-]?\d+)(?![\w.])")


def sub_once(text, old, new, what, count=1):
    n = text.count(old)
    if n != count:
        raise SystemExit(f"transpile: expected {count} x {what!r}, found {n}")
    return text.replace(old, new)


def transpile(src, variant):
    t = src.lstrip("﻿")
    t, n = re.subn(r"^#version[^\n]*\n", "\n", t, count=1, flags=re.M)
    assert n == 1, "#version"
    # 7 (before literals are touched)
    t = sub_once(
        t,
        "vec2 AA = vec2((rand()-0.5)/float(width), (rand()-0.5)/float(height));",
        "float AA_x_ = (rand()-0.5)/float(width);\n    float AA_y_ = (rand()-0.5)/float(height);\n    vec2 AA = vec2(AA_x_, AA_y_);",
        "AA jitter",
    )
    if VARIANTS[variant] == "p5":
        t = sub_once(
            t,
            "hdrTestRay.direction = SampleHdr(rand(), rand());",
            "float xi_a_ = rand();\n        float xi_b_ = rand();\n        hdrTestRay.direction = SampleHdr(xi_a_, xi_b_);",
            "SampleHdr(rand(), rand())",
        )
    leftovers = [ln for ln in t.splitlines() if ln.count("rand()") > 1 and not ln.strip().startswith("//")]
    if leftovers:
        raise SystemExit(f"transpile: unsequenced rand() pair left: {leftovers}")
    # 8, 9
    if variant == "p5sobol":
        t = sub_once(t, "vec3 Li = pathTracingImportanceSampling(firstHit, maxBounce);",
                     "vec3 Li = pathTracing(firstHit, maxBounce);", "IS call")
    if VARIANTS[variant] == "p5":
        t = sub_once(t, "int maxBounce = 2;", "int maxBounce = max_bounce_in;", "maxBounce literal")
    else:
        t, n = re.subn(r"vec3 Li = pathTracing\(firstHit, \d+\);", "vec3 Li = pathTracing(firstHit, max_bounce_in);", t)
        assert n == 1, "pathTracing(firstHit, <literal>)"  # the literal is 2 in P3, 4 in P4
    # 2
    t = FLOAT_LIT.sub(lambda m: m.group(1) + "f", t)
    # 3
    t = re.sub(r"\.(xyz|rgb|xy|rg)\b(?!\s*\()", r".\1()", t)
    # 4
    t, n = re.subn(r"\binout\s+(\w+)\s+", r"\1& ", t)
    assert n >= 1, "inout"
    t = re.sub(r"([(,]\s*)in\s+(?=\w+\s+\w+\s*[,)])", r"\1", t)
    # 5
    t, n = re.subn(r"^uniform\s+(\w+)\s+(\w+)\s*;", r"\1 \2 = U.\2;", t, flags=re.M)
    assert n >= 10, "uniforms"
    t = sub_once(t, "in vec3 pix;", "vec3 pix = pix_in;", "in vec3 pix")
    t = sub_once(t, "out vec4 fragColor;", "vec4 fragColor;", "out vec4 fragColor")
    # 6
    if VARIANTS[variant] == "p5":
        t, n = re.subn(r"^const uint V\[8\*32\] = \{", "static constexpr uint V[8*32] = {", t, flags=re.M)
        assert n == 2, "Sobol table (one live, one inside a comment)"
    t = sub_once(t, "void main() {", "void shader_main() {", "main")
    if re.search(r"^\s*(uniform|in|out|inout)\s", t, flags=re.M):
        raise SystemExit("transpile: a storage qualifier survived")
    return t


def transpile_pass3(src):
    """shaders/pass3.fsh (tone mapping + gamma, identical in parts 3, 4 and 5): edits 1-6 only"""
    t = src.lstrip("\ufeff")
    t, n = re.subn(r"^#version[^\n]*\n", "\n", t, count=1, flags=re.M)
    assert n == 1, "#version"
    t = FLOAT_LIT.sub(lambda m: m.group(1) + "f", t)
    t = re.sub(r"\.(xyz|rgb|xy|rg)\b(?!\s*\()", r".\1()", t)
    t = re.sub(r"([(,]\s*)in\s+(?=\w+\s+\w+\s*[,)])", r"\1", t)
    t, n = re.subn(r"^uniform\s+(\w+)\s+(\w+)\s*;", r"\1 \2 = U.\2;", t, flags=re.M)
    assert n == 7, "texPass uniforms"
    t = sub_once(t, "in vec3 pix;", "vec3 pix = pix_in;", "in vec3 pix")
    t = sub_once(t, "out vec4 fragColor;", "vec4 fragColor;", "out vec4 fragColor")
    t = sub_once(t, "void main() {", "void shader_main() {", "main")
    if re.search(r"^\s*(uniform|in|out|inout)\s", t, flags=re.M):
        raise SystemExit("transpile: a storage qualifier survived")
    return t


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    out_dir = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "_ref")
    os.makedirs(out_dir, exist_ok=True)
    for variant, part in VARIANTS.items():
        path = os.path.join(ref, PARTS[part], "source code", "shaders", "fshader.fsh")
        with open(path, encoding="utf-8") as f:
            src = f.read()
        body = transpile(src, variant)
        dst = os.path.join(out_dir, f"shader_{variant}.inc")
        tmp = dst + f".tmp{os.getpid()}"
        with open(tmp, "w", encoding="utf-8") as f:
            f.write(f"// GENERATED by oracle/ref_shader/transpile.py from {path}\n// build output -- do not commit\n")
            f.write(body)
        os.replace(tmp, dst)
        print(f"transpile: {variant}: {len(src.splitlines())} lines -> {dst}")
    path = os.path.join(ref, PARTS["p5"], "source code", "shaders", "pass3.fsh")
    with open(path, encoding="utf-8") as f:
        body = transpile_pass3(f.read())
    dst = os.path.join(out_dir, "shader_pass3.inc")
    with open(dst, "w", encoding="utf-8") as f:
        f.write(f"// GENERATED by oracle/ref_shader/transpile.py from {path}\n// build output -- do not commit\n")
        f.write(body)
    print(f"transpile: pass3 -> {dst}")


if __name__ == "__main__":
    main()
