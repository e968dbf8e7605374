#!/usr/bin/env python
"""oracle/build_ref.py -- THE RECIPE for everything under oracle/ (test infrastructure, never product):

  oracle/libezrt_oracle.so          the CPU oracle (oracle/ezrt_oracle.cpp)
  oracle/_ref/libhdrloader_ref.so   the reference's own lib/hdrloader.cpp
  oracle/_ref/libezrt_refshader.so  the reference's own fragment shaders, transpiled to C++ (oracle/ref_shader/)
  oracle/_ref/libezrt_refhost*.so   the reference's own main.cpp (parts 3, 4, 5) against stand-in GL/GLUT/glm headers (oracle/ref_stubs/)

Reference sources are compiled FROM WHERE THEY LIE under /root/reference (never copied); outputs go only into the
git-ignored oracle/_ref/ (they travel to the GPU box with the snapshot).  Where /root/reference is absent the functions
return the prebuilt library if it is there, else None.   usage: python oracle/build_ref.py [--force]
ezrt_b200/build.py delegates its build_oracle / build_reference_* entry points to this file."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INCLUDE = os.path.join(ROOT, "include")
ORACLE_SO = os.path.join(ROOT, "oracle", "libezrt_oracle.so")
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
REF_HDR_SO = os.path.join(REF_DIR, "libhdrloader_ref.so")
REF_SHADER_SO = os.path.join(REF_DIR, "libezrt_refshader.so")
REF_HOST_SO = os.path.join(REF_DIR, "libezrt_refhost.so")
REFERENCE_ROOT = "/root/reference"
REFERENCE_PARTS = ("part 3 -- OpenGL Raytracing", "part 4 -- Disney Principle BRDF",
                   "part 5 -- Importance Sampling & Low Discrepancy Sequence")
REFERENCE_P5 = os.path.join(REFERENCE_ROOT, REFERENCE_PARTS[2], "source code")
# bit-identical fp32 on host and device: no contraction, FMA only where ezrt_math.h spells it
HOST_FLAGS = ["-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-mfma", "-Wall"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources if os.path.exists(s))


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("build failed: " + " ".join(cmd))
    return r.stdout + r.stderr


def build_oracle(force=False):
    src = os.path.join(ROOT, "oracle", "ezrt_oracle.cpp")
    deps = [src] + [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE)]
    if force or _newer(ORACLE_SO, deps):
        _run(["g++"] + HOST_FLAGS + ["-fopenmp", "-Wno-misleading-indentation", "-shared", "-I", INCLUDE, src, "-o", ORACLE_SO])
    return ORACLE_SO


def build_reference_hdrloader(force=False):
    """oracle/_ref: compile the reference's own hdrloader.cpp where it lies (never copied)."""
    src = os.path.join(REFERENCE_P5, "lib", "hdrloader.cpp")
    shim = os.path.join(ROOT, "oracle", "ref_hdrloader_shim.cpp")
    if not os.path.exists(src):
        return REF_HDR_SO if os.path.exists(REF_HDR_SO) else None
    os.makedirs(REF_DIR, exist_ok=True)
    if force or _newer(REF_HDR_SO, [src, shim]):
        # the prefix header reroutes the source's sscanf("%ld" into int) call, which is UB on LP64
        prefix = os.path.join(ROOT, "oracle", "ref_hdrloader_prefix.h")
        obj = os.path.join(REF_DIR, "hdrloader.o")
        _run(["g++", "-O2", "-fPIC", "-w", "-include", prefix, "-I", os.path.join(REFERENCE_P5, "lib"), "-c", src, "-o", obj])
        _run(["g++", "-O2", "-fPIC", "-shared", "-w", "-I", os.path.join(REFERENCE_P5, "lib"), obj, shim, "-o", REF_HDR_SO])
        os.remove(obj)
    return REF_HDR_SO


def build_reference_shaders(force=False):
    """oracle/_ref: the reference's own fragment shaders (P3/P4/P5 shaders/fshader.fsh), transpiled from
    where they lie by oracle/ref_shader/transpile.py and compiled against oracle/ref_shader/glsl_emul.h.
    Test infrastructure: pins the hand-written oracle to the reference's statements (tests/test_ref_shader.py)."""
    rs = os.path.join(ROOT, "oracle", "ref_shader")
    srcs = [os.path.join(REFERENCE_ROOT, part, "source code", "shaders", "fshader.fsh") for part in REFERENCE_PARTS]
    if not all(os.path.exists(s) for s in srcs):
        return REF_SHADER_SO if os.path.exists(REF_SHADER_SO) else None
    os.makedirs(REF_DIR, exist_ok=True)
    deps = srcs + [os.path.join(rs, f) for f in ("transpile.py", "glsl_emul.h", "ref_shader_host.cpp")] + [os.path.join(INCLUDE, "ezrt_math.h"), os.path.join(INCLUDE, "ezrt.h")]
    if force or _newer(REF_SHADER_SO, deps):
        _run([sys.executable, os.path.join(rs, "transpile.py"), REFERENCE_ROOT, REF_DIR])
        tmp = REF_SHADER_SO + ".tmp%d" % os.getpid()
        _run(["g++"] + HOST_FLAGS + ["-fopenmp", "-w", "-shared", "-I", INCLUDE, "-I", rs, "-I", REF_DIR, os.path.join(rs, "ref_shader_host.cpp"), "-o", tmp])
        os.replace(tmp, REF_SHADER_SO)
        for f in os.listdir(REF_DIR):  # the transpiled text is a build intermediate: keep only the binary
            if f.startswith("shader_") and f.endswith(".inc"):
                os.remove(os.path.join(REF_DIR, f))
    return REF_SHADER_SO


def build_reference_host(force=False, part=5):
    """oracle/_ref: the reference's own host code (main.cpp of tutorial part 3, 4 or 5: readObj, buildBVH,
    buildBVHwithSAH, calculateHdrCache (P5), main()'s scene set-up and uploads), compiled from where it lies
    together with its hdrloader.cpp; GL/GLUT/glm come from the stand-ins in oracle/ref_stubs/.  Test
    infrastructure (tests/test_ref_host.py)."""
    src_dir = os.path.join(REFERENCE_ROOT, REFERENCE_PARTS[part - 3], "source code")
    target = REF_HOST_SO if part == 5 else REF_HOST_SO.replace(".so", "_p%d.so" % part)
    main_cpp = os.path.join(src_dir, "main.cpp")
    hdr_cpp = os.path.join(src_dir, "lib", "hdrloader.cpp")
    if not (os.path.exists(main_cpp) and os.path.exists(hdr_cpp)):
        return target if os.path.exists(target) else None
    os.makedirs(REF_DIR, exist_ok=True)
    orc = os.path.join(ROOT, "oracle")
    stubs = os.path.join(orc, "ref_stubs")
    shim, hshim, prefix = (os.path.join(orc, f) for f in ("ref_host_shim.cpp", "ref_hdrloader_shim.cpp", "ref_hdrloader_prefix.h"))
    deps = [main_cpp, hdr_cpp, shim, hshim, prefix, os.path.join(INCLUDE, "ezrt_math.h")]
    for d, _, files in os.walk(stubs):
        deps += [os.path.join(d, f) for f in files]
    if force or _newer(target, deps):
        obj = os.path.join(REF_DIR, "hdrloader_host_p%d.o" % part)
        _run(["g++", "-O2", "-fPIC", "-w", "-include", prefix, "-I", os.path.join(src_dir, "lib"), "-c", hdr_cpp, "-o", obj])
        tmp = target + ".tmp%d" % os.getpid()
        _run(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-mfma", "-w", "-Dmain=ezrt_ref_main", "-DEZRT_REF_PART=%d" % part,
              '-DEZRT_REF_MAIN_CPP="%s"' % main_cpp, "-I", stubs, "-I", INCLUDE, "-I", src_dir, "-I", os.path.join(src_dir, "lib"),
              "-shared", shim, obj, hshim, "-o", tmp])
        os.replace(tmp, target)
        os.remove(obj)
    return target


def build_all(force=False):
    out = [build_oracle(force), build_reference_hdrloader(force), build_reference_shaders(force)]
    out += [build_reference_host(force, part) for part in (3, 4, 5)]
    return out


if __name__ == "__main__":
    for so in build_all("--force" in sys.argv):
        print(so)
