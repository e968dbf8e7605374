"""Build recipes for the in-tree native libraries.

  ezrt_b200/libezrt_b200.so   the product: C ABI (include/ezrt.h) + host scene pipeline + sm_100a kernels
  oracle/libezrt_oracle.so    the CPU oracle (test infrastructure, see oracle/README.md)
  oracle/_ref/libhdrloader_ref.so   the one reference translation unit that compiles stand-alone
                                     (P5/lib/hdrloader.cpp), built only where /root/reference exists
  oracle/_ref/libezrt_refhost.so    the reference's own host code (P5 main.cpp + hdrloader.cpp) compiled from where
                                     it lies against stand-in GL/GLUT/glm headers (oracle/ref_stubs/), same condition
  oracle/_ref/libezrt_refshader.so  the reference's own fragment shaders (P3/P4/P5 fshader.fsh) transpiled
                                     to C++ from where they lie (oracle/ref_shader/), same condition

Parity needs bit-identical fp32 arithmetic on host and device, hence
  host  : -ffp-contract=off -mfma      (FMA only where ezrt_math.h spells it)
  device: -fmad=false, IEEE div/sqrt, no FTZ (nvcc defaults), never -use_fast_math
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "ezrt_b200", "csrc")
INCLUDE = os.path.join(ROOT, "include")
PRODUCT_SO = os.path.join(ROOT, "ezrt_b200", "libezrt_b200.so")
ORACLE_SO = os.path.join(ROOT, "oracle", "libezrt_oracle.so")
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
REF_HDR_SO = os.path.join(REF_DIR, "libhdrloader_ref.so")
REF_SHADER_SO = os.path.join(REF_DIR, "libezrt_refshader.so")
REF_HOST_SO = os.path.join(REF_DIR, "libezrt_refhost.so")
REFERENCE_ROOT = "/root/reference"
REFERENCE_PARTS = ("part 3 -- OpenGL Raytracing", "part 4 -- Disney Principle BRDF",
                   "part 5 -- Importance Sampling & Low Discrepancy Sequence")
REFERENCE_P5 = os.path.join(REFERENCE_ROOT, REFERENCE_PARTS[2], "source code")

HOST_FLAGS = ["-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-mfma", "-Wall"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-mfma", "-Xptxas", "-v",
]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd, log=None):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if log:
        with open(log, "w") as f:
            f.write(" ".join(cmd) + "\n" + r.stdout)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("build failed: " + " ".join(cmd))
    return r.stdout


def _headers():
    hs = [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE)]
    hs += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh", ".inc"))]
    return hs


def build_product(force=False, verbose=False, variant=None, defines=()):
    """nvcc + g++ -> ezrt_b200/libezrt_b200.so (cross-compiles for sm_100a without a GPU).
    variant/defines build an experiment copy libezrt_b200_<variant>.so with extra -D macros
    (selected at import time by env EZRT_LIB_VARIANT; used for A/B runs on the GPU box)."""
    nvcc = _nvcc()
    if nvcc is None:
        raise RuntimeError("nvcc not found: cannot build libezrt_b200.so")
    target = PRODUCT_SO if not variant else PRODUCT_SO.replace(".so", "_%s.so" % variant)
    cu = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))
    cpp = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cpp"))
    if not force and not _newer(target, cu + cpp + _headers()):
        return target
    objdir = os.path.join(ROOT, "build", "obj" + ("_" + variant if variant else ""))
    os.makedirs(objdir, exist_ok=True)
    # one builder at a time (several ranks / pytest workers may import the package at once)
    import fcntl
    lock = open(os.path.join(ROOT, "build", ".lock"), "w")
    fcntl.flock(lock, fcntl.LOCK_EX)
    try:
        if not force and not _newer(target, cu + cpp + _headers()):
            return target
        return _build_product_locked(nvcc, target, cu, cpp, objdir, defines, verbose)
    finally:
        fcntl.flock(lock, fcntl.LOCK_UN)
        lock.close()


def _build_product_locked(nvcc, target, cu, cpp, objdir, defines, verbose):
    dflags = ["-D" + d for d in defines]
    objs = []
    for src in cu:
        obj = os.path.join(objdir, os.path.basename(src) + ".o")
        out = _run([nvcc] + NVCC_FLAGS + dflags + ["-I", INCLUDE, "-I", CSRC, "-c", src, "-o", obj],
                   log=os.path.join(objdir, os.path.basename(src) + ".ptxas.log"))
        if verbose:
            print(out)
        objs.append(obj)
    for src in cpp:
        obj = os.path.join(objdir, os.path.basename(src) + ".o")
        _run(["g++"] + HOST_FLAGS + dflags + ["-pthread", "-I", INCLUDE, "-I", CSRC, "-c", src, "-o", obj])
        objs.append(obj)
    tmp = target + ".tmp%d" % os.getpid()
    _run([nvcc, "-shared", "-o", tmp] + objs + ["-Xcompiler", "-pthread", "-cudart", "static"])
    os.replace(tmp, target)  # atomic: a concurrent import never sees a half-written library
    return target


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


EXAMPLE_BIN = os.path.join(ROOT, "examples", "ezrt_main")


def build_example(force=False):
    """examples/ezrt_main: the reference's main() + display() as a C++ host over the C ABI (links libezrt_b200.so)."""
    src = os.path.join(ROOT, "examples", "ezrt_main.cpp")
    if force or _newer(EXAMPLE_BIN, [src, PRODUCT_SO, os.path.join(INCLUDE, "ezrt.h")]):
        tmp = EXAMPLE_BIN + ".tmp%d" % os.getpid()
        _run(["g++", "-O2", "-std=c++17", "-Wall", "-I", INCLUDE, src, "-L", os.path.dirname(PRODUCT_SO), "-lezrt_b200",
              "-Wl,-rpath,$ORIGIN/../ezrt_b200", "-o", tmp])
        os.replace(tmp, EXAMPLE_BIN)
    return EXAMPLE_BIN


def build_all(force=False, verbose=False):
    build_product(force=force, verbose=verbose)
    build_oracle(force=force)
    build_example(force=force)
    build_reference_hdrloader(force=force)
    build_reference_shaders(force=force)
    for part in (3, 4, 5):
        build_reference_host(force=force, part=part)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built", PRODUCT_SO, ORACLE_SO)
