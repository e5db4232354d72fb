"""Build libbvht_cuda.so (the C-ABI library of include/bvht.h) in-tree with nvcc for sm_100a.

    python -m bvhtracer_b200.build [--force]

The trace kernels are compiled twice from the same source with different floating-point flags:
strict (--fmad=false, IEEE div/sqrt, no FTZ) and fast (--fmad=true).  nvcc cross-compiles without a GPU.
"""
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB_DIR = os.path.join(PKG, "lib")
OBJ_DIR = os.path.join(PKG, "build")
LIB_PATH = os.path.join(LIB_DIR, "libbvht_cuda.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
          "-Xcompiler", "-fno-fast-math", "-Xcompiler", "-ffp-contract=off"]
IEEE = ["-prec-div=true", "-prec-sqrt=true", "-ftz=false"]
STRICT = ["--fmad=false"] + IEEE
FAST = ["--fmad=true"] + IEEE

UNITS = [
    # source,             extra flags
    ("trace_strict.cu", STRICT),
    ("trace_fast.cu", FAST),
    ("trace_stats.cu", STRICT),
    ("upload_kernels.cu", STRICT),
    ("refit_kernels.cu", STRICT),
    ("scene_kernels.cu", STRICT),
    ("cover_kernels.cu", STRICT),
    ("build_kernels.cu", STRICT),
    ("bvht_api.cu", STRICT),
    ("leaf_accel.cpp", []),
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _deps():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(os.path.dirname(PKG), "include", "bvht.h")]


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(d) > t for d in _deps())


def build(force=False, verbose=False, variant=None, defs=()):
    """variant / defs: tuning builds (tools/build_variants.py): the same sources with extra -D knobs, written to
    lib/variants/libbvht_cuda_<variant>.so; the product library is always built without them."""
    lib_path, obj_dir = LIB_PATH, OBJ_DIR
    if variant:
        lib_path = os.path.join(LIB_DIR, "variants", f"libbvht_cuda_{variant}.so")
        obj_dir = os.path.join(OBJ_DIR, "variant_" + variant)
        os.makedirs(os.path.dirname(lib_path), exist_ok=True)
        force = True
    if not force and not needs_build():
        return LIB_PATH
    nvcc = _nvcc()
    extra_defs = ["-D" + d for d in defs]
    if os.environ.get("BVHT_MIN_BLOCKS"):            # tuning knob: register budget of the trace kernels
        extra_defs.append("-DBVHT_MIN_BLOCKS=" + os.environ["BVHT_MIN_BLOCKS"])
    if os.environ.get("BVHT_GRAB"):                  # tuning knob: 32-pixel slices pulled per work-counter atomic
        extra_defs.append("-DBVHT_GRAB=" + os.environ["BVHT_GRAB"])
    for knob in ("BVHT_SUB_CH", "BVHT_TLAS_PRUNE", "BVHT_SLICE_LOG"):                    # experiment knobs passed through to every translation unit
        if os.environ.get(knob):
            extra_defs.append("-D%s=%s" % (knob, os.environ[knob]))
    os.makedirs(LIB_DIR, exist_ok=True)
    os.makedirs(obj_dir, exist_ok=True)
    objs = []
    procs = []
    for src, extra in UNITS:
        obj = os.path.join(obj_dir, os.path.splitext(src)[0] + ".o")
        cmd = [nvcc] + ARCH + COMMON + extra + extra_defs + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, cmd, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"[bvht build] {src} FAILED\n{' '.join(cmd)}\n{out}\n")
        elif verbose and out.strip():
            sys.stderr.write(f"[bvht build] {src}\n{out}\n")
    if failed:
        raise RuntimeError("nvcc compilation failed")
    tmp = lib_path + ".tmp"
    cmd = [nvcc] + ARCH + ["-shared", "-o", tmp] + objs
    subprocess.check_call(cmd)
    os.replace(tmp, lib_path)
    return lib_path


HOST_DIR = os.path.join(PKG, "host")
HOST_LIB = os.path.join(LIB_DIR, "libbvht_host.so")


def build_host(force=False):
    """C++ host mirror (bvhtracer_b200/host): g++ with rustc's arithmetic model, linked against libbvht_cuda.so."""
    srcs = [os.path.join(HOST_DIR, f) for f in ("capi.cpp", "bvhtracer.hpp")] + [os.path.join(os.path.dirname(PKG), "include", "bvht.h")]
    if not force and os.path.exists(HOST_LIB) and all(os.path.getmtime(s) <= os.path.getmtime(HOST_LIB) for s in srcs) \
            and os.path.getmtime(LIB_PATH) <= os.path.getmtime(HOST_LIB):
        return HOST_LIB
    cxx = shutil.which("g++") or "g++"
    tmp = HOST_LIB + ".tmp"
    cmd = [cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-fvisibility=hidden", "-Wall",
           os.path.join(HOST_DIR, "capi.cpp"), "-o", tmp, "-L" + LIB_DIR, "-lbvht_cuda", "-Wl,-rpath,$ORIGIN"]
    subprocess.check_call(cmd)
    os.replace(tmp, HOST_LIB)
    return HOST_LIB


def build_all(force=False, verbose=False):
    build(force=force, verbose=verbose)
    build_host(force=force)
    return LIB_PATH, HOST_LIB


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv, verbose="-v" in sys.argv))
