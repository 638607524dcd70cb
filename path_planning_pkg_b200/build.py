"""In-tree build of the native libraries (explicit nvcc / g++ command lines, no JIT cache).

  lib/libpp_b200.so            CUDA kernels + C ABI (include/pp_b200.h), sm_100a
  lib/libpath_planning_b200.so host C++ classes with the reference's API (include/path_planning_pkg/*.h)

Flags that matter for parity: no FMA contraction anywhere (-fmad=false on the
device, -ffp-contract=off on the host), IEEE division and square root, no flush-to-zero.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "lib")
CSRC = os.path.join(HERE, "csrc")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
              "-Xcompiler", "-fPIC,-ffp-contract=off,-O2", "-cudart", "static"]
CXX_FLAGS = ["-std=c++14", "-O2", "-Wall", "-Wno-unknown-pragmas", "-fPIC", "-ffp-contract=off"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _sources(*dirs):
    out = []
    for d in dirs:
        for r, _, fs in os.walk(d):
            out += [os.path.join(r, f) for f in fs if f.endswith((".h", ".cuh", ".cu", ".cpp", ".c"))]
    return out


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)


def build_cuda(force=False, verbose=True, extra=()):
    """nvcc -> lib/libpp_b200.so"""
    os.makedirs(LIB, exist_ok=True)
    target = os.path.join(LIB, "libpp_b200.so")
    srcs = _sources(CSRC, os.path.join(ROOT, "include"))
    if force or _newer(target, srcs):
        nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
        _nvcc_link(nvcc, list(extra), target, verbose)
    return target


def _nvcc_link(nvcc, defines, target, verbose):
    """pp_cabi.cu (bit-exact arithmetic: -fmad=false) + pp_fields.cu (tolerance-bound field kernels: FMA on) -> one .so"""
    obj_dir = os.path.join(LIB, "obj")
    os.makedirs(obj_dir, exist_ok=True)
    tag = os.path.basename(target).replace(".so", "")
    o1 = os.path.join(obj_dir, tag + "_cabi.o")
    o2 = os.path.join(obj_dir, tag + "_fields.o")
    _run([nvcc] + NVCC_FLAGS + defines + ["-c", "-o", o1, os.path.join(CSRC, "cabi", "pp_cabi.cu")], verbose)
    fields_flags = [f for f in NVCC_FLAGS if f != "-fmad=false"]
    _run([nvcc] + fields_flags + defines + ["-c", "-o", o2, os.path.join(CSRC, "kernels", "pp_fields.cu")], verbose)
    _run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static", "-o", target, o1, o2], verbose)


def build_cuda_profile(force=False, verbose=True):
    """Development variant with per-phase cycle counters: lib/libpp_b200_prof.so (never loaded by the package)."""
    os.makedirs(LIB, exist_ok=True)
    target = os.path.join(LIB, "libpp_b200_prof.so")
    srcs = _sources(CSRC, os.path.join(ROOT, "include"))
    if force or _newer(target, srcs):
        nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
        _nvcc_link(nvcc, ["-DPP_PROFILE"], target, verbose)
    return target


def build_cuda_variant(name, defines, force=True, verbose=True):
    """Development variants for A/B runs on the GPU box: lib/libpp_b200_<name>.so (never loaded by the package)."""
    os.makedirs(LIB, exist_ok=True)
    target = os.path.join(LIB, f"libpp_b200_{name}.so")
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    _nvcc_link(nvcc, [f"-D{d}" for d in defines], target, verbose)
    return target


def build_host(force=False, verbose=True):
    """g++ -> lib/libpath_planning_b200.so (reference-compatible C++ classes over the C ABI)"""
    os.makedirs(LIB, exist_ok=True)
    target = os.path.join(LIB, "libpath_planning_b200.so")
    hdir = os.path.join(CSRC, "host")
    cpps = sorted(os.path.join(hdir, f) for f in os.listdir(hdir) if f.endswith(".cpp"))
    if not cpps:
        return None
    srcs = _sources(CSRC, os.path.join(ROOT, "include"))
    if force or _newer(target, srcs):
        cxx = os.environ.get("CXX", "g++")
        _run([cxx] + CXX_FLAGS + ["-DSTORE_GRID_AS_REFERENCE",   # as the reference's CMakeLists.txt:128
                                  "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "include", "path_planning_pkg"),
                                  "-shared", "-o", target] + cpps +
             ["-L", LIB, "-lpp_b200", "-Wl,-rpath,$ORIGIN"], verbose)
    return target


def build_host_emul(force=False, verbose=True):
    """g++ -> tests/cpp/bin/libpp_host_emul.so (TEST ONLY: the PP_HD core on one host lane)"""
    out_dir = os.path.join(ROOT, "tests", "cpp", "bin")
    os.makedirs(out_dir, exist_ok=True)
    target = os.path.join(out_dir, "libpp_host_emul.so")
    src = os.path.join(ROOT, "tests", "cpp", "host_emul.cpp")
    srcs = _sources(CSRC, os.path.join(ROOT, "include")) + [src, os.path.join(ROOT, "oracle", "oracle_api.h")]
    if force or _newer(target, srcs):
        cxx = os.environ.get("CXX", "g++")
        _run([cxx] + CXX_FLAGS + ["-shared", "-o", target, src], verbose)
    return target


def build_gmath_check(force=False, verbose=True):
    """g++ -> tests/cpp/bin/gmath_check (TEST ONLY): exhaustive check of csrc/core/pp_gmath.h against the installed libm.
    -mfma so that fma() is the hardware instruction glibc's own __sinf_fma / __cosf_fma use; -fno-builtin keeps the libm calls calls."""
    out_dir = os.path.join(ROOT, "tests", "cpp", "bin")
    os.makedirs(out_dir, exist_ok=True)
    cxx = os.environ.get("CXX", "g++")
    drv = os.path.join(ROOT, "tests", "cpp", "gmath_check.cpp")
    target = os.path.join(out_dir, "gmath_check")
    if force or _newer(target, [drv, os.path.join(CSRC, "core", "pp_gmath.h"), os.path.join(CSRC, "core", "pp_defs.h")]):
        _run([cxx, "-std=c++14", "-O2", "-ffp-contract=off", "-mfma", "-fno-builtin", drv, "-o", target, "-lpthread", "-lm"], verbose)
    return target


def build_cpp_tests(force=False, verbose=True, reference="/root/reference"):
    """g++ -> tests/cpp/bin/api_driver (C++ API test driver) and, when the reference tree is present,
    tests/cpp/bin/local_planner_b200 = the reference's UNMODIFIED src/local_planner.cpp compiled against this repo's
    headers (with the in-process ROS stand-in tests/ros_stubs) and linked against libpath_planning_b200.so: the link check
    of the drop-in boundary and, driven by a script, the replay harness (tests/test_gpu_replay.py)."""
    out_dir = os.path.join(ROOT, "tests", "cpp", "bin")
    os.makedirs(out_dir, exist_ok=True)
    cxx = os.environ.get("CXX", "g++")
    inc = os.path.join(ROOT, "include", "path_planning_pkg")
    link = ["-L", LIB, "-lpath_planning_b200", "-lpp_b200", "-Wl,-rpath,$ORIGIN/../../../path_planning_pkg_b200/lib"]
    built = []
    drv = os.path.join(ROOT, "tests", "cpp", "api_driver.cpp")
    target = os.path.join(out_dir, "api_driver")
    if force or _newer(target, _sources(CSRC, os.path.join(ROOT, "include")) + [drv]):
        _run([cxx, "-std=c++14", "-O1", "-DSTORE_GRID_AS_REFERENCE", "-I", inc, drv, "-o", target] + link, verbose)
    built.append(target)
    # SURVEY 8(f) N3 / N4 parity driver (host classes only, runs without a GPU)
    drv = os.path.join(ROOT, "tests", "cpp", "velped_driver.cpp")
    target = os.path.join(out_dir, "velped_b200")
    if force or _newer(target, _sources(CSRC, os.path.join(ROOT, "include")) + [drv]):
        _run([cxx, "-std=c++14", "-O1", "-DSTORE_GRID_AS_REFERENCE", "-I", inc, drv, "-o", target] + link, verbose)
    built.append(target)
    built.append(build_gmath_check(force, verbose))
    lp = os.path.join(reference, "src", "local_planner.cpp")
    if os.path.exists(lp):
        target = os.path.join(out_dir, "local_planner_b200")
        if force or _newer(target, _sources(CSRC, os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "ros_stubs")) + [lp]):
            _run([cxx, "-std=c++14", "-O1", "-DSTORE_GRID_AS_REFERENCE", "-I", os.path.join(ROOT, "tests", "ros_stubs"), "-I", inc,
                  "-I", os.path.join(reference, "src"), lp, "-o", target] + link, verbose)
        built.append(target)
    return built


def build_oracle(verbose=True):
    """make -C oracle: CPU restatement + (when /root/reference is present) the compiled reference"""
    cmd = ["make", "-C", os.path.join(ROOT, "oracle"), "all"]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd, stdout=None if verbose else subprocess.DEVNULL)


def build_cuda_pinned(force=False, verbose=True):
    """lib/libpp_b200_pinned.so: the same library with the round-1 math policy (-DPP_MATH_PINNED: float transcendentals evaluated
    in double and rounded once; equals oracle/_ref/libref_oracle_crm.so).  Loaded only through PP_B200_LIB (tests/test_gpu_pinned_variant.py)."""
    target = os.path.join(LIB, "libpp_b200_pinned.so")
    srcs = _sources(CSRC, os.path.join(ROOT, "include"))
    if force or _newer(target, srcs):
        nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
        _nvcc_link(nvcc, ["-DPP_MATH_PINNED"], target, verbose)
    return target


def build_all(force=False, verbose=True):
    build_cuda(force, verbose)
    build_cuda_pinned(force, verbose)
    build_host(force, verbose)
    build_host_emul(force, verbose)
    build_cpp_tests(force, verbose)
    build_oracle(verbose)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
