"""Build the in-tree native libraries (nvcc, sm_100a).  Used by __graft_entry__.build()."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libdpgo_b200.so")
CUDA_HOME = os.environ.get("CUDA_HOME", "/usr/local/cuda")
NVCC = os.path.join(CUDA_HOME, "bin", "nvcc")

CU_SOURCES = ["device_lib.cu", "fused_rtr.cu", "precon_dd.cu"]
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-cudart", "shared", "-diag-suppress", "177"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_device_lib(force=False, verbose=False, trace=False):
    """trace=True: measurement build (-DDPGO_TRACE: per-CTA phase times in the fused solver, read with
    dpgo_phase_trace); always rebuilt, and the next normal build replaces it."""
    flags = FLAGS + (["-DDPGO_TRACE"] if trace else [])
    marker = os.path.join(CSRC, ".trace_build")
    if trace or os.path.exists(marker):
        force = True
    if trace:
        open(marker, "w").close()
    elif os.path.exists(marker):
        os.remove(marker)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if not f.startswith(".") and not f.endswith(".o")]
    deps.append(os.path.join(HERE, "..", "include", "dpgo_b200.h"))
    objs = []
    for src in CU_SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(CSRC, src.replace(".cu", ".o"))
        if force or _stale(o, deps):
            cmd = [NVCC] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            print(" ".join(cmd), flush=True)
            subprocess.check_call(cmd)
        objs.append(o)
    if force or _stale(LIB, objs):
        cmd = [NVCC, "-shared", "-cudart", "shared", "-o", LIB] + objs + [
            "-L" + os.path.join(CUDA_HOME, "lib64"), "-lcusolver", "-lcublas",
            "-Xlinker", "-rpath," + os.path.join(CUDA_HOME, "lib64")]
        print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return LIB


HOST = os.path.join(HERE, "host")
REFERENCE_EXAMPLES = "/root/reference/examples"


def build_host(force=False):
    """C++ drop-in host API (libDPGO.so), its acceptance tests, and -- when the reference tree is
    present (build container only) -- the reference's own example drivers compiled UNMODIFIED
    against the drop-in headers (only the built binaries travel to the GPU box)."""
    inc = os.path.join(HOST, "include")
    srcs = sorted(os.path.join(HOST, "src", f) for f in os.listdir(os.path.join(HOST, "src")) if f.endswith(".cpp"))
    hdrs = []
    for root, _, files in os.walk(inc):
        hdrs += [os.path.join(root, f) for f in files]
    hdrs += [os.path.join(HOST, "src", "check.h"), os.path.join(HERE, "..", "include", "dpgo_b200.h"), LIB]
    lib = os.path.join(HOST, "libDPGO.so")
    link = ["-L" + HERE, "-ldpgo_b200", "-Wl,-rpath,$ORIGIN/..", "-Wl,-rpath,$ORIGIN/../..", "-pthread"]
    if force or _stale(lib, srcs + hdrs):
        cmd = ["g++", "-std=c++17", "-O2", "-fPIC", "-Wall", "-Wextra", "-shared", "-I" + inc] + srcs + link + ["-o", lib]
        print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    bindir = os.path.join(HOST, "bin")
    os.makedirs(bindir, exist_ok=True)
    targets = [(os.path.join(HOST, "tests", "host_tests.cpp"), "host_tests"),
               (os.path.join(HOST, "tests", "robust_pgo_test.cpp"), "robust_pgo_test")]
    if os.path.isdir(REFERENCE_EXAMPLES):
        targets += [(os.path.join(REFERENCE_EXAMPLES, "MultiRobotExample.cpp"), "multi-robot-example"),
                    (os.path.join(REFERENCE_EXAMPLES, "SingleRobotExample.cpp"), "single-robot-example"),
                    (os.path.join(REFERENCE_EXAMPLES, "ChordalInitializationExample.cpp"),
                     "chordal-initialization-example")]
    for src, name in targets:
        out = os.path.join(bindir, name)
        if force or _stale(out, [src, lib] + hdrs):
            cmd = ["g++", "-std=c++17", "-O2", "-I" + inc, src, "-L" + HOST, "-lDPGO"] + link + ["-o", out]
            print(" ".join(cmd), flush=True)
            subprocess.check_call(cmd)
    return lib


if __name__ == "__main__":
    build_device_lib(force="--force" in sys.argv, verbose="-v" in sys.argv, trace="--trace" in sys.argv)
    build_host(force="--force" in sys.argv)
