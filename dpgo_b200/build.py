"""Build the in-tree native libraries (nvcc, sm_100a).  Used by __graft_entry__.build()."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libdpgo_b200.so")
CUDA_HOME = os.environ.get("CUDA_HOME", "/usr/local/cuda")
NVCC = os.path.join(CUDA_HOME, "bin", "nvcc")

CU_SOURCES = ["device_lib.cu", "fused_rtr.cu", "precon_dd.cu", "dense_la.cu", "exchange.cu"]
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-cudart", "shared", "-diag-suppress", "177"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_device_lib(force=False, verbose=False, trace=False, variant=None, defines=()):
    """The product library libdpgo_b200.so, or a measurement build beside it (own object files, loaded only
    when DPGO_B200_LIB points at it): trace=True -> libdpgo_b200_trace.so with -DDPGO_TRACE (per-CTA phase
    times in the fused solver, tools/phase_trace.py); variant="name", defines=["X=1", ...] ->
    libdpgo_b200_<name>.so with -DX=1 ... (A/B runs of compile-time choices on one box)."""
    if trace:
        variant, defines = "trace", ["DPGO_TRACE"]
    flags = FLAGS + ["-D" + d for d in defines]
    suffix = f".{variant}.o" if variant else ".o"
    out = LIB.replace(".so", f"_{variant}.so") if variant else LIB
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if not f.startswith(".") and not f.endswith(".o")]
    deps.append(os.path.join(HERE, "..", "include", "dpgo_b200.h"))
    objs = []
    procs = []
    for src in CU_SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(CSRC, src.replace(".cu", suffix))
        if force or _stale(o, deps):
            cmd = [NVCC] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            print(" ".join(cmd), flush=True)
            procs.append((cmd, subprocess.Popen(cmd)))     # the translation units compile side by side
        objs.append(o)
    for cmd, pr in procs:
        if pr.wait() != 0:
            raise subprocess.CalledProcessError(pr.returncode, cmd)
    if force or _stale(out, objs):
        # no library on the link line but the CUDA runtime: the dense factorizations of the set-up are in-tree
        # (dense_la.cu); libcusolver / libcublas / libcublasLt (1.2 GB, minutes to page in on a fresh box) are gone
        cmd = [NVCC, "-shared", "-cudart", "shared", "-o", out] + objs + [
            "-ldl", "-Xlinker", "-rpath," + os.path.join(CUDA_HOME, "lib64")]
        print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return out


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
               (os.path.join(HOST, "tests", "robust_pgo_test.cpp"), "robust_pgo_test"),
               (os.path.join(HOST, "tests", "host_cli.cpp"), "host_cli")]
    if os.path.isdir(REFERENCE_EXAMPLES):
        targets += [(os.path.join(REFERENCE_EXAMPLES, "MultiRobotExample.cpp"), "multi-robot-example"),
                    (os.path.join(REFERENCE_EXAMPLES, "SingleRobotExample.cpp"), "single-robot-example"),
                    (os.path.join(REFERENCE_EXAMPLES, "ChordalInitializationExample.cpp"),
                     "chordal-initialization-example")]
    stamp = host_abi_stamp()
    for src, name in targets:
        out = os.path.join(bindir, name)
        if force or _stale(out, [src, lib] + hdrs) or _read(out + ".abi") != stamp:
            cmd = ["g++", "-std=c++17", "-O2", "-I" + inc, src, "-L" + HOST, "-lDPGO"] + link + ["-o", out]
            print(" ".join(cmd), flush=True)
            subprocess.check_call(cmd)
            with open(out + ".abi", "w") as f:
                f.write(stamp)
    return lib


def _read(path):
    try:
        with open(path) as f:
            return f.read()
    except OSError:
        return None


def host_abi_stamp():
    """Hash of everything a driver binary in host/bin compiles in or links against by layout: the drop-in
    headers (inline code, class layouts), the C-ABI header and the sources of libDPGO.  A binary whose
    recorded stamp differs is stale (the reference's example drivers can only be rebuilt where the reference
    tree is present, so they travel to the GPU box as built files): tests refuse to run it."""
    import hashlib
    h = hashlib.sha256()
    files = [os.path.join(HERE, "..", "include", "dpgo_b200.h")]
    for root in (os.path.join(HOST, "include"), os.path.join(HOST, "src")):
        for r, _, fs in os.walk(root):
            files += [os.path.join(r, f) for f in fs]
    for f in sorted(os.path.abspath(x) for x in files):
        h.update(os.path.relpath(f, HERE).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def host_binary(name):
    """Path of a driver binary in host/bin, or raises when it is missing or stale against the drop-in sources."""
    out = os.path.join(HOST, "bin", name)
    if not os.path.exists(out):
        raise FileNotFoundError(f"{out} not built: run __graft_entry__.build() where the sources are present")
    if _read(out + ".abi") != host_abi_stamp():
        raise RuntimeError(f"{out} is stale against dpgo_b200/host (headers or sources changed since it was built)")
    return out


if __name__ == "__main__":
    if "--variant" in sys.argv:    # python build.py --variant name -DX=1 -DY ...
        name = sys.argv[sys.argv.index("--variant") + 1]
        build_device_lib(variant=name, defines=[a[2:] for a in sys.argv if a.startswith("-D")])
    else:
        build_device_lib(force="--force" in sys.argv, verbose="-v" in sys.argv, trace="--trace" in sys.argv)
        if "--trace" not in sys.argv:
            build_host(force="--force" in sys.argv)
