"""Build the compiled CPU port of the oracle (oracle/cpu_port/libdpgo_cpu.so).  Called by
__graft_entry__.build(); building the checker is not using it."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "cpu_port", "dpgo_cpu.cpp")
LIB = os.path.join(HERE, "cpu_port", "libdpgo_cpu.so")
DEPS = [SRC]


def build(force=False):
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in DEPS):
        return LIB
    # x86-64-v3 (AVX2 + FMA) rather than -march=native: the .so is built in one container and
    # runs on the GPU box's host CPU
    cmd = ["g++", "-O3", "-march=x86-64-v3", "-std=c++17", "-fPIC", "-shared", SRC, "-o", LIB]
    print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    build(force=True)
