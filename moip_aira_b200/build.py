"""Builds libmoip_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmoip_b200.so")
SOURCES = ["model.cpp", "capi.cpp", "generator.cpp", "solver.cu", "k1_pdhg.cu", "k1_fast.cu", "k1_small.cu", "k1_reg.cu", "k1_reg_kd2.cu", "k1_reg_kd3.cu", "k1_reg_kd4.cu", "k1_reg_kd5.cu", "k2_nodepool.cu", "k3_k4.cu", "k5_chain.cu"]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "moip_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(HERE, "build", src + ".o")
        cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
               "-Xcompiler", "-fPIC", "-x", "cu", "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- {src}\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + objs
    subprocess.check_call(cmd)
    return LIB


SEAM1 = os.path.join(HERE, "seam1")
SEAM1_LIB = os.path.join(SEAM1, "libcplex_moip_b200.so")


def build_seam1(force=False):
    """Link-level CPLEX seam (Seam 1): seam1/cpx_shim.cpp -> seam1/libcplex_moip_b200.so, host code only
    (g++), linked against libmoip_b200.so next to it."""
    src = os.path.join(SEAM1, "cpx_shim.cpp")
    deps = [src, os.path.join(SEAM1, "include", "ilcplex", "cplex.h"), os.path.join(HERE, "..", "include", "moip_b200.h"), LIB]
    if not force and os.path.exists(SEAM1_LIB) and all(os.path.getmtime(d) <= os.path.getmtime(SEAM1_LIB) for d in deps):
        return SEAM1_LIB
    cxx = os.environ.get("CXX", "g++")
    subprocess.check_call([cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", src, "-o", SEAM1_LIB,
                           "-L" + HERE, "-lmoip_b200", "-Wl,-rpath,$ORIGIN/.."])
    return SEAM1_LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_seam1(force="--force" in sys.argv))
