"""In-tree build of libzkfhe_b200.so with nvcc for sm_100a (no torch, no cmake).

    python zk-fhe_b200/build.py [--force] [--verbose]

The shared library lands in zk-fhe_b200/lib/ (git-ignored, but it travels to the
GPU box with the gpurun snapshot).  `ff_ptx_gen.cuh` is regenerated from
gen_ff_ptx.py first, which re-runs the instruction-level self check.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libzkfhe_b200.so")
SOURCES = ["capi.cu", "ntt.cu", "msm.cu", "witness.cu", "keygen.cu", "prover.cu", "verifier.cu", "comm.cu"]
HOST_SOURCES = [("poseidon_ifma.cpp", ["-mavx512f", "-mavx512ifma", "-mavx512vl", "-mbmi2", "-madx"])]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC",
    "-Xcompiler", "-Wall", "-Xcompiler", "-Wno-unknown-pragmas", "-diag-suppress", "177",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _stamp():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include"), os.path.join(HERE, "host")):
        for name in sorted(os.listdir(root)):
            if name.endswith((".cu", ".cuh", ".h", ".py", ".hpp", ".cpp")):
                with open(os.path.join(root, name), "rb") as f:
                    h.update(name.encode() + f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    stamp_path = os.path.join(LIBDIR, ".stamp")
    stamp = _stamp()
    if (not force and os.path.exists(LIB) and os.path.exists(stamp_path)
            and open(stamp_path).read() == stamp):
        return LIB
    for gen in ("gen_ff_ptx.py", "gen_poseidon.py", "gen_pairing_consts.py"):
        subprocess.run([sys.executable, os.path.join(CSRC, gen)], check=True,
                       stdout=None if verbose else subprocess.DEVNULL)
    stamp = _stamp()
    nvcc = _nvcc()
    # headers + flags hash: an object is rebuilt when its source or any header changed
    hh = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for name in sorted(os.listdir(root)):
            if name.endswith((".cuh", ".h")):
                hh.update(open(os.path.join(root, name), "rb").read())
    objs = []
    procs = []
    for src in SOURCES:
        path = os.path.join(CSRC, src)
        if not os.path.exists(path):
            continue
        obj = os.path.join(LIBDIR, src.replace(".cu", ".o"))
        ostamp = hashlib.sha256(hh.digest() + open(path, "rb").read()).hexdigest()
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.exists(obj + ".stamp") and open(obj + ".stamp").read() == ostamp:
            continue
        cmd = [nvcc, *NVCC_FLAGS, "-c", path, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd))
        procs.append((src, obj, ostamp, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, obj, ostamp, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            print(out)
        if p.returncode:
            raise RuntimeError(f"nvcc failed on {src}")
        with open(obj + ".stamp", "w") as f:
            f.write(ostamp)
    # host-only translation units with their own ISA flags (the callers dispatch on cpuid): the AVX-512 IFMA Poseidon
    for src, flags in HOST_SOURCES:
        path = os.path.join(CSRC, src)
        obj = os.path.join(LIBDIR, src.replace(".cpp", ".o"))
        ostamp = hashlib.sha256(hh.digest() + open(path, "rb").read() + " ".join(flags).encode()).hexdigest()
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.exists(obj + ".stamp") and open(obj + ".stamp").read() == ostamp:
            continue
        cmd = ["g++", "-O3", "-std=c++17", "-fPIC", "-Wall", *flags, "-c", path, "-o", obj]
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
        with open(obj + ".stamp", "w") as f:
            f.write(ostamp)
    cmd = [nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart", "-ldl"]
    subprocess.run(cmd, check=True)
    # the `bfv` example entrypoint (host C++ mirror of examples/bfv.rs over the C ABI)
    bindir = os.path.join(HERE, "bin")
    os.makedirs(bindir, exist_ok=True)
    subprocess.run(["g++", "-O2", "-std=c++17", "-Wall", os.path.join(HERE, "host", "bfv.cpp"), "-o",
                    os.path.join(bindir, "bfv"), "-L" + LIBDIR, "-lzkfhe_b200", "-Wl,-rpath,$ORIGIN/../lib",
                    "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"], check=True)
    with open(stamp_path, "w") as f:
        f.write(stamp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
