"""Build libadni_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

The library is compiled from multimodal_alzheimer_b200/csrc/*.cu into
multimodal_alzheimer_b200/csrc/libadni_b200.so.  It is git-ignored but travels to the GPU box.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libadni_b200.so")
SOURCES = [
    "runtime.cu",
    "conv_igemm_kernels.cu",
    "conv_halo.cu",
    "conv_wgrad2.cu",
    "conv_wgrad_halo.cu",
    "conv_api.cu",
    "conv_direct.cu",
    "conv_small.cu",
    "conv_stem.cu",
    "stem_fused.cu",
    "elementwise.cu",
    "dropout.cu",
    "pool_small.cu",
    "head_loss.cu",
    "metrics.cu",
    "normalize.cu",
    "optimizer.cu",
    "fusion_ops.cu",
    "peer_reduce.cu",
]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (needed to build libadni_b200.so)")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "adni_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a and link the shared library."""
    if not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    objdir = os.path.join(CSRC, "build")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    objs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [nvcc] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"[adni_b200 build] {src} failed:\n{out}\n")
        elif verbose and out:
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("nvcc compilation failed")
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-cudart", "static"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout)
    return LIB


STAGE_SRC = os.path.join(CSRC, "stage", "nifti_stage.cpp")
STAGE_LIB = os.path.join(CSRC, "libadni_stage.so")
STAGE_FLAGS = ["-O3", "-std=c++17", "-fPIC", "-shared", "-pthread", "-ffp-contract=off", "-Wall", "-Wl,--exclude-libs,ALL"]


def stage_needs_build():
    if not os.path.exists(STAGE_LIB):
        return True
    t = os.path.getmtime(STAGE_LIB)
    deps = [STAGE_SRC, os.path.join(os.path.dirname(HERE), "include", "adni_staging.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_stage(force=False):
    """g++ build of the host-side input staging library (NIfTI decode; zlib, no CUDA) -> csrc/libadni_stage.so."""
    if not force and not stage_needs_build():
        return STAGE_LIB
    cxx = os.environ.get("CXX") or shutil.which("g++") or "g++"
    cmd = [cxx] + STAGE_FLAGS + [STAGE_SRC, "-o", STAGE_LIB, "-lz"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("staging library build failed:\n" + r.stdout)
    return STAGE_LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_stage(force="--force" in sys.argv))
