"""In-tree builds: nvcc for the CUDA library (sm_100a), g++ for the synthetic-workload generator.
The built .so files live next to their sources (git-ignored, shipped to the GPU box by gpurun)."""
import contextlib
import fcntl
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBNPT = os.path.join(CSRC, "libnpt.so")
LIBSYNTH = os.path.join(HERE, "synth", "libnptsynth.so")
LIBBAM = os.path.join(HERE, "bam", "libnptbam.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--use_fast_math",
              "-Xcompiler", "-fPIC", "-shared", "-cudart", "shared", "-ldl", "-lpthread"]


@contextlib.contextmanager
def _locked(target):
    """One builder at a time per target (ranks of a torchrun job all import this module), and never a half-written .so:
    callers compile to a temporary name and os.replace() it under this lock."""
    with open(target + ".lock", "w") as f:
        fcntl.flock(f, fcntl.LOCK_EX)
        try:
            yield
        finally:
            fcntl.flock(f, fcntl.LOCK_UN)


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _nvcc():
    for c in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    return None


def build_libnpt(force=False, verbose=False):
    srcs = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h", ".cpp", ".inl")))
    srcs.append(os.path.join(ROOT, "include", "npt.h"))
    cus = [s for s in srcs if s.endswith((".cu", ".cpp"))]
    if not force and not _stale(LIBNPT, srcs):
        return LIBNPT
    with _locked(LIBNPT):
        if not force and not _stale(LIBNPT, srcs):
            return LIBNPT  # another process built it while we waited
        nvcc = _nvcc()
        if nvcc is None:
            if os.path.exists(LIBNPT):
                return LIBNPT  # GPU box without a toolchain on PATH: use the shipped build
            raise RuntimeError("nvcc not found and libnpt.so not built")
        tmp = LIBNPT + f".{os.getpid()}.tmp"
        cmd = [nvcc] + NVCC_FLAGS + os.environ.get("NPT_NVCC_EXTRA", "").split() + ["-I", os.path.join(ROOT, "include"), "-o", tmp] + cus
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        os.replace(tmp, LIBNPT)
        if verbose:
            print(r.stderr)
    return LIBNPT


def build_synth(force=False):
    src = os.path.join(HERE, "synth", "synth.cpp")
    if not force and not _stale(LIBSYNTH, [src]):
        return LIBSYNTH
    with _locked(LIBSYNTH):
        if not force and not _stale(LIBSYNTH, [src]):
            return LIBSYNTH
        cxx = shutil.which("g++") or "g++"
        tmp = LIBSYNTH + f".{os.getpid()}.tmp"
        cmd = [cxx, "-O2", "-std=c++17", "-fPIC", "-pthread", "-shared", "-o", tmp, src]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("g++ failed:\n" + r.stdout + r.stderr)
        os.replace(tmp, LIBSYNTH)
    return LIBSYNTH


def build_bam(force=False):
    """BAM (BGZF) -> SoA packer: host C++ + zlib, no CUDA."""
    src = os.path.join(HERE, "bam", "bam_pack.cpp")
    deps = [src, os.path.join(os.path.dirname(src), "fast_inflate.h"), os.path.join(ROOT, "include", "npt_bam.h")]
    if not force and not _stale(LIBBAM, deps):
        return LIBBAM
    with _locked(LIBBAM):
        if not force and not _stale(LIBBAM, deps):
            return LIBBAM
        cxx = shutil.which("g++") or "g++"
        tmp = LIBBAM + f".{os.getpid()}.tmp"
        cmd = [cxx, "-O3", "-std=c++17", "-fPIC", "-pthread", "-shared", "-Wall", "-Wextra", "-o", tmp, src, "-lz"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("g++ failed:\n" + r.stdout + r.stderr)
        os.replace(tmp, LIBBAM)
    return LIBBAM


def build_all(force=False):
    return build_libnpt(force), build_synth(force), build_bam(force)
