"""Builds libgalah_b200.so (CUDA kernels + C ABI + C++ host side) in-tree with nvcc for sm_100a.

The built .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
import glob
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libgalah_b200.so")

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function,-pthread", "--expt-relaxed-constexpr",
]


def _nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; cannot build libgalah_b200.so")
    return nvcc


def sources():
    cu = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    cpp = sorted(glob.glob(os.path.join(CSRC, "host", "*.cpp")))
    return cu + cpp


def _deps():
    deps = sources()
    deps += glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "host", "*.hpp"))
    deps += glob.glob(os.path.join(PKG_DIR, "..", "include", "*.h"))
    return deps


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile the shared library. Returns the .so path."""
    if force or _stale(LIB_PATH, _deps()):
        objdir = os.path.join(PKG_DIR, "build")
        os.makedirs(objdir, exist_ok=True)
        objs = []
        procs = []
        for src in sources():
            obj = os.path.join(objdir, os.path.basename(src) + ".o")
            objs.append(obj)
            if not force and not _stale(obj, [src] + [d for d in _deps() if d.endswith((".cuh", ".hpp", ".h"))]):
                continue
            cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-x", "cu"] * src.endswith(".cu") + ["-c", src, "-o", obj]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        for src, p in procs:
            out, _ = p.communicate()
            if verbose or p.returncode:
                sys.stderr.write(out)
            if p.returncode:
                raise RuntimeError(f"nvcc failed on {src}")
        link = [_nvcc(), "-shared", "-o", LIB_PATH] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                               "-Xcompiler", "-pthread", "-l:libz.so.1"]
        subprocess.check_call(link)
    build_cli(force)
    build_tools(force)
    return LIB_PATH


def build_tools(force=False):
    """tools/bin/pipe_bench: the integer pipe-rate micro-benchmark behind K1's instruction roofline (tools/pipe_bench.cu)."""
    src = os.path.join(PKG_DIR, "..", "tools", "pipe_bench.cu")
    out = os.path.join(PKG_DIR, "..", "tools", "bin", "pipe_bench")
    if os.path.exists(src) and (force or _stale(out, [src])):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call([_nvcc(), "-O3", "-diag-suppress", "177", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out, src])
    return out


CLI_PATH = os.path.join(PKG_DIR, "galah-b200")


def build_cli(force=False):
    """galah-b200: the `galah cluster` command line over the C ABI (csrc/cli/main.cpp), next to the library."""
    src = os.path.join(CSRC, "cli", "main.cpp")
    hdr = os.path.join(PKG_DIR, "..", "include", "galah_b200.h")
    if force or _stale(CLI_PATH, [src, hdr, LIB_PATH]):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", "-I", os.path.join(PKG_DIR, "..", "include"), src, "-o", CLI_PATH,
                               "-L", PKG_DIR, "-l:libgalah_b200.so", "-Wl,-rpath,$ORIGIN"])
    return CLI_PATH


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB_PATH)
