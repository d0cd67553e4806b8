"""Build the engine's native library in-tree: csrc/usvmpc_api.cu -> libusvmpc.so (sm_100a only).

    python -m mpc_collisionavoidance_b200.build [--force]

nvcc cross-compiles without a GPU, so this also runs on the CPU-only build box; the resulting .so is
git-ignored but travels to the GPU box with the snapshot.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libusvmpc.so")
SOURCES = [os.path.join(CSRC, f) for f in ("usvmpc_api.cu", "cta_kernel.cuh", "models.cuh", "cta_layout.h", "cta_compat.h")]
SOURCES.append(os.path.join(HERE, "..", "include", "usvmpc.h"))


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the engine is CUDA-only and cannot be built without the CUDA toolkit")


def up_to_date(lib=LIB):
    return os.path.exists(lib) and all(os.path.getmtime(lib) >= os.path.getmtime(s) for s in SOURCES)


def build(force=False, min_ctas=None, out=LIB, verbose=False, defines=()):
    """Compile for sm_100a.  `min_ctas` = resident CTAs (4 warps each) per SM the register allocator must allow."""
    if not force and min_ctas is None and not defines and up_to_date(out):
        return out
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-Xptxas", "-v", "-shared", "-Xcompiler", "-fPIC", "-o", out, SOURCES[0]]
    if min_ctas is not None:
        cmd.insert(1, f"-DUSVMPC_MIN_CTAS={int(min_ctas)}")
    for d in defines:
        cmd.insert(1, "-D" + d)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
