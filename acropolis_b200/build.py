"""In-tree build of the CUDA library:  python -m acropolis_b200.build [--force]

nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ...  -> acropolis_b200/libacropolis_b200.so
(cross-compiles without a GPU; the .so travels to the GPU box with the repo snapshot)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = [os.path.join(HERE, "csrc", f) for f in ("acro_kernels.cu", "acro_api.cu")]
DEPS = [os.path.join(HERE, "csrc", f) for f in sorted(os.listdir(os.path.join(HERE, "csrc")))] + \
    [os.path.join(HERE, "..", "include", "acropolis_b200.h")]
OUT = os.path.join(HERE, "libacropolis_b200.so")
OUT_XCHECK = os.path.join(HERE, "libacropolis_b200_xcheck.so")   # same sources + the cross-check kernel modes (tests only)
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def nvcc_path():
    for c in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def up_to_date(out=OUT):
    if not os.path.exists(out):
        return False
    t = os.path.getmtime(out)
    return all(os.path.getmtime(d) <= t for d in DEPS)


def build_xcheck_library(force=False):
    """The test build: product sources with -DACRO_CROSSCHECK (kernel_mode PLAIN_GL / PER_T available)."""
    if not force and up_to_date(OUT_XCHECK):
        return OUT_XCHECK
    return build_library(force=True, out=OUT_XCHECK, defines=("-DACRO_CROSSCHECK",))


def build_library(force=False, verbose=False, out=None, defines=()):
    """``out``/``defines`` build a variant library next to the product one (kernel A/B runs, tools/)."""
    if out is None and not force and up_to_date():
        return OUT
    out = OUT if out is None else out
    cmd = [nvcc_path()] + NVCC_FLAGS + list(defines) + (["-Xptxas", "-v"] if verbose else []) + ["-o", out] + SRC
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return out


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_xcheck_library(force="--force" in sys.argv))
