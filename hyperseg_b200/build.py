"""Build libhsb200.so (sm_100a) in-tree with nvcc.

The shared library is the product's only native artefact: hand-written CUDA kernels behind the C ABI
declared in include/hsb200.h.  It is built next to this file so that it travels with the source tree
(the GPU box has no compiler cache) and is git-ignored.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
REPO_ROOT = PKG_DIR.parent
CSRC = PKG_DIR / "csrc"
INCLUDE = REPO_ROOT / "include"
LIB_PATH = PKG_DIR / "libhsb200.so"
BUILD_DIR = PKG_DIR / "csrc" / "build"

ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "--use_fast_math=false", "-Xcompiler", "-fPIC",
              "-Xptxas", "-v", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libhsb200.so cannot be built (there is no non-CUDA fallback)")
    return exe


def sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _digest() -> str:
    h = hashlib.sha256()
    for f in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(INCLUDE.glob("*.h"))):
        h.update(f.name.encode())
        h.update(f.read_bytes())
    h.update(" ".join(ARCH_FLAGS + NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every csrc/*.cu for sm_100a and link libhsb200.so. Returns the library path."""
    BUILD_DIR.mkdir(parents=True, exist_ok=True)
    stamp = BUILD_DIR / "digest.txt"
    digest = _digest()
    if not force and LIB_PATH.exists() and stamp.exists() and stamp.read_text() == digest:
        return LIB_PATH
    nvcc = _nvcc()
    objs = []
    procs = []
    for src in sources():
        obj = BUILD_DIR / (src.stem + ".o")
        cmd = [nvcc, *ARCH_FLAGS, *[f for f in NVCC_FLAGS if f != "--use_fast_math=false"],
               "-I", str(INCLUDE), "-I", str(CSRC), "-c", str(src), "-o", str(obj)]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    log = []
    failed = False
    for src, pr in procs:
        out, _ = pr.communicate()
        log.append(f"==== {src.name} ====\n{out}")
        if pr.returncode != 0:
            failed = True
    (BUILD_DIR / "nvcc.log").write_text("\n".join(log))
    if verbose or failed:
        print("\n".join(log))
    if failed:
        raise RuntimeError("nvcc failed, see " + str(BUILD_DIR / "nvcc.log"))
    link = [nvcc, *ARCH_FLAGS, "-shared", "-Xcompiler", "-fPIC", "-o", str(LIB_PATH), *map(str, objs)]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        print(r.stdout)
        raise RuntimeError("link of libhsb200.so failed")
    stamp.write_text(digest)
    return LIB_PATH


def build_variant(name: str, defines: list[str]) -> Path:
    """Experiment builds (scripts/): libhsb200_<name>.so with extra -D flags, selected at run time with the
    HSB_LIBRARY environment variable.  Not part of the product build."""
    out_dir = BUILD_DIR / name
    out_dir.mkdir(parents=True, exist_ok=True)
    nvcc = _nvcc()
    procs = []
    for src in sources():
        cmd = [nvcc, *ARCH_FLAGS, *[f for f in NVCC_FLAGS if f != "--use_fast_math=false"], *[f"-D{d}" for d in defines],
               "-I", str(INCLUDE), "-I", str(CSRC), "-c", str(src), "-o", str(out_dir / (src.stem + ".o"))]
        procs.append(subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for pr in procs:
        out, _ = pr.communicate()
        if pr.returncode != 0:
            print(out)
            raise RuntimeError("nvcc failed for variant " + name)
    lib = PKG_DIR / f"libhsb200_{name}.so"
    subprocess.run([nvcc, *ARCH_FLAGS, "-shared", "-Xcompiler", "-fPIC", "-o", str(lib),
                    *[str(out_dir / (s.stem + ".o")) for s in sources()]], check=True)
    return lib


if __name__ == "__main__":
    if "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], sys.argv[i + 2:]))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
