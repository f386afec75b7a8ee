"""Build libmsd_b200.so in-tree with nvcc for sm_100a (no JIT cache, so the .so travels with the repo).

    python multi-spectral-satellite-image-denoising_b200/build.py [--force] [--verbose]

Each csrc/*.cu is compiled to build/<name>.o only when it (or a header) is newer than the object, then linked
with the static CUDA runtime.  libcuda is NOT linked: the one driver entry point the library needs
(cuTensorMapEncodeTiled) is resolved at run time through cudaGetDriverEntryPoint, so the library also loads on a
machine without a GPU driver (the CPU-side ABI test relies on that).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
BUILD = PKG_DIR / "build"
LIB = PKG_DIR / "libmsd_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-std=c++17", "-O3", "-lineinfo",
    "--expt-relaxed-constexpr", "--extended-lambda",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found; libmsd_b200 cannot be built")


def _stale(obj: Path, deps: list[Path]) -> bool:
    if not obj.exists():
        return True
    t = obj.stat().st_mtime
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    nvcc = _nvcc()
    BUILD.mkdir(exist_ok=True)
    headers = sorted(CSRC.glob("*.cuh")) + sorted((PKG_DIR.parent / "include").glob("*.h"))
    sources = sorted(CSRC.glob("*.cu"))
    if not sources:
        raise RuntimeError(f"no CUDA sources under {CSRC}")

    def compile_one(src: Path) -> tuple[Path, str]:
        obj = BUILD / (src.stem + ".o")
        log = ""
        if force or _stale(obj, [src, Path(__file__)] + headers):
            cmd = [nvcc, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
            r = subprocess.run(cmd, capture_output=True, text=True)
            log = r.stdout + r.stderr
            (BUILD / (src.stem + ".ptxas.log")).write_text(log)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {src.name}:\n{log}")
        return obj, log

    with ThreadPoolExecutor(max_workers=min(8, len(sources))) as ex:
        results = list(ex.map(compile_one, sources))
    objs = [o for o, _ in results]
    if verbose:
        for _, log in results:
            if log:
                print(log)
    if force or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB), *map(str, objs)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}{r.stderr}")
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(p)
