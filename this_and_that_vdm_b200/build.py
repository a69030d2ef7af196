"""In-tree build of libttvdm_sm100.so (hand-written sm_100a kernels + C ABI, see include/ttvdm.h).

nvcc cross-compiles without a GPU, so this runs in the authoring container and the resulting .so travels to the
GPU box with the repo snapshot. Objects are cached under this_and_that_vdm_b200/csrc/_build/ keyed by mtime.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
BUILD = CSRC / "_build"
LIB_PATH = PKG_DIR / "libttvdm_sm100.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]
# debug builds only (e.g. TTVDM_EXTRA_NVCC_FLAGS=-DTTVDM_ATTN_TRACE for tools/attn_trace.py); use with --force
NVCC_FLAGS += os.environ.get("TTVDM_EXTRA_NVCC_FLAGS", "").split()


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _needs(obj: Path, deps: list[Path]) -> bool:
    if not obj.exists():
        return True
    t = obj.stat().st_mtime
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    BUILD.mkdir(parents=True, exist_ok=True)
    sources = sorted(CSRC.glob("*.cu"))
    headers = sorted(CSRC.glob("*.h")) + sorted(CSRC.glob("*.cuh")) + [PKG_DIR.parent / "include" / "ttvdm.h"]
    nvcc = _nvcc()

    def compile_one(src: Path) -> Path:
        obj = BUILD / (src.stem + ".o")
        if force or _needs(obj, [src] + headers):
            cmd = [nvcc, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
            if verbose:
                print(" ".join(cmd), flush=True)
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(sources) or 1)) as ex:
        objs = list(ex.map(compile_one, sources))
    if force or _needs(LIB_PATH, objs):
        cmd = [nvcc, "-shared", "-o", str(LIB_PATH), *map(str, objs)]
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB_PATH


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose=True)
    print(p)
