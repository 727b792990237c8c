"""In-tree build of libpcuda.so (sm_100a only) and of the test oracle.

`python -m pointcloududa_b200.build` or `__graft_entry__.build()`.  nvcc cross-compiles without a
GPU; the resulting .so is git-ignored but travels with the repo snapshot to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
LIB = CSRC / "libpcuda.so"
ORACLE = ROOT / "oracle"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC,-O2,-Wall,-Wno-unused-function",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: libpcuda needs the CUDA 12.9 toolchain")


def _stale(target: Path, deps: list[Path]) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(d.stat().st_mtime > t for d in deps if d.exists())


def build_cuda(force: bool = False, verbose: bool = False) -> Path:
    nvcc = _nvcc()
    sources = sorted(CSRC.glob("*.cu"))
    headers = sorted(CSRC.glob("*.cuh")) + [ROOT / "include" / "pcuda.h"]
    objdir = CSRC / "build"
    objdir.mkdir(exist_ok=True)

    def compile_one(src: Path) -> Path:
        obj = objdir / (src.stem + ".o")
        if force or _stale(obj, [src] + headers):
            cmd = [nvcc, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
                print(" ".join(cmd), flush=True)
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
            if verbose:
                print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(sources))) as ex:
        objs = list(ex.map(compile_one, sources))
    if force or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-o", str(LIB), *map(str, objs), "-lcudart", "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


def build_oracle(force: bool = False) -> Path:
    """Compile the C restatement used by tests / smoke / the cpu_baseline leg (never by the product)."""
    src = ORACLE / "pcuda_oracle.c"
    out = ORACLE / "libpcuda_oracle.so"
    if force or _stale(out, [src]):
        cmd = ["gcc", "-O3", "-std=c11", "-fPIC", "-shared", "-fopenmp", "-mavx2", "-mfma",
               "-ffp-contract=off", "-fno-fast-math", "-o", str(out), str(src), "-lm"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"gcc failed for oracle:\n{r.stdout}\n{r.stderr}")
    return out


def main(argv: list[str]) -> int:
    force = "--force" in argv
    verbose = "-v" in argv
    lib = build_cuda(force=force, verbose=verbose)
    print(f"built {lib}")
    if (ORACLE / "pcuda_oracle.c").exists():
        print(f"built {build_oracle(force=force)}")
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
