"""In-tree build of the native engine: nvcc -> bayesssm_b200/libbayesssm_b200.so (sm_100a only)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libbayesssm_b200.so")
OBJ_DIR = os.path.join(HERE, "csrc", "_obj")
SOURCES = ["bssm_engine.cu", "bssm_pmmh.cu", "bssm_fast.cu", "bssm_nvrtc.cu"]
NVCC_FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found")
    return nvcc


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_native(force: bool = False, verbose: bool = False) -> str:
    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "bayesssm_b200.h"))
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    objs = [os.path.join(OBJ_DIR, s.replace(".cu", ".o")) for s in srcs]

    def compile_one(pair):
        src, obj = pair
        if not force and not _stale(obj, [os.path.join(CSRC, src)] + headers):
            return ""
        cmd = [nvcc] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        with open(obj + ".ptxas.log", "w") as fh:
            fh.write(r.stderr)
        return r.stderr

    with ThreadPoolExecutor(max_workers=4) as ex:
        logs = list(ex.map(compile_one, zip(srcs, objs)))
    if force or _stale(OUT, objs):
        cmd = [nvcc, "-shared", "-cudart", "static", "-o", OUT] + objs + ["-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if verbose:
        sys.stderr.write("\n".join(logs))
    return OUT


if __name__ == "__main__":
    print(build_native(force="--force" in sys.argv, verbose="-v" in sys.argv))
