"""Build libqasr_b200.so in-tree with nvcc for sm_100a (no torch, no pybind: a plain C-ABI library).

    python -m qwen3_asr_b200.build [--force]

The objects go to qwen3_asr_b200/csrc/build/, the library to qwen3_asr_b200/libqasr_b200.so (git-ignored;
it travels to the GPU box with the gpurun snapshot).
"""

from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ_DIR = os.path.join(CSRC, "build")
LIB_PATH = os.path.join(HERE, "libqasr_b200.so")
SOURCES = ["qasr.cu", "gemm.cu", "mel.cu", "elementwise.cu", "quant.cu", "attention_tc.cu", "prefrontend.cu", "pool.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _deps() -> list[str]:
    out = [os.path.join(HERE, "..", "include", "qasr_b200.h"), os.path.abspath(__file__)]
    for f in os.listdir(CSRC):
        if f.endswith((".cuh", ".h", ".cu")):
            out.append(os.path.join(CSRC, f))
    return out


def is_fresh() -> bool:
    if not os.path.exists(LIB_PATH):
        return False
    t = os.path.getmtime(LIB_PATH)
    return all(os.path.getmtime(d) <= t for d in _deps() if os.path.exists(d))


def build(force: bool = False, verbose: bool = False, variant: str | None = None, defines: tuple[str, ...] = ()) -> str:
    """Default: the product library.  `variant` + `defines` build an A/B copy (libqasr_<variant>.so) with extra -D flags;
    tools/ab.sh selects it through QASR_B200_LIB."""
    lib_path = LIB_PATH if variant is None else os.path.join(HERE, f"libqasr_{variant}.so")
    obj_dir = OBJ_DIR if variant is None else os.path.join(OBJ_DIR, variant)
    if variant is None and not force and is_fresh():
        return LIB_PATH
    nvcc = _nvcc()
    os.makedirs(obj_dir, exist_ok=True)

    def compile_one(src: str) -> str:
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *defines, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "--cudart", "static", "-o", lib_path, *objs, "-Xlinker", "--no-undefined", "-lpthread", "-ldl", "-lrt"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return lib_path


if __name__ == "__main__":
    # python -m qwen3_asr_b200.build [--force] [-v] [--variant NAME -DX=1 ...]
    var = sys.argv[sys.argv.index("--variant") + 1] if "--variant" in sys.argv else None
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv, variant=var, defines=tuple(a for a in sys.argv if a.startswith("-D")))
    print(path)
