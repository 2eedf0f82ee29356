"""Build libmedtok_vq.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m medtok_b200.build [--force]

The .so lands next to this file (medtok_b200/libmedtok_vq.so) so that it travels with the source
tree; there is no torch dependency in the library (plain extern "C", see include/medtok_vq.h).
"""
import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ_DIR = os.path.join(HERE, "_build")
SO_PATH = os.path.join(HERE, "libmedtok_vq.so")
SOURCES = ["api.cu", "assign_tc.cu", "assign_exact.cu", "group.cu", "rows.cu", "soft.cu", "small.cu"]
PROBE_SOURCES = ["microbench.cu"]          # diagnostics: only in builds made with -DMTVQ_WITH_MICROBENCH
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC,-fvisibility=hidden", "--expt-relaxed-constexpr"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _newest_dep():
    t = 0.0
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in os.listdir(root):
            t = max(t, os.path.getmtime(os.path.join(root, f)))
    return t


def build(force: bool = False, verbose: bool = False, variant: str = "", defines=()) -> str:
    """`variant` + `defines` (-D flags) build an experimental copy into _build/variants/ for A/B timing
    (tools_ab.py, MEDTOK_VQ_LIB); the product library is the default, flag-less build."""
    so_path, obj_dir = SO_PATH, OBJ_DIR
    if variant:
        obj_dir = os.path.join(OBJ_DIR, "variants", variant)
        so_path = os.path.join(OBJ_DIR, "variants", f"libmedtok_vq_{variant}.so")
        force = True
    if not force and os.path.exists(so_path) and os.path.getmtime(so_path) >= _newest_dep():
        return so_path
    os.makedirs(obj_dir, exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src):
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

    sources = SOURCES + (PROBE_SOURCES if "-DMTVQ_WITH_MICROBENCH" in defines else [])
    with concurrent.futures.ThreadPoolExecutor(max_workers=len(sources)) as ex:
        objs = list(ex.map(compile_one, sources))
    # api.cu marks the exported symbols; everything else stays hidden
    link = [nvcc, "-shared", "-o", so_path, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
            "-Xcompiler", "-fPIC", "-lcudart_static", "-lpthread", "-ldl", "-lrt"]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return so_path


if __name__ == "__main__":
    _variant = sys.argv[sys.argv.index("--variant") + 1] if "--variant" in sys.argv else ""
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, variant=_variant,
                defines=[a for a in sys.argv[1:] if a.startswith("-D")]))
