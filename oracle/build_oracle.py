"""Compile oracle/vq_oracle.c -> oracle/_build/libvq_oracle.so (test infrastructure only).

gcc flags matter for parity: -ffp-contract=off keeps every non-FMA multiply/add a separate
rounding (the reference's ATen reductions do not fuse them); the explicit fmaf() calls map to
hardware FMA with -mfma (glibc's software fmaf is also correctly rounded, just slow).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_build")
SO_PATH = os.path.join(OUT_DIR, "libvq_oracle.so")


def _cpu_flags():
    try:
        with open("/proc/cpuinfo") as f:
            txt = f.read()
    except OSError:
        return set()
    for line in txt.splitlines():
        if line.startswith("flags"):
            return set(line.split(":", 1)[1].split())
    return set()


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "vq_oracle.c")
    if (not force and os.path.exists(SO_PATH)
            and os.path.getmtime(SO_PATH) >= os.path.getmtime(src)):
        return SO_PATH
    os.makedirs(OUT_DIR, exist_ok=True)
    # The .so travels to the GPU box prebuilt, so do not use -march=native: stick to
    # FMA + AVX2, which every x86 host that can drive a B200 has.
    flags = ["-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared",
             "-fvisibility=hidden", "-fopenmp"]
    cpu = _cpu_flags()
    if not cpu or ("fma" in cpu and "avx2" in cpu):
        flags += ["-mfma", "-mavx2", "-ftree-vectorize"]
    cmd = ["gcc", *flags, "-o", SO_PATH, src, "-lm"]
    subprocess.run(cmd, check=True)
    return SO_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
