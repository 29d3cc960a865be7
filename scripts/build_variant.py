#!/usr/bin/env python
"""Build a variant of the engine library with extra nvcc flags (usually -D experiment macros) next to the in-tree one:

    python scripts/build_variant.py ranks8 -DJEN1_RANKS_IN_FLIGHT=8
    JEN1_B200_LIB=jen1_b200/_C/variants/ranks8/libjen1_b200.so python bench.py --quick ...

The variants live under jen1_b200/_C/ (git-ignored, travels to the GPU box with the snapshot).
"""
import concurrent.futures
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jen1_b200 import build as B  # noqa: E402


def main(name, *flags):
    out = os.path.join(B.OUT_DIR, "variants", name)
    os.makedirs(out, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(B.CSRC, "*.cu")))

    def cc(src):
        obj = os.path.join(out, os.path.basename(src)[:-3] + ".o")
        r = subprocess.run([B.NVCC] + B.FLAGS + list(flags) + ["-c", src, "-o", obj], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(r.stdout + r.stderr)
        return obj

    with concurrent.futures.ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(cc, srcs))
    lib = os.path.join(out, "libjen1_b200.so")
    subprocess.run([B.NVCC, "-shared", "-o", lib] + objs + ["-lcudart"], check=True)
    print("built", lib)


if __name__ == "__main__":
    main(*sys.argv[1:])
