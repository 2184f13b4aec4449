#!/usr/bin/env python
"""Build A/B variants of the library that differ only in compile-time knobs of csrc/uvd.cu (Gram sweep plan), into
psgd_tf_b200/_C/variants/libpsgd_b200_<name>.so, for one-call comparisons on the GPU box:

    python tools/uvd_variants.py                       # here (CPU container; nvcc cross-compiles)
    PSGD_B200_LIB=psgd_tf_b200/_C/variants/libpsgd_b200_<name>.so python bench.py --workload uvd ...   # on the box
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from psgd_tf_b200 import build as B   # noqa: E402

VARIANTS = {
    "rpl2": ["-DPSGD_GRAM_RPL=2"],
    "rpl4": ["-DPSGD_GRAM_RPL=4"],
    "map2": ["-DPSGD_MAP_RPL=2"],
}


def main():
    B.build()
    vdir = os.path.join(B.OUT_DIR, "variants")
    os.makedirs(vdir, exist_ok=True)
    others = [os.path.join(B.OUT_DIR, s.replace(".cu", ".o")) for s in B.SOURCES if s != "uvd.cu"]

    def one(item):
        name, flags = item
        obj = os.path.join(vdir, f"uvd_{name}.o")
        lib = os.path.join(vdir, f"libpsgd_b200_{name}.so")
        r = subprocess.run([B._nvcc(), *B.NVCC_FLAGS, *flags, "-Xptxas", "-v", "-c", os.path.join(B.CSRC, "uvd.cu"), "-o", obj],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(r.stderr[-3000:])
        lines = r.stderr.splitlines()
        info = [lines[i + 2] for i, l in enumerate(lines) if "gram_sweep_kernelILi10ELi0" in l and i + 2 < len(lines)]
        subprocess.run([B._nvcc(), "-shared", "-o", lib, obj, *others, "-gencode", "arch=compute_100a,code=sm_100a"], check=True)
        os.remove(obj)
        return name, info

    names = sys.argv[1:] or list(VARIANTS)
    with ThreadPoolExecutor(max_workers=4) as ex:
        for name, info in ex.map(one, [(n, VARIANTS[n]) for n in names]):
            print(name, info)


if __name__ == "__main__":
    main()
