#!/usr/bin/env python3
"""A/B builds of libmcq.so: recompiles the listed sources with extra -D flags and links them with the regular
objects into quantization_b200/libmcq_<name>.so (git-ignored; travels to the GPU box).  Select at run time with
MCQ_LIB=quantization_b200/libmcq_<name>.so.

    python tools/build_variant.py nopair search2.cu -DMCQ_MERGE_PAIR=0
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quantization_b200 import build as b  # noqa: E402


def main():
    name = sys.argv[1]
    srcs = [a for a in sys.argv[2:] if a.endswith(".cu")]
    flags = [a for a in sys.argv[2:] if not a.endswith(".cu")]
    b.build()
    objs = []
    for src in b.SOURCES:
        obj = os.path.join(b.BUILD, src.replace(".cu", ".o"))
        if src in srcs:
            obj = os.path.join(b.BUILD, f"{name}_{src.replace('.cu', '.o')}")
            cmd = [b.NVCC] + b.FLAGS + flags + ["-c", os.path.join(b.CSRC, src), "-o", obj]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise SystemExit(r.stderr)
        objs.append(obj)
    lib = os.path.join(b.HERE, f"libmcq_{name}.so")
    cmd = [b.NVCC, "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart_static",
                                                   "-ldl", "-lpthread", "-lrt"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise SystemExit(r.stderr)
    print(lib)


if __name__ == "__main__":
    main()
