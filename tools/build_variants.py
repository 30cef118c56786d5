#!/usr/bin/env python3
"""Build-time variants of the tiled decode kernel for A/B runs: nl_tile.cu is recompiled with the given -D flags and linked with the
regular objects into nanollama_b200/build/variants/lib_<name>.so; NL_LIB=<path> makes nanollama_b200.capi load that library.

    python tools/build_variants.py base: xb2:-DNL_TL_XB_SINGLE=0 tr:-DNL_TL_FINE_TRACE=1
    python tools/decode_ab.py --variants "NL_LIB=nanollama_b200/build/variants/lib_base.so;NL_LIB=nanollama_b200/build/variants/lib_xb2.so"
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nanollama_b200 import build as B


def main():
    B.build()
    vdir = os.path.join(B.OBJ, "variants")
    os.makedirs(vdir, exist_ok=True)
    objs = [os.path.join(B.OBJ, s[:-3] + ".o") for s in B._sources() if s != "nl_tile.cu"]
    for spec in sys.argv[1:]:
        name, _, flags = spec.partition(":")
        obj = os.path.join(vdir, f"nl_tile_{name}.o")
        r = subprocess.run([B.nvcc(), *B.NVCC_FLAGS, *[f for f in flags.split(",") if f], "-c", os.path.join(B.CSRC, "nl_tile.cu"), "-o", obj], capture_output=True, text=True)
        if r.returncode:
            raise SystemExit(r.stderr)
        lib = os.path.join(vdir, f"lib_{name}.so")
        r2 = subprocess.run([B.nvcc(), "-shared", "-o", lib, *objs, obj, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static", "-Xlinker", "--no-undefined"],
                            capture_output=True, text=True)
        if r2.returncode:
            raise SystemExit(r2.stderr)
        info = []
        lines = r.stderr.splitlines()
        for i, l in enumerate(lines):
            m = re.search(r"Function properties for (\S+)", l)
            if m and ("ILi2E" in m.group(1)) and any(k in m.group(1) for k in ("decode_tiled", "stream_band", "input_frags")):
                fn = "tiled" if "decode_tiled" in m.group(1) else ("stream_band" if "stream_band" in m.group(1) else ("frags_exch" if "exch" in m.group(1) else "frags"))
                info.append(f"{fn}: {lines[i + 1].strip()}")
        print(f"[{name}] {flags or '(default)'} -> {os.path.relpath(lib, ROOT)}\n    " + "\n    ".join(info))


if __name__ == "__main__":
    main()
