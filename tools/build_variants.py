#!/usr/bin/env python3
"""Build variant libraries (acropolis_b200/variants/lib_<tag>.so) for kernel A/B runs on the GPU box:
    python tools/build_variants.py tag1:-DX=1,-DY=2 tag2:-DZ=3 ...
and time them there with  ACROPOLIS_B200_LIB=$PWD/acropolis_b200/variants/lib_<tag>.so python bench.py --no-cpu ..."""
import os
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from acropolis_b200.build import build_library   # noqa: E402


def one(arg):
    tag, _, defs = arg.partition(":")
    out = os.path.join(ROOT, "acropolis_b200", "variants", "lib_%s.so" % tag)
    os.makedirs(os.path.dirname(out), exist_ok=True)
    build_library(force=True, out=out, defines=tuple(d for d in defs.split(",") if d))
    return out


if __name__ == "__main__":
    with ThreadPoolExecutor(4) as ex:
        for o in ex.map(one, sys.argv[1:]):
            print(o)
