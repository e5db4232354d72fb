#!/usr/bin/env python3
"""Tuning helper: cross-compile several builds of the library with different compile-time knobs (here, no GPU needed), so
that ONE gpurun call can time them all (tools/variant_bench.sh).

    python tools/build_variants.py name:KNOB=v,KNOB=v  name2:...
"""
import os
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bvhtracer_b200 import build as B  # noqa: E402


def one(spec):
    name, _, defs = spec.partition(":")
    defs = [d for d in defs.split(",") if d]
    return name, B.build(variant=name, defs=defs)


if __name__ == "__main__":
    with ThreadPoolExecutor(4) as ex:
        for name, path in ex.map(one, sys.argv[1:]):
            print(name, path, flush=True)
