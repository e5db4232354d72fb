#!/usr/bin/env python3
"""L2 and HBM read bandwidth of this box with the library's own streaming kernel (SURVEY.md 8d ii): python tools/bw_probe.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bvhtracer_b200 import Engine

with Engine() as eng:
    for mib, passes in ((8, 400), (16, 200), (32, 100), (64, 50), (96, 40), (256, 10), (2048, 3)):
        g = max(eng.debug_read_bandwidth(mib << 20, passes) for _ in range(3))
        print(f"{mib:5d} MiB x {passes:3d} passes: {g:9.1f} GB/s", flush=True)
