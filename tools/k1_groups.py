#!/usr/bin/env python3
"""Where K1's issue slots and lanes go, by the event that executes the instructions.

    python tools/k1_groups.py <ncu source page csv (ncu -i X.ncu-rep --page source --csv --print-source cuda,sass)> [trace_kernels.cuh]

SASS instructions are counted once (by address) under the smallest source line of trace_kernels.cuh that lists them (the callee,
tools/lane_op_weights.py) and grouped like the weights bench.py uses.  Per group: share of the warp-level instructions (issue
slots), share of the thread-level instructions (lane work), active lanes per instruction, share of the warp-state samples."""
import csv
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from lane_op_weights import ROOT, line_groups

NAMES = {"sub_pair": "sub-BVH pair step (two child boxes + ordering + push)", "blas_pair": "reference BLAS pair step", "slab_ref": "reference slab test (Aabb::intersect)",
         "slab_tight": "tight-box slab test (TLAS)", "inst_entry": "instance entry (world -> model ray, Ray::new)", "tri_filter": "triangle filter (Moeller-Trumbore up to the u/v/det tests)",
         "tri_finish": "triangle finish (t, record update)", "sub_leaf_entry": "reference leaf entry -> sub-BVH root", "tlas": "TLAS walk", "ray_gen": "primary ray generation",
         "block": "per-block work (pull, candidate masks, skip table, band flags)", "shade": "pixel shading", "other": "other / toolkit headers"}


def main():
    page = sys.argv[1]
    cuh = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "bvhtracer_b200", "csrc", "trace_kernels.cuh")
    rows = list(csv.reader(open(page)))
    hdr = next(r for r in rows if r and r[0] == "Line No")
    i_thr, i_wrp, i_addr = hdr.index("Thread Instructions Executed"), hdr.index("Instructions Executed"), hdr.index("Address")
    i_smp = hdr.index("# Samples")
    sass, cur_line, in_kernel_file = {}, None, False
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            in_kernel_file = r[1].endswith("trace_kernels.cuh")
            cur_line = None
        elif r[0].isdigit():
            cur_line = int(r[0])
        elif r[0] == "" and len(r) > i_thr and r[i_addr].startswith("0x"):
            try:
                t, w_, s_ = int(r[i_thr]), int(r[i_wrp]), int(r[i_smp] or 0)
            except ValueError:
                continue
            e = sass.setdefault(r[i_addr], [None, t, w_, s_])
            if in_kernel_file and cur_line is not None and (e[0] is None or cur_line < e[0]):
                e[0] = cur_line
    groups = line_groups(open(cuh).read().split("\n"))
    agg = {}
    for line, t, w_, s_ in sass.values():
        g = groups.get(line, "other") if line is not None else "other"
        a = agg.setdefault(g, [0, 0, 0, 0])
        a[0] += t; a[1] += w_; a[2] += s_; a[3] += 1
    T, W, S = (sum(a[i] for a in agg.values()) for i in range(3))
    print(f"# {os.path.basename(page)}: {len(sass)} SASS instructions, {W:.4g} warp-level / {T:.4g} thread-level instructions executed, "
          f"{T / W:.2f} lanes per instruction, {S} warp-state samples")
    print(f"{'group':72s} {'SASS':>5s} {'issue slots':>12s} {'lane work':>10s} {'lanes/instr':>12s} {'samples':>8s}")
    for g, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{NAMES.get(g, g):72s} {a[3]:5d} {100 * a[1] / W:11.1f}% {100 * a[0] / T:9.1f}% {a[0] / max(a[1], 1):12.2f} {100 * a[2] / max(S, 1):7.1f}%")


if __name__ == "__main__":
    main()
