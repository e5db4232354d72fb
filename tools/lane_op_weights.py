#!/usr/bin/env python3
"""Calibrate the per-event lane-instruction weights bench.py's roofline uses.

    python tools/lane_op_weights.py <ncu source page csv> <stats counters json> [out json] [trace_kernels.cuh of the capture]

Inputs: (1) the `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` dump of ONE launch of the strict chain-skipping
trace kernel (compiled with -lineinfo): thread-level instructions executed per source line of csrc/trace_kernels.cuh; (2) the
work counters of the instrumented build for the SAME frame (Engine.debug_trace_stats).  Source lines are grouped by the event
that executes them (function markers in the source, so the grouping survives edits), and weight = thread instructions / events.
bench.py then estimates the lane-instructions a frame PERFORMED as sum(weight * counter) -- checked against ncu's
smsp__thread_inst_executed.sum of other frames and workloads in profiles/."""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def line_groups(src_lines):
    """source line -> group name, from the function each line belongs to (+ sub-ranges inside the big functions)"""
    marks = [  # (regex that starts a range, group)
        (r"void ray_prepare_fma\(", "inst_entry"),
        (r"bool slab_test\(", "slab_ref"),
        (r"bool slab_test_sub\(", "slab_tight"),
        (r"bool slab_test_ch\(", "sub_pair"),
        (r"bool mt_filter\(", "tri_filter"),
        (r"bool mt_finish\(", "tri_finish"),
        (r"void leaf_brute\(", "tri_filter"),
        (r"void leaf_accel\(", "sub_leaf_entry"),
        (r"uint32_t first = ref & 0x0FFFFFFFu;", "tri_filter"),
        (r"st.add\(6\);", "sub_pair"),
        (r"if \(lfound\) \{ best_t = lt;", "sub_leaf_entry"),
        (r"void blas_intersect\(", "blas_pair"),
        (r"HitRec scene_intersect\(", "tlas"),
        (r"uint32_t inst = __float_as_uint\(n1.w\);", "inst_entry"),
        (r"RayM primary_ray\(", "ray_gen"),
        (r"uint32_t texel_index\(", "shade"),
        (r"uint32_t tlas_skip_init\(", "block"),
        (r"^trace_primary_kernel\(", "block"),
        (r"^classify_fill_kernel\(", "other"),
    ]
    groups, cur = {}, "other"
    in_scene = False
    for i, text in enumerate(src_lines, 1):
        for rx, g in marks:
            if g and re.search(rx, text):
                cur = g
                in_scene = g in ("tlas",) or (in_scene and g == "inst_entry")
                break
        # inside scene_intersect the instance-entry block ends where the walk pops the TLAS stack again
        if in_scene and cur == "inst_entry" and re.search(r"uint32_t ni = stack\[--sp\];", text):
            cur = "tlas"
        groups[i] = cur
    return groups


def main():
    page, counters_path = sys.argv[1], sys.argv[2]
    out = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "profiles", "lane_op_weights.json")
    rows = list(csv.reader(open(page)))
    hdr = next(r for r in rows if r and r[0] == "Line No")
    i_thr = hdr.index("Thread Instructions Executed")
    i_wrp = hdr.index("Instructions Executed")
    # The page has one section per source FILE, and under every source line the SASS instructions correlated with it.  With
    # inlining one SASS instruction appears under SEVERAL lines (the callee's line and every call site up the inline stack), so
    # instructions are counted once, by address, under the callee -- the smallest line number of trace_kernels.cuh that lists
    # them (callees are defined above their callers).  Instructions only listed under toolkit headers (__ldg, shuffles) are
    # spread over the groups in proportion.
    i_addr = hdr.index("Address")
    sass, cur_line, in_kernel_file = {}, None, False       # address -> [line or None, thread instr, warp instr]
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
                t, w_ = int(r[i_thr]), int(r[i_wrp])
            except ValueError:
                continue
            e = sass.setdefault(r[i_addr], [None, t, w_])
            if in_kernel_file and cur_line is not None and (e[0] is None or cur_line < e[0]):
                e[0] = cur_line
    thr, wrp, other_thr, other_wrp = {}, {}, 0, 0
    for line, t, w_ in sass.values():
        if line is None:
            other_thr += t; other_wrp += w_
        else:
            thr[line] = thr.get(line, 0) + t; wrp[line] = wrp.get(line, 0) + w_
    cuh = sys.argv[4] if len(sys.argv) > 4 else os.path.join(ROOT, "bvhtracer_b200", "csrc", "trace_kernels.cuh")
    groups = line_groups(open(cuh).read().split("\n"))
    tot = {}
    for n, v in thr.items():
        tot[groups[n]] = tot.get(groups[n], 0) + v
    spread = 1.0 + other_thr / max(sum(tot.values()), 1)
    tot = {g: v * spread for g, v in tot.items()}
    c = json.load(open(counters_path))
    ev = {
        "ray_gen": c["rays_traced"],
        "block": c["blocks_pulled"],
        "tlas": c["tlas_pairs"] + c["tlas_chain_heads"],
        "slab_tight": c["tlas_pairs"] + c["tlas_chain_heads"],
        "inst_entry": c["instance_entries"],
        "blas_pair": c["blas_pairs"],
        "slab_ref": 2 * c["blas_pairs"] + 2 * c["tlas_pairs"] + c["tlas_chain_heads"],
        "sub_leaf_entry": c["ref_leaves"],
        "sub_pair": c["sub_pairs"],
        "tri_filter": c["sub_tris"] + c["brute_tris"],
        "tri_finish": c["mt_finishes"],
    }
    w = {g: tot.get(g, 0) / max(ev[g], 1) for g in ev}
    # fold the shared slab tests into their callers (2 per BLAS pair test, ~2 per TLAS pair test, 1 per chain head)
    weights = {
        "rays_traced": w["ray_gen"],
        "blocks_pulled": w["block"] + tot.get("other", 0) / max(c["blocks_pulled"], 1),
        "tlas_pairs": w["tlas"] + w["slab_tight"] + 2 * w["slab_ref"],
        "tlas_chain_heads": w["tlas"] + w["slab_tight"] + w["slab_ref"],
        "instance_entries": w["inst_entry"],
        "blas_pairs": w["blas_pair"] + 2 * w["slab_ref"],
        "ref_leaves": w["sub_leaf_entry"],
        "sub_pairs": w["sub_pair"],
        "sub_tris": w["tri_filter"], "brute_tris": w["tri_filter"],
        "mt_finishes": w["tri_finish"],
    }
    est = sum(weights[k] * c[k] for k in weights)
    total_thr, total_wrp = sum(thr.values()) + other_thr, sum(wrp.values()) + other_wrp
    res = {"weights_thread_instructions_per_event": weights, "calibration": {
        "ncu_thread_instructions": total_thr, "ncu_warp_instructions": total_wrp, "lanes_per_instruction": total_thr / total_wrp,
        "model_thread_instructions": est, "model_over_ncu": est / total_thr, "group_thread_instructions": tot,
        "thread_instructions_of_inlined_toolkit_headers_spread_proportionally": other_thr, "counters": c,
        "source_page": os.path.basename(page)}}
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res["calibration"], indent=1)[:1500])
    print(json.dumps(weights, indent=1))


if __name__ == "__main__":
    main()
