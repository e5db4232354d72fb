"""tools/compare_dump.py is what closes oracle <-> Rust-reference parity wherever cargo exists (rust/bvhtracer/examples/dump_hits.rs
writes the files).  Here: its file format and plumbing, oracle against oracle, on two small cases -- and that a corrupted
record is reported."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOL = os.path.join(ROOT, "tools", "compare_dump.py")


def test_dump_round_trip_and_detection(tmp_path):
    d = str(tmp_path)
    cases = ["cube:0:96x96", "sixteen_armadillos:1:96x54"]
    r = subprocess.run([sys.executable, TOOL, "--write-oracle", d, *cases], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    files = sorted(os.listdir(d))
    assert files == ["cube_f0_96x96.hits", "sixteen_armadillos_f1_96x54.hits"]
    assert os.path.getsize(os.path.join(d, files[0])) == 96 * 96 * 16
    r = subprocess.run([sys.executable, TOOL, d, "--against", "oracle"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "TOTAL differing records: 0" in r.stdout, r.stdout + r.stderr[-2000:]
    # flip one record: the tool must notice and exit 1
    p = os.path.join(d, files[0])
    a = np.fromfile(p, dtype=np.uint32)
    a[4 * 5000 + 3] ^= 1
    a.tofile(p)
    r = subprocess.run([sys.executable, TOOL, d, "--against", "oracle"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 1 and "TOTAL differing records: 1" in r.stdout, r.stdout


def test_rust_dump_example_lists_the_same_default_cases():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import importlib
    cd = importlib.import_module("compare_dump")
    rs = open(os.path.join(ROOT, "rust", "bvhtracer", "examples", "dump_hits.rs")).read()
    for case in cd.DEFAULT_CASES:
        assert f'"{case}"' in rs, case
