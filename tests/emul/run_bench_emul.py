#!/usr/bin/env python3
"""Runs bench.py's own main() against the cuemu build on a tiny mesh — a syntax/plumbing check of
the measurement script (argument handling, e2e path, JSON assembly) in a container without a GPU.
The numbers it prints are meaningless (host emulation) and must never be reported.  TEST
INFRASTRUCTURE ONLY.  usage: run_bench_emul.py [bench.py arguments]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "cfd-lite_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import conftest  # noqa: E402

conftest.use_emulated_library()
import bench  # noqa: E402

sys.argv = ["bench.py"] + sys.argv[1:]
bench.main()
