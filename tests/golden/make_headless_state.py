#!/usr/bin/env python3
"""make_headless_state.py -- a contact-rich state of the `headless` scene (examples/headless.py at res 24) for the tests.

Unlike the other directories under tests/golden/, this one holds NO output of the reference: the state is produced by
the CPU oracle (68 steps from rest: the pinned sphere has pushed into the first three curtains) and only serves as a
common starting point for the oracle-vs-backend comparison of tests/test_zz_gpu_headless.py and the CPU test in
tests/test_host.py.  Runs in seconds on the CPU:   python tests/golden/make_headless_state.py
"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(REPO, "tests"))
import _worker  # noqa: E402

out = os.path.join(REPO, "tests", "golden", "headless_n24")
os.makedirs(out, exist_ok=True)
r = _worker.run({"lib": "oracle", "scene": "headless", "n": 24, "state": None, "steps": 68, "save_state": os.path.join(out, "state_00068.bin"),
                 "out": "/tmp/ppf_headless_state"})
assert bool(r["ok"].all())
print("wrote", os.path.join(out, "state_00068.bin"), "contacts at the last Newton iteration:", int(r["log_num_contact"][-1]))
