"""Positive path of the intersection check on the CPU oracle (contact.cu:3056-3175, main.cu:247-263, 2341): a start state
that is already intersecting makes initialize() return false with fatal code 1 and leaves records that name the elements."""
import os
import sys

import numpy as np

from conftest import REPO

sys.path.insert(0, os.path.join(REPO, "tests"))
import pierce  # noqa: E402


def test_oracle_reports_a_pierced_curtain(ppf, tmp_path):
    path = str(tmp_path / "pierced.bin")
    v, x, y, face, edge = pierce.pierced_curtain(ppf, ppf.LIB_ORACLE, path)
    s = ppf.Session("headless", n=24, lib=ppf.LIB_ORACLE, state=path, allow_init_failure=True)
    assert not s.init_ok and s.fatal_code() == 1
    rec = s.intersection_records()
    assert 1 <= len(rec) <= 12
    touched = set()
    for t, f, e, pos in rec:
        assert t == 0                                                # face-edge
        p = np.array(pos, np.float32).reshape(5, 3)
        assert np.array_equal(p[:3], y[face[f]]) and np.array_equal(p[3:], y[edge[e]])   # the elements' own positions, verbatim
        assert pierce.segment_pierces_triangle(p[3], p[4], p[0], p[1], p[2])
        assert v in edge[e] or v in face[f]                          # every pierce involves the displaced vertex
        touched.add(int(e))
    assert any(v in edge[e] for e in touched)
    # the unmodified start state passes, and the records are gone
    ok = ppf.Session("headless", n=24, lib=ppf.LIB_ORACLE)
    assert ok.init_ok and ok.intersection_records() == []


def test_oracle_reports_rods_inside_their_contact_offsets(ppf, tmp_path):
    path = str(tmp_path / "crowded.bin")
    i, j, off, y, edge = pierce.crowded_rods(ppf, ppf.LIB_ORACLE, path)
    s = ppf.Session("woven", n=32, lib=ppf.LIB_ORACLE, state=path, allow_init_failure=True)
    assert not s.init_ok and s.fatal_code() == 1
    rec = s.intersection_records()
    kinds = {r[0] for r in rec}
    assert kinds <= {1, 3} and 3 in kinds and 1 in kinds
    pp = [r for r in rec if r[0] == 3]
    assert (3, max(i, j), min(i, j)) in {(r[0], r[1], r[2]) for r in pp}              # query vertex first, smaller index second
    for t, a, b, pos in pp:
        p = np.array(pos, np.float64).reshape(2, 3)
        assert np.linalg.norm(p[0] - p[1]) < off
    for t, a, b, pos in rec:
        if t == 1:
            assert b < a and (i in edge[a] or i in edge[b])
