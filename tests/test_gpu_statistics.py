"""Per-object contact statistics on a real B200 (the last third of fetch()'s contract: main.cu:1926-1937; recorded at
contact.cu:379-430, 617, 722, 800, 879, 956, 1051, 1146, 1250-1334; reset per assembly at main.cu:825) against the LIVE reference
build and the CPU oracle, on three golden states: moving sphere mesh against curtains (six objects), a sheet on analytic colliders,
a sheet on a static collision mesh.  Bar: the per-object counts after one step equal the reference's up to the tie-decided
pairs of the last assembly (|delta| <= 2 + 1 %)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu
CASES = [("headless", 24, os.path.join(GOLDEN, "headless_n24", "state_00068.bin"), 6),
         ("drape", 32, os.path.join(GOLDEN, "drape_n32", "state_00140.bin"), 1),
         ("drape-mesh", 24, os.path.join(GOLDEN, "drape_mesh_n24", "state_00040.bin"), 2)]


@pytest.mark.parametrize("scene,n,state,n_obj", CASES, ids=[c[0] for c in CASES])
def test_contact_statistics_match_the_reference(run_job, ppf, scene, n, state, n_obj):
    job = {"scene": scene, "n": n, "state": state, "steps": 1, "statistics": True}
    libs = ["ours", "oracle"] + (["reference"] if os.path.exists(ppf.LIB_REF) else [])
    r = {lib: run_job(dict(job, lib=lib), force_subprocess=True) for lib in libs}
    g = r["ours"]["statistics"].astype(np.int64)
    assert g.shape == (n_obj,) and g.sum() > 10
    for lib, o in r.items():
        assert bool(o["ok"].all())
        c = o["statistics"].astype(np.int64)
        assert c.shape == g.shape and (np.abs(c - g) <= 2 + 0.01 * np.maximum(c, g)).all(), (lib, g.tolist(), c.tolist())
