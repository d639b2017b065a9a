"""Per-object contact statistics (DataSet.statistics_*; contact.cu:379-430, main.cu:825, 1933-1936) on the CPU oracle: the
counts fetch() brings back are those of the LAST contact assembly of the step -- a contact between two objects counts once for
each, a contact inside one object once; analytic colliders and the static collision mesh count for the dynamic side (and for
the static object)."""
import os

import numpy as np

from conftest import GOLDEN


def test_oracle_counts_contacts_per_object(run_job):
    state = os.path.join(GOLDEN, "headless_n24", "state_00068.bin")         # five curtains (objects 0..4) pressed by the moving sphere (object 5)
    r = run_job({"lib": "oracle", "scene": "headless", "n": 24, "state": state, "steps": 1, "statistics": True, "detect": "final"}, force_subprocess=True)
    c = r["statistics"].astype(np.int64)
    assert c.shape == (6,) and c[5] > 20 and c[:5].sum() >= c[5]            # every sphere contact also counts for a curtain
    last = int(r["log_num_contact"][-1])
    # each contact adds 1 (same object) or 2 (two objects): sum of the counts lies between the contact count and twice it
    assert last <= c.sum() <= 2 * last


def test_oracle_counts_collider_and_static_mesh_contacts(run_job):
    a = run_job({"lib": "oracle", "scene": "drape", "n": 32, "state": os.path.join(GOLDEN, "drape_n32", "state_00140.bin"), "steps": 1, "statistics": True}, force_subprocess=True)
    assert a["statistics"].shape == (1,) and int(a["statistics"][0]) == int(a["log_num_contact"][-1]) > 10      # analytic contacts: the one dynamic object
    m = run_job({"lib": "oracle", "scene": "drape-mesh", "n": 24, "state": os.path.join(GOLDEN, "drape_mesh_n24", "state_00040.bin"), "steps": 1, "statistics": True},
                force_subprocess=True)
    c = m["statistics"].astype(np.int64)
    assert c.shape == (2,) and c[1] > 10 and c[0] >= c[1]                    # static object last; every static contact also counts for the sheet
