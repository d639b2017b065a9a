"""Shared fixtures.  Tests that need a B200 are marked @pytest.mark.gpu; everything else runs on CPU."""
import os
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(REPO, "ppf-contact-solver_b200")
sys.path.insert(0, PKG)
sys.path.insert(0, REPO)
GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """Native artefacts: the CUDA backend, the C++ host and the CPU oracle (all built in-tree)."""
    need = [os.path.join(PKG, "lib", "libsimbackend_cuda.so"), os.path.join(PKG, "lib", "libppf_host.so"),
            os.path.join(REPO, "oracle", "build", "libppf_oracle.so"), os.path.join(REPO, "oracle", "build", "liboracle_math.so")]
    if not all(os.path.exists(p) for p in need):
        import __graft_entry__
        __graft_entry__.build()
    for p in need:
        assert os.path.exists(p), "missing native artefact %s" % p
    return True


@pytest.fixture(scope="session")
def ppf(built):
    import ppf_b200
    return ppf_b200


_USED_IN_PROCESS = set()


@pytest.fixture(scope="session")
def run_job(built, tmp_path_factory):
    """Runs a backend job (tests/_worker.py).  The first job on each library runs inside the pytest
    process (so the CUDA .so is loaded by pytest itself); later ones get a fresh process because the
    backends keep global state, like the reference's."""
    import json

    import numpy as np

    sys.path.insert(0, os.path.join(REPO, "tests"))
    import _worker

    def go(job, force_subprocess=False):
        job = dict(job)
        job["out"] = str(tmp_path_factory.mktemp("job"))
        key = job["lib"]
        if key not in _USED_IN_PROCESS and not force_subprocess:
            _USED_IN_PROCESS.add(key)
            res = _worker.run(job)
            res = {k: np.asarray(v) for k, v in res.items()}
        else:
            jp = os.path.join(job["out"], "job.json")
            os.makedirs(job["out"], exist_ok=True)
            json.dump(job, open(jp, "w"))
            r = subprocess.run([sys.executable, os.path.join(REPO, "tests", "_worker.py"), jp], capture_output=True, text=True, timeout=1500)
            assert r.returncode == 0, "worker failed:\n%s\n%s" % (r.stdout[-2000:], r.stderr[-4000:])
            res = dict(np.load(os.path.join(job["out"], "result.npz"), allow_pickle=False))
        res["out_dir"] = job["out"]
        return res

    return go


def golden_cases():
    import json

    out = []
    # second line: reference golden for the features written late in round 1 (make_golden.sh next-round, run on a B200 in round 2)
    for name in ("fast_twist_n16", "fast_twist_n24", "five_twist_n24", "hang_n32", "woven_n32", "woven_sl_n32", "drape_n32",
                 "headless_ref_n24", "lift_n24", "sheet_arap_n24", "sheet_stvk_n24", "sheet_snhk_n24", "sheet_damped_n24", "woven_damped_n32",
                 "seam_n16", "sheet_plastic_n24", "sheet_fixxz_n24", "drape_mesh_n24", "sheet_torque_n24", "balloon_n24",
                 "fast_twist_quad_n16", "fast_twist_log_n16"):
        if not os.path.exists(os.path.join(GOLDEN, name, "meta.json")):
            continue
        d = os.path.join(GOLDEN, name)
        meta = json.load(open(os.path.join(d, "meta.json")))
        for c in meta["cases"]:
            out.append((name, meta["scene"], meta["n"], os.path.join(d, c["state"]), os.path.join(d, c["dir"])))
    return out
