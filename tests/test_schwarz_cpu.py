"""CPU twin (oracle/ppf_oracle.cpp, MasHierarchy / MasLevels) of the multilevel additive aggregation preconditioner that the CUDA
backend runs when ParamSet.precond = Schwarz (csrc/pcg.cu).  It is this backend's own construction -- nested aggregates of 16
shell vertices in Morton order inside every connected component, block-diagonal Galerkin operators on every level -- so what
is pinned here is its defining behaviour:
  * a fixed SPD operator: PCG converges to the block-Jacobi solution within what two solves at cg_tol = 1e-3 differ by;
  * the long solves get shorter (that is what coarse levels are for), while a solve block-Jacobi finishes in a few iterations
    may take a few more (additive levels bound the condition number from below by ~ the number of levels);
  * pinned vertices stay exactly on their targets (they are left out of every aggregate);
  * rod scenes have no aggregates at all: the Schwarz switch then changes nothing, bit for bit."""
import numpy as np
import pytest


@pytest.mark.parametrize("scene,n", [("twist", 96), ("large-twist", 200)])
def test_oracle_schwarz_mode_shortens_the_long_solves(run_job, scene, n):
    job = {"lib": "oracle", "scene": scene, "n": n, "state": None, "steps": 1, "fix_positions": True}
    a = run_job(dict(job), force_subprocess=True)
    b = run_job(dict(job, param_flags={"precond": 1}), force_subprocess=True)
    assert bool(a["ok"].all()) and bool(b["ok"].all()) and len(a["log_iter"]) == len(b["log_iter"])
    assert a["log_iter"].max() > 100 and b["log_iter"].sum() < 0.85 * a["log_iter"].sum(), (list(a["log_iter"]), list(b["log_iter"]))
    assert float(np.abs(a["curr"] - b["curr"]).max()) < 2e-3
    idx = b["fix_index"].astype(np.int64)
    assert idx.size > 0 and np.array_equal(b["curr"][idx], b["fix_target"])


def test_oracle_schwarz_mode_costs_little_on_easy_systems_and_nothing_on_rods(run_job):
    job = {"lib": "oracle", "scene": "hang", "n": 32, "state": None, "steps": 2}
    a = run_job(dict(job), force_subprocess=True)
    b = run_job(dict(job, param_flags={"precond": 1}), force_subprocess=True)
    assert a["log_iter"].max() <= 5 and b["log_iter"].max() <= a["log_iter"].max() + 8     # mass-dominated (dt = 1e-3): kappa <= levels + 1
    assert float(np.abs(a["curr"] - b["curr"]).max()) < 1e-4
    job = {"lib": "oracle", "scene": "woven", "n": 32, "state": None, "steps": 2}
    a = run_job(dict(job), force_subprocess=True)
    b = run_job(dict(job, param_flags={"precond": 1}), force_subprocess=True)
    assert np.array_equal(a["log_iter"], b["log_iter"]) and np.array_equal(a["curr"], b["curr"])
