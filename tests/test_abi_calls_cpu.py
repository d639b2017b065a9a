"""The per-step entry points a host may call between steps -- override_velocity, override_angular_velocity,
gather_current_positions, update_rest_shape (main.cu:2079-2228; S/backend.rs velocity keyframes and streamed rest shapes) --
on the CPU oracle: the state right after the calls is the closed form of the reference's kernels, and the step that follows
moves the way the injected velocity / the new rest shape says."""
import numpy as np

IDX = list(range(40, 120))


def test_velocity_overrides_rewrite_prev_exactly_and_drive_the_next_step(run_job):
    dt = 0.01
    v, w, c = [0.3, -0.2, 0.1], [0.0, 0.0, 2.0], [0.1, 0.2, 0.0]
    job = {"lib": "oracle", "scene": "sheet", "n": 16, "state": None, "steps": 1,
           "abi_calls": [["gather", IDX], ["override_velocity", IDX, v, dt], ["override_angular_velocity", IDX, w, c, dt]]}
    r = run_job(job, force_subprocess=True)
    x, p = r["after_calls_curr"], r["after_calls_prev"]
    assert np.array_equal(r["gather0"], x[IDX])                                       # gather = the start-of-step positions
    f = np.float32
    lin = np.array([f(v[0]) * f(dt), f(v[1]) * f(dt), f(v[2]) * f(dt)], np.float32)
    rr = x[IDX] - np.array(c, np.float32)
    spin = np.stack([f(w[1]) * rr[:, 2] - f(w[2]) * rr[:, 1], f(w[2]) * rr[:, 0] - f(w[0]) * rr[:, 2], f(w[0]) * rr[:, 1] - f(w[1]) * rr[:, 0]], axis=1)
    expect = (x[IDX] - lin) - spin * f(dt)
    assert np.array_equal(p[IDX], expect.astype(np.float32))                          # prev = curr - v dt - (w x r) dt, fp32 as written
    others = np.setdiff1d(np.arange(len(x)), IDX)
    assert np.array_equal(p[others], x[others])                                       # nothing else is touched (the sheet starts at rest)
    plain = run_job({"lib": "oracle", "scene": "sheet", "n": 16, "state": None, "steps": 1}, force_subprocess=True)
    moved = r["curr"][IDX] - plain["curr"][IDX]
    assert float(np.abs(moved).max()) > 5e-4 and float(np.abs(r["curr"][others] - plain["curr"][others]).max()) < float(np.abs(moved).max())


def test_update_rest_shape_changes_the_rest_pose_and_masks_faces(run_job):
    base = {"lib": "oracle", "scene": "sheet", "n": 16, "state": None, "steps": 3}
    plain = run_job(dict(base), force_subprocess=True)
    shrunk = run_job(dict(base, abi_calls=[["update_rest_shape", 1.25, None]]), force_subprocess=True)   # inv_rest x 1.25: rest lengths x 0.8
    masked = run_job(dict(base, abi_calls=[["update_rest_shape", 1.25, list(range(0, 450))]]), force_subprocess=True)
    assert bool(plain["ok"].all()) and bool(shrunk["ok"].all()) and bool(masked["ok"].all())
    ext = lambda x: float(np.ptp(x[:, 0]) + np.ptp(x[:, 2]))
    assert ext(shrunk["curr"]) < ext(plain["curr"]) - 1e-3                            # the sheet contracts toward its smaller rest shape
    assert not np.array_equal(masked["curr"], shrunk["curr"])                         # excluded faces carry no membrane energy
