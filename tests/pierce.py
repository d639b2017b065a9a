"""Deliberately intersecting start states for the positive path of check_intersection (contact.cu:3056-3175).

  pierced_curtain   `headless` at res 24: one free vertex in the middle of the first curtain is pushed through the second one
                    (0.25 away), by a generic offset so that no crossing lies on an edge of the other mesh -> its incident edges
                    pierce triangles of the neighbouring curtain (records of type 0, face-edge)
  crowded_rods      `woven` at n 32 (rods with contact offsets): one rod vertex is moved to within 0.3 of the sum of the contact
                    offsets of a vertex of another strand -> point-point (type 3) and edge-edge (type 1) records

Both functions need a Session only for the rest positions and the topology (any backend; the CPU oracle in the tests).
"""
import numpy as np


def pierced_curtain(ppf, lib, path):
    s = ppf.Session("headless", n=24, lib=lib)
    x, _ = s.fetch()
    face, edge, _ = s.topology()
    idx, _ = s.fix_targets()
    pinned = np.zeros(len(x), bool); pinned[idx] = True
    first = np.arange(24 * 24)                                                     # the first curtain (the sphere comes last)
    assert np.ptp(x[first, 0]) < 1e-6 and abs(float(x[24 * 24, 0] - x[0, 0]) - 0.25) < 1e-5
    c = x[first].mean(axis=0)
    v = int(first[np.argmin(np.linalg.norm(x[first] - c, axis=1) + 10.0 * pinned[first])])
    h = float(np.linalg.norm(x[edge[0, 0]] - x[edge[0, 1]]))                        # a grid edge length
    y = x.copy()
    y[v] += np.array([0.30, 0.37 * h, 0.21 * h], np.float32)
    ppf.write_state_file(path, y, y, time=0.0, prev_dt=0.01)
    return v, x, y, face, edge


def crowded_rods(ppf, lib, path):
    s = ppf.Session("woven", n=32, lib=lib)
    x, _ = s.fetch()
    face, edge, vp = s.topology()                                                    # vp: per-vertex (ghat, offset)
    idx, _ = s.fix_targets()
    pinned = np.zeros(len(x), bool); pinned[idx] = True
    nbr = [set() for _ in range(len(x))]
    for a, b in edge:
        nbr[a].add(int(b)); nbr[b].add(int(a))
    free = np.where(~pinned)[0]
    i = int(free[len(free) // 2])
    d = np.linalg.norm(x - x[i], axis=1)
    d[i] = np.inf
    for j in nbr[i] | {k for j in nbr[i] for k in nbr[j]}:                           # not a neighbour along the same rod
        d[j] = np.inf
    j = int(np.argmin(d))
    off = float(vp[i, 1] + vp[j, 1])
    assert off > 0.0
    y = x.copy()
    u = (x[i] - x[j]) / np.linalg.norm(x[i] - x[j])
    y[i] = x[j] + np.float32(0.3 * off) * u.astype(np.float32)
    ppf.write_state_file(path, y, y, time=0.0, prev_dt=1.0 / 120.0)
    return i, j, off, y, edge


def segment_pierces_triangle(p0, p1, a, b, c):
    """float64 check used to validate a face-edge record"""
    p0, p1, a, b, c = (np.asarray(v, np.float64) for v in (p0, p1, a, b, c))
    n = np.cross(b - a, c - a)
    d0, d1 = np.dot(p0 - a, n), np.dot(p1 - a, n)
    if d0 * d1 > 0:
        return False
    t = d0 / (d0 - d1)
    q = p0 + t * (p1 - p0)
    m = np.column_stack([b - a, c - a, n])
    u, v, _ = np.linalg.solve(m, q - a)
    return u >= -1e-9 and v >= -1e-9 and u + v <= 1 + 1e-9
