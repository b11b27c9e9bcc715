"""CPU: known-answer rays for the Bresenham restatement (kernel text gennbv/utils.py:48-167), derived by hand
from the kernel's update rules, plus an independent pure-Python transliteration on random rays."""
import numpy as np

import oracle as c_oracle


def test_axis_ray_in_grid():
    out = c_oracle.bresenham3d([0, 0, 0], [[3, 0, 0]], 8)
    assert out.tolist() == [[0, 0, 0], [1, 0, 0], [2, 0, 0], [3, 0, 0]]


def test_diagonal_and_tiebreak():
    # dx == dy == dz: x is the driving axis; p1 = p2 = dx >= 0 so y and z step every iteration
    out = c_oracle.bresenham3d([1, 1, 1], [[3, 3, 3]], 8)
    assert out.tolist() == [[1, 1, 1], [2, 2, 2], [3, 3, 3]]
    # dx = 4, dy = 2: p1 = 0 -> y steps on iterations 0 and 2
    out = c_oracle.bresenham3d([0, 0, 0], [[4, 2, 0]], 8)
    assert out.tolist() == [[0, 0, 0], [1, 1, 0], [2, 1, 0], [3, 2, 0], [4, 2, 0]]
    # dy == dz > dx: y drives (tie-break dx -> dy -> dz)
    out = c_oracle.bresenham3d([0, 0, 0], [[0, 2, 2]], 8)
    assert out.tolist() == [[0, 0, 0], [0, 1, 1], [0, 2, 2]]


def test_source_outside_only_inbounds_emitted():
    out = c_oracle.bresenham3d([-3, 2, 2], [[2, 2, 2]], 4)
    assert out.tolist() == [[0, 2, 2], [1, 2, 2], [2, 2, 2]]
    # negative direction, source beyond the far face
    out = c_oracle.bresenham3d([6, 1, 0], [[2, 1, 0]], 4)
    assert out.tolist() == [[3, 1, 0], [2, 1, 0]]
    # degenerate ray: source == target
    assert c_oracle.bresenham3d([1, 2, 3], [[1, 2, 3]], 4).tolist() == [[1, 2, 3]]
    assert c_oracle.bresenham3d([9, 9, 9], [[9, 9, 9]], 4).shape == (0, 3)


def _py_bresenham(s, t, G):
    """Independent transliteration of the kernel's x-driving branch generalised by axis permutation."""
    d = [abs(t[i] - s[i]) for i in range(3)]
    sg = [1 if s[i] < t[i] else -1 for i in range(3)]
    dm = max(d)
    a = 0 if dm == d[0] else (1 if dm == d[1] else 2)
    o = [i for i in range(3) if i != a]
    p = list(s)
    e = [2 * d[o[0]] - d[a], 2 * d[o[1]] - d[a]]
    pts = []
    inb = lambda q: all(0 <= c < G for c in q)
    if inb(p):
        pts.append(list(p))
    for _ in range(d[a]):
        if len(pts) >= 3 * G:
            break
        for k in range(2):
            if e[k] >= 0:
                p[o[k]] += sg[o[k]]
                e[k] -= 2 * d[a]
        p[a] += sg[a]
        e[0] += 2 * d[o[0]]
        e[1] += 2 * d[o[1]]
        if inb(p):
            pts.append(list(p))
    return pts


def test_random_rays_against_python_transliteration():
    rng = np.random.default_rng(0)
    G = 12
    for _ in range(300):
        s = rng.integers(-20, 32, 3).tolist()
        t = rng.integers(0, G, 3).tolist()
        got = c_oracle.bresenham3d(s, [t], G).tolist()
        assert got == _py_bresenham(s, t, G), (s, t)
        assert len(got) <= G


def test_pose_to_idx_unclamped():
    rg = np.array([[1.0, -1.0, 2.0, -2.0, 3.0, 0.0]], np.float32)
    vs = np.array([[0.1, 0.2, 0.3]], np.float32)
    out = c_oracle.pose_to_idx(np.array([[-8.0, 8.0, 10.1]], np.float32), rg, vs)
    lo = rg[0, [1, 3, 5]] - np.float32(0.5) * vs[0]
    want = np.floor((np.array([-8.0, 8.0, 10.1], np.float32) - lo) / vs[0]).astype(np.int64)
    assert out[0].tolist() == want.tolist() and out[0, 0] < 0 and out[0, 1] > 20


def test_gae_matches_numpy_loop():
    rng = np.random.default_rng(1)
    T, N = 9, 5
    r, v = rng.standard_normal((T, N)).astype(np.float32), rng.standard_normal((T, N)).astype(np.float32)
    es = (rng.random((T, N)) < 0.2).astype(np.uint8)
    lv, dn = rng.standard_normal(N).astype(np.float32), (rng.random(N) < 0.5).astype(np.uint8)
    adv, ret = c_oracle.gae(r, v, es, lv, dn, 0.99, 0.95)
    g32, gl32 = np.float32(0.99), np.float32(0.99 * 0.95)
    last = np.zeros(N, np.float32)
    for t in reversed(range(T)):
        nnt = (1 - dn if t == T - 1 else 1 - es[t + 1]).astype(np.float32)
        nv = lv if t == T - 1 else v[t + 1]
        delta = r[t] + g32 * nv * nnt - v[t]
        last = delta + gl32 * nnt * last
        assert (adv[t] == last).all()
    assert (ret == adv + v).all()
