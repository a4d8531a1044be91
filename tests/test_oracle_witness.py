"""A second, independent restatement of the Chipmunk2D query routines (plain numpy, written from the published
algorithm, sharing no code with ``oracle/cat_oracle.c``) run against the C oracle on random queries.

This does not pin the oracle to pymunk (which cannot be installed here — parity stays "unpinned"); it guards
against transcription slips in the C restatement that the analytic vectors do not reach: plane indexing, the
tangential extent test, the bevel circles, the start-inside rule, the thin-ray leaf test, static-before-dynamic
tie order.  Routines restated: cpBBSegmentQuery, cpClosetPointOnSegment, cpPolyShapePointQuery,
CircleSegmentQuery, cpPolyShapeSegmentQuery, cpShapeSegmentQuery, cpSpaceSegmentQueryFirst, and one body's
cpSpaceStep against a single wall (cpArbiterPreStep / cpArbiterApplyImpulse).
"""
import numpy as np
import pytest

import parity_utils as pu
from oracle.cat_oracle import Oracle

WALL_R, AGENT_R = 1.0, 5.0


def _hulls(cm):
    return [cm.vert[cm.hull_off[h]:cm.hull_off[h + 1]] for h in range(cm.n_hulls)]


def bb_segment_query(bb, a, b):
    l, bt, r, t = bb
    d = b - a
    tmin, tmax = -np.inf, np.inf
    if d[0] == 0.0:
        if a[0] < l or r < a[0]:
            return np.inf
    else:
        t1, t2 = (l - a[0]) / d[0], (r - a[0]) / d[0]
        tmin, tmax = max(tmin, min(t1, t2)), min(tmax, max(t1, t2))
    if d[1] == 0.0:
        if a[1] < bt or t < a[1]:
            return np.inf
    else:
        t1, t2 = (bt - a[1]) / d[1], (t - a[1]) / d[1]
        tmin, tmax = max(tmin, min(t1, t2)), min(tmax, max(t1, t2))
    return max(tmin, 0.0) if (tmin <= tmax and 0.0 <= tmax and tmin <= 1.0) else np.inf


def closest_on_segment(p, a, b):
    delta = a - b
    t = np.clip(delta @ (p - b) / (delta @ delta), 0.0, 1.0)
    return b + delta * t


def poly_point_distance(v, p):
    """Signed distance to the raw hull (negative inside), as cpPolyShapePointQuery before subtracting r."""
    n = len(v)
    outside, best = False, np.inf
    for i in range(n):
        v0, v1 = v[i - 1], v[i]
        e = v1 - v0
        nrm = np.array([e[1], -e[0]]) / np.hypot(*e)
        outside = outside or (nrm @ (p - v1) > 0.0)
        best = min(best, np.hypot(*(p - closest_on_segment(p, v0, v1))))
    return best if outside else -best


def circle_segment_query(c, r1, a, b, r2):
    da, db = a - c, b - c
    rsum = r1 + r2
    qa = da @ da - 2.0 * (da @ db) + db @ db
    qb = da @ db - da @ da
    det = qb * qb - qa * (da @ da - rsum * rsum)
    if det >= 0.0:
        t = (-qb - np.sqrt(det)) / qa
        if 0.0 <= t <= 1.0:
            nn = da * (1 - t) + db * t
            nn = nn / np.hypot(*nn)
            return t, a * (1 - t) + b * t - nn * r2
    return None


def poly_segment_query(v, r, a, b, r2):
    n = len(v)
    rsum = r + r2
    alpha, point = 1.0, None
    for i in range(n):
        v0, v1 = v[i - 1], v[i]
        e = v1 - v0
        nrm = np.array([e[1], -e[0]]) / np.hypot(*e)
        an = a @ nrm
        d = an - v1 @ nrm - rsum
        if d < 0.0:
            continue
        bn = b @ nrm
        with np.errstate(over="ignore"):
            t = d / max(an - bn, np.finfo(float).tiny)
        if t < 0.0 or 1.0 < t:
            continue
        pt = a * (1 - t) + b * t
        cross = lambda n_, p_: n_[0] * p_[1] - n_[1] * p_[0]
        if cross(nrm, v0) <= cross(nrm, pt) <= cross(nrm, v1):
            alpha, point = t, pt - nrm * r2
    if rsum > 0.0:
        for i in range(n):
            hit = circle_segment_query(v[i], r, a, b, r2)
            if hit is not None and hit[0] < alpha:
                alpha, point = hit
    return None if point is None else (alpha, point)


def shape_segment_query_poly(v, r, a, b, r2):
    if poly_point_distance(v, a) - r <= r2:
        return 0.0, b.copy()
    return poly_segment_query(v, r, a, b, r2)


def shape_segment_query_circle(c, r, a, b, r2):
    if np.hypot(*(a - c)) - r <= r2:
        return 0.0, b.copy()
    return circle_segment_query(c, r, a, b, r2)


def space_segment_query_first(hulls, centres, self_agent, a, b, r2):
    best, shape, point = 1.0, -1, b.copy()
    for h, v in enumerate(hulls):
        bb = (v[:, 0].min() - WALL_R, v[:, 1].min() - WALL_R, v[:, 0].max() + WALL_R, v[:, 1].max() + WALL_R)
        if not bb_segment_query(bb, a, b) < 1.0:
            continue
        hit = shape_segment_query_poly(v, WALL_R, a, b, r2)
        if hit is not None and hit[0] < best:
            best, point, shape = hit[0], hit[1], h
    if self_agent >= 0:
        for j, c in enumerate(centres):
            if j == self_agent:
                continue
            hit = shape_segment_query_circle(c, AGENT_R, a, b, r2)
            if hit is not None and hit[0] < best:
                best, point, shape = hit[0], hit[1], len(hulls) + j
    return shape, best, point


@pytest.mark.parametrize("name,free", [("squarinth", False), ("labyrinth", True), ("agh-map", True)])
def test_segment_queries_agree_with_the_witness(name, free):
    cm = pu.named_cmap(name, free_spawn=free)
    orc = Oracle(cm)
    hulls = _hulls(cm)
    rng = np.random.default_rng(7)
    lo = np.array([cm.grid_x0, cm.grid_y0])
    hi = lo + np.array([cm.nx, cm.ny]) * cm.cell
    n_q = 400 if name == "agh-map" else 1500
    kinds = {"none": 0, "wall": 0, "agent": 0, "alpha0": 0}
    for q in range(n_q):
        a = rng.uniform(lo, hi)
        centres = rng.uniform(lo, hi, size=(3, 2))
        if q % 3 == 0:                      # make agent hits likely
            centres[1] = a + rng.uniform(-120, 120, 2)
        if q % 7 == 0:                      # start next to a wall vertex: exercises the start-inside and bevel paths
            v = hulls[rng.integers(len(hulls))]
            a = v[rng.integers(len(v))] + rng.uniform(-2.5, 2.5, 2)
        ang = rng.uniform(0, 2 * np.pi)
        radius = 1.0 if q % 5 else 0.0      # sensor rays (radius 1) and the capture line of sight (radius 0)
        self_agent = 0 if radius > 0 else -1
        b = a + 400.0 * np.array([np.cos(ang), np.sin(ang)]) if radius > 0 else centres[2]
        shape, alpha, point = orc.segment_query_first(centres.reshape(-1), self_agent, a, b, radius)
        wshape, walpha, wpoint = space_segment_query_first(hulls, centres, self_agent, a, b, radius)
        assert (shape < 0) == (wshape < 0), (q, shape, wshape)
        if shape < 0:
            kinds["none"] += 1
            continue
        assert abs(alpha - walpha) <= 1e-9, (q, alpha, walpha)
        assert np.allclose(point, wpoint, atol=1e-7), (q, point, wpoint)
        if abs(alpha - walpha) == 0 and alpha > 0:
            assert (shape >= len(hulls)) == (wshape >= len(hulls))     # same kind of shape wins
        kinds["alpha0" if alpha == 0 else ("agent" if shape >= len(hulls) else "wall")] += 1
    # coverage of the query outcomes (random points of agh-map mostly start inside a building: few agent hits there)
    assert kinds["wall"] > 50 and kinds["alpha0"] > 5 and kinds["agent"] > (0 if name == "agh-map" else 5), kinds


def test_single_wall_contact_step_agrees_with_the_witness():
    """One body pushed into one wall for a few steps: position integrate, contact at d <= 6, bias from the
    penetration, normal velocity removed, v_bias applied on the next step (cpSpaceStep order)."""
    cm = pu.named_cmap("squarinth")
    orc = Oracle(cm)
    hulls = _hulls(cm)
    dt, slop, bias_coef = 1 / 60, 0.1, 1 - 0.9 ** (60 * (1 / 60))
    st = orc.new_state(1)
    st.pos[0] = [[115.0, 400.0], [400.0, 400.0], [600.0, 600.0]]     # agent 0 next to the left wall (x in [100, 105])
    st.tc[0] = st.pos[0]
    st.vel[0] = [[-60.0, 7.0], [0.0, 0.0], [0.0, 0.0]]
    p, v, vb = st.pos[0, 0].copy(), st.vel[0, 0].copy(), np.zeros(2)
    jn_cached, touching_prev = 0.0, False
    for _ in range(6):
        orc.step(st, np.full((1, 3), 9, np.int32))                      # action 9: no impulse
        # ---- witness
        p = p + (v + vb) * dt
        vb = np.zeros(2)
        d, nrm = np.inf, None
        for hv in hulls:
            dd = poly_point_distance(hv, p)
            if dd < d:
                d = dd
                # closest feature direction (outside, edge region): from the body towards the hull
                best = None
                for i in range(len(hv)):
                    c = closest_on_segment(p, hv[i - 1], hv[i])
                    if best is None or np.hypot(*(p - c)) < np.hypot(*(p - best)):
                        best = c
                nrm = (best - p) / np.hypot(*(best - p))
        if d <= AGENT_R + WALL_R:
            dist = d - (AGENT_R + WALL_R)
            bias = -bias_coef * min(0.0, dist + slop) / dt
            jn = jn_cached if touching_prev else 0.0
            if touching_prev:
                v = v - nrm * jn                                          # warm start (mass 1)
            jb = 0.0
            for _it in range(10):
                jb_new = max(jb + (bias - (-(vb @ nrm))), 0.0)
                vb = vb - nrm * (jb_new - jb)
                jb = jb_new
                jn_new = max(jn - (-(v @ nrm)), 0.0)
                v = v - nrm * (jn_new - jn)
                jn = jn_new
            jn_cached, touching_prev = jn, True
        else:
            touching_prev = False
        assert np.allclose(st.pos[0, 0], p, atol=1e-9) and np.allclose(st.vel[0, 0], v, atol=1e-9)
        assert np.allclose(st.vbias[0, 0], vb, atol=1e-9)
    assert touching_prev and abs(v[0]) < 1e-9 and abs(v[1] - 7.0) < 1e-12     # e = 0, mu = 0
