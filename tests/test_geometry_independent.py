"""Third-party check of the oracle's geometry PREDICATES (oracle/geos_lite.py) with sympy's exact geometry.

shapely / GEOS exist neither in the build container nor on the GPU boxes, so `test_geos_conformance.py`'s
shapely layer skips there.  sympy does exist, and its `sympy.geometry` computes intersections and distances of segments and
polygons in exact rational arithmetic, written by other people with another method.  On the call sites the hot path uses
(ATT = MADDPG_ownENV_randomOD_radar_one_model_att/env_simulator_randomOD_radar_sur_drones_oneModel_att.py:1077-1164 ray vs a
drone's 64-gon; V2 = ...tdCPA_forV2/env_simulator_...:1210-1300 ray vs a cell's boundary; ATT:2243-2250 64-gon vs cell;
ATT:2266-2269 goal contact) this file takes the polygons geos_lite BUILDS (the construction is pinned separately: published
areas and vertex layout, `test_geos_conformance.py`) and checks what geos_lite then DECIDES about them - intersects, the
intersection's nearest point, distances - against sympy on the same vertices, including the cases a misreading would get
wrong: ray origin inside the polygon, grazing a vertex, collinear with an edge, touching squares.

What this does not pin: that GEOS places the buffer vertices where geos_lite does (only real shapely can).
"""
import math

import numpy as np
import pytest

from oracle import geos_lite as G

sp = pytest.importorskip("sympy")
cv2 = pytest.importorskip("cv2")
from sympy import Rational  # noqa: E402
from sympy.geometry import Point2D, Polygon, Segment2D  # noqa: E402


def spoly(g):
    pts = [Point2D(Rational(x), Rational(y)) for x, y in list(g.exterior.coords)[:-1]]
    return Polygon(*pts)


def sseg(a, b):
    return Segment2D(Point2D(Rational(a[0]), Rational(a[1])), Point2D(Rational(b[0]), Rational(b[1])))


def inside_closed(poly, p):
    """p in the CLOSED convex polygon: every edge sees it on the same side or on its line (exact rationals; sympy's general
    `encloses_point` takes seconds per call on a 64-gon)."""
    v = poly.vertices
    sign = 0
    for a, b in zip(v, v[1:] + v[:1]):
        cr = (b.x - a.x) * (p.y - a.y) - (b.y - a.y) * (p.x - a.x)
        if cr != 0:
            if sign and (cr > 0) != (sign > 0):
                return False
            sign = 1 if cr > 0 else -1
    return True


def d2(it, a):
    """Squared distance (exact) from point a to a sympy intersection item (point, or segment of a collinear overlap)."""
    if isinstance(it, Point2D):
        return (it.x - a.x) ** 2 + (it.y - a.y) ** 2
    return min((q.x - a.x) ** 2 + (q.y - a.y) ** 2 for q in (it.p1, it.p2))      # overlaps are collinear with the ray: an end is nearest


def boundary_hits(poly, seg):
    return [it for side in poly.sides for it in side.intersection(seg)]


def test_ray_against_a_drones_64_gon_matches_sympy():
    """ATT:1077-1164: `line.intersects(circle)`, then the distance from the host to the nearest point of the intersection."""
    rng = np.random.default_rng(11)
    n_hit = n_inside = 0
    for n in range(20):
        c = rng.uniform(-20, 20, 2)
        ang = math.radians(10 * int(rng.integers(0, 36)))
        d = np.array([math.cos(ang), math.sin(ang)])
        kind = n % 4
        if kind == 0:
            o = c + rng.uniform(0, 14) * d + rng.uniform(-4, 4, 2)
        elif kind == 1:                                   # host centre inside the other drone's polygon
            o = c + rng.uniform(-1.7, 1.7, 2)
        elif kind == 2:                                   # grazing: tangent to the polygon's circle up to 1e-3
            nrm = np.array([-d[1], d[0]]) * rng.choice([-1, 1])
            o = c + rng.uniform(0, 15) * d + nrm * (2.5 * math.cos(math.pi / 64) + rng.uniform(-1e-3, 1e-3))
        else:                                             # the other centre on the ray's axis: the ray runs through a vertex
            o = c + rng.uniform(-5, 20) * d
        end = c + 15.0 * d
        line, circle, ctr = G.LineString([tuple(c), tuple(end)]), G.Point(*o).buffer(2.5), G.Point(*c)
        poly, seg = spoly(circle), sseg(c, end)
        hits = boundary_hits(poly, seg)
        origin_inside = inside_closed(poly, seg.p1)
        want = 0.0 if origin_inside else (math.sqrt(float(min(d2(it, seg.p1) for it in hits))) if hits else None)
        assert line.intersects(circle) == (want is not None), (n, c, o)
        if want is None:
            continue
        n_hit += 1
        n_inside += origin_inside
        inter = line.intersection(circle)
        assert not inter.is_empty
        got = ctr.distance(G.nearest_points(ctr, inter)[1])
        assert abs(got - want) <= 1e-9, (n, c, o, got, want)
    assert n_hit >= 10 and n_inside >= 5


def test_ray_against_a_cell_boundary_matches_sympy():
    """V2:1210-1300: `line.intersects(cell)` and `host.distance(line.intersection(cell.boundary))`."""
    rng = np.random.default_rng(12)
    n_hit = n_miss = 0
    for n in range(80):
        c = rng.uniform(0, 60, 2)
        if n % 3 == 0:
            c = np.round(c / 5.0) * 5.0                  # on grid lines / cell centres: collinear and corner cases
        ang = math.radians(5 * int(rng.integers(0, 72)))
        end = c + 15.0 * np.array([math.cos(ang), math.sin(ang)])
        near = c + rng.uniform(0, 18) * np.array([math.cos(ang), math.sin(ang)]) + rng.uniform(-8, 8, 2)
        cell = (10.0 * math.floor(near[0] / 10.0) + 5.0, 10.0 * math.floor(near[1] / 10.0) + 5.0)
        line, sq, ctr = G.LineString([tuple(c), tuple(end)]), G.Point(*cell).buffer(5, cap_style=3), G.Point(*c)
        poly, seg = spoly(sq), sseg(c, end)
        hits = boundary_hits(poly, seg)
        assert line.intersects(sq) == (bool(hits) or inside_closed(poly, seg.p1)), (n, c, ang, cell)
        if not hits:
            n_miss += 1
            continue
        n_hit += 1
        want = math.sqrt(float(min(d2(it, seg.p1) for it in hits)))
        inter = line.intersection(sq.boundary)
        assert not inter.is_empty
        assert abs(ctr.distance(inter) - want) <= 1e-9, (n, c, ang, cell)
    assert n_hit >= 25 and n_miss >= 5


def _cv_overlap(a, b):
    """Area of the intersection of two convex polygons by OpenCV (float32 vertices): > 0 overlapping, 0.0 disjoint."""
    pa = np.array(list(a.exterior.coords)[:-1], dtype=np.float32)
    pb = np.array(list(b.exterior.coords)[:-1], dtype=np.float32)
    return float(cv2.intersectConvexConvex(pa, pb)[0])


def test_gon_against_cell_and_goal_contact_match_opencv_and_sympy():
    """ATT:2243-2250 (`cell.intersection(drone_circle)` non-empty) and ATT:2266-2269 (drone 64-gon vs goal 64-gon): OpenCV's
    convex-polygon intersection on every case whose margin exceeds its float32 vertices (overlap deeper than 1e-3 m or a
    clear gap), sympy's exact edge intersections on a handful of 64-gon / square cases."""
    rng = np.random.default_rng(13)
    n_touch = n_apart = n_exact = 0
    for n in range(120):
        cell = (10.0 * int(rng.integers(0, 4)) + 5.0, 10.0 * int(rng.integers(0, 4)) + 5.0)
        side = n % 3                                      # faces, corners, anywhere
        gap = rng.choice([-1, 1]) * rng.uniform(2e-3, 0.05)       # signed distance between the outlines, away from float32's reach
        if side == 0:
            c = np.array([cell[0] + 5.0 + 2.5 + gap, cell[1] + rng.uniform(-4, 4)])      # the 64-gon has a VERTEX on the -x axis: reach 2.5
        elif side == 1:
            a = rng.uniform(0.2, math.pi / 2 - 0.2)
            c = np.array([cell[0] + 5.0, cell[1] + 5.0]) + (2.5 + gap) * np.array([math.cos(a), math.sin(a)])
        else:
            c = np.array(cell) + rng.uniform(-9, 9, 2)
        sq, circle = G.Point(*cell).buffer(5, cap_style=3), G.Point(*c).buffer(2.5)
        got = bool(sq.intersection(circle))
        area = _cv_overlap(sq, circle)
        if side == 0:                                     # face contact: the margin is known exactly
            assert got == (gap < 0), (n, cell, c, gap)
        if area > 1e-3:
            assert got, (n, cell, c, area)
            n_touch += 1
        elif area == 0.0 and (side != 1 or gap > 0.01):
            assert not got or side == 2, (n, cell, c)     # (anywhere-cases with zero float32 area may still touch: not judged)
            n_apart += not got
        if n < 8:                                         # exact: edges cross, or one polygon holds a vertex of the other
            ps, pc = spoly(sq), spoly(circle)
            want = any(s1.intersection(s2) for s1 in ps.sides for s2 in pc.sides) or inside_closed(ps, pc.vertices[0]) or inside_closed(pc, ps.vertices[0])
            assert got == bool(want), (n, cell, c)
            n_exact += 1
        # goal contact: the same drone polygon against a 1 m 64-gon a signed margin away from first contact in direction a
        a = 2 * math.pi * rng.integers(0, 64) / 64 + rng.choice([0.0, math.pi / 64])     # vertex-to-vertex and edge-to-edge directions
        frac = (a * 64 / (2 * math.pi)) % 1.0
        reach = 3.5 if min(frac, 1.0 - frac) < 0.25 else 3.5 * math.cos(math.pi / 64)       # vertex to vertex, or edge to edge
        g = c + (reach + gap) * np.array([math.cos(a), math.sin(a)])
        goal = G.Point(*g).buffer(1)
        assert (not circle.intersection(goal).is_empty) == (gap < 0), (n, c, g, gap)
        area_g = _cv_overlap(circle, goal)
        if area_g > 1e-3:
            assert gap < 0
    assert n_touch >= 30 and n_apart >= 20 and n_exact == 8


def test_swept_capsule_against_a_boundary_line_matches_sympy():
    """ATT:2172-2173, :2507: `boundary_line.intersects(LineString([pre, pos]).buffer(2.5, cap_style='round'))` - the stadium
    geos_lite builds (convex) against the long boundary segment, decided exactly by sympy on the same vertices."""
    rng = np.random.default_rng(14)
    n_hit = n_miss = 0
    for n in range(40):
        xb = float(rng.choice([0.0, 60.0]))
        sign = 1.0 if xb == 0.0 else -1.0
        pos = np.array([xb + sign * (2.5 + rng.uniform(-0.03, 0.03)), rng.uniform(5, 55)])     # around first contact of the cap
        step = rng.uniform(0, 2.5) * np.array([math.cos(a := rng.uniform(0, 2 * math.pi)), math.sin(a)])
        pre = pos + sign * np.array([abs(step[0]), step[1]]) if n % 4 else pos.copy()         # moving towards the line, or standing still
        cap = G.LineString([tuple(pre), tuple(pos)]).buffer(2.5, cap_style="round")
        bound = G.LineString([(xb, -9999.0), (xb, 9999.0)])
        poly, seg = spoly(cap), sseg((xb, -9999.0), (xb, 9999.0))
        want = bool(boundary_hits(poly, seg)) or inside_closed(poly, seg.p1)
        assert bound.intersects(cap) == want, (n, pre, pos, xb)
        n_hit += want
        n_miss += not want
    assert n_hit >= 8 and n_miss >= 8


def test_reference_line_projection_matches_sympy():
    """ATT:3203-3214, UV2:413-441: `line.project(p)` (arc length of the nearest point of the reference polyline),
    `line.interpolate(d)` and `p.distance(line)` on axis-parallel polylines like the reference lines (cell centres joined
    where the direction changes) - against sympy's exact point / segment distances and projections, leg by leg."""
    rng = np.random.default_rng(15)
    for n in range(60):
        k = int(rng.integers(2, 7))
        pts = [np.array([5.0 + 10.0 * int(rng.integers(0, 8)), 5.0 + 10.0 * int(rng.integers(0, 8))])]
        horiz = bool(rng.integers(0, 2))
        while len(pts) < k:
            step = 10.0 * int(rng.integers(1, 5)) * rng.choice([-1, 1])
            pts.append(pts[-1] + (np.array([step, 0.0]) if horiz else np.array([0.0, step])))
            horiz = not horiz
        p = pts[int(rng.integers(0, k))] + rng.uniform(-12, 12, 2) + 1e-3 * rng.uniform(0.1, 1.0, 2)   # (off the bisectors: no exact ties)
        line, P = G.LineString([tuple(q) for q in pts]), Point2D(Rational(p[0]), Rational(p[1]))
        best, arc, run = None, None, Rational(0)
        for a, b in zip(pts[:-1], pts[1:]):
            seg = sseg(a, b)
            q = seg.projection(P)                       # foot on the carrier line, clamped to the leg below
            if not seg.contains(q):
                q = min((seg.p1, seg.p2), key=lambda e: (e.x - P.x) ** 2 + (e.y - P.y) ** 2)
            dd = (q.x - P.x) ** 2 + (q.y - P.y) ** 2
            if best is None or dd < best:               # first leg attaining the minimum wins
                best, arc = dd, run + abs(q.x - seg.p1.x) + abs(q.y - seg.p1.y)     # legs are axis-parallel: L1 = length
            run += abs(seg.p2.x - seg.p1.x) + abs(seg.p2.y - seg.p1.y)
        got_arc = line.project(G.Point(*p))
        assert abs(got_arc - float(arc)) <= 1e-9, (n, pts, p)
        assert abs(G.Point(*p).distance(line) - math.sqrt(float(best))) <= 1e-9, (n, pts, p)
        back = line.interpolate(got_arc)
        assert abs(G.Point(*p).distance(back) - math.sqrt(float(best))) <= 1e-9, (n, pts, p)
        assert abs(line.length - float(run)) <= 1e-9
