"""Geometry pin of the oracle's GEOS restatement (oracle/geos_lite.py).

Two layers:
  * known answers that need no shapely: closed forms of the constructions GEOS documents (a quad_segs = 16 round
    buffer is the regular 64-gon INSCRIBED in the circle with a vertex on the +x axis; the shapely manual's own
    example `Point(0, 0).buffer(10.0).area` prints 313.65484905459...) - these always run;
  * conformance against the REAL shapely, on the call sites the hot path uses (ATT = MADDPG_ownENV_randomOD_radar_
    one_model_att/env_simulator_randomOD_radar_sur_drones_oneModel_att.py: 1077-1164 drone radar, 2172-2176 swept
    capsule, 2243-2250 building contact, 2266-2269 goal contact, 2507 boundary test; V2 1210-1300 grid radar) over 10^4
    random configurations including centre-inside-polygon, grazing and collinear cases.  shapely is absent from the
    build container and from the GPU boxes (probed in round 2): the layer SKIPS there and runs
    wherever `import shapely` works.
"""
import math

import numpy as np
import pytest

from oracle import geos_lite as G


# ------------------------------------------------------------------ known answers (always run)

def test_point_buffer_is_the_inscribed_64_gon_starting_on_the_x_axis():
    p = G.Point(3.0, -2.0).buffer(2.5)
    xy = np.array(p.exterior.coords)
    assert len(xy) == 65 and np.allclose(xy[0], xy[-1])            # 64 vertices + closing point (GEOS createCircle)
    assert np.allclose(xy[0], (5.5, -2.0))                          # first vertex at angle 0
    ang = np.unwrap(np.arctan2(xy[:-1, 1] + 2.0, xy[:-1, 0] - 3.0))
    assert np.allclose(np.diff(ang), -2 * math.pi / 64)             # clockwise, 5.625 degree pitch
    assert np.allclose(np.hypot(xy[:, 0] - 3.0, xy[:, 1] + 2.0), 2.5)   # vertices ON the circle: the polygon is inscribed


def test_published_buffer_areas():
    # shapely manual, object.buffer: Point(0, 0).buffer(10.0).area -> 313.65484905459...  (= 32 r^2 sin(2 pi / 64))
    def area(poly):
        xy = np.array(poly.exterior.coords)
        return 0.5 * abs(np.sum(xy[:-1, 0] * xy[1:, 1] - xy[1:, 0] * xy[:-1, 1]))
    assert abs(area(G.Point(0, 0).buffer(10.0)) - 313.6548490545939) < 1e-9
    assert abs(area(G.Point(0, 0).buffer(10.0)) - 3200.0 * math.sin(math.pi / 32)) < 1e-9
    # square cap (grid cells, ATT/grid_env_generation:151,174): exact 10 x 10 square
    sq = G.Point(5, 5).buffer(5, cap_style=3)
    assert abs(area(sq) - 100.0) < 1e-12 and sq.bounds == (0.0, 0.0, 10.0, 10.0)


def test_round_capped_line_buffer_extents():
    # LineString.buffer(r): a stadium whose caps are 32-segment fans starting perpendicular to the segment; for an
    # axis-parallel segment the cap reaches exactly r beyond the end point (a fan vertex lies on the axis)
    poly = G.LineString([(0.0, 0.0), (4.0, 0.0)]).buffer(2.5, cap_style="round")
    x0, y0, x1, y1 = poly.bounds
    assert np.allclose((x0, y0, x1, y1), (-2.5, -2.5, 6.5, 2.5))
    # a zero-length segment degenerates to the point's 64-gon (GEOS drops the repeated point)
    deg = G.LineString([(1.0, 1.0), (1.0, 1.0)]).buffer(2.5, cap_style="round")
    assert np.allclose(np.array(deg.exterior.coords), np.array(G.Point(1.0, 1.0).buffer(2.5).exterior.coords))


def test_goal_contact_band():
    # two 64-gons sharing vertex angles (radii 2.5 and 1): touching for centre distance <= 3.5 cos(pi/64) in every
    # direction, apart beyond 3.5 in every direction; geometry_test.py's sample is a contact
    for ang in np.linspace(0, 2 * math.pi, 97):
        c, s = math.cos(ang), math.sin(ang)
        near, far = 3.5 * math.cos(math.pi / 64) - 1e-6, 3.5 + 1e-6
        assert not G.Point(0, 0).buffer(2.5).intersection(G.Point(near * c, near * s).buffer(1)).is_empty
        assert G.Point(0, 0).buffer(2.5).intersection(G.Point(far * c, far * s).buffer(1)).is_empty
    assert not G.Point(534.12, 355.86).buffer(2.5).intersection(G.Point(536, 356).buffer(1)).is_empty   # ATT/geometry_test.py:12-14


# ------------------------------------------------------------------ conformance against real shapely

def _shapely():
    return pytest.importorskip("shapely", reason="shapely is not installed here: geometry parity stays unpinned (DESIGN.md section 5)")


def _coords(g):
    return np.array(g.exterior.coords)


def test_buffers_match_shapely_vertex_for_vertex():
    sh = _shapely()
    from shapely.geometry import LineString, Point
    rng = np.random.default_rng(0)
    for _ in range(2000):
        x, y, r = rng.uniform(-500, 500), rng.uniform(-500, 500), rng.choice([1.0, 2.5, 5.0])
        assert np.allclose(_coords(G.Point(x, y).buffer(r)), _coords(Point(x, y).buffer(r)), rtol=0, atol=1e-9)
        assert np.allclose(_coords(G.Point(x, y).buffer(r, cap_style=3)), _coords(Point(x, y).buffer(r, cap_style=3)), rtol=0, atol=1e-9)
        dx, dy = rng.uniform(-3, 3, 2) * rng.choice([0.0, 1.0], 2, p=[0.1, 0.9])
        a, b = (x, y), (x + dx, y + dy)
        got, want = G.LineString([a, b]).buffer(2.5, cap_style="round"), LineString([a, b]).buffer(2.5, cap_style="round")
        assert np.allclose(got.bounds, want.bounds, rtol=0, atol=1e-9)   # the path tests the capsule against lines: extents decide
        assert abs(len(_coords(got)) - len(_coords(want))) <= 1


def test_drone_radar_call_site_matches_shapely():
    """ATT:1077-1164: ray vs another drone's 64-gon: intersects, intersection type, nearest point distance."""
    _shapely()
    from shapely.geometry import LineString, Point
    from shapely.ops import nearest_points
    rng = np.random.default_rng(1)
    for n in range(10000):
        c = rng.uniform(-20, 20, 2)
        ang = math.radians(10 * rng.integers(0, 36))
        end = c + 15.0 * np.array([math.cos(ang), math.sin(ang)])
        kind = n % 4
        if kind == 0:      # anywhere
            o = c + rng.uniform(-20, 20, 2)
        elif kind == 1:    # host centre inside the other drone's polygon
            o = c + rng.uniform(-1.7, 1.7, 2)
        elif kind == 2:    # grazing: the polygon's circle is tangent to the ray up to 1e-3
            t = rng.uniform(0, 15)
            nrm = np.array([-math.sin(ang), math.cos(ang)]) * rng.choice([-1, 1])
            o = c + t * np.array([math.cos(ang), math.sin(ang)]) + nrm * (2.5 * math.cos(math.pi / 64) + rng.uniform(-1e-3, 1e-3))
        else:              # collinear with an edge: centre on the ray's axis
            o = c + rng.uniform(-5, 20) * np.array([math.cos(ang), math.sin(ang)])
        res = []
        for mod_point, mod_line, np_fn in ((G.Point, G.LineString, G.nearest_points), (Point, LineString, nearest_points)):
            line, circle, ctr = mod_line([tuple(c), tuple(end)]), mod_point(*o).buffer(2.5), mod_point(*c)
            if not line.intersects(circle):
                res.append(None)
                continue
            inter = line.intersection(circle)
            res.append(ctr.distance(np_fn(ctr, inter)[1]) if not inter.is_empty else -1.0)
        assert (res[0] is None) == (res[1] is None), (n, c, o, res)
        if res[0] is not None:
            assert abs(res[0] - res[1]) <= 1e-9, (n, c, o, res)


def test_grid_radar_and_contact_call_sites_match_shapely():
    """V2:1210-1300 (ray vs cell boundary / boundary line), ATT:2243-2250 (64-gon vs cell), ATT:2266-2269 (goal),
    ATT:2172-2176 + :2507 (swept capsule vs boundary line)."""
    _shapely()
    from shapely.geometry import LineString, Point
    rng = np.random.default_rng(2)
    for n in range(10000):
        c = rng.uniform(0, 60, 2)
        if n % 5 == 0:
            c = np.round(c / 5.0) * 5.0          # on grid lines / cell centres: collinear and corner cases
        ang = math.radians(5 * rng.integers(0, 72))
        end = c + 15.0 * np.array([math.cos(ang), math.sin(ang)])
        cell = (10.0 * rng.integers(0, 6) + 5.0, 10.0 * rng.integers(0, 6) + 5.0)
        pre = c - rng.uniform(-2.5, 2.5, 2) * rng.choice([0.0, 1.0])
        xb = float(rng.choice([0.0, 60.0]))
        out = []
        for mod_point, mod_line in ((G.Point, G.LineString), (Point, LineString)):
            line, sq, ctr = mod_line([tuple(c), tuple(end)]), mod_point(*cell).buffer(5, cap_style=3), mod_point(*c)
            r = [line.intersects(sq)]
            if r[0]:
                inter = line.intersection(sq.boundary)
                r.append(float("nan") if inter.is_empty else ctr.distance(inter))
            bound = mod_line([(xb, -9999), (xb, 9999)])
            r.append(line.intersects(bound))
            if r[-1]:
                r.append(ctr.distance(line.intersection(bound)))
            circle = ctr.buffer(2.5)
            r.append(bool(sq.intersection(circle)))                                   # ATT:2245
            r.append(not circle.intersection(mod_point(*cell).buffer(1, cap_style="round")).is_empty)   # ATT:2266-2269 (goal at the cell centre)
            r.append(bound.intersects(mod_line([tuple(pre), tuple(c)]).buffer(2.5, cap_style="round")))   # ATT:2172-2173, :2507
            out.append(r)
        assert len(out[0]) == len(out[1]), (n, out)
        for a, b in zip(*out):
            if isinstance(a, float):
                assert (math.isnan(a) and math.isnan(b)) or abs(a - b) <= 1e-9, (n, c, ang, cell, out)
            else:
                assert a == b, (n, c, ang, cell, out)
