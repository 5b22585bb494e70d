"""geos_lite -- TEST INFRASTRUCTURE ONLY (never imported by the product path).

A small float64 restatement of the shapely 2.0.1 / GEOS 3.11 geometry semantics that the
reference's hot path calls.  shapely itself is not installable in the build container
(SURVEY.md section 8c), so this module lets `tests/golden/gen_golden.py` import and run the
*unmodified* reference `env_simulator_*` classes from /root/reference: the reference's own
control flow, observation layout and reward branches execute verbatim, and only the geometry
primitives below are restated.  Parity status: the geometry layer is "parity unpinned"
(no GEOS binary here to compare against); everything above it is pinned by the real reference.
What is checked without GEOS: the constructions against the answers GEOS / shapely publish, and
the predicates against sympy's exact geometry and OpenCV on the polygons built here
(tests/test_geos_conformance.py, tests/test_geometry_independent.py).

Call sites this module has to serve (ATT = MADDPG_ownENV_randomOD_radar_one_model_att/
env_simulator_randomOD_radar_sur_drones_oneModel_att.py):
  STRtree(...)                         ATT:89,96,150,1082     .query :1106,2243   .nearest :1053
  Point.buffer / LineString.buffer     ATT:1080,1199,2173-2175,2266
  intersects / intersection            ATT:1111-1114,2245,2269,2507
  nearest_points / project / interpolate  ATT:1136,1201-1231,3205; V2/Utilities_own:429-438

GEOS rules restated here:
  * Point.buffer(r) (quad_segs=16): 64-gon, vertices p + r*(cos(-i*2pi/64), sin(-i*2pi/64)), i=0..63
    (OffsetSegmentGenerator::createCircle + addDirectedFillet, clockwise from angle 0).
  * LineString([a,b]).buffer(r, round caps): stadium; each cap is a 32-segment clockwise fan that
    starts at angle(a->b)+pi/2; a zero-length line degenerates to the point buffer.
  * Point.buffer(r, cap_style=3): axis-aligned square of half-width r.
  * predicates are closed-set (touching counts as intersecting).
  * nearest_points / project: first segment attaining the minimum wins (strict '<' scan).
  * STRtree.query returns candidates by envelope overlap.  GEOS returns them in tree-traversal
    order, which is not reproducible from source; this shim returns ascending
    insertion index.
"""
import math

import numpy as np

_TWO_PI = 2.0 * math.pi


def _xy(c):
    """Coordinate from Point / tuple / list / ndarray."""
    if isinstance(c, Point):
        return (c.x, c.y)
    return (float(c[0]), float(c[1]))


def _cross(ax, ay, bx, by):
    return ax * by - ay * bx


def _pt_seg_closest(px, py, ax, ay, bx, by):
    """Closest point on closed segment ab to p (GEOS Distance::pointToSegment layout)."""
    dx, dy = bx - ax, by - ay
    len2 = dx * dx + dy * dy
    if len2 == 0.0:
        return ax, ay
    r = ((px - ax) * dx + (py - ay) * dy) / len2
    if r <= 0.0:
        return ax, ay
    if r >= 1.0:
        return bx, by
    return ax + r * dx, ay + r * dy


def _seg_seg(p, q, a, b):
    """Intersection of closed segments pq and ab: [] | [pt] | [pt1, pt2] (collinear overlap)."""
    rx, ry = q[0] - p[0], q[1] - p[1]
    sx, sy = b[0] - a[0], b[1] - a[1]
    rxs = _cross(rx, ry, sx, sy)
    qpx, qpy = a[0] - p[0], a[1] - p[1]
    qpxr = _cross(qpx, qpy, rx, ry)
    if rxs == 0.0:
        if qpxr != 0.0:
            return []
        rr = rx * rx + ry * ry
        if rr == 0.0:  # pq is a point
            if _on_segment(p, a, b):
                return [p]
            return []
        t0 = (qpx * rx + qpy * ry) / rr
        t1 = t0 + (sx * rx + sy * ry) / rr
        lo, hi = min(t0, t1), max(t0, t1)
        lo, hi = max(lo, 0.0), min(hi, 1.0)
        if lo > hi:
            return []
        p0 = (p[0] + lo * rx, p[1] + lo * ry)
        if lo == hi:
            return [p0]
        return [p0, (p[0] + hi * rx, p[1] + hi * ry)]
    t = _cross(qpx, qpy, sx, sy) / rxs
    u = qpxr / rxs
    if 0.0 <= t <= 1.0 and 0.0 <= u <= 1.0:
        return [(p[0] + t * rx, p[1] + t * ry)]
    return []


def _on_segment(p, a, b):
    if _cross(b[0] - a[0], b[1] - a[1], p[0] - a[0], p[1] - a[1]) != 0.0:
        return False
    return (min(a[0], b[0]) <= p[0] <= max(a[0], b[0])) and (min(a[1], b[1]) <= p[1] <= max(a[1], b[1]))


def _point_in_ring(p, ring):
    """Closed test: inside or on the boundary of the ring (list of coords, closed)."""
    inside = False
    n = len(ring) - 1
    for i in range(n):
        a, b = ring[i], ring[i + 1]
        if _on_segment(p, a, b):
            return True
        if (a[1] > p[1]) != (b[1] > p[1]):
            xint = a[0] + (p[1] - a[1]) * (b[0] - a[0]) / (b[1] - a[1])
            if p[0] < xint:
                inside = not inside
    return inside


class _CoordSeq(list):
    """list of (x, y) tuples with the `.xy` accessor shapely's CoordinateSequence has."""

    @property
    def xy(self):
        return [c[0] for c in self], [c[1] for c in self]


class BaseGeometry:
    geom_type = "GeometryCollection"

    @property
    def type(self):   # shapely 2.0's deprecated alias of geom_type (used at MM:932)
        return self.geom_type
    is_empty = False

    def __bool__(self):
        return not self.is_empty

    def _segments(self):
        return []

    def _points(self):
        return []

    @property
    def bounds(self):
        pts = self._points()
        xs = [c[0] for c in pts]
        ys = [c[1] for c in pts]
        return (min(xs), min(ys), max(xs), max(ys))

    # -- distance: min over component primitives ------------------------------------------
    def distance(self, other):
        if self.is_empty or other.is_empty:
            return float("nan")
        if isinstance(other, Point) and not isinstance(self, Point):
            return other.distance(self)
        raise NotImplementedError("distance %s-%s" % (self.geom_type, other.geom_type))


class EmptyGeometry(BaseGeometry):
    is_empty = True

    def __init__(self, geom_type="GeometryCollection"):
        self.geom_type = geom_type

    length = 0.0
    geoms = ()
    coords = _CoordSeq()


class Point(BaseGeometry):
    geom_type = "Point"

    def __init__(self, *args):
        if len(args) == 1:
            x, y = _xy(args[0])
        else:
            x, y = float(args[0]), float(args[1])
        self.x = x
        self.y = y

    @property
    def coords(self):
        return _CoordSeq([(self.x, self.y)])

    @property
    def xy(self):
        return [self.x], [self.y]

    def _points(self):
        return [(self.x, self.y)]

    def buffer(self, distance, quad_segs=16, cap_style="round", **_kw):
        if cap_style in (3, "square"):
            d = distance
            ring = [(self.x + d, self.y + d), (self.x + d, self.y - d), (self.x - d, self.y - d),
                    (self.x - d, self.y + d)]
            return Polygon(ring)
        n = 4 * quad_segs
        inc = _TWO_PI / n
        ring = []
        for i in range(n):
            ang = -1.0 * i * inc
            ring.append((self.x + distance * math.cos(ang), self.y + distance * math.sin(ang)))
        return Polygon(ring)

    def distance(self, other):
        if other.is_empty:
            return float("nan")
        p = (self.x, self.y)
        if isinstance(other, Point):
            return math.hypot(self.x - other.x, self.y - other.y)
        if isinstance(other, Polygon):
            if _point_in_ring(p, other._ring):
                return 0.0
            return min(_pt_seg_dist(p, a, b) for a, b in other._segments())
        if isinstance(other, (MultiPoint, MultiLineString, GeometryCollection)):
            return min(self.distance(g) for g in other.geoms)
        segs = other._segments()
        if segs:
            return min(_pt_seg_dist(p, a, b) for a, b in segs)
        raise NotImplementedError

    def intersects(self, other):
        return other.intersects(self)


def _pt_seg_dist(p, a, b):
    cx, cy = _pt_seg_closest(p[0], p[1], a[0], a[1], b[0], b[1])
    return math.hypot(p[0] - cx, p[1] - cy)


class LineString(BaseGeometry):
    geom_type = "LineString"

    def __init__(self, coords):
        self._c = _CoordSeq(_xy(c) for c in coords)

    @property
    def coords(self):
        return self._c

    @property
    def xy(self):
        return self._c.xy

    def _points(self):
        return list(self._c)

    def _segments(self):
        return [(self._c[i], self._c[i + 1]) for i in range(len(self._c) - 1)]

    @property
    def length(self):
        return sum(math.hypot(b[0] - a[0], b[1] - a[1]) for a, b in self._segments())

    @property
    def is_empty(self):
        return len(self._c) == 0

    # -- buffer ------------------------------------------------------------------------------
    def buffer(self, distance, quad_segs=16, cap_style="round", **_kw):
        pts = [self._c[0]]
        for c in self._c[1:]:
            if c != pts[-1]:
                pts.append(c)
        if len(pts) == 1:  # repeated points removed -> point curve
            return Point(pts[0]).buffer(distance, quad_segs)
        if len(pts) != 2 or cap_style not in ("round", 1):
            raise NotImplementedError("geos_lite buffers 2-point round-capped lines only")
        p0, p1 = pts
        ring = []
        quantum = (math.pi / 2.0) / quad_segs

        def offset(a, b, side):
            dx, dy = b[0] - a[0], b[1] - a[1]
            ln = math.sqrt(dx * dx + dy * dy)
            ux = side * distance * dx / ln
            uy = side * distance * dy / ln
            return (a[0] - uy, a[1] + ux), (b[0] - uy, b[1] + ux)

        def end_cap(a, b):
            o_l = offset(a, b, 1)
            o_r = offset(a, b, -1)
            ang = math.atan2(b[1] - a[1], b[0] - a[0])
            ring.append(o_l[1])
            start, end = ang + math.pi / 2.0, ang - math.pi / 2.0
            total = abs(start - end)
            nseg = int(total / quantum + 0.5)
            inc = total / nseg
            for i in range(nseg):
                t = start - i * inc
                ring.append((b[0] + distance * math.cos(t), b[1] + distance * math.sin(t)))
            ring.append(o_r[1])

        ring.append(offset(p0, p1, 1)[0])
        end_cap(p0, p1)
        ring.append(offset(p1, p0, 1)[0])
        end_cap(p1, p0)
        return Polygon(ring)

    # -- predicates ---------------------------------------------------------------------------
    def intersects(self, other):
        if isinstance(other, Point):
            p = (other.x, other.y)
            return any(_on_segment(p, a, b) for a, b in self._segments())
        if isinstance(other, Polygon):
            for c in self._c:
                if _point_in_ring(c, other._ring):
                    return True
        for a, b in self._segments():
            for c, d in other._segments():
                if _seg_seg(a, b, c, d):
                    return True
        return False

    # -- overlay ------------------------------------------------------------------------------
    def intersection(self, other):
        if isinstance(other, Polygon):
            return _line_clip_polygon(self, other)
        if isinstance(other, LineString):
            return _line_line_intersection(self, other)
        raise NotImplementedError

    # -- linear referencing -----------------------------------------------------------------
    def project(self, point):
        p = (point.x, point.y)
        best = math.inf
        best_len = 0.0
        run = 0.0
        for a, b in self._segments():
            cx, cy = _pt_seg_closest(p[0], p[1], a[0], a[1], b[0], b[1])
            d = math.hypot(p[0] - cx, p[1] - cy)
            if d < best:
                best = d
                best_len = run + math.hypot(cx - a[0], cy - a[1])
            run += math.hypot(b[0] - a[0], b[1] - a[1])
        return best_len

    def interpolate(self, dist):
        total = self.length
        if dist < 0.0:
            dist = total + dist
        if dist <= 0.0:
            return Point(self._c[0])
        if dist >= total:
            return Point(self._c[-1])
        run = 0.0
        for a, b in self._segments():
            seg = math.hypot(b[0] - a[0], b[1] - a[1])
            if run + seg >= dist and seg > 0.0:
                f = (dist - run) / seg
                return Point(a[0] + f * (b[0] - a[0]), a[1] + f * (b[1] - a[1]))
            run += seg
        return Point(self._c[-1])

    def distance(self, other):
        if isinstance(other, Point):
            return other.distance(self)
        return super().distance(other)


class LinearRing(LineString):
    geom_type = "LinearRing"


class Polygon(BaseGeometry):
    geom_type = "Polygon"

    def __init__(self, shell):
        ring = [_xy(c) for c in shell]
        if ring[0] != ring[-1]:
            ring.append(ring[0])
        self._ring = ring

    @property
    def wkb(self):
        """A hashable byte image of the shell (the reference only hashes it to drop duplicate footprints, ATT/grid_env_generation:112-114)."""
        return b"GLP1" + np.asarray(self._ring, dtype=np.float64).tobytes()

    def disjoint(self, other):
        return not self.intersects(other)

    @property
    def exterior(self):
        return LinearRing(self._ring)

    @property
    def boundary(self):
        return LineString(self._ring)

    def _points(self):
        return self._ring

    def _segments(self):
        return [(self._ring[i], self._ring[i + 1]) for i in range(len(self._ring) - 1)]

    @property
    def centroid(self):
        # area-weighted centroid about the first vertex (GEOS Centroid::addShell)
        bx, by = self._ring[0]
        a2 = 0.0
        cx = cy = 0.0
        for i in range(len(self._ring) - 1):
            x0, y0 = self._ring[i][0] - bx, self._ring[i][1] - by
            x1, y1 = self._ring[i + 1][0] - bx, self._ring[i + 1][1] - by
            cr = x0 * y1 - x1 * y0
            a2 += cr
            cx += (x0 + x1) * cr
            cy += (y0 + y1) * cr
        return Point(bx + cx / (3.0 * a2), by + cy / (3.0 * a2))

    def intersects(self, other):
        if isinstance(other, Point):
            return _point_in_ring((other.x, other.y), self._ring)
        if isinstance(other, LineString):
            return other.intersects(self)
        if isinstance(other, Polygon):
            if _point_in_ring(other._ring[0], self._ring) or _point_in_ring(self._ring[0], other._ring):
                return True
            b0, b1 = self.bounds, other.bounds
            if b0[0] > b1[2] or b1[0] > b0[2] or b0[1] > b1[3] or b1[1] > b0[3]:
                return False
            for a, b in self._segments():
                for c, d in other._segments():
                    if _seg_seg(a, b, c, d):
                        return True
            return False
        raise NotImplementedError

    def intersection(self, other):
        """Only emptiness is consumed by the reference (ATT:2245, :2269)."""
        if self.intersects(other):
            return _Region()
        return EmptyGeometry("Polygon")

    def distance(self, other):
        if isinstance(other, Point):
            return other.distance(self)
        return super().distance(other)


class _Region(BaseGeometry):
    """Non-empty areal/lineal/puntal overlay result whose shape is never inspected."""
    geom_type = "Polygon"


class MultiPoint(BaseGeometry):
    geom_type = "MultiPoint"

    def __init__(self, pts):
        self.geoms = [p if isinstance(p, Point) else Point(p) for p in pts]

    def _points(self):
        return [(p.x, p.y) for p in self.geoms]


class MultiLineString(BaseGeometry):
    geom_type = "MultiLineString"

    def __init__(self, lines):
        self.geoms = list(lines)

    @property
    def length(self):
        return sum(g.length for g in self.geoms)

    def _segments(self):
        return [s for g in self.geoms for s in g._segments()]

    def _points(self):
        return [c for g in self.geoms for c in g._points()]

    def interpolate(self, dist):
        run = 0.0
        for g in self.geoms:
            if dist <= run + g.length:
                return g.interpolate(dist - run)
            run += g.length
        return Point(self.geoms[-1].coords[-1])


class GeometryCollection(BaseGeometry):
    geom_type = "GeometryCollection"

    def __init__(self, geoms=()):
        self.geoms = list(geoms)

    @property
    def is_empty(self):
        return len(self.geoms) == 0

    def _segments(self):
        return [s for g in self.geoms for s in g._segments()]


def _dedupe(pts):
    out = []
    for p in pts:
        if p not in out:
            out.append(p)
    return out


def _line_line_intersection(l1, l2):
    pts, overlaps = [], []
    for a, b in l1._segments():
        for c, d in l2._segments():
            r = _seg_seg(a, b, c, d)
            if len(r) == 1:
                pts.append(r[0])
            elif len(r) == 2:
                overlaps.append(LineString(r))
    pts = _dedupe(pts)
    # points that lie on an overlap piece are absorbed by it
    pts = [p for p in pts if not any(_on_segment(p, o.coords[0], o.coords[1]) for o in overlaps)]
    if overlaps and not pts:
        return overlaps[0] if len(overlaps) == 1 else MultiLineString(overlaps)
    if overlaps:
        return GeometryCollection([Point(p) for p in pts] + overlaps)
    if not pts:
        return EmptyGeometry("LineString")
    if len(pts) == 1:
        return Point(pts[0])
    return MultiPoint(pts)


def _line_clip_polygon(line, poly):
    """LineString n Polygon -> Point | LineString | MultiLineString | empty (OverlayNG result shape)."""
    pieces = []  # list of coordinate lists
    lone_pts = []
    for a, b in line._segments():
        dx, dy = b[0] - a[0], b[1] - a[1]
        ts = [0.0, 1.0]
        for c, d in poly._segments():
            for r in _seg_seg(a, b, c, d):
                if dx * dx + dy * dy > 0.0:
                    ts.append(((r[0] - a[0]) * dx + (r[1] - a[1]) * dy) / (dx * dx + dy * dy))
        ts = sorted(set(min(max(t, 0.0), 1.0) for t in ts))
        cur = None
        for t0, t1 in zip(ts[:-1], ts[1:]):
            tm = 0.5 * (t0 + t1)
            mid = (a[0] + tm * dx, a[1] + tm * dy)
            if _point_in_ring(mid, poly._ring):
                if cur is not None and cur[1] == t0:
                    cur[1] = t1
                else:
                    cur = [t0, t1]
                    pieces.append(cur)
                    cur_owner = (a, dx, dy)
                    cur.append(cur_owner)
            else:
                cur = None
        # isolated touch points (segment grazes a vertex / edge endpoint)
        for t in ts:
            pt = (a[0] + t * dx, a[1] + t * dy)
            if _point_in_ring(pt, poly._ring) and not any(p[0] <= t <= p[1] and p[2][0] == a for p in pieces):
                lone_pts.append(pt)
    lines = []
    for t0, t1, (a, dx, dy) in pieces:
        c0 = (a[0] + t0 * dx, a[1] + t0 * dy)
        c1 = (a[0] + t1 * dx, a[1] + t1 * dy)
        if lines and lines[-1][-1] == c0:
            lines[-1].append(c1)
        else:
            lines.append([c0, c1])
    lone_pts = [p for p in _dedupe(lone_pts) if not any(p in ln for ln in lines)]
    if not lines:
        if not lone_pts:
            return EmptyGeometry("LineString")
        return Point(lone_pts[0]) if len(lone_pts) == 1 else MultiPoint(lone_pts)
    geoms = [LineString(ln) for ln in lines]
    if lone_pts:
        return GeometryCollection([Point(p) for p in lone_pts] + geoms)
    return geoms[0] if len(geoms) == 1 else MultiLineString(geoms)


def nearest_points(g1, g2):
    """shapely.ops.nearest_points for the Point/LineString/Multi* combinations the path uses."""
    if isinstance(g1, Point) and isinstance(g2, Point):
        return (g1, g2)
    if isinstance(g1, Point):
        p = (g1.x, g1.y)
        best, best_pt = math.inf, None
        for a, b in g2._segments():
            cx, cy = _pt_seg_closest(p[0], p[1], a[0], a[1], b[0], b[1])
            d = math.hypot(p[0] - cx, p[1] - cy)
            if d < best:
                best, best_pt = d, (cx, cy)
        for q in (g2.geoms if isinstance(g2, (MultiPoint, GeometryCollection)) else ()):
            if isinstance(q, Point):
                d = g1.distance(q)
                if d < best:
                    best, best_pt = d, (q.x, q.y)
        return (g1, Point(best_pt))
    if isinstance(g2, Point):
        b, a = nearest_points(g2, g1)
        return (a, b)
    raise NotImplementedError


class _GeomArray(np.ndarray):
    pass


class STRtree:
    def __init__(self, geoms):
        self._geoms = list(geoms)
        self.geometries = np.empty(len(self._geoms), dtype=object)
        for i, g in enumerate(self._geoms):
            self.geometries[i] = g
        self._bounds = [g.bounds for g in self._geoms]

    def query(self, geom, predicate=None):
        q = geom.bounds
        out = [i for i, b in enumerate(self._bounds)
               if not (b[0] > q[2] or q[0] > b[2] or b[1] > q[3] or q[1] > b[3])]
        return np.array(out, dtype=np.int64)

    def nearest(self, geom):
        best, best_i = math.inf, None
        for i, g in enumerate(self._geoms):
            d = geom.distance(g)
            if d < best:
                best, best_i = d, i
        return best_i


def install_as_shapely():
    """Register this module under the shapely names the reference imports (ATT:12-16,27)."""
    import sys
    import types

    me = sys.modules[__name__]
    root = types.ModuleType("shapely")
    root.__version__ = "2.0.1-geos_lite"
    geometry = types.ModuleType("shapely.geometry")
    for name in ("Point", "LineString", "LinearRing", "Polygon", "MultiPoint", "MultiLineString",
                 "GeometryCollection"):
        setattr(geometry, name, getattr(me, name))
        setattr(root, name, getattr(me, name))
    strtree = types.ModuleType("shapely.strtree")
    strtree.STRtree = STRtree
    ops = types.ModuleType("shapely.ops")
    ops.nearest_points = nearest_points
    affinity = types.ModuleType("shapely.affinity")
    affinity.scale = lambda g, *a, **k: g
    wkb = types.ModuleType("shapely.wkb")
    wkb.loads = lambda b: Polygon(np.frombuffer(b[4:], dtype=np.float64).reshape(-1, 2).tolist())
    point_mod = types.ModuleType("shapely.geometry.point")   # `from shapely.geometry.point import Point` (ATT/grid_env_generation:20)
    point_mod.Point = Point
    geometry.__path__ = []
    geometry.point = point_mod
    root.geometry, root.strtree, root.ops, root.affinity, root.wkb = geometry, strtree, ops, affinity, wkb
    sys.modules.update({"shapely": root, "shapely.geometry": geometry, "shapely.geometry.point": point_mod, "shapely.strtree": strtree,
                        "shapely.ops": ops, "shapely.affinity": affinity, "shapely.wkb": wkb})
    return root
