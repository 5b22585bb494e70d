"""Static world: 10 m occupancy grids, bounds and quadrant target pools.

The reference derives its grid from a Singapore shapefile that is not in the repository
(ATT/parameters:39, ATT/grid_env_generation:108-185; ATT = MADDPG_ownENV_randomOD_radar_one_model_att),
so maps here are synthetic but follow the reference's grid conventions exactly:

* cell centres sit on multiples of `grid_length` inside the closed bound
  (ATT/grid_env_generation:169-174), each cell is the axis-aligned square centre +/- grid_length/2;
* occupancy is indexed [ix, iy], ix-major -- the order in which the reference appends cells to
  `world_map_2D_polyList[0][0]` / `[0][1]` (ATT/grid_env_generation:169-179);
* free cells whose centre lies on a boundary line are not OD candidates, the rest fall into four
  quadrant pools split at the bound's mid lines (ATT/env_simulator:143-197).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

# bound table of the multipleMap variant (MM/parameters_randomOD_radar_multipleMap.py:52-55)
MULTIMAP_BOUNDS = [
    [0, 250, 550, 700], [230, 530, 1000, 1200], [815, 1015, 270, 385], [455, 680, 255, 385],
    [250, 450, 260, 385], [585, 695, 165, 300], [1395, 1535, 615, 715], [815, 1000, 950, 1055],
    [1005, 1155, 535, 620], [1535, 1675, 225, 345], [905, 1085, 105, 205], [1105, 1195, 385, 515],
    [715, 845, 255, 355], [685, 825, 595, 705],
]
DEFAULT_BOUND = [455, 680, 255, 385]  # ATT/parameters:32-36


@dataclass
class GridMap:
    bound: list            # [xmin, xmax, ymin, ymax]
    grid_length: int       # 10
    occ: np.ndarray        # uint8 [GX, GY], 1 = occupied
    x0c: float = field(init=False)   # centre of cell column 0
    y0c: float = field(init=False)

    def __post_init__(self):
        g = self.grid_length
        self.x0c = float(math.ceil(self.bound[0] / g) * g)
        self.y0c = float(math.ceil(self.bound[2] / g) * g)
        gx, gy = grid_shape(self.bound, g)
        assert self.occ.shape == (gx, gy), (self.occ.shape, gx, gy)

    @property
    def gx(self):
        return self.occ.shape[0]

    @property
    def gy(self):
        return self.occ.shape[1]

    @property
    def origin(self):
        """Local-coordinate origin used by the kernels (bound centre)."""
        return (0.5 * (self.bound[0] + self.bound[1]), 0.5 * (self.bound[2] + self.bound[3]))

    def cell_centre(self, ix, iy):
        return (self.x0c + ix * self.grid_length, self.y0c + iy * self.grid_length)

    def target_pools(self):
        """Four lists of free-cell centres (int tuples), reference order (ATT/env_simulator:154-197)."""
        xmin, xmax, ymin, ymax = self.bound
        xs = (xmax - xmin) / 2 + xmin
        ys = (ymax - ymin) / 2 + ymin
        pools = [[], [], [], []]
        for ix in range(self.gx):
            for iy in range(self.gy):
                if self.occ[ix, iy]:
                    continue
                cx, cy = self.cell_centre(ix, iy)
                if cx == xmin or cx == xmax or cy == ymin or cy == ymax:
                    continue
                c = (int(cx), int(cy))
                if cx < xs and cy < ys:
                    pools[0].append(c)
                elif cx > xs and cy < ys:
                    pools[1].append(c)
                elif cx > xs and cy > ys:
                    pools[2].append(c)
                else:
                    pools[3].append(c)
        return pools

    def cell_of(self, cx, cy):
        g = self.grid_length
        return (int(round((cx - self.x0c) / g)), int(round((cy - self.y0c) / g)))


def grid_shape(bound, grid_length=10):
    g = grid_length
    ix0, ix1 = math.ceil(bound[0] / g), math.floor(bound[1] / g)
    iy0, iy1 = math.ceil(bound[2] / g), math.floor(bound[3] / g)
    return ix1 - ix0 + 1, iy1 - iy0 + 1


def _fill_holes(occ):
    """scipy.ndimage.binary_fill_holes without scipy: free cells not 4-connected to the border fill."""
    gx, gy = occ.shape
    reach = np.zeros_like(occ, dtype=bool)
    stack = [(i, j) for i in range(gx) for j in (0, gy - 1)] + [(i, j) for j in range(gy) for i in (0, gx - 1)]
    while stack:
        i, j = stack.pop()
        if i < 0 or j < 0 or i >= gx or j >= gy or reach[i, j] or occ[i, j]:
            continue
        reach[i, j] = True
        stack += [(i + 1, j), (i - 1, j), (i, j + 1), (i, j - 1)]
    return np.where(reach, 0, 1).astype(np.uint8)


def synthetic_map(bound=None, seed=0, grid_length=10, fill=0.25):
    """Random-rectangle buildings, ~`fill` occupied, border ring free, holes filled.

    The border ring stays free so every free cell is 4-connected (the reference's A*,
    ATT/jps_straight.py:17-70, returns None on an unreachable goal) and all four quadrant
    pools are non-empty."""
    bound = list(DEFAULT_BOUND if bound is None else bound)
    gx, gy = grid_shape(bound, grid_length)
    rng = np.random.default_rng(seed)
    occ = np.zeros((gx, gy), dtype=np.uint8)
    target = fill * gx * gy
    guard = 0
    while occ.sum() < target and guard < 1000:
        guard += 1
        w, h = int(rng.integers(1, 4)), int(rng.integers(1, 4))
        if gx - 2 - w < 1 or gy - 2 - h < 1:
            w = h = 1
        i = int(rng.integers(1, gx - 1 - w + 1))
        j = int(rng.integers(1, gy - 1 - h + 1))
        occ[i:i + w, j:j + h] = 1
    occ[0, :] = occ[-1, :] = 0
    occ[:, 0] = occ[:, -1] = 0
    occ = _fill_holes(occ)
    m = GridMap(bound, grid_length, occ)
    assert all(len(p) > 0 for p in m.target_pools())
    return m


def multimap_set(seed=0, grid_length=10):
    """14 synthetic maps on the multipleMap variant's bound table (MM/parameters:52-55)."""
    return [synthetic_map(b, seed=seed * 100 + k, grid_length=grid_length) for k, b in enumerate(MULTIMAP_BOUNDS)]


# ---------------------------------------------------------------------------------------------------------------------
# Map ingestion: building footprints -> occupancy grid, as ATT/grid_env_generation:108-185 does it.

def _polygon_touches_squares(poly, cx, cy, h):
    """Closed-set test `not polygon.disjoint(square)` (ATT/grid_env_generation:83,90) of one simple polygon (vertex array
    [V, 2], closed or open ring) against the axis-aligned squares centred at (cx[k], cy[k]) with half side h: a polygon
    vertex in a square, a square corner in the polygon, or a polygon edge crossing a square edge."""
    p = np.asarray(poly, dtype=np.float64)
    if np.allclose(p[0], p[-1]):
        p = p[:-1]
    a, b = p, np.roll(p, -1, axis=0)                                    # edges a -> b
    cx, cy = np.asarray(cx, np.float64)[:, None], np.asarray(cy, np.float64)[:, None]
    hit = ((np.abs(p[None, :, 0] - cx) <= h) & (np.abs(p[None, :, 1] - cy) <= h)).any(1)

    def inside(x, y):                                                    # even-odd rule, boundary counts
        x, y = x[:, None], y[:, None]
        ax, ay, bx, by = a[None, :, 0], a[None, :, 1], b[None, :, 0], b[None, :, 1]
        cross = (bx - ax) * (y - ay) - (by - ay) * (x - ax)
        on = (cross == 0) & (x >= np.minimum(ax, bx)) & (x <= np.maximum(ax, bx)) & (y >= np.minimum(ay, by)) & (y <= np.maximum(ay, by))
        with np.errstate(divide="ignore", invalid="ignore"):
            xi = ax + (y - ay) * (bx - ax) / (by - ay)
        crossing = ((ay > y) != (by > y)) & (x < xi)
        return (crossing.sum(1) % 2 == 1) | on.any(1)

    def seg_hits(x0, y0, x1, y1):                                        # closed segments (x0,y0)-(x1,y1) vs every polygon edge
        x0, y0, x1, y1 = (v[:, None] for v in (x0, y0, x1, y1))
        ax, ay, bx, by = a[None, :, 0], a[None, :, 1], b[None, :, 0], b[None, :, 1]
        o = lambda px, py, qx, qy, rx, ry: np.sign((qx - px) * (ry - py) - (qy - py) * (rx - px))
        d1, d2 = o(x0, y0, x1, y1, ax, ay), o(x0, y0, x1, y1, bx, by)
        d3, d4 = o(ax, ay, bx, by, x0, y0), o(ax, ay, bx, by, x1, y1)
        proper = (d1 * d2 < 0) & (d3 * d4 < 0)
        between = lambda px, py, qx, qy, rx, ry: (np.minimum(px, qx) <= rx) & (rx <= np.maximum(px, qx)) & (np.minimum(py, qy) <= ry) & (ry <= np.maximum(py, qy))
        touch = ((d1 == 0) & between(x0, y0, x1, y1, ax, ay)) | ((d2 == 0) & between(x0, y0, x1, y1, bx, by)) | \
                ((d3 == 0) & between(ax, ay, bx, by, x0, y0)) | ((d4 == 0) & between(ax, ay, bx, by, x1, y1))
        return (proper | touch).any(1)

    cxf, cyf = cx[:, 0], cy[:, 0]
    for sx, sy in ((-h, -h), (h, -h), (h, h), (-h, h)):
        hit |= inside(cxf + sx, cyf + sy)
    corners = [(-h, -h), (h, -h), (h, h), (-h, h), (-h, -h)]
    for (x0, y0), (x1, y1) in zip(corners[:-1], corners[1:]):
        hit |= seg_hits(cxf + x0, cyf + y0, cxf + x1, cyf + y1)
    return hit


def gridmap_from_polygons(polygons, bound, grid_length=10, extent=(1800, 1300), heights=None):
    """Building footprints (vertex lists in metres) -> GridMap, the way the reference builds `world_map_2D`
    (ATT/grid_env_generation:140-171): grid points at multiples of `grid_length` over `extent`, a cell (the square of
    half side grid_length / 2 around its grid point) is occupied when it is not disjoint from some footprint, holes
    enclosed by occupied cells are filled (`ndimage.binary_fill_holes`), and the bounded map is the cells whose
    grid point lies inside the closed bound.  `heights` (one per footprint): the reference fills the cell's column up to
    ceil(mean height of the footprints it meets / grid_length) layers and reads layer 0 (:150-153), so a cell counts
    only when that mean height is positive; None = every footprint counts."""
    from scipy import ndimage
    g, h = grid_length, grid_length / 2.0
    nx, ny = int(math.ceil(extent[0] / g)), int(math.ceil(extent[1] / g))   # initialize_3d_array_environment's x / y extents
    if heights is not None:
        hsum, hcnt = np.zeros((nx, ny)), np.zeros((nx, ny), dtype=np.int64)
        for poly, hgt in zip(polygons, heights):
            p = np.asarray(poly, dtype=np.float64)
            ix0, ix1 = max(int(math.floor((p[:, 0].min() - h) / g)), 0), min(int(math.ceil((p[:, 0].max() + h) / g)), nx - 1)
            iy0, iy1 = max(int(math.floor((p[:, 1].min() - h) / g)), 0), min(int(math.ceil((p[:, 1].max() + h) / g)), ny - 1)
            if ix1 < ix0 or iy1 < iy0:
                continue
            ix, iy = np.meshgrid(np.arange(ix0, ix1 + 1), np.arange(iy0, iy1 + 1), indexing="ij")
            hit = _polygon_touches_squares(p, ix.ravel() * g, iy.ravel() * g, h)
            np.add.at(hsum, (ix.ravel()[hit], iy.ravel()[hit]), float(hgt))
            np.add.at(hcnt, (ix.ravel()[hit], iy.ravel()[hit]), 1)
        env = (hcnt > 0) & (np.ceil(np.divide(hsum, np.maximum(hcnt, 1)) / g) >= 1)
        polygons = []
    else:
        env = np.zeros((nx, ny), dtype=bool)
    for poly in polygons:
        p = np.asarray(poly, dtype=np.float64)
        ix0, ix1 = max(int(math.floor((p[:, 0].min() - h) / g)), 0), min(int(math.ceil((p[:, 0].max() + h) / g)), nx - 1)
        iy0, iy1 = max(int(math.floor((p[:, 1].min() - h) / g)), 0), min(int(math.ceil((p[:, 1].max() + h) / g)), ny - 1)
        if ix1 < ix0 or iy1 < iy0:
            continue
        ix, iy = np.meshgrid(np.arange(ix0, ix1 + 1), np.arange(iy0, iy1 + 1), indexing="ij")
        hit = _polygon_touches_squares(p, ix.ravel() * g, iy.ravel() * g, h)
        env[ix.ravel()[hit], iy.ravel()[hit]] = True
    env = ndimage.binary_fill_holes(env)
    # the cells of the bounded map: grid points inside the closed bound, as the reference lists them for its occupied /
    # free polygons (ATT/grid_env_generation:166-176; its `env_map_bounded` slice stops one index short of a bound that
    # falls on a grid point, the polygon lists do not)
    xl, yl = math.ceil(bound[0] / g), math.ceil(bound[2] / g)
    gx, gy = grid_shape(bound, g)
    occ = np.zeros((gx, gy), dtype=np.uint8)
    sub = env[xl:xl + gx, yl:yl + gy]
    occ[:sub.shape[0], :sub.shape[1]] = sub
    return GridMap(list(bound), g, occ)


# ---- shapefile ingestion (ATT/grid_env_generation:108-134; the reference reads it with geopandas) ---------------------

SVY21_X = (14550.0, 16262.89690000005, 1800.0)   # min, max, span of the reference's area in SVY21 easting  (ATT/grid_env_generation:127)
SVY21_Y = (36200.0, 37448.60029999912, 1300.0)   # ... northing                                            (:128)


def svy21_to_metres(x, y):
    """The reference's `coordinate_to_meter` (ATT/grid_env_generation:27-31) with its hard-coded area: note that it
    multiplies by (max - min) / span, as the source does."""
    return ((np.asarray(x, dtype=np.float64) - SVY21_X[0]) * ((SVY21_X[1] - SVY21_X[0]) / SVY21_X[2]),
            (np.asarray(y, dtype=np.float64) - SVY21_Y[0]) * ((SVY21_Y[1] - SVY21_Y[0]) / SVY21_Y[2]))


def read_shapefile(path):
    """Minimal ESRI shapefile reader for building footprints: polygon records (shape types 5 / 15 / 25) of `path`.shp and
    the attribute rows of the .dbf next to it.  Returns (rings, rows, field_names): rings[k] = exterior ring of record k
    as an [n, 2] array (the first part, which is what the reference uses: `row[6].exterior`), rows[k] = its attribute
    values (numbers parsed, text stripped)."""
    import struct
    base = path[:-4] if path.lower().endswith(".shp") else path
    raw = open(base + ".shp", "rb").read()
    if struct.unpack(">i", raw[:4])[0] != 9994:
        raise ValueError("%s.shp is not a shapefile" % base)
    rings, pos = [], 100
    while pos + 8 <= len(raw):
        _, words = struct.unpack(">ii", raw[pos:pos + 8])
        rec = raw[pos + 8:pos + 8 + 2 * words]
        pos += 8 + 2 * words
        stype = struct.unpack("<i", rec[:4])[0]
        if stype == 0:
            rings.append(np.zeros((0, 2)))
            continue
        if stype not in (5, 15, 25):
            raise ValueError("shape type %d: only polygon shapefiles are supported" % stype)
        n_parts, n_points = struct.unpack("<ii", rec[36:44])
        parts = list(struct.unpack("<%di" % n_parts, rec[44:44 + 4 * n_parts])) + [n_points]
        pts = np.frombuffer(rec, dtype="<f8", count=2 * n_points, offset=44 + 4 * n_parts).reshape(n_points, 2)
        rings.append(np.array(pts[parts[0]:parts[1]], dtype=np.float64))
    rows, names = [], []
    try:
        dbf = open(base + ".dbf", "rb").read()
    except FileNotFoundError:
        return rings, [[] for _ in rings], names
    n_rec, hdr_len, rec_len = struct.unpack("<IHH", dbf[4:12])
    fields, off = [], 32
    while dbf[off] != 0x0D:
        name = dbf[off:off + 11].split(b"\x00")[0].decode("latin1")
        fields.append((name, chr(dbf[off + 11]), dbf[off + 16]))
        off += 32
    names = [f[0] for f in fields]
    for k in range(n_rec):
        rec = dbf[hdr_len + k * rec_len:hdr_len + (k + 1) * rec_len]
        vals, o = [], 1
        for _, ftype, flen in fields:
            txt = rec[o:o + flen].decode("latin1").strip()
            o += flen
            if ftype in "NF":
                try:
                    vals.append(float(txt) if txt else 0.0)
                except ValueError:
                    vals.append(0.0)
            else:
                vals.append(txt)
        rows.append(vals)
    return rings, rows, names


def gridmap_from_shapefile(path, bound, height_field=2, grid_length=10, extent=(1800, 1300), transform=svy21_to_metres):
    """`env_generation(shapeFilePath, bound)` (ATT/grid_env_generation:108-185) without geopandas: read the footprints,
    drop duplicate geometries (:110-115), convert SVY21 to the reference's metre frame (:124-130), take the height from
    attribute column `height_field` (the reference's `row[2]`), rasterise and fill holes (`gridmap_from_polygons`)."""
    rings, rows, _ = read_shapefile(path)
    seen, polys, heights = set(), [], []
    for ring, row in zip(rings, rows):
        if len(ring) < 3:
            continue
        key = ring.tobytes()
        if key in seen:
            continue
        seen.add(key)
        x, y = transform(ring[:, 0], ring[:, 1]) if transform else (ring[:, 0], ring[:, 1])
        polys.append(np.stack([x, y], -1))
        heights.append(float(row[height_field]) if len(row) > height_field else 1.0)
    return gridmap_from_polygons(polys, bound, grid_length, extent, heights=heights)
