"""Golden occupancy grid from the UNMODIFIED reference `env_generation` (ATT = MADDPG_ownENV_randomOD_radar_one_model_att/
grid_env_generation_newframe_randomOD_radar_sur_drones_oneModel_att.py:108-185), build container only.

geopandas is absent: `gpd.read_file` is stubbed with a function that returns the pandas frame geopandas would build from
the shapefile (7 columns, height in column 2, geometry in column 6 - the positions the reference indexes), fed by this
repo's own .shp / .dbf reader; shapely is the real one when importable, else oracle/geos_lite.py (recorded).

    python tests/golden/gen_golden_mapgen.py        # writes tests/golden/mapgen_ref.{shp,dbf,npz}
"""
import importlib
import os
import struct
import sys
from unittest import mock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as H  # noqa: E402
from multi_agent_aac_b200 import maps  # noqa: E402

BOUND = [455, 680, 255, 385]


def write_shapefile(base, rings, rows, fields):
    """A polygon shapefile (type 5, one part per record) + dbf with numeric / text fields: just enough for the test data."""
    recs = []
    for k, ring in enumerate(rings):
        pts = np.asarray(ring, dtype="<f8")
        body = struct.pack("<i4d2i", 5, pts[:, 0].min(), pts[:, 1].min(), pts[:, 0].max(), pts[:, 1].max(), 1, len(pts)) + struct.pack("<i", 0) + pts.tobytes()
        recs.append(struct.pack(">ii", k + 1, len(body) // 2) + body)
    allp = np.concatenate([np.asarray(r) for r in rings])
    total = 100 + sum(len(r) for r in recs)
    hdr = struct.pack(">i5ii", 9994, 0, 0, 0, 0, 0, total // 2) + struct.pack("<ii4d4d", 1000, 5, allp[:, 0].min(), allp[:, 1].min(), allp[:, 0].max(), allp[:, 1].max(), 0, 0, 0, 0)
    open(base + ".shp", "wb").write(hdr + b"".join(recs))
    rec_len = 1 + sum(f[2] for f in fields)
    dh = struct.pack("<BBBBIHH20x", 3, 124, 1, 1, len(rows), 32 + 32 * len(fields) + 1, rec_len)
    for name, ftype, flen, dec in fields:
        dh += name.encode().ljust(11, b"\x00") + ftype.encode() + b"\x00" * 4 + struct.pack("<BB", flen, dec) + b"\x00" * 14
    body = b""
    for row in rows:
        body += b" "
        for (name, ftype, flen, dec), v in zip(fields, row):
            body += (("%*.*f" % (flen, dec, v)) if ftype == "N" else str(v).ljust(flen)[:flen]).encode()
    open(base + ".dbf", "wb").write(dh + b"\x0d" + body + b"\x1a")


def test_data():
    rng = np.random.default_rng(7)
    rings, rows = [], []
    inv = lambda xm, ym: (xm / ((maps.SVY21_X[1] - maps.SVY21_X[0]) / maps.SVY21_X[2]) + maps.SVY21_X[0],
                          ym / ((maps.SVY21_Y[1] - maps.SVY21_Y[0]) / maps.SVY21_Y[2]) + maps.SVY21_Y[0])
    for k in range(22):
        cx, cy = rng.uniform(440, 700), rng.uniform(240, 400)
        ang = np.sort(rng.uniform(0, 2 * np.pi, rng.integers(3, 9)))
        rad = rng.uniform(3, 26, len(ang))
        xm, ym = cx + rad * np.cos(ang), cy + rad * np.sin(ang)
        x, y = inv(np.append(xm, xm[0]), np.append(ym, ym[0]))
        rings.append(np.stack([x, y], -1))
        rows.append([k, "B%02d" % k, float(rng.choice([0.0, 7.5, 12.0, 33.0])), 1.0, 2.0, "x"])
    rings.append(rings[3].copy()); rows.append(list(rows[3]))            # a duplicate footprint (dropped, :110-115)
    for ring_m in ([(560, 330), (620, 330), (620, 340), (560, 340)], [(560, 370), (620, 370), (620, 380), (560, 380)],
                   [(560, 330), (570, 330), (570, 380), (560, 380)], [(610, 330), (620, 330), (620, 380), (610, 380)]):   # a courtyard: filled (:156)
        xm, ym = np.array([p[0] for p in ring_m + [ring_m[0]]], float), np.array([p[1] for p in ring_m + [ring_m[0]]], float)
        x, y = inv(xm, ym)
        rings.append(np.stack([x, y], -1)); rows.append([99, "ring", 20.0, 1.0, 2.0, "x"])
    fields = [("ID", "N", 6, 0), ("NAME", "C", 8, 0), ("HEIGHT", "N", 12, 3), ("A", "N", 6, 1), ("B", "N", 6, 1), ("C", "C", 4, 0)]
    return rings, rows, fields


def run_reference(shp):
    import pandas as pd
    H._install_stubs()
    geom = sys.modules["shapely.geometry"]

    def read_file(path):
        rings, rows, names = maps.read_shapefile(path)
        return pd.DataFrame([list(r) + [geom.Polygon([tuple(p) for p in ring])] for ring, r in zip(rings, rows)], columns=names + ["geometry"])
    gpd = mock.MagicMock(name="geopandas")
    gpd.read_file = read_file
    sys.modules["geopandas"] = gpd
    d = os.path.join(H.REFERENCE, H.VARIANTS["att"][0])
    sys.path.insert(0, d)
    try:
        mod = importlib.import_module("grid_env_generation_newframe_randomOD_radar_sur_drones_oneModel_att")
    finally:
        sys.path.remove(d)
    # the reference indexes rows positionally (`row[2]`, `row[6]`), which pandas < 3 allowed on a labelled Series: emulate that
    class PosRow:
        def __init__(self, row):
            self.row = row

        def __getitem__(self, k):
            return self.row.iloc[k] if isinstance(k, int) else self.row[k]
    orig = pd.DataFrame.iterrows
    pd.DataFrame.iterrows = lambda self: ((i, PosRow(r)) for i, r in orig(self))
    try:
        env_map_bounded, polys, grid_length, out_poly, extent = mod.env_generation(shp, BOUND)
    finally:
        pd.DataFrame.iterrows = orig
    ones = sorted((int(round(p.centroid.x)), int(round(p.centroid.y))) for p in out_poly[0][0])
    zeros = sorted((int(round(p.centroid.x)), int(round(p.centroid.y))) for p in out_poly[0][1])
    return np.asarray(env_map_bounded), ones, zeros, grid_length, extent


def main():
    base = os.path.join(HERE, "mapgen_ref")
    rings, rows, fields = test_data()
    write_shapefile(base, rings, rows, fields)
    emb, ones, zeros, g, extent = run_reference(base + ".shp")
    np.savez_compressed(base + ".npz", bound=np.array(BOUND), env_map_bounded=emb.astype(np.uint8), ones=np.array(ones), zeros=np.array(zeros),
                        geometry=np.array(H.GEOMETRY))
    print("wrote", base + ".npz", "occupied cells", len(ones), "free", len(zeros), "geometry:", H.GEOMETRY)


if __name__ == "__main__":
    main()
