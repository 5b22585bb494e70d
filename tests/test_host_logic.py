"""CPU: host-side logic (maps, reset planner, scenario bank), the C-ABI library's exports, the
multi-rank statistics reduction (gloo, world_size 2).  No CUDA calls."""
import ctypes
import os
import random
import re
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from multi_agent_aac_b200 import _capi
from multi_agent_aac_b200.maps import MULTIMAP_BOUNDS, GridMap, grid_shape, multimap_set, synthetic_map
from multi_agent_aac_b200.reset import ScenarioBank, plan_path, prune_collinear, sample_episode_reference_order
from multi_agent_aac_b200.stats import reduce_episode_stats, shard_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_default_grid_shape_matches_survey():
    assert grid_shape([455, 680, 255, 385]) == (23, 13)      # 23 x 13 = 299 cells
    m = synthetic_map(seed=0)
    assert m.occ.shape == (23, 13) and 0.1 < m.occ.mean() < 0.5
    assert m.occ[0].sum() == 0 and m.occ[:, 0].sum() == 0     # border ring free
    assert all(len(p) > 0 for p in m.target_pools())


def test_multimap_set_fits_device_limits():
    maps = multimap_set(seed=0)
    assert len(maps) == len(MULTIMAP_BOUNDS) == 14
    for m in maps:
        assert m.gx * m.gy <= _capi.MAP_STRIDE and (m.gx + 8) * (m.gy + 8) + 32 <= 2048


def test_plan_path_and_prune():
    occ = np.zeros((6, 5), dtype=np.uint8)
    occ[2, 0:4] = 1
    path = plan_path(occ, (0, 0), (5, 0))
    assert path[0] == (0, 0) and path[-1] == (5, 0)
    assert all(abs(a[0] - b[0]) + abs(a[1] - b[1]) == 1 for a, b in zip(path, path[1:]))
    assert all(occ[c] == 0 for c in path)
    pruned = prune_collinear(path)
    assert pruned[0] == path[0] and pruned[-1] == path[-1]
    for a, b in zip(pruned, pruned[1:]):
        assert a[0] == b[0] or a[1] == b[1]
    occ[:, 4] = 0
    occ[2, :] = 1
    assert plan_path(occ, (0, 0), (5, 0)) is None


def test_reference_order_reset_is_seeded_and_separated():
    m = synthetic_map(seed=0)
    a = sample_episode_reference_order(random.Random(5), m, 4)
    b = sample_episode_reference_order(random.Random(5), m, 4)
    assert a.starts == b.starts and a.cells == b.cells
    for i in range(4):
        for j in range(i):
            assert np.hypot(a.starts[i][0] - a.starts[j][0], a.starts[i][1] - a.starts[j][1]) > 5.0
    for line, c in zip(a.lines, a.cells):
        assert len(line) == len(c) >= 2


def test_scenario_bank_packing():
    m = synthetic_map(seed=0)
    bank = ScenarioBank(m, 3, 16, w_max=32, seed=1)
    assert bank.cells.shape == (16, 3, 32) and bank.cells.dtype == np.uint16 and bank.w.min() >= 2
    for s in range(16):
        for i in range(3):
            w = bank.w[s, i]
            ix, iy = bank.cells[s, i, :w] >> 8, bank.cells[s, i, :w] & 255
            assert (m.occ[ix, iy] == 0).all()
            assert ((np.diff(ix.astype(int)) == 0) | (np.diff(iy.astype(int)) == 0)).all()


def test_library_exports_every_declared_symbol():
    _capi.build()
    header = open(os.path.join(ROOT, "include", "aac_env.h")).read()
    declared = set(re.findall(r"\b(aac_[a-z_]+)\s*\(", header))
    assert declared == set(_capi.EXPORTS), declared ^ set(_capi.EXPORTS)
    lib = ctypes.CDLL(_capi.LIB_PATH)
    for name in declared:
        assert getattr(lib, name) is not None
    lib.aac_own_dim.argtypes = [ctypes.c_int32, ctypes.c_int32]
    assert lib.aac_own_dim(0, 3) == 14 and lib.aac_own_dim(1, 10) == 7      # pure host helper
    lib.aac_last_error.restype = ctypes.c_char_p
    h = ctypes.c_void_p()
    cfg = _capi.AacConfig()
    cfg.abi_version = 999
    lib.aac_create.argtypes = [ctypes.POINTER(_capi.AacConfig), ctypes.POINTER(ctypes.c_void_p)]
    assert lib.aac_create(ctypes.byref(cfg), ctypes.byref(h)) == -1 and b"abi_version" in lib.aac_last_error()


def test_struct_sizes_match_header_layout():
    # ABI 3: + radar_targets, n_nbr_obs, n_clouds, clouds[8][6]; 300 bytes padded to the int64 members' alignment
    assert ctypes.sizeof(_capi.AacConfig) == (12 * 4 + 16 + 6 * 4 + 2 * 4 + 3 * 4 + 8 * 6 * 4 + 7) // 8 * 8
    assert ctypes.sizeof(_capi.AacState) == 14 * 8 and ctypes.sizeof(_capi.AacOut) == 20 * 8
    assert ctypes.sizeof(_capi.AacMapDesc) == 2 * 4 + 4 * 4 + 5 * 4


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "multi_agent_aac_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("# oracle", ""), f


def test_env_requires_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from multi_agent_aac_b200.env import BatchedDroneEnv, preset
    with pytest.raises(_capi.AacError):
        BatchedDroneEnv(preset("att"), synthetic_map(seed=0))


def test_shard_range_partitions():
    for world in (1, 2, 3, 8):
        spans = [shard_range(r, world, 65536 + 5) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == 65541
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def _stats_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    local = np.zeros(_capi.N_STATS)
    local[0], local[1], local[2], local[3 + rank] = 10 * (rank + 1), 500.0, -3.0 * (rank + 1), 4
    lo, hi = shard_range(rank, world, 1001)
    out = reduce_episode_stats(local)
    out["span"] = (lo, hi)
    q.put((rank, out))
    dist.destroy_process_group()


def test_episode_stats_allreduce_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_stats_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in (0, 1):
        assert res[r]["episodes"] == 30 and res[r]["steps"] == 1000 and res[r]["return_sum"] == -9
        assert res[r]["bound_crash"] == 4 and res[r]["building_crash"] == 4
        assert abs(res[r]["mean_return"] + 0.3) < 1e-12
    assert res[0]["span"] == (0, 501) and res[1]["span"] == (501, 1001)


def test_host_planner_matches_python_restatement_of_the_reference_search():
    """aac_plan_path (C++, host only) against reset.ref_line_cells (the Python restatement that reproduces the
    reference's episodes in tests/test_gpu_parity.py::test_ref_compat_env_reproduces_reference_episode)."""
    from multi_agent_aac_b200.reset import OdTable, ref_line_cells
    _capi.build()
    for m in (synthetic_map(seed=0), multimap_set(seed=0)[5]):
        tab = OdTable(m, w_max=32)
        P = tab.n_cells
        assert tab.pool_off[0] == 0 and tab.pool_off[4] == P and (np.diff(tab.pool_off) > 0).all()
        rng = np.random.default_rng(0)
        checked = 0
        while checked < 150:
            s, t = int(rng.integers(0, P)), int(rng.integers(0, P))
            if tab.path_len[s * P + t] == 0:
                qs = np.searchsorted(tab.pool_off, s, side="right") - 1
                qt = np.searchsorted(tab.pool_off, t, side="right") - 1
                assert qs == qt                      # only same-quadrant pairs are left out
                continue
            want = ref_line_cells(m, m.cell_centre(*divmod(int(tab.cell_code[s]), 256)), m.cell_centre(*divmod(int(tab.cell_code[t]), 256)))
            got = [divmod(int(c), 256) for c in tab.path(s, t)]
            assert got == [tuple(c) for c in want], (s, t)
            checked += 1


def test_bench_reference_arm_prints_the_contract_line():
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                          "--cpu-envs", "64"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-500:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "agent_steps_per_sec" and line["unit"] == "agent-steps/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"] and line["cpu_baseline"]["cores"] >= 1


def test_bench_c1_and_profile_gating():
    """BASELINE config 1 (one env, one core) prints the contract line on the CPU; the ncu-derived fields of the GPU line are
    tied to the kernel sources: a capture of another build or another instantiation yields None."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "c1", "--steps", "60"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-500:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["cpu_baseline"]["cores"] == 1 and line["steps"] == 60 and line["value"] > 0 and "1 env" in line["cpu_baseline"]["sample"]
    sys.path.insert(0, ROOT)
    import bench
    names = bench.launched_kernels("v2", 10, 36, 1, 2)
    assert names == ["env_kernel<1,0,1,10,36,0,3,1,0,0>", "env_kernel<1,0,1,10,36,0,2,1,0,0>"] and bench.launched_kernels("att", 3, 36, 0, 2) is None
    phased = bench.launched_kernels("v2", 10, 36, 1, 1)
    assert phased == ["env_kernel<1,0,1,10,36,0,4,1,0,0>"]
    for ks in (names, phased):
        rows = bench._profiled_rows(ks)
        if rows is not None:      # the committed summaries belong to this very source: every launch of the step, named exactly
            assert all(r["Source Hash"][1] == bench.source_hash() for r in rows) and bench.profiled_traffic(ks) > 3e8
    assert bench.profiled_traffic(["env_kernel<1,0,1,10,36,0,3,0,0,0>", names[1]]) is None      # another radar mode: not these kernels
    assert bench.profiled_traffic(["env_kernel<1,0,1,10,36,0,4,0,0,0>"]) is None
    assert bench.profiled_traffic(bench.launched_kernels("v2", 20, 72, 1, 2)) is None


def test_actor_library_exports_every_declared_symbol():
    from multi_agent_aac_b200 import _actor_capi
    _actor_capi.build()
    header = open(os.path.join(ROOT, "include", "aac_actor.h")).read()
    declared = set(re.findall(r"\b(aac_actor_[a-z_]+)\s*\(", header))
    assert declared == set(_actor_capi.EXPORTS), declared ^ set(_actor_capi.EXPORTS)
    lib = ctypes.CDLL(_actor_capi.LIB_PATH)
    for name in declared:
        assert getattr(lib, name) is not None
    lib.aac_actor_last_error.restype = ctypes.c_char_p
    h = ctypes.c_void_p()
    cfg = _actor_capi.AacActorConfig(999, 7, 45, 36, 10)
    lib.aac_actor_create.argtypes = [ctypes.POINTER(_actor_capi.AacActorConfig), ctypes.POINTER(ctypes.c_void_p)]
    assert lib.aac_actor_create(ctypes.byref(cfg), ctypes.byref(h)) == -1 and b"abi_version" in lib.aac_actor_last_error()
    assert ctypes.sizeof(_actor_capi.AacActorConfig) == 5 * 4 and ctypes.sizeof(_actor_capi.AacActorParams) == 12 * 8


def test_actor_oracle_matches_reference_fixture():
    """The float64 restatement against the outputs of the unmodified reference class (gen_golden_actor.py)."""
    import numpy as np
    from oracle import actor_oracle
    for name in ("actor_v2", "actor_v2_r18_n4"):
        d = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
        d_own, d_nbr, d_grid, seed = (int(v) for v in d["dims"])
        sd = actor_oracle.reference_like_params(d_own, d_nbr, d_grid, seed)
        act = actor_oracle.forward(sd, d["own"], d["nbr"], d["grid"])
        assert np.abs(act - d["act"]).max() < 1e-12
    noisy = actor_oracle.explore(np.array([[0.9, -0.2]]), np.array([[1.0, -1.0]]), 0.5)
    assert np.allclose(noisy, [[1.0, -0.7]])


def test_att_actor_oracle_matches_reference_fixture():
    import numpy as np
    from oracle import actor_oracle
    for name in ("actor_att", "actor_att_n5_r18"):
        d = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
        d_own, d_grid, n_nei, seed = (int(v) for v in d["dims"])
        act = actor_oracle.forward_att(actor_oracle.reference_like_params_att(d_own, d_grid, seed), d["own"], d["grid"], d["nei"])
        assert np.abs(act - d["act"]).max() < 1e-12


def test_gridmap_from_polygons_matches_geometry_restatement():
    """Map ingestion (ATT/grid_env_generation:140-171): the vectorised polygon-vs-cell test against the GEOS restatement's
    Polygon.intersects on random star-shaped (non-convex) footprints, including cells that are only touched; holes fill."""
    import math
    import numpy as np
    from scipy import ndimage
    from multi_agent_aac_b200.maps import gridmap_from_polygons, grid_shape
    from oracle import geos_lite as G
    rng = np.random.default_rng(5)
    bound = [455, 680, 255, 385]
    polys = []
    for _ in range(12):
        cx, cy = rng.uniform(470, 670), rng.uniform(265, 375)
        ang = np.sort(rng.uniform(0, 2 * np.pi, rng.integers(3, 9)))
        rad = rng.uniform(4, 28, len(ang))
        polys.append([(cx + r * math.cos(a), cy + r * math.sin(a)) for a, r in zip(ang, rad)])
    polys.append([(500, 300), (520, 300), (520, 305), (500, 305)])          # edges exactly on cell borders: touching cells count
    m = gridmap_from_polygons(polys, bound)
    gx, gy = grid_shape(bound, 10)
    brute = np.zeros((180, 130), dtype=bool)
    for ix in range(180):
        for iy in range(130):
            if not (440 <= ix * 10 <= 700 and 240 <= iy * 10 <= 400):
                continue
            sq = G.Point(ix * 10, iy * 10).buffer(5, cap_style=3)
            brute[ix, iy] = any(G.Polygon(p).intersects(sq) for p in polys)
    brute = ndimage.binary_fill_holes(brute)
    want = brute[46:46 + gx, 26:26 + gy]                                      # grid points 460..680 x 260..380
    assert m.occ.shape == (gx, gy) and np.array_equal(m.occ.astype(bool), want)
    assert m.occ[(500 - 460) // 10, (310 - 260) // 10] == 1                  # the cell above the strip touches its top edge
    ring = [[(560, 330), (620, 330), (620, 340), (560, 340)], [(560, 370), (620, 370), (620, 380), (560, 380)],
            [(560, 330), (570, 330), (570, 380), (560, 380)], [(610, 330), (620, 330), (620, 380), (610, 380)]]
    assert gridmap_from_polygons(ring, bound).occ[(590 - 460) // 10, (355 - 260) // 10] == 1   # enclosed courtyard is filled


def test_shapefile_ingestion_matches_reference_env_generation():
    """Map ingestion end to end: shapefile -> occupancy grid without geopandas, against the occupied /
    free cell lists the UNMODIFIED `env_generation` (ATT/grid_env_generation:108-185) produced for the same shapefile
    (tests/golden/mapgen_ref.*, made by tests/golden/gen_golden_mapgen.py): duplicate footprints dropped, SVY21 -> metres,
    zero-height footprints do not occupy layer 0, enclosed courtyards are filled, bounded to the closed bound."""
    import numpy as np
    from multi_agent_aac_b200.maps import gridmap_from_shapefile, read_shapefile
    base = os.path.join(ROOT, "tests", "golden", "mapgen_ref")
    d = np.load(base + ".npz")
    bound = [int(v) for v in d["bound"]]
    rings, rows, names = read_shapefile(base + ".shp")
    assert names == ["ID", "NAME", "HEIGHT", "A", "B", "C"] and len(rings) == len(rows) == 27
    assert rows[1][1] == "B01" and isinstance(rows[1][2], float) and np.allclose(rings[0][0], rings[0][-1])
    m = gridmap_from_shapefile(base + ".shp", bound)
    ones = {(int(x), int(y)) for x, y in d["ones"]}
    zeros = {(int(x), int(y)) for x, y in d["zeros"]}
    assert len(ones) + len(zeros) == m.gx * m.gy and len(ones) > 50
    for ix in range(m.gx):
        for iy in range(m.gy):
            c = tuple(int(v) for v in m.cell_centre(ix, iy))
            assert (c in ones) == bool(m.occ[ix, iy]) and (c in zeros) != bool(m.occ[ix, iy]), c
    assert any(h == 0.0 for h in (r[2] for r in rows))              # the data does contain zero-height footprints
    # where the reference is present (build container), the golden is re-derived from it
    if os.path.isdir("/root/reference"):
        sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
        import gen_golden_mapgen as gen
        emb, ones2, zeros2, g, extent = gen.run_reference(base + ".shp")
        assert {tuple(c) for c in ones2} == ones and {tuple(c) for c in zeros2} == zeros and g == 10 and tuple(extent) == (1800, 1300)


def test_build_is_gated_on_source_content_not_mtime():
    """A library next to a digest of other sources is rebuilt by build() and refused by lib() (no silent stale binary)."""
    from multi_agent_aac_b200 import _capi
    _capi.build()
    side = _capi.LIB_PATH + ".srchash"
    good = open(side).read()
    assert good.strip() == _capi.sources_digest()
    try:
        open(side, "w").write("0" * 64 + "\n")
        _capi._lib = None
        with pytest.raises(_capi.AacError, match="other sources"):
            _capi.lib()
    finally:
        open(side, "w").write(good)
        _capi._lib = None
    assert _capi.lib() is not None


def test_pools_only_od_table_has_the_same_pools_and_no_paths():
    """OdTable(paths=False): the quadrant pools of the full table, no P^2 paths (the device searches per episode)."""
    from multi_agent_aac_b200.maps import synthetic_map
    from multi_agent_aac_b200.reset import OdTable
    gmap = synthetic_map(seed=0)
    full, pools = OdTable(gmap, w_max=32), OdTable(gmap, w_max=32, paths=False)
    assert pools.n_cells == full.n_cells and np.array_equal(pools.pool_off, full.pool_off) and np.array_equal(pools.cell_code, full.cell_code)
    assert pools.path_cells is None and pools.path_off is None and pools.path_len is None and not pools.has_paths and full.has_paths
    assert full.path_cells.size % 8 == 0 and int(full.path_len.max()) <= 32
