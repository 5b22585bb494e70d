"""Golden paths from the UNMODIFIED reference search `jps_find_path` (MADDPG_ownENV_randomOD_radar_one_model_att/
jps_straight.py:17-70 - a self-contained pure-Python module, imported from /root/reference as it lies), build container only.

For two of this repo's occupancy grids (the synthetic single map and one of the multipleMap set) and a small grid with a closed
room (unreachable goals), 400 seeded pairs of free cells each - any two free cells, not only the quadrant pools - the reference's full cell-by-cell path, stored flat.
tests/test_oracle_golden.py::test_planner_equals_the_reference_jps expands the planner's pruned line back to cells and
compares it with these, cell for cell; tests/test_gpu_planner.py does the same with the device planner.

    python tests/golden/gen_golden_jps.py        # writes tests/golden/jps_paths.npz
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
REF = "/root/reference/MADDPG_ownENV_randomOD_radar_one_model_att/jps_straight.py"


def main():
    from multi_agent_aac_b200.maps import multimap_set, synthetic_map
    spec = importlib.util.spec_from_file_location("ref_jps_straight", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    out = {}
    walled = np.zeros((12, 10), dtype=np.uint8)       # a closed room: goals inside are unreachable from outside (the search returns None)
    walled[3:9, 2] = walled[3:9, 7] = 1
    walled[3, 2:8] = walled[8, 2:8] = 1
    walled[5, 4] = 1
    for name, src in (("single", synthetic_map(seed=0)), ("multi5", multimap_set(seed=0)[5]), ("walled", walled)):
        occ = np.ascontiguousarray(src if isinstance(src, np.ndarray) else src.occ, dtype=np.uint8)   # [gx][gy], 1 = building: the grid reset_world passes (ATT:317)
        grid = occ.tolist()
        free = np.argwhere(occ == 0)
        rng = np.random.default_rng(7)
        pairs, flat, off = [], [], [0]
        while len(pairs) < 400:
            s, t = free[rng.integers(len(free))], free[rng.integers(len(free))]
            if (s == t).all():
                continue
            path = ref.jps_find_path((int(s[0]), int(s[1])), (int(t[0]), int(t[1])), grid)
            pairs.append([s[0], s[1], t[0], t[1]])
            cells = [] if path is None else [c[0] * 256 + c[1] for c in path]
            flat.extend(cells)
            off.append(len(flat))
        out[name + "_occ"] = occ
        out[name + "_pairs"] = np.array(pairs, dtype=np.int32)
        out[name + "_cells"] = np.array(flat, dtype=np.uint16)
        out[name + "_off"] = np.array(off, dtype=np.int64)
        print(name, occ.shape, "pairs", len(pairs), "unreachable", sum(1 for k in range(len(pairs)) if off[k + 1] == off[k]), "longest", max(np.diff(off)))
    np.savez_compressed(os.path.join(HERE, "jps_paths.npz"), **out)


if __name__ == "__main__":
    main()
