"""ctypes binding of the C ABI declared in include/aac_env.h (libaac_env.so, built in-tree).

There is no fallback: if the shared library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AAC_LIB") or os.path.join(_HERE, "libaac_env.so")   # AAC_LIB: A/B builds while tuning kernels
SOURCES = [os.path.join(_HERE, "csrc", f) for f in ("aac_kernels.cu", "aac_capi.cu")]
HEADERS = [os.path.join(_HERE, "csrc", "aac_kernels.cuh"), os.path.join(_HERE, "csrc", "aac_radar.cuh"), os.path.join(_HERE, "csrc", "aac_plan.cuh"),
           os.path.join(os.path.dirname(_HERE), "include", "aac_env.h")]

ABI_VERSION = 3
VARIANT_ATT, VARIANT_V2, VARIANT_MM = 0, 1, 2
RADAR_MIN, RADAR_LAST_HIT = 0, 1
OUT_RAW, OUT_NBR6, OUT_TCPA_PAIR, OUT_RADAR_AUX, OUT_PARTS = 0x01, 0x02, 0x04, 0x08, 0x10
TARGET_CELLS, TARGET_BOUNDS, TARGET_CLOUDS, TARGET_AIRCRAFT = 0x1, 0x2, 0x4, 0x8   # the later fork's radar target classes
MAX_CLOUDS = 8
MAP_STRIDE = 1024
N_STATS = 16
STAT_NAMES = ["episodes", "steps", "return_sum", "bound_crash", "building_crash", "drone_crash", "drone_crash_nearest",
              "all_reached", "drones_reached", "step_cap", "plan_fallback"]

EXPORTS = ["aac_create", "aac_destroy", "aac_set_maps", "aac_set_radar_table", "aac_set_bank", "aac_set_od_tables", "aac_plan_path", "aac_plan_paths_device", "aac_bind_state", "aac_reset", "aac_observe",
           "aac_step", "aac_step_autoreset", "aac_step_fused", "aac_autoreset", "aac_step_host", "aac_read_stats", "aac_launch_count", "aac_own_dim",
           "aac_last_error"]


class AacConfig(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("variant", C.c_int32), ("n_envs", C.c_int32), ("n_agents", C.c_int32),
                ("n_rays", C.c_int32), ("w_max", C.c_int32), ("radar_mode", C.c_int32), ("sum_reward", C.c_int32),
                ("episode_length", C.c_int32), ("out_flags", C.c_int32), ("tile_envs", C.c_int32),
                ("block_threads", C.c_int32), ("env_id_base", C.c_int64), ("seed", C.c_uint64),
                ("dt", C.c_float), ("vmax", C.c_float), ("acc_max", C.c_float), ("prot", C.c_float),
                ("ray_len", C.c_float), ("goal_r", C.c_float), ("eval_by_step", C.c_int32), ("autoreset_launches", C.c_int32),
                ("radar_targets", C.c_int32), ("n_nbr_obs", C.c_int32), ("n_clouds", C.c_int32), ("clouds", (C.c_float * 6) * MAX_CLOUDS)]


class AacMapDesc(C.Structure):
    _fields_ = [("gx", C.c_int32), ("gy", C.c_int32), ("bound", C.c_float * 4), ("x0c", C.c_float), ("y0c", C.c_float),
                ("cell", C.c_float), ("origin_x", C.c_float), ("origin_y", C.c_float)]


STATE_FIELDS = ["px", "py", "vx", "vy", "heading", "meta", "ref_cells", "ref_w", "wall_count", "ep_step", "ep_index",
                "ep_return", "map_id", "wp_mask"]
OUT_FIELDS = ["norm_own", "norm_nbr", "radar", "norm_nbr6", "raw_own", "raw_nbr", "raw_nbr6", "reward", "done",
              "check_goal", "bbc", "terminated", "tcpa_min", "tcpa_pair", "nbr_order", "radar_min", "radar_hit", "parts",
              "branch", "cloud_contact"]


class AacState(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in STATE_FIELDS]


class AacOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in OUT_FIELDS]


class AacBank(C.Structure):
    _fields_ = [("n_scenarios", C.c_int32), ("cells", C.c_void_p), ("w", C.c_void_p), ("map_id", C.c_void_p)]


class AacOdTable(C.Structure):
    _fields_ = [("n_cells", C.c_int32), ("pool_off", C.c_int32 * 5), ("cell_code", C.c_void_p), ("path_off", C.c_void_p),
                ("path_len", C.c_void_p), ("path_cells", C.c_void_p), ("n_path_cells", C.c_int64)]


class AacError(RuntimeError):
    pass


def nvcc_command(out=LIB_PATH):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not os.path.exists(nvcc):
        nvcc = "nvcc"
    return [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-prec-div=false", "-prec-sqrt=false",
            "-Xcompiler", "-fPIC", "-shared", "-o", out] + SOURCES


def sources_digest():
    """sha256 over the sources, the headers and the compile command: what a built library is good for."""
    import hashlib
    flags = [a for a in nvcc_command("-")[1:] if not os.path.isabs(a)]   # the tree may live anywhere: no paths in the digest
    h = hashlib.sha256(" ".join(flags).encode())
    for p in SOURCES + HEADERS:
        h.update(os.path.basename(p).encode())
        h.update(open(p, "rb").read())
    return h.hexdigest()


def _built_digest():
    try:
        return open(LIB_PATH + ".srchash").read().strip()
    except OSError:
        return None


def build(force=False):
    """Compile libaac_env.so for sm_100a (cross-compiles without a GPU).  Gated on the CONTENT of the sources (a digest
    written next to the library), not on modification times: a library left over from other sources is rebuilt."""
    want = sources_digest()
    if not force and os.path.exists(LIB_PATH) and _built_digest() == want:
        return LIB_PATH
    subprocess.check_call(nvcc_command())
    with open(LIB_PATH + ".srchash", "w") as f:
        f.write(want + "\n")
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AacError("libaac_env.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(there is no CPU fallback)")
    if not os.environ.get("AAC_LIB") and _built_digest() not in (None, sources_digest()):
        raise AacError(os.path.basename(LIB_PATH) + " was built from other sources than the ones in this tree: run "
                       "`python -c 'import __graft_entry__ as g; g.build()'`")
    L = C.CDLL(LIB_PATH)
    P = C.c_void_p
    L.aac_create.argtypes = [C.POINTER(AacConfig), C.POINTER(P)]
    L.aac_destroy.argtypes = [P]
    L.aac_destroy.restype = None
    L.aac_set_maps.argtypes = [P, C.POINTER(AacMapDesc), P, C.c_int32]
    L.aac_set_radar_table.argtypes = [P, P, P, P, P]
    L.aac_set_bank.argtypes = [P, C.POINTER(AacBank)]
    L.aac_set_od_tables.argtypes = [P, C.POINTER(AacOdTable), C.c_int32]
    L.aac_plan_path.argtypes = [P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, P, C.c_int32]
    L.aac_plan_paths_device.argtypes = [P, C.c_int32, C.c_int32, P, C.c_int64, P, P, C.c_int32, P]
    L.aac_bind_state.argtypes = [P, C.POINTER(AacState)]
    L.aac_reset.argtypes = [P, P, C.POINTER(AacOut), P]
    L.aac_observe.argtypes = [P, C.POINTER(AacOut), P]
    L.aac_step.argtypes = [P, P, C.POINTER(AacOut), P]
    L.aac_step_autoreset.argtypes = [P, P, C.POINTER(AacOut), P]
    L.aac_step_fused.argtypes = [P, P, C.POINTER(AacOut), P]
    L.aac_autoreset.argtypes = [P, C.POINTER(AacOut), P]
    L.aac_step_host.argtypes = [P, P, C.POINTER(AacOut), C.POINTER(AacOut), C.c_int32, P]
    L.aac_read_stats.argtypes = [P, P, C.c_int32, P]
    L.aac_launch_count.argtypes = [P]
    L.aac_launch_count.restype = C.c_int64
    L.aac_own_dim.argtypes = [C.c_int32, C.c_int32]
    L.aac_last_error.restype = C.c_char_p
    for name in EXPORTS:
        if name not in ("aac_destroy", "aac_launch_count", "aac_last_error"):
            getattr(L, name).restype = C.c_int
    _lib = L
    return L


def check(rc, what):
    if rc != 0:
        raise AacError("%s failed (%d): %s" % (what, rc, lib().aac_last_error().decode()))
