"""ctypes binding of the C ABI declared in include/aac_actor.h (libaac_actor.so, built in-tree).

There is no fallback: if the shared library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AAC_ACTOR_LIB") or os.path.join(_HERE, "libaac_actor.so")
SOURCES = [os.path.join(_HERE, "csrc", f) for f in ("aac_actor.cu", "aac_actor_att.cu")]
HEADERS = [os.path.join(os.path.dirname(_HERE), "include", "aac_actor.h")]

ABI_VERSION = 1
H1, H2, H3, NACT = 128, 512, 256, 2
EXPORTS = ["aac_actor_create", "aac_actor_destroy", "aac_actor_load", "aac_actor_forward", "aac_actor_hidden", "aac_actor_prof", "aac_actor_launch_count",
           "aac_actor_last_error", "aac_actor_att_create", "aac_actor_att_destroy", "aac_actor_att_load", "aac_actor_att_forward",
           "aac_actor_att_launch_count", "aac_actor_att_last_error"]
ATT_PARAM_FIELDS = ["w_own", "b_own", "w_grid", "b_grid", "w_nei", "b_nei", "w_q", "w_k", "w_v", "w_merge", "b_merge", "w_out", "b_out"]
PARAM_FIELDS = ["w_own", "b_own", "w_nbr", "b_nbr", "w_grid", "b_grid", "w_merge", "b_merge", "w_hid", "b_hid", "w_out", "b_out"]


class AacActorConfig(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("d_own", C.c_int32), ("d_nbr", C.c_int32), ("d_grid", C.c_int32), ("max_rows", C.c_int32)]


class AacActorParams(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in PARAM_FIELDS]


class AacActorAttConfig(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("d_own", C.c_int32), ("d_grid", C.c_int32), ("d_nei", C.c_int32), ("n_nei", C.c_int32)]


class AacActorAttParams(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ATT_PARAM_FIELDS]


class AacActorError(RuntimeError):
    pass


def nvcc_command(out=LIB_PATH):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not os.path.exists(nvcc):
        nvcc = "nvcc"
    return [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-shared", "-o", out] + SOURCES


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
    """Compile libaac_actor.so for sm_100a (cross-compiles without a GPU).  Gated on the CONTENT of the sources (a digest
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
        raise AacActorError("libaac_actor.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(there is no CPU fallback)")
    if not os.environ.get("AAC_ACTOR_LIB") and _built_digest() not in (None, sources_digest()):
        raise AacActorError(os.path.basename(LIB_PATH) + " was built from other sources than the ones in this tree: run "
                       "`python -c 'import __graft_entry__ as g; g.build()'`")
    L = C.CDLL(LIB_PATH)
    P = C.c_void_p
    L.aac_actor_create.argtypes = [C.POINTER(AacActorConfig), C.POINTER(P)]
    L.aac_actor_destroy.argtypes = [P]
    L.aac_actor_destroy.restype = None
    L.aac_actor_load.argtypes = [P, C.POINTER(AacActorParams)]
    L.aac_actor_forward.argtypes = [P, P, P, P, C.c_int32, C.c_float, C.c_uint64, P, P]
    L.aac_actor_hidden.argtypes = [P, P, P, P, C.c_int32, C.c_int32, P, P]
    L.aac_actor_prof.argtypes = [P, P]
    L.aac_actor_launch_count.argtypes = [P]
    L.aac_actor_launch_count.restype = C.c_int64
    L.aac_actor_last_error.restype = C.c_char_p
    L.aac_actor_att_create.argtypes = [C.POINTER(AacActorAttConfig), C.POINTER(P)]
    L.aac_actor_att_destroy.argtypes = [P]
    L.aac_actor_att_destroy.restype = None
    L.aac_actor_att_load.argtypes = [P, C.POINTER(AacActorAttParams)]
    L.aac_actor_att_forward.argtypes = [P, P, P, P, C.c_int32, C.c_float, C.c_uint64, P, P]
    L.aac_actor_att_launch_count.argtypes = [P]
    L.aac_actor_att_launch_count.restype = C.c_int64
    L.aac_actor_att_last_error.restype = C.c_char_p
    for name in ("aac_actor_att_create", "aac_actor_att_load", "aac_actor_att_forward"):
        getattr(L, name).restype = C.c_int
    for name in ("aac_actor_create", "aac_actor_load", "aac_actor_forward", "aac_actor_hidden"):
        getattr(L, name).restype = C.c_int
    _lib = L
    return L


def check(rc, what):
    if rc != 0:
        raise AacActorError("%s failed (%d): %s" % (what, rc, lib().aac_actor_last_error().decode()))
