"""ctypes binding of the CPU oracle (oracle/libeslam_oracle.so).  TEST INFRASTRUCTURE ONLY: imported by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs -- never by the product."""
import ctypes as C
import os
import subprocess

import numpy as np

from slam_eslam_b200 import abi
from slam_eslam_b200.filter import FilterBase

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libeslam_oracle.so")
LITERAL, SHARED = 0, 1

_lib = None

_EXTRA = {
    "create": (C.c_int, [C.POINTER(abi.Config), C.c_int, C.POINTER(abi.Shard), C.POINTER(C.c_void_p)]),
    "destroy": (None, [C.c_void_p]),
    "set_fake_map": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_uint8)]),
    "get_queried_points": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.c_int]),
    "lowest_points": (C.c_int, [C.POINTER(abi.Contact), C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_float)]),
    "math": (C.c_double, [C.c_int, C.c_double, C.c_double]),
}


def build(force=False):
    """Compile the oracle with oracle/Makefile (g++)."""
    if force and os.path.exists(LIB_PATH):
        os.remove(LIB_PATH)
    subprocess.check_call(["make", "-s", "-C", HERE])


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        lib = C.CDLL(LIB_PATH)
        abi.bind(lib, "eslam_oracle_", abi.SIGNATURES)
        abi.bind(lib, "eslam_oracle_", _EXTRA)
        _lib = lib
    return _lib


class OraclePoseEstimator(FilterBase):
    """The reference's PoseEstimator restated on the CPU; policy LITERAL (line-by-line) or SHARED (GPU arithmetic)."""

    def __init__(self, config=None, policy=SHARED, shard=None):
        lib = load()
        cfg = config if config is not None else abi.default_config()
        ctx = C.c_void_p()
        st = lib.eslam_oracle_create(C.byref(cfg), int(policy), None if shard is None else C.byref(shard), C.byref(ctx))
        assert st == abi.OK
        super().__init__(lib, "eslam_oracle_", ctx)
        self.config, self.policy = cfg, policy

    def close(self):
        if self._ctx:
            self._lib.eslam_oracle_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def setFakeMap(self, z, stddev, res=None):
        """FakeMLSAccess of test/testContactModel.cpp:8-38."""
        zz = (C.c_double * 4)(*z)
        ss = (C.c_double * 4)(*stddev)
        rr = None if res is None else (C.c_uint8 * 4)(*[1 if r else 0 for r in res])
        self._check(self._lib.eslam_oracle_set_fake_map(self._ctx, zz, ss, rr))

    def queriedPoints(self, max_points=1024):
        buf = np.zeros((max_points, 3))
        n = self._lib.eslam_oracle_get_queried_points(self._ctx, buf.ctypes.data_as(C.POINTER(C.c_double)), max_points)
        return buf[:min(n, max_points)]


def lowest_points(contacts, update_probabilities=False):
    """ContactModel::getLowestPointPerGroup / updateContactStateUsingLowestPointHeuristic."""
    lib = load()
    n = len(contacts)
    xyz = np.zeros((n, 3))
    probs = np.zeros(n, dtype=np.float32)
    cnt = lib.eslam_oracle_lowest_points(contacts, n, 1 if update_probabilities else 0,
                                         xyz.ctypes.data_as(C.POINTER(C.c_double)),
                                         probs.ctypes.data_as(C.POINTER(C.c_float)))
    return xyz[:cnt], probs


def math_fn(fn, a, b=0.0):
    return load().eslam_oracle_math(int(fn), float(a), float(b))
