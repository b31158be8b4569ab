"""ctypes binding of oracle/_ref/libeslam_ref.so: the reference's OWN particle-filter sources (src/ContactModel.cpp,
src/PoseEstimator.cpp and their headers), compiled unmodified against the stand-in dependency headers of oracle/ref_shim
(oracle/ref/Makefile, oracle/ref/ref_driver.cpp).  TEST INFRASTRUCTURE ONLY: imported by tests/ and bench.py's CPU legs.

/root/reference exists only in the build container: there the library is (re)built; on the GPU box the prebuilt file
that travelled with the snapshot is loaded; with neither, :func:`available` is False and callers skip."""
import ctypes as C
import os
import subprocess

import numpy as np

from slam_eslam_b200 import abi
from slam_eslam_b200.filter import FilterBase

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "libeslam_ref.so")
TEST_BIN = os.path.join(HERE, "_ref", "testContactModel")
REFERENCE = os.environ.get("ESLAM_REFERENCE_DIR", "/root/reference")

_lib = None

# the subset of abi.SIGNATURES the reference can answer (no running-sum tap, no patches: those are not in its code)
_NAMES = ["set_grid", "init_gaussian", "set_particles", "get_particles", "particle_count", "project", "update",
          "normalize_weights", "resample", "get_ancestors", "centroid", "best_index", "evaluate_poses"]
_EXTRA = {
    "create": (C.c_int, [C.POINTER(abi.Config), C.c_int, C.POINTER(C.c_void_p)]),
    "destroy": (None, [C.c_void_p]),
    "set_fake_map": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_uint8)]),
    "clear_fake_map": (None, []),
    "get_queried_points": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.c_int]),
    "lowest_points": (C.c_int, [C.POINTER(abi.Contact), C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_float)]),
}


def have_sources():
    return os.path.exists(os.path.join(REFERENCE, "src", "PoseEstimator.cpp"))


def build(force=False):
    """Compile the reference's sources where they lie (only possible where /root/reference exists)."""
    if not have_sources():
        return False
    if force and os.path.exists(LIB_PATH):
        os.remove(LIB_PATH)
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "ref"), "REF=" + REFERENCE,
                           "../_ref/libeslam_ref.so", "../_ref/testContactModel"])
    return True


def available():
    return os.path.exists(LIB_PATH) or have_sources()


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH) and not build():
            raise RuntimeError("oracle/_ref/libeslam_ref.so is not built and %s is not here to build it from" % REFERENCE)
        lib = C.CDLL(LIB_PATH)
        abi.bind(lib, "eslam_ref_", {k: abi.SIGNATURES[k] for k in _NAMES})
        abi.bind(lib, "eslam_ref_", _EXTRA)
        _lib = lib
    return _lib


class RefPoseEstimator(FilterBase):
    """eslam::PoseEstimator itself (the reference's compiled code) behind the PoseEstimator mirror API of FilterBase.
    debug=True keeps the per-particle debug vectors (n_found, ancestor tags) at the cost of speed."""

    def __init__(self, config=None, debug=True):
        lib = load()
        cfg = config if config is not None else abi.default_config()
        ctx = C.c_void_p()
        st = lib.eslam_ref_create(C.byref(cfg), 1 if debug else 0, C.byref(ctx))
        if st != abi.OK:
            raise RuntimeError("the reference has no such mode (systematic resampling, patches and the hash are not in its update path)")
        super().__init__(lib, "eslam_ref_", ctx)
        self.config = cfg

    def close(self):
        if self._ctx:
            self._lib.eslam_ref_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def setFakeMap(self, z, stddev, res=None):
        """FakeMLSAccess of test/testContactModel.cpp:8-38 behind ContactModel::evaluatePose's map callback."""
        zz = (C.c_double * 4)(*z)
        ss = (C.c_double * 4)(*stddev)
        rr = None if res is None else (C.c_uint8 * 4)(*[1 if r else 0 for r in res])
        self._check(self._lib.eslam_ref_set_fake_map(self._ctx, zz, ss, rr))

    def clearFakeMap(self):
        self._lib.eslam_ref_clear_fake_map()

    def queriedPoints(self, max_points=1024):
        buf = np.zeros((max_points, 3))
        n = self._lib.eslam_ref_get_queried_points(self._ctx, buf.ctypes.data_as(C.POINTER(C.c_double)), max_points)
        return buf[:min(n, max_points)]


def lowest_points(contacts, update_probabilities=False):
    lib = load()
    n = len(contacts)
    xyz = np.zeros((n, 3))
    probs = np.zeros(n, dtype=np.float32)
    cnt = lib.eslam_ref_lowest_points(contacts, n, 1 if update_probabilities else 0,
                                      xyz.ctypes.data_as(C.POINTER(C.c_double)), probs.ctypes.data_as(C.POINTER(C.c_float)))
    return xyz[:cnt], probs
