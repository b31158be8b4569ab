"""ctypes view of include/eslam_gpu.h (the C ABI of libeslam_gpu.so).

PyTorch/NumPy are plumbing for tests and the benchmark; the product is the shared library.  There is no
CPU fallback: :func:`load_library` raises if the CUDA library has not been built.
"""
import ctypes as C
import math
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(ROOT, "slam-eslam_b200", "libeslam_gpu.so")

MAX_CONTACTS = 128
MAX_CP = 8
OK, ERR_INVALID, ERR_CUDA, ERR_NO_ENVIRONMENT, ERR_ZERO_VARIANCE, ERR_EMPTY, ERR_COMM = range(7)
RESAMPLE_STRATIFIED, RESAMPLE_SYSTEMATIC = 0, 1
SHARD_EXACT_GLOBAL, SHARD_LOCAL_MIGRATE = 0, 1

ERROR_TEXT = {
    ERR_INVALID: "invalid argument",
    ERR_CUDA: "CUDA failure",
    ERR_NO_ENVIRONMENT: "No environment attached.",  # src/PoseEstimator.cpp:260
    ERR_ZERO_VARIANCE: "using a zero measurement variance leads to singularities",  # src/ContactModel.cpp:123
    ERR_EMPTY: "empty particle set",
    ERR_COMM: "communicator / peer-memory exchange failed",
}


class Config(C.Structure):
    _fields_ = [
        ("seed", C.c_uint64), ("min_effective", C.c_uint64),
        ("measurement_error", C.c_double), ("discount_factor", C.c_double), ("spread_threshold", C.c_double),
        ("spread_translation_factor", C.c_double), ("spread_rotation_factor", C.c_double),
        ("slip_factor", C.c_double), ("max_yaw_deviation", C.c_double),
        ("min_contacts", C.c_uint64), ("contact_likelihood_correction", C.c_double),
        ("contact_point_radius", C.c_double),
        ("use_slip_update", C.c_int32), ("use_shape_update", C.c_int32), ("use_hash", C.c_int32),
        ("resample_mode", C.c_int32), ("patch_size", C.c_int32), ("shard_mode", C.c_int32),
        ("migrate_min_share", C.c_double), ("migrate_min_ess", C.c_double),
        ("virtual_shards", C.c_int32), ("reserved", C.c_int32),
    ]


def default_config(**overrides):
    """eslam::Configuration / ContactModelConfiguration defaults (src/Configuration.hpp:53-59,85-111)."""
    c = Config(seed=42, min_effective=50, measurement_error=0.1, discount_factor=0.9, spread_threshold=0.9,
               spread_translation_factor=0.1, spread_rotation_factor=0.05, slip_factor=0.05,
               max_yaw_deviation=15 * math.pi / 180.0, min_contacts=3, contact_likelihood_correction=0.33,
               contact_point_radius=0.01, use_slip_update=0, use_shape_update=1, use_hash=0,
               resample_mode=RESAMPLE_STRATIFIED, patch_size=0, shard_mode=SHARD_EXACT_GLOBAL,
               migrate_min_share=0.25, migrate_min_ess=0.05, virtual_shards=0, reserved=0)
    for k, v in overrides.items():
        if not hasattr(c, k):
            raise AttributeError(k)
        setattr(c, k, v)
    return c


class Shard(C.Structure):
    _fields_ = [("rank", C.c_int32), ("world", C.c_int32), ("global_count", C.c_uint64)]


class Step(C.Structure):
    _fields_ = [("yaw", C.c_double), ("z_delta", C.c_double), ("z_var", C.c_double),
                ("delta_mean", C.c_double * 3), ("delta_sigma", C.c_double * 3)]


class Contact(C.Structure):
    _fields_ = [("position", C.c_double * 3), ("contact", C.c_float), ("slip", C.c_float),
                ("group_id", C.c_int32), ("pad", C.c_int32)]


class Scan(C.Structure):
    _fields_ = [("cells", C.POINTER(C.c_float)), ("width", C.c_int32), ("height", C.c_int32),
                ("scale_x", C.c_double), ("scale_y", C.c_double), ("offset_x", C.c_double), ("offset_y", C.c_double)]


class Gaussian2d(C.Structure):
    _fields_ = [("weight", C.c_double), ("mean", C.c_double * 2), ("cov", C.c_double * 4)]


class UpdateResult(C.Structure):
    _fields_ = [("ess", C.c_double), ("max_weight", C.c_double), ("floating_weight", C.c_double),
                ("weights_sum", C.c_double), ("data_particles", C.c_uint64), ("total_points", C.c_uint64),
                ("resampled", C.c_int32), ("status", C.c_int32)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


def make_contacts(positions, contact, group_id, slip=None):
    """Contact array from (C,3) positions, (C,) contact probabilities and (C,) group ids
    (odometry::BodyContactPoint fields as used at test/testContactModel.cpp:72-77)."""
    n = len(positions)
    arr = (Contact * n)()
    for i in range(n):
        for k in range(3):
            arr[i].position[k] = float(positions[i][k])
        arr[i].contact = float(contact[i])
        arr[i].slip = 0.0 if slip is None else float(slip[i])
        arr[i].group_id = int(group_id[i])
    return arr


def make_step(yaw=0.0, z_delta=0.0, z_var=0.0, delta_mean=(0, 0, 0), delta_sigma=(0, 0, 0)):
    s = Step(yaw=yaw, z_delta=z_delta, z_var=z_var)
    for i in range(3):
        s.delta_mean[i] = float(delta_mean[i])
        s.delta_sigma[i] = float(delta_sigma[i])
    return s


_dp = C.POINTER(C.c_double)
_u8p = C.POINTER(C.c_uint8)
_i32p = C.POINTER(C.c_int32)
_u32p = C.POINTER(C.c_uint32)
_fp = C.POINTER(C.c_float)

# name -> (restype, argtypes); ctx is a void*.  The oracle exports the same signatures under eslam_oracle_*.
SIGNATURES = {
    "set_grid": (C.c_int, [C.c_void_p, _fp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, _dp]),
    "init_gaussian": (C.c_int, [C.c_void_p, C.c_uint64, _dp, _dp, C.c_double, C.c_double]),
    "set_particles": (C.c_int, [C.c_void_p, C.c_uint64] + [_dp] * 6),
    "get_particles": (C.c_int, [C.c_void_p] + [_dp] * 7 + [_u8p, _u8p]),
    "particle_count": (C.c_uint64, [C.c_void_p]),
    "project": (C.c_int, [C.c_void_p, C.POINTER(Step)]),
    "update": (C.c_int, [C.c_void_p, C.POINTER(Contact), C.c_int, C.POINTER(UpdateResult)]),
    "normalize_weights": (C.c_int, [C.c_void_p, _dp]),
    "resample": (C.c_int, [C.c_void_p]),
    "resample_n": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int]),
    "get_ancestors": (C.c_int, [C.c_void_p, _u32p]),
    "get_cdf": (C.c_int, [C.c_void_p, _dp]),
    "centroid": (C.c_int, [C.c_void_p, _dp]),
    "best_index": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "evaluate_poses": (C.c_int, [C.c_void_p, C.c_uint64] + [_dp] * 6 + [C.POINTER(Contact), C.c_int, _u8p]
                       + [_dp] * 4 + [_i32p, _dp, _dp, _i32p, _dp]),
    "get_patches": (C.c_int, [C.c_void_p, _fp, _i32p]),
    "process_map": (C.c_int, [C.c_void_p, C.POINTER(Scan), C.c_int, C.c_int]),
    "pose_moments": (C.c_int, [C.c_void_p, C.POINTER(Gaussian2d)]),
    "hash_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "hash_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), _u32p]),
    "hash_get": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, _dp, _dp, _dp, _dp, _i32p]),
    "init_from_hash": (C.c_int, [C.c_void_p, C.c_uint64]),
    "sample_from_hash": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.POINTER(Contact), C.c_int, C.POINTER(C.c_uint64)]),
}

GPU_ONLY = {
    "default_config": (None, [C.POINTER(Config)]),
    "create": (C.c_int, [C.POINTER(Config), C.c_int, C.POINTER(Shard), C.POINTER(C.c_void_p)]),
    "nccl_unique_id": (C.c_int, [C.POINTER(C.c_uint8)]),
    "comm_init": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint8)]),
    "destroy": (None, [C.c_void_p]),
    "last_error": (C.c_char_p, [C.c_void_p]),
    "step_fused": (C.c_int, [C.c_void_p, C.POINTER(Step), C.POINTER(Contact), C.c_int, C.POINTER(UpdateResult)]),
    "run_steps": (C.c_int, [C.c_void_p, C.POINTER(Step), C.POINTER(Contact), C.c_int, C.c_int,
                            C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(UpdateResult)]),
    "step_times": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.c_int]),
    "io_bytes": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_uint64)]),
    "synchronize": (C.c_int, [C.c_void_p]),
    "weights_sum": (C.c_int, [C.c_void_p, _dp]),
    "debug_counters": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "path_counters": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "launch_count": (C.c_uint64, [C.c_void_p]),
}

EXPORTED_SYMBOLS = ["eslam_gpu_" + n for n in list(SIGNATURES) + list(GPU_ONLY)]


def bind(lib, prefix, names):
    for name, (res, args) in names.items():
        fn = getattr(lib, prefix + name)
        fn.restype = res
        fn.argtypes = args
    return lib


_lib = None


def load_library(path=None):
    """Load libeslam_gpu.so.  Raises (no fallback) when it has not been built: run __graft_entry__.build()."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError("libeslam_gpu.so not found at %s: build it first "
                           "(python -c 'import __graft_entry__ as g; g.build()'); there is no CPU fallback" % p)
    lib = C.CDLL(p, mode=C.RTLD_GLOBAL)
    bind(lib, "eslam_gpu_", SIGNATURES)
    bind(lib, "eslam_gpu_", GPU_ONLY)
    if path is None:
        _lib = lib
    return lib
