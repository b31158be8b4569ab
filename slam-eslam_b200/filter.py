"""Host-side mirror of the reference's PoseEstimator interface over the C ABI (ctypes).

Method names and argument meaning follow eslam::PoseEstimator / ParticleFilter<PoseParticle>
(src/PoseEstimator.hpp:119-155, src/ParticleFilter.hpp:18-178) so parity tests read like the reference's:
``init`` / ``project`` / ``update`` / ``normalizeWeights`` / ``resample`` / ``getCentroid`` /
``getBestParticleIndex`` / ``getParticles``.  Errors map to the reference's exceptions (RuntimeError with
the reference's message).  The C++ facade with the real class names is include/eslam/*.hpp.
"""
import ctypes as C

import numpy as np

from . import abi


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _as_f64(a, n=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float64)
    if n is not None and a.shape != (n,):
        raise ValueError("expected %d doubles" % n)
    return a


class FilterBase:
    """Calls ``<prefix><name>(ctx, ...)`` on a C library exporting the signatures of abi.SIGNATURES."""

    def __init__(self, lib, prefix, ctx):
        self._lib, self._prefix, self._ctx = lib, prefix, ctx
        self.last_result = abi.UpdateResult()

    def _fn(self, name):
        return getattr(self._lib, self._prefix + name)

    def _check(self, status):
        if status != abi.OK:
            raise RuntimeError(abi.ERROR_TEXT.get(status, "error %d" % status) + self._detail())
        return status

    def _detail(self):
        return ""

    # -- environment ---------------------------------------------------------------------------
    def setEnvironment(self, cells, scale, offset, g2l=None):
        """PoseEstimator::setEnvironment (src/PoseEstimator.cpp:47-62) with a dense grid:
        cells[h, w, 2] float32 {mean, stdev} (NaN mean = empty), scale/offset per axis (x, y)."""
        cells = np.ascontiguousarray(cells, dtype=np.float32)
        h, w, two = cells.shape
        assert two == 2
        g = None if g2l is None else np.ascontiguousarray(g2l, dtype=np.float64).reshape(12)
        self._check(self._fn("set_grid")(self._ctx, cells.ctypes.data_as(C.POINTER(C.c_float)), w, h,
                                         float(scale[0]), float(scale[1]), float(offset[0]), float(offset[1]), _dp(g)))

    # -- particles -----------------------------------------------------------------------------
    def init(self, n, mu, sigma, zpos=0.0, zsigma=0.0):
        """PoseEstimator::init(numParticles, mu, sigma, zpos, zsigma) (src/PoseEstimator.cpp:88-102)."""
        mu = (C.c_double * 3)(*[float(v) for v in mu])
        sg = (C.c_double * 3)(*[float(v) for v in sigma])
        self._check(self._fn("init_gaussian")(self._ctx, int(n), mu, sg, float(zpos), float(zsigma)))

    def size(self):
        return int(self._fn("particle_count")(self._ctx))

    def setParticles(self, x=None, y=None, theta=None, zpos=None, zsigma=None, weight=None):
        arrs = [_as_f64(a) for a in (x, y, theta, zpos, zsigma, weight)]
        n = next(len(a) for a in arrs if a is not None)
        self._check(self._fn("set_particles")(self._ctx, n, *[_dp(a) for a in arrs]))

    def getParticles(self):
        """ParticleFilter::getParticles (src/ParticleFilter.hpp:150-158) as a dict of SoA arrays."""
        n = self.size()
        names = ("x", "y", "theta", "zpos", "zsigma", "weight", "mprob")
        out = {k: np.empty(n, dtype=np.float64) for k in names}
        out["floating"] = np.empty(n, dtype=np.uint8)
        out["n_found"] = np.empty(n, dtype=np.uint8)
        self._check(self._fn("get_particles")(
            self._ctx, *[_dp(out[k]) for k in names],
            out["floating"].ctypes.data_as(C.POINTER(C.c_uint8)), out["n_found"].ctypes.data_as(C.POINTER(C.c_uint8))))
        return out

    # -- filter steps --------------------------------------------------------------------------
    def project(self, step):
        """PoseEstimator::project (src/PoseEstimator.cpp:184-242); ``step`` is an abi.Step."""
        self._check(self._fn("project")(self._ctx, C.byref(step)))

    def update(self, contacts):
        """PoseEstimator::update (src/PoseEstimator.cpp:244-255); ``contacts`` from abi.make_contacts.
        Returns the abi.UpdateResult (ess, resampled, max_weight, ...)."""
        res = abi.UpdateResult()
        self._check(self._fn("update")(self._ctx, contacts, len(contacts), C.byref(res)))
        self.last_result = res
        return res

    def normalizeWeights(self):
        ess = C.c_double()
        self._check(self._fn("normalize_weights")(self._ctx, C.byref(ess)))
        return ess.value

    def resample(self):
        self._check(self._fn("resample")(self._ctx))

    def resampleStratified(self, samples):
        """ParticleFilter::resample_stratified(samples) (src/ParticleFilter.hpp:85-108)."""
        self._check(self._fn("resample_n")(self._ctx, int(samples), 0))

    def resampleMultinomial(self, samples):
        """ParticleFilter::resample_multinomial(samples) (src/ParticleFilter.hpp:120-147)."""
        self._check(self._fn("resample_n")(self._ctx, int(samples), 1))

    def getAncestors(self):
        a = np.empty(self.size(), dtype=np.uint32)
        self._check(self._fn("get_ancestors")(self._ctx, a.ctypes.data_as(C.POINTER(C.c_uint32))))
        return a

    def getCdf(self):
        a = np.empty(self.size(), dtype=np.float64)
        self._check(self._fn("get_cdf")(self._ctx, _dp(a)))
        return a

    def getCentroid(self):
        out = np.empty(4, dtype=np.float64)
        self._check(self._fn("centroid")(self._ctx, _dp(out)))
        return out

    def getBestParticleIndex(self):
        idx = C.c_uint64()
        self._check(self._fn("best_index")(self._ctx, C.byref(idx)))
        return idx.value

    # -- ContactModel seam ---------------------------------------------------------------------
    def evaluatePoses(self, x, y, theta, zpos, meas_var, contacts, z_var_in=None):
        """ContactModel::evaluatePose + evaluateWeight (+ updateZPositionEstimate when z_var_in is given)
        for explicit poses (src/ContactModel.cpp:117-224,262-340)."""
        x, y, theta, zpos, meas_var = (_as_f64(np.atleast_1d(a)) for a in (x, y, theta, zpos, meas_var))
        n, c = len(x), len(contacts)
        zvi = None if z_var_in is None else _as_f64(np.atleast_1d(z_var_in))
        out = {
            "ok": np.zeros(n, dtype=np.uint8), "weight": np.zeros(n), "z_delta": np.zeros(n), "z_var": np.zeros(n),
            "pose_var": np.zeros(n), "n_contacts": np.zeros(n, dtype=np.int32), "zpos_out": np.zeros(n),
            "z_var_out": np.zeros(n), "cell_index": np.zeros((n, c), dtype=np.int32),
            "cp": np.zeros((n, abi.MAX_CP, 5)),
        }
        self._check(self._fn("evaluate_poses")(
            self._ctx, n, _dp(x), _dp(y), _dp(theta), _dp(zpos), _dp(meas_var), _dp(zvi), contacts, c,
            out["ok"].ctypes.data_as(C.POINTER(C.c_uint8)), _dp(out["weight"]), _dp(out["z_delta"]), _dp(out["z_var"]),
            _dp(out["pose_var"]), out["n_contacts"].ctypes.data_as(C.POINTER(C.c_int32)), _dp(out["zpos_out"]),
            _dp(out["z_var_out"]), out["cell_index"].ctypes.data_as(C.POINTER(C.c_int32)), _dp(out["cp"])))
        return out

    # -- per-particle map patches --------------------------------------------------------------
    def processMap(self, scan_cells, scale, offset, match=True, update=True):
        """EmbodiedSlamFilter::processMap (src/EmbodiedSlamFilter.cpp:179-232) with the scan as a dense grid:
        scan_cells[h, w, 2] float32 {mean, stdev} in the yaw-compensated body frame (NaN mean = no surface)."""
        cells = np.ascontiguousarray(scan_cells, dtype=np.float32)
        h, w, two = cells.shape
        assert two == 2
        scan = abi.Scan(cells=cells.ctypes.data_as(C.POINTER(C.c_float)), width=w, height=h, scale_x=float(scale[0]),
                        scale_y=float(scale[1]), offset_x=float(offset[0]), offset_y=float(offset[1]))
        self._check(self._fn("process_map")(self._ctx, C.byref(scan), 1 if match else 0, 1 if update else 0))

    def poseMoments(self):
        """Weighted mean / covariance of (x, y): the one-component mixture of a PoseDistribution (src/PoseParticle.hpp:88-114)."""
        g = abi.Gaussian2d()
        self._check(self._fn("pose_moments")(self._ctx, C.byref(g)))
        return dict(weight=g.weight, mean=np.array(list(g.mean)), cov=np.array(list(g.cov)).reshape(2, 2))

    # -- SurfaceHash (src/SurfaceHash.hpp, src/PoseEstimator.cpp:75-86,130-182) -------------------
    def hashCreate(self, slope_bins=20, angular_steps=16):
        """SurfaceHash::create on the attached grid (SurfaceHashConfig defaults: 20 slope bins, 16 headings)."""
        self._check(self._fn("hash_create")(self._ctx, int(slope_bins), int(angular_steps)))
        self._hash_bins = int(slope_bins)

    def hashInfo(self):
        n = C.c_uint64()
        counts = np.zeros(self._hash_bins ** 2, dtype=np.uint32)
        self._check(self._fn("hash_info")(self._ctx, C.byref(n), counts.ctypes.data_as(C.POINTER(C.c_uint32))))
        return int(n.value), counts

    def hashGet(self, first, count):
        out = {k: np.zeros(count) for k in ("x", "y", "theta", "z")}
        out["bucket"] = np.zeros(count, dtype=np.int32)
        self._check(self._fn("hash_get")(self._ctx, int(first), int(count), _dp(out["x"]), _dp(out["y"]), _dp(out["theta"]), _dp(out["z"]),
                                         out["bucket"].ctypes.data_as(C.POINTER(C.c_int32))))
        return out

    def initFromHash(self, n):
        """PoseEstimator::init(numParticles, hash) (src/PoseEstimator.cpp:75-86)."""
        self._check(self._fn("init_from_hash")(self._ctx, int(n)))

    def sampleFromHash(self, contacts, replace_percentage=0.05, avg_factor=0.1):
        """PoseEstimator::sampleFromHash (src/PoseEstimator.cpp:130-182); returns the number of replaced particles."""
        r = C.c_uint64()
        self._check(self._fn("sample_from_hash")(self._ctx, float(replace_percentage), float(avg_factor), contacts, len(contacts), C.byref(r)))
        return int(r.value)

    def getPatches(self, patch_size):
        n = self.size()
        cells = np.empty((n, patch_size, patch_size, 2), dtype=np.float32)
        origin = np.empty((n, 2), dtype=np.int32)
        self._check(self._fn("get_patches")(self._ctx, cells.ctypes.data_as(C.POINTER(C.c_float)),
                                            origin.ctypes.data_as(C.POINTER(C.c_int32))))
        return cells, origin


class GpuPoseEstimator(FilterBase):
    """eslam::PoseEstimator on one B200 (one context per GPU / rank)."""

    def __init__(self, config=None, device=0, shard=None):
        lib = abi.load_library()
        cfg = config if config is not None else abi.default_config()
        ctx = C.c_void_p()
        self._keepalive = (cfg, shard)
        st = lib.eslam_gpu_create(C.byref(cfg), int(device), None if shard is None else C.byref(shard), C.byref(ctx))
        if st != abi.OK:
            raise RuntimeError("eslam_gpu_create failed with status %d (is a CUDA device visible?)" % st)
        super().__init__(lib, "eslam_gpu_", ctx)
        self.config = cfg

    def _detail(self):
        msg = self._lib.eslam_gpu_last_error(self._ctx)
        return (": " + msg.decode()) if msg else ""

    def commInit(self, unique_id):
        """Join the NCCL communicator of a sharded filter; unique_id = 128 bytes from abi nccl_unique_id()."""
        buf = (C.c_uint8 * 128).from_buffer_copy(bytes(unique_id))
        self._check(self._lib.eslam_gpu_comm_init(self._ctx, buf))

    def close(self):
        if self._ctx:
            self._lib.eslam_gpu_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def stepFused(self, step, contacts, want_result=True):
        """EmbodiedSlamFilter::update(contact) with the measurement threshold met
        (src/EmbodiedSlamFilter.cpp:353-369): project + update in one launch sequence."""
        res = abi.UpdateResult()
        self._check(self._lib.eslam_gpu_step_fused(self._ctx, C.byref(step), contacts, len(contacts),
                                                   C.byref(res) if want_result else None))
        if want_result:
            self.last_result = res
        return res

    def runSteps(self, step, contacts, steps, kernel_events=True):
        """Enqueue ``steps`` fused steps back to back; returns (elapsed_ms, top_kernel_ms, last result).  kernel_events=False:
        no event between the first and the last launch (top_kernel_ms is 0 then)."""
        ms, top = C.c_float(), C.c_float()
        res = abi.UpdateResult()
        self._check(self._lib.eslam_gpu_run_steps(self._ctx, C.byref(step), contacts, len(contacts), int(steps),
                                                  C.byref(ms), C.byref(top) if kernel_events else None, C.byref(res)))
        self.last_result = res
        return ms.value, top.value, res

    def stepTimes(self):
        """Step periods (ms) of the last runSteps."""
        n = self._lib.eslam_gpu_step_times(self._ctx, None, 0)
        buf = (C.c_float * max(n, 1))()
        self._lib.eslam_gpu_step_times(self._ctx, buf, n)
        return [buf[i] for i in range(n)]

    def ioBytes(self, n_pts):
        out = (C.c_uint64 * 2)()
        self._check(self._lib.eslam_gpu_io_bytes(self._ctx, int(n_pts), out))
        return int(out[0]), int(out[1])

    def synchronize(self):
        self._check(self._lib.eslam_gpu_synchronize(self._ctx))

    def getWeightsSum(self):
        """ParticleFilter::getWeightsSum (src/ParticleFilter.hpp:34-39)."""
        v = C.c_double()
        self._check(self._lib.eslam_gpu_weights_sum(self._ctx, C.byref(v)))
        return v.value

    def debugCounters(self):
        """{fallbacks, boundaries, error_flags, resamples} of the exact running-sum machinery."""
        out = (C.c_uint64 * 4)()
        self._check(self._lib.eslam_gpu_debug_counters(self._ctx, out))
        return dict(fallbacks=int(out[0]), boundaries=int(out[1]), error_flags=int(out[2]), resamples=int(out[3]))

    def pathCounters(self):
        """{lean, deferred}: particles binned for the lean weight kernel / evaluated by the generic one in the last update."""
        out = (C.c_uint64 * 2)()
        self._check(self._lib.eslam_gpu_path_counters(self._ctx, out))
        return dict(lean=int(out[0]), deferred=int(out[1]))

    def launchCount(self):
        return int(self._lib.eslam_gpu_launch_count(self._ctx))
