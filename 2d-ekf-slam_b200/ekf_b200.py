"""Thin ctypes binding over the C ABI (include/ekf_slam_b200.h, include/ekf_synth.h).

Used by tests/, bench.py and __graft_entry__.py. It adds nothing to the arithmetic: every method
is one C-ABI call. There is no CPU fallback: if lib/libekf_slam_b200.so is missing, or no B200
is usable, calls raise.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.path.join(HERE, "lib")

OK, ERR_CUDA, ERR_BAD_ARG, ERR_CAPACITY, ERR_NO_DEVICE, ERR_UNSUPPORTED = range(6)
DECISION_NONE, DECISION_NEW, DECISION_OLD, DECISION_IGNORE, DECISION_DROPPED = -1, 0, 1, 2, 3
REGIME_AUTO, REGIME_BATCH, REGIME_LARGE = 0, 1, 2
BATCH_KERNEL_AUTO, BATCH_KERNEL_SMEM, BATCH_KERNEL_TILE, BATCH_KERNEL_STILE, BATCH_KERNEL_DTILE = 0, 1, 2, 3, 4
RECORD_HEADER = 8

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)


def record_len(max_meas):
    return RECORD_HEADER + 6 * max_meas


def _dp(a):
    return a.ctypes.data_as(c_dp) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(c_ip) if a is not None else None


class EkfError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("ekf_slam_b200 error %d: %s" % (code, msg))
        self.code = code


class Config(C.Structure):
    _fields_ = [("sigma_v", C.c_double), ("sigma_w", C.c_double), ("deg2rad_pi", C.c_double),
                ("two_pi", C.c_double), ("cond_max", C.c_double), ("mahal_init", C.c_double),
                ("gamma_max", C.c_int32), ("gamma_min", C.c_int32), ("regime", C.c_int32),
                ("batch_kernel", C.c_int32)]


class RunOutputs(C.Structure):
    _fields_ = [("decision", c_ip), ("lm_index", c_ip), ("mahal", c_dp), ("pose_trace", c_dp),
                ("final_pose", c_dp), ("final_nlm", c_ip)]


class SynthConfig(C.Structure):
    _fields_ = [("n_landmarks", C.c_int32), ("steps_per_lap", C.c_int32), ("max_meas", C.c_int32),
                ("compass_every", C.c_int32), ("dt", C.c_double), ("radius", C.c_double),
                ("ring_offset", C.c_double), ("sigma_v", C.c_double), ("sigma_w", C.c_double),
                ("sigma_range", C.c_double), ("sigma_bearing", C.c_double), ("min_range", C.c_double),
                ("max_range", C.c_double), ("fov", C.c_double), ("sigma_compass", C.c_double),
                ("compass_R", C.c_double), ("seed", C.c_uint64)]


_synth = None
_core = None


def synth_lib():
    global _synth
    if _synth is None:
        L = C.CDLL(os.path.join(LIB_DIR, "libekf_synth.so"))
        L.ekf_synth_default_config.argtypes = [C.POINTER(SynthConfig), C.c_int]
        L.ekf_synth_record_len.argtypes = [C.POINTER(SynthConfig)]
        L.ekf_synth_world.argtypes = [C.POINTER(SynthConfig), c_dp]
        L.ekf_synth_true_pose.argtypes = [C.POINTER(SynthConfig), C.c_long, c_dp]
        L.ekf_synth_generate.argtypes = [C.POINTER(SynthConfig), C.c_long, C.c_int, C.c_long, C.c_int, c_dp, c_ip,
                                         C.c_int]
        L.ekf_synth_measurement_from_feature.argtypes = [C.c_double, C.c_double, c_dp, c_dp]
        _synth = L
    return _synth


def core_lib():
    """The CUDA library. Raises if it has not been built (no fallback)."""
    global _core
    if _core is None:
        # EKF_B200_LIB: profiling tooling only (instrumented builds, `make timing`); never a fallback
        path = os.environ.get("EKF_B200_LIB") or os.path.join(LIB_DIR, "libekf_slam_b200.so")
        if not os.path.exists(path):
            raise EkfError(ERR_NO_DEVICE, "CUDA extension missing: %s (run __graft_entry__.build())" % path)
        L = C.CDLL(path)
        H = C.c_void_p
        L.ekf_default_config.argtypes = [C.POINTER(Config)]
        L.ekf_create.argtypes = [C.POINTER(H), C.c_int, C.c_int, C.c_int, C.POINTER(Config)]
        L.ekf_destroy.argtypes = [H]
        L.ekf_reset.argtypes = [H]
        L.ekf_n_filters.argtypes = [H]
        L.ekf_max_landmarks.argtypes = [H]
        L.ekf_regime.argtypes = [H]
        L.ekf_set_batch_kernel.argtypes = [H, C.c_int]
        if hasattr(L, "ekf_large_downdate_kernel"):
            L.ekf_large_downdate_kernel.argtypes = [H]
        if hasattr(L, "ekf_capacity_flags"):        # absent from older profiling builds (EKF_B200_LIB)
            L.ekf_capacity_flags.argtypes = [H, C.POINTER(C.c_int), C.c_int]
        L.ekf_set_state.argtypes = [H, C.c_int, C.c_int, c_dp, c_dp, C.c_int]
        L.ekf_get_state.argtypes = [H, C.c_int, C.POINTER(C.c_int), c_dp, c_dp, C.c_int]
        L.ekf_get_pose.argtypes = [H, c_dp, c_ip]
        L.ekf_get_cov_block.argtypes = [H, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_dp, C.c_int]
        L.ekf_propagate.argtypes = [H, c_dp, c_dp, c_dp, C.c_int]
        L.ekf_update.argtypes = [H, C.c_int, c_dp, c_dp, c_ip, c_ip, c_dp]
        L.ekf_update_compass.argtypes = [H, c_dp, c_dp, C.POINTER(C.c_uint8)]
        L.ekf_run.argtypes = [H, C.c_int, C.c_int, c_dp, C.POINTER(RunOutputs)]
        L.ekf_upload_records.argtypes = [H, C.c_int, C.c_int, c_dp]
        L.ekf_run_resident.argtypes = [H, C.c_int, C.c_int]
        L.ekf_download_outputs.argtypes = [H, C.POINTER(RunOutputs)]
        L.ekf_sync.argtypes = [H]
        L.ekf_last_error.argtypes = [H]
        L.ekf_last_error.restype = C.c_char_p
        L.ekf_host_alloc.argtypes = [C.c_size_t]
        L.ekf_host_alloc.restype = C.c_void_p
        L.ekf_host_free.argtypes = [C.c_void_p]
        L.ekf_timer_start.argtypes = [H]
        L.ekf_timer_stop.argtypes = [H, C.POINTER(C.c_float)]
        L.ekf_kernel_launches.argtypes = [H]
        L.ekf_kernel_launches.restype = C.c_longlong
        L.ekf_kernel_time.argtypes = [H, C.POINTER(C.c_float), C.POINTER(C.c_int)]
        L.ekf_device_info.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                      C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
        L.ekf_measure_fp64_peak.argtypes = [C.c_int, c_dp]
        L.ekf_debug_phase_cycles.argtypes = [C.POINTER(C.c_longlong)]
        L.ekf_debug_stile_timestamps.argtypes = [C.POINTER(C.c_longlong)]
        L.ekf_debug_dtile_timestamps.argtypes = [C.POINTER(C.c_longlong)]
        c_u32p = C.POINTER(C.c_uint32)
        c_u8p = C.POINTER(C.c_uint8)
        L.ekf_hough_create.argtypes = [C.POINTER(H), C.c_int, C.c_int]
        L.ekf_hough_destroy.argtypes = [H]
        L.ekf_hough_tables.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.ekf_hough_tables.restype = None
        L.ekf_hough_get_lines.argtypes = [H, C.c_int, C.c_int, c_dp, c_dp, c_u32p, c_dp, C.c_int, c_ip, c_ip, c_ip, c_u8p]
        L.ekf_hough_get_features.argtypes = [H, C.c_int, C.c_int, c_dp, c_dp, c_u32p, c_dp, c_dp, c_dp, C.c_int, c_ip, c_dp,
                                             c_dp, C.c_int, c_ip, c_dp, C.c_int, c_ip]
        L.ekf_hough_upload.argtypes = [H, C.c_int, C.c_int, c_dp, c_dp, c_u32p]
        L.ekf_hough_run_resident.argtypes = [H, C.c_int]
        L.ekf_hough_download.argtypes = [H, c_dp, C.c_int, c_ip, c_ip, c_ip]
        L.ekf_hough_kernel_time.argtypes = [H, C.POINTER(C.c_float), C.POINTER(C.c_int)]
        L.ekf_hough_sync.argtypes = [H]
        L.ekf_hough_lines_from_peaks.argtypes = [c_ip, c_ip, c_dp, C.c_int]
        L.ekf_hough_last_error.argtypes = [H]
        L.ekf_hough_last_error.restype = C.c_char_p
        L.ekf_sharded_create.argtypes = [C.POINTER(H), C.c_int, c_ip, C.c_int, C.POINTER(Config)]
        L.ekf_sharded_destroy.argtypes = [H]
        L.ekf_sharded_reset.argtypes = [H]
        L.ekf_sharded_n_shards.argtypes = [H]
        L.ekf_sharded_max_landmarks.argtypes = [H]
        L.ekf_sharded_columns.argtypes = [H, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.ekf_sharded_set_state.argtypes = [H, C.c_int, c_dp, c_dp, C.c_int]
        L.ekf_sharded_get_state.argtypes = [H, C.POINTER(C.c_int), c_dp, c_dp, C.c_int]
        L.ekf_sharded_get_pose.argtypes = [H, c_dp, c_ip]
        L.ekf_sharded_get_replica.argtypes = [H, C.c_int, C.POINTER(C.c_int), c_dp, c_dp]
        L.ekf_sharded_propagate.argtypes = [H, C.c_double, C.c_double, C.c_double]
        L.ekf_sharded_update.argtypes = [H, C.c_int, c_dp, c_dp, c_ip, c_ip, c_dp]
        L.ekf_sharded_update_compass.argtypes = [H, C.c_double, C.c_double]
        L.ekf_sharded_run.argtypes = [H, C.c_int, C.c_int, c_dp, C.POINTER(RunOutputs)]
        L.ekf_sharded_last_run_ms.argtypes = [H, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.ekf_sharded_kernel_launches.argtypes = [H]
        L.ekf_sharded_kernel_launches.restype = C.c_longlong
        L.ekf_sharded_run_mode.argtypes = [H]
        L.ekf_sharded_last_error.argtypes = [H]
        L.ekf_sharded_last_error.restype = C.c_char_p
        _core = L
    return _core


# ---- synthetic driver ---------------------------------------------------------------------------
class Synth:
    def __init__(self, n_landmarks, **overrides):
        self.cfg = SynthConfig()
        synth_lib().ekf_synth_default_config(C.byref(self.cfg), n_landmarks)
        for k, v in overrides.items():
            setattr(self.cfg, k, v)

    @property
    def record_len(self):
        return record_len(self.cfg.max_meas)

    def world(self):
        xy = np.zeros((self.cfg.n_landmarks, 2))
        synth_lib().ekf_synth_world(C.byref(self.cfg), _dp(xy))
        return xy

    def true_pose(self, t):
        p = np.zeros(3)
        synth_lib().ekf_synth_true_pose(C.byref(self.cfg), t, _dp(p))
        return p

    def scan(self, t):
        """Synthetic LMS-200 scan at the true pose of step t -> (local x [181] mm, local y [181] mm, range [181] mm)."""
        x = np.zeros(181)
        y = np.zeros(181)
        r = np.zeros(181, np.uint32)
        L = synth_lib()
        L.ekf_synth_scan.argtypes = [C.c_void_p, C.c_long, c_dp, c_dp, C.POINTER(C.c_uint32)]
        n = L.ekf_synth_scan(C.byref(self.cfg), t, _dp(x), _dp(y), r.ctypes.data_as(C.POINTER(C.c_uint32)))
        assert n == 181
        return x, y, r

    def generate(self, n_filters, n_steps, f0=0, t0=0, out=None, want_ids=False, n_threads=0):
        L = self.record_len
        if out is None:
            out = np.zeros((n_filters, n_steps, L))
        assert out.shape == (n_filters, n_steps, L) and out.dtype == np.float64 and out.flags.c_contiguous
        ids = np.zeros((n_filters, n_steps, max(self.cfg.max_meas, 1)), np.int32) if want_ids else None
        rc = synth_lib().ekf_synth_generate(C.byref(self.cfg), f0, n_filters, t0, n_steps, _dp(out), _ip(ids),
                                            n_threads)
        if rc:
            raise ValueError("ekf_synth_generate: bad arguments")
        return (out, ids) if want_ids else out


def measurement_from_feature(fx_mm, fy_mm):
    z = np.zeros(2)
    R = np.zeros(4)
    synth_lib().ekf_synth_measurement_from_feature(fx_mm, fy_mm, _dp(z), _dp(R))
    return z, R


# ---- pinned host buffers ----------------------------------------------------------------------
class PinnedArray:
    """numpy view over cudaMallocHost memory (ekf_host_alloc)."""

    def __init__(self, shape, dtype=np.float64):
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self.ptr = core_lib().ekf_host_alloc(max(self.nbytes, 1))
        if not self.ptr:
            raise EkfError(ERR_CUDA, "ekf_host_alloc(%d) failed" % self.nbytes)
        buf = (C.c_char * max(self.nbytes, 1)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if self.ptr:
            self.array = None
            core_lib().ekf_host_free(self.ptr)
            self.ptr = None


# ---- filter batch -----------------------------------------------------------------------------
class FilterBatch:
    """n_filters independent EKF-SLAM filters on one GPU (one ekf_handle)."""

    def __init__(self, n_filters, max_landmarks, device=0, regime=REGIME_AUTO, **cfg_overrides):
        self.L = core_lib()
        cfg = Config()
        self.L.ekf_default_config(C.byref(cfg))
        cfg.regime = regime
        for k, v in cfg_overrides.items():
            setattr(cfg, k, v)
        self.h = C.c_void_p()
        rc = self.L.ekf_create(C.byref(self.h), device, n_filters, max_landmarks, C.byref(cfg))
        if rc:
            msg = self.L.ekf_last_error(None)
            self.h = None
            raise EkfError(rc, msg.decode() if msg else "")
        self.F = n_filters
        self.cap_lm = max_landmarks
        self.cap_n = 3 + 2 * max_landmarks
        self._pinned = []

    def close(self):
        if getattr(self, "h", None):
            self.L.ekf_destroy(self.h)
            self.h = None
            for pa in self._pinned:
                pa.free()
            self._pinned = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc:
            msg = self.L.ekf_last_error(self.h)
            raise EkfError(rc, msg.decode() if msg else "")

    @property
    def regime(self):
        return self.L.ekf_regime(self.h)

    def set_batch_kernel(self, kernel):
        self._chk(self.L.ekf_set_batch_kernel(self.h, kernel))

    def reset(self):
        self._chk(self.L.ekf_reset(self.h))

    def sync(self, allow_capacity=False):
        rc = self.L.ekf_sync(self.h)
        if not (allow_capacity and rc == ERR_CAPACITY):
            self._chk(rc)
        return rc

    def large_downdate_kernel(self):
        """'large_downdate_tma' / 'large_downdate' (the plain double2 sweep) / None for the batch regime."""
        v = self.L.ekf_large_downdate_kernel(self.h)
        return {1: "large_downdate_tma", 0: "large_downdate"}.get(v)

    def capacity_flags(self, clear=False):
        """Filters that dropped a New association at capacity since the flags were last cleared."""
        n = C.c_int()
        self._chk(self.L.ekf_capacity_flags(self.h, C.byref(n), 1 if clear else 0))
        return n.value

    def set_state(self, filt, x, P, symmetric=False):
        """x (n), P (n x n, numpy row-major). symmetric=True skips the transpose copy to column-major
        (P must equal P.T bit for bit anyway; the C ABI checks)."""
        x = np.ascontiguousarray(x, np.float64)
        n = len(x)
        Pa = np.asarray(P, np.float64)
        Pc = np.ascontiguousarray(Pa if symmetric else Pa.T)  # column-major
        self._chk(self.L.ekf_set_state(self.h, filt, (n - 3) // 2, _dp(x), _dp(Pc), n))

    def get_state(self, filt):
        nl = C.c_int()
        x = np.zeros(self.cap_n)
        P = np.zeros((self.cap_n, self.cap_n))
        self._chk(self.L.ekf_get_state(self.h, filt, C.byref(nl), _dp(x), _dp(P), self.cap_n))
        n = 3 + 2 * nl.value
        return x[:n].copy(), P[:n, :n].T.copy()

    def get_cov_block(self, filt, r0, c0, nr, nc):
        out = np.zeros((nc, nr))
        self._chk(self.L.ekf_get_cov_block(self.h, filt, r0, c0, nr, nc, _dp(out), nr))
        return out.T.copy()

    def get_pose(self):
        p = np.zeros((self.F, 3))
        nl = np.zeros(self.F, np.int32)
        self._chk(self.L.ekf_get_pose(self.h, _dp(p), _ip(nl)))
        return p, nl

    def propagate(self, vel_mm_s, rotvel_deg_s, dt):
        v = np.ascontiguousarray(np.broadcast_to(vel_mm_s, (self.F,)), np.float64)
        w = np.ascontiguousarray(np.broadcast_to(rotvel_deg_s, (self.F,)), np.float64)
        d = np.ascontiguousarray(np.atleast_1d(dt), np.float64)
        stride = 0 if d.size == 1 else 1
        assert d.size in (1, self.F)
        self._chk(self.L.ekf_propagate(self.h, _dp(v), _dp(w), _dp(d), stride))

    def update(self, z, R, want=True):
        """z [F][n_z][2], R [F][n_z][4]. Returns (decision, lm_index, mahal) arrays [F][n_z]."""
        z = np.ascontiguousarray(z, np.float64).reshape(self.F, -1, 2)
        n_z = z.shape[1]
        R = np.ascontiguousarray(R, np.float64).reshape(self.F, n_z, 4)
        if not want:
            self._chk(self.L.ekf_update(self.h, n_z, _dp(z), _dp(R), None, None, None))
            return None
        dec = np.zeros((self.F, n_z), np.int32)
        idx = np.zeros((self.F, n_z), np.int32)
        mah = np.zeros((self.F, n_z))
        rc = self.L.ekf_update(self.h, n_z, _dp(z), _dp(R), _ip(dec), _ip(idx), _dp(mah))
        if rc != ERR_CAPACITY:
            self._chk(rc)
        return dec, idx, mah

    def update_compass(self, z, R, valid=None):
        z = np.ascontiguousarray(np.broadcast_to(z, (self.F,)), np.float64)
        R = np.ascontiguousarray(np.broadcast_to(R, (self.F,)), np.float64)
        vp = None
        if valid is not None:
            valid = np.ascontiguousarray(valid, np.uint8)
            vp = valid.ctypes.data_as(C.POINTER(C.c_uint8))
        self._chk(self.L.ekf_update_compass(self.h, _dp(z), _dp(R), vp))

    def alloc_outputs(self, T, M, trace=True, pose_trace=False, pinned=False):
        """Host output buffers for run()/download_outputs(); pinned=True uses cudaMallocHost memory."""
        def mk(shape, dtype):
            if pinned:
                pa = PinnedArray(shape, dtype)
                self._pinned.append(pa)
                return pa.array
            return np.zeros(shape, dtype)
        M = max(M, 1)
        o = {
            "decision": mk((self.F, T, M), np.int32) if trace else None,
            "index": mk((self.F, T, M), np.int32) if trace else None,
            "mahal": mk((self.F, T, M), np.float64) if trace else None,
            "pose_trace": mk((self.F, T, 3), np.float64) if pose_trace else None,
            "final_pose": mk((self.F, 3), np.float64),
            "final_nlm": mk((self.F,), np.int32),
        }
        o["_c"] = RunOutputs(_ip(o["decision"]), _ip(o["index"]), _dp(o["mahal"]), _dp(o["pose_trace"]),
                             _dp(o["final_pose"]), _ip(o["final_nlm"]))
        return o

    @staticmethod
    def output_bytes(o):
        return int(sum(v.nbytes for k, v in o.items() if k != "_c" and v is not None))

    def run(self, records, max_meas, trace=True, pose_trace=False, allow_capacity=False, outputs=None):
        """End-to-end fused run: H2D records, T steps per filter, D2H outputs."""
        F, T, L = records.shape
        assert F == self.F and L == record_len(max_meas) and records.dtype == np.float64
        o = outputs if outputs is not None else self.alloc_outputs(T, max_meas, trace, pose_trace)
        rc = self.L.ekf_run(self.h, T, max_meas, _dp(records), C.byref(o["_c"]))
        if not (allow_capacity and rc == ERR_CAPACITY):
            self._chk(rc)
        return o

    def upload_records(self, records, max_meas):
        F, T, L = records.shape
        assert F == self.F and L == record_len(max_meas) and records.dtype == np.float64
        self._rec_shape = (T, max_meas)
        self._chk(self.L.ekf_upload_records(self.h, T, max_meas, _dp(records)))

    def run_resident(self, trace=False, pose_trace=False):
        self._chk(self.L.ekf_run_resident(self.h, int(trace), int(pose_trace)))

    def download_outputs(self, trace=True, pose_trace=False, outputs=None, allow_capacity=False):
        T, M = self._rec_shape
        o = outputs if outputs is not None else self.alloc_outputs(T, M, trace, pose_trace)
        rc = self.L.ekf_download_outputs(self.h, C.byref(o["_c"]))
        if not (allow_capacity and rc == ERR_CAPACITY):
            self._chk(rc)
        return o

    def timer_start(self):
        self._chk(self.L.ekf_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float()
        self._chk(self.L.ekf_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def kernel_launches(self):
        return self.L.ekf_kernel_launches(self.h)

    def kernel_time(self):
        ms = C.c_float()
        n = C.c_int()
        self._chk(self.L.ekf_kernel_time(self.h, C.byref(ms), C.byref(n)))
        return ms.value, n.value


class ShardedMap:
    """ONE large map whose covariance is column-sharded over several GPUs (one ekf_sharded handle;
    devices may repeat an ordinal to put several shards on one GPU)."""

    def __init__(self, devices, max_landmarks, **cfg_overrides):
        self.L = core_lib()
        cfg = Config()
        self.L.ekf_default_config(C.byref(cfg))
        for k, v in cfg_overrides.items():
            setattr(cfg, k, v)
        dev = np.ascontiguousarray(devices, np.int32)
        self.h = C.c_void_p()
        rc = self.L.ekf_sharded_create(C.byref(self.h), len(dev), _ip(dev), max_landmarks, C.byref(cfg))
        if rc:
            msg = self.L.ekf_sharded_last_error(None)
            self.h = None
            raise EkfError(rc, msg.decode() if msg else "")
        self.G = len(dev)
        self.cap_lm = max_landmarks
        self.cap_n = 3 + 2 * max_landmarks

    def close(self):
        if getattr(self, "h", None):
            self.L.ekf_sharded_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc:
            msg = self.L.ekf_sharded_last_error(self.h)
            raise EkfError(rc, msg.decode() if msg else "")

    def columns(self, shard):
        c0, c1 = C.c_int(), C.c_int()
        self._chk(self.L.ekf_sharded_columns(self.h, shard, C.byref(c0), C.byref(c1)))
        return c0.value, c1.value

    def reset(self):
        self._chk(self.L.ekf_sharded_reset(self.h))

    def set_state(self, x, P, symmetric=False):
        x = np.ascontiguousarray(x, np.float64)
        n = len(x)
        Pa = np.asarray(P, np.float64)
        Pc = np.ascontiguousarray(Pa if symmetric else Pa.T)
        self._chk(self.L.ekf_sharded_set_state(self.h, (n - 3) // 2, _dp(x), _dp(Pc), n))

    def get_state(self, want_P=True):
        nl = C.c_int()
        self._chk(self.L.ekf_sharded_get_pose(self.h, None, C.cast(C.byref(nl), c_ip)))
        n = 3 + 2 * nl.value
        x = np.zeros(n)
        P = np.zeros((n, n)) if want_P else None
        self._chk(self.L.ekf_sharded_get_state(self.h, C.byref(nl), _dp(x), _dp(P), n))
        return x, (P.T.copy() if want_P else None)

    def get_replica(self, shard):
        nl = C.c_int()
        x = np.zeros(self.cap_n)
        prr = np.zeros(9)
        self._chk(self.L.ekf_sharded_get_replica(self.h, shard, C.byref(nl), _dp(x), _dp(prr)))
        return nl.value, x[:3 + 2 * nl.value].copy(), prr.reshape(3, 3).T.copy()

    def get_pose(self):
        pose = np.zeros(3)
        nl = np.zeros(1, np.int32)
        self._chk(self.L.ekf_sharded_get_pose(self.h, _dp(pose), _ip(nl)))
        return pose, int(nl[0])

    def propagate(self, vel_mm_s, rotvel_deg_s, dt):
        self._chk(self.L.ekf_sharded_propagate(self.h, vel_mm_s, rotvel_deg_s, dt))

    def update(self, z, R, allow_capacity=False):
        """z [n_z][2], R [n_z][4] (column-major 2x2): ONE doUpdate call. Returns decision, index, mahal [n_z]."""
        z = np.ascontiguousarray(z, np.float64).reshape(-1, 2)
        R = np.ascontiguousarray(R, np.float64).reshape(-1, 4)
        nz = len(z)
        dec = np.zeros(nz, np.int32)
        idx = np.zeros(nz, np.int32)
        mah = np.zeros(nz)
        rc = self.L.ekf_sharded_update(self.h, nz, _dp(z), _dp(R), _ip(dec), _ip(idx), _dp(mah))
        if not (allow_capacity and rc == ERR_CAPACITY):
            self._chk(rc)
        return dec, idx, mah

    def update_compass(self, z, R):
        self._chk(self.L.ekf_sharded_update_compass(self.h, z, R))

    def run(self, records, max_meas, trace=True, pose_trace=False, allow_capacity=False):
        """records [T][L] (or [1][T][L]) of the one map. Returns the outputs dict of FilterBatch.run with F = 1."""
        records = np.ascontiguousarray(records, np.float64)
        if records.ndim == 3:
            assert records.shape[0] == 1
            records = records[0]
        T, L = records.shape
        assert L == record_len(max_meas)
        M = max(max_meas, 1)
        o = {
            "decision": np.zeros((1, T, M), np.int32) if trace else None,
            "index": np.zeros((1, T, M), np.int32) if trace else None,
            "mahal": np.zeros((1, T, M)) if trace else None,
            "pose_trace": np.zeros((1, T, 3)) if pose_trace else None,
            "final_pose": np.zeros((1, 3)),
            "final_nlm": np.zeros((1,), np.int32),
        }
        o["_c"] = RunOutputs(_ip(o["decision"]), _ip(o["index"]), _dp(o["mahal"]), _dp(o["pose_trace"]),
                             _dp(o["final_pose"]), _ip(o["final_nlm"]))
        rc = self.L.ekf_sharded_run(self.h, T, max_meas, _dp(records), C.byref(o["_c"]))
        if not (allow_capacity and rc == ERR_CAPACITY):
            self._chk(rc)
        return o

    def last_run_ms(self):
        a, b = C.c_float(), C.c_float()
        self._chk(self.L.ekf_sharded_last_run_ms(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def kernel_launches(self):
        return int(self.L.ekf_sharded_kernel_launches(self.h))

    def run_mode(self):
        """What ekf_sharded_run does: the sweep kernel it launches and whether the O(n) chain is overlapped."""
        return {2: "large_downdate_tma<2,0> on the shard's slab, look-ahead run",
                1: "shard_downdate<2>, look-ahead run",
                0: "shard_downdate<2>, event chain"}[int(self.L.ekf_sharded_run_mode(self.h))]


HOUGH_THETA, HOUGH_RADIUS, HOUGH_PEAKS = 180, 1601, 200


def hough_tables():
    """COS_ARRAY / SIN_ARRAY of the reference constructor (host only)."""
    c = np.zeros(HOUGH_THETA, np.float32)
    s = np.zeros(HOUGH_THETA, np.float32)
    core_lib().ekf_hough_tables(c.ctypes.data_as(C.POINTER(C.c_float)), s.ctypes.data_as(C.POINTER(C.c_float)))
    return c, s


def hough_lines_from_peaks(peaks, values, max_lines=HOUGH_PEAKS):
    peaks = np.ascontiguousarray(peaks, np.int32)
    values = np.ascontiguousarray(values, np.int32)
    lines = np.zeros((max_lines, 3))
    n = core_lib().ekf_hough_lines_from_peaks(_ip(peaks), _ip(values), _dp(lines), max_lines)
    return lines[:n].copy()


class HoughBatch:
    """HoughTransform::getLines for batches of laser scans (one ekf_hough handle)."""

    def __init__(self, max_scans, device=0):
        self.L = core_lib()
        self.h = C.c_void_p()
        rc = self.L.ekf_hough_create(C.byref(self.h), device, max_scans)
        if rc:
            msg = self.L.ekf_hough_last_error(None)
            self.h = None
            raise EkfError(rc, msg.decode() if msg else "")
        self.max_scans = max_scans

    def close(self):
        if getattr(self, "h", None):
            self.L.ekf_hough_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc:
            msg = self.L.ekf_hough_last_error(self.h)
            raise EkfError(rc, msg.decode() if msg else "")

    @staticmethod
    def _inputs(X, Y, R):
        X = np.ascontiguousarray(X, np.float64)
        Y = np.ascontiguousarray(Y, np.float64)
        R = np.ascontiguousarray(R, np.uint32)
        assert X.ndim == 2 and X.shape == Y.shape == R.shape
        return X, Y, R

    def get_lines(self, X, Y, R, max_lines=64, want_grid=False, want_peaks=True, split=True, out=None):
        """X, Y, R [n_scans][n_points]. -> dict(lines, n_lines, peaks, values, grid); lines is a list of
        per-scan (n, 3) arrays, or the raw [n_scans][max_lines][3] buffer when split=False.
        out = (lines, n_lines) reuses caller buffers (e.g. pinned ones from PinnedArray)."""
        X, Y, R = self._inputs(X, Y, R)
        S, P = X.shape
        lines = out[0] if out is not None else np.empty((S, max_lines, 3))
        n_lines = out[1] if out is not None else np.empty(S, np.int32)
        peaks = np.zeros((S, HOUGH_PEAKS), np.int32) if want_peaks else None
        values = np.zeros((S, HOUGH_PEAKS), np.int32) if want_peaks else None
        grid = np.zeros((S, HOUGH_THETA * HOUGH_RADIUS), np.uint8) if want_grid else None
        self._chk(self.L.ekf_hough_get_lines(self.h, S, P, _dp(X), _dp(Y), R.ctypes.data_as(C.POINTER(C.c_uint32)),
                                             _dp(lines), max_lines, _ip(n_lines), _ip(peaks), _ip(values),
                                             grid.ctypes.data_as(C.POINTER(C.c_uint8)) if want_grid else None))
        return {"lines": [lines[s, :min(n_lines[s], max_lines)].copy() for s in range(S)] if split else lines,
                "n_lines": n_lines, "peaks": peaks, "values": values, "grid": grid}

    def get_features(self, X, Y, R, cur_phi=None, offset=None, max_feats=32, want_compass=True, want_segments=False,
                     max_segs=64, max_lines=64, want_lines=True):
        """FeatureDetector::getFeatures for a batch: -> dict(feats [S][max_feats][2], n_feats, compass, offset,
        lines, n_lines, segments, n_segs). cur_phi / offset: per-scan filter heading and COMPASS_OFFSET (100 = unset)."""
        X, Y, R = self._inputs(X, Y, R)
        S, P = X.shape
        feats = np.zeros((S, max_feats, 2))
        n_feats = np.zeros(S, np.int32)
        compass = np.zeros(S) if want_compass else None
        phi = np.ascontiguousarray(cur_phi, np.float64) if cur_phi is not None else None
        off = np.array(offset, np.float64, copy=True) if offset is not None else (np.full(S, 100.0) if want_compass else None)
        lines = np.zeros((S, max_lines, 3)) if want_lines else None
        n_lines = np.zeros(S, np.int32) if want_lines else None
        segs = np.zeros((S, max_segs, 7)) if want_segments else None
        n_segs = np.zeros(S, np.int32)
        self._chk(self.L.ekf_hough_get_features(self.h, S, P, _dp(X), _dp(Y), R.ctypes.data_as(C.POINTER(C.c_uint32)),
                                                _dp(phi), _dp(off), _dp(feats), max_feats, _ip(n_feats), _dp(compass),
                                                _dp(lines), max_lines, _ip(n_lines), _dp(segs), max_segs, _ip(n_segs)))
        return {"feats": feats, "n_feats": n_feats, "compass": compass, "offset": off, "lines": lines, "n_lines": n_lines,
                "segments": segs, "n_segs": n_segs}

    def upload(self, X, Y, R):
        X, Y, R = self._inputs(X, Y, R)
        self._shape = X.shape
        self._chk(self.L.ekf_hough_upload(self.h, X.shape[0], X.shape[1], _dp(X), _dp(Y),
                                          R.ctypes.data_as(C.POINTER(C.c_uint32))))

    def run_resident(self, max_lines=64):
        self._chk(self.L.ekf_hough_run_resident(self.h, max_lines))

    def download(self, max_lines=64):
        S = self._shape[0]
        lines = np.zeros((S, max_lines, 3))
        n_lines = np.zeros(S, np.int32)
        self._chk(self.L.ekf_hough_download(self.h, _dp(lines), max_lines, _ip(n_lines), None, None))
        return lines, n_lines

    def sync(self):
        self._chk(self.L.ekf_hough_sync(self.h))

    def kernel_time(self):
        ms, n = C.c_float(), C.c_int()
        self._chk(self.L.ekf_hough_kernel_time(self.h, C.byref(ms), C.byref(n)))
        return ms.value, n.value


def device_count():
    n = 0
    while core_lib().ekf_device_info(n, None, None, None, None, None) == 0:
        n += 1
    return n


def device_info(device=0):
    L = core_lib()
    sm, maj, mnr = C.c_int(), C.c_int(), C.c_int()
    smem, mem = C.c_size_t(), C.c_size_t()
    rc = L.ekf_device_info(device, C.byref(sm), C.byref(maj), C.byref(mnr), C.byref(smem), C.byref(mem))
    if rc:
        raise EkfError(rc, "ekf_device_info failed (no CUDA device?)")
    return {"sm_count": sm.value, "cc": (maj.value, mnr.value), "smem_optin": smem.value, "total_mem": mem.value}


def debug_phase_cycles(read=True):
    """Per-phase cycle counters of the register-tile kernel (CTA 0); first call enables them."""
    if not read:
        core_lib().ekf_debug_phase_cycles(None)
        return None
    buf = (C.c_longlong * 16)()
    core_lib().ekf_debug_phase_cycles(buf)
    return list(buf) + [0] * 16


def debug_stile_timestamps():
    out = (C.c_longlong * 128)()
    core_lib().ekf_debug_stile_timestamps(out)
    return np.array(out[:], np.int64).reshape(8, 16)


def debug_dtile_timestamps():
    out = (C.c_longlong * 64)()
    core_lib().ekf_debug_dtile_timestamps(out)
    return np.array(out[:], np.int64).reshape(4, 16)


def measure_fp64_peak(device=0):
    v = C.c_double()
    rc = core_lib().ekf_measure_fp64_peak(device, C.byref(v))
    if rc:
        raise EkfError(rc, "ekf_measure_fp64_peak failed")
    return v.value
