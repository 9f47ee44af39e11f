"""ctypes bindings to the CPU checkers (TEST INFRASTRUCTURE; never imported by the product).

* ``Oracle``  - oracle/libekf_oracle.so, the plain-C restatement ("port").
* ``Ref``     - oracle/_ref/libekf_ref.so, the reference's own odometry/*.cpp compiled unmodified
                over oracle/shim ("reference"). Present wherever oracle/_ref was built
                (this container) or shipped as a prebuilt binary (GPU box).
Both expose the same calls so tests can be parametrised over them.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)


class Trace(C.Structure):
    _fields_ = [("decision", C.c_int32), ("opt_i", C.c_int32), ("mahal", C.c_double),
                ("n_cond_skipped", C.c_int32), ("k_col", C.c_int32),
                ("margin_gmin", C.c_double), ("margin_gmax", C.c_double), ("margin_cond", C.c_double)]


def _dp(a):
    return a.ctypes.data_as(c_dp) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(c_ip) if a is not None else None


def build_oracle():
    """Build the checkers (idempotent). oracle/_ref is only (re)built where /root/reference exists."""
    subprocess.run(["make", "-C", ORACLE_DIR, "--no-print-directory"], check=True,
                   stdout=subprocess.DEVNULL)


def record_len(max_meas):
    return 8 + 6 * max_meas


class Oracle:
    """Plain-C restatement; state lives in numpy arrays owned by the caller-side wrapper."""
    kind = "port"

    def __init__(self):
        path = os.path.join(ORACLE_DIR, "libekf_oracle.so")
        if not os.path.exists(path):
            build_oracle()
        self.lib = C.CDLL(path)
        L = self.lib
        L.ekf_oracle_propagate.argtypes = [C.c_int, c_dp, c_dp, C.c_int, C.c_double, C.c_double, C.c_double]
        L.ekf_oracle_update.argtypes = [C.c_int, c_dp, c_dp, C.c_int, C.c_int, c_dp, c_dp, C.c_int, C.c_int,
                                        C.POINTER(Trace)]
        L.ekf_oracle_update.restype = C.c_int
        L.ekf_oracle_update_chunk.argtypes = [C.c_int, c_dp, c_dp, C.c_int, C.c_int, C.c_int, c_dp, c_dp, C.c_int,
                                              C.c_int, C.POINTER(Trace)]
        L.ekf_oracle_update_chunk.restype = C.c_int
        L.ekf_oracle_update_compass.argtypes = [C.c_int, c_dp, c_dp, C.c_int, C.c_double, C.c_double]
        L.ekf_oracle_measurement_from_feature.argtypes = [C.c_double, C.c_double, c_dp, c_dp]
        L.ekf_oracle_run_batch.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, c_dp, C.c_int, c_ip, c_ip, c_dp,
                                           c_dp, c_dp, c_ip, c_dp, c_dp, C.c_int, C.c_int]
        L.ekf_oracle_run_batch.restype = C.c_double

    def new_filter(self, cap_lm):
        return OracleFilter(self, cap_lm)

    def measurement_from_feature(self, fx_mm, fy_mm):
        z = np.zeros(2)
        R = np.zeros(4)
        self.lib.ekf_oracle_measurement_from_feature(fx_mm, fy_mm, _dp(z), _dp(R))
        return z, R

    def run_batch(self, records, max_meas, cap_lm, n_threads=1, trace=True, pose_trace=False, final_state=False,
                  warm_steps=0):
        F, T, L = records.shape
        assert L == record_len(max_meas) and records.dtype == np.float64 and records.flags.c_contiguous
        out = _alloc_batch_outputs(F, T, max_meas, 3 + 2 * cap_lm, trace, pose_trace, final_state)
        secs = self.lib.ekf_oracle_run_batch(F, T, max_meas, cap_lm, _dp(records), n_threads,
                                             _ip(out["decision"]), _ip(out["index"]), _dp(out["mahal"]),
                                             _dp(out["pose_trace"]), _dp(out["final_pose"]), _ip(out["final_nlm"]),
                                             _dp(out["final_x"]), _dp(out["final_P"]), out["ld"], warm_steps)
        out["seconds"] = abs(secs)
        out["bad"] = secs < 0
        return out


def _alloc_batch_outputs(F, T, M, ld, trace, pose_trace, final_state):
    return {
        "decision": np.full((F, T, M), -7, np.int32) if trace else None,
        "index": np.full((F, T, M), -7, np.int32) if trace else None,
        "mahal": np.zeros((F, T, M)) if trace else None,
        "pose_trace": np.zeros((F, T, 3)) if pose_trace else None,
        "final_pose": np.zeros((F, 3)),
        "final_nlm": np.zeros(F, np.int32),
        "final_x": np.zeros((F, ld)) if final_state else None,
        "final_P": np.zeros((F, ld, ld)) if final_state else None,   # [f][col][row]: column-major slabs
        "ld": ld,
    }


class OracleFilter:
    """One filter driven through the restatement; x/P in fixed-capacity column-major buffers."""

    def __init__(self, oracle, cap_lm):
        self.o = oracle
        self.cap_n = 3 + 2 * cap_lm
        self.n = 3
        self.x = np.zeros(self.cap_n)
        self.Pbuf = np.zeros(self.cap_n * self.cap_n)

    def set_state(self, x, P):
        n = len(x)
        self.n = n
        self.x[:] = 0
        self.x[:n] = x
        self.Pbuf[:] = 0
        Pv = self.Pbuf.reshape(self.cap_n, self.cap_n)  # [col][row]
        Pv[:n, :n] = np.asarray(P).T
        return self

    def get_state(self):
        n = self.n
        Pv = self.Pbuf.reshape(self.cap_n, self.cap_n)
        return self.x[:n].copy(), Pv[:n, :n].T.copy()

    @property
    def num_landmarks(self):
        return (self.n - 3) // 2

    def pose(self):
        return self.x[:3].copy()

    def propagate(self, vel_mm_s, rotvel_deg_s, dt):
        self.o.lib.ekf_oracle_propagate(self.n, _dp(self.x), _dp(self.Pbuf), self.cap_n, vel_mm_s, rotvel_deg_s, dt)

    def update(self, z, R, gamma_max=50, gamma_min=10):
        z = np.ascontiguousarray(z, np.float64)
        R = np.ascontiguousarray(R, np.float64)
        tr = Trace()
        n2 = self.o.lib.ekf_oracle_update(self.n, _dp(self.x), _dp(self.Pbuf), self.cap_n, self.cap_n, _dp(z),
                                          _dp(R), gamma_max, gamma_min, C.byref(tr))
        if n2 < 0:
            raise OverflowError("oracle filter capacity exceeded")
        self.n = n2
        return tr

    def update_chunk(self, zs, Rs, gamma_max=50, gamma_min=10):
        """One doUpdate call with n_z measurements: zs [n_z][2], Rs [n_z][4] (column-major 2x2 each).
        The gating bound is frozen at call entry (Update.cpp:26). Returns the list of traces."""
        zs = np.ascontiguousarray(zs, np.float64).reshape(-1, 2)
        Rs = np.ascontiguousarray(Rs, np.float64).reshape(-1, 4)
        trs = (Trace * len(zs))()
        n2 = self.o.lib.ekf_oracle_update_chunk(self.n, _dp(self.x), _dp(self.Pbuf), self.cap_n, self.cap_n, len(zs),
                                                _dp(zs), _dp(Rs), gamma_max, gamma_min, trs)
        if n2 < 0:
            raise OverflowError("oracle filter capacity exceeded")
        self.n = n2
        return list(trs)

    def update_compass(self, z, R):
        self.o.lib.ekf_oracle_update_compass(self.n, _dp(self.x), _dp(self.Pbuf), self.cap_n, z, R)


class Ref:
    """The reference's own KalmanFilter (compiled unmodified over the stand-in Eigen)."""
    kind = "reference"

    @staticmethod
    def path(check=False):
        return os.path.join(ORACLE_DIR, "_ref", "libekf_ref_check.so" if check else "libekf_ref.so")

    @classmethod
    def available(cls):
        if not os.path.exists(cls.path()) and os.path.isdir("/root/reference/odometry"):
            build_oracle()
        return os.path.exists(cls.path())

    def __init__(self, check=False):
        if not self.available():
            raise FileNotFoundError(self.path())
        self.lib = C.CDLL(self.path(check))
        L = self.lib
        L.ref_create.restype = C.c_void_p
        L.ref_destroy.argtypes = [C.c_void_p]
        L.ref_dim.argtypes = [C.c_void_p]
        L.ref_get_state.argtypes = [C.c_void_p, c_dp, c_dp]
        L.ref_set_state.argtypes = [C.c_void_p, C.c_int, c_dp, c_dp]
        L.ref_get_pose.argtypes = [C.c_void_p, c_dp, C.POINTER(C.c_int)]
        L.ref_propagate.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]
        L.ref_update_compass.argtypes = [C.c_void_p, C.c_double, C.c_double]
        L.ref_update.argtypes = [C.c_void_p, c_dp, c_dp, C.POINTER(Trace)]
        L.ref_update.restype = C.c_int
        L.ref_call_propagate.argtypes = [C.c_int, c_dp, c_dp, C.c_double, C.c_double, c_dp, C.c_double]
        L.ref_call_update.argtypes = [C.c_int, c_dp, c_dp, C.c_int, c_dp, c_dp, C.c_int, C.c_int]
        L.ref_call_update.restype = C.c_int
        L.ref_measurement_from_feature.argtypes = [C.c_double, C.c_double, c_dp, c_dp]
        L.ref_run_batch.argtypes = [C.c_int, C.c_int, C.c_int, c_dp, C.c_int, c_ip, c_ip, c_dp, c_dp, c_dp, c_ip,
                                    c_dp, c_dp, C.c_int, C.c_int]
        L.ref_run_batch.restype = C.c_double
        L.ref_hardware_threads.restype = C.c_int

    def new_filter(self, cap_lm=None):
        return RefFilter(self)

    def measurement_from_feature(self, fx_mm, fy_mm):
        z = np.zeros(2)
        R = np.zeros(4)
        self.lib.ref_measurement_from_feature(fx_mm, fy_mm, _dp(z), _dp(R))
        return z, R

    def run_logged(self, records, max_meas, directory, scans=None):
        """The reference's odomRun / featuresRun / covRun / knownfeaturesRun text files for one filter;
        with scans = (x [T][B], y [T][B], range [T][B]) also scanRun.txt (slam.cpp:184-203)."""
        records = np.ascontiguousarray(records, np.float64)
        T, L = records.shape
        assert L == record_len(max_meas)
        if scans is None:
            self.lib.ref_run_logged.argtypes = [C.c_int, C.c_int, c_dp, C.c_char_p]
            rc = self.lib.ref_run_logged(T, max_meas, _dp(records), str(directory).encode())
        else:
            sx = np.ascontiguousarray(scans[0], np.float64)
            sy = np.ascontiguousarray(scans[1], np.float64)
            sr = np.ascontiguousarray(scans[2], np.uint32)
            assert sx.shape == sy.shape == sr.shape and sx.shape[0] == T
            self.lib.ref_run_logged_scans.argtypes = [C.c_int, C.c_int, c_dp, C.c_char_p, c_dp, c_dp,
                                                      C.POINTER(C.c_uint32), C.c_int]
            rc = self.lib.ref_run_logged_scans(T, max_meas, _dp(records), str(directory).encode(), _dp(sx), _dp(sy),
                                               sr.ctypes.data_as(C.POINTER(C.c_uint32)), sx.shape[1])
        assert rc == 0

    def call_update(self, x, P, z_chunk, R_chunk, gamma_max=50, gamma_min=10):
        """Pure-function call of the private KalmanFilter::Update (n_z >= 1). z_chunk 2 x n_z,
        R_chunk 2 x 2n_z (numpy, any layout). Returns (x', P')."""
        n = len(x)
        z_chunk = np.asarray(z_chunk, np.float64).reshape(2, -1)
        n_z = z_chunk.shape[1]
        R_chunk = np.asarray(R_chunk, np.float64).reshape(2, 2 * n_z)
        m = n + 2 * n_z
        xb = np.zeros(m)
        xb[:n] = x
        Pb = np.zeros(m * m)
        Pb[:n * n] = np.asarray(P, np.float64).T.reshape(-1)   # column-major n x n
        zc = np.ascontiguousarray(z_chunk.T.reshape(-1))       # column-major 2 x n_z
        Rc = np.ascontiguousarray(R_chunk.T.reshape(-1))
        n2 = self.lib.ref_call_update(n, _dp(xb), _dp(Pb), n_z, _dp(zc), _dp(Rc), gamma_max, gamma_min)
        return xb[:n2].copy(), Pb[:n2 * n2].reshape(n2, n2).T.copy()

    def run_batch(self, records, max_meas, cap_lm, n_threads=1, trace=True, pose_trace=False, final_state=False,
                  warm_steps=0):
        F, T, L = records.shape
        assert L == record_len(max_meas) and records.dtype == np.float64 and records.flags.c_contiguous
        out = _alloc_batch_outputs(F, T, max_meas, 3 + 2 * cap_lm, trace, pose_trace, final_state)
        secs = self.lib.ref_run_batch(F, T, max_meas, _dp(records), n_threads,
                                      _ip(out["decision"]), _ip(out["index"]), _dp(out["mahal"]),
                                      _dp(out["pose_trace"]), _dp(out["final_pose"]), _ip(out["final_nlm"]),
                                      _dp(out["final_x"]), _dp(out["final_P"]), out["ld"], warm_steps)
        out["seconds"] = abs(secs)
        out["bad"] = secs < 0
        return out

    def hardware_threads(self):
        return self.lib.ref_hardware_threads()


class RefFilter:
    def __init__(self, ref):
        self.r = ref
        self.h = C.c_void_p(ref.lib.ref_create())

    def __del__(self):
        if getattr(self, "h", None):
            self.r.lib.ref_destroy(self.h)
            self.h = None

    @property
    def n(self):
        return self.r.lib.ref_dim(self.h)

    @property
    def num_landmarks(self):
        return (self.n - 3) // 2

    def set_state(self, x, P):
        x = np.ascontiguousarray(x, np.float64)
        Pc = np.ascontiguousarray(np.asarray(P, np.float64).T)
        self.r.lib.ref_set_state(self.h, len(x), _dp(x), _dp(Pc))
        return self

    def get_state(self):
        n = self.n
        x = np.zeros(n)
        P = np.zeros((n, n))
        self.r.lib.ref_get_state(self.h, _dp(x), _dp(P))
        return x, P.T.copy()

    def pose(self):
        p = np.zeros(3)
        self.r.lib.ref_get_pose(self.h, _dp(p), None)
        return p

    def propagate(self, vel_mm_s, rotvel_deg_s, dt):
        self.r.lib.ref_propagate(self.h, vel_mm_s, rotvel_deg_s, dt)

    def update(self, z, R, gamma_max=50, gamma_min=10):
        assert (gamma_max, gamma_min) == (50, 10), "the reference hard-codes 50/10 (kalmanfilter.cpp:67-68)"
        z = np.ascontiguousarray(z, np.float64)
        R = np.ascontiguousarray(R, np.float64)
        tr = Trace()
        rc = self.r.lib.ref_update(self.h, _dp(z), _dp(R), C.byref(tr))
        if rc:
            raise RuntimeError("reference harness cross-check failed: rc=%d" % rc)
        return tr

    def update_compass(self, z, R):
        self.r.lib.ref_update_compass(self.h, z, R)
