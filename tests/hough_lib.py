"""ctypes access to the Hough checkers (TEST INFRASTRUCTURE): oracle/libhough_oracle.so (C restatement)
and oracle/_ref/libhough_ref.so (the reference's own houghtransform.cpp)."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
THETA, RADIUS, PEAKS = 180, 1601, 200
c_dp = C.POINTER(C.c_double)
c_fp = C.POINTER(C.c_float)
c_up = C.POINTER(C.c_uint)
c_ip = C.POINTER(C.c_int)
c_bp = C.POINTER(C.c_ubyte)


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


class HoughOracle:
    def __init__(self):
        self.lib = C.CDLL(os.path.join(ORACLE_DIR, "libhough_oracle.so"))
        L = self.lib
        L.hough_oracle_tables.argtypes = [c_fp, c_fp]
        L.hough_oracle_get_lines.argtypes = [C.c_int, c_dp, c_dp, c_up, c_fp, c_fp, c_dp, C.c_int, c_bp, c_ip]
        L.hough_oracle_lines_from_peaks.argtypes = [c_ip, c_ip, c_dp, C.c_int]
        L.hough_oracle_run_scans.argtypes = [C.c_int, C.c_int, c_dp, c_dp, c_up]
        L.hough_oracle_run_scans.restype = C.c_long
        self.cos, self.sin = self.tables()

    def tables(self):
        c = np.zeros(THETA, np.float32)
        s = np.zeros(THETA, np.float32)
        self.lib.hough_oracle_tables(_p(c, c_fp), _p(s, c_fp))
        return c, s

    def get_lines(self, x, y, rng, want_grid=False):
        """-> lines [n][3] (radius mm, theta rad, weight), peaks [200], grid (or None)"""
        x = np.ascontiguousarray(x, np.float64)
        y = np.ascontiguousarray(y, np.float64)
        rng = np.ascontiguousarray(rng, np.uint32)
        lines = np.zeros((PEAKS, 3))
        peaks = np.zeros(PEAKS, np.int32)
        grid = np.zeros(THETA * RADIUS, np.uint8) if want_grid else None
        n = self.lib.hough_oracle_get_lines(len(x), _p(x, c_dp), _p(y, c_dp), _p(rng, c_up), _p(self.cos, c_fp),
                                            _p(self.sin, c_fp), _p(lines, c_dp), PEAKS, _p(grid, c_bp), _p(peaks, c_ip))
        return lines[:n].copy(), peaks, grid

    def lines_from_peaks(self, peaks, values):
        peaks = np.ascontiguousarray(peaks, np.int32)
        values = np.ascontiguousarray(values, np.int32)
        lines = np.zeros((PEAKS, 3))
        n = self.lib.hough_oracle_lines_from_peaks(_p(peaks, c_ip), _p(values, c_ip), _p(lines, c_dp), PEAKS)
        return lines[:n].copy()

    def run_scans(self, X, Y, R):
        X = np.ascontiguousarray(X, np.float64)
        Y = np.ascontiguousarray(Y, np.float64)
        R = np.ascontiguousarray(R, np.uint32)
        return self.lib.hough_oracle_run_scans(X.shape[0], X.shape[1], _p(X, c_dp), _p(Y, c_dp), _p(R, c_up))

    def get_features(self, x, y, rng, cur_phi=0.0, offset=100.0):
        """The stages of FeatureDetector::getFeatures after getLines.
        -> features [n][2] (mm, robot frame), segments [m][7], compass, new offset, lines"""
        L = self.lib
        L.features_oracle_segments.argtypes = [C.c_int, c_dp, c_dp, c_up, c_dp, C.c_int, c_dp, C.c_int]
        L.features_oracle_corners.argtypes = [c_dp, C.c_int, c_dp, C.c_int]
        L.features_oracle_compass.argtypes = [c_dp, C.c_int, C.c_double, c_dp]
        L.features_oracle_compass.restype = C.c_double
        x = np.ascontiguousarray(x, np.float64)
        y = np.ascontiguousarray(y, np.float64)
        rng = np.ascontiguousarray(rng, np.uint32)
        lines, _, _ = self.get_lines(x, y, rng)
        lines = np.ascontiguousarray(lines)
        segs = np.zeros((512, 7))
        m = L.features_oracle_segments(len(x), _p(x, c_dp), _p(y, c_dp), _p(rng, c_up), _p(lines, c_dp), len(lines),
                                       _p(segs, c_dp), 512)
        segs = np.ascontiguousarray(segs[:m])
        feats = np.zeros((512, 2))
        n = L.features_oracle_corners(_p(segs, c_dp), m, _p(feats, c_dp), 512)
        off = C.c_double(offset)
        compass = L.features_oracle_compass(_p(lines, c_dp), len(lines), cur_phi, C.byref(off))
        return feats[:n].copy(), segs, compass, off.value, lines


class HoughRef:
    PATH = os.path.join(ORACLE_DIR, "_ref", "libhough_ref.so")

    @classmethod
    def available(cls):
        return os.path.exists(cls.PATH)

    def __init__(self):
        self.lib = C.CDLL(self.PATH)
        L = self.lib
        L.refh_tables.argtypes = [c_fp, c_fp]
        L.refh_get_lines.argtypes = [C.c_int, c_dp, c_dp, c_up, c_dp, C.c_int, c_bp, c_ip]
        L.refh_run_scans.argtypes = [C.c_int, C.c_int, c_dp, c_dp, c_up]
        L.refh_run_scans.restype = C.c_long

    def constants(self):
        return self.lib.refh_theta_size(), self.lib.refh_radius_size(), self.lib.refh_num_peaks()

    def tables(self):
        c = np.zeros(THETA, np.float32)
        s = np.zeros(THETA, np.float32)
        self.lib.refh_tables(_p(c, c_fp), _p(s, c_fp))
        return c, s

    def get_lines(self, x, y, rng, want_grid=False):
        x = np.ascontiguousarray(x, np.float64)
        y = np.ascontiguousarray(y, np.float64)
        rng = np.ascontiguousarray(rng, np.uint32)
        lines = np.zeros((PEAKS, 3))
        peaks = np.zeros(PEAKS, np.int32)
        grid = np.zeros(THETA * RADIUS, np.uint8) if want_grid else None
        n = self.lib.refh_get_lines(len(x), _p(x, c_dp), _p(y, c_dp), _p(rng, c_up), _p(lines, c_dp), PEAKS,
                                    _p(grid, c_bp), _p(peaks, c_ip))
        return lines[:n].copy(), peaks, grid

    def run_scans(self, X, Y, R):
        X = np.ascontiguousarray(X, np.float64)
        Y = np.ascontiguousarray(Y, np.float64)
        R = np.ascontiguousarray(R, np.uint32)
        return self.lib.refh_run_scans(X.shape[0], X.shape[1], _p(X, c_dp), _p(Y, c_dp), _p(R, c_up))

    def get_features(self, x, y, rng, cur_phi=0.0, offset=100.0):
        """getLines, fitLineSegments, extractCorners, getStructCompass of the reference itself."""
        L = self.lib
        L.reff_get_features.argtypes = [C.c_int, c_dp, c_dp, c_up, C.c_double, c_dp, c_dp, C.c_int, c_dp, C.c_int, c_ip, c_dp]
        x = np.ascontiguousarray(x, np.float64)
        y = np.ascontiguousarray(y, np.float64)
        rng = np.ascontiguousarray(rng, np.uint32)
        feats = np.zeros((512, 2))
        segs = np.zeros((512, 7))
        m = C.c_int()
        off = C.c_double(offset)
        compass = C.c_double()
        n = L.reff_get_features(len(x), _p(x, c_dp), _p(y, c_dp), _p(rng, c_up), cur_phi, C.byref(off), _p(feats, c_dp), 512,
                                _p(segs, c_dp), 512, C.byref(m), C.byref(compass))
        return feats[:n].copy(), segs[:m.value].copy(), compass.value, off.value
