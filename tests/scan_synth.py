"""Synthetic LMS-200 scans for the Hough front-end (test + bench infrastructure): a robot at a random
pose inside a rectangular room with a few box obstacles, 181 beams over the front 180 degrees in
1-degree steps (slam.cpp:90), ranges in integer millimetres with Gaussian noise, returns given in
the robot frame as ArSensorReading::getLocalX/Y would (x forward, y left). Beams longer than the
sensor limit are reported at 8,191 mm... i.e. beyond HoughTransform::MAX_DIST and therefore skipped
by the transform (houghtransform.cpp:245)."""
import numpy as np

N_BEAMS = 181


def _ray_segments(px, py, dx, dy, segs):
    """Distance along unit rays (dx, dy) from (px, py) to the nearest of the segments [(x0,y0,x1,y1)]."""
    best = np.full(dx.shape, np.inf)
    for (x0, y0, x1, y1) in segs:
        ex, ey = x1 - x0, y1 - y0
        den = dx * ey - dy * ex
        with np.errstate(divide="ignore", invalid="ignore"):
            t = ((x0 - px) * ey - (y0 - py) * ex) / den          # along the ray
            u = ((x0 - px) * dy - (y0 - py) * dx) / den          # along the segment
        ok = (np.abs(den) > 1e-12) & (t > 1e-6) & (u >= 0.0) & (u <= 1.0)
        best = np.where(ok & (t < best), t, best)
    return best


def _box(cx, cy, w, h):
    x0, x1, y0, y1 = cx - w / 2, cx + w / 2, cy - h / 2, cy + h / 2
    return [(x0, y0, x1, y0), (x1, y0, x1, y1), (x1, y1, x0, y1), (x0, y1, x0, y0)]


def make_scans(n_scans, seed=0, noise_mm=8.0, room=(9000.0, 7000.0), n_boxes=3, max_range=8191):
    """-> x [n_scans][181], y [n_scans][181] (float64, mm, robot frame), range [n_scans][181] (uint32, mm)."""
    rng = np.random.default_rng(seed)
    ang = np.deg2rad(np.arange(N_BEAMS) - 90.0)
    X = np.zeros((n_scans, N_BEAMS))
    Y = np.zeros((n_scans, N_BEAMS))
    R = np.zeros((n_scans, N_BEAMS), np.uint32)
    W, H = room
    for s in range(n_scans):
        segs = _box(W / 2, H / 2, W, H)
        for _ in range(n_boxes):
            segs += _box(rng.uniform(0.15 * W, 0.85 * W), rng.uniform(0.15 * H, 0.85 * H),
                         rng.uniform(400, 1500), rng.uniform(400, 1500))
        while True:
            px, py = rng.uniform(0.1 * W, 0.9 * W), rng.uniform(0.1 * H, 0.9 * H)
            if all(not (min(a[0], a[2]) - 150 <= px <= max(a[0], a[2]) + 150 and
                        min(b[1], b[3]) - 150 <= py <= max(b[1], b[3]) + 150)
                   for a, b in zip(segs[4::4], segs[6::4])):
                break
        phi = rng.uniform(-np.pi, np.pi)
        d = _ray_segments(px, py, np.cos(phi + ang), np.sin(phi + ang), segs)
        d = d + rng.normal(0.0, noise_mm, N_BEAMS)
        r = np.where(np.isfinite(d), np.clip(np.rint(d), 1, max_range), max_range).astype(np.uint32)
        R[s] = r
        X[s] = r * np.cos(ang)
        Y[s] = r * np.sin(ang)
    return X, Y, R


def make_room(seed=0, room=(9000.0, 7000.0), n_boxes=3):
    """A fixed room: the outer rectangle plus a few boxes, as a list of wall segments (mm)."""
    rng = np.random.default_rng(seed)
    W, H = room
    segs = _box(W / 2, H / 2, W, H)
    for _ in range(n_boxes):
        segs += _box(rng.uniform(0.2 * W, 0.8 * W), rng.uniform(0.2 * H, 0.8 * H), rng.uniform(500, 1200), rng.uniform(500, 1200))
    return segs


def scan_from_pose(segs, px, py, phi, rng, noise_mm=8.0, max_range=8191):
    """One LMS-200 scan taken at pose (px, py [mm], phi [rad]) in the room `segs`."""
    ang = np.deg2rad(np.arange(N_BEAMS) - 90.0)
    d = _ray_segments(px, py, np.cos(phi + ang), np.sin(phi + ang), segs)
    d = d + rng.normal(0.0, noise_mm, N_BEAMS)
    r = np.where(np.isfinite(d), np.clip(np.rint(d), 1, max_range), max_range).astype(np.uint32)
    return r * np.cos(ang), r * np.sin(ang), r
