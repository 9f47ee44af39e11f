"""Generate tests/golden/hough/*.npz from the REFERENCE's own Hough front-end
(oracle/_ref/libhough_ref.so: features/houghtransform.cpp compiled unmodified over oracle/shim).
Run in the build container, where /root/reference exists:

    python tests/golden/make_golden_hough.py

The file stores the exact scan inputs and what HoughTransform::getLines produced for them: the
constructor's cos / sin tables, the accumulator (as the sparse list of its non-zero cells), the
order-dependent peak array and the lines. tests/test_hough_golden.py checks the C restatement
against it on the CPU, tests/test_gpu_hough.py the CUDA path on the GPU.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import scan_synth  # noqa: E402
from hough_lib import HoughRef  # noqa: E402


def main():
    ref = HoughRef()
    X, Y, R = scan_synth.make_scans(6, seed=2024, n_boxes=4)
    ang = np.deg2rad(np.arange(181) - 90.0)            # plus two hand-made scans: one wall, nothing in range
    d = 2500.0 / np.maximum(np.cos(ang), 1e-3)
    r = np.where(d < 8000, np.rint(d), 8191).astype(np.uint32)
    X = np.vstack([X, r * np.cos(ang), 8191 * np.cos(ang)])
    Y = np.vstack([Y, r * np.sin(ang), 8191 * np.sin(ang)])
    R = np.vstack([R, r, np.full(181, 8191, np.uint32)])
    c, s = ref.tables()
    peaks, nz_cell, nz_val, nz_off, lines, n_lines = [], [], [], [0], [], []
    # the stages behind getLines (fitLineSegments, extractCorners, getStructCompass), offset carried along
    phi = np.linspace(-5.0, 5.0, len(X))
    feats, n_feats, segs, n_segs, compass, off_in, off_out = [], [], [], [], [], [], []
    off = 100.0
    for k in range(len(X)):
        f, sg, cp, new_off = ref.get_features(X[k], Y[k], R[k], float(phi[k]), off)
        fp = np.zeros((64, 2)); fp[:len(f)] = f
        sp = np.zeros((64, 7)); sp[:len(sg)] = sg
        feats.append(fp); n_feats.append(len(f)); segs.append(sp); n_segs.append(len(sg))
        compass.append(cp); off_in.append(off); off_out.append(new_off)
        off = new_off
    for k in range(len(X)):
        ln, pk, grid = ref.get_lines(X[k], Y[k], R[k], want_grid=True)
        cells = np.flatnonzero(grid).astype(np.int32)
        nz_cell.append(cells)
        nz_val.append(grid[cells])
        nz_off.append(nz_off[-1] + len(cells))
        peaks.append(pk)
        pad = np.zeros((200, 3))
        pad[:len(ln)] = ln
        lines.append(pad)
        n_lines.append(len(ln))
    out = os.path.join(HERE, "hough", "ref_8scans.npz")
    np.savez_compressed(out, x=X, y=Y, range=R, cos=c, sin=s, peaks=np.array(peaks, np.int32),
                        nz_cell=np.concatenate(nz_cell), nz_val=np.concatenate(nz_val), nz_off=np.array(nz_off, np.int64),
                        lines=np.array(lines), n_lines=np.array(n_lines, np.int32),
                        phi=phi, feats=np.array(feats), n_feats=np.array(n_feats, np.int32), segs=np.array(segs),
                        n_segs=np.array(n_segs, np.int32), compass=np.array(compass), off_in=np.array(off_in),
                        off_out=np.array(off_out))
    print(out, os.path.getsize(out), "bytes;", "lines per scan:", n_lines, "features per scan:", n_feats)


if __name__ == "__main__":
    main()
