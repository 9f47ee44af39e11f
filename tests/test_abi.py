"""CPU tests (not gpu): the C-ABI library loads, exports every symbol include/*.h declares, keeps
the reference's literals as config defaults, and fails loudly (no CPU fallback) without a GPU."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ekf_[a-z0-9_]+)\s*\(", txt)))


def test_core_exports_every_declared_symbol(ekf):
    lib = ekf.core_lib()
    names = declared("ekf_slam_b200.h")
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "libekf_slam_b200.so does not export " + n


def test_core_exports_the_hough_front_end(ekf):
    lib = ekf.core_lib()
    names = declared("ekf_hough_b200.h")
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), "libekf_slam_b200.so does not export " + n


def test_synth_exports_every_declared_symbol(ekf):
    lib = ekf.synth_lib()
    for n in declared("ekf_synth.h"):
        assert hasattr(lib, n), "libekf_synth.so does not export " + n


def test_default_config_is_the_reference_literals(ekf):
    cfg = ekf.Config()
    ekf.core_lib().ekf_default_config(C.byref(cfg))
    assert cfg.sigma_v == 0.01 and cfg.sigma_w == 0.04                # kalmanfilter.cpp:28-29
    assert cfg.deg2rad_pi == 3.141592654                              # kalmanfilter.cpp:19
    assert cfg.two_pi == 6.283185307                                  # kalmanfilter.cpp:99
    assert cfg.gamma_max == 50 and cfg.gamma_min == 10                # kalmanfilter.cpp:67-68
    assert cfg.cond_max == 80.0                                       # Update.cpp:131
    assert cfg.mahal_init == 999999999999.0                           # kalmanfilter.h:17


def test_only_sm_100a_code_is_embedded():
    so = os.path.join(ROOT, "2d-ekf-slam_b200", "lib", "libekf_slam_b200.so")
    out = subprocess.run(["cuobjdump", "-lelf", so], stdout=subprocess.PIPE, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_gpu_means_loud_failure(ekf):
    import shutil
    if shutil.which("nvidia-smi") and subprocess.run(["nvidia-smi", "-L"], stdout=subprocess.PIPE).returncode == 0:
        pytest.skip("a GPU is present")
    with pytest.raises(ekf.EkfError) as ei:
        ekf.FilterBatch(2, 10)
    assert ei.value.code == ekf.ERR_NO_DEVICE
    assert "no CPU fallback" in str(ei.value)
    with pytest.raises(ekf.EkfError) as ei:
        ekf.HoughBatch(4)
    assert ei.value.code == ekf.ERR_NO_DEVICE
    with pytest.raises(ekf.EkfError) as ei:
        ekf.ShardedMap([0, 0], 10)
    assert ei.value.code == ekf.ERR_NO_DEVICE


def test_product_never_touches_the_oracle():
    """The product path (package + include) must not reference oracle/ in any way."""
    for base in ("2d-ekf-slam_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for fn in files:
                if fn.endswith((".so", ".log", ".pyc")):
                    continue
                txt = open(os.path.join(dp, fn), errors="replace").read()
                assert "ekf_oracle" not in txt and "libekf_ref" not in txt and "oracle_lib" not in txt, os.path.join(dp, fn)


def test_dropin_header_compiles_against_the_reference_loop_shape():
    """host/kalmanfilter.h + examples/slam_synthetic.cpp (the slam.cpp:127-182 call sequence) build."""
    out = "/tmp/ekf_slam_synthetic_test"
    r = subprocess.run(["/usr/bin/g++", "-std=c++11", "-O1", "-I" + os.path.join(ROOT, "oracle", "shim"),
                        "-I" + os.path.join(ROOT, "2d-ekf-slam_b200", "host"), "-I" + os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "examples", "slam_synthetic.cpp"),
                        "-L" + os.path.join(ROOT, "2d-ekf-slam_b200", "lib"), "-lekf_slam_b200", "-lekf_synth",
                        "-Wl,-rpath," + os.path.join(ROOT, "2d-ekf-slam_b200", "lib"), "-o", out],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout


def test_pipeline_chunk_schedule(ekf):
    """Host logic of ekf_run(): the filters of a batch are cut into chunks whose sizes, in waves of co-resident
    CTAs, ramp up from one wave at both ends (short first copy in, short last copy out). No GPU needed."""
    fn = ekf.core_lib().ekf_debug_pipeline_chunks
    fn.argtypes = [C.c_longlong, C.c_longlong, C.POINTER(C.c_longlong), C.c_int]
    buf = (C.c_longlong * 64)()
    assert fn(0, 592, buf, 64) < 0 and fn(10, 592, buf, 8) < 0
    for F, wave in [(1, 592), (591, 592), (592, 592), (593, 592), (8192, 592), (65536, 592), (65536, 296),
                    (10 ** 6, 592), (10 ** 7, 148), (4096, 1184)]:
        n = fn(F, wave, buf, 64)
        assert 1 <= n <= 48
        b = [buf[i] for i in range(n + 1)]
        assert b[0] == 0 and b[-1] == F and all(x < y for x, y in zip(b, b[1:]))        # a partition of [0, F)
        sizes = [y - x for x, y in zip(b, b[1:])]
        assert all(sz % wave == 0 for sz in sizes[:-1])                                 # whole waves, except the last chunk
        if F // wave > 10000:                                                           # beyond any ramp of 48 chunks: equal chunks
            assert max(sizes) - min(sizes[:-1]) == 0
            continue
        assert sizes[0] <= wave and sizes[-1] <= wave                                   # one wave at both ends
        w = [-(-sz // wave) for sz in sizes]
        peak = w.index(max(w))
        assert all(x <= y for x, y in zip(w[:peak], w[1:peak + 1]))                     # ramps up ...
        last = len(w) - 1 - w[::-1].index(max(w))
        assert all(x >= y for x, y in zip(w[last:], w[last + 1:]))                      # ... and down again
    n = fn(8192, 592, buf, 64)
    assert [buf[i + 1] - buf[i] for i in range(n)] == [592, 592, 1184, 1184, 1184, 1184, 1184, 592, 496]
