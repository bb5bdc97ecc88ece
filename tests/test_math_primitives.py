"""The numeric spec (crowdnav_b200/csrc/cn_math.h), exercised through the oracle's taps on the CPU."""
import ctypes as C

import numpy as np

from oracle.oracle import lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_sincos_bin_accuracy():
    L = lib()
    rng = np.random.default_rng(0)
    a = rng.integers(0, 2 ** 32, 200000, dtype=np.uint64).astype(np.uint32)
    a[:8] = [0, 1, 0x20000000, 0x40000000, 0x7FFFFFFF, 0x80000000, 0xDFFFFFFF, 0xFFFFFFFF]
    s, c = np.zeros(len(a), np.float32), np.zeros(len(a), np.float32)
    L.orc_sincos_bin(_p(a), _p(s), _p(c), len(a))
    ang = a.astype(np.float64) * (2 * np.pi / 2 ** 32)
    assert np.abs(s - np.sin(ang)).max() < 1.5e-7 and np.abs(c - np.cos(ang)).max() < 1.5e-7


def test_sincos_rad_accuracy():
    L = lib()
    x = np.random.default_rng(1).uniform(-50, 50, 100000).astype(np.float32)
    s, c = np.zeros_like(x), np.zeros_like(x)
    L.orc_sincos_rad(_p(x), _p(s), _p(c), len(x))
    assert np.abs(s - np.sin(x.astype(np.float64))).max() < 4e-7 and np.abs(c - np.cos(x.astype(np.float64))).max() < 4e-7


def test_atan2_accuracy_and_quadrants():
    L = lib()
    rng = np.random.default_rng(2)
    y = rng.uniform(-3, 3, 200000).astype(np.float32)
    x = rng.uniform(-3, 3, 200000).astype(np.float32)
    y[:6] = [0, 0, 1, -1, 0, 1e-20]
    x[:6] = [1, -1, 0, 0, 0, 1e-20]
    o = np.zeros_like(x)
    L.orc_atan2(_p(y), _p(x), _p(o), len(x))
    want = np.arctan2(y.astype(np.float64), x.astype(np.float64))
    assert np.abs(o - want).max() < 4e-7     # 1 ulp near +-pi is 2.4e-7
    assert o[4] == 0.0                                   # atan2(0, 0) = 0 by definition here


def test_exp_accuracy():
    L = lib()
    x = np.random.default_rng(3).uniform(-8, 8, 100000).astype(np.float32)
    o = np.zeros_like(x)
    L.orc_exp(_p(x), _p(o), len(x))
    want = np.exp(x.astype(np.float64))
    assert (np.abs(o - want) / want).max() < 4e-7


def test_rounding_matches_numpy_and_python2_semantics():
    """np.around(x, 3) on the float64 image of x, cast to float32; Python-2 round = half away from zero."""
    L = lib()
    rng = np.random.default_rng(4)
    x = np.concatenate([rng.uniform(-3, 3, 300000), rng.uniform(0.08, 0.6, 300000)]).astype(np.float32)
    # exact decimal ties in float32 (representable k + 0.5 thousandths do not exist except .0005 * 2^n; use .125 family)
    x[:6] = [0.125, -0.125, 0.375, 2.5, 0.0625, -0.0625]
    np3, py3, py2 = np.zeros_like(x), np.zeros_like(x), np.zeros_like(x)
    L.orc_round(_p(x), _p(np3), _p(py3), _p(py2), len(x))
    x64 = x.astype(np.float64)
    assert np.array_equal(np3, np.around(x64, 3).astype(np.float32))
    def half_away(v, n):
        s = 10.0 ** n
        return (np.sign(v) * np.floor(np.abs(v) * s + 0.5) / s)
    # away from exact ties floor(|v|*s + 0.5) is the nearest integer; the product |v|*s is float64-exact enough
    assert np.array_equal(py3, half_away(x64, 3).astype(np.float32))
    assert np.array_equal(py2, half_away(x64, 2).astype(np.float32))
    assert py2[0] == np.float32(0.13) and py2[1] == np.float32(-0.13)      # Python-2 round(0.125, 2) = 0.13


def test_division_by_constants_is_ieee_exact():
    """cn_div1000 / cn_div100 (3-instruction Markstein form) equal IEEE division for EVERY integer |k| <= 2^24."""
    L = lib()
    L.orc_div_const_mismatches.restype = C.c_long
    assert L.orc_div_const_mismatches(-(1 << 24), 1 << 24) == 0


def test_philox2x32_known_answers():
    """Random123 known-answer vectors for philox2x32-10."""
    L = lib()
    out = (C.c_uint32 * 2)()
    for c0, c1, k, want in ((0, 0, 0, (0xff1dae59, 0x6cd10df2)),
                            (0xffffffff, 0xffffffff, 0xffffffff, (0x2c3f628b, 0xab4fd7ad)),
                            (0x243f6a88, 0x85a308d3, 0x13198a2e, (0xdd7ce038, 0xf62a4c12))):
        L.orc_philox2x32(C.c_uint32(c0), C.c_uint32(c1), C.c_uint32(k), out)
        assert (out[0], out[1]) == want


def test_16_bit_wire_format_is_lossless_on_the_reference_rows():
    """The fused gather's 16-bit wire format (cn_flat.cu wire16_encode / wire16_decode): every value of an observation
    row is a whole number of thousandths, so int16 thousandths + the correctly rounded quotient k / 1000 (cn_div1000,
    exhaustively equal to IEEE division above) give the fp32 bits back.  Checked here on (a) every k the format can
    carry, with the arithmetic the kernel uses, and (b) the rows the REFERENCE's own get_state returned in the golden
    traces (float32, as the agent sees them) and the oracle's rows next to them."""
    import glob
    import os
    k = np.arange(-32767, 32768, dtype=np.int32)
    v = (k.astype(np.float64) / 1000.0).astype(np.float32)                    # what a row holds: fl32(k / 1000)
    back = np.rint(v * np.float32(1000.0)).astype(np.int32)                   # wire16_encode: rintf(v * 1000.0f) in fp32
    assert np.array_equal(back, k)
    here = os.path.join(os.path.dirname(__file__), "golden")
    n_rows = 0
    for f in sorted(glob.glob(os.path.join(here, "trace_*.npz"))):
        z = np.load(f)
        for name in ("ref_state", "oracle_obs"):
            if name not in z.files:
                continue
            rows = z[name].astype(np.float32)
            if name == "ref_state" and "original" not in f:
                rows = rows[:, :rows.shape[1] - 4 * ((rows.shape[1] - 366) // 4)]   # K block: order / clock dependent, same format
            kk = np.rint(rows.astype(np.float64) * 1000.0)
            assert np.abs(kk).max() <= 32767
            rebuilt = (kk / 1000.0).astype(np.float32)
            zero = rows == 0
            assert np.array_equal(rebuilt[~zero].view(np.uint32), rows[~zero].view(np.uint32)), (f, name)
            n_rows += len(rows)
    assert n_rows > 5000
