"""DETMATH (oracle/detmath.h) against glibc libm: the binary32 results must be the correctly
rounded values except for a vanishing fraction of inputs, and never off by more than 1 ulp."""
import numpy as np
import pytest

from oracle import pyoracle as po


def _ulp_diff(a, b):
    a = np.asarray(a, np.float32)
    b = np.asarray(b, np.float32)
    ia = a.view(np.int32).astype(np.int64)
    ib = b.view(np.int32).astype(np.int64)
    ia = np.where(ia < 0, -(ia & 0x7FFFFFFF), ia)
    ib = np.where(ib < 0, -(ib & 0x7FFFFFFF), ib)
    return np.abs(ia - ib)


def _both(fn, *cols):
    L = po.lib()
    f = getattr(L, "oracle_" + fn)
    out = []
    for mode in (po.MATH_DET, po.MATH_LIBM):
        po.set_modes(mode, po.CONTRACT_FMA)
        out.append(np.array([f(*[float(c[i]) for c in cols]) for i in range(len(cols[0]))], np.float32))
    po.set_modes(po.MATH_DET, po.CONTRACT_FMA)
    return out


RNG = np.random.default_rng(12345)
N = 20000

CASES = {
    "cosf": (RNG.uniform(-7000, 7000, N).astype(np.float32),),
    "sinf": (RNG.uniform(-7000, 7000, N).astype(np.float32),),
    "expf": (RNG.uniform(-100, 80, N).astype(np.float32),),
    "logf": ((10.0 ** RNG.uniform(-38, 38, N)).astype(np.float32),),
    "tanhf": (np.concatenate([RNG.uniform(0, 12, N // 2), 10.0 ** RNG.uniform(-6, 5, N // 2)]).astype(np.float32),),
    "powf": ((10.0 ** RNG.uniform(-6, 3, N)).astype(np.float32), RNG.uniform(-6, 6, N).astype(np.float32)),
    "atan2f": (RNG.standard_normal(N).astype(np.float32), RNG.standard_normal(N).astype(np.float32)),
}


@pytest.mark.parametrize("fn", sorted(CASES))
def test_detmath_vs_libm(fn):
    det, ref = _both(fn, *CASES[fn])
    finite = np.isfinite(ref)
    assert np.array_equal(np.isfinite(det), finite)
    d = _ulp_diff(det[finite], ref[finite])
    assert d.max() <= 1, (fn, d.max())
    assert (d != 0).mean() <= 1e-3, (fn, (d != 0).mean())


def test_detmath_special_values():
    L = po.lib()
    po.set_modes(po.MATH_DET, po.CONTRACT_FMA)
    assert L.oracle_expf(-1e30) == 0.0 and L.oracle_expf(0.0) == 1.0 and np.isinf(L.oracle_expf(100.0))
    assert L.oracle_logf(0.0) == -np.inf and L.oracle_logf(1.0) == 0.0
    assert L.oracle_powf(0.0, 2.5) == 0.0 and L.oracle_powf(3.3, 0.0) == 1.0
    assert L.oracle_tanhf(6.0e4) == 1.0 and L.oracle_tanhf(0.0) == 0.0 and L.oracle_tanhf(9.1) == 1.0
    assert L.oracle_atan2f(0.0, 0.0) == 0.0
    assert L.oracle_cosf(0.0) == 1.0 and L.oracle_sinf(0.0) == 0.0
    # quarter-turn twiddle of fft_butterfly.glsl:27
    half_pi32 = float(np.float32(np.float32(np.pi) / np.float32(2.0)))
    assert np.float32(L.oracle_cosf(half_pi32)) == np.float32(-4.371139e-08)
    assert L.oracle_sinf(half_pi32) == 1.0
    # denormal result survives (no flush to zero)
    assert 0.0 < L.oracle_expf(-100.0) < 1e-40
