"""Pins the oracle to the reference itself.

oracle/_ref/libocean_ref.so is the reference's OWN six compute shaders (assets/shaders/compute/*.glsl, text unmodified
apart from the lexical plumbing listed in oracle/ref/glsl2cpp.py) compiled for the CPU and driven like
assets/water/wave_generator.gd drives them (oracle/pyref.py).  These tests assert that oracle/ocean_oracle.c -- the C
restatement every GPU parity test compares the CUDA path against -- reproduces the shaders' outputs BIT FOR BIT: the
butterfly table, the spectrum texture, both halves of the FFT buffer and both RGBA16F maps, for the BASELINE configs that
a CPU finishes in seconds, the parameter corners, the update/_process interleaving and every numeric-policy mode.
"""
import numpy as np
import pytest

from conftest import EDGE_CASES, demo_params
from oracle import pyoracle as po
from oracle import pyref as pr

pytestmark = pytest.mark.skipif(not pr.available(), reason="oracle/_ref is neither built nor buildable (no /root/reference)")

MODES = [("detmath_fma", po.MATH_DET, po.CONTRACT_FMA), ("detmath_strict", po.MATH_DET, po.CONTRACT_STRICT),
         ("libm_strict", po.MATH_LIBM, po.CONTRACT_STRICT), ("libm_fma", po.MATH_LIBM, po.CONTRACT_FMA)]


@pytest.fixture(autouse=True)
def _restore_modes():
    yield
    po.set_modes(po.MATH_DET, po.CONTRACT_FMA)
    pr.set_modes(po.MATH_DET, po.CONTRACT_FMA)


def _bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint8)


def _assert_same_state(o, r, what):
    assert np.array_equal(_bits(o.butterfly), _bits(r.butterfly)), f"{what}: butterfly table"
    assert np.array_equal(_bits(o.spectrum), _bits(r.spectrum)), f"{what}: spectrum texture"
    assert np.array_equal(_bits(o.fft_buffer), _bits(r.fft_buffer)), f"{what}: fft_buffer (both halves)"
    assert np.array_equal(o.displacement_map, r.displacement_map), f"{what}: displacement map"
    assert np.array_equal(o.normal_map, r.normal_map), f"{what}: normal/foam map"


def _pair(N, C, **over):
    return (po.OracleWaveGenerator(N), pr.RefWaveGenerator(N),
            [demo_params(po.CascadeParams, c, **over) for c in range(C)], [demo_params(po.CascadeParams, c, **over) for c in range(C)])


def test_reference_shaders_compiled():
    L = pr.lib()
    for s in pr.SHADERS:
        assert L.ref_has_shader(s.encode())
    import ctypes as C
    xyz = (C.c_int * 3)()
    assert L.ref_local_size(b"fft_compute", xyz) == 0 and tuple(xyz) == (1024, 1, 1)       # fft_compute.glsl:12
    assert L.ref_local_size(b"fft_unpack", xyz) == 0 and tuple(xyz) == (16, 16, 2)         # fft_unpack.glsl:11


@pytest.mark.parametrize("mode", MODES, ids=[m[0] for m in MODES])
@pytest.mark.parametrize("N,C,frames", [(128, 1, 1), (256, 4, 2)], ids=["cfg1_128x1", "cfg2_256x4"])
def test_oracle_reproduces_reference_shaders(N, C, frames, mode):
    """BASELINE.json configs[0] and configs[1]: every resource bit-identical after each update."""
    _, math_mode, contract = mode
    po.set_modes(math_mode, contract)
    pr.set_modes(math_mode, contract)
    o, r, po_p, pr_p = _pair(N, C)
    for f in range(frames):
        o.update_all(1.0 / 50.0, po_p)
        r.update_all(1.0 / 50.0, pr_p)
        _assert_same_state(o, r, f"{mode[0]} frame {f}")
    assert [p.time for p in po_p] == [p.time for p in pr_p]


@pytest.mark.parametrize("name", sorted(EDGE_CASES))
def test_oracle_reproduces_reference_shaders_on_parameter_corners(name):
    N = 128
    for contract in (po.CONTRACT_FMA, po.CONTRACT_STRICT):
        po.set_modes(po.MATH_DET, contract)
        pr.set_modes(po.MATH_DET, contract)
        o, r, po_p, pr_p = _pair(N, 2, **EDGE_CASES[name])
        for delta in (0.02, 0.0, 0.031):
            o.update_all(delta, po_p)
            r.update_all(delta, pr_p)
        _assert_same_state(o, r, f"{name} contract={contract}")


def test_foam_recurrence_and_scheduling_against_reference_shaders():
    """update()/_process() interleaving (wave_generator.gd:56-63,90-109) and the foam state carried through RGBA16F
    (fft_unpack.glsl:59-67) over 10 frames of the three demo cascades."""
    N, C = 128, 3
    o, r, po_p, pr_p = _pair(N, C)
    rng = np.random.default_rng(11)
    for f in range(10):
        delta = 1.0 / 50.0 + float(rng.uniform(0, 0.004))
        o.update(delta, po_p)
        r.update(delta, pr_p)
        for _ in range(int(rng.integers(0, C + 1))):
            o.process()
            r.process()
        if f == 4:                                   # a parameter change mid-run regenerates one spectrum
            for p in (po_p[1], pr_p[1]):
                p.wind_speed = 12.5
                p.should_generate_spectrum = True
    o.update(0.02, po_p)
    r.update(0.02, pr_p)
    while o.pass_num_cascades_remaining:
        o.process()
        r.process()
    _assert_same_state(o, r, "foam loop")
    assert o.normal_half()[0][..., 3].max() > 0


def test_half_conversion_of_the_oracle_equals_the_compilers():
    """RGBA16F stores: the oracle's hand-written RTNE float->half against _Float16 (what the reference build uses)."""
    L = po.lib()
    rng = np.random.default_rng(5)
    vals = np.concatenate([rng.standard_normal(20000).astype(np.float32) * np.float32(10.0) ** rng.integers(-9, 6, 20000).astype(np.float32),
                           np.array([0.0, -0.0, 65504.0, 65519.9, 65520.0, 1e-8, 5.96e-8, 2.98e-8, 2.9802325e-8, 6.1e-5, np.inf, -np.inf], np.float32)])
    got = np.array([L.oracle_float_to_half(float(v)) for v in vals], np.uint16)
    with np.errstate(over="ignore"):
        ref = vals.astype(np.float16).view(np.uint16)
    assert np.array_equal(got, ref)


def test_oracle_reproduces_reference_shaders_on_random_parameters():
    """hypothesis: random draws over the whole @export_range space of wave_cascade_parameters.gd (and beyond), two updates
    each at 128x128 -- every resource bit-identical between the C oracle and the compiled reference shaders."""
    from hypothesis import given, settings, HealthCheck
    from hypothesis import strategies as st

    pos = dict(allow_nan=False, allow_infinity=False)
    params = st.fixed_dictionaries(dict(
        tile_length=st.tuples(st.floats(0.5, 4000.0, width=32, **pos), st.floats(0.5, 4000.0, width=32, **pos)),
        wind_speed=st.floats(0.0001, 60.0, **pos), wind_direction=st.floats(-360.0, 720.0, **pos),
        fetch_length=st.floats(0.0001, 2000.0, **pos), swell=st.floats(0.0, 2.0, **pos), spread=st.floats(0.0, 1.0, **pos),
        detail=st.floats(0.0, 1.0, **pos), whitecap=st.floats(0.0, 2.0, **pos), foam_amount=st.floats(0.0, 10.0, **pos),
        spectrum_seed=st.tuples(st.integers(-2**31, 2**31 - 1), st.integers(-2**31, 2**31 - 1)),
        time=st.floats(0.0, 50000.0, **pos)))

    @settings(max_examples=12, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.too_slow])
    @given(kw=params, contract=st.sampled_from([po.CONTRACT_FMA, po.CONTRACT_STRICT]), delta=st.floats(0.0, 0.1, **pos))
    def run(kw, contract, delta):
        po.set_modes(po.MATH_DET, contract)
        pr.set_modes(po.MATH_DET, contract)
        o, r = po.OracleWaveGenerator(128), pr.RefWaveGenerator(128)
        a, b = [po.CascadeParams(**kw)], [po.CascadeParams(**kw)]
        for _ in range(2):
            o.update_all(delta, a)
            r.update_all(delta, b)
        # NaN payloads aside (log(0) at u1 == 0 is reachable in principle), the bytes must agree
        _assert_same_state(o, r, f"{kw} contract={contract}")

    run()
