"""What "within 1e-5 of the reference's own pipeline, Jacobian foam-sign mask bit-exact" (BASELINE.json north_star) means
when the reference's GLSL leaves contraction and the transcendental library to the Vulkan driver.

The CUDA path reproduces ONE legal reading bit for bit: DETMATH transcendentals + contraction of x*y +/- z*w ("FMA mode";
tests/test_gpu_parity.py).  The other readings -- no contraction at all (STRICT: what lavapipe's LLVM back end does) and
glibc's correctly rounded libm instead of DETMATH -- are available in the oracle AND in the compiled reference shaders
(oracle/_ref, bit-identical to the oracle in every mode: tests/test_ref_pins_oracle.py).  These tests measure the distance
between the readings with the north star's own metric (per-field max|a-b| / max|b| on the binary32 fields before the
half conversion) and assert it stays two orders of magnitude inside the 1e-5 bar, and that the foam-sign mask
(jacobian < whitecap, visible as foam > 0 after the first update) is identical.  Measured:
  cfg2 256x256x4, 3 updates:          STRICT vs FMA 3.3e-7,  LIBM vs DETMATH 0 (bit-identical),  mask Hamming distance 0
  1024x1024, L = 16 m, t = 120 s:     STRICT vs FMA 6.0e-7,  LIBM vs DETMATH 0,                  mask Hamming distance 0
The foam VALUE (fp16 state) may differ by one half-precision ulp in a few texels per layer between contraction modes
(29/65536 at cfg2): the recurrence re-quantises to half every update (fft_unpack.glsl:61-67).
The CPU tests compare oracle modes; the GPU tests compare the CUDA path with the oracle in the other modes."""
import numpy as np
import pytest

from conftest import demo_params
from oracle import pyoracle as po

BAR = 1e-5          # north_star tolerance
MEASURED_MAX = 1e-6  # what the distance between the legal readings actually is (asserted, with margin over 6.0e-7)

CASES = [("cfg2_256x4", 256, 4, 3, {}), ("worst_phase_1024_L16_t120", 1024, 1, 2, dict(tile_length=(16.0, 16.0), time=120.0))]
OTHER_MODES = [("strict_detmath", po.MATH_DET, po.CONTRACT_STRICT), ("fma_libm", po.MATH_LIBM, po.CONTRACT_FMA),
               ("strict_libm", po.MATH_LIBM, po.CONTRACT_STRICT)]


@pytest.fixture(autouse=True)
def _restore():
    yield
    po.set_modes(po.MATH_DET, po.CONTRACT_FMA)


def _oracle(N, C, frames, math_mode, contract, over):
    po.set_modes(math_mode, contract)
    o = po.OracleWaveGenerator(N)
    p = [demo_params(po.CascadeParams, c, **over) for c in range(C)]
    for _ in range(frames):
        o.update_all(0.02, p)
    return o


def _rel(a, b):
    m = float(np.abs(b).max())
    return float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max()) / (m if m > 0 else 1.0)


def _compare(d32, n32, ref, c):
    """(max relative distance over the seven map fields, foam-mask Hamming distance, texels whose half foam differs)"""
    dist = max([_rel(d32[..., ch], ref.displacement_f32[c][..., ch]) for ch in range(3)] +
               [_rel(n32[..., ch], ref.normal_f32[c][..., ch]) for ch in range(3)])
    hamming = int(((n32[..., 3] > 0) != (ref.normal_f32[c][..., 3] > 0)).sum())
    return dist, hamming


@pytest.mark.parametrize("mode", OTHER_MODES, ids=[m[0] for m in OTHER_MODES])
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_distance_between_legal_readings_of_the_shaders(case, mode):
    _, N, C, frames, over = case
    base = _oracle(N, C, frames, po.MATH_DET, po.CONTRACT_FMA, over)
    other = _oracle(N, C, frames, mode[1], mode[2], over)
    for c in range(C):
        dist, hamming = _compare(base.displacement_f32[c], base.normal_f32[c], other, c)
        assert dist <= MEASURED_MAX <= BAR, (case[0], mode[0], c, dist)
        assert hamming == 0, (case[0], mode[0], c, hamming)
        assert _rel(base.spectrum[c], other.spectrum[c]) <= MEASURED_MAX
        # foam state: at most one half-precision ulp apart, in a handful of texels
        fa, fb = base.normal_half()[c][..., 3].astype(np.float32), other.normal_half()[c][..., 3].astype(np.float32)
        assert float(np.abs(fa - fb).max()) <= 2.0 ** -10 and int((fa != fb).sum()) <= 1e-3 * fa.size


@pytest.mark.gpu
@pytest.mark.parametrize("mode", OTHER_MODES, ids=[m[0] for m in OTHER_MODES])
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_cuda_within_bar_of_every_legal_reading(case, mode):
    """CUDA binary32 taps vs the oracle in STRICT / LIBM modes: <= 1e-5 relative per field (measured <= 6.0e-7), foam-sign
    mask bit-exact."""
    import godotoceanwaves_b200 as gow
    _, N, C, frames, over = case
    ref = _oracle(N, C, frames, mode[1], mode[2], over)
    g = gow.WaveGenerator(); g.map_size = N; g.init_gpu(max(2, C)); g.enable_f32_taps(True)
    p = [demo_params(gow.WaveCascadeParameters, c, **over) for c in range(C)]
    for _ in range(frames):
        g.update_all(0.02, p)
    for c in range(C):
        d32, n32 = g.f32_maps_to_host(c)
        dist, hamming = _compare(d32, n32, ref, c)
        assert dist <= MEASURED_MAX <= BAR, (case[0], mode[0], c, dist)
        assert hamming == 0, (case[0], mode[0], c, hamming)
    g.free()
