"""Property tier of SURVEY section 4 on INJECTED spectra (ocean_set_spectrum_amplitudes): size-independent facts about the
time propagation + 4 packed inverse FFTs + map assembly that hold whatever the spectrum is -- a single wave vector gives a
pure cosine at the right place of the transposed map, the displacement is linear in the amplitudes (exactly so for powers of
two), Parseval's identity, a zero spectrum gives exactly flat maps.  Checked on the binary32 taps at BASELINE sizes."""
import numpy as np
import pytest

from conftest import demo_params

pytestmark = pytest.mark.gpu

N = 256


def _gen(C=2):
    import godotoceanwaves_b200 as gow
    g = gow.WaveGenerator(); g.map_size = N; g.init_gpu(C); g.enable_f32_taps(True)
    p = [demo_params(gow.WaveCascadeParameters, c, foam_amount=0.0) for c in range(C)]
    g.update_all(0.02, p)                       # builds tables, clears the dirty flags
    return gow, g, p


def _maps(g, p, amplitudes, cascade=0, delta=0.0):
    g.set_spectrum_amplitudes(cascade, amplitudes)
    g.update_all(delta, p)
    return g.f32_maps_to_host(cascade)


def _sparse(rng, count, scale=1.0):
    a = np.zeros((N, N), np.complex64)
    ys, xs = rng.integers(1, N, count), rng.integers(1, N, count)
    a[ys, xs] = (rng.standard_normal(count) + 1j * rng.standard_normal(count)) * scale
    a[N // 2, N // 2] = 0
    return a


def test_zero_spectrum_gives_flat_maps():
    _, g, p = _gen()
    d, n = _maps(g, p, np.zeros((N, N), np.complex64))
    assert not d.any() and not n[..., :3].any()
    g.free()


@pytest.mark.parametrize("x0,y0", [(N // 2 + 5, N // 2 + 9), (N // 2 - 17, N // 2 + 2), (3, 250)])
def test_single_wave_vector_is_one_cosine_in_the_transposed_map(x0, y0):
    """A(id0) = a e^{i phi}, everything else 0: the height map must be 2a cos(...) -- exactly two Fourier bins, at the place the
    reference's conventions put them: centred spectrum (DC at N/2, spectrum_compute.glsl:105), ifftshift by the sign trick
    (fft_unpack.glsl:38) and NO second transpose (wave_generator.gd:77-78), i.e. map texel (x, y) holds the field at (y, x)."""
    _, g, p = _gen()
    a = np.zeros((N, N), np.complex64)
    amp, phi = 0.37, 0.8
    a[y0, x0] = amp * np.exp(1j * phi)
    d, _ = _maps(g, p, a)
    hy = d[..., 1].astype(np.float64)
    assert np.abs(hy).max() <= 2 * amp * (1 + 2e-5)                          # h(k) + conj pair -> a cosine of amplitude 2a
    spec = np.fft.fft2(hy) / (N * N)
    mag = np.abs(spec)
    peaks = np.argwhere(mag > 1e-4 * mag.max())
    # wave-vector index (x0 - N/2, y0 - N/2); the map is transposed: its row index runs along the spectrum's x axis
    kx, ky = (x0 - N // 2) % N, (y0 - N // 2) % N
    expect = {(kx, ky), ((-kx) % N, (-ky) % N)}
    assert {tuple(int(v) for v in pk) for pk in peaks} == expect, (peaks.tolist(), expect)
    assert np.allclose(mag[kx, ky], amp, rtol=2e-5)
    g.free()


def test_displacement_is_linear_in_the_amplitudes():
    _, g, p = _gen()
    rng = np.random.default_rng(21)
    a1, a2 = _sparse(rng, 400), _sparse(rng, 300, 0.5)
    d1, _ = _maps(g, p, a1)
    d2, _ = _maps(g, p, a2)
    d12, _ = _maps(g, p, (a1 + a2).astype(np.complex64))
    for ch in range(3):
        s = d1[..., ch].astype(np.float64) + d2[..., ch]
        assert np.abs(d12[..., ch] - s).max() <= 1e-5 * np.abs(s).max()
    # scaling by a power of two commutes with every rounding in the path: EXACT
    d4, n4 = _maps(g, p, (a1 * np.float32(4.0)).astype(np.complex64))
    assert np.array_equal(d4[..., :3], d1[..., :3] * np.float32(4.0))
    g.free()


def test_parseval_identity_of_the_height_field():
    """sum_x hy(x)^2 == N^2 sum_k |h(k, t)|^2 for the unnormalised inverse transform (fft_compute.glsl has no 1/N)."""
    gow, g, p = _gen()
    rng = np.random.default_rng(5)
    a = _sparse(rng, 2000, 0.05)
    d, _ = _maps(g, p, a, delta=0.25)
    # h(k, t) = h0(k) e^{i w t} + conj(h0(-k)) e^{-i w t}  (spectrum_modulate.glsl:65-68), binary64 on the host
    f32 = np.float32
    L, t, depth = p[0].tile_length, p[0].time, 20.0
    ys, xs = np.meshgrid(np.arange(N), np.arange(N), indexing="ij")
    kvx = (xs.astype(f32) - f32(N * 0.5)) * f32(2.0) * f32(np.pi) / f32(L[0])
    kvy = (ys.astype(f32) - f32(N * 0.5)) * f32(2.0) * f32(np.pi) / f32(L[1])
    k = np.sqrt(kvx * kvx + kvy * kvy) + f32(1e-6)
    w = np.sqrt(f32(9.81) * k * np.tanh(k.astype(np.float64) * depth).astype(f32))
    ph = (w * f32(t)).astype(np.float64)
    a64 = a.astype(np.complex128)
    am = np.conj(a64[(-ys) % N, (-xs) % N])
    h = a64 * np.exp(1j * ph) + am * np.exp(-1j * ph)
    lhs = float((d[..., 1].astype(np.float64) ** 2).sum())
    rhs = float(N * N * (np.abs(h) ** 2).sum())
    assert abs(lhs - rhs) <= 2e-5 * rhs, (lhs, rhs)
    g.free()
