"""GPU parity of the map-query op (ocean_sample_maps; water.gdshader:27-39,42-84) against oracle/sampling.py on the
generator's own RGBA16F maps: bit-identical binary32 results."""
import numpy as np
import pytest

from conftest import demo_params
from oracle import sampling as sp

pytestmark = pytest.mark.gpu


def _gen(N, C, frames=2):
    import godotoceanwaves_b200 as gow
    g = gow.WaveGenerator(); g.map_size = N; g.init_gpu(max(2, C))
    params = [demo_params(gow.WaveCascadeParameters, c) for c in range(C)]
    for _ in range(frames):
        g.update_all(1.0 / 50.0, params)
    return gow, g, params


def _points(n, seed, span):
    rng = np.random.default_rng(seed)
    pts = rng.uniform(-span, span, (n, 2)).astype(np.float32)
    # texel centres, texel edges, the origin, whole tiles away: the corners of the addressing logic
    pts[:8] = np.array([[0, 0], [0.34375, 0.34375], [88.0, -88.0], [-0.0, 57.0], [1e-30, -1e-30], [16.0, 16.0], [-1234.5, 987.25],
                        [4096.0, -4096.0]], np.float32)
    return pts


@pytest.mark.parametrize("N,C", [(128, 3), (256, 4), (512, 2)])
def test_sample_maps_bit_exact(N, C):
    gow, g, params = _gen(N, C)
    d16, n16 = g.maps_to_host(0, C)
    scales = gow.WaveGenerator.map_scales(params)
    scales[:, 2] = [1.0, 0.75, 0.0, 0.5][:C]                   # displacement scales of main.tscn:43-83 (+ one more)
    scales[:, 3] = [1.0, 1.0, 0.25, 0.5][:C]
    pts = _points(20000, 11 + N, 300.0)
    d, gr = g.sample(pts, scales)
    d_ref, g_ref = sp.sample_maps(d16, n16, pts, scales)
    assert np.array_equal(d.view(np.uint32), d_ref.view(np.uint32))
    assert np.array_equal(gr.view(np.uint32), g_ref.view(np.uint32))
    # fewer cascades = a prefix sum
    d1, g1 = g.sample(pts[:1000], scales[:1])
    d1_ref, g1_ref = sp.sample_maps(d16[:1], n16[:1], pts[:1000], scales[:1])
    assert np.array_equal(d1.view(np.uint32), d1_ref.view(np.uint32)) and np.array_equal(g1.view(np.uint32), g1_ref.view(np.uint32))
    g.free()


def test_sample_maps_texel_centres_return_the_texels():
    gow, g, params = _gen(128, 2)
    d16, n16 = g.maps_to_host(0, 2)
    N = 128
    L = np.float32(64.0)                                        # dyadic tile: u = x/L is exact
    xs, ys = np.meshgrid(np.arange(N), np.arange(N))
    pts = np.stack([(xs.ravel() + 0.5) * (L / N), (ys.ravel() + 0.5) * (L / N)], 1).astype(np.float32)
    scales = np.array([[1 / L, 1 / L, 1.0, 1.0]], np.float32)
    d, gr = g.sample(pts, scales)
    assert np.array_equal(d, d16[0].astype(np.float32).reshape(-1, 4)[:, :3])
    # ppm = 128/64 = 2 -> t = 0.2: mostly bicubic, so only the oracle comparison applies to the gradient
    _, g_ref = sp.sample_maps(d16[:1], n16[:1], pts, scales)
    assert np.array_equal(gr.view(np.uint32), g_ref.view(np.uint32))
    g.free()


def test_sample_maps_arguments():
    gow, g, params = _gen(128, 2, frames=1)
    scales = gow.WaveGenerator.map_scales(params)
    d, gr = g.sample(np.zeros((0, 2), np.float32), scales)       # empty batch
    assert d.shape == (0, 3) and gr.shape == (0, 3)
    with pytest.raises(gow.OceanError):
        g.sample(np.zeros((4, 2), np.float32), np.zeros((5, 4), np.float32))    # more cascades than layers
    with pytest.raises(gow.OceanError):
        g.sample(np.zeros((4, 2), np.float32), np.zeros((0, 4), np.float32))
    g.free()
