"""CPU checks of the spray-candidate specification (oracle/spray.py; sea_spray_particle.gdshader:45-66,80-94)."""
import numpy as np

from oracle import spray as sy


def _maps(N, C, seed, foam_level):
    rng = np.random.default_rng(seed)
    n = np.zeros((C, N, N, 4), np.float16)
    n[..., 0] = rng.normal(0, 0.08, (C, N, N))
    n[..., 1] = rng.normal(0, 0.08, (C, N, N))
    n[..., 2] = rng.normal(0, 0.3, (C, N, N))
    n[..., 3] = np.clip(rng.normal(foam_level, 0.25, (C, N, N)), 0, 1)
    return n


def test_grid_matches_the_shader_formula():
    pts = sy.spray_grid(10000)
    t = 100
    assert pts.shape == (10000, 2)
    assert pts[0].tolist() == [-5.0, -5.0] and pts[t * t - 1].tolist() == [5.0, 5.0]           # corners of the 10 x 10 box
    assert pts[1, 0] == pts[0, 0] and pts[1, 1] > pts[0, 1]                                   # INDEX % t runs along z
    E = np.array([[2, 0, 0, 7], [0, 1, 0, 0], [0, 0, 3, -4]], np.float32)
    moved = sy.spray_grid(10000, E)
    assert np.allclose(moved[:, 0], pts[:, 0] * 2 + 7) and np.allclose(moved[:, 1], pts[:, 1] * 3 - 4)
    # a non-square count: t = 31, indices beyond t*t run off the box exactly as INDEX / t does in the shader
    odd = sy.spray_grid(1000)
    assert odd[31 * 31, 0] > 5.0


def test_candidates_follow_the_shader_rules():
    N, C = 64, 3
    normal = _maps(N, C, 3, 0.4)
    scales = np.array([[1 / 88.0, 1 / 88.0, 1, 1], [1 / 57.0, 1 / 57.0, 0.75, 1], [1 / 16.0, 1 / 16.0, 0, 0.25]], np.float32)
    pts = sy.spray_grid(4096)
    rec = sy.spray_candidates(normal, pts, scales, (1.0, 2.0, 3.0))
    assert 0 < len(rec) < len(pts)
    assert np.all(np.diff(rec["index"].astype(np.int64)) > 0)                                  # candidate order (stable)
    assert np.all(rec["foam"] > np.float32(0.9))
    assert np.all(rec["scale_factor"] <= 1.0) and np.all(rec["scale_factor"] >= 0.0)
    # independent float64 evaluation of the same rules on the same bilinear gradient
    from oracle.sampling import texture_bilinear
    g = sum(texture_bilinear(normal[i], pts[:, 0] * scales[i, 0], pts[:, 1] * scales[i, 1]).astype(np.float64) for i in range(C))
    ny = 1.0 / np.sqrt(g[:, 0] ** 2 + 1.0 + g[:, 1] ** 2)
    nf = 0.25 + 0.75 * np.minimum((ny - 0.92) / 0.07, 1.0)
    ff = 0.25 + 0.75 * np.minimum((g[:, 3] - 0.9) / 0.1, 1.0)
    active = (nf >= 0) & (nf <= 1) & (g[:, 3] > 0.9)
    margin = np.minimum(np.abs(nf), np.abs(g[:, 3] - 0.9)) > 1e-4                             # away from the decision boundaries
    assert np.array_equal(np.isin(np.arange(len(pts)), rec["index"])[margin], active[margin])
    k = rec["index"]
    assert np.allclose(rec["scale_factor"], (nf * ff)[k], rtol=2e-5)
    assert np.allclose(rec["particle_scale"][:, 1], (ff * 1.001 * nf * 2.0)[k], rtol=2e-5)
    assert np.allclose(rec["particle_scale"][:, 0], (ff * 1.001)[k], rtol=2e-5)
    assert np.allclose(rec["particle_scale"][:, 2], (ff * 1.001 * 3.0)[k], rtol=2e-5)


def test_no_foam_no_candidates():
    N, C = 32, 2
    normal = _maps(N, C, 4, 0.0)
    normal[..., 3] = 0
    scales = np.array([[0.02, 0.02, 1, 1], [0.1, 0.1, 1, 1]], np.float32)
    assert len(sy.spray_candidates(normal, sy.spray_grid(1024), scales, (1, 1, 1))) == 0
