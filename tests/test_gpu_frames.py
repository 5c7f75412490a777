"""Fused multi-frame updates (ocean_update_frames) and the SURVEY 8d cfg3 check: 512x512 x 4 cascades, 1000-frame foam
accumulate/decay loop, foam plane against the oracle at frames {1, 10, 100, 1000} (fft_unpack.glsl:59-64: the recurrence
runs through RGBA16F storage every update)."""
import numpy as np
import pytest

from conftest import demo_params
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


def _gen(N, C):
    import godotoceanwaves_b200 as gow
    g = gow.WaveGenerator(); g.map_size = N; g.init_gpu(max(2, C))
    return gow, g, [demo_params(gow.WaveCascadeParameters, c) for c in range(C)]


@pytest.mark.parametrize("N,C,frames", [(128, 3, 100), (256, 4, 70), (512, 2, 9)])
def test_fused_frames_equal_frame_by_frame(N, C, frames):
    """frames spanning several launches (256 / C frames per launch) == the same number of update_all calls, bit for bit."""
    gow, a, pa = _gen(N, C)
    _, b, pb = _gen(N, C)
    a.update_frames(0.02, pa, frames)
    for _ in range(frames):
        b.update_all(0.02, pb)
    assert [p.time for p in pa] == [p.time for p in pb]
    assert [p.foam_grow_rate for p in pa] == [p.foam_grow_rate for p in pb]
    da, na = a.maps_to_host(0, C)
    db, nb = b.maps_to_host(0, C)
    assert da.tobytes() == db.tobytes() and na.tobytes() == nb.tobytes()
    # and the generator keeps working frame by frame afterwards (counters, pending state)
    a.update_all(0.02, pa); b.update_all(0.02, pb)
    a.update_frames(0.02, pa, 3)
    for _ in range(3):
        b.update_all(0.02, pb)
    da, na = a.maps_to_host(0, C)
    db, nb = b.maps_to_host(0, C)
    assert da.tobytes() == db.tobytes() and na.tobytes() == nb.tobytes()
    assert a.pass_num_cascades_remaining == 0
    a.free(); b.free()


def test_cfg3_foam_loop_1000_frames_against_oracle():
    """BASELINE.json configs[2]: 512x512 x 4, 1000 sequential updates, foam state carried; both RGBA16F maps (the foam
    plane is normal.a) equal to the oracle's at frames 1, 10, 100 and 1000."""
    N, C = 512, 4
    gow, g, pg = _gen(N, C)
    po.set_modes(po.MATH_DET, po.CONTRACT_FMA)
    o = po.OracleWaveGenerator(N)
    o.keep_f32 = False
    pc = [demo_params(po.CascadeParams, c) for c in range(C)]
    at = 0
    for checkpoint in (1, 10, 100, 1000):
        g.update_frames(1.0 / 50.0, pg, checkpoint - at)
        for _ in range(checkpoint - at):
            o.update_all(1.0 / 50.0, pc)
        at = checkpoint
        d, n = g.maps_to_host(0, C)
        foam_gpu = n.view(np.uint16)[..., 3]
        assert np.array_equal(foam_gpu, o.normal_map[:C][..., 3]), f"foam plane, frame {checkpoint}"
        assert np.array_equal(n.view(np.uint16), o.normal_map[:C]) and np.array_equal(d.view(np.uint16), o.displacement_map[:C]), checkpoint
        assert [p.time for p in pg] == [p.time for p in pc]
    assert foam_gpu.view(np.float16).max() > 0
    g.free()
