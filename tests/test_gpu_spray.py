"""GPU parity of the spray-candidate op (ocean_extract_spray; sea_spray_particle.gdshader:80-94) against its numpy
specification oracle/spray.py, on the generator's own maps.  Bar: bit-identical records, identical candidate set."""
import numpy as np
import pytest

from conftest import demo_params
from oracle import spray as sy

pytestmark = pytest.mark.gpu


def _generator(N, C, frames, **over):
    import godotoceanwaves_b200 as gow
    g = gow.WaveGenerator(); g.map_size = N; g.init_gpu(max(2, C))
    p = [demo_params(gow.WaveCascadeParameters, c, **over) for c in range(C)]
    for _ in range(frames):
        g.update_all(1.0 / 50.0, p)
    return gow, g, p


@pytest.mark.parametrize("N,C,particles", [(128, 3, 10000), (256, 4, 65536), (512, 2, 250000)])
def test_spray_records_bit_exact(N, C, particles):
    # a foamy sea: enough updates of a rough sea state for the foam plane to pass 0.9 in places
    gow, g, p = _generator(N, C, 25, whitecap=0.9, foam_amount=10.0)
    _, n16 = g.maps_to_host(0, C)
    scales = gow.WaveGenerator.map_scales(p)
    E = np.array([[7.5, 0, 0, 3.25], [0, 1, 0, 0], [0, 0, 7.5, -11.0]], np.float32)          # a scaled, shifted emitter box
    pts = gow.WaveGenerator.spray_grid(particles, E)
    assert np.array_equal(pts.view(np.uint32), sy.spray_grid(particles, E).view(np.uint32))
    rec, count = g.extract_spray(pts, scales, (0.6, 1.4, 0.6))
    ref = sy.spray_candidates(n16, pts, scales, (0.6, 1.4, 0.6))
    assert count == len(rec) == len(ref) and 0 < count < particles, (count, len(ref))
    assert rec.tobytes() == ref.tobytes()
    # max_records cuts the output, not the count
    few, count2 = g.extract_spray(pts, scales, (0.6, 1.4, 0.6), max_records=7)
    assert count2 == count and few.tobytes() == ref[:7].tobytes()
    g.free()


def test_spray_calm_sea_has_no_candidates_and_empty_input():
    gow, g, p = _generator(128, 2, 2, foam_amount=0.0)
    scales = gow.WaveGenerator.map_scales(p)
    rec, count = g.extract_spray(gow.WaveGenerator.spray_grid(4096), scales)
    assert count == 0 and len(rec) == 0
    rec, count = g.extract_spray(np.zeros((0, 2), np.float32), scales)
    assert count == 0
    g.free()
