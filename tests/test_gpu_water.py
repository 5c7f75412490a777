"""The Water node pair end to end (assets/water/water.gd:75-82,112-114 + assets/water/wave_generator.gd:56-63,90-109): rendered
frames with irregular frame times drive the library's scheduler (ocean_water_frame); the oracle is driven by a verbatim
restatement of the GDScript.  Maps must agree bit for bit after every frame that finished a cascade."""
import numpy as np
import pytest

from conftest import demo_params
from oracle import pyoracle as po
from test_water_scheduler_cpu import GdWater

pytestmark = pytest.mark.gpu


def test_water_frames_against_gdscript_sequencing():
    import godotoceanwaves_b200 as gow
    N, C = 128, 3
    w = gow.Water(map_size=N, updates_per_second=50.0)
    pg = [demo_params(gow.WaveCascadeParameters, c) for c in range(C)]
    w.parameters = pg                                         # start times 120 + PI*i, generator set up, spectra dirty
    pc = [demo_params(po.CascadeParams, c) for c in range(C)]
    assert [p.time for p in pg] == [p.time for p in pc]       # demo_params uses the same water.gd:32 offsets
    sched = GdWater(50.0)
    o = po.OracleWaveGenerator(N)
    o.keep_f32 = False
    rng = np.random.default_rng(17)
    updates = 0
    for frame in range(120):
        delta = float(rng.choice([1 / 144, 1 / 60, 1 / 30]) * rng.uniform(0.8, 1.25))
        if frame == 70:
            w.updates_per_second = 20.0
            sched.set_updates_per_second(20.0)
        # reference order within one frame: Water._process (parent) then WaveGenerator._process (child)
        n0 = len(sched.updates)
        sched._process(delta)
        if len(sched.updates) > n0:
            o.update(sched.updates[-1], pc)
        o.process()
        did = w.process(delta)
        assert did == (len(sched.updates) > n0)
        updates += int(did)
        assert w.wave_generator.pass_num_cascades_remaining == o.pass_num_cascades_remaining
        assert w.time == sched.time and w.next_update_time == sched.next_update_time
    assert updates >= 30
    assert [p.time for p in pg] == [p.time for p in pc]
    d, n = w.wave_generator.maps_to_host(0, C)
    assert np.array_equal(d.view(np.uint16), o.displacement_map[:C]) and np.array_equal(n.view(np.uint16), o.normal_map[:C])
    assert np.array_equal(w.map_scales(), gow.WaveGenerator.map_scales(pg))
    assert len(w.layer_bytes()) == C and len(w.layer_bytes()[0][0]) == N * N * 8
    w.free()
