"""The Water node's fixed-rate update accumulator (assets/water/water.gd:51-54,62-63,75-82) in the library
(ocean_scheduler_*), against a literal Python restatement of the GDScript, on random frame-time sequences.
Host logic only: runs without a GPU (the C ABI library loads anywhere)."""
import ctypes as C

import numpy as np
import pytest

import godotoceanwaves_b200 as gow
from godotoceanwaves_b200.native import CascadeParamsC, SchedulerC


class GdWater:
    """water.gd:51-54,62-63,75-82 verbatim (GDScript floats are binary64 = Python floats)."""

    def __init__(self, ups=50.0):
        self.updates_per_second = ups
        self.time = 0.0
        self.next_update_time = 0.0
        self.updates = []

    def set_updates_per_second(self, value):
        self.next_update_time = self.next_update_time - (1.0 / (self.updates_per_second + 1e-10) - 1.0 / (value + 1e-10))
        self.updates_per_second = value

    def _process(self, delta):
        if self.updates_per_second == 0 or self.time >= self.next_update_time:
            target_update_delta = 1.0 / (self.updates_per_second + 1e-10)
            update_delta = delta if self.updates_per_second == 0 else target_update_delta + (self.time - self.next_update_time)
            self.next_update_time = self.time + target_update_delta
            self.updates.append(update_delta)
        self.time += delta


@pytest.mark.parametrize("ups", [50.0, 60.0, 12.5, 0.0, 1.0])
def test_scheduler_matches_water_gd(ups):
    lib = gow.load_library()
    rng = np.random.default_rng(int(ups * 7) + 1)
    ref = GdWater(ups)
    s = SchedulerC()
    assert lib.ocean_scheduler_init(C.byref(s), ups) == 0
    got = []
    for frame in range(3000):
        delta = float(rng.choice([1 / 60, 1 / 144, 1 / 30, 0.2]) * rng.uniform(0.7, 1.3))
        if frame == 1500:                                  # the setter keeps the phase of the next update (water.gd:52-54)
            new = 25.0 if ups else 40.0
            ref.set_updates_per_second(new)
            assert lib.ocean_scheduler_set_rate(C.byref(s), new) == 0
        ud = C.c_double(-1.0)
        due = lib.ocean_scheduler_tick(C.byref(s), delta, C.byref(ud))
        n_before = len(ref.updates)
        ref._process(delta)
        assert bool(due) == (len(ref.updates) > n_before)
        if due:
            got.append(ud.value)
        assert s.time == ref.time and s.next_update_time == ref.next_update_time      # bit-identical binary64 state
    assert got == ref.updates and len(got) > 10


def test_map_scales_and_start_times():
    lib = gow.load_library()
    ps = [gow.WaveCascadeParameters(tile_length=(88.0, 57.0), displacement_scale=0.75, normal_scale=0.25),
          gow.WaveCascadeParameters(tile_length=(16.0, 3.0))]
    arr = (CascadeParamsC * 2)()
    for i, p in enumerate(ps):
        p.to_c(arr[i])
    out = np.zeros((2, 4), np.float32)
    assert lib.ocean_map_scales(arr, 2, out.ctypes.data) == 0
    assert np.array_equal(out, gow.WaveGenerator.map_scales(ps))                      # water.gd:102-110
    assert out[0, 0] == np.float32(1.0) / np.float32(88.0) and out[1, 1] == np.float32(1.0) / np.float32(3.0)
    import math
    assert [lib.ocean_water_default_time(i) for i in range(3)] == [120.0 + math.pi * i for i in range(3)]   # water.gd:32
