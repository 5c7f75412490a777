"""CPU-side checks of the drop-in boundary: libocean.so loads, exports every symbol include/ocean.h
declares, fails loudly without a GPU, and the host-side mirrors behave like the reference classes."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import godotoceanwaves_b200 as gow
from godotoceanwaves_b200 import build as native_build
from godotoceanwaves_b200 import native
from oracle import pyoracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    native_build.build_native()
    return gow.load_library()


def test_header_symbols_all_exported(lib):
    header = open(os.path.join(ROOT, "include", "ocean.h")).read()
    declared = set(re.findall(r"\b(ocean_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 25
    assert declared == set(native.SIGNATURES), declared ^ set(native.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name


def test_struct_layout_matches_header():
    # 2 floats + 10 doubles + 3 int32 (+pad) + 3 doubles
    assert C.sizeof(native.CascadeParamsC) == 8 + 80 + 16 + 24
    assert native.CascadeParamsC.time.offset == 104
    assert C.sizeof(native.InfoC) == 40


def test_jonswap_statics_match_reference_formulas(lib):
    # wave_generator.gd:116-121, pins from SURVEY 4
    assert lib.ocean_jonswap_alpha(20.0, 550e3) == pytest.approx(0.009380366716946217, rel=1e-15)
    assert lib.ocean_jonswap_peak_angular_frequency(20.0, 550e3) == pytest.approx(0.45331955874140006, rel=1e-15)
    for U, F in [(10, 150e3), (5, 1e3), (30, 1000e3)]:
        assert lib.ocean_jonswap_alpha(U, F) == po.JONSWAP_alpha(U, F)
        assert lib.ocean_jonswap_peak_angular_frequency(U, F) == po.JONSWAP_peak_angular_frequency(U, F)


def test_defaults_match_wave_cascade_parameters(lib):
    p = native.CascadeParamsC()
    assert lib.ocean_default_cascade_params(C.byref(p)) == 0
    d = gow.WaveCascadeParameters()
    assert (p.tile_length[0], p.tile_length[1]) == d.tile_length == (50.0, 50.0)
    for f in ("displacement_scale", "normal_scale", "wind_speed", "wind_direction", "fetch_length", "swell", "spread",
              "detail", "whitecap", "foam_amount"):
        assert getattr(p, f) == getattr(d, f), f
    assert p.should_generate_spectrum == 1 and d.should_generate_spectrum


def test_no_cpu_fallback_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    g = gow.WaveGenerator()
    g.map_size = 256
    with pytest.raises(gow.OceanError, match="no CPU fallback"):
        g.init_gpu(2)
    assert lib.ocean_destroy(None) != 0 and b"NULL" in lib.ocean_last_error()
    with pytest.raises(gow.OceanError):
        g.maps_to_host()


def test_wave_cascade_parameters_setters():
    # wave_cascade_parameters.gd:7-35: every setter but the two scales raises the dirty flag; clamps
    p = gow.WaveCascadeParameters()
    p.should_generate_spectrum = False
    p.displacement_scale = 0.5
    p.normal_scale = 0.5
    assert not p.should_generate_spectrum
    for name, val in [("tile_length", (10, 20)), ("wind_speed", 3.0), ("wind_direction", 45.0), ("fetch_length", 10.0),
                      ("swell", 0.1), ("spread", 0.3), ("detail", 0.9), ("whitecap", 0.7), ("foam_amount", 2.0)]:
        p.should_generate_spectrum = False
        setattr(p, name, val)
        assert p.should_generate_spectrum, name
    p.wind_speed = -5
    p.fetch_length = 0
    assert p.wind_speed == 0.0001 and p.fetch_length == 0.0001
    c = native.CascadeParamsC()
    p.spectrum_seed = (3, -4)
    p.time = 12.5
    p.to_c(c)
    assert (c.spectrum_seed[0], c.spectrum_seed[1]) == (3, -4) and c.time == 12.5 and c.tile_length[1] == 20.0


def test_push_constant_mirror_matches_oracle_restatement():
    data = [1234, -5678, 88.0, 88.0, 0.009202510754677472, 0.8807208260620296, 10.0, 0.3490658503988659, 20.0, 0.8,
            1.0, 0.2, 3]
    assert gow.RenderingContext.create_push_constant(data) == po.create_push_constant(data)
    assert len(gow.RenderingContext.create_push_constant(data)) == 64
    with pytest.raises(AssertionError):
        gow.RenderingContext.create_push_constant([0.0] * 33)


def test_host_detmath_exp_matches_the_oracle_bit_for_bit():
    """exp(-foam_decay_rate) is evaluated inside libocean.so on the HOST (ocean_api.cu: exp_det_host); it must be the very
    DETMATH function the oracle (and the device code) define -- checked here without a GPU."""
    import numpy as np
    lib = gow.load_library()
    po.set_modes(po.MATH_DET, po.CONTRACT_FMA)
    ol = po.lib()
    rng = np.random.default_rng(5)
    xs = np.concatenate([
        -np.abs(rng.standard_normal(20000)).astype(np.float32),                 # typical: -delta*(10-foam_amount)*1.15
        rng.uniform(-120.0, 100.0, 20000).astype(np.float32),                   # across both clamps
        np.array([0.0, -0.0, 1.0, -1.0, -0.0115, -0.23, -110.0, -110.5, 90.0, 90.5, 1e-30, -1e-30, 88.7, -87.3, -103.9,
                  np.inf, -np.inf], np.float32)])
    a = np.array([lib.ocean_detmath_expf(float(x)) for x in xs], np.float32)
    b = np.array([ol.oracle_expf(float(x)) for x in xs], np.float32)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    n = lib.ocean_detmath_expf(float("nan"))
    assert n != n


@pytest.mark.parametrize("map_size,count,group,lag", [(256, 128, 0, 0), (256, 128, 16, 1), (256, 7, 2, 3), (128, 9, 3, 4), (512, 16, 0, 0),
                                                     (1024, 8, 0, 0), (256, 5, 8, 1), (128, 260 % 256 + 4, 1, 1)])
def test_work_queue_order_is_deadlock_free(map_size, count, group, lag):
    """The persistent kernel hands work items out in table order and a column-pass (B) item waits for the row pass of its cascade:
    progress is guaranteed because every row-pass (A) item of a cascade precedes every B item of that cascade (a waiting team only
    waits for items that were handed out before its own).  Checked on the host for the default shapes and for shapes with more lag
    than groups: each item exactly once, A before B per cascade, and never more than (lag + 1) groups between the two passes."""
    import ctypes as C
    lib = native.load_library()
    total = lib.ocean_debug_work_queue(map_size, count, group, lag, 0, None, 0)
    assert total > 0
    items = (C.c_int32 * total)()
    assert lib.ocean_debug_work_queue(map_size, count, group, lag, 0, items, total) == total
    codes = np.frombuffer(items, dtype=np.int32).astype(np.int64) & 0xFFFFFFFF
    is_b = (codes >> 31) & 1
    slot = (codes >> 16) & 0x7FFF
    block = codes & 0xFFFF
    assert slot.max() == count - 1
    seen = set()
    last_a = {}
    first_b = {}
    for pos, (b, s, x) in enumerate(zip(is_b, slot, block)):
        key = (int(b), int(s), int(x))
        assert key not in seen
        seen.add(key)
        if b:
            first_b.setdefault(int(s), pos)
        else:
            last_a[int(s)] = pos
    a_per = sum(1 for k in seen if k[0] == 0 and k[1] == 0)
    b_per = sum(1 for k in seen if k[0] == 1 and k[1] == 0)
    assert len(seen) == count * (a_per + b_per) and a_per > 0 and b_per > 0
    for s in range(count):
        assert last_a[s] < first_b[s], s
    assert is_b[0] == 0 and is_b[-1] == 1


@pytest.mark.parametrize("map_size,count,frames", [(512, 4, 64), (256, 4, 3), (128, 1, 1), (256, 7, 2), (1024, 8, 5)])
def test_fused_frames_queue_order_is_deadlock_free(map_size, count, frames):
    """ocean_update_frames: frames alternate between the two halves of the row-pass scratch.  A row-pass item of (frame f, cascade c)
    waits for the column pass of (f - 2, c) -- the last reader of its half --, a column-pass item of (f, c) for the row pass of
    (f, c) and for the column pass of (f - 1, c) (foam plane, maps).  Every one of those must precede the waiting item in the
    hand-out order; every item appears exactly once."""
    import ctypes as C
    lib = native.load_library()
    total = lib.ocean_debug_work_queue(map_size, count, 0, 0, frames, None, 0)
    assert total > 0
    items = (C.c_int32 * total)()
    assert lib.ocean_debug_work_queue(map_size, count, 0, 0, frames, items, total) == total
    codes = np.frombuffer(items, dtype=np.int32).astype(np.int64) & 0xFFFFFFFF
    is_b, slot, block = (codes >> 31) & 1, (codes >> 16) & 0x7FFF, codes & 0xFFFF
    first = {}      # (kind, frame, cascade) -> first position
    last = {}
    seen = set()
    for pos, (b, s, x) in enumerate(zip(is_b, slot, block)):
        key = (int(b), int(s), int(x))
        assert key not in seen
        seen.add(key)
        k = (int(b), int(s) // count, int(s) % count)
        first.setdefault(k, pos)
        last[k] = pos
    assert len(first) == 2 * frames * count
    per = {0: sum(1 for k in seen if k[0] == 0 and k[1] == 0), 1: sum(1 for k in seen if k[0] == 1 and k[1] == 0)}
    assert len(seen) == frames * count * (per[0] + per[1])
    for f in range(frames):
        for c in range(count):
            assert last[(0, f, c)] < first[(1, f, c)]
            if f >= 1:
                assert last[(1, f - 1, c)] < first[(1, f, c)]
            if f >= 2:
                assert last[(1, f - 2, c)] < first[(0, f, c)]
