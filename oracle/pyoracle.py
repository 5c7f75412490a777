"""oracle/pyoracle.py -- TEST INFRASTRUCTURE (not part of the product).

ctypes front-end of the CPU oracle (oracle/ocean_oracle.c) plus a restatement of the
reference's host-side sequencing:

  * create_push_constant      <- assets/render_context.gd:122-135
  * JONSWAP_alpha / _peak_... <- assets/water/wave_generator.gd:116-121
  * CascadeParams             <- assets/water/wave_cascade_parameters.gd:2-56
  * OracleWaveGenerator       <- assets/water/wave_generator.gd:17-109

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this.
PARITY UNPINNED by the reference (it has no tests/golden data); see ocean_oracle.c header.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import struct
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libocean_oracle.so")

G = 9.81          # wave_generator.gd:5
DEPTH = 20.0      # wave_generator.gd:6

MATH_DET, MATH_LIBM = 0, 1
CONTRACT_STRICT, CONTRACT_FMA = 0, 1


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (oracle/Makefile). Returns the .so path."""
    src = [os.path.join(_HERE, f) for f in ("ocean_oracle.c", "detmath.h", "Makefile")]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B", "libocean_oracle.so"], check=True,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return _LIB_PATH


class PcSpectrumCompute(C.Structure):
    _fields_ = [("seed", C.c_int32 * 2), ("tile_length", C.c_float * 2), ("alpha", C.c_float),
                ("peak_frequency", C.c_float), ("wind_speed", C.c_float), ("angle", C.c_float),
                ("depth", C.c_float), ("swell", C.c_float), ("detail", C.c_float), ("spread", C.c_float),
                ("cascade_index", C.c_uint32)]


class PcSpectrumModulate(C.Structure):
    _fields_ = [("tile_length", C.c_float * 2), ("depth", C.c_float), ("time", C.c_float),
                ("cascade_index", C.c_uint32)]


class PcFftUnpack(C.Structure):
    _fields_ = [("cascade_index", C.c_uint32), ("whitecap", C.c_float), ("foam_grow_rate", C.c_float),
                ("foam_decay_rate", C.c_float)]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        fp = C.POINTER(C.c_float)
        hp = C.POINTER(C.c_uint16)
        L.oracle_set_modes.argtypes = [C.c_int, C.c_int]
        L.oracle_num_threads.restype = C.c_int
        L.oracle_set_num_threads.argtypes = [C.c_int]
        for name in ("cosf", "sinf", "expf", "logf", "tanhf"):
            f = getattr(L, "oracle_" + name); f.argtypes = [C.c_float]; f.restype = C.c_float
        for name in ("powf", "atan2f"):
            f = getattr(L, "oracle_" + name); f.argtypes = [C.c_float, C.c_float]; f.restype = C.c_float
        L.oracle_float_to_half.argtypes = [C.c_float]; L.oracle_float_to_half.restype = C.c_uint16
        L.oracle_half_to_float.argtypes = [C.c_uint16]; L.oracle_half_to_float.restype = C.c_float
        L.oracle_hash.argtypes = [C.c_uint32, C.c_uint32, fp]
        L.oracle_hash_n.argtypes = [C.c_uint32, C.c_uint32]; L.oracle_hash_n.restype = C.c_uint32
        L.oracle_amplitude_factor.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(PcSpectrumCompute)]
        L.oracle_amplitude_factor.restype = C.c_float
        L.oracle_spectrum_compute.argtypes = [fp, C.c_int, C.POINTER(PcSpectrumCompute)]
        L.oracle_spectrum_modulate.argtypes = [fp, fp, C.c_int, C.POINTER(PcSpectrumModulate)]
        L.oracle_fft_butterfly.argtypes = [fp, C.c_int]
        L.oracle_fft_compute.argtypes = [fp, fp, C.c_int]
        L.oracle_transpose.argtypes = [fp, C.c_int]
        L.oracle_fft_unpack.argtypes = [fp, hp, hp, fp, fp, C.c_int, C.POINTER(PcFftUnpack)]
        L.oracle_cascade_update.argtypes = [fp, fp, fp, hp, hp, fp, fp, C.c_int, C.c_int,
                                            C.POINTER(PcSpectrumCompute), C.POINTER(PcSpectrumModulate),
                                            C.POINTER(PcFftUnpack)]
        L.oracle_cascade_update_batch.argtypes = [fp, fp, fp, hp, hp, C.c_int, C.c_int, C.POINTER(C.c_int),
                                                  C.POINTER(PcSpectrumCompute), C.POINTER(PcSpectrumModulate),
                                                  C.POINTER(PcFftUnpack)]
        _lib = L
    return _lib


def set_modes(math_mode: int = MATH_DET, contract_mode: int = CONTRACT_FMA) -> None:
    lib().oracle_set_modes(math_mode, contract_mode)


def _fp(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags.c_contiguous
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _hp(a: np.ndarray):
    assert a.dtype == np.uint16 and a.flags.c_contiguous
    return a.ctypes.data_as(C.POINTER(C.c_uint16))


# --------------------------------------------------------------------------------------
# Host-side restatement
# --------------------------------------------------------------------------------------
def create_push_constant(data) -> bytes:
    """assets/render_context.gd:122-135: ints -> s32, floats -> binary32, pad to 16 B."""
    packed_size = len(data) * 4
    assert packed_size <= 128, "Push constant size must be at most 128 bytes!"
    padding = math.ceil(packed_size / 16.0) * 16 - packed_size
    out = bytearray(packed_size + (padding if padding > 0 else 0))
    for i, v in enumerate(data):
        if isinstance(v, (bool, int, np.integer)):
            struct.pack_into("<i", out, i * 4, int(v))
        else:
            # encode_float rounds the GDScript float (binary64) to binary32 (RN)
            struct.pack_into("<f", out, i * 4, float(np.float32(v)))
    return bytes(out)


def JONSWAP_alpha(wind_speed: float = 20.0, fetch_length: float = 550e3) -> float:
    """wave_generator.gd:116-117 (binary64)."""
    return 0.076 * math.pow(wind_speed ** 2 / (fetch_length * G), 0.22)


def JONSWAP_peak_angular_frequency(wind_speed: float = 20.0, fetch_length: float = 550e3) -> float:
    """wave_generator.gd:120-121 (binary64)."""
    return 22.0 * math.pow(G * G / (wind_speed * fetch_length), 1.0 / 3.0)


def deg_to_rad(d: float) -> float:
    return d * (math.pi / 180.0)


@dataclass
class CascadeParams:
    """assets/water/wave_cascade_parameters.gd:2-42 (fields only; the Godot setters clamp
    wind_speed/fetch_length to >= 1e-4 and raise should_generate_spectrum)."""
    tile_length: tuple = (50.0, 50.0)
    displacement_scale: float = 1.0
    normal_scale: float = 1.0
    wind_speed: float = 20.0
    wind_direction: float = 0.0
    fetch_length: float = 550.0
    swell: float = 0.8
    spread: float = 0.2
    detail: float = 1.0
    whitecap: float = 0.5
    foam_amount: float = 5.0
    spectrum_seed: tuple = (0, 0)
    should_generate_spectrum: bool = True
    time: float = 0.0
    foam_grow_rate: float = 0.0
    foam_decay_rate: float = 0.0


def pc_spectrum_compute(p, cascade_index: int) -> PcSpectrumCompute:
    """Push constant of wave_generator.gd:69-71 as the struct of spectrum_compute.glsl:18-30."""
    alpha = JONSWAP_alpha(p.wind_speed, p.fetch_length * 1e3)
    omega = JONSWAP_peak_angular_frequency(p.wind_speed, p.fetch_length * 1e3)
    # Vector2 components are binary32 in Godot
    tl = (float(np.float32(p.tile_length[0])), float(np.float32(p.tile_length[1])))
    raw = create_push_constant([int(p.spectrum_seed[0]), int(p.spectrum_seed[1]), tl[0], tl[1], alpha, omega,
                                float(p.wind_speed), deg_to_rad(p.wind_direction), DEPTH, float(p.swell),
                                float(p.detail), float(p.spread), int(cascade_index)])
    return PcSpectrumCompute.from_buffer_copy(raw[:C.sizeof(PcSpectrumCompute)])


def pc_spectrum_modulate(p, cascade_index: int) -> PcSpectrumModulate:
    tl = (float(np.float32(p.tile_length[0])), float(np.float32(p.tile_length[1])))
    raw = create_push_constant([tl[0], tl[1], DEPTH, float(p.time), int(cascade_index)])
    return PcSpectrumModulate.from_buffer_copy(raw[:C.sizeof(PcSpectrumModulate)])


def pc_fft_unpack(p, cascade_index: int) -> PcFftUnpack:
    raw = create_push_constant([int(cascade_index), float(p.whitecap), float(p.foam_grow_rate),
                                float(p.foam_decay_rate)])
    return PcFftUnpack.from_buffer_copy(raw[:C.sizeof(PcFftUnpack)])


class OracleWaveGenerator:
    """assets/water/wave_generator.gd restated on numpy buffers + the C oracle."""

    def __init__(self, map_size: int):
        self.map_size = int(map_size)
        self.num_cascades = 0
        self.pass_parameters = []
        self.pass_num_cascades_remaining = 0
        self.context = False
        self.keep_f32 = True

    # wave_generator.gd:17-54
    def init_gpu(self, num_cascades: int) -> None:
        N = self.map_size
        S = int(round(math.log2(N)))
        self.num_cascades = num_cascades
        self.spectrum = np.zeros((num_cascades, N, N, 4), np.float32)              # :31
        self.butterfly = np.zeros((S, N, 4), np.float32)                           # :32
        self.fft_buffer = np.zeros((num_cascades, 2, 4, N, N, 2), np.float32)      # :33
        self.displacement_map = np.zeros((num_cascades, N, N, 4), np.uint16)       # :34
        self.normal_map = np.zeros((num_cascades, N, N, 4), np.uint16)             # :35
        self.displacement_f32 = np.zeros((num_cascades, N, N, 4), np.float32)
        self.normal_f32 = np.zeros((num_cascades, N, N, 4), np.float32)
        lib().oracle_fft_butterfly(_fp(self.butterfly), N)                          # :52-54
        self.context = True

    # wave_generator.gd:65-85
    def _update(self, cascade_index: int, parameters) -> None:
        p = parameters[cascade_index]
        gen = bool(p.should_generate_spectrum)
        pcg = pc_spectrum_compute(p, cascade_index)
        if gen:
            p.should_generate_spectrum = False
        pcm = pc_spectrum_modulate(p, cascade_index)
        pcu = pc_fft_unpack(p, cascade_index)
        i = cascade_index
        lib().oracle_cascade_update(_fp(self.spectrum[i]), _fp(self.fft_buffer[i]), _fp(self.butterfly),
                                    _hp(self.displacement_map[i]), _hp(self.normal_map[i]),
                                    _fp(self.displacement_f32[i]) if self.keep_f32 else None,
                                    _fp(self.normal_f32[i]) if self.keep_f32 else None,
                                    self.map_size, int(gen), C.byref(pcg), C.byref(pcm), C.byref(pcu))

    # wave_generator.gd:56-63
    def process(self) -> None:
        if self.pass_num_cascades_remaining == 0:
            return
        self.pass_num_cascades_remaining -= 1
        self._update(self.pass_num_cascades_remaining, self.pass_parameters)

    # wave_generator.gd:90-109
    def update(self, delta: float, parameters) -> None:
        assert len(parameters) != 0
        if not self.context:
            self.init_gpu(max(2, len(parameters)))
        elif self.pass_num_cascades_remaining != 0:
            for i in range(self.pass_num_cascades_remaining):
                self._update(i, self.pass_parameters)
        for p in parameters:
            p.time += delta
            p.foam_grow_rate = delta * p.foam_amount * 7.5
            p.foam_decay_rate = delta * max(0.5, 10.0 - p.foam_amount) * 1.15
        self.pass_parameters = parameters
        self.pass_num_cascades_remaining = len(parameters)

    def update_all(self, delta: float, parameters) -> None:
        """update() followed by draining every pending cascade (what N frames of _process do)."""
        self.update(delta, parameters)
        while self.pass_num_cascades_remaining:
            self.process()

    def update_all_batched(self, delta: float, parameters) -> None:
        """Same result as update_all, with the pending cascades run as ONE batch, one OpenMP thread per cascade
        (oracle_cascade_update_batch) -- the CPU baseline's way of using every host core on a many-cascade step."""
        self.update(delta, parameters)
        n = self.pass_num_cascades_remaining
        assert n == len(parameters) == self.num_cascades or n <= self.num_cascades
        gen = (C.c_int * n)()
        pcg = (PcSpectrumCompute * n)()
        pcm = (PcSpectrumModulate * n)()
        pcu = (PcFftUnpack * n)()
        for i in range(n):
            p = self.pass_parameters[i]
            gen[i] = int(bool(p.should_generate_spectrum))
            pcg[i] = pc_spectrum_compute(p, i)
            p.should_generate_spectrum = False
            pcm[i] = pc_spectrum_modulate(p, i)
            pcu[i] = pc_fft_unpack(p, i)
        lib().oracle_cascade_update_batch(_fp(self.spectrum), _fp(self.fft_buffer), _fp(self.butterfly), _hp(self.displacement_map),
                                          _hp(self.normal_map), self.map_size, n, gen, pcg, pcm, pcu)
        self.pass_num_cascades_remaining = 0

    # convenience views
    def displacement_half(self) -> np.ndarray:
        return self.displacement_map.view(np.float16)

    def normal_half(self) -> np.ndarray:
        return self.normal_map.view(np.float16)
