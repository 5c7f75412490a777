"""oracle/pyref.py -- TEST INFRASTRUCTURE (not part of the product).

ctypes front-end of oracle/_ref/libocean_ref.so: the reference's OWN six compute shaders
(/root/reference/assets/shaders/compute/*.glsl), compiled for the CPU by the recipe under oracle/ref/ (glsl2cpp.py,
glsl_shim.hpp, shader_tu.cpp, ref_runtime.cpp, Makefile).  `RefWaveGenerator` drives them exactly the way
assets/water/wave_generator.gd does: the resources of :31-35, the uniform sets of :37-41, the workgroup counts of
:44-49 and the push constants of :71,73,76,85 packed by create_push_constant (render_context.gd:122-135).

This is what pins the oracle: tests/test_ref_pins_oracle.py asserts that oracle/ocean_oracle.c (the C restatement the
GPU parity tests compare against) reproduces these shaders' outputs bit for bit.

The library can only be BUILT where /root/reference exists (this container); the GPU box uses the prebuilt file that
travels with the gpurun snapshot.  Only tests/, tools/ and bench.py's CPU-baseline legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np

from . import pyoracle as po

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libocean_ref.so")
REFERENCE = os.environ.get("OCEAN_REFERENCE", "/root/reference")
SHADER_DIR = os.path.join(REFERENCE, "assets", "shaders", "compute")
SHADERS = ("spectrum_compute", "spectrum_modulate", "fft_butterfly", "fft_compute", "transpose", "fft_unpack")

FORMAT_RGBA32F, FORMAT_RGBA16F = 0, 1      # DATA_FORMAT_R32G32B32A32_SFLOAT / R16G16B16A16_SFLOAT


def reference_present() -> bool:
    return all(os.path.exists(os.path.join(SHADER_DIR, s + ".glsl")) for s in SHADERS)


def available() -> bool:
    return os.path.exists(_LIB_PATH) or reference_present()


def build(force: bool = False) -> str | None:
    """Runs oracle/ref/Makefile when the reference sources are present (make decides what is stale).
    Returns the library path, or None when neither the sources nor a prebuilt library exist."""
    if reference_present():
        args = ["make", "-C", os.path.join(_HERE, "ref"), f"REFERENCE={REFERENCE}"]
        if force:
            args.append("-B")
        subprocess.run(args, check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return _LIB_PATH if os.path.exists(_LIB_PATH) else None


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        path = build()
        if path is None:
            raise RuntimeError("oracle/_ref/libocean_ref.so is missing and /root/reference is not available to build it")
        L = C.CDLL(path)
        L.ref_set_modes.argtypes = [C.c_int, C.c_int]
        L.ref_set_num_threads.argtypes = [C.c_int]
        L.ref_num_threads.restype = C.c_int
        L.ref_has_shader.argtypes = [C.c_char_p]
        L.ref_bind_buffer.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_void_p]
        L.ref_bind_image.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.ref_dispatch.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.ref_local_size.argtypes = [C.c_char_p, C.POINTER(C.c_int)]
        for s in SHADERS:
            assert L.ref_has_shader(s.encode()), s
        _lib = L
    return _lib


def set_modes(math_mode: int = po.MATH_DET, contract_mode: int = po.CONTRACT_FMA) -> None:
    lib().ref_set_modes(math_mode, contract_mode)


class RefWaveGenerator(po.OracleWaveGenerator):
    """wave_generator.gd with the reference's shaders underneath (same host-side sequencing as the oracle's generator:
    update / process / update_all are inherited, init_gpu and _update dispatch the compiled GLSL)."""

    # wave_generator.gd:17-54
    def init_gpu(self, num_cascades: int) -> None:
        L = lib()
        N = self.map_size
        S = int(math.log(N) / math.log(2) + 1e-9)                                   # :29 int(log(map_size) / log(2))
        assert 1 << S == N
        self.num_cascades = num_cascades
        self.spectrum = np.zeros((num_cascades, N, N, 4), np.float32)               # :31 R32G32B32A32_SFLOAT x layers
        self.butterfly = np.zeros((S, N, 4), np.float32)                            # :32
        self.fft_buffer = np.zeros((num_cascades, 2, 4, N, N, 2), np.float32)       # :33
        self.displacement_map = np.zeros((num_cascades, N, N, 4), np.uint16)        # :34 R16G16B16A16_SFLOAT
        self.normal_map = np.zeros((num_cascades, N, N, 4), np.uint16)              # :35
        self.displacement_f32 = None                                                # the shaders have no binary32 taps
        self.normal_f32 = None

        def img(shader, set_, binding, arr, fmt):
            rc = L.ref_bind_image(shader.encode(), set_, binding, arr.ctypes.data, N, N, num_cascades, fmt)
            assert rc == 0, (shader, set_, binding, rc)

        def buf(shader, set_, binding, arr):
            rc = L.ref_bind_buffer(shader.encode(), set_, binding, arr.ctypes.data)
            assert rc == 0, (shader, set_, binding, rc)

        # uniform sets :37-41 as bound by the pipelines of :44-49 (binding index = position in the descriptor list)
        img("spectrum_compute", 0, 0, self.spectrum, FORMAT_RGBA32F)                # spectrum_set
        img("spectrum_modulate", 0, 0, self.spectrum, FORMAT_RGBA32F)               # spectrum_set, fft_buffer_set
        buf("spectrum_modulate", 1, 0, self.fft_buffer)
        buf("fft_butterfly", 0, 0, self.butterfly)                                  # fft_butterfly_set
        buf("fft_compute", 0, 0, self.butterfly)                                    # fft_compute_set
        buf("fft_compute", 0, 1, self.fft_buffer)
        buf("transpose", 0, 0, self.butterfly)                                      # fft_compute_set again (:48)
        buf("transpose", 0, 1, self.fft_buffer)
        img("fft_unpack", 0, 0, self.displacement_map, FORMAT_RGBA16F)              # unpack_set, fft_buffer_set
        img("fft_unpack", 0, 1, self.normal_map, FORMAT_RGBA16F)
        buf("fft_unpack", 1, 0, self.fft_buffer)
        self.groups = {                                                             # :44-49
            "spectrum_compute": (N // 16, N // 16, 1), "spectrum_modulate": (N // 16, N // 16, 1),
            "fft_butterfly": (N // 2 // 64, S, 1), "fft_compute": (1, N, 4), "transpose": (N // 32, N // 32, 4),
            "fft_unpack": (N // 16, N // 16, 1)}
        self._call("fft_butterfly", b"")                                            # :52-54
        self.context = True

    def _call(self, shader: str, push_constant: bytes) -> None:
        g = self.groups[shader]
        rc = lib().ref_dispatch(shader.encode(), push_constant, len(push_constant), g[0], g[1], g[2])
        assert rc == 0, (shader, rc)

    # wave_generator.gd:65-85
    def _update(self, cascade_index: int, parameters) -> None:
        p = parameters[cascade_index]
        cpc = po.create_push_constant
        tl = (float(np.float32(p.tile_length[0])), float(np.float32(p.tile_length[1])))   # Vector2 components are binary32
        if p.should_generate_spectrum:
            alpha = po.JONSWAP_alpha(p.wind_speed, p.fetch_length * 1e3)
            omega = po.JONSWAP_peak_angular_frequency(p.wind_speed, p.fetch_length * 1e3)
            self._call("spectrum_compute", cpc([int(p.spectrum_seed[0]), int(p.spectrum_seed[1]), tl[0], tl[1], alpha, omega,
                                                float(p.wind_speed), po.deg_to_rad(p.wind_direction), po.DEPTH, float(p.swell),
                                                float(p.detail), float(p.spread), int(cascade_index)]))          # :71
            p.should_generate_spectrum = False
        self._call("spectrum_modulate", cpc([tl[0], tl[1], po.DEPTH, float(p.time), int(cascade_index)]))        # :73
        fft_pc = cpc([int(cascade_index)])                                                                      # :76
        self._call("fft_compute", fft_pc)                                                                       # :79
        self._call("transpose", fft_pc)                                                                         # :80
        self._call("fft_compute", fft_pc)                                                                       # :82
        self._call("fft_unpack", cpc([int(cascade_index), float(p.whitecap), float(p.foam_grow_rate),
                                      float(p.foam_decay_rate)]))                                               # :85
