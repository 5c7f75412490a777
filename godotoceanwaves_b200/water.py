"""Water -- host-side mirror of the reference's Water node for the wave path (assets/water/water.gd): owns the generator,
runs the fixed-rate update accumulator (:51-54,75-82), gives every cascade its start time (:32), publishes map_scales
(:102-110) and hands the two layered RGBA16F maps over as the bytes ``RenderingDevice.texture_update(rid, layer, bytes)``
wants (the reference's textures carry TEXTURE_USAGE_CAN_UPDATE_BIT, wave_generator.gd:34-35).

All arithmetic of the scheduler runs in the library (ocean_scheduler_* / ocean_water_frame, C ABI) so that a C# or
GDExtension host binds the same code; this class is the thin Python face of it.  Rendering (meshes, materials, colours,
water.gd:6-18,43-46,65-66) is the engine's business and not mirrored.  The reference draws the spectrum seeds from Godot's
RandomNumberGenerator seeded 1234 (:31,69); callers pass seeds explicitly here."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .native import SchedulerC, check, load_library
from .wave_generator import WaveGenerator


class Water:
    def __init__(self, device: int = 0, map_size: int = 1024, updates_per_second: float = 50.0):
        self.device = device
        self._map_size = int(map_size)                     # water.gd:38 (default 1024)
        self._sched = SchedulerC()
        check(load_library().ocean_scheduler_init(C.byref(self._sched), float(updates_per_second)))
        self._parameters = []
        self.wave_generator: WaveGenerator | None = None

    # ---- water.gd:22-35
    @property
    def parameters(self):
        return self._parameters

    @parameters.setter
    def parameters(self, value):
        lib = load_library()
        for i, p in enumerate(value):
            p.time = lib.ocean_water_default_time(i)       # :32  120.0 + PI*i
        self._parameters = list(value)
        self._setup_wave_generator()

    # ---- water.gd:38-41
    @property
    def map_size(self) -> int:
        return self._map_size

    @map_size.setter
    def map_size(self, value: int) -> None:
        self._map_size = int(value)
        self._setup_wave_generator()

    # ---- water.gd:51-54
    @property
    def updates_per_second(self) -> float:
        return self._sched.updates_per_second

    @updates_per_second.setter
    def updates_per_second(self, value: float) -> None:
        check(load_library().ocean_scheduler_set_rate(C.byref(self._sched), float(value)))

    @property
    def time(self) -> float:
        return self._sched.time

    @property
    def next_update_time(self) -> float:
        return self._sched.next_update_time

    # ---- water.gd:84-100
    def _setup_wave_generator(self) -> None:
        if len(self._parameters) <= 0:
            return
        for p in self._parameters:
            p.should_generate_spectrum = True              # :86-87
        if self.wave_generator is not None:
            self.wave_generator.free()
        self.wave_generator = WaveGenerator(device=self.device)
        self.wave_generator.map_size = self._map_size
        self.wave_generator.init_gpu(max(2, len(self._parameters)))     # :91

    # ---- water.gd:102-110
    def map_scales(self) -> np.ndarray:
        return WaveGenerator.map_scales(self._parameters)

    # ---- one rendered frame: Water._process (:75-82) then the child generator's _process (wave_generator.gd:56-63)
    def process(self, delta: float) -> bool:
        """Returns True when this frame started a new wave update."""
        if self.wave_generator is None:
            self._setup_wave_generator()
        g = self.wave_generator
        arr = g._marshal(self._parameters)
        did = C.c_int(0)
        check(load_library().ocean_water_frame(g.context, C.byref(self._sched), float(delta), arr, len(self._parameters), C.byref(did)))
        g._readback(self._parameters)
        g.pass_parameters = self._parameters
        return bool(did.value)

    # ---- the hand-off: bytes per layer for RenderingDevice.texture_update
    def layer_bytes(self):
        """[(displacement_bytes, normal_bytes)] per cascade, each map_size*map_size*8 bytes of RGBA16F."""
        d, n = self.wave_generator.maps_to_host(0, len(self._parameters))
        return [(d[i].tobytes(), n[i].tobytes()) for i in range(len(self._parameters))]

    def free(self) -> None:                                # water.gd:116-119
        if self.wave_generator is not None:
            self.wave_generator.free()
            self.wave_generator = None
