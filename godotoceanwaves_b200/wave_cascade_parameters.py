"""WaveCascadeParameters -- mirror of assets/water/wave_cascade_parameters.gd (a Godot Resource).

Same field names, defaults, clamping and dirty-flag behaviour: every setter except the two
render-side scales raises ``should_generate_spectrum`` (including whitecap / foam_amount,
wave_cascade_parameters.gd:33,35)."""
from __future__ import annotations

from .native import CascadeParamsC

_DIRTYING = frozenset(("tile_length", "wind_speed", "wind_direction", "fetch_length", "swell", "spread", "detail",
                       "whitecap", "foam_amount"))


def _prop(name, conv):
    def fget(self):
        return self._v[name]

    def fset(self, value):
        self._v[name] = conv(value)
        self._version += 1                      # lets WaveGenerator skip re-marshalling untouched objects
        if name in _DIRTYING:
            self._v["should_generate_spectrum"] = True
    return property(fget, fset)


class WaveCascadeParameters:
    __slots__ = ("_v", "_version", "_synced")

    tile_length = _prop("tile_length", lambda v: (float(v[0]), float(v[1])))          # :7
    displacement_scale = _prop("displacement_scale", float)                           # :9
    normal_scale = _prop("normal_scale", float)                                       # :11
    wind_speed = _prop("wind_speed", lambda v: max(0.0001, float(v)))                 # :15-16
    wind_direction = _prop("wind_direction", float)                                   # :17 (degrees)
    fetch_length = _prop("fetch_length", lambda v: max(0.0001, float(v)))             # :20-21 (km)
    swell = _prop("swell", float)                                                     # :22
    spread = _prop("spread", float)                                                   # :25
    detail = _prop("detail", float)                                                   # :28
    whitecap = _prop("whitecap", float)                                               # :32
    foam_amount = _prop("foam_amount", float)                                         # :34
    spectrum_seed = _prop("spectrum_seed", lambda v: (int(v[0]), int(v[1])))          # :37
    should_generate_spectrum = _prop("should_generate_spectrum", bool)                # :38
    time = _prop("time", float)                                                       # :40
    foam_grow_rate = _prop("foam_grow_rate", float)                                   # :41
    foam_decay_rate = _prop("foam_decay_rate", float)                                 # :42

    def __init__(self, tile_length=(50.0, 50.0), displacement_scale=1.0, normal_scale=1.0, wind_speed=20.0,
                 wind_direction=0.0, fetch_length=550.0, swell=0.8, spread=0.2, detail=1.0, whitecap=0.5,
                 foam_amount=5.0, spectrum_seed=(0, 0), time=0.0):
        self._v = {}
        self._version = 0
        self._synced = None                     # the C record that last exchanged the library-mutated fields with this object
        self.spectrum_seed = spectrum_seed
        self.time = time
        self.foam_grow_rate = 0.0
        self.foam_decay_rate = 0.0
        self.tile_length = tile_length
        self.displacement_scale = displacement_scale
        self.normal_scale = normal_scale
        self.wind_speed = wind_speed
        self.wind_direction = wind_direction
        self.fetch_length = fetch_length
        self.swell = swell
        self.spread = spread
        self.detail = detail
        self.whitecap = whitecap
        self.foam_amount = foam_amount
        self.should_generate_spectrum = True

    # -- C ABI marshalling (struct ocean_cascade_params, include/ocean.h)
    def to_c(self, out: CascadeParamsC) -> None:
        v = self._v
        out.tile_length[0], out.tile_length[1] = v["tile_length"]       # Vector2: binary32 components
        out.displacement_scale = v["displacement_scale"]
        out.normal_scale = v["normal_scale"]
        out.wind_speed = v["wind_speed"]
        out.wind_direction = v["wind_direction"]
        out.fetch_length = v["fetch_length"]
        out.swell = v["swell"]
        out.spread = v["spread"]
        out.detail = v["detail"]
        out.whitecap = v["whitecap"]
        out.foam_amount = v["foam_amount"]
        out.spectrum_seed[0], out.spectrum_seed[1] = v["spectrum_seed"]
        out.should_generate_spectrum = 1 if v["should_generate_spectrum"] else 0
        out.time = v["time"]
        out.foam_grow_rate = v["foam_grow_rate"]
        out.foam_decay_rate = v["foam_decay_rate"]

    def from_c(self, src: CascadeParamsC) -> None:
        """Reads back the fields the generator mutates (wave_generator.gd:72,103-106) without touching the version
        counter; remembers WHICH C record they came from, so that a second generator driving the same object re-marshals
        it instead of trusting its own stale copy of time / dirty flag."""
        self._synced = src
        v = self._v
        v["should_generate_spectrum"] = bool(src.should_generate_spectrum)
        v["time"] = src.time
        v["foam_grow_rate"] = src.foam_grow_rate
        v["foam_decay_rate"] = src.foam_decay_rate

    def __repr__(self):
        kv = ", ".join(f"{k}={val!r}" for k, val in self._v.items())
        return f"WaveCascadeParameters({kv})"
