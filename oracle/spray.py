"""CPU restatement of the spray-particle spawn test of the reference.

TEST INFRASTRUCTURE ONLY -- the product path (godotoceanwaves_b200/csrc) never imports or calls this module.

Follows assets/shaders/spatial/sea_spray_particle.gdshader:
  * start()   :45-66  candidate start positions: a t x t grid, t = uint(sqrt(float(num_particles))),
                      coords = (vec2(uvec2(INDEX / t, INDEX % t)) / (float(t) - 1.0) - 0.5) * 10.0, moved by EMISSION_TRANSFORM
  * process() :80-94  gradient = sum_i texture(normals, vec3(START_POS.xz * map_scales[i].xy, i)).xyw
                      normal   = normalize(vec3(-gradient.x, 1.0, -gradient.y)),  foam = gradient.z
                      normal_factor = mix(0.25, 1.0, min((normal.y - 0.92) / (0.99 - 0.92), 1.0))
                      foam_factor   = mix(0.25, 1.0, min((foam - 0.9) / (1.0 - 0.9), 1.0))
                      ACTIVE = normal_factor >= 0.0 && normal_factor <= 1.0 && foam > 0.9
                      SCALE_FACTOR = normal_factor * foam_factor
                      PARTICLE_SCALE = vec3(foam_factor * (float(ACTIVE) + 1e-3)) * vec3(1, normal_factor, 1) * particle_scale
The reference evaluates this for EVERY particle and culls the inactive ones afterwards (README.md:29 "most particles are
culled"); the op returns the active candidates only, in candidate order (a stable stream compaction).

Numeric policy (as oracle/sampling.py): binary32, round to nearest, the shader's operation order, no contraction;
texture() = exact-weight bilinear, REPEAT; normalize(v) = v / sqrt(v.x*v.x + v.y*v.y + v.z*v.z) with IEEE sqrt and division;
constant sub-expressions (0.99 - 0.92, 1.0 - 0.9) are evaluated in binary32.
"""
from __future__ import annotations

import numpy as np

from .sampling import F, _mix, texture_bilinear

RECORD = np.dtype([("index", np.uint32), ("start_x", np.float32), ("start_z", np.float32), ("scale_factor", np.float32),
                   ("particle_scale", np.float32, 3), ("foam", np.float32)])


def spray_grid(num_particles: int, emission_transform=None) -> np.ndarray:
    """sea_spray_particle.gdshader:47,52-54: START_POS.xz of every particle INDEX (float32 [num_particles][2]).
    emission_transform: 3 x 4 (rows x, y, z; columns basis x, basis y, basis z, origin), identity when None;
    position = ((E[:,0]*cx + E[:,1]*0) + E[:,2]*cz) + E[:,3]."""
    t = np.uint32(np.sqrt(F(num_particles)))
    idx = np.arange(num_particles, dtype=np.uint32)
    c = np.stack([(idx // t).astype(np.float32), (idx % t).astype(np.float32)], axis=1)
    c = (c / (F(t) - F(1.0)) - F(0.5)) * F(10.0)
    E = np.eye(3, 4, dtype=np.float32) if emission_transform is None else np.asarray(emission_transform, np.float32).reshape(3, 4)
    cx, cz = c[:, 0], c[:, 1]
    out = np.empty((num_particles, 2), np.float32)
    for k, row in enumerate((0, 2)):
        out[:, k] = ((E[row, 0] * cx + E[row, 1] * F(0.0)) + E[row, 2] * cz) + E[row, 3]
    return out


def spray_candidates(normal: np.ndarray, points_xz: np.ndarray, map_scales: np.ndarray, particle_scale) -> np.ndarray:
    """normal: [C][N][N][4] float16; points_xz: [n][2] START_POS.xz; map_scales: [C][4]; particle_scale: 3 floats.
    Returns the active candidates as RECORD rows in candidate order."""
    pts = np.ascontiguousarray(points_xz, np.float32)
    sc = np.ascontiguousarray(map_scales, np.float32)
    ps = np.asarray(particle_scale, np.float32)
    n = pts.shape[0]
    grad = np.zeros((n, 3), np.float32)
    for i in range(normal.shape[0]):
        t = texture_bilinear(normal[i], pts[:, 0] * sc[i, 0], pts[:, 1] * sc[i, 1])            # :83
        grad = grad + t[:, [0, 1, 3]]
    gx, gy, foam = -grad[:, 0], -grad[:, 1], grad[:, 2]
    with np.errstate(invalid="ignore", over="ignore"):
        ny = F(1.0) / np.sqrt((gx * gx + F(1.0) * F(1.0)) + gy * gy)                            # normalize(...).y  :84
        nf = _mix(F(0.25), F(1.0), np.minimum((ny - F(0.92)) / (F(0.99) - F(0.92)), F(1.0)))    # :86
        ff = _mix(F(0.25), F(1.0), np.minimum((foam - F(0.9)) / (F(1.0) - F(0.9)), F(1.0)))     # :87
        active = (nf >= F(0.0)) & (nf <= F(1.0)) & (foam > F(0.9))                              # :89
    s = ff * (F(1.0) + F(1e-3))                                                                  # :92 (ACTIVE == true)
    out = np.zeros(int(active.sum()), RECORD)
    k = np.nonzero(active)[0]
    out["index"] = k.astype(np.uint32)
    out["start_x"], out["start_z"] = pts[k, 0], pts[k, 1]
    out["scale_factor"] = (nf * ff)[k]                                                           # :90
    out["particle_scale"][:, 0] = ((s * F(1.0)) * ps[0])[k]                                      # :93-94
    out["particle_scale"][:, 1] = ((s * nf) * ps[1])[k]
    out["particle_scale"][:, 2] = ((s * F(1.0)) * ps[2])[k]
    out["foam"] = foam[k]
    return out
