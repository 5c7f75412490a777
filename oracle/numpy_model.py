"""oracle/numpy_model.py -- TEST INFRASTRUCTURE (not part of the product).

An INDEPENDENT float64 numpy model of the reference pipeline, used only to pin the C
oracle's *semantics* (it shares no code with ocean_oracle.c and uses numpy.fft for the
transforms).  It follows the formulas of

  assets/shaders/compute/spectrum_compute.glsl:34-125   (spectrum)
  assets/shaders/compute/spectrum_modulate.glsl:52-90   (time propagation + packing)
  assets/shaders/compute/fft_compute.glsl / transpose.glsl (== (N^2 * ifft2(X))^T)
  assets/shaders/compute/fft_unpack.glsl:33-70          (maps, Jacobian, foam)

in binary64 (except where noted), so it agrees with the binary32 oracle to ~1e-5 of each
field's maximum, not bit for bit.
"""
from __future__ import annotations

import math

import numpy as np

G = 9.81
PI32 = float(np.float32(math.pi))       # the GLSL literal PI as binary32
G32 = float(np.float32(9.81))


def hash_uniforms(ix: np.ndarray, iy: np.ndarray):
    """spectrum_compute.glsl:34-41 on uint32 arrays; returns binary32-valued uniforms."""
    x = ix.astype(np.uint32)
    y = iy.astype(np.uint32)
    with np.errstate(over="ignore"):
        h = y + np.uint32(374761393) + x * np.uint32(3266489917)
        h = np.uint32(2246822519) * (h ^ (h >> np.uint32(15)))
        h = np.uint32(3266489917) * (h ^ (h >> np.uint32(13)))
        n = h ^ (h >> np.uint32(16))
        n2 = n * np.uint32(48271)
    u1 = ((n >> np.uint32(1)) & np.uint32(0x7FFFFFFF)).astype(np.float32) / np.float32(2147483648.0)
    u2 = ((n2 >> np.uint32(1)) & np.uint32(0x7FFFFFFF)).astype(np.float32) / np.float32(2147483648.0)
    return u1.astype(np.float64), u2.astype(np.float64), n


def amplitude_factor(idx, idy, N, tile_length, alpha, w_p, wind_speed, angle, depth, swell, detail, spread):
    """sqrt(2*S*D*w_norm), spectrum_compute.glsl:103-114, float64."""
    idx = np.asarray(idx, np.float64)
    idy = np.asarray(idy, np.float64)
    dkx = 2.0 * PI32 / tile_length[0]
    dky = 2.0 * PI32 / tile_length[1]
    kx = (idx - N * 0.5) * dkx
    ky = (idy - N * 0.5) * dky
    k = np.sqrt(kx * kx + ky * ky) + 1e-6
    theta = np.arctan2(kx, ky)
    a = k * depth
    b = np.tanh(a)
    w = np.sqrt(G32 * k * b)
    dw = 0.5 * G32 * (b + a * (1.0 - b * b)) / w
    w_norm = dw / k * dkx * dky
    # TMA
    sigma = np.where(w <= w_p, 0.07, 0.09)
    r = np.exp(-(w - w_p) ** 2 / (2.0 * sigma * sigma * w_p * w_p))
    with np.errstate(over="ignore", under="ignore", divide="ignore", invalid="ignore"):
        jonswap = (alpha * G32 * G32) / w ** 5 * np.exp(-1.25 * (w_p / w) ** 4) * 3.3 ** r
    w_h = np.minimum(w * math.sqrt(depth / G32), 2.0)
    kit = np.where(w_h <= 1.0, 0.5 * w_h * w_h, 1.0 - 0.5 * (2.0 - w_h) ** 2)
    S = jonswap * kit
    # Hasselmann + Longuet-Higgins
    p = w / w_p
    with np.errstate(over="ignore", under="ignore"):
        s = np.where(w <= w_p, 6.97 * np.abs(p) ** 4.06,
                     9.77 * np.abs(p) ** (-2.33 - 1.45 * (wind_speed * w_p / G32 - 1.17)))
    s_xi = 16.0 * np.tanh(w_p / w) * swell * swell
    ss = s + s_xi
    sa = np.sqrt(ss)
    norm = np.where(ss < 0.4, 0.5 / PI32 + ss * (0.220636 + ss * (-0.109 + ss * 0.090)),
                    (1.0 / math.sqrt(PI32)) * (sa * 0.5 + (1.0 / sa) * 0.0625))
    D = norm * np.abs(np.cos((theta - angle) * 0.5)) ** (2.0 * ss)
    mixa = 1.0 - spread
    d = ((0.5 / PI32) * (1.0 - mixa) + D * mixa) * np.exp(-(1.0 - detail) ** 2 * k * k)
    return np.sqrt(2.0 * S * d * w_norm)


def spectrum(N, seed, tile_length, alpha, w_p, wind_speed, angle, depth, swell, detail, spread):
    """Returns complex h0(k) and conj(h0(-k)) as two (N,N) complex128 arrays [y,x]."""
    ys, xs = np.meshgrid(np.arange(N), np.arange(N), indexing="ij")

    def amp(ix, iy):
        f = amplitude_factor(ix, iy, N, tile_length, alpha, w_p, wind_speed, angle, depth, swell, detail, spread)
        u1, u2, _ = hash_uniforms(ix + seed[0], iy + seed[1])
        with np.errstate(divide="ignore"):
            rr = np.sqrt(-2.0 * np.log(u1))
        th = 2.0 * PI32 * u2
        return (rr * np.cos(th) + 1j * rr * np.sin(th)) * f

    h0 = amp(xs, ys)
    h0m = np.conj(amp((-xs) % N, (-ys) % N))
    return h0, h0m


def modulate(h0, h0m, N, tile_length, depth, time, phase_fp32=True):
    """spectrum_modulate.glsl:52-90 -> 4 packed complex layers [4,y,x]."""
    ys, xs = np.meshgrid(np.arange(N), np.arange(N), indexing="ij")
    if phase_fp32:
        f32 = np.float32
        kvx = ((xs.astype(f32) - f32(N * 0.5)) * f32(2.0) * f32(PI32) / f32(tile_length[0]))
        kvy = ((ys.astype(f32) - f32(N * 0.5)) * f32(2.0) * f32(PI32) / f32(tile_length[1]))
        k = np.sqrt(kvx * kvx + kvy * kvy) + f32(1e-6)
        th = np.tanh((k * f32(depth)).astype(np.float64)).astype(f32)
        phase = (np.sqrt(f32(G32) * k * th) * f32(time)).astype(np.float64)
        kux = (kvx / k).astype(np.float64)
        kuy = (kvy / k).astype(np.float64)
        kvx = kvx.astype(np.float64)
        kvy = kvy.astype(np.float64)
    else:
        kvx = (xs - N * 0.5) * 2.0 * PI32 / tile_length[0]
        kvy = (ys - N * 0.5) * 2.0 * PI32 / tile_length[1]
        k = np.sqrt(kvx * kvx + kvy * kvy) + 1e-6
        phase = np.sqrt(G32 * k * np.tanh(k * depth)) * time
        kux, kuy = kvx / k, kvy / k
    m = np.exp(1j * phase)
    h = h0 * m + h0m * np.conj(m)
    hi = 1j * h
    hx, hy, hz = hi * kuy, h, hi * kux
    dhy_dx, dhy_dz = hi * kvy, hi * kvx
    dhx_dx, dhz_dz, dhz_dx = -h * kvy * kuy, -h * kvx * kux, -h * kvy * kux
    return np.stack([hx + 1j * hy, hz + 1j * dhy_dx, dhy_dz + 1j * dhx_dx, dhz_dz + 1j * dhz_dx])


def ifft_maps(layers):
    """(N^2 * ifft2(X))^T * (-1)^(x+y) for each layer (fft_compute x2 + transpose + unpack sign)."""
    N = layers.shape[-1]
    out = np.fft.ifft2(layers, axes=(-2, -1)) * (N * N)
    out = np.swapaxes(out, -1, -2)
    ys, xs = np.meshgrid(np.arange(N), np.arange(N), indexing="ij")
    sign = 1.0 - 2.0 * ((xs ^ ys) & 1)
    return out * sign


def unpack(fields, foam_prev, whitecap, grow, decay):
    """fft_unpack.glsl:45-68 in float64; fields = ifft_maps(...) [4,y,x] complex."""
    hx, hy = fields[0].real, fields[0].imag
    hz, dhy_dx = fields[1].real, fields[1].imag
    dhy_dz, dhx_dx = fields[2].real, fields[2].imag
    dhz_dz, dhz_dx = fields[3].real, fields[3].imag
    jac = (1.0 + dhx_dx) * (1.0 + dhz_dz) - dhz_dx * dhz_dx
    foam_factor = -np.minimum(0.0, jac - whitecap)
    foam = np.clip(foam_prev * math.exp(-decay) + foam_factor * grow, 0.0, 1.0)
    disp = np.stack([hx, hy, hz, np.zeros_like(hx)], axis=-1)
    normal = np.stack([dhy_dx / (1.0 + np.abs(dhx_dx)), dhy_dz / (1.0 + np.abs(dhz_dz)), dhx_dx, foam], axis=-1)
    return disp, normal, jac
