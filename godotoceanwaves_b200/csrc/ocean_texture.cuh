// ocean_texture.cuh -- texture() of the reference's spatial shaders on the generator's RGBA16F maps: exact-weight bilinear
// filter with REPEAT addressing, binary32, round to nearest, no contraction (oracle/sampling.py is the specification).
// Shared by the map-query op (ocean_sample.cu) and the spray-candidate op (ocean_spray.cu).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace ocean {
namespace {

__device__ __forceinline__ float4 texel(const uint2* __restrict__ layer, int N, int x, int y) {
    const uint2 t = __ldg(&layer[(size_t)y * N + x]);
    const __half2 lo = *reinterpret_cast<const __half2*>(&t.x), hi = *reinterpret_cast<const __half2*>(&t.y);
    const float2 a = __half22float2(lo), b = __half22float2(hi);
    return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }   // GLSL mix
__device__ __forceinline__ float4 mix4(const float4 a, const float4 b, float t) {
    return make_float4(mixf(a.x, b.x, t), mixf(a.y, b.y, t), mixf(a.z, b.z, t), mixf(a.w, b.w, t));
}

// texture(): bilinear, REPEAT.  N is a power of two (128..1024), so the wrap is a mask.
__device__ __forceinline__ float4 texture_bilinear(const uint2* __restrict__ layer, int N, float u, float v) {
    const float n = (float)N;
    const float x = u * n - 0.5f, y = v * n - 0.5f;
    const float x0 = floorf(x), y0 = floorf(y);
    const float fx = x - x0, fy = y - y0;
    const int ix0 = (int)(long long)x0 & (N - 1), iy0 = (int)(long long)y0 & (N - 1);
    const int ix1 = (ix0 + 1) & (N - 1), iy1 = (iy0 + 1) & (N - 1);
    const float4 t00 = texel(layer, N, ix0, iy0), t10 = texel(layer, N, ix1, iy0);
    const float4 t01 = texel(layer, N, ix0, iy1), t11 = texel(layer, N, ix1, iy1);
    return mix4(mix4(t00, t10, fx), mix4(t01, t11, fx), fy);
}

}  // namespace
}  // namespace ocean
