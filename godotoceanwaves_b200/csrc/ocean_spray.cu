// ocean_spray.cu -- foam-driven spray candidates as a stream-compaction op.
//
// Reference: assets/shaders/spatial/sea_spray_particle.gdshader
//   process() :80-94   gradient = sum_i texture(normals, vec3(START_POS.xz * map_scales[i].xy, i)).xyw
//                      normal = normalize(vec3(-gradient.x, 1, -gradient.y)), foam = gradient.z
//                      normal_factor / foam_factor = mix(0.25, 1, min((v - lo) / (hi - lo), 1)),  ACTIVE = factors in range && foam > 0.9
//                      SCALE_FACTOR, PARTICLE_SCALE
// The reference runs this for every particle of the emitter and culls the inactive ones (README.md:29: "most particles are
// culled"); here the candidates are evaluated once and only the active ones are written out, in candidate order:
//   pass 1  one thread per candidate: evaluate, count the active ones per block of 256
//   scan    exclusive prefix sum of the block counts (one block)
//   pass 2  evaluate again (four texel gathers per cascade, L2-resident) and write the record at its final position
//           (block offset + rank inside the block from warp ballots) -- a STABLE compaction, deterministic output.
// Numeric policy: oracle/spray.py is the specification (binary32, shader operation order, no contraction, -fmad=false).
#include "ocean_kernels.cuh"
#include "ocean_texture.cuh"

namespace ocean {

namespace {

struct SprayEval {
    bool active;
    float scale_factor, psx, psy, psz, foam;
};

__device__ __forceinline__ SprayEval spray_eval(const uint2* __restrict__ normal, int N, int C, float2 p, const float4* __restrict__ scales,
                                                float3 particle_scale) {
    float gx = 0.0f, gy = 0.0f, gf = 0.0f;
    for (int c = 0; c < C; ++c) {
        const float4 s = __ldg(&scales[c]);
        const float4 t = texture_bilinear(normal + (size_t)c * N * N, N, p.x * s.x, p.y * s.y);     // :83
        gx = gx + t.x;
        gy = gy + t.y;
        gf = gf + t.w;
    }
    const float nx = -gx, nz = -gy;
    const float ny = __fdiv_rn(1.0f, __fsqrt_rn((nx * nx + 1.0f * 1.0f) + nz * nz));               // normalize(...).y  :84
    const float nf = mixf(0.25f, 1.0f, fminf(__fdiv_rn(ny - 0.92f, 0.99f - 0.92f), 1.0f));          // :86
    const float ff = mixf(0.25f, 1.0f, fminf(__fdiv_rn(gf - 0.9f, 1.0f - 0.9f), 1.0f));             // :87
    SprayEval e;
    e.active = nf >= 0.0f && nf <= 1.0f && gf > 0.9f;                                               // :89
    e.scale_factor = nf * ff;                                                                       // :90
    const float s0 = ff * (1.0f + 1e-3f);                                                           // :92
    e.psx = (s0 * 1.0f) * particle_scale.x;                                                         // :93-94
    e.psy = (s0 * nf) * particle_scale.y;
    e.psz = (s0 * 1.0f) * particle_scale.z;
    e.foam = gf;
    return e;
}

constexpr int kSprayBlock = 256;

__global__ void __launch_bounds__(kSprayBlock) k_spray_count(const uint2* __restrict__ normal, int N, int C, const float2* __restrict__ points, int n,
                                                             const float4* __restrict__ scales, float3 particle_scale, int* __restrict__ block_counts) {
    const int i = blockIdx.x * kSprayBlock + threadIdx.x;
    bool active = false;
    if (i < n) active = spray_eval(normal, N, C, points[i], scales, particle_scale).active;
    const int total = __syncthreads_count(active);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = total;
}

// exclusive scan of `blocks` counts in place; counts[blocks] receives the total
__global__ void __launch_bounds__(1024) k_spray_scan(int* __restrict__ counts, int blocks) {
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < blocks; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = (i < blocks) ? counts[i] : 0;
        int incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, d);
            if ((threadIdx.x & 31) >= d) incl += t;
        }
        if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            int w = s_warp[threadIdx.x];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, w, d);
                if (threadIdx.x >= d) w += t;
            }
            s_warp[threadIdx.x] = w;
        }
        __syncthreads();
        const int warp_off = (threadIdx.x >> 5) ? s_warp[(threadIdx.x >> 5) - 1] : 0;
        const int carry = s_carry;
        if (i < blocks) counts[i] = carry + warp_off + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + warp_off + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) counts[blocks] = s_carry;
}

struct SprayRecord {        // == ocean_spray_record (include/ocean.h), 32 bytes
    uint32_t index;
    float start_x, start_z, scale_factor;
    float particle_scale[3];
    float foam;
};

__global__ void __launch_bounds__(kSprayBlock) k_spray_write(const uint2* __restrict__ normal, int N, int C, const float2* __restrict__ points, int n,
                                                             const float4* __restrict__ scales, float3 particle_scale,
                                                             const int* __restrict__ block_offsets, SprayRecord* __restrict__ out, int max_records) {
    __shared__ int s_warp[kSprayBlock / 32];
    const int i = blockIdx.x * kSprayBlock + threadIdx.x;
    SprayEval e;
    e.active = false;
    float2 p = make_float2(0.f, 0.f);
    if (i < n) {
        p = points[i];
        e = spray_eval(normal, N, C, p, scales, particle_scale);
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, e.active);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_warp[warp] = __popc(ballot);
    __syncthreads();
    int off = block_offsets[blockIdx.x];
    for (int w = 0; w < warp; ++w) off += s_warp[w];
    off += __popc(ballot & ((1u << lane) - 1u));
    if (e.active && off < max_records) {
        SprayRecord r;
        r.index = (uint32_t)i;
        r.start_x = p.x;
        r.start_z = p.y;
        r.scale_factor = e.scale_factor;
        r.particle_scale[0] = e.psx;
        r.particle_scale[1] = e.psy;
        r.particle_scale[2] = e.psz;
        r.foam = e.foam;
        out[off] = r;
    }
}

}  // namespace

int spray_blocks(int n) { return (n + kSprayBlock - 1) / kSprayBlock; }

// counts_dev: [spray_blocks(n) + 1] ints of scratch; on completion counts_dev[spray_blocks(n)] = number of active
// candidates (which may exceed max_records: the records beyond it are dropped, the count is not clamped).
cudaError_t launch_extract_spray(const DeviceBuffers& b, int num_cascades, const float2* points_dev, int n, const float4* scales_dev,
                                 float3 particle_scale, int* counts_dev, void* records_dev, int max_records, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    const int blocks = spray_blocks(n);
    k_spray_count<<<blocks, kSprayBlock, 0, stream>>>(b.normal, b.map_size, num_cascades, points_dev, n, scales_dev, particle_scale, counts_dev);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    k_spray_scan<<<1, 1024, 0, stream>>>(counts_dev, blocks);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    k_spray_write<<<blocks, kSprayBlock, 0, stream>>>(b.normal, b.map_size, num_cascades, points_dev, n, scales_dev, particle_scale, counts_dev,
                                                      static_cast<SprayRecord*>(records_dev), max_records);
    return cudaGetLastError();
}

}  // namespace ocean
