// tools/microbench/pipe_rates.cu -- issue rates of the packed / scalar binary32 instructions the IFFT is made of, per SM, at the
// occupancy the update kernel runs at (and above).  Not part of the product; built and run by tools/microbench/run.sh on the B200.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o pipe_rates pipe_rates.cu && ./pipe_rates
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

constexpr int ACC = 12;     // independent chains per thread
constexpr int ITERS = 4096;

template <int MODE>
__global__ void k(float* out, float seed) {
    u64 a[ACC];
    float s[ACC];
#pragma unroll
    for (int i = 0; i < ACC; ++i) { a[i] = pk(seed + i, seed - i); s[i] = seed * (i + 1); }
    const u64 b = pk(1.0000001f, 0.9999999f), c = pk(1e-7f, -1e-7f);
    const float bs = 1.0000001f, cs = 1e-7f;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ACC; ++i) {
            if (MODE == 0) a[i] = fma2(a[i], b, c);                    // FFMA2, 3 register pairs
            if (MODE == 1) a[i] = add2(a[i], c);                       // FADD2
            if (MODE == 2) a[i] = mul2(a[i], b);                       // FMUL2
            if (MODE == 3) s[i] = __fmaf_rn(s[i], bs, cs);             // scalar FFMA
            if (MODE == 4) s[i] = __fadd_rn(s[i], cs);                 // scalar FADD
            if (MODE == 5) { a[i] = fma2(a[i], b, c); s[i] = __fmaf_rn(s[i], bs, cs); }        // 1 packed + 1 scalar
            if (MODE == 6) { a[i] = add2(a[i], c); a[i] = fma2(a[i], b, c); }                  // butterfly-like mix
            if (MODE == 7) a[i] = fma2(a[i], pk(bs, bs), c);           // FFMA2 with a broadcast scalar operand
        }
    }
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < ACC; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a[i])); acc += lo + hi + s[i]; }
    if (acc == 12345.678f) out[threadIdx.x] = acc;
}

template <int MODE>
void run(const char* name, int per_thread_instr, int ctas_per_sm, float* out) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148 * ctas_per_sm;
    k<MODE><<<grid, 128>>>(out, 1.0f);
    cudaEventRecord(e0);
    k<MODE><<<grid, 128>>>(out, 1.0f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double warp_instr = (double)grid * 4 * ITERS * ACC * per_thread_instr;
    const double cycles = ms * 1e-3 * 1.965e9;
    printf("%-44s %2d warps/SM: %7.3f ms  %.3f warp-instr/clk/SM  (%.3f per SMSP)\n", name, ctas_per_sm * 4, ms, warp_instr / cycles / 148,
           warp_instr / cycles / 148 / 4);
}

int main() {
    float* out;
    cudaMalloc(&out, 4096);
    for (int c : {4, 8, 16}) {
        run<0>("FFMA2 (3 register pairs)", 1, c, out);
        run<7>("FFMA2 (broadcast scalar multiplicand)", 1, c, out);
        run<1>("FADD2", 1, c, out);
        run<2>("FMUL2", 1, c, out);
        run<3>("FFMA scalar", 1, c, out);
        run<4>("FADD scalar", 1, c, out);
        run<5>("FFMA2 + FFMA scalar interleaved", 2, c, out);
        run<6>("FADD2 + FFMA2 interleaved", 2, c, out);
    }
    return 0;
}
