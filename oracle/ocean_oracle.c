/*
 * oracle/ocean_oracle.c -- TEST INFRASTRUCTURE. CPU restatement of the reference's
 * wave pipeline (2Retr0/GodotOceanWaves).  NOT part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library, and only as the checker / the CPU baseline.
 *
 * PARITY UNPINNED BY THE REFERENCE: the reference ships no tests, golden vectors or
 * runnable headless build (no Godot / Vulkan / glslang in this image), so nothing in
 * /root/reference pins results.  This oracle is pinned instead by (a) the anchors
 * derived from the shader text in SURVEY.md section 4 (tests/test_oracle_pins.py),
 * (b) an independent float64 numpy model that uses numpy.fft (oracle/numpy_model.py),
 * (c) glibc libm for the transcendental functions (math mode 1).
 *
 * Every function below cites the reference lines (relative to /root/reference) it
 * restates.  Arithmetic follows the GLSL text operation for operation
 * (SURVEY.md appendix D), all in binary32 unless the shader says otherwise.
 *
 * Numeric policy (the reference leaves these to the Vulkan driver; fixed here):
 *   1. literals are binary32; constant sub-expressions are evaluated in binary32.
 *   2. + - * / sqrt are IEEE-754 round-to-nearest-even.
 *   3. transcendentals: math mode 0 = DETMATH (oracle/detmath.h, default; the mode
 *      the bit-exact claims are made against), math mode 1 = glibc binary64 libm
 *      rounded to binary32 (cross-check).
 *   4. contraction: contract mode 1 = FMA (default; every mul_complex is
 *      (fma(ax,bx,-(ay*by)), fma(ax,by,ay*bx)); the Jacobian is fma(a,b,-(c*c)); the
 *      foam accumulate is fma(factor,grow,foam)), contract mode 0 = STRICT (no
 *      contraction anywhere).  Nothing else is ever contracted.
 *   5. float -> half is round-to-nearest-even; foam state lives in half.
 *   6. the spectrum texture is RGBA32F (assets/water/wave_generator.gd:31).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC).
 */
#include "detmath.h"
#include <stdlib.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define PI_F 3.141592653589793f   /* GLSL "#define PI (3.141592653589793)" as binary32 = 0x40490FDB */
#define G_F 9.81f
#define NUM_SPECTRA 4

static int g_math_mode = 0;     /* 0 = DETMATH, 1 = libm */
static int g_contract = 1;      /* 1 = FMA, 0 = STRICT   */

void oracle_set_modes(int math_mode, int contract_mode) { g_math_mode = math_mode; g_contract = contract_mode; }
int oracle_get_math_mode(void) { return g_math_mode; }
int oracle_get_contract_mode(void) { return g_contract; }
int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void oracle_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ---------- binary32 transcendental wrappers (policy item 3) ---------- */
static inline float t_cos(float x) { if (g_math_mode) return (float)cos((double)x); double s, c; dm_sincos((double)x, &s, &c); return (float)c; }
static inline float t_sin(float x) { if (g_math_mode) return (float)sin((double)x); double s, c; dm_sincos((double)x, &s, &c); return (float)s; }
static inline float t_exp(float x) { return (float)(g_math_mode ? exp((double)x) : dm_exp((double)x)); }
static inline float t_log(float x) { return (float)(g_math_mode ? log((double)x) : dm_log((double)x)); }
static inline float t_pow(float x, float y) { return (float)(g_math_mode ? pow((double)x, (double)y) : dm_pow((double)x, (double)y)); }
static inline float t_tanh(float x) { return (float)(g_math_mode ? tanh((double)x) : dm_tanh((double)x)); }
static inline float t_atan2(float y, float x) {
    if (g_math_mode) return (x == 0.0f && y == 0.0f) ? 0.0f : (float)atan2((double)y, (double)x);
    return (float)dm_atan2((double)y, (double)x);
}
static inline float t_inversesqrt(float x) { return (float)(1.0 / sqrt((double)x)); }

float oracle_cosf(float x) { return t_cos(x); }
float oracle_sinf(float x) { return t_sin(x); }
float oracle_expf(float x) { return t_exp(x); }
float oracle_logf(float x) { return t_log(x); }
float oracle_powf(float x, float y) { return t_pow(x, y); }
float oracle_tanhf(float x) { return t_tanh(x); }
float oracle_atan2f(float y, float x) { return t_atan2(y, x); }

/* ---------- complex helpers ---------- */
typedef struct { float x, y; } vec2;

/* mul_complex: spectrum_modulate.glsl:37-39, fft_compute.glsl:29-31 (policy item 4) */
static inline vec2 mul_complex(vec2 a, vec2 b) {
    vec2 r;
    if (g_contract) {
        r.x = fmaf(a.x, b.x, -(a.y * b.y));
        r.y = fmaf(a.x, b.y, a.y * b.x);
    } else {
        float p0 = a.x * b.x, p1 = a.y * b.y, p2 = a.x * b.y, p3 = a.y * b.x;
        r.x = p0 - p1;
        r.y = p2 + p3;
    }
    return r;
}

/* ---------- float <-> half, round-to-nearest-even (policy item 5) ---------- */
uint16_t oracle_float_to_half(float f) {
    uint32_t x; memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u;
    uint32_t ax = x & 0x7fffffffu;
    if (ax >= 0x7f800000u) return (uint16_t)(sign | 0x7c00u | ((ax > 0x7f800000u) ? 0x200u : 0u));
    if (ax >= 0x477ff000u) return (uint16_t)(sign | 0x7c00u);         /* rounds to >= 65520 -> inf */
    if (ax < 0x33000001u) return (uint16_t)sign;                       /* <= 2^-25 -> 0 (tie to even) */
    int32_t e = (int32_t)(ax >> 23) - 127;
    uint32_t m = (ax & 0x7fffffu) | 0x800000u;
    uint32_t shift, half_e;
    if (e < -14) { shift = (uint32_t)(13 + (-14 - e)); half_e = 0; } else { shift = 13; half_e = (uint32_t)(e + 15); }
    uint32_t q = m >> shift;
    uint32_t rem = m & ((1u << shift) - 1u);
    uint32_t halfway = 1u << (shift - 1);
    if (rem > halfway || (rem == halfway && (q & 1u))) q++;
    uint32_t h;
    if (half_e == 0) h = q;                       /* denormal (may carry into exponent 1: correct) */
    else h = ((half_e - 1) << 10) + q;            /* q includes the implicit bit (0x400) */
    return (uint16_t)(sign | h);
}
float oracle_half_to_float(uint16_t h) {
    uint32_t sign = ((uint32_t)h & 0x8000u) << 16;
    uint32_t e = (h >> 10) & 0x1f, m = h & 0x3ffu, x;
    if (e == 0) {
        if (m == 0) x = sign;
        else { float v = (float)m * 0x1p-24f; memcpy(&x, &v, 4); x |= sign; }
    } else if (e == 31) x = sign | 0x7f800000u | (m << 13);
    else x = sign | ((e + 112) << 23) | (m << 13);
    float f; memcpy(&f, &x, 4); return f;
}

/* =====================================================================
 * Push-constant blocks (assets/render_context.gd:122-135 packs ints as s32 and
 * floats as binary32; layouts per SURVEY.md 8a row a3)
 * ===================================================================== */
typedef struct {            /* spectrum_compute.glsl:18-30, packed at wave_generator.gd:71 */
    int32_t seed[2];
    float tile_length[2];
    float alpha, peak_frequency, wind_speed, angle, depth, swell, detail, spread;
    uint32_t cascade_index;
} pc_spectrum_compute;
typedef struct {            /* spectrum_modulate.glsl:24-29, wave_generator.gd:73 */
    float tile_length[2];
    float depth, time;
    uint32_t cascade_index;
} pc_spectrum_modulate;
typedef struct {            /* fft_unpack.glsl:20-25, wave_generator.gd:85 */
    uint32_t cascade_index;
    float whitecap, foam_grow_rate, foam_decay_rate;
} pc_fft_unpack;

/* =====================================================================
 * spectrum_compute.glsl
 * ===================================================================== */
/* hash: spectrum_compute.glsl:34-41 */
void oracle_hash(uint32_t x, uint32_t y, float out[2]) {
    uint32_t h32 = y + 374761393U + x * 3266489917U;
    h32 = 2246822519U * (h32 ^ (h32 >> 15));
    h32 = 3266489917U * (h32 ^ (h32 >> 13));
    uint32_t n = h32 ^ (h32 >> 16);
    uint32_t rz0 = n, rz1 = n * 48271U;
    const float denom = (float)0x7FFFFFFF;                      /* = 2147483648.0f */
    out[0] = (float)((rz0 >> 1) & 0x7FFFFFFFU) / denom;
    out[1] = (float)((rz1 >> 1) & 0x7FFFFFFFU) / denom;
}
uint32_t oracle_hash_n(uint32_t x, uint32_t y) {
    uint32_t h32 = y + 374761393U + x * 3266489917U;
    h32 = 2246822519U * (h32 ^ (h32 >> 15));
    h32 = 3266489917U * (h32 ^ (h32 >> 13));
    return h32 ^ (h32 >> 16);
}

/* gaussian: spectrum_compute.glsl:44-49 */
static inline vec2 gaussian(const float u[2]) {
    float r = sqrtf(-2.0f * t_log(u[0]));
    float theta = (2.0f * PI_F) * u[1];
    vec2 g = { r * t_cos(theta), r * t_sin(theta) };
    return g;
}

/* dispersion_relation: spectrum_compute.glsl:58-66 */
static inline void dispersion_relation2(float k, float depth, float *w, float *dw) {
    float a = k * depth;
    float b = t_tanh(a);
    float disp = sqrtf(G_F * k * b);
    float d_disp = (0.5f * G_F) * (b + a * (1.0f - b * b)) / disp;
    *w = disp; *dw = d_disp;
}

/* longuet_higgins_normalization: spectrum_compute.glsl:69-73 */
static inline float lh_normalization(float s) {
    float a = sqrtf(s);
    return (s < 0.4f) ? (0.5f / PI_F) + s * (0.220636f + s * (-0.109f + s * 0.090f))
                      : t_inversesqrt(PI_F) * (a * 0.5f + (1.0f / a) * 0.0625f);
}
/* longuet_higgins_function: spectrum_compute.glsl:76-78 */
static inline float lh_function(float s, float theta) {
    return lh_normalization(s) * t_pow(fabsf(t_cos(theta * 0.5f)), 2.0f * s);
}
/* hasselmann_directional_spread: spectrum_compute.glsl:81-86 */
static inline float hasselmann(float w, float w_p, float wind_speed, float theta, float swell, float angle) {
    float p = w / w_p;
    float s = (w <= w_p) ? 6.97f * t_pow(fabsf(p), 4.06f)
                         : 9.77f * t_pow(fabsf(p), -2.33f - 1.45f * (wind_speed * w_p / G_F - 1.17f));
    float s_xi = 16.0f * t_tanh(w_p / w) * swell * swell;
    return lh_function(s + s_xi, theta - angle);
}
/* TMA_spectrum: spectrum_compute.glsl:89-101 */
static inline float tma_spectrum(float w, float w_p, float alpha, float depth) {
    const float beta = 1.25f, gamma = 3.3f;
    float sigma = (w <= w_p) ? 0.07f : 0.09f;
    float r = t_exp(-(w - w_p) * (w - w_p) / (2.0f * sigma * sigma * w_p * w_p));
    float jonswap = (alpha * G_F * G_F) / t_pow(w, 5.0f) * t_exp(-beta * t_pow(w_p / w, 4.0f)) * t_pow(gamma, r);
    float w_h = fminf(w * sqrtf(depth / G_F), 2.0f);
    float kit = (w_h <= 1.0f) ? 0.5f * w_h * w_h : 1.0f - 0.5f * (2.0f - w_h) * (2.0f - w_h);
    return jonswap * kit;
}

/* deterministic part of get_spectrum_amplitude: sqrt(2*s*d*w_norm), spectrum_compute.glsl:103-114 */
float oracle_amplitude_factor(int idx, int idy, int map_size, const pc_spectrum_compute *pc) {
    float dkx = (2.0f * PI_F) / pc->tile_length[0], dky = (2.0f * PI_F) / pc->tile_length[1];
    float half = (float)map_size * 0.5f;
    float kx = ((float)idx - half) * dkx, ky = ((float)idy - half) * dky;
    float k = sqrtf(kx * kx + ky * ky) + 1e-6f;
    float theta = t_atan2(kx, ky);                                  /* atan(k_vec.x, k_vec.y) */
    float w, dw;
    dispersion_relation2(k, pc->depth, &w, &dw);
    float w_norm = dw / k * dkx * dky;
    float s = tma_spectrum(w, pc->peak_frequency, pc->alpha, pc->depth);
    float D = hasselmann(w, pc->peak_frequency, pc->wind_speed, theta, pc->swell, pc->angle);
    float a = 1.0f - pc->spread;
    float mixv = (0.5f / PI_F) * (1.0f - a) + D * a;                 /* mix(x,y,a) = x*(1-a) + y*a */
    float d = mixv * t_exp(-(1.0f - pc->detail) * (1.0f - pc->detail) * k * k);
    return sqrtf(2.0f * s * d * w_norm);
}
/* get_spectrum_amplitude: spectrum_compute.glsl:103-115 */
static inline vec2 get_spectrum_amplitude(int idx, int idy, int map_size, const pc_spectrum_compute *pc) {
    float f = oracle_amplitude_factor(idx, idy, map_size, pc);
    float u[2];
    oracle_hash((uint32_t)(idx + pc->seed[0]), (uint32_t)(idy + pc->seed[1]), u);   /* uvec2(id + seed) wraps */
    vec2 g = gaussian(u);
    vec2 r = { g.x * f, g.y * f };
    return r;
}
/* main: spectrum_compute.glsl:117-125.  spectrum: N*N*4 floats, texel (x,y) at (y*N+x)*4 */
void oracle_spectrum_compute(float *spectrum, int map_size, const pc_spectrum_compute *pc) {
    const int N = map_size;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < N; ++y)
        for (int x = 0; x < N; ++x) {
            /* id1 = ivec2(mod(-id0, dims)): x - y*floor(x/y) in binary32, exact for these integers */
            float fx = -(float)x, fy = -(float)y, fn = (float)N;
            int x1 = (int)(fx - fn * floorf(fx / fn));
            int y1 = (int)(fy - fn * floorf(fy / fn));
            vec2 a0 = get_spectrum_amplitude(x, y, N, pc);
            vec2 a1 = get_spectrum_amplitude(x1, y1, N, pc);
            float *t = spectrum + ((size_t)y * N + x) * 4;
            t[0] = a0.x; t[1] = a0.y; t[2] = a1.x; t[3] = -a1.y;       /* conj_complex, :52-54 */
        }
}

/* =====================================================================
 * spectrum_modulate.glsl:52-90.  fft_buffer holds 2 halves x 4 layers x N*N vec2
 * (index in vec2 units: half*4*N*N + layer*N*N + y*N + x; one cascade)
 * ===================================================================== */
void oracle_spectrum_modulate(const float *spectrum, float *fft_buffer, int map_size, const pc_spectrum_modulate *pc) {
    const int N = map_size;
    const size_t NN = (size_t)N * N;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < N; ++y)
        for (int x = 0; x < N; ++x) {
            float half = (float)N * 0.5f;
            float kvx = ((float)x - half) * 2.0f * PI_F / pc->tile_length[0];     /* :59 */
            float kvy = ((float)y - half) * 2.0f * PI_F / pc->tile_length[1];
            float k = sqrtf(kvx * kvx + kvy * kvy) + 1e-6f;                         /* :60 */
            float kux = kvx / k, kuy = kvy / k;                                     /* :61 */
            const float *h0 = spectrum + ((size_t)y * N + x) * 4;                   /* :64 */
            float dispersion = sqrtf(G_F * k * t_tanh(k * pc->depth)) * pc->time;   /* :49,65 */
            vec2 m = { t_cos(dispersion), t_sin(dispersion) };                      /* :66 */
            vec2 mc = { m.x, m.y * -1.0f };                                         /* conj: :42-45 */
            vec2 h0a = { h0[0], h0[1] }, h0b = { h0[2], h0[3] };
            vec2 pa = mul_complex(h0a, m), pb = mul_complex(h0b, mc);
            vec2 h = { pa.x + pb.x, pa.y + pb.y };                                  /* :68 */
            vec2 hi = { -h.y, h.x };                                                /* :69 */
            vec2 hx = { hi.x * kuy, hi.y * kuy };                                   /* :72 */
            vec2 hy = h;                                                            /* :73 */
            vec2 hz = { hi.x * kux, hi.y * kux };                                   /* :74 */
            vec2 dhy_dx = { hi.x * kvy, hi.y * kvy };                               /* :78 */
            vec2 dhy_dz = { hi.x * kvx, hi.y * kvx };                               /* :79 */
            vec2 dhx_dx = { -h.x * kvy * kuy, -h.y * kvy * kuy };                   /* :80 */
            vec2 dhz_dz = { -h.x * kvx * kux, -h.y * kvx * kux };                   /* :81 */
            vec2 dhz_dx = { -h.x * kvy * kux, -h.y * kvy * kux };                   /* :82 */
            size_t o = (size_t)y * N + x;
            float *d = fft_buffer;
            d[(0 * NN + o) * 2 + 0] = hx.x - hy.y;         d[(0 * NN + o) * 2 + 1] = hx.y + hy.x;          /* :86 */
            d[(1 * NN + o) * 2 + 0] = hz.x - dhy_dx.y;     d[(1 * NN + o) * 2 + 1] = hz.y + dhy_dx.x;      /* :87 */
            d[(2 * NN + o) * 2 + 0] = dhy_dz.x - dhx_dx.y; d[(2 * NN + o) * 2 + 1] = dhy_dz.y + dhx_dx.x;  /* :88 */
            d[(3 * NN + o) * 2 + 0] = dhz_dz.x - dhz_dx.y; d[(3 * NN + o) * 2 + 1] = dhz_dz.y + dhz_dx.x;  /* :89 */
        }
}

/* =====================================================================
 * fft_butterfly.glsl:18-34.  butterfly: S*N vec4 = (bitcast r0, bitcast r1, tw.re, tw.im)
 * ===================================================================== */
void oracle_fft_butterfly(float *butterfly, int map_size) {
    const uint32_t N = (uint32_t)map_size;
    uint32_t S = 0; while ((1u << S) < N) ++S;
    for (uint32_t stage = 0; stage < S; ++stage)
        for (uint32_t col = 0; col < N / 2; ++col) {
            uint32_t stride = 1u << stage, mid = N >> (stage + 1);
            uint32_t i = col >> stage, j = col % stride;
            float ang = PI_F / (float)stride * (float)j;               /* :27 */
            float twx = t_cos(ang), twy = t_sin(ang);
            uint32_t r0 = stride * (i + 0) + j, r1 = stride * (i + mid) + j;
            uint32_t w0 = stride * (2 * i + 0) + j, w1 = stride * (2 * i + 1) + j;
            float f0, f1; memcpy(&f0, &r0, 4); memcpy(&f1, &r1, 4);  /* uintBitsToFloat :31 */
            float *b0 = butterfly + ((size_t)stage * N + w0) * 4, *b1 = butterfly + ((size_t)stage * N + w1) * 4;
            b0[0] = f0; b0[1] = f1; b0[2] = twx;  b0[3] = twy;          /* :33 */
            b1[0] = f0; b1[1] = f1; b1[2] = -twx; b1[3] = -twy;         /* :34 */
        }
}

/* =====================================================================
 * fft_compute.glsl:37-60: row-wise radix-2 Stockham, half 0 -> half 1 (one cascade)
 * ===================================================================== */
void oracle_fft_compute(const float *butterfly, float *fft_buffer, int map_size) {
    const int N = map_size;
    const size_t NN = (size_t)N * N;
    int S = 0; while ((1 << S) < N) ++S;
#pragma omp parallel for schedule(static) collapse(2)
    for (int spectrum = 0; spectrum < NUM_SPECTRA; ++spectrum)
        for (int row = 0; row < N; ++row) {
            vec2 row_shared[2][1024];
            const float *in = fft_buffer + ((size_t)spectrum * NN + (size_t)row * N) * 2;
            float *out = fft_buffer + ((NUM_SPECTRA + (size_t)spectrum) * NN + (size_t)row * N) * 2;
            for (int col = 0; col < N; ++col) { row_shared[0][col].x = in[2 * col]; row_shared[0][col].y = in[2 * col + 1]; }
            for (int stage = 0; stage < S; ++stage) {
                int rd = stage % 2, wr = (stage + 1) % 2;
                for (int col = 0; col < N; ++col) {
                    const float *bd = butterfly + ((size_t)stage * N + col) * 4;
                    uint32_t r0, r1; memcpy(&r0, &bd[0], 4); memcpy(&r1, &bd[1], 4);
                    vec2 tw = { bd[2], bd[3] };
                    vec2 upper = row_shared[rd][r0], lower = row_shared[rd][r1];
                    vec2 p = mul_complex(lower, tw);
                    row_shared[wr][col].x = upper.x + p.x;              /* :57 */
                    row_shared[wr][col].y = upper.y + p.y;
                }
            }
            for (int col = 0; col < N; ++col) { out[2 * col] = row_shared[S % 2][col].x; out[2 * col + 1] = row_shared[S % 2][col].y; }
        }
}

/* transpose.glsl:29-40: half 1 -> half 0, out[x][y] = in[y][x] per layer */
void oracle_transpose(float *fft_buffer, int map_size) {
    const int N = map_size;
    const size_t NN = (size_t)N * N;
#pragma omp parallel for schedule(static) collapse(2)
    for (int spectrum = 0; spectrum < NUM_SPECTRA; ++spectrum)
        for (int y = 0; y < N; ++y)
            for (int x = 0; x < N; ++x) {
                const float *in = fft_buffer + ((NUM_SPECTRA + (size_t)spectrum) * NN + (size_t)y * N + x) * 2;
                float *out = fft_buffer + ((size_t)spectrum * NN + (size_t)x * N + y) * 2;
                out[0] = in[0]; out[1] = in[1];
            }
}

/* =====================================================================
 * fft_unpack.glsl:33-70.  displacement/normal: N*N*4 halves (RGBA16F, one layer).
 * normal_map.a is the persistent foam state.  disp_f32/normal_f32 (optional, may be
 * NULL) receive the binary32 values before the half conversion.
 * ===================================================================== */
void oracle_fft_unpack(const float *fft_buffer, uint16_t *displacement, uint16_t *normal,
                       float *disp_f32, float *normal_f32, int map_size, const pc_fft_unpack *pc) {
    const int N = map_size;
    const size_t NN = (size_t)N * N;
    const float *half1 = fft_buffer + NUM_SPECTRA * NN * 2;
    const float decay = t_exp(-pc->foam_decay_rate);                    /* :62 (uniform across texels) */
#pragma omp parallel for schedule(static)
    for (int y = 0; y < N; ++y)
        for (int x = 0; x < N; ++x) {
            size_t o = (size_t)y * N + x;
            float sign_shift = (float)(-2 * ((x & 1) ^ (y & 1)) + 1);   /* :38 */
            const float *l0 = half1 + (0 * NN + o) * 2, *l1 = half1 + (1 * NN + o) * 2;
            const float *l2 = half1 + (2 * NN + o) * 2, *l3 = half1 + (3 * NN + o) * 2;
            float hx = l0[0], hy = l0[1], hz = l1[0];                   /* :47-49 */
            float d0 = hx * sign_shift, d1 = hy * sign_shift, d2 = hz * sign_shift, d3 = 0.0f * sign_shift; /* :50 */
            float dhy_dx = l1[1] * sign_shift, dhy_dz = l2[0] * sign_shift, dhx_dx = l2[1] * sign_shift;    /* :53-55 */
            float dhz_dz = l3[0] * sign_shift, dhz_dx = l3[1] * sign_shift;                                /* :56-57 */
            float a = 1.0f + dhx_dx, b = 1.0f + dhz_dz, jacobian;
            if (g_contract) jacobian = fmaf(a, b, -(dhz_dx * dhz_dx));  /* :59 */
            else { float ab = a * b, cc = dhz_dx * dhz_dx; jacobian = ab - cc; }
            float jw = jacobian - pc->whitecap;
            float foam_factor = -((jw < 0.0f) ? jw : 0.0f);             /* -min(0, jacobian - whitecap) :60 */
            float foam = oracle_half_to_float(normal[o * 4 + 3]);       /* :61 */
            foam = foam * decay;                                         /* :62 */
            if (g_contract) foam = fmaf(foam_factor, pc->foam_grow_rate, foam);    /* :63 */
            else { float fg = foam_factor * pc->foam_grow_rate; foam = foam + fg; }
            foam = fminf(fmaxf(foam, 0.0f), 1.0f);                       /* :64 */
            float gx = dhy_dx / (1.0f + fabsf(dhx_dx)), gy = dhy_dz / (1.0f + fabsf(dhz_dz));  /* :66 */
            displacement[o * 4 + 0] = oracle_float_to_half(d0); displacement[o * 4 + 1] = oracle_float_to_half(d1);
            displacement[o * 4 + 2] = oracle_float_to_half(d2); displacement[o * 4 + 3] = oracle_float_to_half(d3);
            normal[o * 4 + 0] = oracle_float_to_half(gx); normal[o * 4 + 1] = oracle_float_to_half(gy);   /* :67 */
            normal[o * 4 + 2] = oracle_float_to_half(dhx_dx); normal[o * 4 + 3] = oracle_float_to_half(foam);
            if (disp_f32) { float *p = disp_f32 + o * 4; p[0] = d0; p[1] = d1; p[2] = d2; p[3] = d3; }
            if (normal_f32) { float *p = normal_f32 + o * 4; p[0] = gx; p[1] = gy; p[2] = dhx_dx; p[3] = foam; }
        }
}

/* =====================================================================
 * One cascade update = wave_generator.gd:65-85 (_update) for one layer.
 * spectrum/fft_buffer/displacement/normal point at THIS cascade's storage.
 * ===================================================================== */
void oracle_cascade_update(float *spectrum, float *fft_buffer, const float *butterfly,
                           uint16_t *displacement, uint16_t *normal, float *disp_f32, float *normal_f32,
                           int map_size, int generate_spectrum, const pc_spectrum_compute *pc_gen,
                           const pc_spectrum_modulate *pc_mod, const pc_fft_unpack *pc_unpack) {
    if (generate_spectrum) oracle_spectrum_compute(spectrum, map_size, pc_gen);      /* :68-72 */
    oracle_spectrum_modulate(spectrum, fft_buffer, map_size, pc_mod);                /* :73 */
    oracle_fft_compute(butterfly, fft_buffer, map_size);                              /* :79 */
    oracle_transpose(fft_buffer, map_size);                                           /* :80 */
    oracle_fft_compute(butterfly, fft_buffer, map_size);                              /* :82 */
    oracle_fft_unpack(fft_buffer, displacement, normal, disp_f32, normal_f32, map_size, pc_unpack); /* :85 */
}

/* =====================================================================
 * A batch of independent cascade updates, one OpenMP thread per cascade (the loops inside the stages then run
 * serially: nested parallelism is off).  Same arithmetic as `count` calls of oracle_cascade_update; used by the CPU
 * baseline of bench.py, where a step is many cascades (wave_generator.gd:96-97: cascades are independent).
 * Layer c of every array belongs to cascade c.
 * ===================================================================== */
void oracle_cascade_update_batch(float *spectrum, float *fft_buffer, const float *butterfly,
                                 uint16_t *displacement, uint16_t *normal, int map_size, int count,
                                 const int *generate_spectrum, const pc_spectrum_compute *pc_gen,
                                 const pc_spectrum_modulate *pc_mod, const pc_fft_unpack *pc_unpack) {
    const size_t NN = (size_t)map_size * map_size;
#pragma omp parallel for schedule(dynamic, 1)
    for (int c = 0; c < count; ++c)
        oracle_cascade_update(spectrum + (size_t)c * NN * 4, fft_buffer + (size_t)c * NN * 16, butterfly,
                              displacement + (size_t)c * NN * 4, normal + (size_t)c * NN * 4, NULL, NULL, map_size,
                              generate_spectrum[c], &pc_gen[c], &pc_mod[c], &pc_unpack[c]);
}
