// oracle/ref/glsl_shim.hpp -- TEST INFRASTRUCTURE (recipe of oracle/_ref; not part of the product).
//
// A small GLSL-on-CPU execution model, just large enough to compile the reference's six compute shaders
// (/root/reference/assets/shaders/compute/*.glsl, pre-processed lexically by glsl2cpp.py) with g++ and run
// them invocation by invocation exactly as a Vulkan device would dispatch them:
//   * vector types with the swizzles, constructors and operators those shaders use;
//   * gl_GlobalInvocationID & co., storage buffers, RGBA32F / RGBA16F storage images, push constants;
//   * barrier() -- every invocation of a workgroup is a fiber (own stack); barrier() switches to the next one, so
//     the shader body runs unmodified, `shared` arrays included;
//   * the numeric policy the GLSL text leaves to the driver (see oracle/ocean_oracle.c header; same policy here):
//       - `float` is IEEE binary32, every operation rounded separately (the file is compiled with -ffp-contract=off);
//       - contraction mode 1 ("FMA mode"): an expression of the form x*y +/- z*w becomes fma(x, y, +/-(z*w)) and
//         v += x*y becomes fma(x, y, v), nothing else is contracted; mode 0 ("STRICT"): no contraction at all.
//         Implemented by the product proxy `Prod` below -- the shader source decides where such expressions occur;
//       - transcendentals: mode 0 = DETMATH (oracle/detmath.h), mode 1 = glibc binary64 libm rounded to binary32;
//       - float -> half conversion on RGBA16F image stores: round-to-nearest-even (the compiler's _Float16).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../detmath.h"

namespace glsl {

extern int g_contract;   // 1 = FMA mode, 0 = STRICT
extern int g_math;       // 0 = DETMATH, 1 = libm

// ---------------------------------------------------------------------------------------------------------------
// binary32 scalar with the contraction policy
// ---------------------------------------------------------------------------------------------------------------
struct Prod;
struct Real {
    float v;
    Real() = default;
    Real(float x) : v(x) {}
    Real(double x) : v((float)x) {}
    Real(int x) : v((float)x) {}
    Real(unsigned x) : v((float)x) {}
    inline Real(const Prod& p);
    explicit operator int() const { return (int)v; }
    explicit operator unsigned() const { return (unsigned)v; }
    Real& operator+=(Real o) { v = v + o.v; return *this; }
    inline Real& operator+=(const Prod& p);
    Real& operator-=(Real o) { v = v - o.v; return *this; }
    Real& operator*=(Real o) { v = v * o.v; return *this; }
    Real& operator/=(Real o) { v = v / o.v; return *this; }
};
// an unrounded-yet product a*b; converting it to Real rounds it
struct Prod {
    float a, b;
    float value() const { return a * b; }
};
inline Real::Real(const Prod& p) : v(p.a * p.b) {}
inline Real& Real::operator+=(const Prod& p) {
    v = g_contract ? fmaf(p.a, p.b, v) : v + p.a * p.b;
    return *this;
}
inline Prod operator*(Real a, Real b) { return Prod{a.v, b.v}; }
inline Prod operator*(const Prod& a, Real b) { return Prod{a.value(), b.v}; }
inline Prod operator*(Real a, const Prod& b) { return Prod{a.v, b.value()}; }
inline Prod operator*(const Prod& a, const Prod& b) { return Prod{a.value(), b.value()}; }
inline Real operator+(Real a, Real b) { return Real(a.v + b.v); }
inline Real operator-(Real a, Real b) { return Real(a.v - b.v); }
inline Real operator/(Real a, Real b) { return Real(a.v / b.v); }
inline Real operator-(Real a) { return Real(-a.v); }
inline Prod operator-(const Prod& p) { return Prod{-p.a, p.b}; }
inline Real operator+(const Prod& a, const Prod& b) { return Real(g_contract ? fmaf(a.a, a.b, b.a * b.b) : a.a * a.b + b.a * b.b); }
inline Real operator-(const Prod& a, const Prod& b) { return Real(g_contract ? fmaf(a.a, a.b, -(b.a * b.b)) : a.a * a.b - b.a * b.b); }
inline bool operator<(Real a, Real b) { return a.v < b.v; }
inline bool operator<=(Real a, Real b) { return a.v <= b.v; }
inline bool operator>(Real a, Real b) { return a.v > b.v; }
inline bool operator>=(Real a, Real b) { return a.v >= b.v; }
inline bool operator==(Real a, Real b) { return a.v == b.v; }
inline bool operator!=(Real a, Real b) { return a.v != b.v; }

// transcendentals and the other scalar built-ins (policy: see header)
inline Real cos(Real x) { if (g_math) return Real(::cos((double)x.v)); double s, c; dm_sincos((double)x.v, &s, &c); return Real(c); }
inline Real sin(Real x) { if (g_math) return Real(::sin((double)x.v)); double s, c; dm_sincos((double)x.v, &s, &c); return Real(s); }
inline Real exp(Real x) { return Real(g_math ? ::exp((double)x.v) : dm_exp((double)x.v)); }
inline Real log(Real x) { return Real(g_math ? ::log((double)x.v) : dm_log((double)x.v)); }
inline Real pow(Real x, Real y) { return Real(g_math ? ::pow((double)x.v, (double)y.v) : dm_pow((double)x.v, (double)y.v)); }
inline Real tanh(Real x) { return Real(g_math ? ::tanh((double)x.v) : dm_tanh((double)x.v)); }
inline Real atan(Real y, Real x) {
    if (g_math) return (x.v == 0.0f && y.v == 0.0f) ? Real(0.0f) : Real(::atan2((double)y.v, (double)x.v));
    return Real(dm_atan2((double)y.v, (double)x.v));
}
inline Real sqrt(Real x) { return Real(::sqrtf(x.v)); }
inline Real inversesqrt(Real x) { return Real(1.0 / ::sqrt((double)x.v)); }
inline Real abs(Real x) { return Real(::fabsf(x.v)); }
inline Real min(Real a, Real b) { return Real(::fminf(a.v, b.v)); }
inline Real max(Real a, Real b) { return Real(::fmaxf(a.v, b.v)); }
inline Real clamp(Real x, Real lo, Real hi) { return Real(::fminf(::fmaxf(x.v, lo.v), hi.v)); }
inline Real mix(Real x, Real y, Real a) { return Real(x.v * (1.0f - a.v) + y.v * a.v); }   // x*(1-a) + y*a
inline Real floor(Real x) { return Real(::floorf(x.v)); }
inline int findMSB(unsigned v) { return v ? 31 - __builtin_clz(v) : -1; }
inline Real uintBitsToFloat(unsigned u) { Real r; std::memcpy(&r.v, &u, 4); return r; }
inline unsigned floatBitsToUint(Real f) { unsigned u; std::memcpy(&u, &f.v, 4); return u; }

// ---------------------------------------------------------------------------------------------------------------
// vectors.  Components are ordinary members; a swizzle (.xy, .zw, .yx, .a) is an EMPTY proxy member placed at offset 0
// of its owner ([[no_unique_address]]) that reaches the owner's components through its own address.  Only what the six
// shaders use is provided.
// ---------------------------------------------------------------------------------------------------------------
template <typename V, typename T, int A, int B>
struct Swz2 {                      // two-component swizzle proxy
    operator V() const { const T* d = reinterpret_cast<const T*>(this); return V(d[A], d[B]); }
    Swz2& operator=(const V& o) { T* d = reinterpret_cast<T*>(this); const T a = o.x, b = o.y; d[A] = a; d[B] = b; return *this; }
};
template <typename T, int A>
struct Swz1 {                      // one-component alias (.a of a vec4)
    operator T() const { return reinterpret_cast<const T*>(this)[A]; }
};

struct ivec2;
struct uvec2;
struct vec2 {
    Real x, y;
    vec2() = default;
    vec2(Real a, Real b) : x(a), y(b) {}
    explicit vec2(Real a) : x(a), y(a) {}
    inline vec2(const ivec2& o);               // implicit int -> float conversion of GLSL
    inline explicit vec2(const uvec2& o);
    Real& operator[](int i) { return (&x)[i]; }
    const Real& operator[](int i) const { return (&x)[i]; }
};
struct ivec2 {
    int x, y;
    ivec2() = default;
    ivec2(int a, int b) : x(a), y(b) {}
    explicit ivec2(const vec2& o) : x((int)o.x.v), y((int)o.y.v) {}
};
struct uvec2 {
    [[no_unique_address]] Swz2<uvec2, unsigned, 0, 1> xy;
    [[no_unique_address]] Swz2<uvec2, unsigned, 1, 0> yx;
    unsigned x, y;
    uvec2() = default;
    uvec2(unsigned a, unsigned b) : x(a), y(b) {}
    explicit uvec2(unsigned a) : x(a), y(a) {}
    explicit uvec2(const ivec2& o) : x((unsigned)o.x), y((unsigned)o.y) {}
    unsigned& operator[](int i) { return (&x)[i]; }
    const unsigned& operator[](int i) const { return (&x)[i]; }
};
inline vec2::vec2(const ivec2& o) : x((float)o.x), y((float)o.y) {}
inline vec2::vec2(const uvec2& o) : x((float)o.x), y((float)o.y) {}

struct ivec3 {
    [[no_unique_address]] Swz2<ivec2, int, 0, 1> xy;
    int x, y, z;
    ivec3() = default;
    ivec3(int a, int b, int c) : x(a), y(b), z(c) {}
    ivec3(const uvec2& a, unsigned c) : x((int)a.x), y((int)a.y), z((int)c) {}
};
struct uvec3 {
    [[no_unique_address]] Swz2<uvec2, unsigned, 0, 1> xy;
    [[no_unique_address]] Swz2<uvec2, unsigned, 1, 0> yx;
    unsigned x, y, z;
    uvec3() = default;
    uvec3(unsigned a, unsigned b, unsigned c) : x(a), y(b), z(c) {}
    uvec3(const uvec2& a, unsigned c) : x(a.x), y(a.y), z(c) {}
};
struct vec4 {
    [[no_unique_address]] Swz2<vec2, Real, 0, 1> xy;
    [[no_unique_address]] Swz2<vec2, Real, 2, 3> zw;
    [[no_unique_address]] Swz1<Real, 3> a;
    Real x, y, z, w;
    vec4() = default;
    vec4(Real a_, Real b_, Real c_, Real d_) : x(a_), y(b_), z(c_), w(d_) {}
    vec4(const vec2& p, const vec2& q) : x(p.x), y(p.y), z(q.x), w(q.y) {}
    vec4(const vec2& p, Real c_, Real d_) : x(p.x), y(p.y), z(c_), w(d_) {}
    Real& operator[](int i) { return (&x)[i]; }
    const Real& operator[](int i) const { return (&x)[i]; }
};
static_assert(sizeof(vec2) == 8 && sizeof(vec4) == 16, "buffer element sizes (std430 vec2 / vec4)");
static_assert(__builtin_offsetof(vec4, x) == 0 && __builtin_offsetof(uvec3, x) == 0 && __builtin_offsetof(ivec3, x) == 0 &&
                  __builtin_offsetof(uvec2, x) == 0,
              "swizzle proxies must sit at offset 0 of their owner");

// float vectors (component-wise; a product of components is rounded on the spot, like any vector operation)
inline vec2 operator+(const vec2& a, const vec2& b) { return vec2(a.x + b.x, a.y + b.y); }
inline vec2 operator-(const vec2& a, const vec2& b) { return vec2(a.x - b.x, a.y - b.y); }
inline vec2 operator*(const vec2& a, const vec2& b) { return vec2(Real(a.x * b.x), Real(a.y * b.y)); }
inline vec2 operator/(const vec2& a, const vec2& b) { return vec2(a.x / b.x, a.y / b.y); }
inline vec2 operator*(const vec2& a, Real s) { return vec2(Real(a.x * s), Real(a.y * s)); }
inline vec2 operator*(Real s, const vec2& a) { return vec2(Real(s * a.x), Real(s * a.y)); }
inline vec2 operator/(const vec2& a, Real s) { return vec2(a.x / s, a.y / s); }
inline vec2 operator/(Real s, const vec2& a) { return vec2(s / a.x, s / a.y); }
inline vec2 operator+(Real s, const vec2& a) { return vec2(s + a.x, s + a.y); }
inline vec2 operator+(const vec2& a, Real s) { return vec2(a.x + s, a.y + s); }
inline vec2 operator-(const vec2& a) { return vec2(-a.x, -a.y); }
inline vec4 operator*(const vec4& a, Real s) { return vec4(Real(a.x * s), Real(a.y * s), Real(a.z * s), Real(a.w * s)); }
inline vec2 abs(const vec2& a) { return vec2(abs(a.x), abs(a.y)); }
inline Real length(const vec2& a) { return Real(::sqrtf(a.x.v * a.x.v + a.y.v * a.y.v)); }
inline vec2 mod(const vec2& x, const vec2& y) {   // x - y*floor(x/y)
    return vec2(Real(x.x.v - y.x.v * ::floorf(x.x.v / y.x.v)), Real(x.y.v - y.y.v * ::floorf(x.y.v / y.y.v)));
}
// integer vectors
inline ivec2 operator+(const ivec2& a, const ivec2& b) { return ivec2(a.x + b.x, a.y + b.y); }
inline ivec2 operator-(const ivec2& a) { return ivec2(-a.x, -a.y); }
inline vec2 operator*(const ivec2& a, Real s) { return vec2(Real(Real(a.x) * s), Real(Real(a.y) * s)); }
inline vec2 operator-(const ivec2& a, const vec2& b) { return vec2(Real(a.x) - b.x, Real(a.y) - b.y); }
inline uvec2 operator>>(const uvec2& a, int s) { return uvec2(a.x >> s, a.y >> s); }
inline uvec2 operator&(const uvec2& a, const uvec2& b) { return uvec2(a.x & b.x, a.y & b.y); }
inline uvec2 operator*(const uvec2& a, unsigned s) { return uvec2(a.x * s, a.y * s); }
inline uvec2 operator+(const uvec2& a, const uvec2& b) { return uvec2(a.x + b.x, a.y + b.y); }
inline uvec2 floatBitsToUint(const vec2& f) { return uvec2(floatBitsToUint(f.x), floatBitsToUint(f.y)); }

// ---------------------------------------------------------------------------------------------------------------
// storage images (texture arrays created by wave_generator.gd:31,34-35)
// ---------------------------------------------------------------------------------------------------------------
enum ImageFormat { RGBA32F = 0, RGBA16F = 1 };
struct image2DArray {
    void* data = nullptr;      // [layers][height][width][4] float or IEEE half
    int width = 0, height = 0, layers = 0;
    int format = RGBA32F;      // the format the texture was CREATED with (not the shader's format qualifier)
};
inline ivec3 imageSize(const image2DArray& im) { return ivec3(im.width, im.height, im.layers); }
inline vec4 imageLoad(const image2DArray& im, const ivec3& p) {
    const size_t o = (((size_t)p.z * im.height + p.y) * im.width + p.x) * 4;
    vec4 r;
    if (im.format == RGBA32F) {
        const float* f = static_cast<const float*>(im.data) + o;
        for (int i = 0; i < 4; ++i) r[i] = Real(f[i]);
    } else {
        const _Float16* h = static_cast<const _Float16*>(im.data) + o;
        for (int i = 0; i < 4; ++i) r[i] = Real((float)h[i]);
    }
    return r;
}
inline void imageStore(const image2DArray& im, const ivec3& p, const vec4& v) {
    const size_t o = (((size_t)p.z * im.height + p.y) * im.width + p.x) * 4;
    if (im.format == RGBA32F) {
        float* f = static_cast<float*>(im.data) + o;
        for (int i = 0; i < 4; ++i) f[i] = v[i].v;
    } else {
        _Float16* h = static_cast<_Float16*>(im.data) + o;
        for (int i = 0; i < 4; ++i) h[i] = (_Float16)v[i].v;      // round-to-nearest-even
    }
}

// ---------------------------------------------------------------------------------------------------------------
// invocation built-ins and the workgroup executor
// ---------------------------------------------------------------------------------------------------------------
extern thread_local uvec3 gl_NumWorkGroups, gl_WorkGroupID, gl_LocalInvocationID, gl_GlobalInvocationID;

void barrier();                                   // switches to the next invocation of the workgroup

// Runs groups.x * groups.y * groups.z workgroups of `local` invocations of entry() (OpenMP over workgroups).
// with_barrier = the shader calls barrier(): every invocation then gets its own fiber.
void dispatch(void (*entry)(), uvec3 local, uvec3 groups, bool with_barrier);

// resource registry of one shader (filled by the GLSL_BUFFER / GLSL_IMAGE declarations)
struct Binding {
    int set, binding;
    bool is_image;
    void* slot;            // T** of a buffer, image2DArray* of an image
};
struct ShaderModule {
    const char* name;
    void (*entry)();
    uvec3 local_size;
    bool has_barrier;
    void* push_constants;
    size_t push_constant_size;
    std::vector<Binding> bindings;
};
void register_shader(ShaderModule* m);
struct BindingRegistrar {
    BindingRegistrar(ShaderModule& m, int set, int binding, bool is_image, void* slot) { m.bindings.push_back(Binding{set, binding, is_image, slot}); }
};

}  // namespace glsl
