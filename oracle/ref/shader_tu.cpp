// oracle/ref/shader_tu.cpp -- TEST INFRASTRUCTURE (recipe of oracle/_ref; not part of the product).
//
// One translation unit per reference shader: compiled with -DSHADER_NAME=<name> -DSHADER_INC="<oracle/_ref/gen/name.inc>",
// it includes the (lexically pre-processed, otherwise unmodified) GLSL text inside its own namespace, on top of
// glsl_shim.hpp, and registers the resulting module (entry point, workgroup size, bindings, push-constant block).
#include "glsl_shim.hpp"

#define GLSL_CAT2(a, b) a##b
#define GLSL_CAT(a, b) GLSL_CAT2(a, b)
#define GLSL_STR2(x) #x
#define GLSL_STR(x) GLSL_STR2(x)

namespace SHADER_NAME {
using namespace glsl;

static ShaderModule module_;

// ---- what glsl2cpp.py turned the layout() declarations into ----
#define GLSL_LOCAL_SIZE(X, Y, Z) static const uvec3 gl_WorkGroupSize((unsigned)(X), (unsigned)(Y), (unsigned)(Z));
#define GLSL_BUFFER(SET, BINDING, TYPE, NAME) \
    static TYPE* NAME;                        \
    static BindingRegistrar GLSL_CAT(bind_, NAME)(module_, SET, BINDING, false, &NAME);
#define GLSL_IMAGE(SET, BINDING, NAME) \
    static image2DArray NAME;          \
    static BindingRegistrar GLSL_CAT(bind_, NAME)(module_, SET, BINDING, true, &NAME);
// push constants: the block becomes a struct that receives the raw bytes RenderingContext.create_push_constant packs
// (assets/render_context.gd:122-135; std430 offsets == natural C++ offsets for the blocks of these shaders: 4-byte scalars
// and 8-byte-aligned vec2/ivec2 in non-decreasing alignment order), and every member gets an unqualified alias.
#define GLSL_PUSH_CONSTANTS_BEGIN struct PushConstants_ {
#define GLSL_PUSH_CONSTANTS_END }; static PushConstants_ push_constants_;
#define GLSL_PUSH_CONSTANT(TYPE, NAME) static TYPE& NAME = push_constants_.NAME;

// ---- GLSL keywords that are not C++ ----
#define float Real
#define in
#define shared static thread_local
typedef unsigned int uint;

#include SHADER_INC

#undef float
#undef in
#undef shared

static void entry_() { main(); }
static struct Registrar {
    Registrar() {
        module_.name = GLSL_STR(SHADER_NAME);
        module_.entry = entry_;
        module_.local_size = gl_WorkGroupSize;
        module_.has_barrier = GLSL_HAS_BARRIER != 0;
#if GLSL_HAS_PUSH_CONSTANTS
        module_.push_constants = &push_constants_;
        module_.push_constant_size = sizeof push_constants_;
#else
        module_.push_constants = nullptr;
        module_.push_constant_size = 0;
#endif
        register_shader(&module_);
    }
} registrar_;

}  // namespace SHADER_NAME
