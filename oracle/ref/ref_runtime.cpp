// oracle/ref/ref_runtime.cpp -- TEST INFRASTRUCTURE (recipe of oracle/_ref; not part of the product).
//
// Executes the reference's compute shaders (compiled from their own text by shader_tu.cpp) on the CPU:
//   * workgroup executor: invocations of a workgroup are fibers when the shader uses barrier(), plain calls otherwise;
//     workgroups are independent and run under OpenMP;
//   * a tiny "rendering device": the resources wave_generator.gd:31-35 creates, the uniform sets of :37-41 and
//     dispatches with raw push-constant bytes, exposed through a C ABI for oracle/pyref.py.
// Nothing here knows the algorithm: every number the maps end up holding is produced by the reference's GLSL.
#include "glsl_shim.hpp"

#include <cstdio>
#include <map>
#include <string>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace glsl {

int g_contract = 1;
int g_math = 0;
thread_local uvec3 gl_NumWorkGroups, gl_WorkGroupID, gl_LocalInvocationID, gl_GlobalInvocationID;

// ---------------------------------------------------------------------------------------------------------------
// fibers (x86-64 System V): a context is a stack pointer; the callee-saved registers live on the fiber's own stack
// ---------------------------------------------------------------------------------------------------------------
#if !defined(__x86_64__)
#error "oracle/ref fibers are written for x86-64 (System V ABI)"
#endif
extern "C" void glsl_fiber_switch(void** save_sp, void* load_sp);
asm(R"(
    .text
    .globl glsl_fiber_switch
    .type glsl_fiber_switch, @function
glsl_fiber_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
    .size glsl_fiber_switch, .-glsl_fiber_switch
)");

namespace {
constexpr size_t kStackBytes = 32 * 1024;
struct Fiber {
    void* sp = nullptr;
    bool done = false;
    uvec3 local_id, global_id;
};
struct WorkgroupRunner {              // one per OS thread, kept for the life of the thread
    char* stacks = nullptr;           // uninitialised; pages are touched only as far as the fibers use them
    size_t stack_count = 0;
    std::vector<Fiber> fibers;
    ~WorkgroupRunner() { std::free(stacks); }
    void* scheduler_sp = nullptr;
    Fiber* current = nullptr;
    void (*entry)() = nullptr;
};
thread_local WorkgroupRunner* t_runner = nullptr;

[[noreturn]] void fiber_main() {
    WorkgroupRunner* r = t_runner;
    r->entry();
    r->current->done = true;
    void* dummy;
    glsl_fiber_switch(&dummy, r->scheduler_sp);
    __builtin_unreachable();
}

void run_workgroup_fibers(WorkgroupRunner& r, uvec3 local, uvec3 group_id) {
    const size_t n = (size_t)local.x * local.y * local.z;
    if (r.stack_count < n) {
        r.fibers.resize(n);
        std::free(r.stacks);
        r.stacks = static_cast<char*>(std::aligned_alloc(64, n * kStackBytes));
        if (!r.stacks) std::abort();
        r.stack_count = n;
    }
    char* base = r.stacks;
    size_t i = 0;
    for (unsigned lz = 0; lz < local.z; ++lz)
        for (unsigned ly = 0; ly < local.y; ++ly)
            for (unsigned lx = 0; lx < local.x; ++lx, ++i) {
                Fiber& f = r.fibers[i];
                f.done = false;
                f.local_id = uvec3(lx, ly, lz);
                f.global_id = uvec3(group_id.x * local.x + lx, group_id.y * local.y + ly, group_id.z * local.z + lz);
                // initial frame: six callee-saved registers, the entry address `ret` pops, a null return address
                void** top = reinterpret_cast<void**>(base + (i + 1) * kStackBytes);     // 16-byte aligned
                top[-1] = nullptr;
                top[-2] = reinterpret_cast<void*>(&fiber_main);
                for (int k = 3; k <= 8; ++k) top[-k] = nullptr;
                f.sp = &top[-8];
            }
    size_t live = n;
    while (live) {                       // one sweep = every live invocation runs up to its next barrier() (or to its end)
        live = 0;
        for (size_t k = 0; k < n; ++k) {
            Fiber& f = r.fibers[k];
            if (f.done) continue;
            r.current = &f;
            gl_LocalInvocationID = f.local_id;
            gl_GlobalInvocationID = f.global_id;
            glsl_fiber_switch(&r.scheduler_sp, f.sp);
            if (!f.done) ++live;
        }
    }
}
}  // namespace

void barrier() {
    WorkgroupRunner* r = t_runner;
    Fiber* self = r->current;
    glsl_fiber_switch(&self->sp, r->scheduler_sp);
}

void dispatch(void (*entry)(), uvec3 local, uvec3 groups, bool with_barrier) {
    const long total = (long)groups.x * groups.y * groups.z;
#pragma omp parallel
    {
        static thread_local WorkgroupRunner runner;
        runner.entry = entry;
        t_runner = &runner;
#pragma omp for schedule(dynamic, 1)
        for (long g = 0; g < total; ++g) {
            const uvec3 gid((unsigned)(g % groups.x), (unsigned)((g / groups.x) % groups.y), (unsigned)(g / ((long)groups.x * groups.y)));
            gl_NumWorkGroups = groups;
            gl_WorkGroupID = gid;
            if (with_barrier) {
                run_workgroup_fibers(runner, local, gid);
            } else {
                for (unsigned lz = 0; lz < local.z; ++lz)
                    for (unsigned ly = 0; ly < local.y; ++ly)
                        for (unsigned lx = 0; lx < local.x; ++lx) {
                            gl_LocalInvocationID = uvec3(lx, ly, lz);
                            gl_GlobalInvocationID = uvec3(gid.x * local.x + lx, gid.y * local.y + ly, gid.z * local.z + lz);
                            entry();
                        }
            }
        }
        t_runner = nullptr;
    }
}

static std::map<std::string, ShaderModule*>& registry() {
    static std::map<std::string, ShaderModule*> r;
    return r;
}
void register_shader(ShaderModule* m) { registry()[m->name] = m; }

}  // namespace glsl

// ---------------------------------------------------------------------------------------------------------------
// C ABI (oracle/pyref.py)
// ---------------------------------------------------------------------------------------------------------------
using namespace glsl;

extern "C" {

void ref_set_modes(int math_mode, int contract_mode) { g_math = math_mode; g_contract = contract_mode; }
void ref_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}
int ref_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
int ref_has_shader(const char* shader) { return registry().count(shader) ? 1 : 0; }

// uniform_set_create for one binding (render_context.gd:89-97): storage buffer
int ref_bind_buffer(const char* shader, int set, int binding, void* data) {
    auto it = registry().find(shader);
    if (it == registry().end()) return -1;
    for (Binding& b : it->second->bindings)
        if (b.set == set && b.binding == binding && !b.is_image) {
            *static_cast<void**>(b.slot) = data;
            return 0;
        }
    return -2;
}
// ... storage image; format: 0 = R32G32B32A32_SFLOAT, 1 = R16G16B16A16_SFLOAT (the format the texture was created with)
int ref_bind_image(const char* shader, int set, int binding, void* data, int width, int height, int layers, int format) {
    auto it = registry().find(shader);
    if (it == registry().end()) return -1;
    for (Binding& b : it->second->bindings)
        if (b.set == set && b.binding == binding && b.is_image) {
            image2DArray* im = static_cast<image2DArray*>(b.slot);
            im->data = data;
            im->width = width;
            im->height = height;
            im->layers = layers;
            im->format = format;
            return 0;
        }
    return -2;
}
// compute_list_set_push_constant + compute_list_dispatch (render_context.gd:110-118)
int ref_dispatch(const char* shader, const void* push_constant, int push_constant_size, int gx, int gy, int gz) {
    auto it = registry().find(shader);
    if (it == registry().end()) return -1;
    ShaderModule* m = it->second;
    // the packed array is padded to a multiple of 16 bytes (render_context.gd:126-129); the block reads its own size
    if (push_constant_size < 0 || (size_t)push_constant_size < m->push_constant_size) return -3;
    if (m->push_constant_size) std::memcpy(m->push_constants, push_constant, m->push_constant_size);
    for (const Binding& b : m->bindings) {
        const bool bound = b.is_image ? static_cast<image2DArray*>(b.slot)->data != nullptr : *static_cast<void**>(b.slot) != nullptr;
        if (!bound) return -4;
    }
    dispatch(m->entry, m->local_size, uvec3((unsigned)gx, (unsigned)gy, (unsigned)gz), m->has_barrier);
    return 0;
}
int ref_local_size(const char* shader, int* xyz) {
    auto it = registry().find(shader);
    if (it == registry().end()) return -1;
    xyz[0] = (int)it->second->local_size.x;
    xyz[1] = (int)it->second->local_size.y;
    xyz[2] = (int)it->second->local_size.z;
    return 0;
}

}  // extern "C"
