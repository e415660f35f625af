// TEST INFRASTRUCTURE (never part of the product): a minimal host emulation of the CUDA execution model, enough to run the plain
// (no TMA / tcgen05 / mbarrier) kernels of ttts_b200/csrc on CPU from their own source, so that index arithmetic, shared-memory
// choreography, barrier placement and the host-side launch sequence are exercised without a GPU.
//
//   * a grid runs block after block; inside a block every CUDA thread is an OS thread, so __syncthreads() and warp shuffles are real
//     rendezvous points (std::barrier) and a missing / divergent barrier shows up as a wrong answer or a deadlock, as on the device;
//   * `__shared__` variables become function-local statics (one block at a time, so one instance is right); dynamic shared memory is
//     reached through TTTS_DYN_SMEM;
//   * bf16 conversions round to nearest even like the hardware; packed bf16x2 arithmetic rounds once per operation;
//   * griddepcontrol (programmatic dependent launch) is a no-op: launches are synchronous and in order.
// What it cannot show: memory-model races between blocks, coalescing, occupancy, anything about performance.
#pragma once
#include <algorithm>
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#include "../../include/ttts_b200.h"

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))
#define TTTS_DEVICE inline
#define TTTS_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(::ttts_emu::tl.dyn_smem)

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct uint4 { uint32_t x, y, z, w; };
struct uint2 { uint32_t x, y; };
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
inline uint4 make_uint4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) { return uint4{a, b, c, d}; }
inline uint2 make_uint2(uint32_t a, uint32_t b) { return uint2{a, b}; }
inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }
inline float2 make_float2(float a, float b) { return float2{a, b}; }
typedef void* cudaStream_t;
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <typename F> inline int cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return 0; }
typedef int cudaError_t;
constexpr int cudaSuccess = 0;

namespace ttts_emu {
struct Warp {
    float x[32];
    std::barrier<> bar;
    explicit Warp(int n) : bar(n) {}
};
struct ThreadCtx {
    dim3 tid, bid, bdim, gdim;
    std::barrier<>* block_bar = nullptr;
    Warp* warp = nullptr;
    int lane = 0;
    void* dyn_smem = nullptr;
};
inline thread_local ThreadCtx tl;
inline unsigned long long launches = 0;

template <typename F>
void run_grid(dim3 grid, dim3 block, size_t smem, F&& body) {
    const int nt = (int)(block.x * block.y * block.z);
    ++launches;
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                std::vector<uint64_t> dyn((smem + 7) / 8 + 1);
                std::barrier<> bb(nt);
                std::vector<std::unique_ptr<Warp>> warps;
                for (int w = 0; w * 32 < nt; ++w) warps.emplace_back(new Warp(std::min(32, nt - w * 32)));
                std::vector<std::thread> th;
                th.reserve(nt);
                for (int t = 0; t < nt; ++t) {
                    th.emplace_back([&, t] {
                        tl.tid = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
                        tl.bid = dim3(bx, by, bz);
                        tl.bdim = block; tl.gdim = grid;
                        tl.block_bar = &bb; tl.warp = warps[t / 32].get(); tl.lane = t % 32; tl.dyn_smem = dyn.data();
                        body();
                        tl.warp->bar.arrive_and_drop();      // an exited thread no longer takes part in barriers (CUDA semantics)
                        bb.arrive_and_drop();
                    });
                }
                for (auto& x : th) x.join();
            }
}
}  // namespace ttts_emu

#define threadIdx (::ttts_emu::tl.tid)
#define blockIdx (::ttts_emu::tl.bid)
#define blockDim (::ttts_emu::tl.bdim)
#define gridDim (::ttts_emu::tl.gdim)

inline void __syncthreads() { ::ttts_emu::tl.block_bar->arrive_and_wait(); }
inline float __shfl_xor_sync(unsigned, float v, int o) {
    auto& t = ::ttts_emu::tl;
    t.warp->x[t.lane] = v;
    t.warp->bar.arrive_and_wait();
    const float r = t.warp->x[t.lane ^ o];
    t.warp->bar.arrive_and_wait();
    return r;
}
template <typename T> inline T __ldg(const T* p) { return *p; }
inline float atomicAdd(float* p, float v) { return std::atomic_ref<float>(*p).fetch_add(v, std::memory_order_relaxed); }
inline float __expf(float x) { return expf(x); }
inline void sincospif(float x, float* s, float* c) { *s = (float)sin(3.14159265358979323846 * (double)x); *c = (float)cos(3.14159265358979323846 * (double)x); }
inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
using std::max;
using std::min;

// ---- bf16 ----
struct __nv_bfloat16 { uint16_t x; };
inline __nv_bfloat16 __float2bfloat16_rn(float f) {
    uint32_t u = __float_as_uint(f);
    __nv_bfloat16 h;
    if ((u & 0x7fffffffu) > 0x7f800000u) { h.x = 0x7fff; return h; }      // NaN
    u += 0x7fffu + ((u >> 16) & 1u);
    h.x = (uint16_t)(u >> 16);
    return h;
}
inline float __bfloat162float(__nv_bfloat16 h) { return __uint_as_float((uint32_t)h.x << 16); }

namespace ttts {
typedef __nv_bfloat16 bf16;

// ---- the subset of common.cuh / host_util.h / kernels.h that the emulated sources use ----
inline float warp_sum(float v) { for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o); return v; }
inline float warp_max(float v) { for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o)); return v; }
inline float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
inline float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }
inline float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
inline uint32_t pack_bf16(float lo, float hi) { return (uint32_t)__float2bfloat16_rn(lo).x | ((uint32_t)__float2bfloat16_rn(hi).x << 16); }
// gelu_new of two packed bf16 values with bf16 intermediates (common.cuh: gelu_new_bf2; fma rounds once, tanh.approx ~ bf16(tanh))
inline float emu_gelu_bf16(float x) {
    const float k0 = bf16_hi(0x3F4C0000u), k0k1 = bf16_hi(0x3D120000u);
    const float x2 = bf16_round(x * x);
    const float u = bf16_round(x * bf16_round(fmaf(x2, k0k1, k0)));
    const float t = bf16_round(tanhf(u));
    const float hx = bf16_round(x * 0.5f);
    return bf16_round(fmaf(hx, t, hx));
}
inline uint32_t gelu_new_bf2(uint32_t x) { return pack_bf16(emu_gelu_bf16(bf16_lo(x)), emu_gelu_bf16(bf16_hi(x))); }
// cp.async as a synchronous copy (zero fill when !pred); commit / wait are no-ops.  The emulation therefore cannot see a MISSING wait;
// it does see wrong addresses, wrong stage indices and misplaced barriers.
inline void cp_async4(void* smem_dst, const void* gsrc, bool pred) {
    if (pred) memcpy(smem_dst, gsrc, 4); else memset(smem_dst, 0, 4);
}
inline void cp_async16(void* smem_dst, const void* gsrc, bool pred) {
    if (pred) memcpy(smem_dst, gsrc, 16); else memset(smem_dst, 0, 16);
}
inline void cp_async_commit() {}
template <int N> inline void cp_async_wait() {}
inline void pdl_wait() {}
inline void pdl_launch_dependents() {}

inline char g_err[512];
inline void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
inline int num_sms() { return 2; }      // keeps the SM-capped grids of the element-wise kernels small: every launch costs one OS thread per CUDA thread here
inline void count_launch() {}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t, Args... args) {
    ::ttts_emu::run_grid(grid, block, smem, [&] { kern(static_cast<KArgs>(args)...); });
    return cudaSuccess;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_plain(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    return launch_pdl(kern, grid, block, smem, st, args...);
}
inline void prof_begin(int, cudaStream_t, double) {}
inline void prof_end(int, cudaStream_t) {}
}  // namespace ttts

#define TTTS_CHECK_ARG(cond, ...)            \
    do {                                     \
        if (!(cond)) {                       \
            ::ttts::set_error(__VA_ARGS__);  \
            return TTTS_ERR_INVALID;         \
        }                                    \
    } while (0)
#define TTTS_CUDA(expr)                              \
    do {                                             \
        if ((expr) != cudaSuccess) return TTTS_ERR_CUDA; \
    } while (0)
#define TTTS_LAUNCH_CHECK(name) do { } while (0)
#define TTTS_RUN(expr)                  \
    do {                                \
        int _rc = (expr);               \
        if (_rc != TTTS_OK) return _rc; \
    } while (0)
