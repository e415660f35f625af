// Host-side helpers shared by the C-ABI translation units: error reporting, launch checks,
// device properties, TMA tensor-map encoding (driver entry point resolved at run time so the
// library has no link-time dependency on libcuda and builds on GPU-less boxes).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include "../../include/ttts_b200.h"

namespace ttts {

void set_error(const char* fmt, ...);
int fail_cuda(cudaError_t e, const char* what);
int num_sms();
void count_launch();
// GEMM profiling (bench.py roofline): when enabled, every GEMM launch is bracketed by CUDA events on its stream
bool prof_enabled();
void prof_gemm_begin(cudaStream_t st, double flops);
void prof_gemm_end(cudaStream_t st);
// the same bracket for other kernel families (ttts_prof_gemm_enable(kind): 2 conv1d dgrad, 3 conv1d wgrad, 4 conv1d forward, 5 conv1d_tcs)
void prof_begin(int kind, cudaStream_t st, double flops);
void prof_end(int kind, cudaStream_t st);

#define TTTS_CHECK_ARG(cond, ...)                    \
    do {                                             \
        if (!(cond)) {                               \
            ::ttts::set_error(__VA_ARGS__);          \
            return TTTS_ERR_INVALID;                 \
        }                                            \
    } while (0)

#define TTTS_CUDA(expr)                                           \
    do {                                                          \
        cudaError_t _e = (expr);                                  \
        if (_e != cudaSuccess) return ::ttts::fail_cuda(_e, #expr); \
    } while (0)

#define TTTS_LAUNCH_CHECK(name)                                   \
    do {                                                          \
        cudaError_t _e = cudaGetLastError();                      \
        if (_e != cudaSuccess) return ::ttts::fail_cuda(_e, name); \
        ::ttts::count_launch();                                   \
    } while (0)

#define TTTS_RUN(expr)                  \
    do {                                \
        int _rc = (expr);               \
        if (_rc != TTTS_OK) return _rc; \
    } while (0)

// NVTX range (header-only nvtx3: no link dependency, a no-op unless a profiler injects itself): the kernel groups of the step show up
// named in Nsight Systems / ncu --nvtx (the reference only has commented-out torch.autograd.profiler stubs)
struct NvtxRange {
    explicit NvtxRange(const char* name);
    NvtxRange(const char* name, int index);       // "name index"
    ~NvtxRange();
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};

// TTTS_PDL=0 disables programmatic dependent launch (common.cuh: pdl_wait) for A/B measurements
bool pdl_enabled();
// <<<grid, block, smem, st>>> with the PDL attribute; only for kernels that call pdl_wait() before their first global-memory access
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// plain <<<grid, block, smem, st>>> launch as a function call (sources that are also compiled by the CPU emulation, tests/emu, use it)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_plain(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// 2-D bf16/fp32 tiled tensor map; swizzle: 0 = none, 1 (true) = 128B, 2 = 64B.
//   inner/outer: tensor extents in elements (inner = contiguous dim); ld_elems: row stride in elements
//   box_inner/box_outer: box extents in elements
int make_tmap_2d(CUtensorMap* out, const void* gptr, int elem_bytes, uint64_t inner, uint64_t outer, uint64_t ld_elems,
                 uint32_t box_inner, uint32_t box_outer, int swizzle);

// 3-D view of an MN-major operand [K rows][MN contiguous] as {64 (mn), K, MN/64 atoms}: one TMA box {64, box_k, atoms}
// lands in shared memory as [atom][k][64] -- exactly the MN-major 128B-swizzle UMMA layout.  Requires MN % 64 == 0.
int make_tmap_mn3d(CUtensorMap* out, const void* gptr, uint64_t mn, uint64_t k_rows, uint64_t ld_elems, uint32_t box_k, uint32_t atoms);

}  // namespace ttts
