#include "host_util.h"
#include <stdlib.h>
#include <stdarg.h>
#include <string.h>
#include <mutex>
#include <unordered_map>
#include <atomic>
#include <vector>

#include <nvtx3/nvToolsExt.h>

namespace ttts {

NvtxRange::NvtxRange(const char* name) { nvtxRangePushA(name); }
NvtxRange::NvtxRange(const char* name, int index) {
    char buf[64];
    snprintf(buf, sizeof(buf), "%s %d", name, index);
    nvtxRangePushA(buf);
}
NvtxRange::~NvtxRange() { nvtxRangePop(); }


static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int fail_cuda(cudaError_t e, const char* what) {
    set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    return TTTS_ERR_CUDA;
}

static std::atomic<unsigned long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

struct ProfRec { cudaEvent_t e0, e1; double flops; };
static int g_prof = 0;             // 0 off | 1 tcgen05 GEMMs | 2 conv1d dgrad | 3 conv1d wgrad | 4 conv1d forward (fp32 kernels) | 5 conv1d_tcs
static std::vector<ProfRec> g_prof_recs;
static size_t g_prof_used = 0;
bool prof_enabled() { return g_prof == 1; }
static bool g_prof_open = false;
void prof_begin(int kind, cudaStream_t st, double flops) {
    if (g_prof != kind) return;
    g_prof_open = true;
    if (g_prof_used == g_prof_recs.size()) {
        ProfRec r; cudaEventCreate(&r.e0); cudaEventCreate(&r.e1); r.flops = 0; g_prof_recs.push_back(r);
    }
    g_prof_recs[g_prof_used].flops = flops;
    cudaEventRecord(g_prof_recs[g_prof_used].e0, st);
}
void prof_gemm_begin(cudaStream_t st, double flops) { prof_begin(1, st, flops); }
void prof_end(int kind, cudaStream_t st) {
    if (g_prof != kind || !g_prof_open) return;
    g_prof_open = false;
    cudaEventRecord(g_prof_recs[g_prof_used].e1, st);
    ++g_prof_used;
}
void prof_gemm_end(cudaStream_t st) { prof_end(1, st); }

bool pdl_enabled() {
    static int on = -1;
    if (on < 0) { const char* e = getenv("TTTS_PDL"); on = (e && e[0] == '0') ? 0 : 1; }
    return on != 0;
}

int num_sms() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, []() {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    });
    return fn;
}

struct TmapKey {
    const void* p; int eb; uint64_t inner, outer, ld; uint32_t bi, bo; int sw; int dev;
    bool operator==(const TmapKey& o) const {
        return p == o.p && eb == o.eb && inner == o.inner && outer == o.outer && ld == o.ld && bi == o.bi && bo == o.bo &&
               sw == o.sw && dev == o.dev;
    }
};
struct TmapKeyHash {
    size_t operator()(const TmapKey& k) const {
        uint64_t h = (uint64_t)(uintptr_t)k.p * 0x9E3779B97F4A7C15ULL;
        h ^= k.inner * 0xff51afd7ed558ccdULL + k.outer * 0xc4ceb9fe1a85ec53ULL + k.ld * 31 + k.bi * 131 + k.bo * 17 + k.eb + k.sw * 7 + k.dev * 1315423911ULL;
        return (size_t)h;
    }
};

int make_tmap_2d(CUtensorMap* out, const void* gptr, int elem_bytes, uint64_t inner, uint64_t outer, uint64_t ld_elems,
                 uint32_t box_inner, uint32_t box_outer, int swizzle) {
    static std::mutex mu;
    static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
    int dev = 0;
    cudaGetDevice(&dev);
    TmapKey key{gptr, elem_bytes, inner, outer, ld_elems, box_inner, box_outer, swizzle, dev};
    {
        std::lock_guard<std::mutex> g(mu);
        auto it = cache.find(key);
        if (it != cache.end()) { *out = it->second; return TTTS_OK; }
    }
    PFN_encodeTiled enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled driver entry point unavailable (no CUDA driver?)"); return TTTS_ERR_CUDA; }
    TTTS_CHECK_ARG(((uintptr_t)gptr & 15) == 0, "TMA base pointer %p not 16B aligned", gptr);
    TTTS_CHECK_ARG((ld_elems * (uint64_t)elem_bytes) % 16 == 0, "TMA row stride %llu B not a multiple of 16", (unsigned long long)(ld_elems * elem_bytes));
    TTTS_CHECK_ARG(box_inner <= 256 && box_outer <= 256, "TMA box too large");
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {ld_elems * (uint64_t)elem_bytes};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    CUresult r = enc(out, dt, 2, const_cast<void*>(gptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     swizzle == 1 ? CU_TENSOR_MAP_SWIZZLE_128B : (swizzle == 2 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed: %d (inner=%llu outer=%llu ld=%llu box=%ux%u)", (int)r, (unsigned long long)inner,
                  (unsigned long long)outer, (unsigned long long)ld_elems, box_inner, box_outer);
        return TTTS_ERR_CUDA;
    }
    {
        std::lock_guard<std::mutex> g(mu);
        if (cache.size() > 65536) cache.clear();
        cache[key] = *out;
    }
    return TTTS_OK;
}

int make_tmap_mn3d(CUtensorMap* out, const void* gptr, uint64_t mn, uint64_t k_rows, uint64_t ld_elems, uint32_t box_k, uint32_t atoms) {
    static std::mutex mu;
    static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
    int dev = 0;
    cudaGetDevice(&dev);
    TmapKey key{gptr, 3, mn, k_rows, ld_elems, box_k, atoms, true, dev};
    {
        std::lock_guard<std::mutex> g(mu);
        auto it = cache.find(key);
        if (it != cache.end()) { *out = it->second; return TTTS_OK; }
    }
    PFN_encodeTiled enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled driver entry point unavailable (no CUDA driver?)"); return TTTS_ERR_CUDA; }
    TTTS_CHECK_ARG(mn % 64 == 0, "3-D MN-major tensor map needs MN %% 64 == 0");
    TTTS_CHECK_ARG(((uintptr_t)gptr & 15) == 0 && (ld_elems * 2) % 16 == 0, "TMA alignment");
    cuuint64_t dims[3] = {64, k_rows, mn / 64};
    cuuint64_t strides[2] = {ld_elems * 2, 128};
    cuuint32_t box[3] = {64, box_k, atoms};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(gptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(3d) failed: %d (mn=%llu k=%llu ld=%llu box_k=%u atoms=%u)", (int)r, (unsigned long long)mn,
                  (unsigned long long)k_rows, (unsigned long long)ld_elems, box_k, atoms);
        return TTTS_ERR_CUDA;
    }
    {
        std::lock_guard<std::mutex> g(mu);
        if (cache.size() > 65536) cache.clear();
        cache[key] = *out;
    }
    return TTTS_OK;
}

}  // namespace ttts

extern "C" {

int ttts_version(void) { return 100; }
const char* ttts_last_error(void) { return ttts::g_err; }

unsigned long long ttts_launch_count(void) { return ttts::g_launches.load(); }

void ttts_prof_gemm_enable(int on) { ttts::g_prof = on; if (on) ttts::g_prof_used = 0; }
/* synchronises the device, sums the bracketed GEMM launches since enable: total ms, total FLOPs, launch count */
int ttts_prof_gemm_read(double* ms_total, double* flops_total, long long* launches) {
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return ttts::fail_cuda(e, "prof sync");
    double ms = 0, fl = 0;
    for (size_t i = 0; i < ttts::g_prof_used; ++i) {
        float t = 0.f;
        cudaEventElapsedTime(&t, ttts::g_prof_recs[i].e0, ttts::g_prof_recs[i].e1);
        ms += t; fl += ttts::g_prof_recs[i].flops;
    }
    if (ms_total) *ms_total = ms;
    if (flops_total) *flops_total = fl;
    if (launches) *launches = (long long)ttts::g_prof_used;
    return TTTS_OK;
}

int ttts_device_ok(void) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
    return major == 10 ? 1 : 0;
}
}
