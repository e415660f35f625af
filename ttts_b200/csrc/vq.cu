// RVQ (n_q = 1) codebook lookup for the VQ-VAE encode front end, exact fp32:
//   dist = -(|x|^2 - 2 x.E^T + |e|^2) ; index = argmax (first max wins)          ttts/vqvae/core_vq.py:174-182
//   dequantize (row gather), straight-through value x + (q - x), commitment-loss partials   core_vq.py:205-230,303-322
//   training extras: code histogram, embed_sum scatter-add, EMA + Laplace-smoothed renormalisation  core_vq.py:212-228
//
// The reference computes the distances with an fp32 GEMM; to keep "bit-exact indices" meaningful this kernel stays on the
// FP32 FMA pipe (sequential-k fused multiply-add, no TF32/bf16 rounding).  At K=1024, D=192 that makes the kernel
// FMA-bound, not HBM-bound (255 FLOP/B) -- reported as such in DESIGN.md.
//
// Tiling: CTA = 64 vectors x all K codes; the 64x192 fp32 x-tile stays in shared memory (read from HBM exactly once),
// codebook streamed through shared memory in [16 dims x 128 codes] chunks (L2-resident, 786 KB); each thread owns a
// 4-vector x 8-code register tile; the per-vector argmax is reduced across the 16 lanes that share a vector with shuffles.
#include "common.cuh"
#include "host_util.h"
#include "kernels.h"

namespace ttts {

constexpr int VQ_TM = 64;      // vectors per CTA
constexpr int VQ_TN = 128;     // codes per tile
constexpr int VQ_DC = 16;      // dims per chunk
constexpr int VQ_THREADS = 256;

__global__ void vq_code_norms_kernel(const float* __restrict__ E, int K, int D, float* __restrict__ ee) {
    const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (k >= K) return;
    // same left-to-right association as a simple row sum would not match torch's pairwise sum bit-for-bit either; the
    // index test tolerates near-ties (tests/test_vq_gpu.py), everything else is insensitive at 1e-7.
    float s = 0.f;
    for (int d = lane; d < D; d += 32) { float v = E[(size_t)k * D + d]; s += v * v; }
    s = warp_sum(s);
    if (lane == 0) ee[k] = s;
}

struct VqLayout { long long sB, sD, sN; int Nn; };   // element (b, d, n) of x at b*sB + d*sD + n*sN ; vector v = b*Nn + n

// x tile [d][64 vectors].  SWZ: the 16 four-vector groups of row d are permuted by d mod 16, so a warp that writes (or reads back) 32
// consecutive dims of one vector -- the coalesced order for row-major [N, D] input -- touches 16 different bank groups instead of one;
// the main loop's float4 reads (one row, group ty) only see a different group number.
template <bool SWZ>
TTTS_DEVICE int vq_xs_index(int d, int vl) { return SWZ ? d * VQ_TM + ((((vl >> 2) ^ d) & 15) << 2) + (vl & 3) : d * VQ_TM + vl; }

// x tile -> shared memory (transposed to [d][v]) and |x|^2 per vector.  Ends with the tile visible to the whole CTA.
template <bool SWZ>
TTTS_DEVICE void vq_load_x_tile(const float* __restrict__ x, const VqLayout& lay, int N, int D, int v0, float* Xs, float* xx, int tid) {
    for (int i = tid; i < D * VQ_TM; i += VQ_THREADS) {
        int d, vl;
        if (lay.sD == 1) { vl = i / D; d = i - vl * D; }        // row-major [N, D]: consecutive threads along d
        else { d = i / VQ_TM; vl = i - d * VQ_TM; }             // [B, D, Nn]: consecutive threads along n
        const int v = v0 + vl;
        float val = 0.f;
        if (v < N) {
            const int b = v / lay.Nn, n = v - b * lay.Nn;
            val = x[(size_t)(b * lay.sB + d * lay.sD + n * lay.sN)];
        }
        Xs[vq_xs_index<SWZ>(d, vl)] = val;
    }
    __syncthreads();
    if (tid < VQ_TM) {
        float s = 0.f;
        for (int d = 0; d < D; ++d) { float t = Xs[vq_xs_index<SWZ>(d, tid)]; s += t * t; }
        xx[tid] = s;
    }
}

// Per-thread (best, index) of 4 vectors -> CTA-wide argmax (lowest index among equals), then the fused epilogue: indices, dequantised
// rows, straight-through values, commit-loss partial, code histogram and embedding sums for the EMA update.
template <bool SWZ>
TTTS_DEVICE void vq_finish_tile(float (&best)[4], int (&besti)[4], const VqLayout& lay, int N, int D,
                                const float* __restrict__ E, int v0, const float* Xs, int* sidx, int64_t* __restrict__ idx_out,
                                float* __restrict__ q_out, int straight_through, float* __restrict__ commit_partial, float* __restrict__ hist,
                                float* __restrict__ embed_sum, int tid) {
    const int tx = tid & 15, ty = tid >> 4;
    // reduce across the 16 lanes (tx) that share the same 4 vectors
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best[i], o);
            const int oi = __shfl_xor_sync(0xffffffffu, besti[i], o);
            if (ob > best[i] || (ob == best[i] && oi < besti[i])) { best[i] = ob; besti[i] = oi; }
        }
        if (tx == 0) sidx[ty * 4 + i] = besti[i];
    }
    __syncthreads();
    if (tid < VQ_TM && v0 + tid < N) {
        idx_out[v0 + tid] = (int64_t)sidx[tid];
        if (hist) atomicAdd(hist + sidx[tid], 1.0f);
    }
    float csum = 0.f;
    for (int i = tid; i < D * VQ_TM; i += VQ_THREADS) {
        int d, vl;
        if (lay.sD == 1) { vl = i / D; d = i - vl * D; }
        else { d = i / VQ_TM; vl = i - d * VQ_TM; }
        const int v = v0 + vl;
        if (v < N) {
            const int k = sidx[vl];
            const float q = __ldg(E + (size_t)k * D + d);
            const float xv = Xs[vq_xs_index<SWZ>(d, vl)];
            const float diff = q - xv;
            csum += diff * diff;
            if (q_out) {
                const int b = v / lay.Nn, n = v - b * lay.Nn;
                q_out[(size_t)(b * lay.sB + d * lay.sD + n * lay.sN)] = straight_through ? (xv + diff) : q;
            }
            if (embed_sum) atomicAdd(embed_sum + (size_t)k * D + d, xv);
        }
    }
    if (commit_partial) {
        __shared__ float red[VQ_THREADS / 32];
        csum = warp_sum(csum);
        if ((tid & 31) == 0) red[tid >> 5] = csum;
        __syncthreads();
        if (tid < 32) {
            float s = tid < VQ_THREADS / 32 ? red[tid] : 0.f;
            s = warp_sum(s);
            if (tid == 0) commit_partial[blockIdx.x] = s;
        }
    }
}

__global__ void __launch_bounds__(VQ_THREADS) vq_argmin_kernel(const float* __restrict__ x, VqLayout lay, int N, int D, const float* __restrict__ E,
                                                               const float* __restrict__ ee, int K, int64_t* __restrict__ idx_out,
                                                               float* __restrict__ q_out, int straight_through, float* __restrict__ commit_partial,
                                                               float* __restrict__ hist, float* __restrict__ embed_sum) {
    extern __shared__ float vq_smem[];
    float* Xs = vq_smem;                         // [D][VQ_TM]
    float* Es = Xs + (size_t)D * VQ_TM;          // [2][VQ_DC][VQ_TN]
    float* xx = Es + 2 * VQ_DC * VQ_TN;          // [VQ_TM]
    int* sidx = reinterpret_cast<int*>(xx + VQ_TM);   // [VQ_TM]
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int v0 = blockIdx.x * VQ_TM;

    vq_load_x_tile<false>(x, lay, N, D, v0, Xs, xx, tid);

    float best[4]; int besti[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { best[i] = -INFINITY; besti[i] = 0; }
    const int nchunks = (D + VQ_DC - 1) / VQ_DC;

    for (int ct = 0; ct < K; ct += VQ_TN) {
        float acc[4][8];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
        for (int ch = 0; ch < nchunks; ++ch) {
            float* es = Es + (ch & 1) * VQ_DC * VQ_TN;
            // load codebook chunk [16 dims][128 codes] (transposing read of E[K, D])
            for (int i = tid; i < VQ_DC * VQ_TN; i += VQ_THREADS) {
                const int dd = i & (VQ_DC - 1), c = i >> 4;
                const int d = ch * VQ_DC + dd, k = ct + c;
                es[dd * VQ_TN + c] = (d < D && k < K) ? __ldg(E + (size_t)k * D + d) : 0.f;
            }
            __syncthreads();
#pragma unroll
            for (int dd = 0; dd < VQ_DC; ++dd) {
                const int d = ch * VQ_DC + dd;
                if (d < D) {
                    const float4 xv = *reinterpret_cast<const float4*>(Xs + d * VQ_TM + ty * 4);
                    const float4 e0 = *reinterpret_cast<const float4*>(es + dd * VQ_TN + tx * 8);
                    const float4 e1 = *reinterpret_cast<const float4*>(es + dd * VQ_TN + tx * 8 + 4);
                    const float xa[4] = {xv.x, xv.y, xv.z, xv.w};
                    const float ea[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(xa[i], ea[j], acc[i][j]);
                }
            }
            // double-buffered Es: the next chunk writes the other half, so one barrier per chunk is enough
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = ct + tx * 8 + j;
            if (k < K) {
                const float e2 = __ldg(ee + k);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float dist = -((xx[ty * 4 + i] - 2.0f * acc[i][j]) + e2);
                    if (dist > best[i]) { best[i] = dist; besti[i] = k; }      // strict >: earlier (lower) index wins
                }
            }
        }
    }
    vq_finish_tile<false>(best, besti, lay, N, D, E, v0, Xs, sidx, idx_out, q_out, straight_through, commit_partial, hist, embed_sum, tid);
}

// Pipelined form (default when D % 16 == 0 and E is 16-byte aligned; TTTS_VQ_V1=1 selects the kernel above).  The kernel above reads every
// codebook chunk with a synchronous, transposing gather (8 scalar L2 loads per thread -> st.shared -> barrier -> 512 FMAs): the FMA pipe idles
// for an L2 round trip per chunk unless another CTA covers it (measured 19.6 TFLOP/s = 27 % of the fp32 peak at N = 2^20, r1g bench).
// Here the chunk [128 codes x 16 dims] is copied as it lies in E (64 contiguous bytes per code) with two 16-byte cp.async per thread into a
// double buffer, one chunk ahead of the FMAs; the tile is stored code-major with a 20-float row pitch, thread tx owns codes tx, tx + 16, ...
// so that its float4 reads along the dims are bank-conflict free.  Accumulation order over d is unchanged: distances and indices are
// bit-identical to the kernel above.
constexpr int VQ_EP = VQ_DC + 4;      // row pitch (floats) of the code-major chunk

__global__ void __launch_bounds__(VQ_THREADS, 3) vq_argmin_pipe_kernel(const float* __restrict__ x, VqLayout lay, int N, int D, const float* __restrict__ E,
                                                                    const float* __restrict__ ee, int K, int64_t* __restrict__ idx_out,
                                                                    float* __restrict__ q_out, int straight_through,
                                                                    float* __restrict__ commit_partial, float* __restrict__ hist,
                                                                    float* __restrict__ embed_sum) {
    extern __shared__ __align__(16) float vq_smem[];
    float* Xs = vq_smem;                         // [D][VQ_TM], 16-byte groups of a row XOR-swizzled by d (vq_xs_index)
    float* Es = Xs + (size_t)D * VQ_TM;          // [2][VQ_TN][VQ_EP]
    float* xx = Es + 2 * VQ_TN * VQ_EP;          // [VQ_TM]
    int* sidx = reinterpret_cast<int*>(xx + VQ_TM);   // [VQ_TM]
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int v0 = blockIdx.x * VQ_TM;
    const int nchunks = D / VQ_DC;
    const int ntiles = (K + VQ_TN - 1) / VQ_TN;
    const int total = ntiles * nchunks;

    // chunk q = (code tile q / nchunks, dims (q % nchunks) * 16 ...): thread copies 16 bytes of codes c and c + 64
    auto issue = [&](int q, int ct, int d0) {
        float* es = Es + (q & 1) * VQ_TN * VQ_EP;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int c = (tid >> 2) + h * 64, part = (tid & 3) * 4;
            const bool ok = ct + c < K;
            cp_async16(es + c * VQ_EP + part, ok ? E + (size_t)(ct + c) * D + d0 + part : E, ok);
        }
    };
    issue(0, 0, 0);
    cp_async_commit();

    vq_load_x_tile<true>(x, lay, N, D, v0, Xs, xx, tid);

    float best[4]; int besti[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { best[i] = -INFINITY; besti[i] = 0; }
    float acc[4][8];
    int ch = 0, ct = 0;                          // chunk q = dims ch * 16 ... of code tile ct (no division in the loop)
    for (int q = 0; q < total; ++q) {
        if (ch == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
        }
        cp_async_wait<0>();                      // chunk q has landed (this thread's part)
        __syncthreads();                         // ... everybody's part; and all threads are done with chunk q - 1 (the buffer refilled next)
        if (q + 1 < total) { const bool wrap = ch + 1 == nchunks; issue(q + 1, wrap ? ct + VQ_TN : ct, wrap ? 0 : (ch + 1) * VQ_DC); }
        cp_async_commit();
        const float* es = Es + (q & 1) * VQ_TN * VQ_EP;
#pragma unroll
        for (int d4 = 0; d4 < VQ_DC; d4 += 4) {
            float4 ev[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) ev[j] = *reinterpret_cast<const float4*>(es + (j * 16 + tx) * VQ_EP + d4);
#pragma unroll
            for (int dd = 0; dd < 4; ++dd) {
                const float4 xv = *reinterpret_cast<const float4*>(Xs + (ch * VQ_DC + d4 + dd) * VQ_TM + ((ty ^ (d4 + dd)) << 2));
                const float xa[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float e = dd == 0 ? ev[j].x : dd == 1 ? ev[j].y : dd == 2 ? ev[j].z : ev[j].w;
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc[i][j] = fmaf(xa[i], e, acc[i][j]);
                }
            }
        }
        if (ch == nchunks - 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {        // ascending k within the thread: strict > keeps the lowest index among equals
                const int k = ct + j * 16 + tx;
                if (k < K) {
                    const float e2 = __ldg(ee + k);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float dist = -((xx[ty * 4 + i] - 2.0f * acc[i][j]) + e2);
                        if (dist > best[i]) { best[i] = dist; besti[i] = k; }
                    }
                }
            }
            ch = 0; ct += VQ_TN;
        } else {
            ++ch;
        }
    }
    vq_finish_tile<true>(best, besti, lay, N, D, E, v0, Xs, sidx, idx_out, q_out, straight_through, commit_partial, hist, embed_sum, tid);
}

__global__ void __launch_bounds__(1024) vq_commit_final_kernel(const float* __restrict__ partial, int n, float scale, float* __restrict__ out) {
    __shared__ float sm[32];
    float s = 0.f;
    for (int i = threadIdx.x; i < n; i += 1024) s += partial[i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = warp_sum(sm[threadIdx.x]);
        if (threadIdx.x == 0) out[0] = s * scale;
    }
}

// cluster_size <- decay*cs + (1-decay)*hist ; total = sum(cs)            core_vq.py:217, 46-47
__global__ void __launch_bounds__(1024) vq_ema_counts_kernel(float* __restrict__ cluster_size, const float* __restrict__ hist, int K, float decay,
                                                             float* __restrict__ total) {
    __shared__ float sm[32];
    float s = 0.f;
    for (int k = threadIdx.x; k < K; k += 1024) {
        const float c = cluster_size[k] * decay + (1.0f - decay) * hist[k];
        cluster_size[k] = c;
        s += c;
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = warp_sum(sm[threadIdx.x]);
        if (threadIdx.x == 0) total[0] = s;
    }
}
// embed_avg <- decay*ea + (1-decay)*embed_sum ; embed <- ea / (laplace(cs) * total)      core_vq.py:218-228, 50-51
__global__ void vq_ema_embed_kernel(float* __restrict__ embed, float* __restrict__ embed_avg, const float* __restrict__ embed_sum,
                                    const float* __restrict__ cluster_size, const float* __restrict__ total, int K, int D, float decay, float eps) {
    const int k = blockIdx.x;
    const float tot = total[0];
    const float smoothed = (cluster_size[k] + eps) / (tot + (float)K * eps) * tot;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        const size_t i = (size_t)k * D + d;
        const float ea = embed_avg[i] * decay + (1.0f - decay) * embed_sum[i];
        embed_avg[i] = ea;
        embed[i] = ea / smoothed;
    }
}

// dx = dquantized + dcommit * 2 (x - q) / (N*D)       (straight-through + commitment loss, core_vq.py:311-318)
__global__ void vq_bwd_kernel(const float* __restrict__ x, VqLayout lay, int N, int D, const float* __restrict__ E, const int64_t* __restrict__ idx,
                              const float* __restrict__ dq, const float* __restrict__ dcommit, float* __restrict__ dx) {
    const size_t total = (size_t)N * D;
    const float g = (dcommit ? dcommit[0] : 0.f) * 2.0f / (float)total;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        // iterate in memory order of the [B, D, Nn] / [N, D] tensor
        int v, d;
        if (lay.sD == 1) { v = (int)(i / D); d = (int)(i - (size_t)v * D); }
        else { const int b = (int)(i / ((size_t)D * lay.Nn)); const int r = (int)(i - (size_t)b * D * lay.Nn); d = r / lay.Nn; v = b * lay.Nn + (r - d * lay.Nn); }
        const int b = v / lay.Nn, n = v - b * lay.Nn;
        const size_t off = (size_t)(b * lay.sB + d * lay.sD + n * lay.sN);
        const float q = __ldg(E + (size_t)idx[v] * D + d);
        dx[off] = (dq ? dq[off] : 0.f) + g * (x[off] - q);
    }
}

static VqLayout make_layout(int B, int D, int Nn, int bdn) {
    VqLayout l;
    if (bdn) { l.sB = (long long)D * Nn; l.sD = Nn; l.sN = 1; l.Nn = Nn; }
    else { l.sB = D; l.sD = 1; l.sN = 0; l.Nn = 1; }
    return l;
}

}  // namespace ttts

using namespace ttts;

extern "C" {

// workspace floats needed by ttts_vq_forward: K (code norms) + ceil(N/64) (commit partials) + 1
int64_t ttts_vq_workspace_floats(int32_t N, int32_t K) { return (int64_t)K + (N + VQ_TM - 1) / VQ_TM + 8; }

int ttts_vq_forward(const float* x, int32_t B, int32_t D, int32_t Nn, int32_t layout_bdn, const float* embed, int32_t K, int64_t* codes,
                    float* quantized, int32_t straight_through, float* commit_out, float* hist, float* embed_sum, float* workspace, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    TTTS_CHECK_ARG(x && embed && codes && workspace, "vq: null pointer");
    TTTS_CHECK_ARG(B > 0 && D > 0 && Nn > 0 && K > 0, "vq: bad shape");
    TTTS_CHECK_ARG(D % 4 == 0 || true, "vq: D");
    const int N = layout_bdn ? B * Nn : B;
    VqLayout lay = make_layout(B, D, layout_bdn ? Nn : 1, layout_bdn);
    float* ee = workspace;
    float* partial = workspace + K;
    const int blocks = (N + VQ_TM - 1) / VQ_TM;
    vq_code_norms_kernel<<<(K + 7) / 8, 256, 0, st>>>(embed, K, D, ee);
    TTTS_LAUNCH_CHECK("vq_code_norms");
    static int v1 = -1;
    if (v1 < 0) { const char* e = getenv("TTTS_VQ_V1"); v1 = (e && e[0] == '1') ? 1 : 0; }
    const bool pipe = !v1 && D % VQ_DC == 0 && (reinterpret_cast<uintptr_t>(embed) & 15) == 0;
    const size_t smem = ((size_t)D * VQ_TM + (pipe ? 2 * VQ_TN * VQ_EP : 2 * VQ_DC * VQ_TN) + VQ_TM) * sizeof(float) + VQ_TM * sizeof(int);
    TTTS_CHECK_ARG(smem <= 200 * 1024, "vq: D too large for the shared-memory x tile");
    static size_t attr_smem = 0, attr_smem_pipe = 0;
    if (pipe) {
        if (smem > attr_smem_pipe) {
            TTTS_CUDA(cudaFuncSetAttribute(vq_argmin_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_smem_pipe = smem;
        }
        vq_argmin_pipe_kernel<<<blocks, VQ_THREADS, smem, st>>>(x, lay, N, D, embed, ee, K, codes, quantized, straight_through,
                                                                commit_out ? partial : nullptr, hist, embed_sum);
    } else {
        if (smem > attr_smem) {
            TTTS_CUDA(cudaFuncSetAttribute(vq_argmin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_smem = smem;
        }
        vq_argmin_kernel<<<blocks, VQ_THREADS, smem, st>>>(x, lay, N, D, embed, ee, K, codes, quantized, straight_through, commit_out ? partial : nullptr,
                                                           hist, embed_sum);
    }
    TTTS_LAUNCH_CHECK("vq_argmin");
    if (commit_out) {
        vq_commit_final_kernel<<<1, 1024, 0, st>>>(partial, blocks, 1.0f / ((float)N * (float)D), commit_out);
        TTTS_LAUNCH_CHECK("vq_commit_final");
    }
    return TTTS_OK;
}

int ttts_vq_ema_update(float* embed, float* embed_avg, float* cluster_size, const float* hist, const float* embed_sum, int32_t K, int32_t D,
                       float decay, float eps, float* scratch1, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    TTTS_CHECK_ARG(embed && embed_avg && cluster_size && hist && embed_sum && scratch1, "vq_ema: null pointer");
    vq_ema_counts_kernel<<<1, 1024, 0, st>>>(cluster_size, hist, K, decay, scratch1);
    TTTS_LAUNCH_CHECK("vq_ema_counts");
    vq_ema_embed_kernel<<<K, 64, 0, st>>>(embed, embed_avg, embed_sum, cluster_size, scratch1, K, D, decay, eps);
    TTTS_LAUNCH_CHECK("vq_ema_embed");
    return TTTS_OK;
}

int ttts_vq_backward(const float* x, int32_t B, int32_t D, int32_t Nn, int32_t layout_bdn, const float* embed, const int64_t* codes,
                     const float* dquantized, const float* dcommit, float* dx, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    TTTS_CHECK_ARG(x && embed && codes && dx, "vq_bwd: null pointer");
    const int N = layout_bdn ? B * Nn : B;
    VqLayout lay = make_layout(B, D, layout_bdn ? Nn : 1, layout_bdn);
    const size_t total = (size_t)N * D;
    int blocks = (int)((total + 255) / 256);
    if (blocks > num_sms() * 8) blocks = num_sms() * 8;
    vq_bwd_kernel<<<blocks, 256, 0, st>>>(x, lay, N, D, embed, codes, dquantized, dcommit, dx);
    TTTS_LAUNCH_CHECK("vq_bwd");
    return TTTS_OK;
}
}
