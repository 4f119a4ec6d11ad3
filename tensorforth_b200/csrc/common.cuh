// common.cuh — shared device/host helpers for libt4k (sm_100a only)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <float.h>
#include "../../include/t4k.h"

#define T4K_SMS          148                  // B200: 148 SMs (grid sizing; real count read at runtime)
#define T4K_THREADS      256
#define DU_EPS           1.0e-6f              // src/ten4_types.h:85
#define DU_LNX           1.0e-12f             // src/t4math.cu:172

namespace t4k {

extern long g_launches;                       // counted kernel launches (t4k_launch_count)
int  sm_count();                              // cudaDevAttrMultiProcessorCount of the CURRENT device (cached per device)
int  cur_device();                            // cudaGetDevice, -1 on error
// function attributes (max dynamic shared memory, non-portable cluster sizes) are PER DEVICE: "set once" flags are kept per device
struct DevFlag { bool f[16]; };
static inline bool dev_first(DevFlag &x) { const int d = cur_device(); if (d < 0 || d >= 16) return true; if (x.f[d]) return false; x.f[d] = true; return true; }
struct DevSize { size_t v[16]; };
static inline bool dev_grow(DevSize &x, size_t s) { const int d = cur_device(); if (d < 0 || d >= 16) return true; if (s <= x.v[d]) return false; x.v[d] = s; return true; }
int  check_launch();                          // cudaGetLastError() → rc, ++g_launches
void *workspace(size_t bytes, int slot);      // library-owned per-device scratch (grown on demand)
float *reduce_slot(cudaStream_t st);          // 4 KiB partials + counter, ring of slots, zeroed counter

struct GemmDeferred { const float *part; int splits; };   // gemm_simt(..., defer): caller-side split-K finish (splits == 0: nothing left to do, see GemmEpilogue)
// bias + activation epilogue a GEMM engine MAY apply itself (only the cluster variant of gemm_tcf does): Y = product + bias (written to O),
// A = act(Y), F = saved derivative / mask.  An engine that applied it reports defer->splits = 0.
struct GemmEpilogue { const float *bias; float *A, *F; int layer; float alpha; };
int gemm_simt(const float *A, const float *B, float *O, float alpha, float beta, int tA, int tB,
              int M, int N, int K, int C, int batch, int64_t sA, int64_t sB, int64_t sO, cudaStream_t st, GemmDeferred *defer = nullptr);

// mid-size single-launch tensor-core GEMM (gemm_tcf.cu): in-kernel 3xTF32 split, split-K; same deferred-finish contract
bool gemm_tcf_ok(int tA, int tB, int M, int N, int K, int C, int batch);
int  gemm_tcf(const float *A, const float *B, float *O, float alpha, float beta, int tA, int tB,
              int M, int N, int K, cudaStream_t st, GemmDeferred *defer = nullptr, const GemmEpilogue *epi = nullptr);

// the layer GEMM (gemm_tl.cu): TMA-fed 3xTF32 on tcgen05, any transposition native, split-K in a thread-block cluster reduced over
// distributed shared memory, epilogue fused (TlEpi.mode: 0 alpha/beta, 1 bias + activation, 2 bias + activation + classifier head,
// 3 dX and dX * F).  One launch, nothing deferred.
struct TlEpi {
    int mode; const float *bias; float *actA, *actF; int layer; float act_alpha;       // mode 1, 2: Y = acc + bias (to O), A = act(Y), F = derivative / mask
    const float *W2, *B2; float *Y2, *P, *P2; int E2;                                 // mode 2: Y2 = A @ W2^T + B2 [M,E2], P = softmax(Y2), P2 = copy of P (may be null)
    const float *F; float *O2;                                                        // mode 3: O = acc, O2 = acc * F
    const float *T; float *Ylin, *hpart;                                              // mode 4 (train tail, see gemm_tl.cu): target [M,E2], head linear's output tensor, partials [ctas][nEp]
    // generated A operand (gP != nullptr; the A argument of gemm_tl is ignored): with dY1[n][e] = (Σ_j (gP[n][j] - gT[n][j]) * gW2[j][e]) * gF[n][e]
    // (gF may be null, j < gE2 <= 32, e < EH <= 128, EH % 4 == 0):  tA == 0: A[m][k] = dY1[m][k] (EH = K);  tA == 1: A(m,k) = dY1[k][m] (EH = M)
    const float *gP, *gT, *gW2, *gF; int gE2;
};
struct TlJob { const float *A, *B; float *O; float alpha, beta; int tA, tB, M, N, K; const TlEpi *epi; };
// one launch for one or two independent problems (e.g. dW and dX of a layer): T4K_ENOSUP when they do not fit one co-resident wave
int  gemm_tl_multi(const TlJob *jobs, int njobs, cudaStream_t st, int *ctas_out = nullptr);
int  gemm_tl_pair_inplace(const TlJob *jobs, cudaStream_t st);   // two problems, jobs[0] stores over an operand jobs[1] reads (its stores wait for those reads)
int  gemm_tl_ctas(const TlJob *jobs, int njobs);          // CTAs gemm_tl_multi would launch for these problems (<= 0: not supported)
bool gemm_tl_ok(const float *A, const float *B, const float *O, int tA, int tB, int M, int N, int K, int C, int batch);
int  gemm_tl(const float *A, const float *B, float *O, float alpha, float beta, int tA, int tB,
             int M, int N, int K, cudaStream_t st, const TlEpi *epi = nullptr);

static inline cudaStream_t STRM(t4k_stream_t s) { return (cudaStream_t)s; }

// ---- programmatic dependent launch (PDL).  A train step is a chain of ~12 short dependent kernels; with a plain
// launch each one pays the previous grid's drain + its own launch/ramp latency (2-3 us per boundary).  Kernels launched
// with launch_pdl() may be scheduled as soon as every CTA of the previous grid has passed pdl_trigger(); they run their
// prologue (index math, shared-memory halos) and block in pdl_wait() until the previous grid has completed and its
// memory is visible.  Rule kept everywhere: NO global access before pdl_wait().  Works in eager streams and under
// CUDA-graph capture (programmatic edges).  Off by default (see below); T4K_PDL=1 / t4k_set_pdl(1) turns it on.
extern int g_pdl;
__device__ __forceinline__ void pdl_wait()    { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// Measured on the MNIST step (B200, 200 steps): the attribute on EVERY kernel made the step slower (100.3 -> 108.5 us):
// early-resident CTAs of a machine-filling kernel (GEMM, conv block) pile onto the first SMs the previous grid frees and
// the grid starts unbalanced (linear_bwd 24.6 -> 33.9 us), while the short kernels gain 0.3-0.5 us each.  So only the
// short, latency-bound kernels are launched with it (launch_pdl); machine-filling ones use launch_std (they still execute
// pdl_trigger, so the short kernel behind them is scheduled early).
// Shared-memory carve-out.  An SM changes its L1 / shared-memory split only when it is EMPTY: a short kernel that asks for no shared memory
// gets the smallest split, and while its CTAs are resident the big-shared-memory kernels of another stream (fused conv blocks, layer GEMM) cannot
// be placed on that SM — the side-stream branches of a step (loss, head gradients, optimizer) would push the critical path's kernels off the SMs
// they touch.  The short kernels are therefore launched asking for the LARGEST shared-memory split (they stream through L2, L1 size is irrelevant
// to them); T4K_CARVEOUT=-1 turns the attribute off, 0..100 picks another percentage.
extern int g_carve;
template<bool PDL, typename... KA, typename... A>
static inline void launch_k(void (*kern)(KA...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[2];
    int na = 0;
    if (PDL && g_pdl) { at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[na].val.programmaticStreamSerializationAllowed = 1; na++; }
    if (g_carve >= 0) { at[na].id = cudaLaunchAttributePreferredSharedMemoryCarveout; at[na].val.sharedMemCarveout = (unsigned)g_carve; na++; }
    cfg.attrs = at; cfg.numAttrs = na;
    cudaLaunchKernelEx(&cfg, kern, KA(args)...);
}
template<typename... KA, typename... A>
static inline void launch_pdl(void (*kern)(KA...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A... args) { launch_k<true>(kern, grid, block, smem, st, args...); }
template<typename... KA, typename... A>
static inline void launch_std(void (*kern)(KA...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A... args) { launch_k<false>(kern, grid, block, smem, st, args...); }

// persistent-ish grid for HBM streaming kernels: enough CTAs to cover n items at `per_thread`
// each, capped at 8 resident CTAs of 256 threads per SM (148 x 8 = 1184).
static inline int stream_grid(int64_t n_items, int per_thread = 1) {
    int64_t need = (n_items + (int64_t)T4K_THREADS * per_thread - 1) / ((int64_t)T4K_THREADS * per_thread);
    int64_t cap  = (int64_t)sm_count() * 8;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}
__host__ __device__ static inline bool aligned16(const void *p) { return (((uintptr_t)p) & 15) == 0; }

__device__ __forceinline__ float warp_sum(float v) {
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// block-wide sum of `v` (blockDim.x multiple of 32, <= 1024); result valid in thread 0
__device__ __forceinline__ float block_sum(float v, float *sm32) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_sum(v);
    if (lane == 0) sm32[w] = v;
    __syncthreads();
    if (w == 0) {
        v = (lane < (int)(blockDim.x >> 5)) ? sm32[lane] : 0.0f;
        v = warp_sum(v);
    }
    __syncthreads();
    return v;
}
// ------------------------------------------------------------------ asynchronous global -> shared copies (LDGSTS)
__device__ __forceinline__ void cp_async16(float *dst, const float *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
// 4-byte copy; on == false writes a zero (src-size 0) and reads nothing
__device__ __forceinline__ void cp_async4(float *dst, const float *src, bool on) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" :: "r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src), "r"(on ? 4 : 0) : "memory");
}
__device__ __forceinline__ void cp_async_commit()   { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// 128-bit streaming global access
__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void   stg4(float *p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

} // namespace t4k
