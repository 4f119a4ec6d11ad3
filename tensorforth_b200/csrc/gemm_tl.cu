// gemm_tl.cu — the LAYER GEMM: single-launch FP32 GEMM on the 5th-gen tensor cores for the linear layers' products at batch
// 64-4096 (forward X @ W^T, dW = dY^T @ X, dX = dY @ W: 0.02-2 GFLOP), 3xTF32, latency first.
//   replaces k_gemm_tile_claude (src/t4math.cu:478-583) as launched by Tensor::linear / Model::_flinear / _blinear
//   (src/mu/tensor.cu:74-87, src/nn/forward.cu:158-198, src/nn/backprop.cu:194-254), with the epilogues of k_bias
//   (src/nn/nmath.cu:27-35), k_activate (nmath.cu:37-70), the small head linear + k_softmax_small (nmath.cu:74-118)
//   and the activation backward (backprop.cu:257-263) fused in.
//
// What is different from gemm_tcf.cu (whose ~10 us of fixed cost made it tie with the FP32-FMA kernel at these sizes):
//   * operands arrive by TMA (cp.async.bulk.tensor.2d of the RAW FP32 tiles, SWIZZLE_128B, out-of-range rows/cols/k zero filled by
//     the TMA unit): one elected thread keeps a whole ring (3 x 32 KiB) in flight, nothing is staged through registers;
//   * ANY transposition is native: an operand that is contiguous along M/N is loaded as [k][32 x m] boxes and handed to the tensor
//     core as an MN-major UMMA operand (instruction-descriptor bits 15/16; 32-bit MN-major operands use the SWIZZLE_128B_BASE32B
//     layout = TMA's 128B_ATOM_32B swizzle: LBO = 4 KiB between 32-wide m blocks, SBO = 512 B between 4-row k groups) — dW (both
//     operands MN-major) and dX (B MN-major) no longer pay scattered 4-byte shared stores;
//   * the raw FP32 plane IS the hi operand: kind::tf32 reads the upper 19 bits of each word (the 13 low mantissa bits are ignored:
//     hi = truncate(a)); eight converter warps only derive the lo plane, lo = tf32(a - hi), elementwise on the swizzled image
//     (same offset in a second plane, whatever the layout); product = lo*hi + hi*lo + hi*hi with FP32 accumulation in TMEM;
//   * split-K lives in a THREAD-BLOCK CLUSTER (1 x 1 x S, S <= 16): every CTA parks its accumulator tile in its own shared memory and
//     CTA r reduces rows [r*128/S, (r+1)*128/S) of all S tiles in rank order over distributed shared memory — no partials in HBM, no
//     finish launch, deterministic — and applies the epilogue there, a warp per output row:
//       mode 0  O = alpha * acc + beta * O                                  (Tensor::mm / gemm words, dW with beta = 1)
//       mode 1  Y = acc + bias ; A = act(Y) ; F = saved derivative / mask   (Model::_flinear + _factivate)
//       mode 2  mode 1, then Y2 = A @ W2^T + B2 ; P = softmax(Y2) (+ dup)   (... + the classifier head: the row never leaves the warp)
//       mode 3  O = acc ; O2 = acc * F                                      (Model::_blinear's dX + the _bactivate in front of it)
// Bound: launch + one HBM/L2 round trip + a handful of MMAs; at the largest layer shapes (0.8 GFLOP) shared-memory bandwidth
// (operand reads of the SS MMAs + the lo pass).
#include "tc_ptx.cuh"
#include "act.cuh"
#include <cuda.h>
#include <cstdlib>

namespace t4k {

constexpr int L_BM = 128, L_BN = 128, L_BK = 32, L_UK = 8;
constexpr uint32_t L_PLANE_B = 128u * L_BK * 4u;         // 16 KiB: one raw (= hi) or lo plane of a 128 x 32 operand tile
constexpr uint32_t L_OP_B = 2u * L_PLANE_B;              // raw + lo
constexpr uint32_t L_STAGE_B = 2u * L_OP_B;              // A + B = 64 KiB
constexpr int L_STAGES = 3;
constexpr int L_DRAIN_KB = 8;                            // k-blocks per accumulator chain (gemm_tc.cu: DRAIN_KB)
constexpr int L_NCONV = 6, L_NEPI = 8;                   // 16 warps: register files are granted per 4 warps, 512 threads leave 128 registers each
constexpr int L_WARPS = 2 + L_NCONV + L_NEPI;            // 16
constexpr int L_THREADS = L_WARPS * 32;                  // 512
constexpr int L_CONV_T = L_NCONV * 32;                   // 192 converter threads
constexpr int L_W2_FLTS = 32 * 128;                      // head weights [E2 <= 32][EH <= 128] in shared memory (mode 2, generated A)
constexpr int L_D_FLTS = 128 * 32;                       // generated A: p - y rows of the tile [128][E2 <= 32]

struct TlP {
    float *O; float alpha, beta;
    int M, N, K;
    int KT, kt_per;              // k-blocks of 32: total, per cluster rank
    int a_mn, b_mn;              // operand is contiguous along M / N in memory (tA / !tB)
    int mask_hi;                 // debug: store the masked hi back over the raw plane (does not rely on the MMA ignoring the low bits)
    uint32_t mn_type, mn_lbo, mn_sbo, mn_kstep;      // MN-major operand descriptor: layout type, LBO / SBO / start-address step per 8 k (bytes)
    int mode;
    const float *bias; float *actA, *actF; int layer; float act_alpha;          // mode 1, 2
    const float *W2, *B2; float *Y2, *P, *P2; int E2;                          // mode 2
    const float *F; float *O2;                                                  // mode 3
    // generated A operand (gen != 0; A is K-major [M][K], K <= 128): A[m][k] = (Σ_j (gP[m][j] - gT[m][j]) * gW2[j][k]) * gF[m][k] — the classifier
    // head's backward (Model::_bprep + the small linear's dX + the activation backward, backprop.cu:76-140,194-263) evaluated in the
    // converter warps instead of being loaded: the dX GEMM of the hidden linear layer no longer waits for a head-backward kernel
    int gen, gE2; const float *gP, *gT, *gW2, *gF;
    float *part;                 // split-K partials [cluster][rank][128][128] in the library workspace (L2-resident); nullptr: reduce over distributed shared memory
    long long *trace;            // bring-up: clock64 stamps of CTA (0,0,0)'s phases (t4k_gemm_tl_trace), nullptr in production
};
#define TL_TRACE(ev) do { if (p.trace && (blockIdx.x | blockIdx.y | blockIdx.z) == 0) p.trace[ev] = clock64(); } while (0)

__device__ __forceinline__ void tl_tma_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void tl_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tl_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float4 tl_ld_dsmem4(uint32_t local_saddr, uint32_t rank) {        // the same shared-memory offset in CTA `rank` of the cluster
    uint32_t r; float4 v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_saddr), "r"(rank));
    asm volatile("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(r) : "memory");
    return v;
}
// parked accumulator tile [128 rows][128 cols] fp32 over the (idle) operand ring: 16-byte chunk c4 of row r lives at chunk (c4 ^ (r & 31)):
// the epilogue warps' per-row float4 stores (lanes = rows) and the reduction's per-row reads (lanes = chunks) are both conflict-free
__device__ __forceinline__ int tl_park_off(int row, int c4) { return row * L_BN + ((c4 ^ (row & 31)) << 2); }

// shared-memory matrix descriptors (cute::UMMA::SmemDescriptor), version 1
//   K-major : SWIZZLE_128B (type 2: 16-byte chunks XOR row & 7); rows of 128 B (32 tf32 along k), 8-row groups 1024 B apart (SBO); a k-step
//             of 8 advances the start address by 32 B
//   MN-major: 32-bit operands have ONE legal layout, SWIZZLE_128B_BASE32B (type 1: 32-byte chunks XOR row & 3 — the TMA mode
//             CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): rows of 128 B (32 tf32 along m/n) at one k, 4 k-rows = one 512 B atom (SBO = 512 B
//             between k groups), 32-wide m/n blocks LBO = 4 KiB apart (one TMA box of 32 k-rows); a k-step of 8 = two atoms = 1024 B
__device__ __forceinline__ uint64_t tl_desc(uint32_t saddr, bool mn, const TlP &p) {
    const uint64_t lbo = mn ? (p.mn_lbo >> 4) : 1u, sbo = mn ? (p.mn_sbo >> 4) : (1024u >> 4), type = mn ? p.mn_type : 2u;
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (lbo << 16) | (sbo << 32) | (1ull << 46) | (type << 61);
}

template<int L> __device__ __forceinline__ void tl_act4(const float (&y)[4], float alpha, float (&a)[4], float (&f)[4]) {
    #pragma unroll
    for (int e = 0; e < 4; e++) act<L>(y[e], alpha, a[e], f[e]);
}
__device__ __forceinline__ void tl_act(int layer, const float (&y)[4], float alpha, float (&a)[4], float (&f)[4]) {
    switch (layer) {
    case T4K_L_RELU:    tl_act4<T4K_L_RELU>(y, alpha, a, f); break;
    case T4K_L_TANH:    tl_act4<T4K_L_TANH>(y, alpha, a, f); break;
    case T4K_L_SIGMOID: tl_act4<T4K_L_SIGMOID>(y, alpha, a, f); break;
    case T4K_L_SELU:    tl_act4<T4K_L_SELU>(y, alpha, a, f); break;
    case T4K_L_LEAKYRL: tl_act4<T4K_L_LEAKYRL>(y, alpha, a, f); break;
    case T4K_L_ELU:     tl_act4<T4K_L_ELU>(y, alpha, a, f); break;
    default:            tl_act4<T4K_L_DROPOUT>(y, alpha, a, f); break;
    }
}
// transposing warp reduction (nn.cu): lane k ends up with Σ_lanes v[k]
__device__ __forceinline__ float tl_treduce32(float (&v)[32], int lane) {
    #pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        const bool up = (lane & o) != 0;
        #pragma unroll
        for (int i = 0; i < o; i++) {
            const float send = up ? v[i] : v[i + o];
            const float keep = up ? v[i + o] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
    }
    return v[0];
}

__global__ void __launch_bounds__(L_THREADS, 1) k_gemm_tl(const TlP p, const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap bmap) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);      // SWIZZLE_128B tiles: 1024-byte aligned
    uint64_t *bars = (uint64_t*)(smem + L_STAGES * L_STAGE_B);                        // raw_full[S], lo_full[S], empty[S], acc_full[2], acc_empty[2]
    uint32_t *tmem_slot = (uint32_t*)(bars + 3 * L_STAGES + 4);
    float *sW2 = reinterpret_cast<float*>(smem + L_STAGES * L_STAGE_B + 256);         // mode 2 / generated A: head weights
    float *sD = sW2 + L_W2_FLTS;                                                      // generated A: p - y
    pdl_wait(); pdl_trigger();
    if (threadIdx.x == 0) TL_TRACE(0);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mt = blockIdx.y, nt = blockIdx.x, zs = blockIdx.z, S = (int)gridDim.z;
    const int kt0 = zs * p.kt_per;
    const int kt1 = min(p.KT, kt0 + p.kt_per);
    const int nkb = max(0, kt1 - kt0);
    const int nchunk = (nkb + L_DRAIN_KB - 1) / L_DRAIN_KB;
    const uint32_t raw0 = smem_u32(bars), lo0 = smem_u32(bars + L_STAGES), empty0 = smem_u32(bars + 2 * L_STAGES);
    const uint32_t afull0 = smem_u32(bars + 3 * L_STAGES), aempty0 = smem_u32(bars + 3 * L_STAGES + 2);

    // one k-block's TMA loads (raw FP32 tiles) into stage s; complete on raw_full[s]
    auto tma_issue = [&](int i) {
        const int s = i % L_STAGES;
        const uint32_t sa = smem_u32(smem + (size_t)s * L_STAGE_B), sb = sa + L_OP_B, bar = raw0 + 8 * s;
        const int m0 = mt * L_BM, n0 = nt * L_BN, k0 = (kt0 + i) * L_BK;
        mbar_expect_tx(bar, p.gen ? L_PLANE_B : 2 * L_PLANE_B);
        if (i == 0) TL_TRACE(2);
        if (i == nkb - 1) TL_TRACE(6);
        if (p.gen) { /* A is written by the converter warps */ }
        else if (!p.a_mn) tl_tma_2d(sa, &amap, k0, m0, bar);                          // box {32 k, 128 rows}
        else {
            #pragma unroll
            for (int j = 0; j < 4; j++) tl_tma_2d(sa + j * 4096u, &amap, m0 + 32 * j, k0, bar);              // box {32 m, 32 k}
        }
        if (!p.b_mn) tl_tma_2d(sb, &bmap, k0, n0, bar);
        else {
            #pragma unroll
            for (int j = 0; j < 4; j++) tl_tma_2d(sb + j * 4096u, &bmap, n0 + 32 * j, k0, bar);
        }
    };
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" :: "l"(&amap) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(&bmap) : "memory");
        for (int s = 0; s < L_STAGES; s++) { mbar_init(raw0 + 8 * s, 1); mbar_init(lo0 + 8 * s, L_NCONV); mbar_init(empty0 + 8 * s, 1); }
        for (int b = 0; b < 2; b++) { mbar_init(afull0 + 8 * b, 1); mbar_init(aempty0 + 8 * b, L_NEPI); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // the first ring of loads leaves before the set-up barrier (TMEM allocation, head weights): their latency overlaps it
        for (int i = 0; i < nkb && i < L_STAGES; i++) tma_issue(i);
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 2 * L_BN);                         // two accumulators of 128 fp32 columns
    if (p.mode == 2 && warp >= 2) {                                                   // head weights: asynchronous copies, published by the barrier in front of the reduction
        const int EH = p.N, tot = p.E2 * EH;
        for (int t = threadIdx.x - 64; t < tot; t += L_THREADS - 64) cp_async4(sW2 + t, p.W2 + t, true);
        cp_async_commit();
    }
    if (p.gen && warp >= 2) {                                                         // generated A: W2 [E2][K] and d = p - y of the tile's rows
        const int tot = p.gE2 * p.K;
        for (int t = threadIdx.x - 64; t < tot; t += L_THREADS - 64) cp_async4(sW2 + t, p.gW2 + t, true);
        cp_async_commit();
        const int nd = L_BM * p.gE2, m0 = blockIdx.y * L_BM;
        for (int t = threadIdx.x - 64; t < nd; t += L_THREADS - 64) {
            const int r = t / p.gE2, j = t - r * p.gE2, gr = m0 + r;
            sD[t] = (gr < p.M) ? __fsub_rn(__ldg(p.gP + (int64_t)gr * p.gE2 + j), __ldg(p.gT + (int64_t)gr * p.gE2 + j)) : 0.0f;
        }
        cp_async_wait_all();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) TL_TRACE(1);

    if (warp == 0) {
        // ===== TMA producer: raw FP32 tiles, the whole ring in flight =====
        if (lane == 0) {
            for (int i = L_STAGES; i < nkb; i++) {
                const int s = i % L_STAGES, it = i / L_STAGES;
                mbar_wait(empty0 + 8 * s, (it & 1) ^ 1);
                tma_issue(i);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        int nv = p.N - nt * L_BN; if (nv > L_BN) nv = L_BN;
        const int un = (nv + 15) & ~15;                                               // UMMA N: the valid columns of this tile, rounded to 16
        const uint32_t idesc = idesc_tf32(L_BM, un) | ((uint32_t)(p.a_mn ? 1 : 0) << 15) | ((uint32_t)(p.b_mn ? 1 : 0) << 16);
        const uint64_t ka = p.a_mn ? (p.mn_kstep >> 4) : (32u >> 4), kb = p.b_mn ? (p.mn_kstep >> 4) : (32u >> 4);      // start-address step per 8 k
        for (int i = 0; i < nkb; i++) {
            const int s = i % L_STAGES, it = i / L_STAGES;
            const int c = i / L_DRAIN_KB, ib = i % L_DRAIN_KB, b = c & 1;
            if (ib == 0 && c >= 2) { mbar_wait(aempty0 + 8 * b, ((c >> 1) - 1) & 1); tc_fence_after(); }      // chunk c-2 drained
            mbar_wait(raw0 + 8 * s, it & 1);
            mbar_wait(lo0 + 8 * s, it & 1);
            tc_fence_after();
            if (lane == 0 && i == 0) TL_TRACE(5);
            if (lane == 0 && i == nkb - 1) TL_TRACE(8);
            if (elect_one()) {
                const uint32_t acc = tmem_base + (uint32_t)(b * L_BN);
                const uint32_t sa = smem_u32(smem + (size_t)s * L_STAGE_B), sb = sa + L_OP_B;
                const uint64_t a_hi = tl_desc(sa, p.a_mn, p), a_lo = tl_desc(sa + L_PLANE_B, p.a_mn, p);
                const uint64_t b_hi = tl_desc(sb, p.b_mn, p), b_lo = tl_desc(sb + L_PLANE_B, p.b_mn, p);
                #pragma unroll
                for (int k = 0; k < L_BK / L_UK; k++) {
                    tc_mma_tf32(acc, a_lo + k * ka, b_hi + k * kb, idesc, (ib | k) ? 1u : 0u);
                    tc_mma_tf32(acc, a_hi + k * ka, b_lo + k * kb, idesc, 1u);
                    tc_mma_tf32(acc, a_hi + k * ka, b_hi + k * kb, idesc, 1u);
                }
            }
            __syncwarp();
            if (elect_one()) {
                tc_commit(empty0 + 8 * s);
                if (ib == L_DRAIN_KB - 1 || i == nkb - 1) tc_commit(afull0 + 8 * b);
            }
            __syncwarp();
        }
    } else if (warp < 2 + L_NCONV) {
        // ===== converters: lo = tf32(a - truncate(a)) on the swizzled image, same offset in the second plane =====
        const int t = threadIdx.x - 64;                                               // 0..191
        constexpr int NV = (2048 + L_CONV_T - 1) / L_CONV_T;                           // 16-byte words of a stage's two raw planes per thread (11)
        for (int i = 0; i < nkb; i++) {
            const int s = i % L_STAGES, it = i / L_STAGES;
            if (p.gen) {
                // the A tile of this k-block, computed: thread -> (row, 16-byte chunk); 8 threads write one 128-byte row (swizzled: conflict-free)
                mbar_wait(empty0 + 8 * s, (it & 1) ^ 1);                              // the stage's previous MMAs are done
                uint8_t *abase = smem + (size_t)s * L_STAGE_B;
                const int k0 = (kt0 + i) * L_BK, c = t & 7, k = k0 + 4 * c, E2 = p.gE2, EH = p.K;
                #pragma unroll 1
                for (int r = t >> 3; r < L_BM; r += L_CONV_T / 8) {
                    const int gr = mt * L_BM + r;
                    float x[4] = {0.0f, 0.0f, 0.0f, 0.0f};
                    if (gr < p.M && k < EH) {                                         // EH % 4 == 0 (host check): whole chunks
                        for (int j = 0; j < E2; j++) {                                // class order as k_head_bwd: the same bits as the stored dY1
                            const float dj = sD[r * E2 + j];
                            const float4 w = *reinterpret_cast<const float4*>(sW2 + j * EH + k);
                            x[0] = fmaf(dj, w.x, x[0]); x[1] = fmaf(dj, w.y, x[1]); x[2] = fmaf(dj, w.z, x[2]); x[3] = fmaf(dj, w.w, x[3]);
                        }
                        if (p.gF) {
                            const float4 f = ldg4(p.gF + (int64_t)gr * EH + k);
                            x[0] = __fmul_rn(x[0], f.x); x[1] = __fmul_rn(x[1], f.y); x[2] = __fmul_rn(x[2], f.z); x[3] = __fmul_rn(x[3], f.w);
                        }
                    }
                    float hi[4], lo[4];
                    #pragma unroll
                    for (int e = 0; e < 4; e++) { hi[e] = __uint_as_float(__float_as_uint(x[e]) & 0xFFFFE000u); lo[e] = to_tf32(x[e] - hi[e]); }
                    uint8_t *q = abase + (size_t)r * 128 + (size_t)((c ^ (r & 7)) << 4);
                    *reinterpret_cast<float4*>(q) = make_float4(x[0], x[1], x[2], x[3]);
                    *reinterpret_cast<float4*>(q + L_PLANE_B) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                }
            }
            mbar_wait(raw0 + 8 * s, it & 1);
            if (t == 0 && i == 0) TL_TRACE(3);
            if (t == 0 && i == nkb - 1) TL_TRACE(7);
            uint8_t *base = smem + (size_t)s * L_STAGE_B;
            float4 v[NV];
            #pragma unroll
            for (int j = 0; j < NV; j++) {
                const int idx = t + L_CONV_T * j + (p.gen ? 1024 : 0);                 // generated A: only the B planes are converted
                if (idx < 2048) v[j] = *reinterpret_cast<const float4*>(base + (size_t)(idx >> 10) * L_OP_B + (size_t)(idx & 1023) * 16);
            }
            #pragma unroll
            for (int j = 0; j < NV; j++) {
                const int idx = t + L_CONV_T * j + (p.gen ? 1024 : 0);
                if (idx >= 2048) break;
                uint8_t *q = base + (size_t)(idx >> 10) * L_OP_B + (size_t)(idx & 1023) * 16;
                const float x[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
                float hi[4], lo[4];
                #pragma unroll
                for (int e = 0; e < 4; e++) {
                    hi[e] = __uint_as_float(__float_as_uint(x[e]) & 0xFFFFE000u);
                    lo[e] = to_tf32(x[e] - hi[e]);
                }
                *reinterpret_cast<float4*>(q + L_PLANE_B) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                if (p.mask_hi) *reinterpret_cast<float4*>(q) = make_float4(hi[0], hi[1], hi[2], hi[3]);
            }
            tl_fence_proxy_async();                      // generic-proxy stores → visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(lo0 + 8 * s);
            if (t == 0 && i == 0) TL_TRACE(4);
        }
    } else {
        // ===== epilogue warps: drain the accumulator chains into registers (round-to-nearest adds), park the tile =====
        constexpr int CW = L_BN / 2;
        const int q = warp & 3, h = (warp - (2 + L_NCONV)) >> 2;
        float acc[CW];
        #pragma unroll
        for (int j = 0; j < CW; j++) acc[j] = 0.0f;
        for (int c = 0; c < nchunk; c++) {
            const int b = c & 1;
            mbar_wait(afull0 + 8 * b, (c >> 1) & 1);
            tc_fence_after();
            if (c == nchunk - 1 && warp == 2 + L_NCONV && lane == 0) TL_TRACE(9);
            #pragma unroll
            for (int gq = 0; gq < CW / 16; gq++) {
                uint32_t v[16];
                tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * L_BN + h * CW + gq * 16), v);
                tmem_ld_wait();
                #pragma unroll
                for (int j = 0; j < 16; j++) acc[gq * 16 + j] += __uint_as_float(v[j]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(aempty0 + 8 * b);
        }
        // the operand ring is idle: the last accumulator commit covers every MMA that read it, every TMA write was consumed
        float *park = reinterpret_cast<float*>(smem);
        const int r = q * 32 + lane;
        #pragma unroll
        for (int j = 0; j < CW; j += 4)
            *reinterpret_cast<float4*>(park + tl_park_off(r, (h * CW + j) >> 2)) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
        if (warp == 2 + L_NCONV && lane == 0) TL_TRACE(10);
    }
    if (p.mode == 2) cp_async_wait_all();
    tc_fence_before();
    __syncthreads();
    // split-K partials through L2 (default): distributed shared memory serves ~20 B/clk per SM — 60 KiB of peer tiles cost 1.6 us and every CTA
    // has to stay resident until its last reader is done (measured, profiles/r02_tl_trace.txt); the L2 takes the same bytes at 3x the rate
    // and nobody waits at the exit.  The tile goes out coalesced (a warp per row), the cluster barrier (release/acquire) publishes it.
    float *mypart = nullptr;
    if (S > 1 && p.part) {
        mypart = p.part + ((size_t)(blockIdx.y * gridDim.x + blockIdx.x) * S) * (size_t)(L_BM * L_BN);
        const float *park = reinterpret_cast<const float*>(smem);
        float *dst = mypart + (size_t)zs * (L_BM * L_BN);
        const int rows = min(L_BM, p.M - mt * L_BM);
        if (nt * L_BN + lane * 4 < p.N)
            for (int r = warp; r < rows; r += L_WARPS)
                __stcg(reinterpret_cast<float4*>(dst + r * L_BN + lane * 4), *reinterpret_cast<const float4*>(park + tl_park_off(r, lane)));
    }
    if (S > 1) tl_cluster_sync();                                                     // every CTA's tile is parked / stored and visible cluster-wide
    if (threadIdx.x == 0) TL_TRACE(11);

    // ===== reduction over the cluster + epilogue: CTA `zs` finishes rows [zs*rpr, (zs+1)*rpr) of the tile, a warp per row, lane = 16-byte chunk =====
    {
        const int rpr = L_BM / S;
        const uint32_t park_s = smem_u32(smem);
        const float *park = reinterpret_cast<const float*>(smem);
        const int gc = nt * L_BN + lane * 4;
        const bool o_vec = ((p.N & 3) == 0);
        for (int rr = warp; rr < rpr; rr += L_WARPS) {
            const int r = zs * rpr + rr, gr = mt * L_BM + r;
            if (gr >= p.M) break;                                                     // rows ascend with rr
            float4 sum;
            if (S == 1) sum = *reinterpret_cast<const float4*>(park + tl_park_off(r, lane));
            else {
                float4 v[16];
                if (mypart) {
                    const float *src = mypart + r * L_BN + lane * 4;
                    const bool ld = gc < p.N;
                    #pragma unroll
                    for (int qk = 0; qk < 16; qk++) if (qk < S) v[qk] = ld ? __ldcg(reinterpret_cast<const float4*>(src + (size_t)qk * (L_BM * L_BN))) : make_float4(0.f, 0.f, 0.f, 0.f);
                } else {
                    const uint32_t a = park_s + (uint32_t)tl_park_off(r, lane) * 4u;
                    #pragma unroll
                    for (int qk = 0; qk < 16; qk++) if (qk < S) v[qk] = tl_ld_dsmem4(a, (uint32_t)qk);
                }
                sum = v[0];
                #pragma unroll
                for (int qk = 1; qk < 16; qk++) if (qk < S) { sum.x += v[qk].x; sum.y += v[qk].y; sum.z += v[qk].z; sum.w += v[qk].w; }
            }
            if (threadIdx.x == 0 && rr == 0) TL_TRACE(12);
            const bool in = gc < p.N;
            const bool full = o_vec && gc + 3 < p.N;
            const int64_t at = (int64_t)gr * p.N + gc;
            float out[4] = {sum.x, sum.y, sum.z, sum.w};
            if (p.mode == 0) {
                #pragma unroll
                for (int e = 0; e < 4; e++) out[e] *= p.alpha;
                if (full) {
                    if (p.beta != 0.0f) { const float4 old = *reinterpret_cast<const float4*>(p.O + at); out[0] += old.x * p.beta; out[1] += old.y * p.beta; out[2] += old.z * p.beta; out[3] += old.w * p.beta; }
                    stg4(p.O + at, make_float4(out[0], out[1], out[2], out[3]));
                } else if (in) {
                    #pragma unroll
                    for (int e = 0; e < 4; e++) if (gc + e < p.N) p.O[at + e] = (p.beta != 0.0f) ? out[e] + p.O[at + e] * p.beta : out[e];
                }
            } else if (p.mode == 3) {
                if (full) {
                    stg4(p.O + at, make_float4(out[0], out[1], out[2], out[3]));
                    const float4 f = ldg4(p.F + at);
                    stg4(p.O2 + at, make_float4(__fmul_rn(out[0], f.x), __fmul_rn(out[1], f.y), __fmul_rn(out[2], f.z), __fmul_rn(out[3], f.w)));
                } else if (in) {
                    #pragma unroll
                    for (int e = 0; e < 4; e++) if (gc + e < p.N) { p.O[at + e] = out[e]; p.O2[at + e] = __fmul_rn(out[e], p.F[at + e]); }
                }
            } else {
                // linear layer epilogue (k_linear_fin's arithmetic: Σ splits, + bias, activation)
                float y[4] = {0.0f, 0.0f, 0.0f, 0.0f}, a[4] = {0.0f, 0.0f, 0.0f, 0.0f}, f[4] = {0.0f, 0.0f, 0.0f, 0.0f};
                if (in) {
                    #pragma unroll
                    for (int e = 0; e < 4; e++) if (gc + e < p.N) y[e] = out[e] + __ldg(p.bias + gc + e);
                    if (p.layer == T4K_L_DROPOUT) {
                        #pragma unroll
                        for (int e = 0; e < 4; e++) if (gc + e < p.N) f[e] = p.actF[at + e];
                    }
                    if (p.layer != T4K_L_NONE) tl_act(p.layer, y, p.act_alpha, a, f);
                    else { a[0] = y[0]; a[1] = y[1]; a[2] = y[2]; a[3] = y[3]; }
                    if (full) {
                        stg4(p.O + at, make_float4(y[0], y[1], y[2], y[3]));
                        if (p.layer != T4K_L_NONE) { stg4(p.actA + at, make_float4(a[0], a[1], a[2], a[3])); stg4(p.actF + at, make_float4(f[0], f[1], f[2], f[3])); }
                    } else {
                        #pragma unroll
                        for (int e = 0; e < 4; e++) if (gc + e < p.N) {
                            p.O[at + e] = y[e];
                            if (p.layer != T4K_L_NONE) { p.actA[at + e] = a[e]; p.actF[at + e] = f[e]; }
                        }
                    }
                }
                if (p.mode == 2) {
                    // classifier head on the finished row (one n-tile: the lane's four hidden units are columns gc..gc+3)
                    const int EH = p.N, E2 = p.E2;
                    float hacc[32];
                    #pragma unroll
                    for (int k = 0; k < 32; k++) hacc[k] = 0.0f;
                    if (gc < EH) {                                                     // EH % 4 == 0 (host check): 128-bit conflict-free reads of W2 rows
                        #pragma unroll
                        for (int k = 0; k < 32; k++) if (k < E2) {
                            const float4 w = *reinterpret_cast<const float4*>(sW2 + k * EH + gc);
                            hacc[k] = fmaf(a[0], w.x, hacc[k]); hacc[k] = fmaf(a[1], w.y, hacc[k]);
                            hacc[k] = fmaf(a[2], w.z, hacc[k]); hacc[k] = fmaf(a[3], w.w, hacc[k]);
                        }
                    }
                    float y2 = tl_treduce32(hacc, lane);
                    const bool on = lane < E2;
                    if (on) y2 += __ldg(p.B2 + lane);
                    const float mx = warp_max(on ? y2 : -FLT_MAX);                    // k_softmax_small (nmath.cu:74-118): exp(x - max) / Σ
                    const float ex = on ? __expf(y2 - mx) : 0.0f;
                    const float sm = warp_sum(ex);
                    if (on) {
                        const float pv = ex / sm; const int64_t o2 = (int64_t)gr * E2 + lane;
                        p.Y2[o2] = y2; p.P[o2] = pv; if (p.P2) p.P2[o2] = pv;
                    }
                }
            }
        }
    }
    if (threadIdx.x == 0) TL_TRACE(13);
    if (S > 1 && !mypart) tl_cluster_sync();                                          // distributed shared memory: nobody leaves while its tile is still being read
    if (threadIdx.x == 0) TL_TRACE(14);
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 2 * L_BN); }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*PFN_tl_encode)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_tl_encode tl_encoder() {
    static PFN_tl_encode enc = nullptr;
    if (!enc) {
        void *fp = nullptr; cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qr) != cudaSuccess || !fp) { cudaGetLastError(); return nullptr; }
        enc = (PFN_tl_encode)fp;
    }
    return enc;
}
// 2-D FP32 matrix [outer][inner] (row pitch = inner floats), box = box_outer x 32 floats, 128-byte swizzle (16-byte chunks, or 32-byte
// chunks for the MN-major operands), zero fill out of bounds
static int tl_map(CUtensorMap *m, const float *X, int64_t inner, int64_t outer, int box_outer, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
    PFN_tl_encode enc = tl_encoder();
    if (!enc) return T4K_ENOSUP;
    const cuuint64_t gdim[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
    const cuuint64_t gstr[1] = {(cuuint64_t)inner * 4};
    const cuuint32_t box[2] = {32, (cuuint32_t)box_outer}, estr[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)X, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS ? 0 : T4K_EINVAL;
}

constexpr size_t L_SMEM = (size_t)L_STAGES * L_STAGE_B + 256 + (size_t)(L_W2_FLTS + L_D_FLTS) * 4 + 1024;
static_assert(L_SMEM <= 227 * 1024, "shared memory budget");
#define TL_MAX_DEV 16
static int g_tl_maxcl[TL_MAX_DEV][5];                   // [device][log2 S]: co-resident clusters of size S (0: not queried yet, -1: unavailable)

static int tl_device() { const int d = cur_device(); return (d < 0 || d >= TL_MAX_DEV) ? -1 : d; }
static int tl_prepare(int dev) {
    static bool attr[TL_MAX_DEV];
    if (attr[dev]) return 0;
    cudaError_t e = cudaFuncSetAttribute(k_gemm_tl, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_gemm_tl, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) { cudaGetLastError(); return (int)e; }
    for (int l = 0; l < 5; l++) {
        const int S = 1 << l;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(1, 1, S); cfg.blockDim = dim3(L_THREADS); cfg.dynamicSmemBytes = L_SMEM;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = (unsigned)S;
        cfg.attrs = at; cfg.numAttrs = 1;
        int n = 0;
        if (S == 1) n = sm_count();
        else if (cudaOccupancyMaxActiveClusters(&n, k_gemm_tl, &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
        g_tl_maxcl[dev][l] = n > 0 ? n : -1;
    }
    attr[dev] = true;
    return 0;
}
static int tl_env(const char *name, int dflt) { const char *e = getenv(name); return e ? atoi(e) : dflt; }
// knobs (t4k_set_gemm_tl): 0 engine on/off for AUTO, 1 mask_hi (debug), 2 largest cluster size; 3-7 (bring-up probes only): MN-major
// descriptor layout type, TMA swizzle mode, LBO, SBO, start-address step per 8 k; 8: split-K partials through L2 (1, default) or reduced
// over distributed shared memory (0)
#define TL_NKNOB 9
static long long *g_tl_trace = nullptr;
static int g_tl_knob[TL_NKNOB] = {-1, -1, -1, -1, -1, -1, -1, -1, -1};
static int tl_knob(int k) {
    static const char *name[TL_NKNOB] = {"T4K_GEMM_TL", "T4K_TL_MASKHI", "T4K_TL_SMAX", "T4K_TL_MN_TYPE", "T4K_TL_MN_SWZ", "T4K_TL_MN_LBO", "T4K_TL_MN_SBO", "T4K_TL_MN_KSTEP", "T4K_TL_L2RED"};
    static const int dflt[TL_NKNOB] = {1, 0, 16, 1, (int)CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, 4096, 512, 1024, 1};
    if (g_tl_knob[k] < 0) g_tl_knob[k] = tl_env(name[k], dflt[k]);
    return g_tl_knob[k];
}

bool gemm_tl_ok(const float *A, const float *B, const float *O, int tA, int tB, int M, int N, int K, int C, int batch) {
    if (!tl_knob(0) || C != 1 || batch != 1 || M < 32 || N < 16 || K < 16) return false;
    if (!aligned16(A) || !aligned16(B) || !O) return false;
    if (((tA ? M : K) & 3) || ((tB ? K : N) & 3)) return false;               // TMA: row pitch a multiple of 16 bytes
    const double w = (double)M * N * K;
    return w >= 4.0e6 && w < 2.0e10;
}

int gemm_tl(const float *A, const float *B, float *O, float alpha, float beta, int tA, int tB, int M, int N, int K, cudaStream_t st, const TlEpi *epi) {
    const int dev = tl_device();
    if (dev < 0) return T4K_EINVAL;
    int rc = tl_prepare(dev); if (rc) return rc;
    const int mtiles = (M + L_BM - 1) / L_BM, ntiles = (N + L_BN - 1) / L_BN, KT = (K + L_BK - 1) / L_BK, T = mtiles * ntiles;
    if (epi && epi->mode == 2 && (ntiles != 1 || (N & 3) || epi->E2 > 32 || epi->E2 < 1)) return T4K_ENOSUP;
    // cluster size = split-K factor: the largest power of two that keeps the whole grid in one wave of co-resident clusters and
    // leaves every rank at least one k-block (two when there is a choice)
    int S = 1;
    const int smax = tl_knob(2);
    for (int l = 4; l >= 1; l--) {
        const int s = 1 << l, cap = g_tl_maxcl[dev][l];
        if (s > smax || cap <= 0 || T > cap) continue;
        const int per = (KT + s - 1) / s;
        if ((KT + per - 1) / per != s) continue;                               // an empty rank
        if (per < 2 && l > 1 && KT >= 4) continue;
        S = s; break;
    }
    const int kt_per = (KT + S - 1) / S;
    const int mask_hi = tl_knob(1);
    TlP p{};
    p.O = O; p.alpha = alpha; p.beta = beta; p.M = M; p.N = N; p.K = K; p.KT = KT; p.kt_per = kt_per;
    p.a_mn = tA ? 1 : 0; p.b_mn = tB ? 0 : 1; p.mask_hi = mask_hi; p.mode = 0;
    p.mn_type = (uint32_t)tl_knob(3); p.mn_lbo = (uint32_t)tl_knob(5); p.mn_sbo = (uint32_t)tl_knob(6); p.mn_kstep = (uint32_t)tl_knob(7);
    const CUtensorMapSwizzle mn_swz = (CUtensorMapSwizzle)tl_knob(4);
    p.trace = g_tl_trace;
    if (S > 1 && tl_knob(8)) {
        p.part = (float*)workspace((size_t)T * S * L_BM * L_BN * sizeof(float), 7);
        if (!p.part) return T4K_ENOMEM;
    }
    if (epi) {
        p.mode = epi->mode; p.bias = epi->bias; p.actA = epi->actA; p.actF = epi->actF; p.layer = epi->layer; p.act_alpha = epi->act_alpha;
        p.W2 = epi->W2; p.B2 = epi->B2; p.Y2 = epi->Y2; p.P = epi->P; p.P2 = epi->P2; p.E2 = epi->E2; p.F = epi->F; p.O2 = epi->O2;
        if (epi->gP) {
            if (tA || K > 128 || (K & 3) || epi->gE2 < 1 || epi->gE2 > 32 || !epi->gT || !epi->gW2) return T4K_ENOSUP;
            p.gen = 1; p.gE2 = epi->gE2; p.gP = epi->gP; p.gT = epi->gT; p.gW2 = epi->gW2; p.gF = epi->gF;
        }
    }
    CUtensorMap amap, bmap;
    // op(A)(m,k): A stored [M][K] (K-major: box 128 rows x 32 k) or [K][M] when tA (M-major: box 32 k-rows x 32 m)
    if (p.gen) rc = tl_map(&amap, B, tB ? K : N, tB ? N : K, tB ? 128 : 32, tB ? CU_TENSOR_MAP_SWIZZLE_128B : mn_swz);       // unused: a valid map
    else rc = tA ? tl_map(&amap, A, M, K, 32, mn_swz) : tl_map(&amap, A, K, M, 128);
    if (rc) return rc;
    // op(B)(k,n): B stored [N][K] when tB (K-major) or [K][N] (N-major)
    rc = tB ? tl_map(&bmap, B, K, N, 128) : tl_map(&bmap, B, N, K, 32, mn_swz); if (rc) return rc;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ntiles, mtiles, S); cfg.blockDim = dim3(L_THREADS); cfg.dynamicSmemBytes = L_SMEM; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = (unsigned)S;
    cfg.attrs = at; cfg.numAttrs = S > 1 ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, k_gemm_tl, p, amap, bmap);
    ++g_launches;
    if (e != cudaSuccess) { cudaGetLastError(); return (int)e; }
    return (int)cudaGetLastError();
}

} // namespace t4k

extern "C" int t4k_gemm_tl_trace(long long *dev16) { t4k::g_tl_trace = dev16; return 0; }   // bring-up: 16 clock64 stamps of CTA (0,0,0), nullptr = off
extern "C" int t4k_set_gemm_tl(int what, int value) {
    if (what < 0 || what >= TL_NKNOB) return T4K_EINVAL;
    const int was = t4k::tl_knob(what);
    t4k::g_tl_knob[what] = value;
    return was;
}
