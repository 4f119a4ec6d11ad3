// gemm_tl.cu — the LAYER GEMM: single-launch FP32 GEMM on the 5th-gen tensor cores for the linear layers' products at batch
// 64-4096 (forward X @ W^T, dW = dY^T @ X, dX = dY @ W: 0.02-2 GFLOP), 3xTF32, latency first.
//   replaces k_gemm_tile_claude (src/t4math.cu:478-583) as launched by Tensor::linear / Model::_flinear / _blinear
//   (src/mu/tensor.cu:74-87, src/nn/forward.cu:158-198, src/nn/backprop.cu:194-254), with the epilogues of k_bias
//   (src/nn/nmath.cu:27-35), k_activate (nmath.cu:37-70), the small head linear + k_softmax_small (nmath.cu:74-118)
//   and the activation backward (backprop.cu:257-263) fused in.
//
// What is different from gemm_tcf.cu (whose ~10 us of fixed cost made it tie with the FP32-FMA kernel at these sizes):
//   * operands arrive by TMA (cp.async.bulk.tensor.2d of the RAW FP32 tiles, out-of-range rows/cols/k zero filled by the TMA unit): one
//     elected thread keeps a ring of FOUR k-blocks (4 x 32 KiB) in flight, nothing is staged through registers;
//   * ANY transposition is native: an operand that is contiguous along M/N is loaded as [k][32 x m] boxes and handed to the tensor
//     core as an MN-major UMMA operand (instruction-descriptor bits 15/16; 32-bit MN-major operands use the SWIZZLE_128B_BASE32B
//     layout = TMA's 128B_ATOM_32B swizzle: LBO = 4 KiB between 32-wide m blocks, SBO = 512 B between 4-row k groups) — dW (both
//     operands MN-major) and dX (B MN-major) no longer pay scattered 4-byte shared stores;
//   * the raw FP32 plane IS the hi operand: kind::tf32 reads the upper 19 bits of each word (the 13 low mantissa bits are ignored:
//     hi = truncate(a), verified bit-for-bit on the device); six converter warps only derive the lo plane, lo = tf32(a - hi), elementwise on
//     the swizzled image (same offset in a second ring of TWO slots, whatever the layout); product = lo*hi + hi*lo + hi*hi, FP32 in TMEM;
//   * an operand can be GENERATED instead of loaded: A = ((P - T) @ W2) * F — the classifier head's backward (Model::_bprep, the small
//     linear's dX and the activation backward, backprop.cu:76-140,194-263) evaluated by the converter warps, K-major (the dX GEMM of the
//     hidden layer) or MN-major (its dW GEMM): neither waits for a head-backward kernel any more;
//   * split-K lives in a THREAD-BLOCK CLUSTER (S <= 16 CTAs): every CTA parks its accumulator tile in shared memory, writes it to an
//     L2-resident workspace, and after the cluster barrier CTA r adds rows [r*128/S, (r+1)*128/S) of the S tiles in rank order —
//     deterministic, no finish launch — and applies the epilogue there, a warp per output row:
//       mode 0  O = alpha * acc + beta * O                                  (Tensor::mm / gemm words, dW with beta = 1)
//       mode 1  Y = acc + bias ; A = act(Y) ; F = saved derivative / mask   (Model::_flinear + _factivate)
//       mode 2  mode 1, then Y2 = A @ W2^T + B2 ; P = softmax(Y2) (+ dup)   (... + the classifier head: the row never leaves the warp)
//       mode 3  O = acc ; O2 = acc * F                                      (Model::_blinear's dX + the _bactivate in front of it)
//       mode 4  the TRAIN TAIL: mode 2, and on the same row, still in the warp, the head's backward (Model::_bprep p - y, the head linear's
//               dX and the activation backward): the layer tensors receive their BACKWARD values directly (inside a fused train step the
//               forward values are never observable), the head's parameter gradients leave as per-CTA partials (k_head_grad_fin adds them)
//   * TWO problems can share one launch (dW and dX of a layer when X has a duplicate): a cluster is either the K-split of one tile or S
//     independent tiles, so both fill the machine together instead of queueing behind each other (this kernel owns its SM: 512 threads x
//     128 registers, 225 KiB of shared memory — nothing else is co-resident).
// Measured timeline of a CTA (profiles/r02_tl_trace.txt): set-up 0.4 us, first tile landed at 1.2 us, 0.65 us per k-block (shared-memory
// bound: the lo pass and the operand reads of the SS MMAs), reduction 3.5 us.
#include "tc_ptx.cuh"
#include "act.cuh"
#include <cuda.h>
#include <cstdlib>

namespace t4k {

constexpr int L_BM = 128, L_BN = 128, L_BK = 32, L_UK = 8;
constexpr uint32_t L_PLANE_B = 128u * L_BK * 4u;         // 16 KiB: one raw (= hi) or lo plane of a 128 x 32 operand tile
constexpr uint32_t L_SLOT_B = 2u * L_PLANE_B;            // A plane + B plane = 32 KiB: one slot of the raw ring or of the lo ring
constexpr int L_RAW = 4, L_LO = 2;                       // ring depths: raw tiles in flight / lo planes
constexpr uint32_t L_RING_B = (L_RAW + L_LO) * L_SLOT_B; // 192 KiB
constexpr int L_DRAIN_KB = 8;                            // k-blocks per accumulator chain (gemm_tc.cu: DRAIN_KB)
constexpr int L_NCONV = 6, L_NEPI = 8;                   // 16 warps: register files are granted per 4 warps, 512 threads leave 128 registers each
constexpr int L_WARPS = 2 + L_NCONV + L_NEPI;            // 16
constexpr int L_THREADS = L_WARPS * 32;                  // 512
constexpr int L_CONV_T = L_NCONV * 32;                   // 192 converter threads
constexpr int L_W2_FLTS = 32 * 128;                      // head weights [E2 <= 32][EH <= 128] in shared memory (mode 2, generated A)
constexpr int L_D_FLTS = 128 * 32;                       // generated A: p - y rows [128][E2 <= 32]
constexpr int L_NBAR = 2 * L_RAW + 2 * L_LO + 4;         // raw_full, raw_empty, lo_full, lo_empty, acc_full[2], acc_empty[2]

struct TlP {
    float *O; float alpha, beta;
    int M, N, K;
    int KT, kt_per, split;       // k-blocks of 32: total, per rank; ranks per tile (1 or the cluster size)
    int mtiles, ntiles;
    int a_mn, b_mn;              // operand is contiguous along M / N in memory (tA / !tB)
    int mode;
    const float *bias; float *actA, *actF; int layer; float act_alpha;          // mode 1, 2
    const float *W2, *B2; float *Y2, *P, *P2; int E2;                          // mode 2
    const float *F; float *O2;                                                  // mode 3
    const float *T; float *Ylin, *hpart;                                        // mode 4 (train tail): target rows, the head linear's output tensor, per-CTA gradient partials
    // generated A operand (gen 1: K-major [M][K], K <= 128; gen 2: M-major, A(m,k) = the same matrix transposed, M <= 128):
    //   dY1[n][e] = (Σ_j (gP[n][j] - gT[n][j]) * gW2[j][e]) * gF[n][e];  gen 1: A[m=n][k=e];  gen 2: A[m=e][k=n]
    int gen, gE2; const float *gP, *gT, *gW2, *gF;
    float *part;                 // split-K partials [tile][rank][128][128] in the library workspace (L2-resident); nullptr: reduce over distributed shared memory
};
struct TlG {
    TlP p[2];
    int ncl0;                    // clusters of problem 0 (the rest run problem 1)
    int mask_hi;                 // debug: store the masked hi back over the raw plane (does not rely on the MMA ignoring the low bits)
    uint32_t mn_type, mn_lbo, mn_sbo, mn_kstep;      // MN-major operand descriptor: layout type, LBO / SBO / start-address step per 8 k (bytes)
    long long *trace;            // bring-up: clock64 stamps of CTA 0's phases (t4k_gemm_tl_trace), nullptr in production
    // in-place pair (problem 0 overwrites an operand of problem 1: dX = dY @ W stored over X while dW += dY^T @ X reads X): every CTA of problem 1
    // counts itself in xsync[0] once its LAST operand tile has landed in shared memory; the CTAs of problem 0 hold their stores until all
    // `nread` have (the grid is one co-resident wave, so nobody waits for a CTA that has not started); the last CTA out re-zeroes the counters
    unsigned *xsync; unsigned nread;
};
#define TL_TRACE(ev) do { if (g.trace && blockIdx.x == 0) g.trace[ev] = clock64(); } while (0)

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
// parked accumulator tile [128 rows][128 cols] fp32 over the (idle) raw ring: 16-byte chunk c4 of row r lives at chunk (c4 ^ (r & 31)):
// the epilogue warps' per-row float4 stores (lanes = rows) and the reduction's per-row reads (lanes = chunks) are both conflict-free
__device__ __forceinline__ int tl_park_off(int row, int c4) { return row * L_BN + ((c4 ^ (row & 31)) << 2); }

// shared-memory matrix descriptors (cute::UMMA::SmemDescriptor), version 1
//   K-major : SWIZZLE_128B (type 2: 16-byte chunks XOR row & 7); rows of 128 B (32 tf32 along k), 8-row groups 1024 B apart (SBO); a k-step
//             of 8 advances the start address by 32 B
//   MN-major: 32-bit operands have ONE legal layout, SWIZZLE_128B_BASE32B (type 1: 32-byte chunks XOR row & 3 — the TMA mode
//             CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): rows of 128 B (32 tf32 along m/n) at one k, 4 k-rows = one 512 B atom (SBO = 512 B
//             between k groups), 32-wide m/n blocks LBO = 4 KiB apart (one TMA box of 32 k-rows); a k-step of 8 = two atoms = 1024 B
__device__ __forceinline__ uint64_t tl_desc(uint32_t saddr, bool mn, const TlG &g) {
    const uint64_t lbo = mn ? (g.mn_lbo >> 4) : 1u, sbo = mn ? (g.mn_sbo >> 4) : (1024u >> 4), type = mn ? g.mn_type : 2u;
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

__global__ void __launch_bounds__(L_THREADS, 1) k_gemm_tl(const __grid_constant__ TlG g,
                                                          const __grid_constant__ CUtensorMap amap0, const __grid_constant__ CUtensorMap bmap0,
                                                          const __grid_constant__ CUtensorMap amap1, const __grid_constant__ CUtensorMap bmap1) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);      // swizzled tiles: 1024-byte aligned
    uint8_t *lo_ring = smem + L_RAW * L_SLOT_B;
    uint64_t *bars = (uint64_t*)(smem + L_RING_B);
    uint32_t *tmem_slot = (uint32_t*)(bars + L_NBAR);
    float *sW2 = reinterpret_cast<float*>(smem + L_RING_B + 256);                     // mode 2 / generated A: head weights
    float *sD = sW2 + L_W2_FLTS;                                                      // generated A: p - y
    pdl_wait(); pdl_trigger();
    if (threadIdx.x == 0) TL_TRACE(0);

    // ---- which problem, which tile, which k range: cluster `cl` is the K-split of ONE tile (split == S) or S tiles side by side (split == 1)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t S, rk;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(S));
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rk));
    const int cl = (int)(blockIdx.x / S);
    const bool second = cl >= g.ncl0;
    const TlP &p = g.p[second ? 1 : 0];
    const CUtensorMap *amap = second ? &amap1 : &amap0, *bmap = second ? &bmap1 : &bmap0;
    const int split = p.split, tpc = (int)S / split;                                  // tiles per cluster
    const int tile = (second ? cl - g.ncl0 : cl) * tpc + (int)rk / split;
    const int zs = (int)rk % split, rk0 = (int)rk - zs;                               // rank inside the tile's group, first cluster rank of the group
    const bool active = tile < p.mtiles * p.ntiles;
    const int mt = active ? tile / p.ntiles : 0, nt = active ? tile % p.ntiles : 0;
    const int kt0 = zs * p.kt_per;
    const int kt1 = min(p.KT, kt0 + p.kt_per);
    const int nkb = active ? max(0, kt1 - kt0) : 0;
    const int nchunk = (nkb + L_DRAIN_KB - 1) / L_DRAIN_KB;
    const uint32_t rawf0 = smem_u32(bars), rawe0 = rawf0 + 8 * L_RAW, lof0 = rawe0 + 8 * L_RAW, loe0 = lof0 + 8 * L_LO;
    const uint32_t afull0 = loe0 + 8 * L_LO, aempty0 = afull0 + 16;
    const int gen = active ? p.gen : 0;

    // one k-block's TMA loads (raw FP32 tiles) into raw slot i % L_RAW; complete on raw_full
    auto tma_issue = [&](int i) {
        const int s = i % L_RAW;
        const uint32_t sa = smem_u32(smem + (size_t)s * L_SLOT_B), sb = sa + L_PLANE_B, bar = rawf0 + 8 * s;
        const int m0 = mt * L_BM, n0 = nt * L_BN, k0 = (kt0 + i) * L_BK;
        mbar_expect_tx(bar, gen ? L_PLANE_B : 2 * L_PLANE_B);
        if (i == 0) TL_TRACE(2);
        if (i == nkb - 1) TL_TRACE(6);
        if (gen) { /* A is written by the converter warps */ }
        else if (!p.a_mn) tl_tma_2d(sa, amap, k0, m0, bar);                           // box {32 k, 128 rows}
        else {
            #pragma unroll
            for (int j = 0; j < 4; j++) tl_tma_2d(sa + j * 4096u, amap, m0 + 32 * j, k0, bar);               // box {32 m, 32 k}
        }
        if (!p.b_mn) tl_tma_2d(sb, bmap, k0, n0, bar);
        else {
            #pragma unroll
            for (int j = 0; j < 4; j++) tl_tma_2d(sb + j * 4096u, bmap, n0 + 32 * j, k0, bar);
        }
    };
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" :: "l"(amap) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(bmap) : "memory");
        for (int s = 0; s < L_RAW; s++) { mbar_init(rawf0 + 8 * s, 1); mbar_init(rawe0 + 8 * s, 1); }
        for (int s = 0; s < L_LO; s++) { mbar_init(lof0 + 8 * s, L_NCONV); mbar_init(loe0 + 8 * s, 1); }
        for (int b = 0; b < 2; b++) { mbar_init(afull0 + 8 * b, 1); mbar_init(aempty0 + 8 * b, L_NEPI); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // the first ring of loads leaves before the set-up barrier (TMEM allocation, head weights): their latency overlaps it
        for (int i = 0; i < nkb && i < L_RAW; i++) tma_issue(i);
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 2 * L_BN);                         // two accumulators of 128 fp32 columns
    if (active && (p.mode == 2 || p.mode == 4) && warp >= 2) {                                         // head weights: asynchronous copies, published by the barrier in front of the reduction
        const int EH = p.N, tot = p.E2 * EH;
        for (int t = threadIdx.x - 64; t < tot; t += L_THREADS - 64) cp_async4(sW2 + t, p.W2 + t, true);
        cp_async_commit();
    }
    // generated A: dY1 = ((P - T) @ W2) * F with dY1 [Ng][EHg]; gen 1: A rows = dY1 rows of this M tile; gen 2: A(m,k) = dY1[k][m], rows of this k range
    const int EHg = (gen == 2) ? p.M : p.K, Ng = (gen == 2) ? p.K : p.M;
    const int drow0 = (gen == 2) ? kt0 * L_BK : mt * L_BM;                            // first dY1 row held in sD
    if (gen && warp >= 2) {
        const int tot = p.gE2 * EHg;
        for (int t = threadIdx.x - 64; t < tot; t += L_THREADS - 64) cp_async4(sW2 + t, p.gW2 + t, true);
        cp_async_commit();
        const int nd = L_BM * p.gE2;
        for (int t = threadIdx.x - 64; t < nd; t += L_THREADS - 64) {
            const int r = t / p.gE2, j = t - r * p.gE2, gr = drow0 + r;
            sD[t] = (gr < Ng) ? __fsub_rn(__ldg(p.gP + (int64_t)gr * p.gE2 + j), __ldg(p.gT + (int64_t)gr * p.gE2 + j)) : 0.0f;
        }
        cp_async_wait_all();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) TL_TRACE(1);

    if (g.xsync && second && nkb == 0 && threadIdx.x == 0) atomicAdd(g.xsync, 1u);               // nothing to read: counted at once
    if (warp == 0) {
        // ===== TMA producer: raw FP32 tiles, four k-blocks in flight =====
        if (lane == 0) {
            for (int i = L_RAW; i < nkb; i++) {
                mbar_wait(rawe0 + 8 * (i % L_RAW), ((i / L_RAW) & 1) ^ 1);
                tma_issue(i);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        int nv = p.N - nt * L_BN; if (nv > L_BN) nv = L_BN;
        const int un = (nv + 15) & ~15;                                               // UMMA N: the valid columns of this tile, rounded to 16
        const uint32_t idesc = idesc_tf32(L_BM, un) | ((uint32_t)(p.a_mn ? 1 : 0) << 15) | ((uint32_t)(p.b_mn ? 1 : 0) << 16);
        const uint64_t ka = p.a_mn ? (g.mn_kstep >> 4) : (32u >> 4), kb = p.b_mn ? (g.mn_kstep >> 4) : (32u >> 4);       // start-address step per 8 k
        for (int i = 0; i < nkb; i++) {
            const int s = i % L_RAW, sl = i % L_LO;
            const int c = i / L_DRAIN_KB, ib = i % L_DRAIN_KB, b = c & 1;
            if (ib == 0 && c >= 2) { mbar_wait(aempty0 + 8 * b, ((c >> 1) - 1) & 1); tc_fence_after(); }      // chunk c-2 drained
            mbar_wait(rawf0 + 8 * s, (i / L_RAW) & 1);
            mbar_wait(lof0 + 8 * sl, (i / L_LO) & 1);
            tc_fence_after();
            if (lane == 0 && i == 0) TL_TRACE(5);
            if (lane == 0 && i == nkb - 1) TL_TRACE(8);
            if (g.xsync && second && i == nkb - 1 && lane == 0) { __threadfence(); atomicAdd(g.xsync, 1u); }      // this CTA's reads of the shared operand are over
            if (elect_one()) {
                const uint32_t acc = tmem_base + (uint32_t)(b * L_BN);
                const uint32_t sa = smem_u32(smem + (size_t)s * L_SLOT_B), sb = sa + L_PLANE_B;
                const uint32_t la = smem_u32(lo_ring + (size_t)sl * L_SLOT_B), lb = la + L_PLANE_B;
                const uint64_t a_hi = tl_desc(sa, p.a_mn, g), a_lo = tl_desc(la, p.a_mn, g);
                const uint64_t b_hi = tl_desc(sb, p.b_mn, g), b_lo = tl_desc(lb, p.b_mn, g);
                #pragma unroll
                for (int k = 0; k < L_BK / L_UK; k++) {
                    tc_mma_tf32(acc, a_lo + k * ka, b_hi + k * kb, idesc, (ib | k) ? 1u : 0u);
                    tc_mma_tf32(acc, a_hi + k * ka, b_lo + k * kb, idesc, 1u);
                    tc_mma_tf32(acc, a_hi + k * ka, b_hi + k * kb, idesc, 1u);
                }
            }
            __syncwarp();
            if (elect_one()) {
                tc_commit(rawe0 + 8 * s);
                tc_commit(loe0 + 8 * sl);
                if (ib == L_DRAIN_KB - 1 || i == nkb - 1) tc_commit(afull0 + 8 * b);
            }
            __syncwarp();
        }
    } else if (warp < 2 + L_NCONV) {
        // ===== converters: lo = tf32(a - truncate(a)) on the swizzled image, same offset in the lo ring; generated A tiles =====
        const int t = threadIdx.x - 64;                                               // 0..191
        constexpr int NV = (2048 + L_CONV_T - 1) / L_CONV_T;                           // 16-byte words of a k-block's two raw planes per thread (11)
        for (int i = 0; i < nkb; i++) {
            const int s = i % L_RAW, sl = i % L_LO;
            uint8_t *rbase = smem + (size_t)s * L_SLOT_B, *lbase = lo_ring + (size_t)sl * L_SLOT_B;
            mbar_wait(loe0 + 8 * sl, ((i / L_LO) & 1) ^ 1);                           // the lo slot's previous MMAs are done
            if (gen) {
                mbar_wait(rawe0 + 8 * s, ((i / L_RAW) & 1) ^ 1);                      // ... and the raw slot's
                const int E2 = p.gE2;
                auto dy1 = [&](int dr, int gr, int e0, float (&x)[4]) {               // dY1[gr][e0 .. e0+3]; dr = row in sD.  Class order as k_head_bwd: the same bits
                    x[0] = x[1] = x[2] = x[3] = 0.0f;
                    if (gr < Ng && e0 < EHg) {                                        // EHg % 4 == 0 (host check): whole chunks
                        for (int j = 0; j < E2; j++) {
                            const float dj = sD[dr * E2 + j];
                            const float4 w = *reinterpret_cast<const float4*>(sW2 + j * EHg + e0);
                            x[0] = fmaf(dj, w.x, x[0]); x[1] = fmaf(dj, w.y, x[1]); x[2] = fmaf(dj, w.z, x[2]); x[3] = fmaf(dj, w.w, x[3]);
                        }
                        if (p.gF) {
                            const float4 f = ldg4(p.gF + (int64_t)gr * EHg + e0);
                            x[0] = __fmul_rn(x[0], f.x); x[1] = __fmul_rn(x[1], f.y); x[2] = __fmul_rn(x[2], f.z); x[3] = __fmul_rn(x[3], f.w);
                        }
                    }
                };
                auto put = [&](uint32_t off, const float (&x)[4]) {
                    float lo[4];
                    #pragma unroll
                    for (int e = 0; e < 4; e++) lo[e] = to_tf32(x[e] - __uint_as_float(__float_as_uint(x[e]) & 0xFFFFE000u));
                    *reinterpret_cast<float4*>(rbase + off) = make_float4(x[0], x[1], x[2], x[3]);
                    *reinterpret_cast<float4*>(lbase + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                };
                if (gen == 1) {
                    // K-major tile [128 rows = samples][32 k = hidden units]: thread -> (row, 16-byte chunk); 8 threads write one swizzled 128-byte row
                    const int c = t & 7, e0 = (kt0 + i) * L_BK + 4 * c;
                    #pragma unroll 1
                    for (int r = t >> 3; r < L_BM; r += L_CONV_T / 8) {
                        float x[4]; dy1(r, mt * L_BM + r, e0, x);
                        put((uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4), x);
                    }
                } else {
                    // M-major tile: 32 k-rows (samples) x 4 blocks of 32 m (hidden units); 16-byte chunk cc of a row sits in 32-byte chunk (cc >> 1) ^ (row & 3)
                    #pragma unroll 1
                    for (int idx = t; idx < 1024; idx += L_CONV_T) {
                        const int row = idx >> 5, c = idx & 31, j = c >> 3, cc = c & 7;
                        float x[4]; dy1(i * L_BK + row, (kt0 + i) * L_BK + row, 4 * c, x);
                        put((uint32_t)j * 4096u + (uint32_t)row * 128u + (uint32_t)((((cc >> 1) ^ (row & 3)) << 5) | ((cc & 1) << 4)), x);
                    }
                }
            }
            mbar_wait(rawf0 + 8 * s, (i / L_RAW) & 1);
            if (t == 0 && i == 0) TL_TRACE(3);
            if (t == 0 && i == nkb - 1) TL_TRACE(7);
            float4 v[NV];
            #pragma unroll
            for (int j = 0; j < NV; j++) {
                const int idx = t + L_CONV_T * j + (gen ? 1024 : 0);                  // generated A: only the B plane is converted
                if (idx < 2048) v[j] = *reinterpret_cast<const float4*>(rbase + (size_t)idx * 16);
            }
            #pragma unroll
            for (int j = 0; j < NV; j++) {
                const int idx = t + L_CONV_T * j + (gen ? 1024 : 0);
                if (idx >= 2048) break;
                const float x[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
                float hi[4], lo[4];
                #pragma unroll
                for (int e = 0; e < 4; e++) {
                    hi[e] = __uint_as_float(__float_as_uint(x[e]) & 0xFFFFE000u);
                    lo[e] = to_tf32(x[e] - hi[e]);
                }
                *reinterpret_cast<float4*>(lbase + (size_t)idx * 16) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                if (g.mask_hi) *reinterpret_cast<float4*>(rbase + (size_t)idx * 16) = make_float4(hi[0], hi[1], hi[2], hi[3]);
            }
            tl_fence_proxy_async();                      // generic-proxy stores → visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(lof0 + 8 * sl);
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
        // the raw ring is idle: the last accumulator commit covers every MMA that read it, every TMA write was consumed
        float *park = reinterpret_cast<float*>(smem);
        const int r = q * 32 + lane;
        #pragma unroll
        for (int j = 0; j < CW; j += 4)
            *reinterpret_cast<float4*>(park + tl_park_off(r, (h * CW + j) >> 2)) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
        if (warp == 2 + L_NCONV && lane == 0) TL_TRACE(10);
    }
    if (active && (p.mode == 2 || p.mode == 4)) cp_async_wait_all();
    tc_fence_before();
    __syncthreads();
    // split-K partials through L2: distributed shared memory serves ~20 B/clk per SM (60 KiB of peer tiles: 2 us) and keeps every CTA resident
    // until its last reader is done; the L2 path costs about the same for the reader and lets everybody leave early.  The tile goes out
    // coalesced (a warp per row), the cluster barrier (release/acquire) publishes it.
    float *mypart = nullptr;
    if (active && split > 1 && p.part) {
        mypart = p.part + (size_t)tile * split * (size_t)(L_BM * L_BN);
        const float *park = reinterpret_cast<const float*>(smem);
        float *dst = mypart + (size_t)zs * (L_BM * L_BN);
        const int rows = min(L_BM, p.M - mt * L_BM);
        if (nt * L_BN + lane * 4 < p.N)
            for (int r = warp; r < rows; r += L_WARPS)
                __stcg(reinterpret_cast<float4*>(dst + r * L_BN + lane * 4), *reinterpret_cast<const float4*>(park + tl_park_off(r, lane)));
    }
    if (S > 1) tl_cluster_sync();                                                     // every CTA's tile is parked / stored and visible cluster-wide
    if (threadIdx.x == 0) TL_TRACE(11);

    if (g.xsync && !second) {                                                         // in-place pair: the operand this problem overwrites has been read
        if (threadIdx.x == 0) {
            unsigned v;
            do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(g.xsync) : "memory"); } while (v < g.nread);
        }
        __syncthreads();
    }
    // ===== reduction over the tile's ranks + epilogue: rank `zs` finishes rows [zs*rpr, (zs+1)*rpr) of the tile, a warp per row, lane = 16-byte chunk =====
    if (active) {
        const int rpr = L_BM / split;
        const uint32_t park_s = smem_u32(smem);
        const float *park = reinterpret_cast<const float*>(smem);
        const int gc = nt * L_BN + lane * 4;
        const bool o_vec = ((p.N & 3) == 0);
        // mode 4: per-warp gradient partials of the head in the idle ring behind the parked tile, zeroed here
        const int E2p = (p.E2 + 3) & ~3, nEp = p.E2 * p.N + E2p + ((p.N + 3) & ~3);
        float *sAcc = reinterpret_cast<float*>(smem + (size_t)L_BM * L_BN * 4);
        if (p.mode == 4) { for (int t = lane; t < nEp; t += 32) sAcc[(size_t)warp * nEp + t] = 0.0f; __syncwarp(); }
        for (int rr = warp; rr < rpr; rr += L_WARPS) {
            const int r = zs * rpr + rr, gr = mt * L_BM + r;
            if (gr >= p.M) break;                                                     // rows ascend with rr
            float4 sum;
            if (split == 1) sum = *reinterpret_cast<const float4*>(park + tl_park_off(r, lane));
            else {
                float4 v[16];
                if (mypart) {
                    const float *src = mypart + r * L_BN + lane * 4;
                    const bool ld = gc < p.N;
                    #pragma unroll
                    for (int qk = 0; qk < 16; qk++) if (qk < split) v[qk] = ld ? __ldcg(reinterpret_cast<const float4*>(src + (size_t)qk * (L_BM * L_BN))) : make_float4(0.f, 0.f, 0.f, 0.f);
                } else {
                    const uint32_t a = park_s + (uint32_t)tl_park_off(r, lane) * 4u;
                    #pragma unroll
                    for (int qk = 0; qk < 16; qk++) if (qk < split) v[qk] = tl_ld_dsmem4(a, (uint32_t)(rk0 + qk));
                }
                sum = v[0];
                #pragma unroll
                for (int qk = 1; qk < 16; qk++) if (qk < split) { sum.x += v[qk].x; sum.y += v[qk].y; sum.z += v[qk].z; sum.w += v[qk].w; }
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
                    if (p.mode == 4) stg4(p.actF + at, make_float4(f[0], f[1], f[2], f[3]));      // N % 4 == 0 (host check); Y1 / A1 receive their backward values below
                    else if (full) {
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
                if (p.mode == 2 || p.mode == 4) {
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
                    const float pv = on ? ex / sm : 0.0f;
                    const int64_t o2 = (int64_t)gr * E2 + lane;
                    if (p.mode == 2) { if (on) { p.Y2[o2] = y2; p.P[o2] = pv; if (p.P2) p.P2[o2] = pv; } }
                    else {
                        // ---- head backward on the same row (k_head_bwd's arithmetic, class order ascending: the same bits)
                        const float d = on ? __fsub_rn(pv, __ldg(p.T + o2)) : 0.0f;
                        if (on) { p.P[o2] = d; p.Ylin[o2] = d; if (p.P2) p.P2[o2] = pv; }
                        float dx[4] = {0.0f, 0.0f, 0.0f, 0.0f};
                        float *acc_w = sAcc + (size_t)warp * nEp;                      // this warp's gradient partials [E2][EH] | [E2] | [EH]
                        #pragma unroll 1
                        for (int k = 0; k < E2; k++) {
                            const float dk = __shfl_sync(0xffffffffu, d, k);
                            if (gc < EH) {
                                const float4 w = *reinterpret_cast<const float4*>(sW2 + k * EH + gc);
                                dx[0] = fmaf(dk, w.x, dx[0]); dx[1] = fmaf(dk, w.y, dx[1]); dx[2] = fmaf(dk, w.z, dx[2]); dx[3] = fmaf(dk, w.w, dx[3]);
                                float4 *q = reinterpret_cast<float4*>(acc_w + k * EH + gc);      // dW2[k][e] += d_k * a1[e]
                                float4 t4 = *q;
                                t4.x = fmaf(dk, a[0], t4.x); t4.y = fmaf(dk, a[1], t4.y); t4.z = fmaf(dk, a[2], t4.z); t4.w = fmaf(dk, a[3], t4.w);
                                *q = t4;
                            }
                        }
                        if (on) acc_w[E2 * EH + lane] += d;                            // dB2
                        if (gc < EH) {
                            const float dy[4] = {__fmul_rn(dx[0], f[0]), __fmul_rn(dx[1], f[1]), __fmul_rn(dx[2], f[2]), __fmul_rn(dx[3], f[3])};
                            stg4(p.actA + at, make_float4(dx[0], dx[1], dx[2], dx[3]));      // the activation's output tensor <- dX of the head linear
                            stg4(p.O + at, make_float4(dy[0], dy[1], dy[2], dy[3]));         // the hidden linear's output tensor <- dY1
                            float4 *q = reinterpret_cast<float4*>(acc_w + E2 * EH + E2p + gc);   // dB1
                            float4 t4 = *q; t4.x += dy[0]; t4.y += dy[1]; t4.z += dy[2]; t4.w += dy[3]; *q = t4;
                        }
                    }
                }
            }
        }
    }
    if (active && p.mode == 4) {
        // the CTA's partial of (dW2 | dB2 | dB1): warps summed in order; k_head_grad_fin adds the CTAs' partials in CTA order (deterministic)
        __syncthreads();
        const int E2p = (p.E2 + 3) & ~3, nEp = p.E2 * p.N + E2p + ((p.N + 3) & ~3);
        const float *sAcc = reinterpret_cast<const float*>(smem + (size_t)L_BM * L_BN * 4);
        float *dst = p.hpart + (size_t)blockIdx.x * nEp;
        for (int t = threadIdx.x; t < nEp; t += L_THREADS) {
            float s_ = 0.0f;
            #pragma unroll
            for (int w = 0; w < L_WARPS; w++) s_ += sAcc[(size_t)w * nEp + t];
            dst[t] = s_;
        }
    }
    if (threadIdx.x == 0) TL_TRACE(13);
    // distributed shared memory: nobody leaves while its tile is still being read (uniform over the grid: it depends on the problems only)
    const bool dsm_exit = (g.p[0].split > 1 && !g.p[0].part) || (g.p[1].split > 1 && !g.p[1].part);
    if (S > 1 && dsm_exit) tl_cluster_sync();
    if (threadIdx.x == 0) TL_TRACE(14);
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 2 * L_BN); }
    if (g.xsync && threadIdx.x == 0) {
        const unsigned t = atomicAdd(g.xsync + 1, 1u);
        if (t == gridDim.x - 1) { g.xsync[0] = 0u; __threadfence(); g.xsync[1] = 0u; }           // everybody is past its use of the counters
    }
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

constexpr size_t L_SMEM = (size_t)L_RING_B + 256 + (size_t)(L_W2_FLTS + L_D_FLTS) * 4 + 1024;
static_assert(L_SMEM <= 227 * 1024, "shared memory budget");
static_assert(L_NBAR * 8 + 8 <= 256, "barrier block");
#define TL_MAX_DEV 16
static int g_tl_maxcl[TL_MAX_DEV][5];                   // [device][log2 S]: co-resident clusters of size S (-1: unavailable)
static unsigned *g_tl_xsync[TL_MAX_DEV];                // in-place pair counters (two words per ring slot, zero between launches)
static unsigned g_tl_xslot[TL_MAX_DEV];

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
        cfg.gridDim = dim3(S, 1, 1); cfg.blockDim = dim3(L_THREADS); cfg.dynamicSmemBytes = L_SMEM;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = (unsigned)S; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int n = 0;
        if (S == 1) n = sm_count();
        else if (cudaOccupancyMaxActiveClusters(&n, k_gemm_tl, &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
        g_tl_maxcl[dev][l] = n > 0 ? n : -1;
    }
    if (!g_tl_xsync[dev]) {                               // ring of 64 counter pairs: launches of different streams never share one
        if (cudaMalloc((void**)&g_tl_xsync[dev], 64 * 2 * sizeof(unsigned)) != cudaSuccess) { cudaGetLastError(); return T4K_ENOMEM; }
        cudaMemset(g_tl_xsync[dev], 0, 64 * 2 * sizeof(unsigned));
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

// one problem of a launch → TlP (everything but the split) + its two tensor maps
static int tl_fill(const TlJob &j, TlP &p, CUtensorMap *amap, CUtensorMap *bmap, CUtensorMapSwizzle mn_swz) {
    p = TlP{};
    p.O = j.O; p.alpha = j.alpha; p.beta = j.beta; p.M = j.M; p.N = j.N; p.K = j.K;
    p.mtiles = (j.M + L_BM - 1) / L_BM; p.ntiles = (j.N + L_BN - 1) / L_BN; p.KT = (j.K + L_BK - 1) / L_BK;
    p.a_mn = j.tA ? 1 : 0; p.b_mn = j.tB ? 0 : 1; p.mode = 0;
    if (const TlEpi *epi = j.epi) {
        p.mode = epi->mode; p.bias = epi->bias; p.actA = epi->actA; p.actF = epi->actF; p.layer = epi->layer; p.act_alpha = epi->act_alpha;
        p.W2 = epi->W2; p.B2 = epi->B2; p.Y2 = epi->Y2; p.P = epi->P; p.P2 = epi->P2; p.E2 = epi->E2; p.F = epi->F; p.O2 = epi->O2;
        if ((epi->mode == 2 || epi->mode == 4) && (p.ntiles != 1 || (j.N & 3) || epi->E2 > 32 || epi->E2 < 1)) return T4K_ENOSUP;
        if (epi->mode == 4) {
            const int E2p = (epi->E2 + 3) & ~3, nEp = epi->E2 * j.N + E2p + ((j.N + 3) & ~3);
            if ((size_t)L_WARPS * nEp * 4 > (size_t)L_RING_B - (size_t)L_BM * L_BN * 4 || !epi->T || !epi->Ylin || !epi->hpart) return T4K_ENOSUP;
            p.T = epi->T; p.Ylin = epi->Ylin; p.hpart = epi->hpart;
        }
        if (epi->gP) {
            const int EH = j.tA ? j.M : j.K;                                   // hidden width: K of the K-major operand, M of the M-major one
            if (EH > 128 || (EH & 3) || epi->gE2 < 1 || epi->gE2 > 32 || !epi->gT || !epi->gW2) return T4K_ENOSUP;
            p.gen = j.tA ? 2 : 1; p.gE2 = epi->gE2; p.gP = epi->gP; p.gT = epi->gT; p.gW2 = epi->gW2; p.gF = epi->gF;
        }
    }
    int rc;
    // op(A)(m,k): A stored [M][K] (K-major: box 128 rows x 32 k) or [K][M] when tA (M-major: box 32 k-rows x 32 m)
    if (p.gen) rc = tl_map(amap, j.B, j.tB ? j.K : j.N, j.tB ? j.N : j.K, j.tB ? 128 : 32, j.tB ? CU_TENSOR_MAP_SWIZZLE_128B : mn_swz);   // unused: any valid map
    else rc = j.tA ? tl_map(amap, j.A, j.M, j.K, 32, mn_swz) : tl_map(amap, j.A, j.K, j.M, 128);
    if (rc) return rc;
    // op(B)(k,n): B stored [N][K] when tB (K-major) or [K][N] (N-major)
    return j.tB ? tl_map(bmap, j.B, j.K, j.N, 128) : tl_map(bmap, j.B, j.N, j.K, 32, mn_swz);
}

// One launch for one or two problems.  Cluster size S and, per problem, K-split (split == S: a cluster is one tile) or none (split == 1: a
// cluster is S tiles): the combination with the fewest k-blocks per CTA that keeps the whole grid co-resident (one wave).
static int tl_launch(const TlJob *jobs, int njobs, cudaStream_t st, int *ctas_out, bool dry, bool inplace = false) {
    if (njobs < 1 || njobs > 2) return T4K_EINVAL;
    const int dev = tl_device();
    if (dev < 0) return T4K_EINVAL;
    int rc = tl_prepare(dev); if (rc) return rc;
    TlG g{};
    CUtensorMap maps[4];
    const CUtensorMapSwizzle mn_swz = (CUtensorMapSwizzle)tl_knob(4);
    for (int q = 0; q < njobs; q++) { rc = tl_fill(jobs[q], g.p[q], &maps[2 * q], &maps[2 * q + 1], mn_swz); if (rc) return rc; }
    if (njobs == 1) { maps[2] = maps[0]; maps[3] = maps[1]; }
    const int smax = tl_knob(2);
    int bestS = 1, bestSplit[2] = {1, 1}, bestCost = 1 << 30, bestCtas = 0;
    for (int l = 0; l <= 4; l++) {
        const int S = 1 << l, cap = g_tl_maxcl[dev][l];
        if (S > smax || cap <= 0) continue;
        for (int c0 = 0; c0 < 2; c0++) for (int c1 = 0; c1 < (njobs == 2 ? 2 : 1); c1++) {
            const int sp[2] = {c0 ? S : 1, c1 ? S : 1};
            if (S == 1 && (c0 || c1)) continue;
            int ncl = 0, cost = 0; bool okc = true;
            for (int q = 0; q < njobs; q++) {
                const TlP &p = g.p[q];
                const int T = p.mtiles * p.ntiles, per = (p.KT + sp[q] - 1) / sp[q];
                if ((p.KT + per - 1) / per != sp[q]) okc = false;             // an empty rank
                if (p.gen == 2 && per > 4) okc = false;                        // generated M-major A: the k range's p - y rows must fit sD (128 rows)
                ncl += (T * sp[q] + S - 1) / S;
                // k-blocks per CTA, plus what a split costs (store + barrier + reload of the partials: about three k-blocks' worth)
                cost = max(cost, per + (sp[q] > 1 ? 3 : 0));
            }
            if (!okc || ncl > cap) continue;
            if (cost < bestCost || (cost == bestCost && ncl * S < bestCtas)) { bestCost = cost; bestS = S; bestSplit[0] = sp[0]; bestSplit[1] = sp[1]; bestCtas = ncl * S; }
        }
    }
    if (bestCost == (1 << 30)) return T4K_ENOSUP;                              // does not fit one wave in any shape
    const int S = bestS;
    int ncl[2] = {0, 0};
    size_t part_flts = 0;
    for (int q = 0; q < njobs; q++) {
        TlP &p = g.p[q];
        p.split = bestSplit[q]; p.kt_per = (p.KT + p.split - 1) / p.split;
        const int T = p.mtiles * p.ntiles;
        ncl[q] = (T * p.split + S - 1) / S;
        if (p.split > 1 && tl_knob(8)) part_flts += (size_t)T * p.split * L_BM * L_BN;
    }
    if (ctas_out) *ctas_out = (ncl[0] + ncl[1]) * S;
    if (dry) return 0;
    if (part_flts) {
        float *part = (float*)workspace(part_flts * sizeof(float), 7);
        if (!part) return T4K_ENOMEM;
        for (int q = 0; q < njobs; q++) {
            TlP &p = g.p[q];
            if (p.split > 1) { p.part = part; part += (size_t)p.mtiles * p.ntiles * p.split * L_BM * L_BN; }
        }
    }
    if (njobs == 1) g.p[1] = g.p[0];
    g.ncl0 = ncl[0];
    g.mask_hi = tl_knob(1);
    g.mn_type = (uint32_t)tl_knob(3); g.mn_lbo = (uint32_t)tl_knob(5); g.mn_sbo = (uint32_t)tl_knob(6); g.mn_kstep = (uint32_t)tl_knob(7);
    g.trace = g_tl_trace;
    if (inplace && njobs == 2) { g.xsync = g_tl_xsync[dev] + 2 * (g_tl_xslot[dev]++ & 63u); g.nread = (unsigned)(ncl[1] * S); }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)((ncl[0] + ncl[1]) * S)); cfg.blockDim = dim3(L_THREADS); cfg.dynamicSmemBytes = L_SMEM; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = (unsigned)S; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = S > 1 ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, k_gemm_tl, g, maps[0], maps[1], maps[2], maps[3]);
    ++g_launches;
    if (e != cudaSuccess) { cudaGetLastError(); return (int)e; }
    return (int)cudaGetLastError();
}

int gemm_tl_multi(const TlJob *jobs, int njobs, cudaStream_t st, int *ctas_out) { return tl_launch(jobs, njobs, st, ctas_out, false); }
// two problems, the FIRST of which stores over an operand the SECOND reads (see TlG::xsync)
int gemm_tl_pair_inplace(const TlJob *jobs, cudaStream_t st) { return tl_launch(jobs, 2, st, nullptr, false, true); }
int gemm_tl_ctas(const TlJob *jobs, int njobs) { int n = 0; const int rc = tl_launch(jobs, njobs, nullptr, &n, true); return rc ? rc : n; }
int gemm_tl(const float *A, const float *B, float *O, float alpha, float beta, int tA, int tB, int M, int N, int K, cudaStream_t st, const TlEpi *epi) {
    TlJob j{A, B, O, alpha, beta, tA, tB, M, N, K, epi};
    return gemm_tl_multi(&j, 1, st);
}

} // namespace t4k

extern "C" int t4k_gemm_tl_trace(long long *dev16) { t4k::g_tl_trace = dev16; return 0; }   // bring-up: 16 clock64 stamps of CTA 0, nullptr = off
extern "C" int t4k_set_gemm_tl(int what, int value) {
    if (what < 0 || what >= TL_NKNOB) return T4K_EINVAL;
    const int was = t4k::tl_knob(what);
    t4k::g_tl_knob[what] = value;
    return was;
}
