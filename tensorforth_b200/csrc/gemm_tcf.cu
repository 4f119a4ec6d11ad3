// gemm_tcf.cu — single-launch FP32 GEMM on the 5th-gen tensor cores for the mid-size problems of the NN layers
// (linear forward / dW / dX at batch 512-1024: 0.1-1 GFLOP), "3xTF32" as gemm_tc.cu, but with the operand split FUSED
// into the kernel: there is no pack pass and no packed copy of A/B in HBM.
//   replaces k_gemm_tile_claude (src/t4math.cu:478-583) as launched by Tensor::linear / Model::_flinear / _blinear
//   (src/mu/tensor.cu:74-87, src/nn/forward.cu:158-198, src/nn/backprop.cu:194-254).
//
// One 128 x 128 output tile per CTA, split-K over gridDim.z so that at most one CTA lands on every SM.  21 warps:
//   warps 0-11  producers, three groups of four that take k-blocks round robin: coalesced 128-bit global loads of the raw FP32
//               operands (any tA/tB: K-contiguous rows are read as 8 x float4 per row, M/N-contiguous operands as float4s
//               of 4 rows at one k), hi/lo TF32 split in registers, stores into shared memory directly in the UMMA
//               SWIZZLE_128B K-major layout (16-byte chunk index XOR (row & 7)), fence.proxy.async, mbarrier arrive
//   warp  12    tcgen05.mma issuer (kind::tf32, operands from shared memory, FP32 accumulators in TMEM; lo·hi, hi·lo, hi·hi)
//   warps 13-20 epilogue: drain the accumulator chain every DRAIN_KB k-blocks into registers (round-to-nearest adds, see
//               gemm_tc.cu), alpha/beta, 128-bit stores — or split-K partials for the caller's fused finish
// 3-stage ring of 64 KiB stages (A hi+lo 32 KiB, B hi+lo 32 KiB).  Per k-block a CTA moves 32 KiB of raw operands and
// issues 12 MMAs of 128x128x8: producers and tensor pipe are balanced at ~0.5 us per k-block.
// Bound: tensor pipe for large K, launch/ramp latency for the layer shapes (a few k-blocks per CTA).
#include "tc_ptx.cuh"
#include "act.cuh"
#include <cstdlib>

namespace t4k {

constexpr int F_BM = 128, F_BN = 128, F_BK = 32, F_UK = 8;
constexpr int F_PLANE_FLTS = 128 * F_BK;                 // one hi (or lo) plane of a 128-row operand tile
constexpr uint32_t F_PLANE_B = F_PLANE_FLTS * 4;         // 16 KiB
constexpr uint32_t F_OP_B = 2 * F_PLANE_B;               // hi + lo
constexpr uint32_t F_STAGE_B = 2 * F_OP_B;               // A + B = 64 KiB
constexpr int F_STAGES = 3;
constexpr int F_DRAIN_KB = 8;                            // k-blocks per accumulator chain (gemm_tc.cu: DRAIN_KB)
constexpr int F_NGROUP = 3;                             // producer groups (4 warps each), one per ring stage: 96 KiB of loads in flight per CTA
constexpr int F_NPROD = 4 * F_NGROUP, F_NEPI = 8;
constexpr int F_THREADS = (F_NPROD + 1 + F_NEPI) * 32;   // 672

struct TcfP {
    const float *A, *B;          // raw operands
    float *O;                    // [M,N] row-major
    float alpha, beta;
    int M, N, K;
    int64_t a_sr, a_sk;          // op(A)(m,k) = A[m*a_sr + k*a_sk]
    int64_t b_sr, b_sk;          // op(B)(k,n) = B[n*b_sr + k*b_sk]   (row index of the packed tile = n)
    int KT, kt_per_split, splits;
    float *part;                 // [splits][M*N] when splits > 1
};
// CLUSTER variant only: fused bias (+ activation) epilogue of a linear layer, bias == nullptr: none.  Kept out of TcfP so that the
// parameter block — and with it the generated code — of the default kernel stays exactly what was validated on the device.
struct TcfEpi { const float *bias; float *actA, *actF; int layer; float act_alpha; };
template<bool CLUSTER> struct TcfArgs { TcfP p; };
template<> struct TcfArgs<true> { TcfP p; TcfEpi e; };

__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// one 128 x 32 operand tile: rows r0.. (bound R), k from k0 (bound K)  →  hi plane at `dst`, lo plane at dst + PLANE
// executed by the 128 threads of one producer group (t = 0..127)
__device__ __forceinline__ void produce_tile(const float *__restrict__ X, int64_t sr, int64_t sk, int R, int K, int r0, int k0,
                                             float *dst, int t, bool vec) {
    if (sk == 1) {
        // K-contiguous rows: 8 threads x float4 per row, 16 rows per pass, 8 passes
        float4 v[8];
        #pragma unroll
        for (int ps = 0; ps < 8; ps++) {
            const int r = ps * 16 + (t >> 3), c = t & 7;
            const int gr = r0 + r, gk = k0 + c * 4;
            v[ps] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gr < R && gk < K) {
                const float *src = X + (int64_t)gr * sr + gk;
                if (vec && gk + 3 < K) v[ps] = ldg4(src);
                else { v[ps].x = src[0]; if (gk + 1 < K) v[ps].y = src[1]; if (gk + 2 < K) v[ps].z = src[2]; if (gk + 3 < K) v[ps].w = src[3]; }
            }
        }
        #pragma unroll
        for (int ps = 0; ps < 8; ps++) {
            const int r = ps * 16 + (t >> 3), c = t & 7;
            float4 hi, lo;
            hi.x = to_tf32(v[ps].x); hi.y = to_tf32(v[ps].y); hi.z = to_tf32(v[ps].z); hi.w = to_tf32(v[ps].w);
            lo.x = to_tf32(v[ps].x - hi.x); lo.y = to_tf32(v[ps].y - hi.y); lo.z = to_tf32(v[ps].z - hi.z); lo.w = to_tf32(v[ps].w - hi.w);
            const int o = r * 32 + ((c ^ (r & 7)) << 2);
            *reinterpret_cast<float4*>(dst + o) = hi;
            *reinterpret_cast<float4*>(dst + F_PLANE_FLTS + o) = lo;
        }
    } else {
        // row-contiguous operand (sr == 1): float4 = rows 4*rq..4*rq+3 at one k.  A warp covers 8 k x 4 row-quads per pass
        // (64 contiguous bytes per k in global; 2-way bank conflicts at most on the scattered 4-byte shared stores)
        float4 v[8];
        const int w = t >> 5, k_lo = t & 7, rq_lo = (t >> 3) & 3;
        #pragma unroll
        for (int ps = 0; ps < 8; ps++) {
            const int combo = w + 4 * ps, k = (combo & 3) * 8 + k_lo, rq = (combo >> 2) * 4 + rq_lo;
            const int gr = r0 + rq * 4, gk = k0 + k;
            v[ps] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gk < K && gr < R) {
                const float *src = X + (int64_t)gk * sk + (int64_t)gr * sr;
                if (vec && sr == 1 && gr + 3 < R) v[ps] = ldg4(src);
                else { v[ps].x = src[0]; if (gr + 1 < R) v[ps].y = src[sr]; if (gr + 2 < R) v[ps].z = src[2 * sr]; if (gr + 3 < R) v[ps].w = src[3 * sr]; }
            }
        }
        #pragma unroll
        for (int ps = 0; ps < 8; ps++) {
            const int combo = w + 4 * ps, k = (combo & 3) * 8 + k_lo, rq = (combo >> 2) * 4 + rq_lo;
            const float x[4] = {v[ps].x, v[ps].y, v[ps].z, v[ps].w};
            #pragma unroll
            for (int j = 0; j < 4; j++) {
                const int r = rq * 4 + j;
                const float hi = to_tf32(x[j]), lo = to_tf32(x[j] - hi);
                const int o = r * 32 + ((((k >> 2) ^ (r & 7)) << 2) | (k & 3));
                dst[o] = hi; dst[F_PLANE_FLTS + o] = lo;
            }
        }
    }
}

// ---- thread-block cluster helpers (CLUSTER variant: split-K reduced through distributed shared memory)
__device__ __forceinline__ void tcf_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float4 tcf_ld_dsmem4(const float *local, uint32_t rank) {      // the same shared-memory offset in CTA `rank` of the cluster
    uint32_t a = (uint32_t)__cvta_generic_to_shared(local), r; float4 v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
    asm volatile("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(r) : "memory");
    return v;
}
// staging tile [128 rows][128 cols] fp32 over the (idle) operand ring: 16-byte chunk c4 of row r lives at chunk (c4 ^ (r & 31)):
// the epilogue's per-row float4 stores (lanes = rows) and the reduction's per-row reads (lanes = chunks) are both conflict-free
__device__ __forceinline__ int tcf_stage_off(int row, int c4) { return row * F_BN + ((c4 ^ (row & 31)) << 2); }

// CLUSTER = false: split-K partials go to global memory, a second launch (or the caller's fused finish) adds them.
// CLUSTER = true (EXPERIMENTAL, opt-in with T4K_TCF_CLUSTER=1; correct on the device — tests/test_gpu_kernels.py -k tcf_layer_shapes with the
// variable set: 24/24 vs float64 — but not yet TIMED, so the automatic path does not use it; the fused bias/activation epilogue below
// (t4k_linear_act_fwd passes it) was added after that run and has not been on a device yet): the `splits` CTAs of one output tile
// form a thread-block cluster (1 x 1 x splits, splits a power of two <= 8); every CTA parks its accumulator tile in its own shared
// memory, and after a cluster barrier CTA r adds rows [r*128/splits, (r+1)*128/splits) of all the parked tiles in RANK ORDER through
// distributed shared memory and writes alpha*sum + beta*O: no partials in HBM, no finish launch, deterministic.
template<bool CLUSTER>
__global__ void __launch_bounds__(F_THREADS, 1) k_gemm_tcf(const TcfArgs<CLUSTER> args) {
    const TcfP &p = args.p;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);      // SWIZZLE_128B tiles: 1024-byte aligned
    uint64_t *bars = (uint64_t*)(smem + F_STAGES * F_STAGE_B);                        // full[S], empty[S], acc_full[2], acc_empty[2]
    uint32_t *tmem_slot = (uint32_t*)(bars + 2 * F_STAGES + 4);
    pdl_wait(); pdl_trigger();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mt = blockIdx.y, nt = blockIdx.x, zs = blockIdx.z;
    const int kt0 = zs * p.kt_per_split;
    const int kt1 = min(p.KT, kt0 + p.kt_per_split);
    const int nkb = kt1 - kt0;
    const int nchunk = (nkb + F_DRAIN_KB - 1) / F_DRAIN_KB;
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + F_STAGES);
    const uint32_t afull0 = smem_u32(bars + 2 * F_STAGES), aempty0 = smem_u32(bars + 2 * F_STAGES + 2);

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < F_STAGES; s++) { mbar_init(full0 + 8 * s, 4); mbar_init(empty0 + 8 * s, 1); }     // 4 producer warps fill a stage
        for (int b = 0; b < 2; b++) { mbar_init(afull0 + 8 * b, 1); mbar_init(aempty0 + 8 * b, F_NEPI); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == F_NPROD) tmem_alloc(smem_u32(tmem_slot), 2 * F_BN);                   // two accumulators of 128 fp32 columns
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < F_NPROD) {
        // ===== producers: group g = warp / 4 fills the k-blocks i with i % F_NGROUP == g =====
        const int g = warp >> 2, t = threadIdx.x & 127;
        const bool avec = ((p.a_sr & 3) == 0 || p.a_sr == 1) && ((p.a_sk & 3) == 0 || p.a_sk == 1) && ((((uintptr_t)p.A) & 15) == 0);
        const bool bvec = ((p.b_sr & 3) == 0 || p.b_sr == 1) && ((p.b_sk & 3) == 0 || p.b_sk == 1) && ((((uintptr_t)p.B) & 15) == 0);
        for (int i = g; i < nkb; i += F_NGROUP) {
            const int s = i % F_STAGES, it = i / F_STAGES;
            mbar_wait(empty0 + 8 * s, (it & 1) ^ 1);
            float *sa = reinterpret_cast<float*>(smem + (size_t)s * F_STAGE_B);
            float *sb = sa + 2 * F_PLANE_FLTS;
            const int k0 = (kt0 + i) * F_BK;
            produce_tile(p.A, p.a_sr, p.a_sk, p.M, p.K, mt * F_BM, k0, sa, t, avec);
            produce_tile(p.B, p.b_sr, p.b_sk, p.N, p.K, nt * F_BN, k0, sb, t, bvec);
            fence_proxy_async_smem();                    // generic-proxy stores → visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(full0 + 8 * s);
        }
    } else if (warp == F_NPROD) {
        // ===== MMA issuer =====
        constexpr uint32_t idesc = idesc_tf32(F_BM, F_BN);
        for (int i = 0; i < nkb; i++) {
            const int s = i % F_STAGES, it = i / F_STAGES;
            const int c = i / F_DRAIN_KB, ib = i % F_DRAIN_KB, b = c & 1;
            if (ib == 0 && c >= 2) { mbar_wait(aempty0 + 8 * b, ((c >> 1) - 1) & 1); tc_fence_after(); }      // chunk c-2 drained
            mbar_wait(full0 + 8 * s, it & 1);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t acc = tmem_base + (uint32_t)(b * F_BN);
                const uint32_t sa = smem_u32(smem + (size_t)s * F_STAGE_B);
                const uint32_t sb = sa + F_OP_B;
                const uint64_t a_hi = smem_desc_sw128(sa), a_lo = smem_desc_sw128(sa + F_PLANE_B);
                const uint64_t b_hi = smem_desc_sw128(sb), b_lo = smem_desc_sw128(sb + F_PLANE_B);
                #pragma unroll
                for (int k = 0; k < F_BK / F_UK; k++) {
                    const uint64_t ko = (uint64_t)((k * F_UK * 4) >> 4);
                    tc_mma_tf32(acc, a_lo + ko, b_hi + ko, idesc, (ib | k) ? 1u : 0u);
                    tc_mma_tf32(acc, a_hi + ko, b_lo + ko, idesc, 1u);
                    tc_mma_tf32(acc, a_hi + ko, b_hi + ko, idesc, 1u);
                }
            }
            __syncwarp();
            if (elect_one()) {
                tc_commit(empty0 + 8 * s);
                if (ib == F_DRAIN_KB - 1 || i == nkb - 1) tc_commit(afull0 + 8 * b);
            }
            __syncwarp();
        }
    } else {
        // ===== epilogue: 8 warps after the issuer; TMEM lane quarter = warp % 4 (hardware rule), column half = (warp - first) / 4 =====
        constexpr int CW = F_BN / 2;
        const int q = warp & 3, h = (warp - (F_NPROD + 1)) >> 2;
        const int row = mt * F_BM + q * 32 + lane;
        float acc[CW];
        #pragma unroll
        for (int j = 0; j < CW; j++) acc[j] = 0.0f;
        for (int c = 0; c < nchunk; c++) {
            const int b = c & 1;
            mbar_wait(afull0 + 8 * b, (c >> 1) & 1);
            tc_fence_after();
            #pragma unroll
            for (int gq = 0; gq < CW / 16; gq++) {
                uint32_t v[16];
                tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * F_BN + h * CW + gq * 16), v);
                tmem_ld_wait();
                #pragma unroll
                for (int j = 0; j < 16; j++) acc[gq * 16 + j] += __uint_as_float(v[j]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(aempty0 + 8 * b);
        }
        if (CLUSTER) {
            // park the tile: the operand ring is idle (the last accumulator commit covers every MMA that read it)
            float *stage = reinterpret_cast<float*>(smem);
            const int r = q * 32 + lane;
            #pragma unroll
            for (int j = 0; j < CW; j += 4)
                *reinterpret_cast<float4*>(stage + tcf_stage_off(r, (h * CW + j) >> 2)) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
        }
        float *dst; float alpha = p.alpha, beta = p.beta;
        if (p.splits > 1) { dst = p.part + (int64_t)zs * p.M * p.N; alpha = 1.0f; beta = 0.0f; }
        else dst = p.O;
        const bool n_vec = ((p.N & 3) == 0) && ((((uintptr_t)dst) & 15) == 0);
        const int col0 = nt * F_BN + h * CW;
        if (!CLUSTER && row < p.M && col0 < p.N) {
            float *o = dst + (int64_t)row * p.N + col0;
            #pragma unroll
            for (int j = 0; j < CW; j += 4) {
                if (n_vec && col0 + j + 3 < p.N) {
                    float4 r = make_float4(acc[j] * alpha, acc[j + 1] * alpha, acc[j + 2] * alpha, acc[j + 3] * alpha);
                    if (beta != 0.0f) {
                        const float4 old = *reinterpret_cast<const float4*>(o + j);
                        r.x += old.x * beta; r.y += old.y * beta; r.z += old.z * beta; r.w += old.w * beta;
                    }
                    stg4(o + j, r);
                } else {
                    #pragma unroll
                    for (int e = 0; e < 4; e++) {
                        if (col0 + j + e < p.N) {
                            float r = acc[j + e] * alpha;
                            if (beta != 0.0f) r += o[j + e] * beta;
                            o[j + e] = r;
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CLUSTER) {
        const int S = (int)gridDim.z;                                  // cluster = (1, 1, S): rank in the cluster = blockIdx.z
        tcf_cluster_sync();                                            // every CTA's tile is parked and visible cluster-wide
        const float *stage = reinterpret_cast<const float*>(smem);
        const int rpr = F_BM / S;                                      // rows of the tile this CTA finishes (S divides 128)
        const bool o_vec = ((p.N & 3) == 0) && ((((uintptr_t)p.O) & 15) == 0);
        const TcfEpi epi = [&] { if constexpr (CLUSTER) return args.e; else return TcfEpi{}; }();
        for (int idx = threadIdx.x; idx < rpr * (F_BN / 4); idx += F_THREADS) {
            const int r = zs * rpr + (idx >> 5), c4 = idx & 31;
            float4 sum = tcf_ld_dsmem4(stage + tcf_stage_off(r, c4), 0u);
            for (int qk = 1; qk < S; qk++) {
                const float4 v = tcf_ld_dsmem4(stage + tcf_stage_off(r, c4), (uint32_t)qk);
                sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
            }
            const int gr = mt * F_BM + r, gc = nt * F_BN + c4 * 4;
            if (gr >= p.M || gc >= p.N) continue;
            float *o = p.O + (int64_t)gr * p.N + gc;
            float out[4] = {sum.x * p.alpha, sum.y * p.alpha, sum.z * p.alpha, sum.w * p.alpha};
            if (epi.bias) {
                // linear layer epilogue (k_linear_fin's arithmetic: Σ splits, + bias, activation): beta == 0 here
                #pragma unroll
                for (int e = 0; e < 4; e++) {
                    if (gc + e >= p.N) break;
                    const float y = out[e] + __ldg(epi.bias + gc + e);
                    o[e] = y;
                    if (epi.layer != T4K_L_NONE) {
                        const int64_t at = (int64_t)gr * p.N + gc + e;
                        float a = 0.0f, f = (epi.layer == T4K_L_DROPOUT) ? epi.actF[at] : 0.0f;
                        switch (epi.layer) {
                        case T4K_L_RELU:    act<T4K_L_RELU>(y, epi.act_alpha, a, f); break;
                        case T4K_L_TANH:    act<T4K_L_TANH>(y, epi.act_alpha, a, f); break;
                        case T4K_L_SIGMOID: act<T4K_L_SIGMOID>(y, epi.act_alpha, a, f); break;
                        case T4K_L_SELU:    act<T4K_L_SELU>(y, epi.act_alpha, a, f); break;
                        case T4K_L_LEAKYRL: act<T4K_L_LEAKYRL>(y, epi.act_alpha, a, f); break;
                        case T4K_L_ELU:     act<T4K_L_ELU>(y, epi.act_alpha, a, f); break;
                        default:            act<T4K_L_DROPOUT>(y, epi.act_alpha, a, f); break;
                        }
                        epi.actA[at] = a; epi.actF[at] = f;
                    }
                }
                continue;
            }
            if (o_vec && gc + 3 < p.N) {
                float4 w = make_float4(out[0], out[1], out[2], out[3]);
                if (p.beta != 0.0f) { const float4 old = *reinterpret_cast<const float4*>(o); w.x += old.x * p.beta; w.y += old.y * p.beta; w.z += old.z * p.beta; w.w += old.w * p.beta; }
                stg4(o, w);
            } else {
                #pragma unroll
                for (int e = 0; e < 4; e++) if (gc + e < p.N) o[e] = (p.beta != 0.0f) ? out[e] + o[e] * p.beta : out[e];
            }
        }
        tcf_cluster_sync();                                            // nobody leaves while its tile is still being read
    }
    if (warp == F_NPROD) { tc_fence_after(); tmem_dealloc(tmem_base, 2 * F_BN); }
}

// deterministic split-K reduction (fixed order), alpha/beta applied once
__global__ void __launch_bounds__(T4K_THREADS) k_splitk_fin_f(const float *__restrict__ part, float *O, float alpha, float beta, int64_t MN, int splits, int vec) {
    pdl_wait(); pdl_trigger();
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    if (vec) {
        for (int64_t q = tid; q < (MN >> 2); q += nth) {
            float4 s = ldg4(part + 4 * q);
            for (int k = 1; k < splits; k++) { const float4 t = ldg4(part + (int64_t)k * MN + 4 * q); s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w; }
            s.x *= alpha; s.y *= alpha; s.z *= alpha; s.w *= alpha;
            if (beta != 0.0f) { const float4 o = *reinterpret_cast<const float4*>(O + 4 * q); s.x += o.x * beta; s.y += o.y * beta; s.z += o.z * beta; s.w += o.w * beta; }
            stg4(O + 4 * q, s);
        }
    } else {
        for (int64_t e = tid; e < MN; e += nth) {
            float s = 0.0f;
            for (int k = 0; k < splits; k++) s += part[(int64_t)k * MN + e];
            O[e] = (beta == 0.0f) ? s * alpha : s * alpha + O[e] * beta;
        }
    }
}

static bool tcf_cluster_on() {
    static int on = -1;
    if (on < 0) { const char *e = getenv("T4K_TCF_CLUSTER"); on = (e && e[0] == '1') ? 1 : 0; }
    return on == 1;
}
bool gemm_tcf_ok(int tA, int tB, int M, int N, int K, int C, int batch) {
    // Measured on B200 (bench_scripts/gemm_probe.py, profiles/r01_gemm_probe.txt): launch + TMEM/barrier set-up + epilogue + split-K
    // finish cost ~10 us, so the FP32-FMA kernel wins below ~0.25 GFLOP (MNIST 1960->100 = 0.2 GFLOP stays there).  Above it this
    // kernel wins when BOTH operands are K-contiguous (X @ W^T, the linear forward: 0.27 GFLOP 13 vs 16 us, 0.82 GFLOP 20 vs 33 us);
    // an operand that is contiguous along M/N goes through 4-byte scattered shared-memory stores and loses to the packed-plane
    // engine (dW, dX at 0.82 GFLOP: 24-28 vs 21 us), so those shapes are left to the other two engines.
    // (the layer GEMM, gemm_tl.cu, now takes these shapes first; this engine stays for what it refuses.  The opt-in cluster variant below does not
    // change the break-even: a shape that then fails cluster eligibility would run the plain path under its measured threshold — ADVICE r1)
    return C == 1 && batch == 1 && tA == 0 && tB == 1 && M >= 32 && N >= 32 && K >= 32 && (double)M * N * K >= 1.2e8;
}

// defer: as gemm_simt — the caller runs its own split-K finish over defer->part [splits][M*N] (splits == 1: O holds the product)
int gemm_tcf(const float *A, const float *B, float *O, float alpha, float beta, int tA, int tB,
             int M, int N, int K, cudaStream_t st, GemmDeferred *defer, const GemmEpilogue *epi) {
    const int mtiles = (M + F_BM - 1) / F_BM, ntiles = (N + F_BN - 1) / F_BN, KT = (K + F_BK - 1) / F_BK;
    const int sms = sm_count();
    int splits = 1;
    if (2 * mtiles * ntiles <= sms && KT >= 4) {
        splits = sms / (mtiles * ntiles);                          // one wave: never more CTAs than SMs (1 CTA / SM: 192 KiB of smem)
        if (splits > KT / 2) splits = KT / 2;
        if (splits > 32) splits = 32;
        if (splits < 1) splits = 1;
    }
    // EXPERIMENTAL cluster variant (T4K_TCF_CLUSTER=1): split count = cluster size, a power of two <= 8 (portable), every split non-empty
    bool cluster = false;
    if (tcf_cluster_on() && splits >= 2) {
        int S = 1; while (S * 2 <= splits && S * 2 <= 8) S *= 2;
        const int per = (KT + S - 1) / S;
        if (S >= 2 && (KT + per - 1) / per == S) { splits = S; cluster = true; }
    }
    int kt_per = (KT + splits - 1) / splits;
    splits = (KT + kt_per - 1) / kt_per;
    TcfP p{A, B, O, alpha, beta, M, N, K,
           tA ? 1 : (int64_t)K, tA ? (int64_t)M : 1,          // A [M,K] row-major, or stored [K,M] when tA
           tB ? (int64_t)K : 1, tB ? 1 : (int64_t)N,          // B [K,N] row-major → (n,k) at k*N + n; stored [N,K] when tB
           KT, kt_per, splits, nullptr};
    constexpr size_t smem = (size_t)F_STAGES * F_STAGE_B + 1024 + 256;
    if (cluster) {
        static DevFlag cattr;
        if (dev_first(cattr)) {
            cudaError_t e = cudaFuncSetAttribute(k_gemm_tcf<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
        }
        p.splits = 1;                                                  // the kernel writes O itself (its split count is gridDim.z)
        const bool fused_epi = epi && epi->bias && beta == 0.0f && (epi->layer == T4K_L_NONE || (epi->A && epi->F));
        TcfArgs<true> ca{p, TcfEpi{nullptr, nullptr, nullptr, T4K_L_NONE, 0.0f}};
        if (fused_epi) ca.e = TcfEpi{epi->bias, epi->A, epi->F, epi->layer, epi->alpha};
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(ntiles, mtiles, splits); cfg.blockDim = dim3(F_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = (unsigned)splits;
        cfg.attrs = at; cfg.numAttrs = 1;
        const cudaError_t le = cudaLaunchKernelEx(&cfg, k_gemm_tcf<true>, ca);
        int rc = check_launch();
        if (le == cudaSuccess && rc == 0) {
            if (defer) { defer->part = O; defer->splits = fused_epi ? 0 : 1; }
            return 0;
        }
        cudaGetLastError();                                            // the cluster could not be co-scheduled (or the attribute is refused): the plain launch below
        p.splits = splits;
    }
    if (splits > 1) {
        p.part = (float*)workspace((size_t)splits * M * N * sizeof(float), 7);
        if (!p.part) return T4K_ENOMEM;
    }
    static DevFlag attr;
    if (dev_first(attr)) {
        cudaError_t e = cudaFuncSetAttribute(k_gemm_tcf<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    launch_std(k_gemm_tcf<false>, dim3(ntiles, mtiles, splits), dim3(F_THREADS), smem, st, TcfArgs<false>{p});
    int rc = check_launch();
    if (defer) { defer->part = splits > 1 ? p.part : O; defer->splits = splits; return rc; }
    if (rc || splits == 1) return rc;
    const int64_t MN = (int64_t)M * N;
    const int vec = ((MN & 3) == 0) && aligned16(p.part) && aligned16(O);
    launch_pdl(k_splitk_fin_f, dim3(stream_grid(MN, vec ? 4 : 1)), dim3(T4K_THREADS), 0, st, (const float*)p.part, O, alpha, beta, MN, splits, vec);
    return check_launch();
}

} // namespace t4k
