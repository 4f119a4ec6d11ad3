// gemm_mma.cu — FP32 GEMM for the LAYER-SIZED products (0.05 - 1 GFLOP: linear forward / dW / dX at batch 512-1024) on the
// warp-level tensor-core MMA (mma.sync m16n8k8, TF32 operands, FP32 accumulate), "3xTF32": a = hi + lo, a*b ~ lo*hi + hi*lo + hi*hi.
//   replaces k_gemm_tile_claude (src/t4math.cu:478-583) as launched by Tensor::linear / Model::_flinear / _blinear
//   (src/mu/tensor.cu:74-87, src/nn/forward.cu:158-198, src/nn/backprop.cu:194-254).
//
// EXPERIMENT, measured and NOT adopted by the automatic engine choice (see layer_mma below for the numbers).
// Why it was tried next to gemm_tc.cu / gemm_tcf.cu (tcgen05): at these sizes a call lasts 7 - 20 us and is bound by
// its fixed costs, not by the tensor pipe.  The tcgen05 kernels pay TMEM allocation, mbarrier ring set-up, a 672-thread CTA with
// 193 KiB of shared memory (1 CTA / SM, one wave) and a separate split-K finish launch: ~10 us before the first useful flop
// (profiles/r01_gemm_probe.txt).  Here a CTA is 4 warps and 36 KiB (5 CTAs / SM), operands go global -> registers -> shared ->
// fragments with no descriptor or proxy fence in between, and the split-K reduction is done by the LAST CTA of each output tile
// in fixed split order (deterministic, no second launch).  The big-problem engines stay on tcgen05 (4096^3: 394 TFLOP/s).
//
// Tile 64 x 64 x 32, 4 warps as 2 x 2, warp tile 32 x 32 = 2 (m16) x 4 (n8) MMA tiles.  Any (tA, tB): an operand whose rows are
// K-contiguous is staged as [row][36] (fragment loads hit banks 4g + t), one that is contiguous along M / N as [k][72] (banks
// 8t + g): both conflict-free, both filled with 128-bit global loads and 128-bit shared stores.  hi/lo split (cvt.rna) happens on the
// fragments in registers.  The tensor core's accumulator add truncates (gemm_tc.cu), so every 32-deep k-block starts a fresh
// accumulator (12 chained MMAs) that is then added to the running FP32 sum with round-to-nearest adds.
// Bound: launch + one global round trip per k-block for the layer shapes; issue rate of the legacy MMA path for long K.
#include "common.cuh"
#include <mutex>
#include <cstdlib>

namespace t4k {

constexpr int G_BM = 64, G_BN = 64, G_BK = 32;
constexpr int G_LDK = G_BK + 4;                 // [row][36]  K-contiguous staging
constexpr int G_LDR = G_BM + 8;                 // [k][72]    row-contiguous staging
constexpr int G_TILE = 64 * G_LDK;              // = 32 * G_LDR = 2304 floats per operand per stage
static_assert(64 * G_LDK == 32 * G_LDR, "both staging layouts share one buffer size");
constexpr int G_THREADS = 128;

struct MmaP {
    const float *A, *B;
    float *O, *part;             // part: [splits][M*N] when splits > 1
    int *cnt;                    // per output tile arrival counters (fused finish), nullptr: caller finishes
    float alpha, beta;
    int M, N, K;
    int64_t a_sr, a_sk;          // op(A)(m,k) = A[m*a_sr + k*a_sk]
    int64_t b_sr, b_sk;          // op(B)(k,n) = B[n*b_sr + k*b_sk]
    int KT, kb_per_split, splits;
    int avec, bvec;              // 128-bit global loads allowed (alignment + stride)
};

__device__ __forceinline__ uint32_t cvt_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void split3(float x, uint32_t &hi, uint32_t &lo) {
    hi = cvt_tf32(x);
    lo = cvt_tf32(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// global -> registers: this thread's 4 float4 of one 64-row x 32-k operand tile.  KC: rows are K-contiguous (sk == 1)
template<bool KC>
__device__ __forceinline__ void tile_load(const float *__restrict__ X, int64_t sr, int64_t sk, int R, int K, int r0, int k0, int vec, int t, float4 (&v)[4]) {
    #pragma unroll
    for (int i = 0; i < 4; i++) {
        const int idx = t + G_THREADS * i;
        v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (KC) {
            const int r = idx >> 3, gr = r0 + r, gk = k0 + ((idx & 7) << 2);
            if (gr < R && gk < K) {
                const float *src = X + (int64_t)gr * sr + gk;
                if (vec && gk + 3 < K) v[i] = ldg4(src);
                else { v[i].x = __ldg(src); if (gk + 1 < K) v[i].y = __ldg(src + 1); if (gk + 2 < K) v[i].z = __ldg(src + 2); if (gk + 3 < K) v[i].w = __ldg(src + 3); }
            }
        } else {
            const int k = idx >> 4, gk = k0 + k, gr = r0 + ((idx & 15) << 2);
            if (gk < K && gr < R) {
                const float *src = X + (int64_t)gk * sk + (int64_t)gr * sr;
                if (vec && gr + 3 < R) v[i] = ldg4(src);
                else { v[i].x = __ldg(src); if (gr + 1 < R) v[i].y = __ldg(src + sr); if (gr + 2 < R) v[i].z = __ldg(src + 2 * sr); if (gr + 3 < R) v[i].w = __ldg(src + 3 * sr); }
            }
        }
    }
}
template<bool KC>
__device__ __forceinline__ void tile_store(float *s, int t, const float4 (&v)[4]) {
    #pragma unroll
    for (int i = 0; i < 4; i++) {
        const int idx = t + G_THREADS * i;
        const int o = KC ? (idx >> 3) * G_LDK + ((idx & 7) << 2) : (idx >> 4) * G_LDR + ((idx & 15) << 2);
        *reinterpret_cast<float4*>(s + o) = v[i];
    }
}
// element (row r, k) of a staged tile
template<bool KC> __device__ __forceinline__ float tile_at(const float *s, int r, int k) { return KC ? s[r * G_LDK + k] : s[k * G_LDR + r]; }

template<bool AKC, bool BKC>
__global__ void __launch_bounds__(G_THREADS, 3) k_gemm_mma(MmaP p) {
    __shared__ __align__(16) float sA[2][G_TILE];
    __shared__ __align__(16) float sB[2][G_TILE];
    __shared__ int s_last;
    pdl_wait(); pdl_trigger();
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31, g = lane >> 2, tq = lane & 3;
    const int wm = warp >> 1, wn = warp & 1;
    const int nt = blockIdx.x, mt = blockIdx.y, zs = blockIdx.z;
    const int r0 = mt * G_BM, c0 = nt * G_BN;
    const int kb0 = zs * p.kb_per_split;
    const int kb1 = min(p.KT, kb0 + p.kb_per_split);
    const int nkb = kb1 - kb0;

    float sum[2][4][4];
    #pragma unroll
    for (int i = 0; i < 2; i++)
        #pragma unroll
        for (int j = 0; j < 4; j++)
            #pragma unroll
            for (int e = 0; e < 4; e++) sum[i][j][e] = 0.0f;

    float4 ra[4], rb[4];
    if (nkb > 0) {
        tile_load<AKC>(p.A, p.a_sr, p.a_sk, p.M, p.K, r0, kb0 * G_BK, p.avec, t, ra);
        tile_load<BKC>(p.B, p.b_sr, p.b_sk, p.N, p.K, c0, kb0 * G_BK, p.bvec, t, rb);
        tile_store<AKC>(sA[0], t, ra);
        tile_store<BKC>(sB[0], t, rb);
    }
    __syncthreads();
    for (int i = 0; i < nkb; i++) {
        const int cur = i & 1;
        const bool more = (i + 1 < nkb);
        if (more) {
            tile_load<AKC>(p.A, p.a_sr, p.a_sk, p.M, p.K, r0, (kb0 + i + 1) * G_BK, p.avec, t, ra);
            tile_load<BKC>(p.B, p.b_sr, p.b_sk, p.N, p.K, c0, (kb0 + i + 1) * G_BK, p.bvec, t, rb);
        }
        const float *a = sA[cur], *b = sB[cur];
        float c[2][4][4];
        #pragma unroll
        for (int ii = 0; ii < 2; ii++)
            #pragma unroll
            for (int j = 0; j < 4; j++)
                #pragma unroll
                for (int e = 0; e < 4; e++) c[ii][j][e] = 0.0f;
        #pragma unroll
        for (int ks = 0; ks < G_BK / 8; ks++) {
            const int kk = ks * 8;
            uint32_t ahi[2][4], alo[2][4], bhi[4][2], blo[4][2];
            #pragma unroll
            for (int ii = 0; ii < 2; ii++) {
                const int mb = wm * 32 + ii * 16;
                split3(tile_at<AKC>(a, mb + g,     kk + tq),     ahi[ii][0], alo[ii][0]);
                split3(tile_at<AKC>(a, mb + g + 8, kk + tq),     ahi[ii][1], alo[ii][1]);
                split3(tile_at<AKC>(a, mb + g,     kk + tq + 4), ahi[ii][2], alo[ii][2]);
                split3(tile_at<AKC>(a, mb + g + 8, kk + tq + 4), ahi[ii][3], alo[ii][3]);
            }
            #pragma unroll
            for (int j = 0; j < 4; j++) {
                const int nb = wn * 32 + j * 8;
                split3(tile_at<BKC>(b, nb + g, kk + tq),     bhi[j][0], blo[j][0]);
                split3(tile_at<BKC>(b, nb + g, kk + tq + 4), bhi[j][1], blo[j][1]);
            }
            #pragma unroll
            for (int ii = 0; ii < 2; ii++)
                #pragma unroll
                for (int j = 0; j < 4; j++) {
                    mma_tf32(c[ii][j], alo[ii], bhi[j]);          // small terms first
                    mma_tf32(c[ii][j], ahi[ii], blo[j]);
                    mma_tf32(c[ii][j], ahi[ii], bhi[j]);
                }
        }
        #pragma unroll
        for (int ii = 0; ii < 2; ii++)
            #pragma unroll
            for (int j = 0; j < 4; j++)
                #pragma unroll
                for (int e = 0; e < 4; e++) sum[ii][j][e] += c[ii][j][e];
        if (more) {
            tile_store<AKC>(sA[cur ^ 1], t, ra);
            tile_store<BKC>(sB[cur ^ 1], t, rb);
        }
        __syncthreads();
    }

    // ---- epilogue: accumulator fragment (row g / g+8, cols 2*tq, 2*tq+1 of each 16 x 8 tile)
    const bool direct = (p.splits == 1);
    float *dst = direct ? p.O : p.part + (int64_t)zs * p.M * p.N;
    const float alpha = direct ? p.alpha : 1.0f, beta = direct ? p.beta : 0.0f;
    const bool pair = ((p.N & 1) == 0) && ((((uintptr_t)dst) & 7) == 0);
    #pragma unroll
    for (int ii = 0; ii < 2; ii++) {
        #pragma unroll
        for (int h = 0; h < 2; h++) {
            const int row = r0 + wm * 32 + ii * 16 + g + 8 * h;
            if (row >= p.M) continue;
            #pragma unroll
            for (int j = 0; j < 4; j++) {
                const int col = c0 + wn * 32 + j * 8 + 2 * tq;
                if (col >= p.N) continue;
                float *o = dst + (int64_t)row * p.N + col;
                float v0 = sum[ii][j][2 * h] * alpha, v1 = sum[ii][j][2 * h + 1] * alpha;
                if (pair && col + 1 < p.N) {
                    if (beta != 0.0f) { const float2 q = *reinterpret_cast<const float2*>(o); v0 += q.x * beta; v1 += q.y * beta; }
                    *reinterpret_cast<float2*>(o) = make_float2(v0, v1);
                } else {
                    if (beta != 0.0f) v0 += o[0] * beta;
                    o[0] = v0;
                    if (col + 1 < p.N) { if (beta != 0.0f) v1 += o[1] * beta; o[1] = v1; }
                }
            }
        }
    }
    if (direct || !p.cnt) return;

    // ---- fused split-K finish: the last CTA to arrive on this output tile adds the partials in split order
    __threadfence();
    __syncthreads();
    int *cnt = p.cnt + (mt * gridDim.x + nt);
    if (t == 0) { const int old = atomicAdd(cnt, 1); s_last = (old == p.splits - 1); }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const int64_t MN = (int64_t)p.M * p.N;
    const bool v4 = ((p.N & 3) == 0) && ((((uintptr_t)p.part) & 15) == 0) && ((((uintptr_t)p.O) & 15) == 0);
    if (v4) {
        for (int q = t; q < G_BM * (G_BN / 4); q += G_THREADS) {
            const int row = r0 + (q >> 4), col = c0 + ((q & 15) << 2);
            if (row >= p.M || col >= p.N) continue;                   // N % 4 == 0: a quad is all in or all out
            const int64_t e = (int64_t)row * p.N + col;
            float4 s = __ldcg(reinterpret_cast<const float4*>(p.part + e));
            for (int k = 1; k < p.splits; k++) { const float4 x = __ldcg(reinterpret_cast<const float4*>(p.part + (int64_t)k * MN + e)); s.x += x.x; s.y += x.y; s.z += x.z; s.w += x.w; }
            s.x *= p.alpha; s.y *= p.alpha; s.z *= p.alpha; s.w *= p.alpha;
            if (p.beta != 0.0f) { const float4 o = *reinterpret_cast<const float4*>(p.O + e); s.x += o.x * p.beta; s.y += o.y * p.beta; s.z += o.z * p.beta; s.w += o.w * p.beta; }
            stg4(p.O + e, s);
        }
    } else {
        for (int q = t; q < G_BM * G_BN; q += G_THREADS) {
            const int row = r0 + (q >> 6), col = c0 + (q & 63);
            if (row >= p.M || col >= p.N) continue;
            const int64_t e = (int64_t)row * p.N + col;
            float s = __ldcg(p.part + e);
            for (int k = 1; k < p.splits; k++) s += __ldcg(p.part + (int64_t)k * MN + e);
            s *= p.alpha;
            if (p.beta != 0.0f) s += p.O[e] * p.beta;
            p.O[e] = s;
        }
    }
    if (t == 0) *cnt = 0;                                             // ready for the next call (stream order)
}

// arrival counters: ring of zero-initialised ints, every call takes `n` consecutive ones; the finishing CTA re-zeroes its own
#define MMA_CNT_RING (1 << 16)
static int *g_cnt[16];
static unsigned g_cnt_next[16];
static std::mutex g_cnt_mu;
static int *tile_counters(int n) {
    int dev = 0;
    if (n > MMA_CNT_RING || cudaGetDevice(&dev) != cudaSuccess || dev >= 16) return nullptr;
    std::lock_guard<std::mutex> lk(g_cnt_mu);
    if (!g_cnt[dev]) {
        void *q = nullptr;
        if (cudaMalloc(&q, MMA_CNT_RING * sizeof(int)) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        cudaMemset(q, 0, MMA_CNT_RING * sizeof(int));
        g_cnt[dev] = (int*)q;
    }
    if (g_cnt_next[dev] + n > MMA_CNT_RING) g_cnt_next[dev] = 0;
    int *r = g_cnt[dev] + g_cnt_next[dev];
    g_cnt_next[dev] += n;
    return r;
}

bool gemm_mma_ok(int M, int N, int K, int C, int batch) {
    return C == 1 && batch == 1 && M >= 1 && N >= 1 && K >= 1 && (int64_t)((M + G_BM - 1) / G_BM) * ((N + G_BN - 1) / G_BN) <= 60000;
}

// AUTO policy.  MEASURED (bench_scripts/gemm_probe.py -> profiles/r01_gemm_probe_v2.txt): the legacy MMA path of sm_100 sustains
// ~100-135 TFLOP/s of TF32 (1024^3: 33 TFLOP/s x 3 MMAs, 2048^3: 45 x 3), i.e. a 3xTF32 product runs no faster than the FP32-FMA
// kernel, and the serial last-CTA finish costs more than a second launch (1960->100 forward: 22 us vs 14.7).  It wins on two of
// the 18 layer shapes only (128->256 forward 6.8 vs 8.3 us), so AUTO does NOT take it: opt-in with T4K_GEMM_MMA=1 or engine
// T4K_GEMM_MMA.  Kept as the measured reference point for "tensor cores without tcgen05" and as a second FP32-grade checker.
bool layer_mma(int M, int N, int K, int C, int batch) {
    static int on = -1;
    if (on < 0) { const char *e = getenv("T4K_GEMM_MMA"); on = (e && e[0] == '1') ? 1 : 0; }
    if (!on || !gemm_mma_ok(M, N, K, C, batch)) return false;
    const double w = (double)M * N * K;
    return M >= 32 && N >= 32 && K >= 32 && w >= 4.0e6 && w < 1.5e9;
}

// defer: as gemm_simt — the caller runs its own split-K finish over defer->part [splits][M*N] (splits == 1: O holds the product)
int gemm_mma(const float *A, const float *B, float *O, float alpha, float beta, int tA, int tB,
             int M, int N, int K, cudaStream_t st, GemmDeferred *defer) {
    const int mtiles = (M + G_BM - 1) / G_BM, ntiles = (N + G_BN - 1) / G_BN, KT = (K + G_BK - 1) / G_BK;
    const int tiles = mtiles * ntiles, sms = sm_count();
    if (mtiles > 65535) return T4K_EINVAL;
    // split K until ~2 CTAs per SM are in flight (5 fit), never below 4 k-blocks per CTA
    int splits = 1;
    if (tiles < 2 * sms && KT >= 8) {
        splits = (2 * sms + tiles - 1) / tiles;
        if (splits > KT / 4) splits = KT / 4;
        if (splits > 32) splits = 32;
        if (splits < 1) splits = 1;
    }
    const int kb_per = (KT + splits - 1) / splits;
    splits = (KT + kb_per - 1) / kb_per;
    MmaP p{};
    p.A = A; p.B = B; p.O = O; p.alpha = alpha; p.beta = beta; p.M = M; p.N = N; p.K = K;
    p.a_sr = tA ? 1 : (int64_t)K; p.a_sk = tA ? (int64_t)M : 1;            // A [M,K] row-major, or stored [K,M] when tA
    p.b_sr = tB ? (int64_t)K : 1; p.b_sk = tB ? 1 : (int64_t)N;            // B [K,N] row-major, or stored [N,K] when tB
    p.KT = KT; p.kb_per_split = kb_per; p.splits = splits;
    p.avec = aligned16(A) && (((tA ? M : K) & 3) == 0);
    p.bvec = aligned16(B) && (((tB ? K : N) & 3) == 0);
    if (splits > 1) {
        p.part = (float*)workspace((size_t)splits * M * N * sizeof(float), 7);        // slot shared with gemm_tcf (same role, never both in flight)
        if (!p.part) return T4K_ENOMEM;
        if (!defer) { p.cnt = tile_counters(tiles); if (!p.cnt) return T4K_ENOMEM; }
    }
    const dim3 grid(ntiles, mtiles, splits), block(G_THREADS);
    if (tA) { if (tB) launch_std(k_gemm_mma<false, true >, grid, block, 0, st, p); else launch_std(k_gemm_mma<false, false>, grid, block, 0, st, p); }
    else    { if (tB) launch_std(k_gemm_mma<true,  true >, grid, block, 0, st, p); else launch_std(k_gemm_mma<true,  false>, grid, block, 0, st, p); }
    int rc = check_launch();
    if (defer) { defer->part = splits > 1 ? p.part : O; defer->splits = splits; }
    return rc;
}

} // namespace t4k
