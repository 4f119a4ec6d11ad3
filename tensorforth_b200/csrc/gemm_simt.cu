// gemm_simt.cu — general FP32-FMA GEMM (CUDA cores): every (tA,tB), channel-interleaved C>1,
// batch/broadcast, arbitrary M,N,K, deterministic split-K for skinny problems.
//   replaces k_gemm / k_gemm_claude / k_gemm_tile_claude(_x2) (src/t4math.cu:370-734) for the
//   shapes the tcgen05 engine (gemm_tc.cu) does not take: small / unaligned / C>1.
// Tile 64x64x16, 256 threads, 4x4 register micro-tile, register-staged double buffering.
// Bound: FP32 FMA pipe (128 FMA/clk/SM); used where launch latency, not math, dominates.
#include "common.cuh"
#include <cstdlib>

namespace t4k {

#define SBM 64
#define SBN 64
#define SBK 16

struct GemmP {
    const float *A, *B; float *O;
    float alpha, beta;
    int M, N, K, C;
    int64_t sA, sB, sO;          // batch strides (floats), 0 = broadcast
    int splits, kchunk;          // split-K: kchunk = K range per split (multiple of SBK)
    float *part;                 // split-K partials [batch*C][splits][M*N] (nullptr when splits==1)
};

template<bool TA, bool TB>
__global__ void __launch_bounds__(256) k_gemm_simt(GemmP p) {
    __shared__ float sA[2][SBK][SBM + 4];
    __shared__ float sB[2][SBK][SBN + 4];
    pdl_wait(); pdl_trigger();                  // PDL: nothing global before this line
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int zc = blockIdx.z % p.C;                          // channel
    const int zs = (blockIdx.z / p.C) % p.splits;             // k-split
    const int zb = blockIdx.z / (p.C * p.splits);             // batch
    const int C = p.C, M = p.M, N = p.N, K = p.K;
    const float *A = p.A + zb * p.sA + zc;
    const float *B = p.B + zb * p.sB + zc;
    const int m0 = blockIdx.y * SBM, n0 = blockIdx.x * SBN;
    const int kbeg = zs * p.kchunk, kend = min(K, kbeg + p.kchunk);

    // loader coordinates: contiguous global dimension on the fast thread index
    // A normal  [M,K]: (m = tid/16 + 16*i, k = tid%16)     A^T [K,M]: (k = tid/64 + 4*i, m = tid%64)
    // B normal  [K,N]: (k = tid/64 + 4*i, n = tid%64)      B^T [N,K]: (n = tid/16 + 16*i, k = tid%16)
    float ra[4], rb[4];
    auto load = [&](int k0) {
        #pragma unroll
        for (int i = 0; i < 4; i++) {
            int m, k;
            if (TA) { k = (tid >> 6) + 4 * i; m = tid & 63; } else { m = (tid >> 4) + 16 * i; k = tid & 15; }
            const int gm = m0 + m, gk = k0 + k;
            ra[i] = (gm < M && gk < kend) ? __ldg(A + (TA ? ((int64_t)gk * M + gm) : ((int64_t)gm * K + gk)) * C) : 0.0f;
            int n, kb;
            if (TB) { n = (tid >> 4) + 16 * i; kb = tid & 15; } else { kb = (tid >> 6) + 4 * i; n = tid & 63; }
            const int gn = n0 + n, gkb = k0 + kb;
            rb[i] = (gn < N && gkb < kend) ? __ldg(B + (TB ? ((int64_t)gn * K + gkb) : ((int64_t)gkb * N + gn)) * C) : 0.0f;
        }
    };
    auto store = [&](int buf) {
        #pragma unroll
        for (int i = 0; i < 4; i++) {
            int m, k;
            if (TA) { k = (tid >> 6) + 4 * i; m = tid & 63; } else { m = (tid >> 4) + 16 * i; k = tid & 15; }
            sA[buf][k][m] = ra[i];
            int n, kb;
            if (TB) { n = (tid >> 4) + 16 * i; kb = tid & 15; } else { kb = (tid >> 6) + 4 * i; n = tid & 63; }
            sB[buf][kb][n] = rb[i];
        }
    };

    float acc[4][4] = {};
    int buf = 0;
    if (kbeg < kend) { load(kbeg); store(0); }
    __syncthreads();
    for (int k0 = kbeg; k0 < kend; k0 += SBK) {
        const bool more = (k0 + SBK) < kend;
        if (more) load(k0 + SBK);
        #pragma unroll
        for (int k = 0; k < SBK; k++) {
            const float4 a4 = *reinterpret_cast<const float4*>(&sA[buf][k][ty * 4]);
            const float4 b4 = *reinterpret_cast<const float4*>(&sB[buf][k][tx * 4]);
            const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
            #pragma unroll
            for (int i = 0; i < 4; i++)
                #pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (more) { store(buf ^ 1); }
        __syncthreads();
        buf ^= 1;
    }
    if (p.splits == 1) {
        float *O = p.O + zb * p.sO + zc;
        #pragma unroll
        for (int i = 0; i < 4; i++) {
            const int gm = m0 + ty * 4 + i;
            if (gm >= M) continue;
            #pragma unroll
            for (int j = 0; j < 4; j++) {
                const int gn = n0 + tx * 4 + j;
                if (gn >= N) continue;
                const int64_t z = ((int64_t)gm * N + gn) * C;
                O[z] = (p.beta == 0.0f) ? acc[i][j] * p.alpha : acc[i][j] * p.alpha + O[z] * p.beta;
            }
        }
    } else {
        float *P = p.part + ((int64_t)(zb * C + zc) * p.splits + zs) * ((int64_t)M * N);
        #pragma unroll
        for (int i = 0; i < 4; i++) {
            const int gm = m0 + ty * 4 + i;
            if (gm >= M) continue;
            #pragma unroll
            for (int j = 0; j < 4; j++) {
                const int gn = n0 + tx * 4 + j;
                if (gn < N) P[(int64_t)gm * N + gn] = acc[i][j];
            }
        }
    }
}

// split-K epilogue: O = alpha * Σ_s part[s] + beta * O   (fixed order → deterministic)
__global__ void __launch_bounds__(T4K_THREADS) k_splitk_fin(GemmP p, int batch) {
    pdl_wait(); pdl_trigger();                  // PDL: nothing global before this line
    const int64_t MN = (int64_t)p.M * p.N, total = MN * p.C * batch;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = t % MN; const int bc = (int)(t / MN); const int b = bc / p.C, c = bc % p.C;
        const float *P = p.part + (int64_t)bc * p.splits * MN + e;
        float s = 0.0f;
        for (int k = 0; k < p.splits; k++) s += P[(int64_t)k * MN];
        float *o = p.O + b * p.sO + e * p.C + c;
        *o = (p.beta == 0.0f) ? s * p.alpha : s * p.alpha + *o * p.beta;
    }
}

// ====================================================================== v2: 8x8 register tile, 128-bit global loads
// C == 1, operands 16-byte aligned with the contiguous extents a multiple of 4.  128 threads = (BM/8) x (BN/8); each thread
// owns rows {ty*4..+3, BM/2+ty*4..+3} x cols {tx*4..+3, BN/2+tx*4..+3} so every shared-memory read is a conflict-free
// 128-bit access; 64 FMA per 4 LDS.128.  Global tiles arrive as float4 along the contiguous dimension (transposed into the
// k-major smem tile when that dimension is K) and are register-staged one k-tile ahead.
#define VBK 16
// register tile TM x TN (4 or 8 each): an 8-wide side is two 4-wide groups half a tile apart (conflict-free LDS.128)
template<bool TA, bool TB, int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN)) k_gemm_v2(GemmP p) {
    constexpr int TX = BN / TN;
    constexpr int NT = (BM / TM) * (BN / TN);                            // threads per CTA: 128, or 256 for the 4 x 4 register tile
    static_assert(NT == 128 || NT == 256, "128 or 256 threads");
    constexpr int NA = BM * VBK / 4 / NT, NB = BN * VBK / 4 / NT;        // float4 loads per thread per k-tile
    static_assert(NA >= 1 && NB >= 1, "tile too small for the CTA");
    __shared__ __align__(16) float sA[2][VBK][BM + 4];
    __shared__ __align__(16) float sB[2][VBK][BN + 4];
    pdl_wait(); pdl_trigger();                  // PDL: nothing global before this line
    const int tid = threadIdx.x, tx = tid % TX, ty = tid / TX;
    const int zs = blockIdx.z % p.splits, zb = blockIdx.z / p.splits;
    const int M = p.M, N = p.N, K = p.K;
    const float *A = p.A + zb * p.sA, *B = p.B + zb * p.sB;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int kbeg = zs * p.kchunk, kend = min(K, kbeg + p.kchunk);
    float4 ra[NA], rb[NB];
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    auto load = [&](int k0) {
        #pragma unroll
        for (int i = 0; i < NA; i++) {
            const int f = tid + NT * i;
            if (TA) { const int k = f / (BM / 4), mq = f % (BM / 4); const int gk = k0 + k, gm = m0 + mq * 4;       // A^T [K,M]: float4 along m
                      ra[i] = (gk < kend && gm < M) ? ldg4(A + (int64_t)gk * M + gm) : z4; }
            else    { const int m = f / (VBK / 4), kq = f % (VBK / 4); const int gm = m0 + m, gk = k0 + kq * 4;      // A [M,K]: float4 along k
                      ra[i] = (gm < M && gk < kend) ? ldg4(A + (int64_t)gm * K + gk) : z4; }
        }
        #pragma unroll
        for (int i = 0; i < NB; i++) {
            const int f = tid + NT * i;
            if (TB) { const int n = f / (VBK / 4), kq = f % (VBK / 4); const int gn = n0 + n, gk = k0 + kq * 4;      // B^T [N,K]: float4 along k
                      rb[i] = (gn < N && gk < kend) ? ldg4(B + (int64_t)gn * K + gk) : z4; }
            else    { const int k = f / (BN / 4), nq = f % (BN / 4); const int gk = k0 + k, gn = n0 + nq * 4;        // B [K,N]: float4 along n
                      rb[i] = (gk < kend && gn < N) ? ldg4(B + (int64_t)gk * N + gn) : z4; }
        }
    };
    auto store = [&](int buf) {
        #pragma unroll
        for (int i = 0; i < NA; i++) {
            const int f = tid + NT * i;
            if (TA) { const int k = f / (BM / 4), mq = f % (BM / 4); *reinterpret_cast<float4*>(&sA[buf][k][mq * 4]) = ra[i]; }
            else    { const int m = f / (VBK / 4), kq = f % (VBK / 4);
                      sA[buf][kq * 4][m] = ra[i].x; sA[buf][kq * 4 + 1][m] = ra[i].y; sA[buf][kq * 4 + 2][m] = ra[i].z; sA[buf][kq * 4 + 3][m] = ra[i].w; }
        }
        #pragma unroll
        for (int i = 0; i < NB; i++) {
            const int f = tid + NT * i;
            if (TB) { const int n = f / (VBK / 4), kq = f % (VBK / 4);
                      sB[buf][kq * 4][n] = rb[i].x; sB[buf][kq * 4 + 1][n] = rb[i].y; sB[buf][kq * 4 + 2][n] = rb[i].z; sB[buf][kq * 4 + 3][n] = rb[i].w; }
            else    { const int k = f / (BN / 4), nq = f % (BN / 4); *reinterpret_cast<float4*>(&sB[buf][k][nq * 4]) = rb[i]; }
        }
    };
    float acc[TM][TN];
    #pragma unroll
    for (int i = 0; i < TM; i++)
        #pragma unroll
        for (int j = 0; j < TN; j++) acc[i][j] = 0.0f;
    int buf = 0;
    if (kbeg < kend) { load(kbeg); store(0); }
    __syncthreads();
    for (int k0 = kbeg; k0 < kend; k0 += VBK) {
        const bool more = (k0 + VBK) < kend;
        if (more) load(k0 + VBK);
        #pragma unroll
        for (int k = 0; k < VBK; k++) {
            float a[TM], b[TN];
            #pragma unroll
            for (int g = 0; g < TM / 4; g++) {
                const float4 v = *reinterpret_cast<const float4*>(&sA[buf][k][g * (BM / 2) + ty * 4]);
                a[g * 4] = v.x; a[g * 4 + 1] = v.y; a[g * 4 + 2] = v.z; a[g * 4 + 3] = v.w;
            }
            #pragma unroll
            for (int g = 0; g < TN / 4; g++) {
                const float4 v = *reinterpret_cast<const float4*>(&sB[buf][k][g * (BN / 2) + tx * 4]);
                b[g * 4] = v.x; b[g * 4 + 1] = v.y; b[g * 4 + 2] = v.z; b[g * 4 + 3] = v.w;
            }
            #pragma unroll
            for (int i = 0; i < TM; i++)
                #pragma unroll
                for (int j = 0; j < TN; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (more) store(buf ^ 1);
        __syncthreads();
        buf ^= 1;
    }
    const bool fin = p.splits == 1;
    float *O = fin ? p.O + zb * p.sO : p.part + ((int64_t)zb * p.splits + zs) * ((int64_t)M * N);
    const float alpha = fin ? p.alpha : 1.0f, beta = fin ? p.beta : 0.0f;
    #pragma unroll
    for (int i = 0; i < TM; i++) {
        const int gm = m0 + (i / 4) * (BM / 2) + ty * 4 + (i & 3);
        if (gm >= M) continue;
        #pragma unroll
        for (int h = 0; h < TN / 4; h++) {
            const int gn = n0 + h * (BN / 2) + tx * 4;
            if (gn >= N) continue;                                    // N % 4 == 0: a float4 is all in or all out
            float *o = O + (int64_t)gm * N + gn;
            float4 r = make_float4(acc[i][h * 4] * alpha, acc[i][h * 4 + 1] * alpha, acc[i][h * 4 + 2] * alpha, acc[i][h * 4 + 3] * alpha);
            if (beta != 0.0f) { const float4 q = *reinterpret_cast<const float4*>(o); r.x += q.x * beta; r.y += q.y * beta; r.z += q.z * beta; r.w += q.w * beta; }
            stg4(o, r);
        }
    }
}
template<int BM, int BN, int TM, int TN> static void launch_v2(const GemmP &p, dim3 g, int tA, int tB, cudaStream_t st) {
    constexpr int NT = (BM / TM) * (BN / TN);
    if (tA) { if (tB) launch_std(k_gemm_v2<true, true, BM, BN, TM, TN>, g, dim3(NT), 0, st, p); else launch_std(k_gemm_v2<true, false, BM, BN, TM, TN>, g, dim3(NT), 0, st, p); }
    else    { if (tB) launch_std(k_gemm_v2<false, true, BM, BN, TM, TN>, g, dim3(NT), 0, st, p); else launch_std(k_gemm_v2<false, false, BM, BN, TM, TN>, g, dim3(NT), 0, st, p); }
}
static bool v2_ok(const float *A, const float *B, const float *O, int tA, int tB, int M, int N, int K, int C, int64_t sA, int64_t sB, int64_t sO) {
    if (C != 1 || (N & 3) || !aligned16(A) || !aligned16(B) || !aligned16(O) || (sA & 3) || (sB & 3) || (sO & 3)) return false;
    if ((tA ? M : K) & 3) return false;                               // A's contiguous extent
    if ((tB ? K : N) & 3) return false;                               // B's contiguous extent
    return (int64_t)M * N >= 4096 && K >= 16;
}

// defer != nullptr: the caller runs its own split-K finish (fused epilogue): on return defer->part / defer->splits describe
// the partials [splits][M*N] (C == 1, batch == 1); splits == 1 means O already holds alpha*A@B + beta*O.
int gemm_simt(const float *A, const float *B, float *O, float alpha, float beta, int tA, int tB,
              int M, int N, int K, int C, int batch, int64_t sA, int64_t sB, int64_t sO, cudaStream_t st, GemmDeferred *defer) {
    GemmP p{A, B, O, alpha, beta, M, N, K, C, sA, sB, sO, 1, K, nullptr};
    const bool v2 = v2_ok(A, B, O, tA, tB, M, N, K, C, sA, sB, sO);
    // v2 tile: 64x128 (8x8 per thread) when that alone gives >= 2 CTAs per SM, else 64x64 (8x4 per thread: twice the CTAs,
    // so a second CTA computes while the first waits on its next k-tile); 128x64 for narrow outputs
    const bool wide = N > 64;
    const bool big = wide && (int64_t)((M + 63) / 64) * ((N + 127) / 128) * batch >= 2 * sm_count();
    const int TBM = v2 ? (wide ? 64 : 128) : SBM, TBN = v2 ? (wide ? (big ? 128 : 64) : 64) : SBN;
    const int gx = (N + TBN - 1) / TBN, gy = (M + TBM - 1) / TBM;
    const int64_t ctas = (int64_t)gx * gy * C * batch;
    // split K when the output grid cannot fill the machine and K is deep
    int splits = 1;
    const int sms = sm_count();
    static int mult = 0;                                  // CTAs per SM the split aims at (T4K_SIMT_SPLIT_MULT, measured default below)
    if (!mult) { const char *e = getenv("T4K_SIMT_SPLIT_MULT"); mult = e ? atoi(e) : 2; if (mult < 1 || mult > 8) mult = 2; }
    if (ctas < sms && K >= 8 * SBK) {
        splits = (int)((mult * sms + ctas - 1) / ctas);
        const int maxs = K / (4 * SBK);
        if (splits > maxs) splits = maxs;
        if (splits > 64) splits = 64;
        if (splits < 1) splits = 1;
    }
    if (splits > 1) {
        int kchunk = (K + splits - 1) / splits;
        kchunk = (kchunk + SBK - 1) / SBK * SBK;
        splits = (K + kchunk - 1) / kchunk;
        p.kchunk = kchunk;
    }
    p.splits = splits;
    if (splits > 1) {
        p.part = (float*)workspace((size_t)batch * C * splits * M * N * sizeof(float), 0);
        if (!p.part) return T4K_ENOMEM;
    }
    if ((int64_t)C * splits * batch > 65535) return T4K_EINVAL;
    dim3 g(gx, gy, C * splits * batch);
    if (K == 0) { p.splits = 1; }
    // 64 x 64 tile when the grid is small (the layer products): 128 threads with an 8 x 4 register tile.  A 256-thread 4 x 4 variant
    // (twice the warps per SM) was measured SLOWER on the MNIST step (linear_bwd 26.1 vs 24.5 us, step 84.3 vs 80.1 us): opt-in
    // with T4K_SIMT_T256=1
    static int t256 = -1;
    if (t256 < 0) { const char *e = getenv("T4K_SIMT_T256"); t256 = (e && e[0] == '1') ? 1 : 0; }
    if (v2) { if (!wide) launch_v2<128, 64, 8, 8>(p, g, tA, tB, st); else if (big) launch_v2<64, 128, 8, 8>(p, g, tA, tB, st);
              else if (t256) launch_v2<64, 64, 4, 4>(p, g, tA, tB, st); else launch_v2<64, 64, 8, 4>(p, g, tA, tB, st); }
    else if (tA) { if (tB) launch_std(k_gemm_simt<true, true >, g, dim3(256), 0, st, p); else launch_std(k_gemm_simt<true, false>, g, dim3(256), 0, st, p); }
    else    { if (tB) launch_std(k_gemm_simt<false, true>, g, dim3(256), 0, st, p); else launch_std(k_gemm_simt<false, false>, g, dim3(256), 0, st, p); }
    int rc = check_launch();
    if (defer) { defer->part = p.splits > 1 ? p.part : O; defer->splits = p.splits; return rc; }
    if (rc || p.splits == 1) return rc;
    const int64_t total = (int64_t)M * N * C * batch;
    launch_pdl(k_splitk_fin, dim3(stream_grid(total)), dim3(T4K_THREADS), 0, st, p, batch);
    return check_launch();
}

} // namespace t4k
