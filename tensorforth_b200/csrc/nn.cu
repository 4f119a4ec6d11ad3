// nn.cu — linear layer forward/backward on top of the GEMM engines, with the layer-group fusions the launch-bound
// MLP tails of the reference's examples need (examples/t4_40a.4th:12-13 "100 linear relu 10 linear softmax"):
//   Model::_flinear (src/nn/forward.cu:158-198)  : Y = X @ W^T + B      (Tensor::linear tB=true, then k_bias)
//   Model::_factivate (forward.cu:201-209)        : fused into the split-K finish of the GEMM (t4k_linear_act_fwd)
//   Model::_blinear (src/nn/backprop.cu:194-254) : dB += ΣdY ; dW += dY^T @ X (beta=1) ; dX = dY @ W
//   head  forward  (t4k_mlp_head_fwd): small linear + bias + row softmax, one launch (forward.cu:158-198,231-243)
//   head  backward (t4k_mlp_head_bwd): _bprep (p - y), softmax pass-through copy, small _blinear (dB, dW, dX),
//                  the preceding _bactivate (dX * mask) and the dB of the linear before it, one launch
//                  (backprop.cu:76-140,194-263)
// Every layer tensor the per-layer path writes is still written with the same meaning.
#include "act.cuh"

namespace t4k {


// ------------------------------------------------------------------ split-K finish + bias (+ activation)
// part: [splits][MN] partial products (splits == 1: the finished product itself); Y = Σ part + bias; A,F = act(Y)
template<int L>
__global__ void __launch_bounds__(T4K_THREADS) k_linear_fin(const float *__restrict__ part, int splits, int64_t MN, int E0,
                                                            const float *__restrict__ bias, float *Y, float *A, float *F, float alpha, int vec) {
    pdl_wait(); pdl_trigger();                  // PDL: nothing global before this line
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    if (vec) {
        for (int64_t q = tid; q < (MN >> 2); q += nth) {
            float4 s = ldg4(part + 4 * q);
            for (int k = 1; k < splits; k++) { const float4 t = ldg4(part + (int64_t)k * MN + 4 * q); s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w; }
            const float4 b = ldg4(bias + (int)((4 * q) % E0));
            s.x += b.x; s.y += b.y; s.z += b.z; s.w += b.w;
            stg4(Y + 4 * q, s);
            if (L != T4K_L_NONE) {
                float4 o, f;
                if (L == T4K_L_DROPOUT) f = *reinterpret_cast<const float4*>(F + 4 * q);
                act<L>(s.x, alpha, o.x, f.x); act<L>(s.y, alpha, o.y, f.y); act<L>(s.z, alpha, o.z, f.z); act<L>(s.w, alpha, o.w, f.w);
                stg4(A + 4 * q, o); stg4(F + 4 * q, f);
            }
        }
    } else {
        for (int64_t e = tid; e < MN; e += nth) {
            float s = part[e];
            for (int k = 1; k < splits; k++) s += part[(int64_t)k * MN + e];
            s += __ldg(bias + (int)(e % E0));
            Y[e] = s;
            if (L != T4K_L_NONE) { float o, f = (L == T4K_L_DROPOUT) ? F[e] : 0.0f; act<L>(s, alpha, o, f); A[e] = o; F[e] = f; }
        }
    }
}
template<int L> static int launch_fin(const GemmDeferred &d, int64_t MN, int E0, const float *B, float *Y, float *A, float *F, float alpha, cudaStream_t st) {
    const int vec = ((E0 & 3) == 0) && aligned16(d.part) && aligned16(B) && aligned16(Y) && (L == T4K_L_NONE || (aligned16(A) && aligned16(F)));
    launch_pdl(k_linear_fin<L>, dim3(stream_grid(MN, vec ? 4 : 1)), dim3(T4K_THREADS), 0, st, d.part, d.splits, MN, E0, B, Y, A, F, alpha, (int)vec);
    return check_launch();
}
static int linear_act_fwd(int layer, const float *X, const float *W, const float *B, float *Y, float *A, float *F, float alpha,
                          int N, int E0, int E1, cudaStream_t st) {
    if (gemm_tl_ok(X, W, Y, 0, 1, N, E0, E1, 1, 1)) {       // the layer GEMM: bias + activation in its cluster epilogue, one launch
        TlEpi e{}; e.mode = 1; e.bias = B; e.actA = A; e.actF = F; e.layer = layer; e.act_alpha = alpha;
        return gemm_tl(X, W, Y, 1.0f, 0.0f, 0, 1, N, E0, E1, st, &e);
    }
    GemmDeferred d{nullptr, 1};
    const GemmEpilogue epi{B, A, F, layer, alpha};         // taken by the (opt-in) cluster variant of gemm_tcf only: reports d.splits == 0
    int rc = gemm_tcf_ok(0, 1, N, E0, E1, 1, 1) ? gemm_tcf(X, W, Y, 1.0f, 0.0f, 0, 1, N, E0, E1, st, &d, &epi)
                                                : gemm_simt(X, W, Y, 1.0f, 0.0f, 0, 1, N, E0, E1, 1, 1, 0, 0, 0, st, &d);
    if (rc || d.splits == 0) return rc;
    const int64_t MN = (int64_t)N * E0;
    switch (layer) {
    case T4K_L_NONE:    return launch_fin<T4K_L_NONE>(d, MN, E0, B, Y, A, F, alpha, st);
    case T4K_L_RELU:    return launch_fin<T4K_L_RELU>(d, MN, E0, B, Y, A, F, alpha, st);
    case T4K_L_TANH:    return launch_fin<T4K_L_TANH>(d, MN, E0, B, Y, A, F, alpha, st);
    case T4K_L_SIGMOID: return launch_fin<T4K_L_SIGMOID>(d, MN, E0, B, Y, A, F, alpha, st);
    case T4K_L_SELU:    return launch_fin<T4K_L_SELU>(d, MN, E0, B, Y, A, F, alpha, st);
    case T4K_L_LEAKYRL: return launch_fin<T4K_L_LEAKYRL>(d, MN, E0, B, Y, A, F, alpha, st);
    case T4K_L_ELU:     return launch_fin<T4K_L_ELU>(d, MN, E0, B, Y, A, F, alpha, st);
    case T4K_L_DROPOUT: return launch_fin<T4K_L_DROPOUT>(d, MN, E0, B, Y, A, F, alpha, st);
    default:            return T4K_EINVAL;
    }
}

// ------------------------------------------------------------------ transposing warp reduction (see cnn_block.cu)
__device__ __forceinline__ float warp_treduce32(float (&v)[32], int lane) {
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

// ------------------------------------------------------------------ head forward: Y = X @ W^T + B ; P = softmax(Y)
// warp per row, W (E0 x E1, E0 <= 32) in shared memory; lane k ends up owning class k.
#define HEAD_JMAX 8                      // E1 <= 32 * HEAD_JMAX
__global__ void __launch_bounds__(T4K_THREADS) k_head_fwd(const float *__restrict__ X, const float *__restrict__ W, const float *__restrict__ B,
                                                          float *Y, float *P, float *P2, int N, int E0, int E1) {
    extern __shared__ float sW[];                      // [E0][E1]
    pdl_wait(); pdl_trigger();                  // PDL: nothing global before this line
    for (int t = threadIdx.x; t < E0 * E1; t += blockDim.x) cp_async4(sW + t, W + t, true);     // all copies in flight at once (a load -> store loop pays one L2 round trip per trip)
    cp_async_commit(); cp_async_wait_all();
    __syncthreads();
    const int lane = threadIdx.x & 31, nw = (gridDim.x * blockDim.x) >> 5;
    for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < N; row += nw) {
        float acc[32];
        #pragma unroll
        for (int k = 0; k < 32; k++) acc[k] = 0.0f;
        const float *x = X + (int64_t)row * E1;
        // a lane owns the inputs 4*lane .. 4*lane+3 (+128 j): the partition (and with it the bits) of the layer GEMM's fused head epilogue
        for (int e0 = 4 * lane; e0 < E1; e0 += 128) {
            #pragma unroll
            for (int u = 0; u < 4; u++) {
                const int e = e0 + u;
                if (e < E1) {
                    const float xv = x[e];
                    #pragma unroll
                    for (int k = 0; k < 32; k++) if (k < E0) acc[k] = fmaf(xv, sW[k * E1 + e], acc[k]);
                }
            }
        }
        float y = warp_treduce32(acc, lane);           // lane k: Σ_e x[e] W[k][e]
        const bool on = lane < E0;
        if (on) y += __ldg(B + lane);
        const float mx = warp_max(on ? y : -FLT_MAX);  // k_softmax_small (nmath.cu:74-118): exp(x - max) / Σ
        const float ex = on ? __expf(y - mx) : 0.0f;
        const float sm = warp_sum(ex);
        if (on) { const float pv = ex / sm; Y[(int64_t)row * E0 + lane] = y; P[(int64_t)row * E0 + lane] = pv; if (P2) P2[(int64_t)row * E0 + lane] = pv; }
    }
}

// ------------------------------------------------------------------ split-K finish of the hidden linear layer + activation, fused with the head
// forward: row n of  Y1 = Σ_k part_k + b1,  A1 = act(Y1) (+ mask F1)  stays in the registers of the warp that owns the row and feeds
// Y2 = A1 @ W2^T + b2, P = softmax(Y2) directly.  Same arithmetic, same order as k_linear_fin followed by k_head_fwd (bit-equal); one
// launch and one round trip of A1 less on the critical path of the step.  E1 (hidden width) <= 128, E0 (classes) <= 32.
template<int L>
__global__ void __launch_bounds__(T4K_THREADS) k_fin_head_fwd(const float *__restrict__ part, int splits, int64_t MN, const float *__restrict__ B1,
                                                              float *Y1, float *A1, float *F1, float alpha,
                                                              const float *__restrict__ W2, const float *__restrict__ B2, float *Y2, float *P, float *P2,
                                                              int N, int E0, int E1) {
    extern __shared__ float sW[];                      // [E0][E1]
    pdl_wait(); pdl_trigger();
    // W2 arrives by asynchronous copies while the split-K partials of the warp's first row are being summed: the barrier that
    // publishes sW sits behind those loads, so the kernel pays one L2 round trip for both instead of one per fill-loop trip
    for (int t = threadIdx.x; t < E0 * E1; t += blockDim.x) cp_async4(sW + t, W2 + t, true);
    cp_async_commit();
    const int lane = threadIdx.x & 31, nw = (gridDim.x * blockDim.x) >> 5;
    // the lane's (up to) four hidden units 4*lane .. 4*lane+3 (the partition of the layer GEMM's fused head epilogue: same bits); eight
    // splits' loads in flight together; sums in split order as k_linear_fin
    auto sum_parts = [&](int64_t r0, float (&sv)[4]) {
        #pragma unroll
        for (int j = 0; j < 4; j++) { const int e = 4 * lane + j; sv[j] = (e < E1) ? part[r0 + e] : 0.0f; }
        #pragma unroll 8
        for (int k = 1; k < splits; k++) {
            float tv[4];
            #pragma unroll
            for (int j = 0; j < 4; j++) { const int e = 4 * lane + j; tv[j] = (e < E1) ? part[(int64_t)k * MN + r0 + e] : 0.0f; }
            #pragma unroll
            for (int j = 0; j < 4; j++) sv[j] += tv[j];
        }
    };
    int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    float sv[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    bool have = row < N;
    if (have) sum_parts((int64_t)row * E1, sv);
    float b1v[4];                                          // biases: requested with everything else, not after the barrier
    #pragma unroll
    for (int j = 0; j < 4; j++) { const int e = 4 * lane + j; b1v[j] = (e < E1) ? __ldg(B1 + e) : 0.0f; }
    const float b2v = (lane < E0) ? __ldg(B2 + lane) : 0.0f;
    cp_async_wait_all();
    __syncthreads();
    for (; row < N; row += nw) {
        float acc[32];
        #pragma unroll
        for (int k = 0; k < 32; k++) acc[k] = 0.0f;
        const int64_t r0 = (int64_t)row * E1;
        if (!have) sum_parts(r0, sv);
        have = false;
        #pragma unroll
        for (int j = 0; j < 4; j++) {
            const int e = 4 * lane + j;
            if (e < E1) {
                const float s = sv[j] + b1v[j];
                Y1[r0 + e] = s;
                float xv = s;
                if (L != T4K_L_NONE) { float o, f = (L == T4K_L_DROPOUT) ? F1[r0 + e] : 0.0f; act<L>(s, alpha, o, f); A1[r0 + e] = o; F1[r0 + e] = f; xv = o; }
                #pragma unroll
                for (int k = 0; k < 32; k++) if (k < E0) acc[k] = fmaf(xv, sW[k * E1 + e], acc[k]);
            }
        }
        float y = warp_treduce32(acc, lane);
        const bool on = lane < E0;
        if (on) y += b2v;
        const float mx = warp_max(on ? y : -FLT_MAX);
        const float ex = on ? __expf(y - mx) : 0.0f;
        const float sm = warp_sum(ex);
        if (on) { const float pv = ex / sm; Y2[(int64_t)row * E0 + lane] = y; P[(int64_t)row * E0 + lane] = pv; if (P2) P2[(int64_t)row * E0 + lane] = pv; }
    }
}

// ------------------------------------------------------------------ head backward
// per row n:  d = P - T  → P (in place, Model::_bprep) and → Ylin (softmax backward: in = out)
//             dX2[e] = Σ_k d[k] W[k][e]                → X2 row (the small linear's input tensor, in place)
//             dY1[e] = dX2[e] * F1[e]                   → Y1 row (the activation's input tensor)      [if F1]
//  gradients: dW[k][e] += Σ_n d[k] x2[e],  dB[k] += Σ_n d[k],  dB1[e] += Σ_n dY1[e]   (x2 = X2 before overwrite)
// ONE THREAD-BLOCK CLUSTER (HB_CTAS CTAs, a warp per group of rows): per-warp register partials → per-CTA shared-memory
// partial → cluster barrier → CTA r adds slice r of the HB_CTAS partials through DISTRIBUTED SHARED MEMORY in CTA order
// and applies it.  No global partials, no atomics, no second launch; deterministic.
#define HB_CTAS 16                                     // non-portable cluster size (opt-in attribute), 16 x 8 warps
#define HB_ROWS 4                                      // rows per warp per trip (all their loads in flight together)
struct HeadB {
    float *P; const float *T; float *Ylin, *X2; const float *F1; float *Y1; const float *W;
    float *dW, *dB, *dB1;
    int N, E0, E1, train;
};
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float ld_dsmem(const float *local, uint32_t rank) {       // the same smem offset in CTA `rank` of the cluster
    uint32_t a = (uint32_t)__cvta_generic_to_shared(local), r; float v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(r) : "memory");
    return v;
}
template<int KM>                                       // KM = compile-time bound on E0 (8, 16 or 32); E1 <= 128
__global__ void __cluster_dims__(HB_CTAS, 1, 1) __launch_bounds__(T4K_THREADS) k_head_bwd(HeadB p) {
    extern __shared__ float sm[];
    pdl_wait(); pdl_trigger();                  // PDL: nothing global before this line
    const int E0 = p.E0, E1 = p.E1, nE = E0 * E1 + E0 + E1;
    float *sW = sm;                                    // [E0][128]       rows zero-padded: no column guards in the hot loop
    float *sPart = sm + E0 * 128;                      // [nE]            this CTA's partial (read by the whole cluster)
    float *sAcc = sPart + ((nE + 3) & ~3);             // [nwarps][nE]    per-warp partials
    // W by asynchronous copies (zero-filled padding), the warp's first rows requested before the barrier that publishes it:
    // one global round trip for both (a load -> store fill loop pays one L2 round trip per trip, five of them here)
    for (int t = threadIdx.x; t < E0 * 128; t += blockDim.x) { const int e = t & 127, k = t >> 7; cp_async4(sW + t, p.W + ((e < E1) ? k * E1 + e : 0), e < E1); }
    cp_async_commit();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int nw = gridDim.x * nwarps;
    float accW[KM][4], accB1[4], accB = 0.0f;
    #pragma unroll
    for (int k = 0; k < KM; k++) { accW[k][0] = accW[k][1] = accW[k][2] = accW[k][3] = 0.0f; }
    accB1[0] = accB1[1] = accB1[2] = accB1[3] = 0.0f;
    float d[HB_ROWS], x[HB_ROWS][4], f[HB_ROWS][4];
    auto load_rows = [&](int row0) {                                               // rows past N read nothing (zeros)
        #pragma unroll
        for (int r = 0; r < HB_ROWS; r++) {
            const int row = row0 + r;
            const bool rv = row < p.N;
            d[r] = 0.0f;
            if (rv && lane < E0) { const int64_t o = (int64_t)row * E0 + lane; d[r] = __fsub_rn(p.P[o], p.T[o]); }
            #pragma unroll
            for (int j = 0; j < 4; j++) {
                const int e = lane + 32 * j;
                const bool ev = rv && e < E1;
                x[r][j] = ev ? p.X2[(int64_t)row * E1 + e] : 0.0f;
                f[r][j] = (ev && p.F1) ? p.F1[(int64_t)row * E1 + e] : 1.0f;
            }
        }
    };
    int row0 = (blockIdx.x * nwarps + warp) * HB_ROWS;
    bool have = true;
    load_rows(row0);
    cp_async_wait_all();
    __syncthreads();
    for (; row0 < p.N; row0 += nw * HB_ROWS) {
        if (!have) load_rows(row0);
        have = false;
        #pragma unroll
        for (int r = 0; r < HB_ROWS; r++) {
            const int row = row0 + r;
            if (row >= p.N) break;
            if (lane < E0) { const int64_t o = (int64_t)row * E0 + lane; p.P[o] = d[r]; p.Ylin[o] = d[r]; accB += d[r]; }
            float dx[4] = {0.0f, 0.0f, 0.0f, 0.0f};
            #pragma unroll
            for (int k = 0; k < KM; k++) {
                if (k < E0) {
                    const float dk = __shfl_sync(0xffffffffu, d[r], k);
                    #pragma unroll
                    for (int j = 0; j < 4; j++) {
                        dx[j] = fmaf(dk, sW[k * 128 + lane + 32 * j], dx[j]);
                        accW[k][j] = fmaf(dk, x[r][j], accW[k][j]);
                    }
                }
            }
            #pragma unroll
            for (int j = 0; j < 4; j++) {
                const int e = lane + 32 * j;
                if (e < E1) {
                    const int64_t o = (int64_t)row * E1 + e;
                    p.X2[o] = dx[j];
                    float g = dx[j];
                    if (p.F1) { g = __fmul_rn(dx[j], f[r][j]); p.Y1[o] = g; }
                    accB1[j] += g;
                }
            }
        }
    }
    if (!p.train) return;                                                          // uniform over the cluster
    float *mine = sAcc + (size_t)warp * nE;
    #pragma unroll
    for (int k = 0; k < KM; k++) if (k < E0) {
        #pragma unroll
        for (int j = 0; j < 4; j++) { const int e = lane + 32 * j; if (e < E1) mine[k * E1 + e] = accW[k][j]; }
    }
    if (lane < E0) mine[E0 * E1 + lane] = accB;
    #pragma unroll
    for (int j = 0; j < 4; j++) { const int e = lane + 32 * j; if (e < E1) mine[E0 * E1 + E0 + e] = accB1[j]; }
    __syncthreads();
    for (int t = threadIdx.x; t < nE; t += blockDim.x) {
        float s = 0.0f;
        for (int w = 0; w < nwarps; w++) s += sAcc[(size_t)w * nE + t];
        sPart[t] = s;
    }
    cluster_sync_all();                                                            // every CTA's sPart is complete and visible
    const int rank = (int)cluster_rank();
    const int chunk = (nE + HB_CTAS - 1) / HB_CTAS;
    for (int t = rank * chunk + threadIdx.x; t < min(nE, (rank + 1) * chunk); t += blockDim.x) {
        float v[HB_CTAS];
        #pragma unroll
        for (int c = 0; c < HB_CTAS; c++) v[c] = ld_dsmem(sPart + t, (uint32_t)c);
        float s = 0.0f;
        #pragma unroll
        for (int c = 0; c < HB_CTAS; c++) s += v[c];
        if (t < E0 * E1) p.dW[t] += s;
        else if (t < E0 * E1 + E0) p.dB[t - E0 * E1] += s;
        else if (p.dB1) p.dB1[t - E0 * E1 - E0] += s;
    }
    cluster_sync_all();                                                            // nobody leaves while its smem is still being read
}

// dW2 | dB2 | dB1 += Σ_cta partial[cta] (CTA order: deterministic); layout of a partial: [E0][EH] | [E0 padded to 4] | [EH padded to 4]
__global__ void __launch_bounds__(T4K_THREADS) k_head_grad_fin(const float *__restrict__ part, int ncta, int E0, int EH, float *dW2, float *dB2, float *dB1) {
    pdl_wait(); pdl_trigger();
    const int E0p = (E0 + 3) & ~3, nEp = E0 * EH + E0p + ((EH + 3) & ~3);
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nEp) return;
    float s_ = 0.0f;
    for (int c = 0; c < ncta; c++) s_ += part[(size_t)c * nEp + t];
    if (t < E0 * EH) dW2[t] += s_;
    else if (t < E0 * EH + E0p) { if (t - E0 * EH < E0) dB2[t - E0 * EH] += s_; }
    else if (t - E0 * EH - E0p < EH && dB1) dB1[t - E0 * EH - E0p] += s_;
}

} // namespace t4k
using namespace t4k;

// dX and dW of a linear layer in one launch of the layer GEMM (T4K_TL_PAIR=0: two launches, as before)
static int g_tl_pair = []{ const char *e = getenv("T4K_TL_PAIR"); return (e && e[0] == '0') ? 0 : 1; }();
static bool head_train_ok(int layer, int N, int EH, int E1, int E0) {
    return (layer == T4K_L_RELU || layer == T4K_L_TANH || layer == T4K_L_SELU || layer == T4K_L_LEAKYRL || layer == T4K_L_ELU) &&
           E0 >= 1 && E0 <= 32 && EH <= 128 && (EH & 3) == 0 && (E1 & 3) == 0 && N >= 32;
}
extern "C" int64_t t4k_head_train_scratch_floats(int layer, int N, int EH, int E1, int E0) {
    if (!head_train_ok(layer, N, EH, E1, E0)) return 0;
    TlEpi e{}; e.mode = 4; e.E2 = E0; e.T = (const float*)16; e.Ylin = (float*)16; e.hpart = (float*)16; e.layer = layer;
    const TlJob j{(const float*)16, (const float*)16, (float*)16, 1.0f, 0.0f, 0, 1, N, EH, E1, &e};
    if ((double)N * EH * E1 < 4.0e6) return 0;
    const int ctas = gemm_tl_ctas(&j, 1);
    if (ctas <= 0) return 0;
    const int E0p = (E0 + 3) & ~3;
    return (int64_t)ctas * (E0 * EH + E0p + ((EH + 3) & ~3));
}
extern "C" int t4k_linear_act_head_train(int layer, const float *X, const float *W1, const float *B1, float *Y1, float *A1, float *F1, float alpha,
                                         const float *W2, const float *B2, float *Ylin, float *P, float *Pdup, const float *T,
                                         float *scratch, int *ncta, int N, int EH, int E1, int E0, t4k_stream_t s) {
    if (!X || !W1 || !B1 || !Y1 || !A1 || !F1 || !W2 || !B2 || !Ylin || !P || !T || !scratch || !ncta) return T4K_EINVAL;
    if (!head_train_ok(layer, N, EH, E1, E0) || !gemm_tl_ok(X, W1, Y1, 0, 1, N, EH, E1, 1, 1) || !aligned16(W2) || !aligned16(F1) || !aligned16(A1) || !aligned16(Y1) || !aligned16(scratch))
        return T4K_ENOSUP;
    TlEpi e{}; e.mode = 4; e.bias = B1; e.actA = A1; e.actF = F1; e.layer = layer; e.act_alpha = alpha;
    e.W2 = W2; e.B2 = B2; e.Y2 = Ylin; e.P = P; e.P2 = Pdup; e.E2 = E0; e.T = T; e.Ylin = Ylin; e.hpart = scratch;
    const TlJob j{X, W1, Y1, 1.0f, 0.0f, 0, 1, N, EH, E1, &e};
    return gemm_tl_multi(&j, 1, STRM(s), ncta);
}
extern "C" int t4k_head_grad_finish(const float *scratch, int ncta, int E0, int EH, float *dW2, float *dB2, float *dB1, t4k_stream_t s) {
    if (!scratch || ncta < 1 || E0 < 1 || EH < 1 || !dW2 || !dB2) return T4K_EINVAL;
    const int E0p = (E0 + 3) & ~3, nEp = E0 * EH + E0p + ((EH + 3) & ~3);
    launch_pdl(k_head_grad_fin, dim3((nEp + T4K_THREADS - 1) / T4K_THREADS), dim3(T4K_THREADS), 0, STRM(s), scratch, ncta, E0, EH, dW2, dB2, dB1);
    return check_launch();
}
extern "C" int t4k_linear_bwd_pair(const float *X, const float *W, const float *dY, float *dX, float *dW, int N, int E0, int E1, t4k_stream_t s) {
    if (!X || !W || !dY || !dX || !dW || N < 1 || E0 < 1 || E1 < 1 || X == dX) return T4K_EINVAL;
    if (!gemm_tl_ok(dY, W, dX, 0, 0, N, E1, E0, 1, 1) || !gemm_tl_ok(dY, X, dW, 1, 0, E0, E1, N, 1, 1)) return T4K_ENOSUP;
    const TlJob jobs[2] = {{dY, W, dX, 1.0f, 0.0f, 0, 0, N, E1, E0, nullptr},           // dX[N,E1]  = dY[N,E0] @ W[E0,E1]
                           {dY, X, dW, 1.0f, 1.0f, 1, 0, E0, E1, N, nullptr}};          // dW[E0,E1] += dY^T @ X[N,E1]
    return gemm_tl_multi(jobs, 2, STRM(s));
}

extern "C" int t4k_linear_fwd(const float *X, const float *W, const float *B, float *Y, int N, int E0, int E1, t4k_stream_t s) {
    if (!X || !W || !B || !Y || N < 1 || E0 < 1 || E1 < 1) return T4K_EINVAL;
    if ((double)N * E0 * E1 >= 2.0e10 && N >= 64 && E0 >= 32 && E1 >= 64) {        // very large: packed-plane tensor-core engine + bias pass
        int rc = t4k_gemm(X, W, Y, 1.0f, 0.0f, 0, 1, N, E0, E1, 1, 1, 0, 0, 0, s);
        if (rc) return rc;
        return t4k_bias(B, Y, N, E0, s);
    }
    return linear_act_fwd(T4K_L_NONE, X, W, B, Y, nullptr, nullptr, 0.0f, N, E0, E1, STRM(s));
}
extern "C" int t4k_linear_act_fwd(int layer, const float *X, const float *W, const float *B, float *Y, float *A, float *F, float alpha,
                                  int N, int E0, int E1, t4k_stream_t s) {
    if (!X || !W || !B || !Y || !A || !F || N < 1 || E0 < 1 || E1 < 1) return T4K_EINVAL;
    if (layer < T4K_L_RELU || layer > T4K_L_DROPOUT) return T4K_EINVAL;
    if ((double)N * E0 * E1 >= 2.0e8 && N >= 64 && E0 >= 32 && E1 >= 64) {
        int rc = t4k_gemm(X, W, Y, 1.0f, 0.0f, 0, 1, N, E0, E1, 1, 1, 0, 0, 0, s);
        if (rc) return rc;
        GemmDeferred d{Y, 1};                                                       // bias + activation in one pass over Y
        const int64_t MN = (int64_t)N * E0;
        switch (layer) {
        case T4K_L_RELU:    return launch_fin<T4K_L_RELU>(d, MN, E0, B, Y, A, F, alpha, STRM(s));
        case T4K_L_TANH:    return launch_fin<T4K_L_TANH>(d, MN, E0, B, Y, A, F, alpha, STRM(s));
        case T4K_L_SIGMOID: return launch_fin<T4K_L_SIGMOID>(d, MN, E0, B, Y, A, F, alpha, STRM(s));
        case T4K_L_SELU:    return launch_fin<T4K_L_SELU>(d, MN, E0, B, Y, A, F, alpha, STRM(s));
        case T4K_L_LEAKYRL: return launch_fin<T4K_L_LEAKYRL>(d, MN, E0, B, Y, A, F, alpha, STRM(s));
        case T4K_L_ELU:     return launch_fin<T4K_L_ELU>(d, MN, E0, B, Y, A, F, alpha, STRM(s));
        default:            return launch_fin<T4K_L_DROPOUT>(d, MN, E0, B, Y, A, F, alpha, STRM(s));
        }
    }
    return linear_act_fwd(layer, X, W, B, Y, A, F, alpha, N, E0, E1, STRM(s));
}
extern "C" int t4k_linear_bwd_ex(const float *X, const float *W, const float *dY, float *dX, float *dW, float *dB,
                                 int N, int E0, int E1, int train, int skip_db, t4k_stream_t s) {
    if (!X || !W || !dY || !dX || N < 1 || E0 < 1 || E1 < 1) return T4K_EINVAL;
    if (train) {
        if (!dW || (!dB && !skip_db)) return T4K_EINVAL;
        int rc = 0;
        if (!skip_db) { rc = t4k_dbias(dY, dB, N, E0, s); if (rc) return rc; }       // dB[E0] += Σ_n dY
        if (g_tl_pair && gemm_tl_ok(dY, W, dX, 0, 0, N, E1, E0, 1, 1) && gemm_tl_ok(dY, X, dW, 1, 0, E0, E1, N, 1, 1)) {
            // dX and dW in ONE launch of the layer GEMM, also when dX is stored over X (Model::_blinear works in place): the dX tiles are held
            // back until every CTA of the dW problem has its last X tile in shared memory
            const TlJob jobs[2] = {{dY, W, dX, 1.0f, 0.0f, 0, 0, N, E1, E0, nullptr}, {dY, X, dW, 1.0f, 1.0f, 1, 0, E0, E1, N, nullptr}};
            rc = (X == dX) ? gemm_tl_pair_inplace(jobs, STRM(s)) : gemm_tl_multi(jobs, 2, STRM(s));
            if (rc != T4K_ENOSUP) return rc;
        }
        rc = t4k_gemm(dY, X, dW, 1.0f, 1.0f, 1, 0, E0, E1, N, 1, 1, 0, 0, 0, s);      // dW[E0,E1] += dY^T[E0,N] @ X[N,E1]
        if (rc) return rc;
    }
    return t4k_gemm(dY, W, dX, 1.0f, 0.0f, 0, 0, N, E1, E0, 1, 1, 0, 0, 0, s);       // dX[N,E1] = dY[N,E0] @ W[E0,E1]
}
extern "C" int t4k_linear_bwd_act(const float *X, const float *W, const float *dY, float *dX, float *dW, float *dB,
                                  const float *Fprev, float *dXprev, int N, int E0, int E1, int train, int skip_db, t4k_stream_t s) {
    if (!X || !W || !dY || !dX || !Fprev || !dXprev || N < 1 || E0 < 1 || E1 < 1) return T4K_EINVAL;
    if (!gemm_tl_ok(dY, W, dX, 0, 0, N, E1, E0, 1, 1)) {                            // per-layer path: _blinear, then _bactivate
        int rc = t4k_linear_bwd_ex(X, W, dY, dX, dW, dB, N, E0, E1, train, skip_db, s);
        if (rc) return rc;
        return t4k_activate_bwd(dX, Fprev, dXprev, (int64_t)N * E1, s);
    }
    TlEpi e{}; e.mode = 3; e.F = Fprev; e.O2 = dXprev;
    if (train) {
        if (!dW || (!dB && !skip_db)) return T4K_EINVAL;
        int rc = 0;
        if (!skip_db) { rc = t4k_dbias(dY, dB, N, E0, s); if (rc) return rc; }
        if (g_tl_pair && gemm_tl_ok(dY, X, dW, 1, 0, E0, E1, N, 1, 1)) {               // dX (+ the activation backward in its epilogue) and dW in one launch, see t4k_linear_bwd_ex
            const TlJob jobs[2] = {{dY, W, dX, 1.0f, 0.0f, 0, 0, N, E1, E0, &e}, {dY, X, dW, 1.0f, 1.0f, 1, 0, E0, E1, N, nullptr}};
            rc = (X == dX || X == dXprev) ? gemm_tl_pair_inplace(jobs, STRM(s)) : gemm_tl_multi(jobs, 2, STRM(s));
            if (rc != T4K_ENOSUP) return rc;
        }
        rc = t4k_gemm(dY, X, dW, 1.0f, 1.0f, 1, 0, E0, E1, N, 1, 1, 0, 0, 0, s);
        if (rc) return rc;
    }
    return gemm_tl(dY, W, dX, 1.0f, 0.0f, 0, 0, N, E1, E0, STRM(s), &e);            // dX = dY @ W ; dXprev = dX * Fprev
}
extern "C" int t4k_linear_dx_from_head(const float *P, const float *T, const float *W2, const float *F1, const float *W1, float *dX,
                                       int N, int E2, int EH, int E1, t4k_stream_t s) {
    if (!P || !T || !W2 || !W1 || !dX || N < 1 || E2 < 1 || EH < 1 || E1 < 1) return T4K_EINVAL;
    if (E2 > 32 || EH > 128 || (EH & 3) || !gemm_tl_ok(W1, W1, dX, 0, 0, N, E1, EH, 1, 1) || !aligned16(W2) || (F1 && !aligned16(F1))) return T4K_ENOSUP;
    TlEpi e{}; e.mode = 0; e.gP = P; e.gT = T; e.gW2 = W2; e.gF = F1; e.gE2 = E2;
    return gemm_tl(nullptr, W1, dX, 1.0f, 0.0f, 0, 0, N, E1, EH, STRM(s), &e);      // dX[N,E1] = A[N,EH] @ W1[EH,E1], A generated
}
extern "C" int t4k_linear_bwd_from_head(const float *P, const float *T, const float *W2, const float *F1, const float *X, const float *W1,
                                        float *dX, float *dW1, int N, int E2, int EH, int E1, t4k_stream_t s) {
    if (!P || !T || !W2 || !X || !W1 || !dX || !dW1 || N < 1 || E2 < 1 || EH < 1 || E1 < 1 || X == dX) return T4K_EINVAL;
    if (E2 > 32 || EH > 128 || (EH & 3) || !aligned16(W2) || (F1 && !aligned16(F1)) ||
        !gemm_tl_ok(W1, W1, dX, 0, 0, N, E1, EH, 1, 1) || !gemm_tl_ok(X, X, dW1, 1, 0, EH, E1, N, 1, 1)) return T4K_ENOSUP;
    TlEpi e{}; e.mode = 0; e.gP = P; e.gT = T; e.gW2 = W2; e.gF = F1; e.gE2 = E2;
    const TlJob jobs[2] = {{nullptr, W1, dX, 1.0f, 0.0f, 0, 0, N, E1, EH, &e},          // dX[N,E1]   = dY1[N,EH] @ W1[EH,E1]
                           {nullptr, X, dW1, 1.0f, 1.0f, 1, 0, EH, E1, N, &e}};         // dW1[EH,E1] += dY1^T @ X[N,E1]
    return gemm_tl_multi(jobs, 2, STRM(s));
}
extern "C" int t4k_linear_bwd(const float *X, const float *W, const float *dY, float *dX, float *dW, float *dB,
                              int N, int E0, int E1, int train, t4k_stream_t s) {
    return t4k_linear_bwd_ex(X, W, dY, dX, dW, dB, N, E0, E1, train, 0, s);
}

extern "C" int t4k_mlp_head_fwd(const float *X, const float *W, const float *B, float *Y, float *P, int N, int E0, int E1, t4k_stream_t s) {
    return t4k_mlp_head_fwd_dup(X, W, B, Y, P, nullptr, N, E0, E1, s);
}
extern "C" int t4k_mlp_head_fwd_dup(const float *X, const float *W, const float *B, float *Y, float *P, float *Pdup, int N, int E0, int E1, t4k_stream_t s) {
    if (!X || !W || !B || !Y || !P || N < 1 || E0 < 1 || E1 < 1) return T4K_EINVAL;
    if (E0 > 32 || (size_t)E0 * E1 * sizeof(float) > 40 * 1024) return T4K_ENOSUP;
    const int rows_per_cta = T4K_THREADS / 32;
    int g = (N + rows_per_cta - 1) / rows_per_cta;
    if (g > 2 * sm_count()) g = 2 * sm_count();
    launch_pdl(k_head_fwd, dim3(g), dim3(T4K_THREADS), (size_t)E0 * E1 * sizeof(float), STRM(s), X, W, B, Y, P, Pdup, N, E0, E1);
    return check_launch();
}
template<int L> static int launch_fin_head(const GemmDeferred &d, int64_t MN, const float *B1, float *Y1, float *A1, float *F1, float alpha,
                                           const float *W2, const float *B2, float *Y2, float *P, float *Pdup, int N, int E0, int E1, cudaStream_t st) {
    const int rows_per_cta = T4K_THREADS / 32;
    int g = (N + rows_per_cta - 1) / rows_per_cta;
    if (g > 2 * sm_count()) g = 2 * sm_count();
    launch_pdl(k_fin_head_fwd<L>, dim3(g), dim3(T4K_THREADS), (size_t)E0 * E1 * sizeof(float), st,
               d.part, d.splits, MN, B1, Y1, A1, F1, alpha, W2, B2, Y2, P, Pdup, N, E0, E1);
    return check_launch();
}
extern "C" int t4k_linear_act_head_fwd(int layer, const float *X, const float *W1, const float *B1, float *Y1, float *A1, float *F1, float alpha,
                                       const float *W2, const float *B2, float *Y2, float *P, float *Pdup,
                                       int N, int EH, int E1, int E0, t4k_stream_t s) {
    if (!X || !W1 || !B1 || !Y1 || !A1 || !F1 || !W2 || !B2 || !Y2 || !P || N < 1 || EH < 1 || E1 < 1 || E0 < 1) return T4K_EINVAL;
    if (E0 > 32 || EH > 128 || (size_t)E0 * EH * sizeof(float) > 40 * 1024) return T4K_ENOSUP;
    cudaStream_t st = STRM(s);
    if (layer != T4K_L_DROPOUT && (EH & 3) == 0 && gemm_tl_ok(X, W1, Y1, 0, 1, N, EH, E1, 1, 1)) {
        // ONE launch: layer GEMM, and in its cluster epilogue bias + activation + the small linear + softmax on the finished rows
        TlEpi e{}; e.mode = 2; e.bias = B1; e.actA = A1; e.actF = F1; e.layer = layer; e.act_alpha = alpha;
        e.W2 = W2; e.B2 = B2; e.Y2 = Y2; e.P = P; e.P2 = Pdup; e.E2 = E0;
        int rc = gemm_tl(X, W1, Y1, 1.0f, 0.0f, 0, 1, N, EH, E1, st, &e);
        if (rc != T4K_ENOSUP) return rc;
    }
    GemmDeferred d{nullptr, 1};
    int rc = gemm_tcf_ok(0, 1, N, EH, E1, 1, 1) ? gemm_tcf(X, W1, Y1, 1.0f, 0.0f, 0, 1, N, EH, E1, st, &d)
                                                : gemm_simt(X, W1, Y1, 1.0f, 0.0f, 0, 1, N, EH, E1, 1, 1, 0, 0, 0, st, &d);
    if (rc) return rc;
    const int64_t MN = (int64_t)N * EH;
    switch (layer) {
    case T4K_L_RELU:    return launch_fin_head<T4K_L_RELU>(d, MN, B1, Y1, A1, F1, alpha, W2, B2, Y2, P, Pdup, N, E0, EH, st);
    case T4K_L_TANH:    return launch_fin_head<T4K_L_TANH>(d, MN, B1, Y1, A1, F1, alpha, W2, B2, Y2, P, Pdup, N, E0, EH, st);
    case T4K_L_SIGMOID: return launch_fin_head<T4K_L_SIGMOID>(d, MN, B1, Y1, A1, F1, alpha, W2, B2, Y2, P, Pdup, N, E0, EH, st);
    case T4K_L_SELU:    return launch_fin_head<T4K_L_SELU>(d, MN, B1, Y1, A1, F1, alpha, W2, B2, Y2, P, Pdup, N, E0, EH, st);
    case T4K_L_LEAKYRL: return launch_fin_head<T4K_L_LEAKYRL>(d, MN, B1, Y1, A1, F1, alpha, W2, B2, Y2, P, Pdup, N, E0, EH, st);
    case T4K_L_ELU:     return launch_fin_head<T4K_L_ELU>(d, MN, B1, Y1, A1, F1, alpha, W2, B2, Y2, P, Pdup, N, E0, EH, st);
    default:            return T4K_ENOSUP;
    }
}
extern "C" int t4k_mlp_head_bwd(float *P, const float *T, float *Ylin, float *X2, const float *F1, float *Y1, const float *W,
                                float *dW, float *dB, float *dB1, int N, int E0, int E1, int train, t4k_stream_t s) {
    if (!P || !T || !Ylin || !X2 || !W || N < 1 || E0 < 1 || E1 < 1 || (F1 && !Y1) || (train && (!dW || !dB))) return T4K_EINVAL;
    if (E0 > 32 || E1 > 128) return T4K_ENOSUP;
    const int nwarps = T4K_THREADS / 32, nE = E0 * E1 + E0 + E1;
    const size_t smem = ((size_t)E0 * 128 + ((nE + 3) & ~3) + (size_t)nwarps * nE) * sizeof(float);
    if (smem > 96 * 1024) return T4K_ENOSUP;
    HeadB p{P, T, Ylin, X2, F1, Y1, W, dW, dB, dB1, N, E0, E1, train};
    #define HEADB(KM_) { static DevFlag attr; if (dev_first(attr)) { cudaFuncSetAttribute(k_head_bwd<KM_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024); \
                                                                 cudaFuncSetAttribute(k_head_bwd<KM_>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1); } \
                         launch_pdl(k_head_bwd<KM_>, dim3(HB_CTAS), dim3(T4K_THREADS), smem, STRM(s), p); }
    if (E0 <= 8) HEADB(8) else if (E0 <= 16) HEADB(16) else HEADB(32)
    return check_launch();
}
