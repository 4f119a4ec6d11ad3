// cnn_block.cu — the fused "conv2d → maxpool(2) → relu (→ flatten)" layer group, second generation.
//   forward : Model::_fconv + _fpool + _factivate (+ flatten copy) (+ the `n0 = input` copy of Model::forward)
//             src/nn/forward.cu:29-43,83-155,201-228 ; kernels k_conv2d / k_pool / k_activate / k_copy
//   backward: flatten copy + _bactivate + _bpool + _bconv            src/nn/backprop.cu:112-191,257-280
// Every tensor the per-layer path writes is still written (n@ / nn.dw show the same values); what changes is
// that each is touched ONCE: HBM-bound streaming kernels, roofline = algorithmic bytes / HBM bandwidth.
//
// Mapping (stride-1 "same" conv, C1 <= 4, C0 <= 16, even H0/W0): one CTA per sample, one THREAD PER 2x2 POOL
// WINDOW.  A thread owns the window's 4 conv pixels x C0 channels in registers, so
//   forward : conv (FMA order of k_conv_fwd_small: bias, then ky,kx,c1 — bit-equal), max-pool and relu never leave
//             registers; the conv output is written as 2 x (2*C0) contiguous floats per thread (128-bit stores,
//             neighbouring threads neighbouring addresses), pooled tensors as 64-bit stores;
//   backward: (C1 == 1) 128-thread CTAs, TWO windows per thread (4 CTAs/SM: a 512-sample batch is one wave).  Every global read of
//             the CTA is requested before its first barrier: dY / input / taps as asynchronous copies into shared memory, the relu
//             mask and the windows' forward conv outputs (128-bit loads) into registers.  The conv outputs redo the
//             arg-max routing (first strict max in y,x order, nmath.tcu:535-549), the routed gradient goes back
//             to HBM from registers and into a zero-haloed channel-major smem tile for the dX gather;
//             dF/dB: per-thread partials in channel passes (9 taps x 5 channels + their dB sums per pass: every routed value is
//             read once for all taps), reduced across the warp with a
//             transposing butterfly (31 shuffles per 32 values), across warps in smem, per-sample partials are
//             summed in sample order by k_wgrad_fin (deterministic; the reference uses atomics, nmath.tcu:307-336).
//   forward with dataset feed (FEED): the sample's pixels come from the staged U8 block, are normalised in registers and written to
//             the dataset tensor, the model's input layer and the conv tile; the CTA also writes the sample's label and one-hot row.
// Shapes outside this envelope fall back to the first-generation kernels in conv.cu (same results).
#include "common.cuh"
#include <cstdlib>

namespace t4k {

int cpr_v1_fwd(const float *I, const float *F, const float *B, float *convO, float *poolO, float *actO, float *actF,
               float *flatO, int N, int H1, int W1, int C1, int H0, int W0, int C0, int KS, int S, int P, cudaStream_t st);
int cpr_v1_bwd(const float *dY, float *actO, const float *actF, float *poolO, float *convO, float *Iio, float *dXbuf,
               const float *F, float *dF, float *dB, int N, int H1, int W1, int C1, int H0, int W0, int C0,
               int KS, int S, int P, int train, cudaStream_t st);
int wgrad_fin_launch(const float *part, float *dF, float *dB, int nF, int C0, int nparts, int KS, int S, cudaStream_t st);
int wgrad_fin_opt_launch(const float *part, float *dF, float *dB, int nF, int C0, int nparts, int KS, int S, const t4k_fused_opt_t *opt, cudaStream_t st);

#ifndef CPR2_FWD_MINB
#define CPR2_FWD_MINB 4        // 64 registers, no spills: 4 x 224-thread CTAs per SM = 592 slots, so the 512 samples of the MNIST batch are ONE wave (3 per SM: 1.15 waves)
#endif
struct Cpr2P {
    const float *I, *F, *B, *dY, *actFc;
    float *Icopy, *convO, *poolO, *actO, *actF, *flatO, *Iio, *dXbuf, *part;
    int H, W, C1, C0, train, zero_all;
    // dataset feed folded into the forward block (t4k_conv_pool_relu_fwd_feed): samples n < feedN take their pixels from the staged
    // U8 block — d = ((float)u8 - mean) * scale as t4k_dataset_load — and write them to the dataset tensor (I) as well
    const uint8_t *u8I, *u8L; float mean, scale; int32_t *lab32; float *hot; int E, feedN;
};

// ------------------------------------------------------------------ forward
// C0T = exact channel count known at compile time (static register indexing for the 128-bit stores), 0 = runtime C0 <= CP
// Global traffic is kept at full-sector granularity: the sample's input arrives as 128-bit loads, the conv pixels
// leave the registers as 128-bit stores of contiguous 2*C0 runs, the pooled values are staged in smem and the four
// pooled tensors (pool, relu, mask, flatten) leave as contiguous 128-bit streams.
template<int KS, int CP, int C0T, int C1T, bool FEED = false>
__global__ void __launch_bounds__(256, CPR2_FWD_MINB) k_cpr2_fwd(Cpr2P p) {
    extern __shared__ __align__(16) float sm[];
    pdl_wait(); pdl_trigger();                  // PDL: nothing global before this line
    constexpr int P = (KS - 1) / 2;
    const int H = p.H, W = p.W, C1 = C1T ? C1T : p.C1, C0 = C0T ? C0T : p.C0;
    const int WP = W + 2 * P, HP = H + 2 * P;
    const int nFp = KS * KS * C1 * CP;
    const int Hp = H / 2, Wp = W / 2, nwin = Hp * Wp, nP = nwin * C0;
    float *sF = sm;                         // [(ky*KS+kx)*C1 + c1][CP]   zero padded channels
    float *sB = sF + nFp;                   // [CP]
    float *sP = sB + CP;                    // [nwin][C0] pooled values (staging), 16-byte aligned
    float *sI = sP + ((nP + 3) & ~3);       // [HP][WP][C1]               zero halo
    const int n = blockIdx.x;
    const int nI = H * W * C1;
    const float *gI = p.I + (int64_t)n * nI;
    // Order of issue (one global round trip instead of three): the sample's input first (HBM, into a register), then taps and bias
    // as asynchronous copies straight into shared memory, then the halo zeros; the input is consumed last.
    const int rowp = WP * C1;
    const int rowf = W * C1;
    const bool vecI = (rowf & 3) == 0 && aligned16(p.I) && (!p.Icopy || aligned16(p.Icopy));
    const int rq = rowf >> 2;
    float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool feed = FEED && n < p.feedN;                // host-checked for FEED: vecI, nI % 16 == 0, 16-byte aligned blocks, E <= blockDim
    const bool has0 = vecI && !feed && (int)threadIdx.x < H * rq;
    if (has0) { const int y = threadIdx.x / rq, q = threadIdx.x - y * rq; v0 = ldg4(gI + y * rowf + 4 * q); }
    uint4 u0 = make_uint4(0u, 0u, 0u, 0u);
    int lab = 0;
    const uint8_t *gU = FEED ? p.u8I + (int64_t)n * nI : nullptr;
    if (FEED && feed) {
        if ((int)threadIdx.x < (nI >> 4)) u0 = __ldg(reinterpret_cast<const uint4*>(gU) + threadIdx.x);
        if ((int)threadIdx.x < p.E) lab = (int)__ldg(p.u8L + n);
    }
    for (int t = threadIdx.x; t < nFp; t += blockDim.x) {
        const int c = t % CP; int r = t / CP; const int c1 = r % C1; r /= C1;        // r = ky*KS+kx
        cp_async4(sF + t, p.F + ((c < C0) ? ((int64_t)c1 * KS * KS + r) * C0 + c : 0), c < C0);
    }
    for (int t = threadIdx.x; t < CP; t += blockDim.x) cp_async4(sB + t, p.B + ((t < C0) ? t : 0), t < C0);
    cp_async_commit();
    // halo zeros (top/bottom rows, left/right columns), then the interior from 128-bit loads (else scalar)
    for (int t = threadIdx.x; t < P * rowp; t += blockDim.x) { sI[t] = 0.0f; sI[(HP - P) * rowp + t] = 0.0f; }
    for (int t = threadIdx.x; t < H * P * C1; t += blockDim.x) {
        const int y = t / (P * C1), q = t - y * (P * C1);
        sI[(y + P) * rowp + q] = 0.0f; sI[(y + P) * rowp + (W + P) * C1 + q] = 0.0f;
    }
    if (FEED && feed) {
        // Dataset::_load on the fly (src/mu/dataset.cu:139-152): 16 pixels per 128-bit load; the normalised pixels go to the dataset
        // tensor, to the model's input layer (n0 = input) and into the conv tile; labels widen to int32 and to their one-hot row
        float *gD = const_cast<float*>(p.I) + (int64_t)n * nI;
        auto putu = [&](int t, const uint4 w) {
            const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
            #pragma unroll
            for (int k = 0; k < 4; k++) {
                float4 o;
                o.x = __fmul_rn(__fsub_rn((float)(int)(ww[k] & 0xffu), p.mean), p.scale);
                o.y = __fmul_rn(__fsub_rn((float)(int)((ww[k] >> 8) & 0xffu), p.mean), p.scale);
                o.z = __fmul_rn(__fsub_rn((float)(int)((ww[k] >> 16) & 0xffu), p.mean), p.scale);
                o.w = __fmul_rn(__fsub_rn((float)(int)(ww[k] >> 24), p.mean), p.scale);
                const int idx = 16 * t + 4 * k, y = idx / rowf, q = idx - y * rowf;       // rowf % 4 == 0: the quad stays in one row
                stg4(gD + idx, o);
                stg4(p.Icopy + (int64_t)n * nI + idx, o);
                float *d = sI + ((y + P) * WP + P) * C1 + q;
                d[0] = o.x; d[1] = o.y; d[2] = o.z; d[3] = o.w;
            }
        };
        if ((int)threadIdx.x < (nI >> 4)) putu(threadIdx.x, u0);
        for (int t = threadIdx.x + blockDim.x; t < (nI >> 4); t += blockDim.x) putu(t, __ldg(reinterpret_cast<const uint4*>(gU) + t));
        if ((int)threadIdx.x < p.E) p.hot[(int64_t)n * p.E + threadIdx.x] = ((int)threadIdx.x == (lab < p.E ? lab : 0)) ? 1.0f : 0.0f;   // loss.cpp:59-68
        if (threadIdx.x == 0) p.lab32[n] = lab;
    } else if (vecI) {
        auto put = [&](int t, const float4 v) {
            const int y = t / rq, q = t - y * rq;
            if (p.Icopy) stg4(p.Icopy + (int64_t)n * nI + y * rowf + 4 * q, v);       // Model::forward: n0 = input
            float *d = sI + ((y + P) * WP + P) * C1 + 4 * q;
            d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
        };
        if (has0) put(threadIdx.x, v0);
        for (int t = threadIdx.x + blockDim.x; t < H * rq; t += blockDim.x) { const int y = t / rq, q = t - y * rq; put(t, ldg4(gI + y * rowf + 4 * q)); }
    } else {
        for (int t = threadIdx.x; t < nI; t += blockDim.x) {
            const int y = t / rowf, q = t - y * rowf;
            const float v = __ldg(gI + t);
            if (p.Icopy) p.Icopy[(int64_t)n * nI + t] = v;
            sI[((y + P) * WP + P) * C1 + q] = v;
        }
    }
    cp_async_wait_all();
    __syncthreads();
    float *gO = p.convO + (int64_t)n * H * W * C0;
    constexpr bool VEC = C0T > 0 && (C0T & 1) == 0;        // host checks 16-byte alignment of the tensors
    for (int w = threadIdx.x; w < nwin; w += blockDim.x) {
        const int j0 = w % Wp, i0 = w / Wp;
        float acc[4][CP];
        #pragma unroll
        for (int q = 0; q < 4; q++)
            #pragma unroll
            for (int c = 0; c < CP; c++) acc[q][c] = sB[c];
        #pragma unroll
        for (int ky = 0; ky < KS; ky++) {
            #pragma unroll
            for (int kx = 0; kx < KS; kx++) {
                const float *px = sI + ((2 * i0 + ky) * WP + 2 * j0 + kx) * C1;
                #pragma unroll
                for (int c1 = 0; c1 < (C1T ? C1T : 4); c1++) {
                    if (c1 >= C1) break;
                    const float v0 = px[c1], v1 = px[C1 + c1], v2_ = px[WP * C1 + c1], v3 = px[(WP + 1) * C1 + c1];
                    const float4 *f = reinterpret_cast<const float4*>(sF + ((ky * KS + kx) * C1 + c1) * CP);
                    #pragma unroll
                    for (int c4 = 0; c4 < CP / 4; c4++) {
                        const float4 fv = f[c4];
                        const float ff[4] = {fv.x, fv.y, fv.z, fv.w};
                        #pragma unroll
                        for (int k = 0; k < 4; k++) {
                            acc[0][c4 * 4 + k] = fmaf(ff[k], v0, acc[0][c4 * 4 + k]);
                            acc[1][c4 * 4 + k] = fmaf(ff[k], v1, acc[1][c4 * 4 + k]);
                            acc[2][c4 * 4 + k] = fmaf(ff[k], v2_, acc[2][c4 * 4 + k]);
                            acc[3][c4 * 4 + k] = fmaf(ff[k], v3, acc[3][c4 * 4 + k]);
                        }
                    }
                }
            }
        }
        // conv output: rows 2*i0, 2*i0+1; per row the pixel pair is 2*C0 contiguous floats
        #pragma unroll
        for (int dy = 0; dy < 2; dy++) {
            float *o = gO + ((2 * i0 + dy) * W + 2 * j0) * C0;
            if constexpr (VEC) {
                // the pixel pair is a run of 2*C0T floats: element e = (pixel e / C0T, channel e % C0T)
                #pragma unroll
                for (int e = 0; e < 2 * C0T; e += 4) {
                    float r[4];
                    #pragma unroll
                    for (int k = 0; k < 4; k++) r[k] = (e + k < C0T) ? acc[dy * 2][(e + k) % CP] : acc[dy * 2 + 1][(e + k >= C0T ? e + k - C0T : 0)];
                    stg4(o + e, make_float4(r[0], r[1], r[2], r[3]));
                }
            } else {
                #pragma unroll
                for (int c = 0; c < CP; c++) if (c < C0) { o[c] = acc[dy * 2][c]; o[C0 + c] = acc[dy * 2 + 1][c]; }
            }
        }
        // max-pool (k_pool<2> order) → staging
        float *sp = sP + w * C0;
        #pragma unroll
        for (int c = 0; c < CP; c++) {
            float v = acc[0][c];
            v = fmaxf(acc[1][c], v); v = fmaxf(acc[2][c], v); v = fmaxf(acc[3][c], v);
            if (c < C0) sp[c] = v;
        }
    }
    __syncthreads();
    // pooled tensors: pool value, relu (+mask, k_activate RELU), flatten copy — contiguous streams
    const int64_t gp = (int64_t)n * nP;
    const bool al = aligned16(p.poolO) && aligned16(p.actO) && aligned16(p.actF) && (!p.flatO || aligned16(p.flatO));
    if ((nP & 3) == 0 && al) {
        for (int t = threadIdx.x; t < (nP >> 2); t += blockDim.x) {
            const float4 v = *reinterpret_cast<const float4*>(sP + 4 * t);
            float4 a, f;
            if (v.x > 0.0f) { f.x = 1.0f; a.x = v.x; } else { f.x = 0.0f; a.x = 0.0f; }
            if (v.y > 0.0f) { f.y = 1.0f; a.y = v.y; } else { f.y = 0.0f; a.y = 0.0f; }
            if (v.z > 0.0f) { f.z = 1.0f; a.z = v.z; } else { f.z = 0.0f; a.z = 0.0f; }
            if (v.w > 0.0f) { f.w = 1.0f; a.w = v.w; } else { f.w = 0.0f; a.w = 0.0f; }
            stg4(p.poolO + gp + 4 * t, v); stg4(p.actO + gp + 4 * t, a); stg4(p.actF + gp + 4 * t, f);
            if (p.flatO) stg4(p.flatO + gp + 4 * t, a);
        }
    } else {
        for (int t = threadIdx.x; t < nP; t += blockDim.x) {
            const float v = sP[t];
            float a, f;
            if (v > 0.0f) { f = 1.0f; a = v; } else { f = 0.0f; a = 0.0f; }
            p.poolO[gp + t] = v; p.actO[gp + t] = a; p.actF[gp + t] = f;
            if (p.flatO) p.flatO[gp + t] = a;
        }
    }
}

// ------------------------------------------------------------------ transposing warp reduction
// v[0..32) per lane → lane l returns Σ_lanes v[l]   (16+8+4+2+1 = 31 shuffles instead of 32 x 5)
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

// ------------------------------------------------------------------ backward (C1 == 1, KS == 3)
// CM = compile-time channel bound (multiple of 2): 10 or 16.   Partials: part[n][nF + C0], nF = 9*C0.
// EXACT: C0 == CM (static register indexing, 128-bit global access)
template<int CM, bool EXACT>
__global__ void __launch_bounds__(256, 2) k_cpr2_bwd(Cpr2P p) {
    extern __shared__ __align__(16) float sm[];
    pdl_wait(); pdl_trigger();                  // PDL: nothing global before this line
    constexpr int KS = 3, P = 1;
    // dF/dB partials: NPASS passes over CH channels each; a pass accumulates the 9 taps x CH channels (+ the CH dB sums) in registers
    constexpr int CH = (CM == 10) ? 5 : 4, NPASS = CM / CH, NACC = 9 * CH;
    constexpr int NGP = (NACC + CH + 31) / 32;          // 32-value groups per pass (45 + 5 -> 2, 36 + 4 -> 2)
    static_assert(CM % CH == 0, "channel passes");
    const int H = p.H, W = p.W, C0 = EXACT ? CM : p.C0;
    const int WP = W + 2, HP = H + 2;
    const int RW = (WP + 1) & ~1;                       // even row stride → 8-byte aligned pairs
    float *sF = sm;                                     // [9][CM] original taps (dX uses the flipped index)
    float *sFx = sF + ((9 * CM + 3) & ~3);              // [CM][12] flipped taps, channel-major (dX gather: 3 x 128-bit per channel)
    float *sI = sFx + 12 * CM;                          // [HP][WP] zero halo (forward input, C1 == 1)
    float *sR = sI + ((HP * WP + 3) & ~3);              // [C0][HP][RW] routed gradient, zero halo
    float *sRed = sR + (((size_t)C0 * HP * RW + 3) & ~(size_t)3);   // [nwarps][NPASS*NGP*32]
    const int Hp = H / 2, Wp = W / 2, nwin = Hp * Wp, nP = nwin * C0;
    float *sD = sRed + (size_t)(blockDim.x >> 5) * (NPASS * NGP * 32);      // [nwin][C0] dY, overwritten with g = dY*mask
    const int n = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int nI = H * W;
    float *gO = p.convO + (int64_t)n * H * W * C0;
    constexpr bool VEC = EXACT && ((2 * CM) & 3) == 0;
    // the forward conv outputs of a window (4 pixels x C0), 128-bit loads: t_[dy*2+dx][c]
    auto load_window = [&](int w, float (&t_)[4][CM]) {
        const int j0 = w % Wp, i0 = w / Wp;
        #pragma unroll
        for (int dy = 0; dy < 2; dy++) {
            const float *o = gO + ((2 * i0 + dy) * W + 2 * j0) * C0;
            if constexpr (VEC) {
                #pragma unroll
                for (int e = 0; e < 2 * CM; e += 4) {
                    const float4 q = *reinterpret_cast<const float4*>(o + e);
                    const float qq[4] = {q.x, q.y, q.z, q.w};
                    #pragma unroll
                    for (int k = 0; k < 4; k++) {
                        if (e + k < CM) t_[dy * 2][(e + k) % CM] = qq[k];
                        else            t_[dy * 2 + 1][(e + k) % CM] = qq[k];
                    }
                }
            } else {
                #pragma unroll
                for (int c = 0; c < CM; c++) if (c < C0) { t_[dy * 2][c] = o[c]; t_[dy * 2 + 1][c] = o[C0 + c]; }
            }
        }
    };
    // Global-latency plan: the conv outputs of this thread's first NPF windows are requested FIRST, so that their round trip
    // overlaps the set-up below (pooled-tensor streams, taps, input tile, zeroing) instead of following its barrier; with
    // 128-thread CTAs (two windows per thread, cpr2_bwd_threads) every HBM read of the CTA is in flight before the barrier.
    // Order of issue: (1) asynchronous copies global -> shared (no registers, no stall): dY -> sD, the forward input -> sI, the
    // taps -> sF / sFx; (2) the relu mask, first chunk, into registers; (3) the window prefetch; (4) shared-memory zeroing while
    // all of that is in flight; then the copies are awaited and consumed.
    constexpr int NPF = (CM <= 10) ? 2 : 1;             // 2 x 4 x CM registers held across the set-up
    constexpr int MU = 4;                               // mask quads per thread held in registers per chunk
    const int64_t gp = (int64_t)n * nP;
    const bool alp = (nP & 3) == 0 && aligned16(p.dY) && aligned16(p.actFc) && aligned16(p.actO) && aligned16(p.poolO);
    const int nq4 = nP >> 2;
    float *gI = p.Iio + (int64_t)n * nI;
    if (alp) for (int t = threadIdx.x; t < nq4; t += blockDim.x) cp_async16(sD + 4 * t, p.dY + gp + 4 * t);
    for (int t = threadIdx.x; t < nI; t += blockDim.x) { const int y = t / W; cp_async4(sI + (y + 1) * WP + 1 + (t - y * W), gI + t, true); }
    for (int t = threadIdx.x; t < 9 * CM; t += blockDim.x) { const int c = t % CM, tap = t / CM; cp_async4(sF + t, p.F + (c < C0 ? tap * C0 + c : 0), c < C0); }
    for (int t = threadIdx.x; t < 12 * CM; t += blockDim.x) { const int k = t % 12, c = t / 12; const bool on = (c < C0 && k < 9); cp_async4(sFx + t, p.F + (on ? (8 - k) * C0 + c : 0), on); }
    cp_async_commit();
    float4 mv[MU];
    if (alp) {
        #pragma unroll
        for (int u = 0; u < MU; u++) { const int t = threadIdx.x + u * blockDim.x; if (t < nq4) mv[u] = ldg4(p.actFc + gp + 4 * t); }
    }
    float tpre[NPF][4][CM];
    #pragma unroll
    for (int wi = 0; wi < NPF; wi++) {
        const int w = threadIdx.x + wi * blockDim.x;
        if (w < nwin) load_window(w, tpre[wi]);
    }
    for (int t = threadIdx.x; t < WP; t += blockDim.x) { sI[t] = 0.0f; sI[(HP - 1) * WP + t] = 0.0f; }
    for (int t = threadIdx.x; t < H; t += blockDim.x) { sI[(t + 1) * WP] = 0.0f; sI[(t + 1) * WP + W + 1] = 0.0f; }
    if (p.zero_all) {   // A/B knob (T4K_CPR_ZERO=1): clear the whole routed tile, as before
        const int nq = (int)((((size_t)C0 * HP * RW + 3) & ~(size_t)3) >> 2);
        for (int t = threadIdx.x; t < nq; t += blockDim.x) *reinterpret_cast<float4*>(sR + 4 * t) = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
        // zero the HALO of the routed tile only: phase A writes all four pixels of every window, i.e. the whole interior, after the barrier —
        // clearing it here as well was 9000 of the CTA's shared-memory stores for nothing.  Rows 0 and HP-1 as 64-bit stores (RW is even), then the
        // left column and the right column (+ row padding) of the rows in between.
        const int rq = RW >> 1;
        for (int t = threadIdx.x; t < C0 * 2 * rq; t += blockDim.x) {
            const int c = t / (2 * rq), k = t - c * 2 * rq;
            float *row = sR + c * (HP * RW) + (k < rq ? 0 : (HP - 1) * RW);
            *reinterpret_cast<float2*>(row + 2 * (k < rq ? k : k - rq)) = make_float2(0.f, 0.f);
        }
        for (int t = threadIdx.x; t < C0 * H; t += blockDim.x) {
            const int c = t / H, y = t - c * H;
            float *row = sR + c * (HP * RW) + (y + 1) * RW;
            row[0] = 0.0f;
            for (int x = WP - 1; x < RW; x++) row[x] = 0.0f;
        }
    }
    cp_async_wait_all();                                // this thread's own copies have landed (it consumes only those before the barrier)
    if (alp) {                                          // coalesced 128-bit streams of the pooled-size tensors
        auto consume = [&](int t, const float4 m) {
            const float4 d = *reinterpret_cast<const float4*>(sD + 4 * t);
            if (p.actO != p.dY) stg4(p.actO + gp + 4 * t, d);                         // flatten backward: in = out
            const float4 g = make_float4(__fmul_rn(d.x, m.x), __fmul_rn(d.y, m.y), __fmul_rn(d.z, m.z), __fmul_rn(d.w, m.w));
            stg4(p.poolO + gp + 4 * t, g);                                            // _bactivate: in = out * mask
            *reinterpret_cast<float4*>(sD + 4 * t) = g;
        };
        #pragma unroll
        for (int u = 0; u < MU; u++) { const int t = threadIdx.x + u * blockDim.x; if (t < nq4) consume(t, mv[u]); }
        for (int t = threadIdx.x + MU * blockDim.x; t < nq4; t += blockDim.x) consume(t, ldg4(p.actFc + gp + 4 * t));
    } else {
        for (int t = threadIdx.x; t < nP; t += blockDim.x) {
            const float d = p.dY[gp + t];
            if (p.actO != p.dY) p.actO[gp + t] = d;
            const float g = __fmul_rn(d, p.actFc[gp + t]);
            p.poolO[gp + t] = g; sD[t] = g;
        }
    }
    __syncthreads();
    constexpr int RSTRIDE = NPASS * NGP * 32;
    const int cs = HP * RW;                               // channel stride of the routed tile
    float accB[CM];
    #pragma unroll
    for (int c = 0; c < CM; c++) accB[c] = 0.0f;
    // ---- phase A (once per window): relu backward is in sD, arg-max routing (first strict max in y,x order, nmath.tcu:535-549),
    //      routed gradient back to HBM (in place of the conv output) and into the haloed channel-major tile
    auto route_window = [&](int w, float (&t_)[4][CM]) {
        const int j0 = w % Wp, i0 = w / Wp;
        const int rb0 = 2 * i0 * RW + 2 * j0;              // window origin in the haloed routed tile (row 2*i0, col 2*j0)
        const float *sg = sD + w * C0;
        #pragma unroll
        for (int c = 0; c < CM; c++) {
            if (c < C0) {
                const float g = sg[c];
                float best = t_[0][c]; int arg = 0;
                if (t_[1][c] > best) { best = t_[1][c]; arg = 1; }
                if (t_[2][c] > best) { best = t_[2][c]; arg = 2; }
                if (t_[3][c] > best) { best = t_[3][c]; arg = 3; }
                t_[0][c] = (arg == 0) ? g : 0.0f; t_[1][c] = (arg == 1) ? g : 0.0f;
                t_[2][c] = (arg == 2) ? g : 0.0f; t_[3][c] = (arg == 3) ? g : 0.0f;
                accB[c] += g;
                float *rr = sR + c * cs + rb0 + RW + 1;
                rr[0] = t_[0][c]; rr[1] = t_[1][c]; rr[RW] = t_[2][c]; rr[RW + 1] = t_[3][c];
            } else { t_[0][c] = t_[1][c] = t_[2][c] = t_[3][c] = 0.0f; }
        }
        #pragma unroll
        for (int dy = 0; dy < 2; dy++) {
            float *o = gO + ((2 * i0 + dy) * W + 2 * j0) * C0;
            if constexpr (VEC) {
                #pragma unroll
                for (int e = 0; e < 2 * CM; e += 4) {
                    float q[4];
                    #pragma unroll
                    for (int k = 0; k < 4; k++) q[k] = (e + k < CM) ? t_[dy * 2][(e + k) % CM] : t_[dy * 2 + 1][(e + k) % CM];
                    stg4(o + e, make_float4(q[0], q[1], q[2], q[3]));
                }
            } else {
                #pragma unroll
                for (int c = 0; c < CM; c++) if (c < C0) { o[c] = t_[dy * 2][c]; o[C0 + c] = t_[dy * 2 + 1][c]; }
            }
        }
    };
    #pragma unroll
    for (int wi = 0; wi < NPF; wi++) {
        const int w = threadIdx.x + wi * blockDim.x;
        if (w < nwin) route_window(w, tpre[wi]);
    }
    for (int w = threadIdx.x + NPF * blockDim.x; w < nwin; w += blockDim.x) {       // more than NPF windows per thread: load in place
        float t_[4][CM];
        load_window(w, t_);
        route_window(w, t_);
    }
    // ---- dF / dB: NPASS passes of CH channels.  Per window and pass: the 4 x 4 input neighbourhood (8 x 64-bit loads) and the window's
    //      own routed values for CH channels (read back from the tile it just wrote) feed 9 taps x CH channels x 4 pixels of FMAs; every
    //      routed value is read ONCE for all nine taps (a ky-major loop would read it three times).  Same FMA order per (tap, channel)
    //      as before: window after window, pixels (0,0) (0,1) (1,0) (1,1).
    if (p.train) {
        #pragma unroll
        for (int pass = 0; pass < NPASS; pass++) {
            float acc[NGP * 32];                          // [(ky*3+kx)*CH + cc], then the CH dB sums
            #pragma unroll
            for (int g = 0; g < NGP * 32; g++) acc[g] = 0.0f;
            for (int w = threadIdx.x; w < nwin; w += blockDim.x) {
                const int j0 = w % Wp, i0 = w / Wp;
                const int rb0 = 2 * i0 * RW + 2 * j0;
                float nb[4][4];
                #pragma unroll
                for (int a = 0; a < 4; a++) {
                    const float *ip = sI + (2 * i0 + a) * WP + 2 * j0;       // WP even (host-checked): 8-byte aligned pairs
                    const float2 u = *reinterpret_cast<const float2*>(ip), v = *reinterpret_cast<const float2*>(ip + 2);
                    nb[a][0] = u.x; nb[a][1] = u.y; nb[a][2] = v.x; nb[a][3] = v.y;
                }
                float r[4][CH];
                #pragma unroll
                for (int cc = 0; cc < CH; cc++) {
                    const int c = pass * CH + cc;
                    if (c < C0) {
                        const float *rr = sR + c * cs + rb0 + RW + 1;
                        r[0][cc] = rr[0]; r[1][cc] = rr[1]; r[2][cc] = rr[RW]; r[3][cc] = rr[RW + 1];
                    } else { r[0][cc] = r[1][cc] = r[2][cc] = r[3][cc] = 0.0f; }
                }
                // dF[ky][kx][c] += Σ_{pixel (dy,dx) of the window} I[2*i0+dy+ky-1][2*j0+dx+kx-1] * r[dy*2+dx][c]
                #pragma unroll
                for (int ky = 0; ky < 3; ky++)
                    #pragma unroll
                    for (int kx = 0; kx < 3; kx++)
                        #pragma unroll
                        for (int cc = 0; cc < CH; cc++) {
                            float a = acc[(ky * 3 + kx) * CH + cc];
                            a = fmaf(nb[ky][kx], r[0][cc], a);     a = fmaf(nb[ky][kx + 1], r[1][cc], a);
                            a = fmaf(nb[ky + 1][kx], r[2][cc], a); a = fmaf(nb[ky + 1][kx + 1], r[3][cc], a);
                            acc[(ky * 3 + kx) * CH + cc] = a;
                        }
            }
            #pragma unroll
            for (int cc = 0; cc < CH; cc++) acc[NACC + cc] = accB[pass * CH + cc];
            #pragma unroll
            for (int g = 0; g < NGP; g++) {
                float v[32];
                #pragma unroll
                for (int k = 0; k < 32; k++) v[k] = acc[g * 32 + k];
                const float sv_ = warp_treduce32(v, lane);
                sRed[warp * RSTRIDE + (pass * NGP + g) * 32 + lane] = sv_;
            }
        }
    }
    __syncthreads();
    if (p.train) {
        const int nF = 9 * C0, nE = nF + C0;
        for (int t = threadIdx.x; t < nE; t += blockDim.x) {
            int c, idx;
            if (t < nF) { c = t % C0; idx = (t / C0) * CH + c % CH; }      // t / C0 = ky*3 + kx
            else        { c = t - nF; idx = NACC + c % CH; }
            const int slot = (c / CH) * NGP * 32 + idx;
            float s_ = 0.0f;
            for (int wv = 0; wv < nwarps; wv++) s_ += sRed[wv * RSTRIDE + slot];
            p.part[(int64_t)n * nE + t] = s_;
        }
    }
    // ---- dX (flipped taps, nmath.tcu:304): dX[y][x] = Σ_c Σ_{ky,kx} F[2-ky][2-kx][c] * R[y+1-ky][x+1-kx][c]
    for (int w = threadIdx.x; w < nwin; w += blockDim.x) {
        const int j0 = w % Wp, i0 = w / Wp;
        float a00 = 0.f, a01 = 0.f, a10 = 0.f, a11 = 0.f;
        for (int c = 0; c < C0; c++) {
            // neighbourhood rows 2*i0 .. 2*i0+3, cols 2*j0 .. 2*j0+3 in halo coordinates
            const float *rb = sR + c * cs + 2 * i0 * RW + 2 * j0;
            float nb[4][4];
            #pragma unroll
            for (int rr = 0; rr < 4; rr++) {
                const float2 u = *reinterpret_cast<const float2*>(rb + rr * RW);
                const float2 v = *reinterpret_cast<const float2*>(rb + rr * RW + 2);
                nb[rr][0] = u.x; nb[rr][1] = u.y; nb[rr][2] = v.x; nb[rr][3] = v.y;
            }
            float fx[12];
            #pragma unroll
            for (int q = 0; q < 3; q++) {
                const float4 fq = *reinterpret_cast<const float4*>(sFx + c * 12 + 4 * q);
                fx[4 * q] = fq.x; fx[4 * q + 1] = fq.y; fx[4 * q + 2] = fq.z; fx[4 * q + 3] = fq.w;
            }
            #pragma unroll
            for (int ky = 0; ky < 3; ky++) {
                #pragma unroll
                for (int kx = 0; kx < 3; kx++) {
                    const float f = fx[ky * 3 + kx];               // = F[2-ky][2-kx][c]
                    // pixel (dy,dx): halo row = 2*i0+dy + 2 - ky → nb row dy + 2 - ky ; col dx + 2 - kx
                    a00 = fmaf(f, nb[2 - ky][2 - kx], a00);
                    a01 = fmaf(f, nb[2 - ky][3 - kx], a01);
                    a10 = fmaf(f, nb[3 - ky][2 - kx], a10);
                    a11 = fmaf(f, nb[3 - ky][3 - kx], a11);
                }
            }
        }
        float *o = gI + (2 * i0) * W + 2 * j0;
        float *b = p.dXbuf + (int64_t)n * nI + (2 * i0) * W + 2 * j0;
        *reinterpret_cast<float2*>(o) = make_float2(a00, a01); *reinterpret_cast<float2*>(o + W) = make_float2(a10, a11);
        *reinterpret_cast<float2*>(b) = make_float2(a00, a01); *reinterpret_cast<float2*>(b + W) = make_float2(a10, a11);
    }
}

static bool cpr2_fwd_ok(int H1, int W1, int C1, int H0, int W0, int C0, int KS, int S, int P, int *CP, size_t *smem) {
    if (!(KS == 3 || KS == 5) || S != 1 || P != (KS - 1) / 2 || H0 != H1 || W0 != W1) return false;
    if (C1 > 4 || C0 > 16 || (H0 & 1) || (W0 & 1)) return false;
    *CP = (C0 <= 4) ? 4 : (C0 <= 8) ? 8 : (C0 <= 12) ? 12 : 16;
    const size_t nP = (size_t)(H0 / 2) * (W0 / 2) * C0;
    *smem = ((size_t)KS * KS * C1 * *CP + *CP + ((nP + 3) & ~(size_t)3) + (size_t)(H1 + 2 * P) * (W1 + 2 * P) * C1) * sizeof(float);
    return *smem <= 96 * 1024;
}
static bool cpr2_bwd_ok(int H1, int W1, int C1, int H0, int W0, int C0, int KS, int S, int P, int *CM, size_t *smem, int threads) {
    if (KS != 3 || S != 1 || P != 1 || H0 != H1 || W0 != W1 || C1 != 1 || C0 > 16 || (H0 & 1) || (W0 & 1)) return false;
    *CM = (C0 <= 10) ? 10 : 16;
    const int CH = (*CM == 10) ? 5 : 4, NPASS = *CM / CH, NGP = (9 * CH + CH + 31) / 32;      // as in k_cpr2_bwd
    const int HP = H1 + 2, WP = W1 + 2, RW = (WP + 1) & ~1;
    const size_t nP = (size_t)(H0 / 2) * (W0 / 2) * C0;
    *smem = ((size_t)((9 * *CM + 3) & ~3) + 12 * *CM + ((HP * WP + 3) & ~3) + (((size_t)C0 * HP * RW + 3) & ~(size_t)3) + (size_t)(threads / 32) * (NPASS * NGP * 32) +
             ((nP + 3) & ~(size_t)3)) * sizeof(float);
    return *smem <= 100 * 1024 && (W1 & 1) == 0;
}
static int cpr2_bwd_threads(int nwin) {
    const int full = (nwin + 31) & ~31;
    if (full <= 128 || full > 256) return full > 256 ? 256 : (full < 32 ? 32 : full);
    return (((nwin + 1) / 2) + 31) & ~31;                  // two windows per thread
}
static int win_threads(int nwin) { int t = (nwin + 31) & ~31; return t > 256 ? 256 : (t < 32 ? 32 : t); }

// feed != nullptr: dataset feed folded in (fields u8I .. feedN of *feed are copied into the launch parameters; caller checked eligibility)
static int cpr2_fwd_launch(const float *I, const float *F, const float *B, float *Icopy, float *convO, float *poolO,
                           float *actO, float *actF, float *flatO, int N, int H1, int W1, int C1, int H0, int W0,
                           int C0, int KS, int S, int P, const Cpr2P *feed, cudaStream_t st) {
    int CP = 0; size_t smem = 0;
    if (!cpr2_fwd_ok(H1, W1, C1, H0, W0, C0, KS, S, P, &CP, &smem)) return T4K_ENOSUP;
    Cpr2P p{}; p.I = I; p.F = F; p.B = B; p.Icopy = (Icopy == I) ? nullptr : Icopy; p.convO = convO; p.poolO = poolO; p.actO = actO;
    p.actF = actF; p.flatO = flatO; p.H = H1; p.W = W1; p.C1 = C1; p.C0 = C0;
    const int threads = win_threads((H0 / 2) * (W0 / 2));
    const bool al = aligned16(convO) && aligned16(poolO) && aligned16(actO) && aligned16(actF) && (!flatO || aligned16(flatO));
    #define CPR2F(K_, CP_, CT_, C1_, FD_) { static DevFlag attr; if (smem > 48 * 1024 && dev_first(attr)) { cudaFuncSetAttribute(k_cpr2_fwd<K_, CP_, CT_, C1_, FD_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024); } \
                                  launch_std(k_cpr2_fwd<K_, CP_, CT_, C1_, FD_>, dim3(N), dim3(threads), smem, st, p); }
    if (feed) {
        // the feed variants exist for the exact single-input-channel 3x3 shapes (the MNIST-style first block)
        if (!(KS == 3 && al && C1 == 1 && (C0 == 10 || C0 == 16 || C0 == 8)) || feed->E > threads) return T4K_ENOSUP;
        p.u8I = feed->u8I; p.u8L = feed->u8L; p.mean = feed->mean; p.scale = feed->scale; p.lab32 = feed->lab32; p.hot = feed->hot; p.E = feed->E; p.feedN = feed->feedN;
        if (C0 == 10) CPR2F(3, 12, 10, 1, true) else if (C0 == 16) CPR2F(3, 16, 16, 1, true) else CPR2F(3, 8, 8, 1, true)
        return check_launch();
    }
    if (KS == 3) {
        if (al && C1 == 1 && C0 == 10) CPR2F(3, 12, 10, 1, false) else if (al && C1 == 1 && C0 == 16) CPR2F(3, 16, 16, 1, false) else if (al && C1 == 1 && C0 == 8) CPR2F(3, 8, 8, 1, false)
        else switch (CP) { case 4: CPR2F(3, 4, 0, 0, false) break; case 8: CPR2F(3, 8, 0, 0, false) break; case 12: CPR2F(3, 12, 0, 0, false) break; default: CPR2F(3, 16, 0, 0, false) break; }
    } else {
        switch (CP) { case 4: CPR2F(5, 4, 0, 0, false) break; case 8: CPR2F(5, 8, 0, 0, false) break; case 12: CPR2F(5, 12, 0, 0, false) break; default: CPR2F(5, 16, 0, 0, false) break; }
    }
    return check_launch();
}

} // namespace t4k
using namespace t4k;

extern "C" int t4k_conv_pool_relu_fwd(const float *I, const float *F, const float *B, float *Icopy, float *convO, float *poolO,
                                      float *actO, float *actF, float *flatO, int N, int H1, int W1, int C1, int H0, int W0,
                                      int C0, int KS, int S, int P, t4k_stream_t s) {
    if (!I || !F || !B || !convO || !poolO || !actO || !actF || N < 1) return T4K_EINVAL;
    int rc = cpr2_fwd_launch(I, F, B, Icopy, convO, poolO, actO, actF, flatO, N, H1, W1, C1, H0, W0, C0, KS, S, P, nullptr, STRM(s));
    if (rc != T4K_ENOSUP) return rc;
    if (Icopy && Icopy != I) { rc = t4k_copy(I, Icopy, (int64_t)N * H1 * W1 * C1, s); if (rc) return rc; }
    return cpr_v1_fwd(I, F, B, convO, poolO, actO, actF, flatO, N, H1, W1, C1, H0, W0, C0, KS, S, P, STRM(s));
}

// The forward block fed straight from a staged U8 mini-batch: Dataset::_load (src/mu/dataset.cu:124-152) + Model::onehot
// (src/nn/loss.cpp:47-72) + the block, one launch.  `data` is the dataset tensor [N,H1,W1,C1]: its first feedN samples are
// rewritten with ((float)u8 - mean) * scale, the rest keep their values (partial batch, as _load); Icopy = the model's input layer.
// Shapes outside the fused envelope run t4k_dataset_load followed by t4k_conv_pool_relu_fwd (same results).
extern "C" int t4k_conv_pool_relu_fwd_feed(const uint8_t *u8I, const uint8_t *u8L, int feedN, float mean, float scale, int32_t *lab32,
                                           float *hot, int E, float *data, const float *F, const float *B, float *Icopy, float *convO,
                                           float *poolO, float *actO, float *actF, float *flatO, int N, int H1, int W1, int C1,
                                           int H0, int W0, int C0, int KS, int S, int P, t4k_stream_t s) {
    if (!u8I || !u8L || !lab32 || !data || !F || !B || !convO || !poolO || !actO || !actF || N < 1 || feedN < 0 || feedN > N || (hot && E < 1)) return T4K_EINVAL;
    const int nI = H1 * W1 * C1;
    const bool fusable = hot && Icopy && Icopy != data && (nI & 15) == 0 && ((W1 * C1) & 3) == 0 && aligned16(u8I) && aligned16(data) && aligned16(Icopy);
    if (fusable) {
        Cpr2P fd{}; fd.u8I = u8I; fd.u8L = u8L; fd.mean = mean; fd.scale = scale; fd.lab32 = lab32; fd.hot = hot; fd.E = E; fd.feedN = feedN;
        int rc = cpr2_fwd_launch(data, F, B, Icopy, convO, poolO, actO, actF, flatO, N, H1, W1, C1, H0, W0, C0, KS, S, P, &fd, STRM(s));
        if (rc != T4K_ENOSUP) return rc;
    }
    int rc = t4k_dataset_load(u8I, data, (int64_t)feedN * nI, mean, scale, u8L, lab32, feedN, hot, E, s);
    if (rc) return rc;
    return t4k_conv_pool_relu_fwd(data, F, B, Icopy, convO, poolO, actO, actF, flatO, N, H1, W1, C1, H0, W0, C0, KS, S, P, s);
}

// One-shot hook: the NEXT fused conv-block backward of this thread records `event` between its two launches — behind the machine-filling main
// kernel, in front of the short finish launch.  A caller hangs side-stream work there that must not share the SMs with the main kernel (it
// fills them in exactly one wave) but may overlap the finish: the data-parallel exchange of the rest of the arena (Model::backprop).
static thread_local void *g_cpr_mid_event = nullptr;
extern "C" int t4k_conv_pool_relu_bwd_mid_event(void *event) { g_cpr_mid_event = event; return 0; }
extern "C" int t4k_conv_pool_relu_bwd(const float *dY, float *actO, const float *actF, float *poolO, float *convO, float *Iio, float *dXbuf,
                                      const float *F, float *dF, float *dB, int N, int H1, int W1, int C1, int H0, int W0, int C0,
                                      int KS, int S, int P, int train, t4k_stream_t s) {
    return t4k_conv_pool_relu_bwd_opt(dY, actO, actF, poolO, convO, Iio, dXbuf, F, dF, dB, N, H1, W1, C1, H0, W0, C0, KS, S, P, train, nullptr, s);
}
extern "C" int t4k_conv_pool_relu_bwd_opt(const float *dY, float *actO, const float *actF, float *poolO, float *convO, float *Iio, float *dXbuf,
                                          const float *F, float *dF, float *dB, int N, int H1, int W1, int C1, int H0, int W0, int C0,
                                          int KS, int S, int P, int train, const t4k_fused_opt_t *opt, t4k_stream_t s) {
    if (!dY || !actO || !actF || !poolO || !convO || !Iio || !dXbuf || !F || N < 1 || (train && (!dF || !dB))) return T4K_EINVAL;
    int CM = 0; size_t smem = 0;
    // CTA width of the backward block: one thread per pool window (224 threads at 14x14 windows, 2 CTAs/SM at 126 registers:
    // 512 samples = 1.73 waves), or HALF the windows per pass (128 threads x 4 CTAs/SM = 592 slots: the batch is one wave and each
    // thread walks two windows).  T4K_CPR2_BWD_THREADS overrides (multiple of 32).
    static int thr_env = -1;
    if (thr_env < 0) { const char *e = getenv("T4K_CPR2_BWD_THREADS"); thr_env = e ? atoi(e) : 0; if (thr_env < 32 || thr_env > 256 || (thr_env & 31)) thr_env = 0; }
    const int threads = thr_env ? thr_env : cpr2_bwd_threads((H0 / 2) * (W0 / 2));
    if (!cpr2_bwd_ok(H1, W1, C1, H0, W0, C0, KS, S, P, &CM, &smem, threads)) {
        if (opt) return T4K_ENOSUP;                       // the fused optimizer rides in this generation's finish launch only
        return cpr_v1_bwd(dY, actO, actF, poolO, convO, Iio, dXbuf, F, dF, dB, N, H1, W1, C1, H0, W0, C0, KS, S, P, train, STRM(s));
    }
    const int nF = 9 * C0;
    Cpr2P p{}; p.F = F; p.dY = dY; p.actFc = actF; p.actO = actO; p.poolO = poolO; p.convO = convO; p.Iio = Iio; p.dXbuf = dXbuf;
    p.H = H1; p.W = W1; p.C1 = 1; p.C0 = C0; p.train = train;
    { static int za = []{ const char *e = getenv("T4K_CPR_ZERO"); return (e && e[0] == '1') ? 1 : 0; }(); p.zero_all = za; }
    if (train) { p.part = (float*)workspace((size_t)N * (nF + C0) * sizeof(float), 4); if (!p.part) return T4K_ENOMEM; }
    const bool ex = (C0 == CM) && aligned16(convO);
    #define CPR2B(CM_, EX_) { static DevFlag attr; if (dev_first(attr)) { cudaFuncSetAttribute(k_cpr2_bwd<CM_, EX_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024); cudaFuncSetAttribute(k_cpr2_bwd<CM_, EX_>, cudaFuncAttributePreferredSharedMemoryCarveout, 100); } \
                              launch_std(k_cpr2_bwd<CM_, EX_>, dim3(N), dim3(threads), smem, STRM(s), p); }
    if (CM == 10) { if (ex) CPR2B(10, true) else CPR2B(10, false) } else { if (ex) CPR2B(16, true) else CPR2B(16, false) }
    int rc = check_launch(); if (rc || !train) return rc;
    if (g_cpr_mid_event) { cudaEventRecord((cudaEvent_t)g_cpr_mid_event, STRM(s)); g_cpr_mid_event = nullptr; }   // t4k_conv_pool_relu_bwd_mid_event
    if (opt) return wgrad_fin_opt_launch(p.part, dF, dB, nF, C0, N, KS, S, opt, STRM(s));
    return wgrad_fin_launch(p.part, dF, dB, nF, C0, N, KS, S, STRM(s));
}
