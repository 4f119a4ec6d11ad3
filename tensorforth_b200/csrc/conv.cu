// conv.cu — conv2d forward / backward, NHWC activations, filter [C1,KS,KS,C0]
//   replaces k_conv2d<TS,KS,S,P> (src/nn/nmath.tcu:34-104, launched by Model::_fconv
//   src/nn/forward.cu:126-155) and k_dconv2d<TS,KS,S,P> (src/nn/nmath.tcu:211-338,
//   Model::_bconv src/nn/backprop.cu:153-191).
// The reference accumulates over c1 with global atomicAdd into a pre-zeroed output and launches
// C0*C1*N tiny blocks; here every output element is produced once by one thread/tile:
//   * small-channel path (MNIST first layer, C0<=16): one thread per output pixel, filter in smem,
//     HBM-bound streaming kernel (reads I once, writes O once, coalesced 128-byte rows of C0);
//   * general path: implicit GEMM on CUDA cores (64x64x16 tiles, on-the-fly im2col gather),
//     M = N*H*W pixels, N = channels, K = KS*KS*C (k order ky,kx,c so NHWC patches are contiguous);
//   * weight gradient: split-K over pixels with per-CTA partials + ordered finalize (deterministic;
//     the reference uses shared + global atomics, nmath.tcu:307-336).
// dX uses the reference's 180-degree flipped filter taps (nmath.tcu:304) — replicated, not "fixed".
#include "common.cuh"
#include "optim.cuh"

namespace t4k {

struct ConvP {
    const float *I, *F, *B, *dO;
    float *O, *dX, *dF, *dB;
    int N, H1, W1, C1, H0, W0, C0, KS, S, P;
};

// ====================================================================== small-channel forward
// thread = one output pixel (n,i,j); acc[C0<=16]; filter (C1*KS*KS*C0 floats) + bias in smem
template<int KS, int C0MAX>
__global__ void __launch_bounds__(T4K_THREADS) k_conv_fwd_small(ConvP p) {
    extern __shared__ float sF[];                       // [C1][KS][KS][C0] then bias[C0]
    const int nF = p.C1 * KS * KS * p.C0;
    for (int t = threadIdx.x; t < nF; t += blockDim.x) sF[t] = __ldg(p.F + t);
    for (int t = threadIdx.x; t < p.C0; t += blockDim.x) sF[nF + t] = __ldg(p.B + t);
    __syncthreads();
    const int64_t npix = (int64_t)p.N * p.H0 * p.W0;
    const int C0 = p.C0, C1 = p.C1, S = p.S, P = p.P;
    for (int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pix < npix; pix += (int64_t)gridDim.x * blockDim.x) {
        const int j = (int)(pix % p.W0); const int64_t t = pix / p.W0;
        const int i = (int)(t % p.H0);   const int n = (int)(t / p.H0);
        float acc[C0MAX];
        #pragma unroll
        for (int c = 0; c < C0MAX; c++) acc[c] = (c < C0) ? sF[nF + c] : 0.0f;
        const float *nI = p.I + (int64_t)n * p.H1 * p.W1 * C1;
        #pragma unroll
        for (int y = 0; y < KS; y++) {
            const int gi = i * S + y - P;
            if (gi < 0 || gi >= p.H1) continue;
            #pragma unroll
            for (int x = 0; x < KS; x++) {
                const int gj = j * S + x - P;
                if (gj < 0 || gj >= p.W1) continue;
                const float *px = nI + ((int64_t)p.W1 * gi + gj) * C1;
                for (int c1 = 0; c1 < C1; c1++) {
                    const float v = __ldg(px + c1);
                    const float *f = sF + ((c1 * KS + y) * KS + x) * C0;
                    #pragma unroll
                    for (int c = 0; c < C0MAX; c++) if (c < C0) acc[c] = fmaf(f[c], v, acc[c]);
                }
            }
        }
        float *o = p.O + pix * C0;
        #pragma unroll
        for (int c = 0; c < C0MAX; c++) if (c < C0) o[c] = acc[c];
    }
}

// ====================================================================== small-channel dgrad
// thread = one input pixel (n,y,x), C1 <= 4 accumulators; flipped taps (nmath.tcu:304)
template<int KS, int C1MAX>
__global__ void __launch_bounds__(T4K_THREADS) k_conv_dgrad_small(ConvP p) {
    extern __shared__ float sF[];
    const int nF = p.C1 * KS * KS * p.C0;
    for (int t = threadIdx.x; t < nF; t += blockDim.x) sF[t] = __ldg(p.F + t);
    __syncthreads();
    const int64_t npix = (int64_t)p.N * p.H1 * p.W1;
    const int C0 = p.C0, C1 = p.C1, S = p.S, P = p.P;
    for (int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pix < npix; pix += (int64_t)gridDim.x * blockDim.x) {
        const int x = (int)(pix % p.W1); const int64_t t = pix / p.W1;
        const int y = (int)(t % p.H1);   const int n = (int)(t / p.H1);
        float acc[C1MAX];
        #pragma unroll
        for (int c = 0; c < C1MAX; c++) acc[c] = 0.0f;
        const float *nO = p.dO + (int64_t)n * p.H0 * p.W0 * C0;
        #pragma unroll
        for (int ky = 0; ky < KS; ky++) {
            const int ti = y + P - ky;                         // = i*S
            if (ti < 0 || (ti % S) != 0) continue;
            const int i = ti / S;
            if (i >= p.H0) continue;
            #pragma unroll
            for (int kx = 0; kx < KS; kx++) {
                const int tj = x + P - kx;
                if (tj < 0 || (tj % S) != 0) continue;
                const int j = tj / S;
                if (j >= p.W0) continue;
                const float *d = nO + ((int64_t)p.W0 * i + j) * C0;
                for (int c0 = 0; c0 < C0; c0++) {
                    const float dv = __ldg(d + c0);
                    #pragma unroll
                    for (int c1 = 0; c1 < C1MAX; c1++)
                        if (c1 < C1) acc[c1] = fmaf(sF[((c1 * KS + (KS - 1 - ky)) * KS + (KS - 1 - kx)) * C0 + c0], dv, acc[c1]);
                }
            }
        }
        float *o = p.dX + pix * C1;
        #pragma unroll
        for (int c1 = 0; c1 < C1MAX; c1++) if (c1 < C1) o[c1] = acc[c1];
    }
}

// ====================================================================== small wgrad (+dB)
// CTA = a strip of WG_ROWS output rows of one sample; dO strip and the matching input patch are
// staged in smem; thread t < nF owns one filter element (c1,ky,kx,c0), threads nF..nF+C0 own dB.
// Partials go to part[cta][nF + C0]; k_wgrad_fin sums them in CTA order and adds into dF / dB.
#define WG_ROWS 4
template<int KS>
__global__ void __launch_bounds__(T4K_THREADS) k_conv_wgrad_small(ConvP p, float *part, int strips) {
    extern __shared__ float sm[];
    const int C0 = p.C0, C1 = p.C1, S = p.S, P = p.P, W0 = p.W0, W1 = p.W1;
    const int n = blockIdx.x / strips, strip = blockIdx.x % strips;
    const int i0 = strip * WG_ROWS, rows = min(WG_ROWS, p.H0 - i0);
    const int prow = (WG_ROWS - 1) * S + KS;                       // input rows covering the strip
    float *sO = sm;                                                // [WG_ROWS][W0][C0]
    float *sI = sm + WG_ROWS * W0 * C0;                            // [prow][W1][C1]
    const float *nO = p.dO + ((int64_t)n * p.H0 + i0) * W0 * C0;
    for (int t = threadIdx.x; t < rows * W0 * C0; t += blockDim.x) sO[t] = __ldg(nO + t);
    const int gi0 = i0 * S - P;
    for (int t = threadIdx.x; t < prow * W1 * C1; t += blockDim.x) {
        const int r = t / (W1 * C1), gi = gi0 + r;
        sI[t] = (gi >= 0 && gi < p.H1) ? __ldg(p.I + ((int64_t)n * p.H1 + gi) * W1 * C1 + (t % (W1 * C1))) : 0.0f;
    }
    __syncthreads();
    const int nF = C1 * KS * KS * C0;
    for (int t = threadIdx.x; t < nF + C0; t += blockDim.x) {
        float acc = 0.0f;
        if (t < nF) {
            const int c0 = t % C0; int r = t / C0;
            const int kx = r % KS; r /= KS; const int ky = r % KS; const int c1 = r / KS;
            for (int i = 0; i < rows; i++) {
                const float *irow = sI + (i * S + ky) * W1 * C1 + c1;
                const float *orow = sO + i * W0 * C0 + c0;
                for (int j = 0; j < W0; j++) {
                    const int gj = j * S + kx - P;
                    if (gj >= 0 && gj < W1) acc = fmaf(irow[gj * C1], orow[j * C0], acc);
                }
            }
        } else {
            const int c0 = t - nF;
            for (int q = 0; q < rows * W0; q++) acc += sO[q * C0 + c0];
        }
        part[(int64_t)blockIdx.x * (nF + C0) + t] = acc;
    }
}
// Reference quirk (nmath.tcu:332-336): the per-tile _df[tap] is flushed once per thread whose
// load_id = ty*TS + tx equals tap, (tx,ty) in [0,16)^2, TS = (16-KS+S)/S.  Exactly one thread for
// KS = 1,3; but taps 12..15,24 of the 5x5/s1 config are flushed twice and taps 7..13 / 14..15 of
// the 4x4/s2 config two / three times.  Verified on the reference kernels (tests/golden); replicated.
__host__ __device__ __forceinline__ int dconv_flush_mult(int KS, int S, int tap) {
    const int TS = (16 - KS + S) / S;
    int m = 0;
    for (int ty = 0; ty < 16; ty++) { const int tx = tap - ty * TS; if (tx >= 0 && tx < 16) m++; }
    return m;
}
// the same + the optimizer step on the finished element (t4k_fused_opt_t): the arithmetic of k_optim_multi on DG = dF / dB, then DG = 0
struct WgOpt { bool mom; OptP p; float *G, *M, *V; int64_t offF, offB; float nwF, nwB; };
template<int KIND>
__global__ void __launch_bounds__(T4K_THREADS) k_wgrad_fin_opt(const float *__restrict__ part, float *dF, float *dB,
                                                               int nF, int C0, int nparts, int KS, int S, WgOpt o) {
    __shared__ float red[T4K_THREADS / 32];
    pdl_wait(); pdl_trigger();
    const int t = blockIdx.x;
    float v = 0.0f;
    for (int c = threadIdx.x; c < nparts; c += blockDim.x) v += part[(int64_t)c * (nF + C0) + t];
    v = block_sum(v, red);
    if (threadIdx.x == 0) {
        float dg, nw; int64_t j; float *d;
        if (t < nF) { d = dF + t; dg = *d; dg += v * (float)dconv_flush_mult(KS, S, (t / C0) % (KS * KS)); j = o.offF + t; nw = o.nwF; }
        else        { d = dB + (t - nF); dg = *d; dg += v; j = o.offB + (t - nF); nw = o.nwB; }
        float g = o.G[j], m = 0.0f, vv = 0.0f;
        if (KIND == 0) { dg = dg / nw; if (o.mom) m = o.M[j]; } else { m = o.M[j]; vv = o.V[j]; }
        opt_step<KIND>(g, dg, m, vv, 1.0f, o.mom, o.p);
        o.G[j] = g; *d = 0.0f;
        if (KIND == 0) { if (o.mom) o.M[j] = m; } else { o.M[j] = m; o.V[j] = vv; }
    }
}
// dF[t] += Σ_cta part[cta][t] (t < nF) ; dB[t-nF] += ... ; ordered → deterministic
__global__ void __launch_bounds__(T4K_THREADS) k_wgrad_fin(const float *__restrict__ part, float *dF, float *dB,
                                                           int nF, int C0, int nparts, int KS, int S) {
    __shared__ float red[T4K_THREADS / 32];
    pdl_wait(); pdl_trigger();                  // PDL: nothing global before this line
    const int t = blockIdx.x;                                      // one output element per CTA
    float v = 0.0f;
    for (int c = threadIdx.x; c < nparts; c += blockDim.x) v += part[(int64_t)c * (nF + C0) + t];
    v = block_sum(v, red);
    if (threadIdx.x == 0) {
        if (t < nF) dF[t] += v * (float)dconv_flush_mult(KS, S, (t / C0) % (KS * KS));
        else dB[t - nF] += v;
    }
}

// ====================================================================== general implicit GEMM (CUDA cores)
// MODE 0: fwd   O[m=(n,i,j), c0]   = B[c0] + Σ_{ky,kx,c1} I[n,i*S+ky-P,j*S+kx-P,c1] * F[c1,ky,kx,c0]
// MODE 1: dgrad dX[m=(n,y,x), c1]  = Σ_{ky,kx,c0} dO[n,(y+P-ky)/S,(x+P-kx)/S,c0] * F[c1,KS-1-ky,KS-1-kx,c0]
// MODE 2: wgrad part[z][m=(c1,ky,kx), c0] = Σ_{pix in split z} I[n,i*S+ky-P,j*S+kx-P,c1] * dO[pix,c0]
#define CBM 64
#define CBN 64
#define CBK 16
template<int MODE>
__global__ void __launch_bounds__(256) k_conv_igemm(ConvP p, float *part, int kchunk) {
    __shared__ float sA[2][CBK][CBM + 4];
    __shared__ float sB[2][CBK][CBN + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int KS = p.KS, S = p.S, P = p.P, C0 = p.C0, C1 = p.C1;
    int64_t Mg; int Ng; int64_t Kg;
    if (MODE == 0)      { Mg = (int64_t)p.N * p.H0 * p.W0; Ng = C0; Kg = (int64_t)KS * KS * C1; }
    else if (MODE == 1) { Mg = (int64_t)p.N * p.H1 * p.W1; Ng = C1; Kg = (int64_t)KS * KS * C0; }
    else                { Mg = (int64_t)C1 * KS * KS;      Ng = C0; Kg = (int64_t)p.N * p.H0 * p.W0; }
    const int64_t m0 = (int64_t)blockIdx.y * CBM; const int n0 = blockIdx.x * CBN;
    const int64_t kbeg = (MODE == 2) ? (int64_t)blockIdx.z * kchunk : 0;
    const int64_t kend = (MODE == 2) ? min(Kg, kbeg + kchunk) : Kg;

    // A gather: thread loads 4 elements: (m = tid/16 + 16*i, k = tid%16)  [k contiguous in NHWC c]
    // for MODE 2 the contiguous axis of A is m?  no: A(m=(c1,ky,kx), k=pix) → neither; keep same map.
    auto loadA = [&](int64_t m, int64_t k) -> float {
        if (m >= Mg || k >= kend) return 0.0f;
        if (MODE == 0) {
            const int j = (int)(m % p.W0); int64_t t = m / p.W0; const int i = (int)(t % p.H0); const int n = (int)(t / p.H0);
            const int c1 = (int)(k % C1); const int r = (int)(k / C1); const int kx = r % KS, ky = r / KS;
            const int gi = i * S + ky - P, gj = j * S + kx - P;
            if (gi < 0 || gi >= p.H1 || gj < 0 || gj >= p.W1) return 0.0f;
            return __ldg(p.I + (((int64_t)n * p.H1 + gi) * p.W1 + gj) * C1 + c1);
        } else if (MODE == 1) {
            const int x = (int)(m % p.W1); int64_t t = m / p.W1; const int y = (int)(t % p.H1); const int n = (int)(t / p.H1);
            const int c0 = (int)(k % C0); const int r = (int)(k / C0); const int kx = r % KS, ky = r / KS;
            const int ti = y + P - ky, tj = x + P - kx;
            if (ti < 0 || tj < 0 || (ti % S) || (tj % S)) return 0.0f;
            const int i = ti / S, j = tj / S;
            if (i >= p.H0 || j >= p.W0) return 0.0f;
            return __ldg(p.dO + (((int64_t)n * p.H0 + i) * p.W0 + j) * C0 + c0);
        } else {
            int r = (int)m; const int kx = r % KS; r /= KS; const int ky = r % KS; const int c1 = r / KS;
            const int j = (int)(k % p.W0); int64_t t = k / p.W0; const int i = (int)(t % p.H0); const int n = (int)(t / p.H0);
            const int gi = i * S + ky - P, gj = j * S + kx - P;
            if (gi < 0 || gi >= p.H1 || gj < 0 || gj >= p.W1) return 0.0f;
            return __ldg(p.I + (((int64_t)n * p.H1 + gi) * p.W1 + gj) * C1 + c1);
        }
    };
    auto loadB = [&](int64_t k, int n) -> float {
        if (n >= Ng || k >= kend) return 0.0f;
        if (MODE == 0) {
            const int c1 = (int)(k % C1); const int r = (int)(k / C1); const int kx = r % KS, ky = r / KS;
            return __ldg(p.F + (((int64_t)c1 * KS + ky) * KS + kx) * C0 + n);
        } else if (MODE == 1) {
            const int c0 = (int)(k % C0); const int r = (int)(k / C0); const int kx = r % KS, ky = r / KS;
            return __ldg(p.F + (((int64_t)n * KS + (KS - 1 - ky)) * KS + (KS - 1 - kx)) * C0 + c0);
        } else {
            return __ldg(p.dO + k * C0 + n);
        }
    };
    float ra[4], rb[4];
    auto load = [&](int64_t k0) {
        #pragma unroll
        for (int i = 0; i < 4; i++) {
            if (MODE == 2) ra[i] = loadA(m0 + (tid & 63), k0 + (tid >> 6) + 4 * i);       // pixels strided, m fast
            else           ra[i] = loadA(m0 + (tid >> 4) + 16 * i, k0 + (tid & 15));      // channel-contiguous k fast
            rb[i] = loadB(k0 + (tid >> 6) + 4 * i, n0 + (tid & 63));                      // n (channel) fast
        }
    };
    auto store = [&](int buf) {
        #pragma unroll
        for (int i = 0; i < 4; i++) {
            if (MODE == 2) sA[buf][(tid >> 6) + 4 * i][tid & 63] = ra[i];
            else           sA[buf][tid & 15][(tid >> 4) + 16 * i] = ra[i];
            sB[buf][(tid >> 6) + 4 * i][tid & 63] = rb[i];
        }
    };
    float acc[4][4] = {};
    int buf = 0;
    if (kbeg < kend) { load(kbeg); store(0); }
    __syncthreads();
    for (int64_t k0 = kbeg; k0 < kend; k0 += CBK) {
        const bool more = (k0 + CBK) < kend;
        if (more) load(k0 + CBK);
        #pragma unroll
        for (int k = 0; k < CBK; k++) {
            const float4 a4 = *reinterpret_cast<const float4*>(&sA[buf][k][ty * 4]);
            const float4 b4 = *reinterpret_cast<const float4*>(&sB[buf][k][tx * 4]);
            const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
            #pragma unroll
            for (int i = 0; i < 4; i++)
                #pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (more) store(buf ^ 1);
        __syncthreads();
        buf ^= 1;
    }
    #pragma unroll
    for (int i = 0; i < 4; i++) {
        const int64_t gm = m0 + ty * 4 + i;
        if (gm >= Mg) continue;
        #pragma unroll
        for (int j = 0; j < 4; j++) {
            const int gn = n0 + tx * 4 + j;
            if (gn >= Ng) continue;
            if (MODE == 0)      p.O[gm * C0 + gn] = acc[i][j] + __ldg(p.B + gn);
            else if (MODE == 1) p.dX[gm * C1 + gn] = acc[i][j];
            else                part[((int64_t)blockIdx.z * Mg + gm) * C0 + gn] = acc[i][j];
        }
    }
}
// dF[(c1,ky,kx),c0] += Σ_z part[z][...]   (MODE 2 finalize; part row m=(c1,ky,kx) matches dF's layout)
__global__ void __launch_bounds__(T4K_THREADS) k_igemm_wgrad_fin(const float *__restrict__ part, float *dF, int64_t nF, int splits, int C0, int KS, int S) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nF; t += (int64_t)gridDim.x * blockDim.x) {
        float s = 0.0f;
        for (int z = 0; z < splits; z++) s += part[(int64_t)z * nF + t];
        dF[t] += s * (float)dconv_flush_mult(KS, S, (int)((t / C0) % (KS * KS)));
    }
}
// dB[c0] += Σ_pix dO[pix,c0] : two-phase column reduce (CTA partials, ordered finalize)
__global__ void __launch_bounds__(T4K_THREADS) k_colsum_part(const float *__restrict__ X, float *part, int64_t rows, int C, int64_t rows_per) {
    // thread (c = tid % CP, r = tid / CP) with CP = smallest pow2 >= min(C,256)... keep simple: loop channels
    __shared__ float red[T4K_THREADS / 32];
    const int64_t r0 = (int64_t)blockIdx.x * rows_per, r1 = min(rows, r0 + rows_per);
    for (int c = 0; c < C; c++) {
        float v = 0.0f;
        for (int64_t r = r0 + threadIdx.x; r < r1; r += blockDim.x) v += __ldg(X + r * C + c);
        v = block_sum(v, red);
        if (threadIdx.x == 0) part[(int64_t)blockIdx.x * C + c] = v;
    }
}
__global__ void __launch_bounds__(T4K_THREADS) k_colsum_fin(const float *__restrict__ part, float *dB, int C, int nparts) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float s = 0.0f;
    for (int k = 0; k < nparts; k++) s += part[(int64_t)k * C + c];
    dB[c] += s;
}
// coalesced variant for C multiple of 4.. general: thread owns channel c = tid % C (C <= 256), strides rows
__global__ void __launch_bounds__(T4K_THREADS) k_colsum_part_c(const float *__restrict__ X, float *part, int64_t rows, int C, int64_t rows_per) {
    __shared__ float sm[T4K_THREADS];
    const int lanes_r = T4K_THREADS / C;                 // rows handled in parallel
    const int c = threadIdx.x % C, rr = threadIdx.x / C;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per, r1 = min(rows, r0 + rows_per);
    float v = 0.0f;
    if (rr < lanes_r) for (int64_t r = r0 + rr; r < r1; r += lanes_r) v += __ldg(X + r * C + c);
    sm[threadIdx.x] = (rr < lanes_r) ? v : 0.0f;
    __syncthreads();
    if (threadIdx.x < C) {
        float s = 0.0f;
        for (int k = 0; k < lanes_r; k++) s += sm[k * C + threadIdx.x];
        part[(int64_t)blockIdx.x * C + threadIdx.x] = s;
    }
}

bool conv_tc_ok(int H1, int W1, int CI, int H0, int W0, int CO, int KS, int S, int P);
int  conv_tc(const float *X, const float *F, const float *bias, float *Y, int N, int H, int W, int C1, int C0,
             int KS, int P, int mode, cudaStream_t st);
int  conv_wgrad_tc(const float *I, const float *dO, float *dF, float *dB, int N, int H, int W, int C1, int C0,
                   int KS, int P, cudaStream_t st);
bool conv_wgrad_tc_ok(int H1, int W1, int C1, int H0, int W0, int C0, int KS, int S, int P);
static int g_conv_engine = T4K_GEMM_AUTO;                 // t4k_set_conv_engine: AUTO / SIMT / TC

static bool conv_cfg_ok(int KS, int S, int P) {           // forward.cu:142-151
    return (KS == 1 && S == 1 && P == 0) || (KS == 3 && S == 1 && P == 1) ||
           (KS == 4 && S == 2 && P == 1) || (KS == 5 && S == 1 && P == 2);
}


// ====================================================================== fused conv → maxpool(2) → relu (→ flatten)
// The canonical block of the reference's CNN examples ("0.5 10 conv2d 2 maxpool relu [flatten]",
// examples/t4_40a.4th:11-12, t4_30e.4th:15-21).  One CTA per sample keeps the whole conv output of
// that sample in shared memory, so the block costs ONE launch each way and touches HBM once per
// layer tensor — every tensor the per-layer path writes is still written (n@ shows the same values).
//   forward : _fconv + _fpool + _factivate (+ flatten copy)      src/nn/forward.cu:83-155,201-228
//   backward: flatten copy + _bactivate + _bpool + _bconv         src/nn/backprop.cu:112-191,257-280
// Arithmetic order is identical to the per-layer kernels above for O/dX/pool routing (bit-equal);
// dF/dB use per-sample partials + the same ordered finalize.
struct CprP {
    const float *I, *F, *B;
    float *convO, *poolO, *actO, *actF, *flatO;     // forward outputs (flatO may be null)
    const float *dY;                                 // backward: gradient arriving at the block output
    float *Iio, *dXbuf, *part;                       // backward: conv input (overwritten with dX), grad[4] copy, wgrad partials
    int H1, W1, C1, H0, W0, C0, S, P, train;
};
template<int KS, int C0MAX>
__global__ void __launch_bounds__(T4K_THREADS) k_cpr_fwd(CprP p) {
    extern __shared__ float sm[];
    const int C0 = p.C0, C1 = p.C1, S = p.S, P = p.P, H0 = p.H0, W0 = p.W0, H1 = p.H1, W1 = p.W1;
    const int nF = C1 * KS * KS * C0, nI = H1 * W1 * C1, nO = H0 * W0 * C0;
    float *sF = sm, *sI = sm + ((nF + C0 + 3) & ~3), *sO = sI + ((nI + 3) & ~3);
    const int n = blockIdx.x;
    const float *gI = p.I + (int64_t)n * nI;
    for (int t = threadIdx.x; t < nF; t += blockDim.x) sF[t] = __ldg(p.F + t);
    for (int t = threadIdx.x; t < C0; t += blockDim.x) sF[nF + t] = __ldg(p.B + t);
    for (int t = threadIdx.x; t < nI; t += blockDim.x) sI[t] = __ldg(gI + t);
    __syncthreads();
    for (int pix = threadIdx.x; pix < H0 * W0; pix += blockDim.x) {
        const int j = pix % W0, i = pix / W0;
        float acc[C0MAX];
        #pragma unroll
        for (int c = 0; c < C0MAX; c++) acc[c] = (c < C0) ? sF[nF + c] : 0.0f;
        #pragma unroll
        for (int y = 0; y < KS; y++) {
            const int gi = i * S + y - P;
            if (gi < 0 || gi >= H1) continue;
            #pragma unroll
            for (int x = 0; x < KS; x++) {
                const int gj = j * S + x - P;
                if (gj < 0 || gj >= W1) continue;
                const float *px = sI + (W1 * gi + gj) * C1;
                for (int c1 = 0; c1 < C1; c1++) {
                    const float v = px[c1];
                    const float *f = sF + ((c1 * KS + y) * KS + x) * C0;
                    #pragma unroll
                    for (int c = 0; c < C0MAX; c++) if (c < C0) acc[c] = fmaf(f[c], v, acc[c]);
                }
            }
        }
        #pragma unroll
        for (int c = 0; c < C0MAX; c++) if (c < C0) sO[pix * C0 + c] = acc[c];
    }
    __syncthreads();
    float *gO = p.convO + (int64_t)n * nO;
    if (((nO & 3) == 0) && aligned16(p.convO)) { for (int t = threadIdx.x; t < (nO >> 2); t += blockDim.x) stg4(gO + 4 * t, *reinterpret_cast<const float4*>(sO + 4 * t)); }
    else for (int t = threadIdx.x; t < nO; t += blockDim.x) gO[t] = sO[t];
    const int Hp = H0 / 2, Wp = W0 / 2, nP = Hp * Wp * C0;
    const int64_t gp = (int64_t)n * nP;
    for (int t = threadIdx.x; t < nP; t += blockDim.x) {
        const int c = t % C0; int r = t / C0; const int j0 = r % Wp, i0 = r / Wp;
        const float *ix = sO + ((i0 * 2) * W0 + j0 * 2) * C0 + c;
        float v = ix[0];
        v = fmaxf(ix[C0], v); v = fmaxf(ix[W0 * C0], v); v = fmaxf(ix[(W0 + 1) * C0], v);      // k_pool<2> order
        p.poolO[gp + t] = v;
        float o, f;
        if (v > 0.0f) { f = 1.0f; o = v; } else { f = 0.0f; o = 0.0f; }                       // k_activate RELU
        p.actO[gp + t] = o; p.actF[gp + t] = f;
        if (p.flatO) p.flatO[gp + t] = o;
    }
}
template<int KS>
__global__ void __launch_bounds__(T4K_THREADS) k_cpr_bwd(CprP p) {
    extern __shared__ float sm[];
    __shared__ float sred[2 * 512];
    const int C0 = p.C0, C1 = p.C1, S = p.S, P = p.P, H0 = p.H0, W0 = p.W0, H1 = p.H1, W1 = p.W1;
    const int nF = C1 * KS * KS * C0, nI = H1 * W1 * C1, nO = H0 * W0 * C0;
    float *sF = sm, *sI = sm + ((nF + 3) & ~3), *sO = sI + ((nI + 3) & ~3);
    const int n = blockIdx.x;
    float *gO = p.convO + (int64_t)n * nO;
    float *gI = p.Iio + (int64_t)n * nI;
    for (int t = threadIdx.x; t < nF; t += blockDim.x) sF[t] = __ldg(p.F + t);
    for (int t = threadIdx.x; t < nI; t += blockDim.x) sI[t] = gI[t];
    if (((nO & 3) == 0) && aligned16(p.convO)) { for (int t = threadIdx.x; t < (nO >> 2); t += blockDim.x) *reinterpret_cast<float4*>(sO + 4 * t) = *reinterpret_cast<const float4*>(gO + 4 * t); }
    else for (int t = threadIdx.x; t < nO; t += blockDim.x) sO[t] = gO[t];
    __syncthreads();
    // flatten copy, relu backward, max-pool routing (first strict max in y,x order), in shared memory
    const int Hp = H0 / 2, Wp = W0 / 2, nP = Hp * Wp * C0;
    const int64_t gp = (int64_t)n * nP;
    for (int t = threadIdx.x; t < nP; t += blockDim.x) {
        const float d = p.dY[gp + t];
        if (p.actO != p.dY) p.actO[gp + t] = d;                      // flatten backward: in = out
        const float g = __fmul_rn(d, p.actF[gp + t]);                // _bactivate: in = out * mask
        p.poolO[gp + t] = g;
        const int c = t % C0; int r = t / C0; const int j0 = r % Wp, i0 = r / Wp;
        float *ix = sO + ((i0 * 2) * W0 + j0 * 2) * C0 + c;
        const float t0 = ix[0], t1 = ix[C0], t2 = ix[W0 * C0], t3 = ix[(W0 + 1) * C0];
        float best = t0; int arg = 0;
        if (t1 > best) { best = t1; arg = 1; }
        if (t2 > best) { best = t2; arg = 2; }
        if (t3 > best) { best = t3; arg = 3; }
        ix[0] = (arg == 0) ? g : 0.0f; ix[C0] = (arg == 1) ? g : 0.0f;
        ix[W0 * C0] = (arg == 2) ? g : 0.0f; ix[(W0 + 1) * C0] = (arg == 3) ? g : 0.0f;
    }
    __syncthreads();
    if (((nO & 3) == 0) && aligned16(p.convO)) { for (int t = threadIdx.x; t < (nO >> 2); t += blockDim.x) stg4(gO + 4 * t, *reinterpret_cast<const float4*>(sO + 4 * t)); }
    else for (int t = threadIdx.x; t < nO; t += blockDim.x) gO[t] = sO[t];
    // weight / bias gradient partials of this sample: 2 threads per element (output-row halves)
    if (p.train) {
        const int nE = nF + C0;
        for (int u = threadIdx.x; u < 2 * nE; u += blockDim.x) {
            const int half = u / nE, t = u - half * nE;
            const int ia = half ? H0 / 2 : 0, ib = half ? H0 : H0 / 2;
            float acc = 0.0f;
            if (t < nF) {
                const int c0 = t % C0; int r = t / C0;
                const int kx = r % KS; r /= KS; const int ky = r % KS; const int c1 = r / KS;
                for (int i = ia; i < ib; i++) {
                    const int gi = i * S + ky - P;
                    if (gi < 0 || gi >= H1) continue;
                    const float *irow = sI + gi * W1 * C1 + c1;
                    const float *orow = sO + i * W0 * C0 + c0;
                    for (int j = 0; j < W0; j++) {
                        const int gj = j * S + kx - P;
                        if (gj >= 0 && gj < W1) acc = fmaf(irow[gj * C1], orow[j * C0], acc);
                    }
                }
            } else {
                const int c0 = t - nF;
                for (int q = ia * W0; q < ib * W0; q++) acc += sO[q * C0 + c0];
            }
            sred[u] = acc;
        }
        __syncthreads();
        for (int t = threadIdx.x; t < nE; t += blockDim.x) p.part[(int64_t)n * nE + t] = sred[t] + sred[nE + t];
    }
    // input gradient (flipped taps, k_conv_dgrad_small order) → conv input tensor and grad[4]
    for (int pix = threadIdx.x; pix < H1 * W1; pix += blockDim.x) {
        const int x = pix % W1, y = pix / W1;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        #pragma unroll
        for (int ky = 0; ky < KS; ky++) {
            const int ti = y + P - ky;
            if (ti < 0 || (ti % S) != 0) continue;
            const int i = ti / S;
            if (i >= H0) continue;
            #pragma unroll
            for (int kx = 0; kx < KS; kx++) {
                const int tj = x + P - kx;
                if (tj < 0 || (tj % S) != 0) continue;
                const int j = tj / S;
                if (j >= W0) continue;
                const float *d = sO + (W0 * i + j) * C0;
                for (int c0 = 0; c0 < C0; c0++) {
                    const float dv = d[c0];
                    #pragma unroll
                    for (int c1 = 0; c1 < 4; c1++)
                        if (c1 < C1) acc[c1] = fmaf(sF[((c1 * KS + (KS - 1 - ky)) * KS + (KS - 1 - kx)) * C0 + c0], dv, acc[c1]);
                }
            }
        }
        #pragma unroll
        for (int c1 = 0; c1 < 4; c1++) if (c1 < C1) { gI[pix * C1 + c1] = acc[c1]; p.dXbuf[(int64_t)n * nI + pix * C1 + c1] = acc[c1]; }
    }
}
static bool cpr_ok(int H1, int W1, int C1, int H0, int W0, int C0, int KS, int S, int P, size_t *smem) {
    if (!conv_cfg_ok(KS, S, P) || C1 > 4 || C0 > 16 || (H0 & 1) || (W0 & 1)) return false;
    const size_t nF = (size_t)C1 * KS * KS * C0, nI = (size_t)H1 * W1 * C1, nO = (size_t)H0 * W0 * C0;
    if (2 * (nF + C0) > 1024) return false;
    *smem = (((nF + C0 + 3) & ~(size_t)3) + ((nI + 3) & ~(size_t)3) + nO) * sizeof(float);
    return *smem <= 100 * 1024;
}
} // namespace t4k
using namespace t4k;

// ====================================================================== C ABI
extern "C" int t4k_set_conv_engine(int engine) {
    if (engine < T4K_GEMM_AUTO || engine > T4K_GEMM_TC) return T4K_EINVAL;
    g_conv_engine = engine; return 0;
}
extern "C" int t4k_conv2d_fwd(const float *I, const float *F, const float *B, float *O,
                              int N, int H1, int W1, int C1, int H0, int W0, int C0, int KS, int S, int P, t4k_stream_t s) {
    if (!I || !F || !B || !O || N < 1 || H1 < 1 || W1 < 1 || C1 < 1 || H0 < 1 || W0 < 1 || C0 < 1) return T4K_EINVAL;
    if (!conv_cfg_ok(KS, S, P)) return T4K_ENOSUP;
    ConvP p{I, F, B, nullptr, O, nullptr, nullptr, nullptr, N, H1, W1, C1, H0, W0, C0, KS, S, P};
    cudaStream_t st = STRM(s);
    const int64_t npix = (int64_t)N * H0 * W0;
    const size_t fbytes = ((size_t)C1 * KS * KS * C0 + C0) * sizeof(float);
    if (C0 <= 16 && C1 <= 8 && fbytes <= 40 * 1024) {
        const int g = stream_grid(npix);
        switch (KS) {
        case 1: k_conv_fwd_small<1, 16><<<g, T4K_THREADS, fbytes, st>>>(p); break;
        case 3: k_conv_fwd_small<3, 16><<<g, T4K_THREADS, fbytes, st>>>(p); break;
        case 4: k_conv_fwd_small<4, 16><<<g, T4K_THREADS, fbytes, st>>>(p); break;
        default: k_conv_fwd_small<5, 16><<<g, T4K_THREADS, fbytes, st>>>(p); break;
        }
        return check_launch();
    }
    if (g_conv_engine != T4K_GEMM_SIMT && conv_tc_ok(H1, W1, C1, H0, W0, C0, KS, S, P))
        return conv_tc(I, F, B, O, N, H1, W1, C1, C0, KS, P, 0, st);
    if (g_conv_engine == T4K_GEMM_TC) return T4K_EINVAL;
    // gridDim.y <= 65535: launch per group of samples so each launch has at most 65535 pixel tiles
    const int64_t pix_per_n = (int64_t)H0 * W0;
    int n_per = (int)((65535LL * CBM) / pix_per_n); if (n_per < 1) return T4K_EINVAL;
    if (n_per > N) n_per = N;
    for (int n0 = 0; n0 < N; n0 += n_per) {
        ConvP r = p; r.N = (N - n0 < n_per) ? N - n0 : n_per;
        r.I = I + (int64_t)n0 * H1 * W1 * C1; r.O = O + (int64_t)n0 * H0 * W0 * C0;
        dim3 g((C0 + CBN - 1) / CBN, (unsigned)(((int64_t)r.N * pix_per_n + CBM - 1) / CBM));
        k_conv_igemm<0><<<g, 256, 0, st>>>(r, nullptr, 0);
        int rc = check_launch(); if (rc) return rc;
    }
    return 0;
}

// the two halves of k_dconv2d: parameter gradients (train: dF += , dB += ; reads I and dO) and input gradient (dX: reads dO and F; skipped when
// dX == nullptr).  t4k_conv2d_bwd runs both; the conv-transpose layer (t4k_dconv2d_*) takes one half at a time with the roles swapped.
static int conv_bwd_impl(const float *I, const float *dO, const float *F, float *dX, float *dF, float *dB,
                         int N, int H1, int W1, int C1, int H0, int W0, int C0, int KS, int S, int P, int train, cudaStream_t st) {
    ConvP p{I, F, nullptr, dO, nullptr, dX, dF, dB, N, H1, W1, C1, H0, W0, C0, KS, S, P};
    const int nF = C1 * KS * KS * C0;
    int rc;
    // ---- weight + bias gradient first (dX may alias nothing, but keep I intact until wgrad has read it)
    if (train) {
        const size_t smem_small = ((size_t)WG_ROWS * W0 * C0 + (size_t)((WG_ROWS - 1) * S + KS) * W1 * C1) * sizeof(float);
        if (g_conv_engine != T4K_GEMM_SIMT && conv_wgrad_tc_ok(H1, W1, C1, H0, W0, C0, KS, S, P)) {
            rc = conv_wgrad_tc(I, dO, dF, dB, N, H1, W1, C1, C0, KS, P, st);
            if (rc) return rc;
        } else if (nF + C0 <= 4096 && smem_small <= 96 * 1024 && C0 <= 16) {
            const int strips = (H0 + WG_ROWS - 1) / WG_ROWS;
            const int ctas = N * strips;
            float *part = (float*)workspace((size_t)ctas * (nF + C0) * sizeof(float), 4);
            if (!part) return T4K_ENOMEM;
            static DevFlag attr[6];
            #define WG_LAUNCH(K_) { if (smem_small > 48 * 1024 && dev_first(attr[K_])) { cudaFuncSetAttribute(k_conv_wgrad_small<K_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024); } \
                                    k_conv_wgrad_small<K_><<<ctas, T4K_THREADS, smem_small, st>>>(p, part, strips); }
            switch (KS) { case 1: WG_LAUNCH(1) break; case 3: WG_LAUNCH(3) break; case 4: WG_LAUNCH(4) break; default: WG_LAUNCH(5) break; }
            rc = check_launch(); if (rc) return rc;
            launch_pdl(k_wgrad_fin, dim3(nF + C0), dim3(T4K_THREADS), 0, st, part, dF, dB, nF, C0, ctas, KS, S);
            rc = check_launch(); if (rc) return rc;
        } else {
            const int64_t Kg = (int64_t)N * H0 * W0;
            const int64_t Mg = (int64_t)C1 * KS * KS;
            const int gx = (C0 + CBN - 1) / CBN, gy = (int)((Mg + CBM - 1) / CBM);
            int splits = (int)((2 * sm_count() + gx * gy - 1) / (gx * gy));
            if ((int64_t)splits * 4 * CBK > Kg) splits = (int)((Kg + 4 * CBK - 1) / (4 * CBK));
            if (splits < 1) splits = 1;
            if (splits > 1024) splits = 1024;
            int64_t kchunk = (Kg + splits - 1) / splits; kchunk = (kchunk + CBK - 1) / CBK * CBK;
            splits = (int)((Kg + kchunk - 1) / kchunk);
            float *part = (float*)workspace((size_t)splits * nF * sizeof(float), 4);
            if (!part) return T4K_ENOMEM;
            k_conv_igemm<2><<<dim3(gx, gy, splits), 256, 0, st>>>(p, part, (int)kchunk);
            rc = check_launch(); if (rc) return rc;
            k_igemm_wgrad_fin<<<stream_grid(nF), T4K_THREADS, 0, st>>>(part, dF, nF, splits, C0, KS, S);
            rc = check_launch(); if (rc) return rc;
            // dB
            int nparts = 4 * sm_count();
            int64_t rows_per = (Kg + nparts - 1) / nparts; if (rows_per < 1) rows_per = 1;
            nparts = (int)((Kg + rows_per - 1) / rows_per);
            float *bp = (float*)workspace((size_t)nparts * C0 * sizeof(float), 5);
            if (!bp) return T4K_ENOMEM;
            if (C0 <= T4K_THREADS) k_colsum_part_c<<<nparts, T4K_THREADS, 0, st>>>(dO, bp, Kg, C0, rows_per);
            else                   k_colsum_part<<<nparts, T4K_THREADS, 0, st>>>(dO, bp, Kg, C0, rows_per);
            rc = check_launch(); if (rc) return rc;
            k_colsum_fin<<<(C0 + T4K_THREADS - 1) / T4K_THREADS, T4K_THREADS, 0, st>>>(bp, dB, C0, nparts);
            rc = check_launch(); if (rc) return rc;
        }
    }
    if (!dX) return 0;
    // ---- input gradient (flipped taps)
    const int64_t npix = (int64_t)N * H1 * W1;
    const size_t fbytes = (size_t)nF * sizeof(float);
    if (C1 <= 4 && C0 <= 64 && fbytes <= 40 * 1024) {
        const int g = stream_grid(npix);
        switch (KS) {
        case 1: k_conv_dgrad_small<1, 4><<<g, T4K_THREADS, fbytes, st>>>(p); break;
        case 3: k_conv_dgrad_small<3, 4><<<g, T4K_THREADS, fbytes, st>>>(p); break;
        case 4: k_conv_dgrad_small<4, 4><<<g, T4K_THREADS, fbytes, st>>>(p); break;
        default: k_conv_dgrad_small<5, 4><<<g, T4K_THREADS, fbytes, st>>>(p); break;
        }
        return check_launch();
    }
    if (g_conv_engine != T4K_GEMM_SIMT && conv_tc_ok(H0, W0, C0, H1, W1, C1, KS, S, P))
        return conv_tc(dO, F, nullptr, dX, N, H1, W1, C1, C0, KS, P, 1, st);
    const int64_t pix_per_n = (int64_t)H1 * W1;
    int n_per = (int)((65535LL * CBM) / pix_per_n); if (n_per < 1) return T4K_EINVAL;
    if (n_per > N) n_per = N;
    for (int n0 = 0; n0 < N; n0 += n_per) {
        ConvP r = p; r.N = (N - n0 < n_per) ? N - n0 : n_per;
        r.dO = dO + (int64_t)n0 * H0 * W0 * C0; r.dX = dX + (int64_t)n0 * H1 * W1 * C1;
        dim3 g((C1 + CBN - 1) / CBN, (unsigned)(((int64_t)r.N * pix_per_n + CBM - 1) / CBM));
        k_conv_igemm<1><<<g, 256, 0, st>>>(r, nullptr, 0);
        rc = check_launch(); if (rc) return rc;
    }
    return 0;
}

extern "C" int t4k_conv2d_bwd(const float *I, const float *dO, const float *F, float *dX, float *dF, float *dB,
                              int N, int H1, int W1, int C1, int H0, int W0, int C0, int KS, int S, int P,
                              int train, t4k_stream_t s) {
    if (!I || !dO || !F || !dX || N < 1 || H1 < 1 || W1 < 1 || C1 < 1 || H0 < 1 || W0 < 1 || C0 < 1) return T4K_EINVAL;
    if (train && (!dF || !dB)) return T4K_EINVAL;
    if (!conv_cfg_ok(KS, S, P)) return T4K_ENOSUP;
    return conv_bwd_impl(I, dO, F, dX, dF, dB, N, H1, W1, C1, H0, W0, C0, KS, S, P, train, STRM(s));
}

// ---- conv-transpose layer (L_DCONV, `dconv2d`: 4x4, stride 2): the reference wires it as the convolution layer with the two kernels' roles
// swapped — forward = k_dconv2d's input-gradient half, backward = k_conv2d (src/nn/forward.cu:110, backprop.cu:137; shapes model.cpp:129-133).
// Layer input I [N,H1,W1,C1] (small), output O [N,H0,W0,C0] (large); F is the filter [C0][K][K][C1] of the convolution (C0 -> C1, K, S, P) that maps
// the large image onto the small one.
extern "C" int t4k_dconv2d_fwd(const float *I, const float *F, const float *B, float *O,
                               int N, int H1, int W1, int C1, int H0, int W0, int C0, int KS, int S, int P, t4k_stream_t s) {
    if (!I || !F || !B || !O || N < 1 || H1 < 1 || W1 < 1 || C1 < 1 || H0 < 1 || W0 < 1 || C0 < 1) return T4K_EINVAL;
    if (!conv_cfg_ok(KS, S, P)) return T4K_ENOSUP;
    if ((H0 - KS + 2 * P) / S + 1 != H1 || (W0 - KS + 2 * P) / S + 1 != W1) return T4K_EINVAL;
    // O = "dX" of the convolution whose output gradient is I (flipped taps, as k_dconv2d computes it), then + bias per output channel
    int rc = conv_bwd_impl(nullptr, I, F, O, nullptr, nullptr, N, H0, W0, C0, H1, W1, C1, KS, S, P, 0, STRM(s));
    if (rc) return rc;
    return t4k_bias(B, O, (int)((int64_t)N * H0 * W0), C0, s);
}
extern "C" int t4k_dconv2d_bwd(const float *I, const float *dO, const float *F, float *dX, float *dF, float *dB,
                               int N, int H1, int W1, int C1, int H0, int W0, int C0, int KS, int S, int P, int train, t4k_stream_t s) {
    if (!I || !dO || !F || !dX || N < 1 || H1 < 1 || W1 < 1 || C1 < 1 || H0 < 1 || W0 < 1 || C0 < 1 || I == dX) return T4K_EINVAL;
    if (train && (!dF || !dB)) return T4K_EINVAL;
    if (!conv_cfg_ok(KS, S, P)) return T4K_ENOSUP;
    if ((H0 - KS + 2 * P) / S + 1 != H1 || (W0 - KS + 2 * P) / S + 1 != W1) return T4K_EINVAL;
    cudaStream_t st = STRM(s);
    int rc;
    float *zb = (float*)workspace((size_t)(2 * C1 + 8) * sizeof(float), 3);        // [C1] zero bias | [C1] the convolution's own dB (discarded)
    if (!zb) return T4K_ENOMEM;
    if (cudaMemsetAsync(zb, 0, (size_t)2 * C1 * sizeof(float), st) != cudaSuccess) return (int)cudaGetLastError();
    if (train) {
        // dF += the convolution's filter gradient with (input, output gradient) = (dO, I): k_dconv2d's parameter half
        rc = conv_bwd_impl(dO, I, F, nullptr, dF, zb + C1, N, H0, W0, C0, H1, W1, C1, KS, S, P, 1, st);
        if (rc) return rc;
        // dB[c0] += sum over the output pixels of dO (the bias is added per output channel in the forward)
        const int64_t rows = (int64_t)N * H0 * W0;
        int nparts = 4 * sm_count();
        int64_t rows_per = (rows + nparts - 1) / nparts; if (rows_per < 1) rows_per = 1;
        nparts = (int)((rows + rows_per - 1) / rows_per);
        float *bp = (float*)workspace((size_t)nparts * C0 * sizeof(float), 5);
        if (!bp) return T4K_ENOMEM;
        if (C0 <= T4K_THREADS) k_colsum_part_c<<<nparts, T4K_THREADS, 0, st>>>(dO, bp, rows, C0, rows_per);
        else                   k_colsum_part<<<nparts, T4K_THREADS, 0, st>>>(dO, bp, rows, C0, rows_per);
        rc = check_launch(); if (rc) return rc;
        k_colsum_fin<<<(C0 + T4K_THREADS - 1) / T4K_THREADS, T4K_THREADS, 0, st>>>(bp, dB, C0, nparts);
        rc = check_launch(); if (rc) return rc;
    }
    // dX = k_conv2d(dO) with the same filter, no bias
    return t4k_conv2d_fwd(dO, F, zb, dX, N, H0, W0, C0, H1, W1, C1, KS, S, P, s);
}

// ---- fused conv → maxpool(2) → relu (→ flatten) block; T4K_ENOSUP when the shape is not eligible (caller: per-layer calls)
namespace t4k {
int cpr_v1_fwd(const float *I, const float *F, const float *B, float *convO, float *poolO, float *actO, float *actF,
               float *flatO, int N, int H1, int W1, int C1, int H0, int W0, int C0, int KS, int S, int P, cudaStream_t s) {
    if (!I || !F || !B || !convO || !poolO || !actO || !actF || N < 1) return T4K_EINVAL;
    size_t smem = 0;
    if (!cpr_ok(H1, W1, C1, H0, W0, C0, KS, S, P, &smem)) return T4K_ENOSUP;
    CprP p{}; p.I = I; p.F = F; p.B = B; p.convO = convO; p.poolO = poolO; p.actO = actO; p.actF = actF; p.flatO = flatO;
    p.H1 = H1; p.W1 = W1; p.C1 = C1; p.H0 = H0; p.W0 = W0; p.C0 = C0; p.S = S; p.P = P;
    static DevFlag attr[6];
    #define CPRF(K_) { if (smem > 48 * 1024 && dev_first(attr[K_])) { cudaFuncSetAttribute(k_cpr_fwd<K_, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024); } \
                       k_cpr_fwd<K_, 16><<<N, T4K_THREADS, smem, s>>>(p); }
    switch (KS) { case 1: CPRF(1) break; case 3: CPRF(3) break; case 4: CPRF(4) break; default: CPRF(5) break; }
    return check_launch();
}
int cpr_v1_bwd(const float *dY, float *actO, const float *actF, float *poolO, float *convO, float *Iio, float *dXbuf,
               const float *F, float *dF, float *dB, int N, int H1, int W1, int C1, int H0, int W0, int C0,
               int KS, int S, int P, int train, cudaStream_t s) {
    if (!dY || !actO || !actF || !poolO || !convO || !Iio || !dXbuf || !F || N < 1 || (train && (!dF || !dB))) return T4K_EINVAL;
    size_t smem = 0;
    if (!cpr_ok(H1, W1, C1, H0, W0, C0, KS, S, P, &smem)) return T4K_ENOSUP;
    const int nF = C1 * KS * KS * C0;
    CprP p{}; p.F = F; p.convO = convO; p.poolO = poolO; p.actO = actO; p.actF = (float*)actF; p.dY = dY; p.Iio = Iio; p.dXbuf = dXbuf;
    p.H1 = H1; p.W1 = W1; p.C1 = C1; p.H0 = H0; p.W0 = W0; p.C0 = C0; p.S = S; p.P = P; p.train = train;
    if (train) { p.part = (float*)workspace((size_t)N * (nF + C0) * sizeof(float), 4); if (!p.part) return T4K_ENOMEM; }
    static DevFlag attr[6];
    #define CPRB(K_) { if (smem > 40 * 1024 && dev_first(attr[K_])) { cudaFuncSetAttribute(k_cpr_bwd<K_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024); } \
                       k_cpr_bwd<K_><<<N, T4K_THREADS, smem, s>>>(p); }
    switch (KS) { case 1: CPRB(1) break; case 3: CPRB(3) break; case 4: CPRB(4) break; default: CPRB(5) break; }
    int rc = check_launch(); if (rc || !train) return rc;
    launch_pdl(k_wgrad_fin, dim3(nF + C0), dim3(T4K_THREADS), 0, s, p.part, dF, dB, nF, C0, N, KS, S);
    return check_launch();
}
int wgrad_fin_opt_launch(const float *part, float *dF, float *dB, int nF, int C0, int nparts, int KS, int S, const t4k_fused_opt_t *opt, cudaStream_t st) {
    if (!opt->G || opt->kind < 0 || opt->kind > 2 || (opt->kind && (!opt->M || !opt->V))) return T4K_EINVAL;
    WgOpt o{!(fabsf(opt->b1) < DU_EPS), OptP{opt->lr, opt->b1, opt->b2, opt->wd}, opt->G, opt->M, opt->V, opt->offF, opt->offB, (float)opt->NwF, (float)opt->NwB};
    if (opt->kind == 0 && o.mom && !opt->M) return T4K_EINVAL;
    if (opt->kind == 0)      launch_pdl(k_wgrad_fin_opt<0>, dim3(nF + C0), dim3(T4K_THREADS), 0, st, part, dF, dB, nF, C0, nparts, KS, S, o);
    else if (opt->kind == 1) launch_pdl(k_wgrad_fin_opt<1>, dim3(nF + C0), dim3(T4K_THREADS), 0, st, part, dF, dB, nF, C0, nparts, KS, S, o);
    else                     launch_pdl(k_wgrad_fin_opt<2>, dim3(nF + C0), dim3(T4K_THREADS), 0, st, part, dF, dB, nF, C0, nparts, KS, S, o);
    return check_launch();
}
int wgrad_fin_launch(const float *part, float *dF, float *dB, int nF, int C0, int nparts, int KS, int S, cudaStream_t st) {
    launch_pdl(k_wgrad_fin, dim3(nF + C0), dim3(T4K_THREADS), 0, st, part, dF, dB, nF, C0, nparts, KS, S);
    return check_launch();
}
} // namespace t4k
