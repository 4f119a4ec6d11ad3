// pool_bn.cu — pooling fwd/bwd and batch-norm fwd/bwd, NHWC (channels innermost → coalesced)
//   replaces k_pool<KS>/k_dpool<KS> (src/nn/nmath.tcu:122-186,475-568; FORK4P launch :110-120) and
//   k_batchnorm_1/2/3, k_dbatchnorm_1/2/3 (src/nn/nmath.cu:177-264,295-414).
// All HBM-bound.  The reference's batch-norm statistics kernels read with stride C (one block per
// (c,n)); here one pass reads whole NHWC rows (C contiguous) and reduces per channel in smem, then
// an ordered finalize over CTA partials (deterministic; the reference uses atomicAdd per block).
#include "common.cuh"

namespace t4k {

// ------------------------------------------------------------------ pooling forward
// thread = one output element (n, i0, j0, c), c fastest → coalesced reads of KS*KS rows of C floats
template<int KS>
__global__ void __launch_bounds__(T4K_THREADS)
k_pool(int layer, const float *__restrict__ I, float *__restrict__ O, int H1, int W1, int H0, int W0, int C, int64_t total) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(t % C); int64_t r = t / C;
        const int j0 = (int)(r % W0); r /= W0;
        const int i0 = (int)(r % H0); const int64_t n = r / H0;
        const float *ix = I + ((n * H1 + (int64_t)i0 * KS) * W1 + (int64_t)j0 * KS) * C + c;
        float tile[KS * KS];
        #pragma unroll
        for (int y = 0; y < KS; y++)
            #pragma unroll
            for (int x = 0; x < KS; x++) tile[y * KS + x] = __ldg(ix + ((int64_t)y * W1 + x) * C);
        float v;
        if (layer == T4K_L_MAXPOOL) { v = tile[0];
            #pragma unroll
            for (int k = 1; k < KS * KS; k++) v = fmaxf(tile[k], v); }
        else if (layer == T4K_L_MINPOOL) { v = tile[0];
            #pragma unroll
            for (int k = 1; k < KS * KS; k++) v = fminf(tile[k], v); }
        else { v = 0.0f;                                   // AVGPOOL and USAMPLE(-backward): Σ / KS²
            #pragma unroll
            for (int k = 0; k < KS * KS; k++) v += tile[k];
            v /= (float)(KS * KS); }
        O[t] = v;
    }
}
// ------------------------------------------------------------------ pooling backward, IN PLACE on the forward input
template<int KS>
__global__ void __launch_bounds__(T4K_THREADS)
k_dpool(int layer, float *I, const float *__restrict__ dO, int H1, int W1, int H0, int W0, int C, int64_t total) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(t % C); int64_t r = t / C;
        const int j0 = (int)(r % W0); r /= W0;
        const int i0 = (int)(r % H0); const int64_t n = r / H0;
        float *ix = I + ((n * H1 + (int64_t)i0 * KS) * W1 + (int64_t)j0 * KS) * C + c;
        const float d = __ldg(dO + t);
        if (layer == T4K_L_AVGPOOL || layer == T4K_L_USAMPLE) {
            const float v = (layer == T4K_L_AVGPOOL) ? d / (float)(KS * KS) : d;
            #pragma unroll
            for (int y = 0; y < KS; y++)
                #pragma unroll
                for (int x = 0; x < KS; x++) ix[((int64_t)y * W1 + x) * C] = v;
        } else {
            // first strict max/min in (y,x) scan order, `dx > best` with best = tile[0] (nmath.tcu:535-549)
            float tile[KS * KS];
            #pragma unroll
            for (int y = 0; y < KS; y++)
                #pragma unroll
                for (int x = 0; x < KS; x++) tile[y * KS + x] = ix[((int64_t)y * W1 + x) * C];
            float best = tile[0]; int arg = 0;
            #pragma unroll
            for (int k = 1; k < KS * KS; k++) {
                const bool better = (layer == T4K_L_MAXPOOL) ? (tile[k] > best) : (tile[k] < best);
                if (better) { best = tile[k]; arg = k; }
            }
            #pragma unroll
            for (int y = 0; y < KS; y++)
                #pragma unroll
                for (int x = 0; x < KS; x++) ix[((int64_t)y * W1 + x) * C] = (y * KS + x == arg) ? d : 0.0f;
        }
    }
}

// ------------------------------------------------------------------ batch norm
// per-channel column sums over NHW rows of C floats.  CTA = rows_per consecutive rows; thread owns
// channel c = tid % CP and row lane tid / CP.  TWO sums per channel in one pass.
//   MODE 0: (Σx, Σx²)        MODE 1: (Σdy, Σdy·x̂)
template<int MODE>
__global__ void __launch_bounds__(T4K_THREADS)
k_bn_colsum2(const float *__restrict__ X, const float *__restrict__ Y, float *part, int64_t rows, int C, int64_t rows_per) {
    __shared__ float s0[T4K_THREADS], s1[T4K_THREADS];
    const int64_t r0 = (int64_t)blockIdx.x * rows_per, r1 = min(rows, r0 + rows_per);
    for (int cb = 0; cb < C; cb += T4K_THREADS) {                  // channel blocks (C > 256 loops)
        const int cw = min(C - cb, T4K_THREADS);
        const int lanes_r = T4K_THREADS / cw;
        const int c = threadIdx.x % cw, rr = threadIdx.x / cw;
        float a = 0.0f, b = 0.0f;
        if (rr < lanes_r) {
            for (int64_t r = r0 + rr; r < r1; r += lanes_r) {
                const float x = __ldg(X + r * C + cb + c);
                if (MODE == 0) { a += x; b += x * x; }
                else           { const float y = __ldg(Y + r * C + cb + c); a += x; b += x * y; }
            }
        }
        s0[threadIdx.x] = a; s1[threadIdx.x] = b;
        __syncthreads();
        if (threadIdx.x < cw) {
            float sa = 0.0f, sb = 0.0f;
            for (int k = 0; k < lanes_r; k++) { sa += s0[k * cw + threadIdx.x]; sb += s1[k * cw + threadIdx.x]; }
            part[((int64_t)blockIdx.x * 2 + 0) * C + cb + threadIdx.x] = sa;
            part[((int64_t)blockIdx.x * 2 + 1) * C + cb + threadIdx.x] = sb;
        }
        __syncthreads();
    }
}
// forward finalize (k_batchnorm_2, nmath.cu:224-237): scratch[0,C)=rvar, [C,2C)=mean
__global__ void __launch_bounds__(T4K_THREADS)
k_bn_fin_fwd(const float *__restrict__ part, float *scratch, int64_t NHW, int C, int nparts) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float sx = 0.0f, sq = 0.0f;
    for (int k = 0; k < nparts; k++) { sx += part[((int64_t)k * 2) * C + c]; sq += part[((int64_t)k * 2 + 1) * C + c]; }
    const float b_avg = sx / (float)NHW;
    const float b_var = sq / (float)NHW - b_avg * b_avg;
    scratch[C + c] = b_avg;
    scratch[c]     = 1.0f / (__fsqrt_rn(fmaxf(b_var, 0.0f)) + DU_EPS);     // eps OUTSIDE the sqrt (nmath.cu:236)
}
// backward finalize (k_dbatchnorm_2, nmath.cu:360-382): s1 = mean(dy) → scratch[C,2C), s2 = mean(dy·x̂) → [2C,3C)
__global__ void __launch_bounds__(T4K_THREADS)
k_bn_fin_bwd(const float *__restrict__ part, float *scratch, float *dW, float *dB, int64_t NHW, int C, int nparts, int train) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float a = 0.0f, b = 0.0f;
    for (int k = 0; k < nparts; k++) { a += part[((int64_t)k * 2) * C + c]; b += part[((int64_t)k * 2 + 1) * C + c]; }
    const float g0 = a / (float)NHW, g1 = b / (float)NHW;
    scratch[C + c] = g0; scratch[2 * C + c] = g1;
    if (train) { dB[c] += g0; dW[c] += g1; }                               // MEANS, not sums (nmath.cu:378-381)
}
// apply: XH = (x-mean)*rvar ; O = XH*gamma + beta      (k_batchnorm_3, nmath.cu:239-264)
template<bool VEC>
__global__ void __launch_bounds__(T4K_THREADS)
k_bn_apply(const float *__restrict__ I, float *O, float *XH, const float *__restrict__ gamma, const float *__restrict__ beta,
           const float *__restrict__ scratch, int C, int64_t total) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    if (VEC) {          // C % 4 == 0: a float4 never straddles a row
        for (int64_t k = tid; k < (total >> 2); k += nth) {
            const int c = (int)((4 * k) % C);
            const float4 x = ldg4(I + 4 * k), av = ldg4(scratch + C + c), rv = ldg4(scratch + c), g = ldg4(gamma + c), b = ldg4(beta + c);
            float4 h, o;
            h.x = (x.x - av.x) * rv.x; h.y = (x.y - av.y) * rv.y; h.z = (x.z - av.z) * rv.z; h.w = (x.w - av.w) * rv.w;
            o.x = h.x * g.x + b.x; o.y = h.y * g.y + b.y; o.z = h.z * g.z + b.z; o.w = h.w * g.w + b.w;
            stg4(XH + 4 * k, h); stg4(O + 4 * k, o);
        }
    } else {
        for (int64_t k = tid; k < total; k += nth) {
            const int c = (int)(k % C);
            const float h = (I[k] - scratch[C + c]) * scratch[c];
            XH[k] = h; O[k] = h * gamma[c] + beta[c];
        }
    }
}
// dX = gamma*rvar*(dy - s1 - xhat*s2)      (k_dbatchnorm_3, nmath.cu:395-414)
template<bool VEC>
__global__ void __launch_bounds__(T4K_THREADS)
k_bn_dx(const float *__restrict__ dO, const float *__restrict__ XH, float *dX, const float *__restrict__ gamma,
        const float *__restrict__ scratch, int C, int64_t total) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    if (VEC) {
        for (int64_t k = tid; k < (total >> 2); k += nth) {
            const int c = (int)((4 * k) % C);
            const float4 d = ldg4(dO + 4 * k), h = ldg4(XH + 4 * k), rv = ldg4(scratch + c), g = ldg4(gamma + c),
                         s1 = ldg4(scratch + C + c), s2 = ldg4(scratch + 2 * C + c);
            float4 o;
            o.x = (rv.x * g.x) * (d.x - s1.x - h.x * s2.x); o.y = (rv.y * g.y) * (d.y - s1.y - h.y * s2.y);
            o.z = (rv.z * g.z) * (d.z - s1.z - h.z * s2.z); o.w = (rv.w * g.w) * (d.w - s1.w - h.w * s2.w);
            stg4(dX + 4 * k, o);
        }
    } else {
        for (int64_t k = tid; k < total; k += nth) {
            const int c = (int)(k % C);
            dX[k] = (scratch[c] * gamma[c]) * (dO[k] - scratch[C + c] - XH[k] * scratch[2 * C + c]);
        }
    }
}

// data parallel: the CTAs' partial column sums -> ONE [2C] vector (the payload of the cross-rank SUM), in CTA order
__global__ void __launch_bounds__(T4K_THREADS)
k_bn_sum_parts(const float *__restrict__ part, float *out2C, int C, int nparts) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * C) return;
    const int which = t / C, c = t - which * C;
    float s = 0.0f;
    for (int k = 0; k < nparts; k++) s += part[((int64_t)k * 2 + which) * C + c];
    out2C[t] = s;
}
// backward finalize, data parallel: s1 / s2 are means over the GLOBAL batch (sums already reduced over the ranks); the parameter gradients
// receive this rank's SHARE, local sum / global rows — the gradient exchange adds the ranks' shares to the reference's means (nmath.cu:378-381)
__global__ void __launch_bounds__(T4K_THREADS)
k_bn_fin_bwd_dp(const float *__restrict__ glob2C, const float *__restrict__ loc2C, float *scratch, float *dW, float *dB, int64_t NHW_global, int C, int train) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    scratch[C + c] = glob2C[c] / (float)NHW_global; scratch[2 * C + c] = glob2C[C + c] / (float)NHW_global;
    if (train) { dB[c] += loc2C[c] / (float)NHW_global; dW[c] += loc2C[C + c] / (float)NHW_global; }
}

static int bn_parts(int64_t rows, int64_t *rows_per) {
    int nparts = 4 * sm_count();
    int64_t rp = (rows + nparts - 1) / nparts; if (rp < 1) rp = 1;
    *rows_per = rp;
    return (int)((rows + rp - 1) / rp);
}

} // namespace t4k
using namespace t4k;

// ====================================================================== C ABI
extern "C" int t4k_pool_fwd(int layer, const float *I, float *O, int N, int H1, int W1, int H0, int W0, int C, int KS, t4k_stream_t s) {
    if (!I || !O || N < 1 || H1 < 1 || W1 < 1 || H0 < 1 || W0 < 1 || C < 1) return T4K_EINVAL;
    if (layer != T4K_L_AVGPOOL && layer != T4K_L_MAXPOOL && layer != T4K_L_MINPOOL && layer != T4K_L_USAMPLE) return T4K_EINVAL;
    if (KS != 2 && KS != 3) return T4K_ENOSUP;                               // forward.cu:223-226
    if ((int64_t)H0 * KS > H1 || (int64_t)W0 * KS > W1) return T4K_EINVAL;   // the reference reads out of bounds here
    const int64_t total = (int64_t)N * H0 * W0 * C;
    if (KS == 2) k_pool<2><<<stream_grid(total), T4K_THREADS, 0, STRM(s)>>>(layer, I, O, H1, W1, H0, W0, C, total);
    else         k_pool<3><<<stream_grid(total), T4K_THREADS, 0, STRM(s)>>>(layer, I, O, H1, W1, H0, W0, C, total);
    return check_launch();
}
extern "C" int t4k_pool_bwd(int layer, float *I, const float *dO, int N, int H1, int W1, int H0, int W0, int C, int KS, t4k_stream_t s) {
    if (!I || !dO || N < 1 || H1 < 1 || W1 < 1 || H0 < 1 || W0 < 1 || C < 1) return T4K_EINVAL;
    if (layer != T4K_L_AVGPOOL && layer != T4K_L_MAXPOOL && layer != T4K_L_MINPOOL && layer != T4K_L_USAMPLE) return T4K_EINVAL;
    if (KS != 2 && KS != 3) return T4K_ENOSUP;
    if ((int64_t)H0 * KS > H1 || (int64_t)W0 * KS > W1) return T4K_EINVAL;
    const int64_t total = (int64_t)N * H0 * W0 * C;
    if (KS == 2) k_dpool<2><<<stream_grid(total), T4K_THREADS, 0, STRM(s)>>>(layer, I, dO, H1, W1, H0, W0, C, total);
    else         k_dpool<3><<<stream_grid(total), T4K_THREADS, 0, STRM(s)>>>(layer, I, dO, H1, W1, H0, W0, C, total);
    return check_launch();
}

extern "C" int t4k_batchnorm_fwd(const float *I, float *O, float *XH, const float *gamma, const float *beta,
                                 float *scratch3C, int N, int HW, int C, t4k_stream_t s) {
    if (!I || !O || !XH || !gamma || !beta || !scratch3C || N < 1 || HW < 1 || C < 1) return T4K_EINVAL;
    const int64_t rows = (int64_t)N * HW, total = rows * C;
    int64_t rows_per; const int nparts = bn_parts(rows, &rows_per);
    float *part = (float*)workspace((size_t)nparts * 2 * C * sizeof(float), 6);
    if (!part) return T4K_ENOMEM;
    k_bn_colsum2<0><<<nparts, T4K_THREADS, 0, STRM(s)>>>(I, nullptr, part, rows, C, rows_per);
    int rc = check_launch(); if (rc) return rc;
    k_bn_fin_fwd<<<(C + T4K_THREADS - 1) / T4K_THREADS, T4K_THREADS, 0, STRM(s)>>>(part, scratch3C, rows, C, nparts);
    rc = check_launch(); if (rc) return rc;
    const bool vec = (C % 4 == 0) && aligned16(I) && aligned16(O) && aligned16(XH) && aligned16(gamma) && aligned16(beta) && aligned16(scratch3C);
    if (vec) k_bn_apply<true ><<<stream_grid(total, 4), T4K_THREADS, 0, STRM(s)>>>(I, O, XH, gamma, beta, scratch3C, C, total);
    else     k_bn_apply<false><<<stream_grid(total, 1), T4K_THREADS, 0, STRM(s)>>>(I, O, XH, gamma, beta, scratch3C, C, total);
    return check_launch();
}
extern "C" int t4k_batchnorm_bwd(const float *dO, const float *XH, float *dX, const float *gamma,
                                 float *dgamma, float *dbeta, float *scratch3C, int N, int HW, int C, int train, t4k_stream_t s) {
    if (!dO || !XH || !dX || !gamma || !scratch3C || N < 1 || HW < 1 || C < 1) return T4K_EINVAL;
    if (train && (!dgamma || !dbeta)) return T4K_EINVAL;
    const int64_t rows = (int64_t)N * HW, total = rows * C;
    int64_t rows_per; const int nparts = bn_parts(rows, &rows_per);
    float *part = (float*)workspace((size_t)nparts * 2 * C * sizeof(float), 6);
    if (!part) return T4K_ENOMEM;
    k_bn_colsum2<1><<<nparts, T4K_THREADS, 0, STRM(s)>>>(dO, XH, part, rows, C, rows_per);
    int rc = check_launch(); if (rc) return rc;
    k_bn_fin_bwd<<<(C + T4K_THREADS - 1) / T4K_THREADS, T4K_THREADS, 0, STRM(s)>>>(part, scratch3C, dgamma, dbeta, rows, C, nparts, train);
    rc = check_launch(); if (rc) return rc;
    const bool vec = (C % 4 == 0) && aligned16(dO) && aligned16(XH) && aligned16(dX) && aligned16(gamma) && aligned16(scratch3C);
    if (vec) k_bn_dx<true ><<<stream_grid(total, 4), T4K_THREADS, 0, STRM(s)>>>(dO, XH, dX, gamma, scratch3C, C, total);
    else     k_bn_dx<false><<<stream_grid(total, 1), T4K_THREADS, 0, STRM(s)>>>(dO, XH, dX, gamma, scratch3C, C, total);
    return check_launch();
}

/* Batch norm over a batch that is SHARDED across the ranks of `comm` (SURVEY §8e collective 2): the per-channel sums of the shard are
 * SUM-all-reduced (2C floats over NVLink peer memory, t4k_allreduce_sum) between the statistics pass and the apply pass, so mean / variance
 * (forward) and mean(dy), mean(dy*xhat) (backward) are those of the global batch of N_global samples — the single-device result of
 * k_batchnorm_1/2/3 and k_dbatchnorm_1/2/3 (src/nn/nmath.cu:177-264,295-414) on the concatenated batch.  `comm` must be a communicator of
 * its own (capacity >= 4C), not the one the gradient arena is exchanged on: its chunks' epochs advance with every call. */
extern "C" int t4k_batchnorm_fwd_dp(t4k_comm_t comm, const float *I, float *O, float *XH, const float *gamma, const float *beta,
                                    float *scratch3C, int N, int N_global, int HW, int C, t4k_stream_t s) {
    if (!comm || !I || !O || !XH || !gamma || !beta || !scratch3C || N < 1 || N_global < N || HW < 1 || C < 1 || t4k_comm_capacity(comm) < 4 * (int64_t)C) return T4K_EINVAL;
    const int64_t rows = (int64_t)N * HW, total = rows * C;
    int64_t rows_per; const int nparts = bn_parts(rows, &rows_per);
    float *part = (float*)workspace(((size_t)nparts * 2 * C + 4 * (size_t)C + 8) * sizeof(float), 6);
    if (!part) return T4K_ENOMEM;
    float *glob = part + (size_t)nparts * 2 * C; glob = (float*)(((uintptr_t)glob + 15) & ~(uintptr_t)15);
    k_bn_colsum2<0><<<nparts, T4K_THREADS, 0, STRM(s)>>>(I, nullptr, part, rows, C, rows_per);
    int rc = check_launch(); if (rc) return rc;
    k_bn_sum_parts<<<(2 * C + T4K_THREADS - 1) / T4K_THREADS, T4K_THREADS, 0, STRM(s)>>>(part, glob, C, nparts);
    rc = check_launch(); if (rc) return rc;
    rc = t4k_allreduce_sum(comm, glob, 2 * (int64_t)C, s); if (rc) return rc;
    k_bn_fin_fwd<<<(C + T4K_THREADS - 1) / T4K_THREADS, T4K_THREADS, 0, STRM(s)>>>(glob, scratch3C, (int64_t)N_global * HW, C, 1);
    rc = check_launch(); if (rc) return rc;
    const bool vec = (C % 4 == 0) && aligned16(I) && aligned16(O) && aligned16(XH) && aligned16(gamma) && aligned16(beta) && aligned16(scratch3C);
    if (vec) k_bn_apply<true ><<<stream_grid(total, 4), T4K_THREADS, 0, STRM(s)>>>(I, O, XH, gamma, beta, scratch3C, C, total);
    else     k_bn_apply<false><<<stream_grid(total, 1), T4K_THREADS, 0, STRM(s)>>>(I, O, XH, gamma, beta, scratch3C, C, total);
    return check_launch();
}
extern "C" int t4k_batchnorm_bwd_dp(t4k_comm_t comm, const float *dO, const float *XH, float *dX, const float *gamma,
                                    float *dgamma, float *dbeta, float *scratch3C, int N, int N_global, int HW, int C, int train, t4k_stream_t s) {
    if (!comm || !dO || !XH || !dX || !gamma || !scratch3C || N < 1 || N_global < N || HW < 1 || C < 1 || t4k_comm_capacity(comm) < 4 * (int64_t)C) return T4K_EINVAL;
    if (train && (!dgamma || !dbeta)) return T4K_EINVAL;
    const int64_t rows = (int64_t)N * HW, total = rows * C;
    int64_t rows_per; const int nparts = bn_parts(rows, &rows_per);
    float *part = (float*)workspace(((size_t)nparts * 2 * C + 4 * (size_t)C + 8) * sizeof(float), 6);
    if (!part) return T4K_ENOMEM;
    float *glob = part + (size_t)nparts * 2 * C; glob = (float*)(((uintptr_t)glob + 15) & ~(uintptr_t)15);
    float *loc = glob + 2 * (size_t)C;
    k_bn_colsum2<1><<<nparts, T4K_THREADS, 0, STRM(s)>>>(dO, XH, part, rows, C, rows_per);
    int rc = check_launch(); if (rc) return rc;
    k_bn_sum_parts<<<(2 * C + T4K_THREADS - 1) / T4K_THREADS, T4K_THREADS, 0, STRM(s)>>>(part, loc, C, nparts);
    rc = check_launch(); if (rc) return rc;
    rc = t4k_copy(loc, glob, 2 * (int64_t)C, s); if (rc) return rc;
    rc = t4k_allreduce_sum(comm, glob, 2 * (int64_t)C, s); if (rc) return rc;
    k_bn_fin_bwd_dp<<<(C + T4K_THREADS - 1) / T4K_THREADS, T4K_THREADS, 0, STRM(s)>>>(glob, loc, scratch3C, dgamma, dbeta, (int64_t)N_global * HW, C, train);
    rc = check_launch(); if (rc) return rc;
    const bool vec = (C % 4 == 0) && aligned16(dO) && aligned16(XH) && aligned16(dX) && aligned16(gamma) && aligned16(scratch3C);
    if (vec) k_bn_dx<true ><<<stream_grid(total, 4), T4K_THREADS, 0, STRM(s)>>>(dO, XH, dX, gamma, scratch3C, C, total);
    else     k_bn_dx<false><<<stream_grid(total, 1), T4K_THREADS, 0, STRM(s)>>>(dO, XH, dX, gamma, scratch3C, C, total);
    return check_launch();
}
