/*
 * t4_oracle.c — CPU restatement (plain C, FP32) of tensorForth's tensor-op hot path.
 *
 * TEST INFRASTRUCTURE ONLY — see t4_oracle.h.  Not a product path, not a fallback.
 *
 * Each function restates one reference kernel / wrapper; the citation (file:line,
 * relative to /root/reference) is on the function.  Where the reference result depends
 * on atomicAdd arrival order (k_sum, k_conv2d, k_dconv2d, k_batchnorm_1 ...) the oracle
 * accumulates in double and rounds once, i.e. it returns the value every legal ordering
 * of the reference is an FP32-rounding-noise away from.
 */
#include "t4_oracle.h"
#include <math.h>
#include <float.h>
#include <string.h>
#include <stdlib.h>

#define DU_EPS  1.0e-6f      /* src/ten4_types.h:85 */
#define DU_LNX  1.0e-12f     /* src/t4math.cu:172   */
#define SELU_L  1.0507       /* src/nn/nmath.h:33   */
#define SELU_LA 1.7581       /* src/nn/nmath.h:34   */

/* ---------------------------------------------------------------------------
 * GEMM: O[M,N,C] = alpha*op(A)@op(B) + beta*O, channel interleaved (stride C).
 * src/t4math.cu:478-583 (k_gemm_tile_claude): FP32 FMA accumulate, k ascending;
 * index maps :525-526 (A), :539-540 (B), :579-580 (O, `acc*alpha + O*beta`).
 * One call == one sample n (the per-sample loop is src/mu/tensor.cu:175-178).
 * ------------------------------------------------------------------------- */
void orc_gemm(const float *A, const float *B, float *O, float alpha, float beta,
              int tA, int tB, int M, int N, int K, int C)
{
    #pragma omp parallel for collapse(2) schedule(static)
    for (int m = 0; m < M; m++) {
        for (int n = 0; n < N; n++) {
            for (int c = 0; c < C; c++) {
                float acc = 0.0f;
                for (int k = 0; k < K; k++) {
                    long ai = tA ? ((long)k * M + m) * C + c : ((long)m * K + k) * C + c;
                    long bi = tB ? ((long)n * K + k) * C + c : ((long)k * N + n) * C + c;
                    acc = fmaf(A[ai], B[bi], acc);
                }
                long z0 = ((long)m * N + n) * C + c;
                /* beta==0 still reads O in the reference (O*0); keep NaN-propagation identical */
                O[z0] = acc * alpha + O[z0] * beta;
            }
        }
    }
}
/* src/t4math.cu:370-391 (k_gemm) and :411-452 (k_gemm_claude): double accumulator */
void orc_gemm_f64acc(const float *A, const float *B, float *O, float alpha, float beta,
                     int M, int N, int K, int C)
{
    #pragma omp parallel for collapse(2) schedule(static)
    for (int m = 0; m < M; m++)
        for (int n = 0; n < N; n++)
            for (int c = 0; c < C; c++) {
                double acc = 0.0;
                for (int k = 0; k < K; k++)
                    acc += A[((long)m * K + k) * C + c] * B[((long)k * N + n) * C + c];
                long z0 = ((long)m * N + n) * C + c;
                O[z0] = (float)(alpha * acc + beta * O[z0]);
            }
}
/* ---------------------------------------------------------------------------
 * k_math — src/t4math.cu:173-202; element macros src/t4math.h:60-104
 * ------------------------------------------------------------------------- */
void orc_map(int op, float *A, float v, long n)
{
    for (long j = 0; j < n; j++) {
        float ak = A[j];
        switch (op) {
        case O_ABS:   A[j] = fabsf(ak);                          break;
        case O_NEG:   A[j] = -ak;                                break;
        case O_EXP:   A[j] = expf(ak);                           break; /* __expf */
        case O_LN:    A[j] = logf(fmaxf(ak, DU_LNX));            break; /* __logf, clamped */
        case O_LOG:   A[j] = log10f(fmaxf(ak, DU_LNX));          break; /* __log10f, clamped */
        case O_TANH:  A[j] = tanhf(ak);                          break;
        case O_RELU:  A[j] = fmaxf(0.0f, ak);                    break;
        case O_SIGM:  A[j] = 1.0f / (1.0f + expf(-ak));          break;
        case O_SQRT:  A[j] = sqrtf(fmaxf(ak, 0.0f));             break;
        case O_RCP:   A[j] = 1.0f / ak;                          break;
        case O_SAT:   A[j] = fminf(1.0f, fmaxf(0.0f, ak));       break;
        case O_FILL:  A[j] = v;                                  break;
        case O_GFILL: A[j] = v * (float)j / (float)n;            break; /* :192 `v * j / n` */
        case O_SCALE: A[j] = ak * v;                             break;
        case O_POW:   A[j] = powf(ak, v);                        break; /* __powf */
        case O_ADD:   A[j] = ak + v;                             break;
        case O_SUB:   A[j] = ak - v;                             break;
        case O_MUL:   A[j] = ak * v;                             break;
        case O_DIV:   A[j] = ak / v;                             break;
        default: break;
        }
    }
}
/* k_ts_op — src/t4math.cu:206-218 */
void orc_ts_op(int op, const float *A, float v, float *O, long n)
{
    for (long j = 0; j < n; j++) {
        switch (op) {
        case O_ADD: O[j] = A[j] + v; break;
        case O_SUB: O[j] = A[j] - v; break;
        case O_MUL: O[j] = A[j] * v; break;
        case O_DIV: O[j] = A[j] / v; break;
        default: break;
        }
    }
}
/* k_tt_op — src/t4math.cu:222-234 (the N-broadcast loop is src/mu/tensor.cu:39-46) */
void orc_tt_op(int op, const float *A, const float *B, float *O, long n)
{
    for (long j = 0; j < n; j++) {
        switch (op) {
        case O_ADD: O[j] = A[j] + B[j]; break;
        case O_SUB: O[j] = A[j] - B[j]; break;
        case O_MUL: O[j] = A[j] * B[j]; break;
        case O_DIV: O[j] = A[j] / B[j]; break;
        default: break;
        }
    }
}
/* k_copy — src/t4math.cu:134-149 */
void orc_copy(const float *src, float *dst, long n) { memmove(dst, src, (size_t)n * sizeof(float)); }
/* k_transpose — src/t4math.cu:150-159 (one sample; per-channel 2-D transpose) */
void orc_transpose(const float *src, float *dst, int H, int W, int C)
{
    for (int i = 0; i < H; i++)
        for (int j = 0; j < W; j++)
            for (int c = 0; c < C; c++)
                dst[((long)H * j + i) * C + c] = src[((long)W * i + j) * C + c];
}
/* k_identity — src/t4math.cu:160-170 */
void orc_identity(float *T, int H, int W, int C)
{
    for (int i = 0; i < H; i++)
        for (int j = 0; j < W; j++)
            for (int c = 0; c < C; c++)
                T[((long)W * i + j) * C + c] = (i == j) ? 1.0f : 0.0f;
}
/* k_sum — src/t4math.cu:23-46 (atomic order is not reproducible → double accumulate) */
float orc_sum(const float *A, long n)
{
    double s = 0.0;
    for (long j = 0; j < n; j++) s += A[j];
    return (float)s;
}
/* k_nvar — src/t4math.cu:48-72: Σ(x-avg)² */
float orc_nvar(const float *A, float avg, long n)
{
    double s = 0.0;
    for (long j = 0; j < n; j++) { float d = A[j] - avg; s += (double)(d * d); }
    return (float)s;
}
/* k_max — src/t4math.cu:102-131 */
float orc_max(const float *A, long n, int find_max)
{
    float m = find_max ? -FLT_MAX : FLT_MAX;
    for (long j = 0; j < n; j++) m = find_max ? fmaxf(m, A[j]) : fminf(m, A[j]);
    return m;
}
/* k_dot — src/t4math.cu:309-365: O[c] = alpha*Σ_k A[k,c]B[k,c] + beta*O[c] */
void orc_dot(const float *A, const float *B, float *O, float alpha, float beta, int K, int C)
{
    for (int c = 0; c < C; c++) {
        double acc = 0.0;
        for (int k = 0; k < K; k++) acc += (double)(A[(long)k * C + c] * B[(long)k * C + c]);
        O[c] = (float)acc * alpha + O[c] * beta;
    }
}
/* k_bce — src/t4math.cu:248-274: Σ t·ln(o+ε) + (1-t)·ln(1-o+ε) */
float orc_bce_sum(const float *T, const float *O, long n)
{
    double s = 0.0;
    for (long j = 0; j < n; j++) {
        float t = T[j], o = O[j];
        s += (double)(t * logf(o + DU_EPS) + (1.0f - t) * logf(1.0f - o + DU_EPS));
    }
    return (float)s;
}
/* Tensor::avg/std/norm — src/mu/tensor.cu:238-259 (std = sqrt(Σ(x-μ)²)/n, sic) */
float orc_avg(const float *A, long n)  { return orc_sum(A, n) / (float)n; }
float orc_std(const float *A, long n)  { return n ? sqrtf(orc_nvar(A, orc_avg(A, n), n)) / (float)n : 0.0f; }
float orc_norm(const float *A, long n) { return sqrtf(orc_nvar(A, 0.0f, n)); }
/* ---------------------------------------------------------------------------
 * Tensor::loss — src/mu/tensor.cu:289-325.  `out_copy` is destroyed (the reference
 * works on the Model::_loss duplicate, src/nn/loss.cpp:129-132).
 * ------------------------------------------------------------------------- */
float orc_loss(int op, float *o, const float *tgt, long numel, int N)
{
    float z = 0.0f;
    switch (op) {
    case OLOSS_MSE:
        orc_tt_op(O_SUB, o, tgt, o, numel);
        orc_tt_op(O_MUL, o, o, o, numel);
        z = orc_sum(o, numel);
        break;
    case OLOSS_BCE:
        z = -orc_bce_sum(tgt, o, numel);
        break;
    case OLOSS_CE:
        orc_map(O_LN, o, 0.0f, numel);            /* fallthrough, :313-315 */
    case OLOSS_NLL:
        orc_tt_op(O_MUL, o, tgt, o, numel);
        z = -orc_sum(o, numel);
        break;
    default: break;
    }
    return z / (float)N;
}
/* k_bias — src/nn/nmath.cu:27-35 */
void orc_bias(const float *B, float *O, int N, int E0)
{
    for (int n = 0; n < N; n++)
        for (int e = 0; e < E0; e++) O[(long)n * E0 + e] += B[e];
}
/* k_dlinear_db — src/nn/nmath.cu:274-280: dB[e] += Σ_n dY[n,e] */
void orc_dlinear_db(const float *dY, float *dB, int N, int E0)
{
    for (int e = 0; e < E0; e++) {
        double s = 0.0;
        for (int n = 0; n < N; n++) s += dY[(long)n * E0 + e];
        dB[e] += (float)s;
    }
}
/* k_activate — src/nn/nmath.cu:37-70 (writes output O and derivative/mask F) */
void orc_activate(int layer, const float *I, float *O, float *F, float alpha, long n)
{
    for (long j = 0; j < n; j++) {
        float i = I[j];
        switch (layer) {
        case OL_RELU:
            if (i > 0.0f) { F[j] = 1.0f; O[j] = i; } else { F[j] = 0.0f; O[j] = 0.0f; }
            break;
        case OL_TANH:
            O[j] = i = tanhf(i); F[j] = 1.0f - i * i;
            break;
        case OL_SIGMOID:
            O[j] = i = 1.0f / (1.0f + expf(-i)); F[j] = i * (1.0f - i);
            break;
        case OL_SELU:   /* positive branch outputs x, not lambda*x (:56-58) */
            if (i > 0.0f) { F[j] = (float)SELU_L; O[j] = i; }
            else { F[j] = (float)(SELU_LA * (double)expf(i)); O[j] = (float)((double)F[j] - SELU_LA); }
            break;
        case OL_LEAKYRL:
            if (i > 0.0f) { F[j] = 1.0f; O[j] = i; } else { F[j] = alpha; O[j] = alpha * i; }
            break;
        case OL_ELU:
            if (i > 0.0f) { F[j] = 1.0f; O[j] = i; }
            else { F[j] = alpha * expf(i); O[j] = F[j] - alpha; }
            break;
        case OL_DROPOUT: /* F holds U(0,1] on entry; keep iff F > p; no 1/(1-p) rescale (:65-67) */
            if (F[j] > alpha) { F[j] = 1.0f; O[j] = i; } else { F[j] = 0.0f; O[j] = 0.0f; }
            break;
        default: break;
        }
    }
}
/* k_softmax_small / k_softmax — src/nn/nmath.cu:74-169: exp(x-max)/Σ per sample */
void orc_softmax(const float *I, float *O, int N, int C)
{
    for (int n = 0; n < N; n++) {
        const float *s = I + (long)n * C; float *d = O + (long)n * C;
        float mx = -FLT_MAX;
        for (int c = 0; c < C; c++) mx = fmaxf(mx, s[c]);
        double sm = 0.0;
        for (int c = 0; c < C; c++) { d[c] = expf(s[c] - mx); sm += d[c]; }
        float fs = (float)sm;
        for (int c = 0; c < C; c++) d[c] /= fs;
    }
}
/* Model::_flogsoftmax — src/nn/forward.cu:246-259 AS CODED:
 * out = exp(x); per sample: out -= log10(max(Σ out, 1e-6))   (LOG is log10, t4math.h:64) */
void orc_logsoftmax(const float *I, float *O, int N, int C)
{
    for (int n = 0; n < N; n++) {
        const float *s = I + (long)n * C; float *d = O + (long)n * C;
        for (int c = 0; c < C; c++) d[c] = expf(s[c]);
        float sum = orc_sum(d, C);
        float logsum = log10f(fmaxf(sum, DU_EPS));
        for (int c = 0; c < C; c++) d[c] -= logsum;
    }
}
/* ---------------------------------------------------------------------------
 * k_conv2d<TS,KS,S,P> — src/nn/nmath.tcu:34-104
 *   O[n,i,j,c0] = B[c0] + Σ_{c1,y,x} F[c1,y,x,c0] * I[n, i*S+y-P, j*S+x-P, c1]   (zero pad)
 * filter layout [C1,KS,KS,C0] (:73-77).  Accumulated over c1 by atomics in the
 * reference (:102) → order-free; oracle sums in double.
 * ------------------------------------------------------------------------- */
void orc_conv2d(const float *I, const float *F, const float *B, float *O,
                int N, int H1, int W1, int C1, int H0, int W0, int C0, int KS, int S, int P)
{
    #pragma omp parallel for collapse(2) schedule(static)
    for (int n = 0; n < N; n++) {
        for (int i = 0; i < H0; i++) {
            const float *nI = I + (long)n * H1 * W1 * C1;
            float *nO = O + (long)n * H0 * W0 * C0;
            for (int j = 0; j < W0; j++) {
                for (int c0 = 0; c0 < C0; c0++) {
                    double sum = B[c0];
                    for (int c1 = 0; c1 < C1; c1++)
                        for (int y = 0; y < KS; y++) {
                            int gi = i * S + y - P;
                            if (gi < 0 || gi >= H1) continue;
                            for (int x = 0; x < KS; x++) {
                                int gj = j * S + x - P;
                                if (gj < 0 || gj >= W1) continue;
                                sum += (double)(F[(((long)c1 * KS + y) * KS + x) * C0 + c0] *
                                                nI[((long)W1 * gi + gj) * C1 + c1]);
                            }
                        }
                    nO[((long)W0 * i + j) * C0 + c0] = (float)sum;
                }
            }
        }
    }
}
/* ---------------------------------------------------------------------------
 * k_dconv2d<TS,KS,S,P> — src/nn/nmath.tcu:211-338
 *   dB[c0] += Σ dO                                   (:274-278, only if train)
 *   dF[c1,ky,kx,c0] += Σ I[n,i*S+ky-P,j*S+kx-P,c1]*dO[n,i,j,c0]   (:307-312,:332-336, train)
 *   dX[n,i*S+ky-P,j*S+kx-P,c1] += F[c1,KS-1-ky,KS-1-kx,c0]*dO[n,i,j,c0]  (:304, FLIPPED taps)
 * dX is pre-zeroed by the caller (src/nn/backprop.cu:169); oracle zeroes it here.
 * dF/dB ACCUMULATE into the caller's buffers.
 * ------------------------------------------------------------------------- */
/* Flush multiplicity of dF taps — src/nn/nmath.tcu:332-336 flushes _df[t] once per thread whose
 * load_id = ty*TS + tx equals t, with (tx,ty) in [0,16)^2 and TS = (16-KS+S)/S (backprop.cu:144).
 * For KS=1,3 every t < KS*KS has exactly one such thread; for (KS,S)=(5,1) taps 12..15,24 are
 * flushed twice and for (4,2) taps 7..13 twice, 14..15 three times.  Reference behaviour
 * (verified against the reference kernels on a B200, tests/golden) → replicated. */
static int dconv_flush_mult(int KS, int S, int t)
{
    const int TS = (16 - KS + S) / S;
    int m = 0;
    for (int ty = 0; ty < 16; ty++)
        for (int tx = 0; tx < 16; tx++) if (ty * TS + tx == t) m++;
    return m;
}
void orc_dconv2d(const float *I, const float *dO, const float *F,
                 float *dX, float *dF, float *dB,
                 int N, int H1, int W1, int C1, int H0, int W0, int C0,
                 int KS, int S, int P, int train)
{
    const long nx = (long)N * H1 * W1 * C1;
    double *ax = (double*)calloc((size_t)nx, sizeof(double));
    #pragma omp parallel for schedule(static)
    for (int n = 0; n < N; n++) {
        const float *nO = dO + (long)n * H0 * W0 * C0;
        double *nX = ax + (long)n * H1 * W1 * C1;
        for (int i = 0; i < H0; i++)
            for (int j = 0; j < W0; j++)
                for (int c0 = 0; c0 < C0; c0++) {
                    float d = nO[((long)W0 * i + j) * C0 + c0];
                    for (int ky = 0; ky < KS; ky++) {
                        int gi = i * S + ky - P;
                        if (gi < 0 || gi >= H1) continue;
                        for (int kx = 0; kx < KS; kx++) {
                            int gj = j * S + kx - P;
                            if (gj < 0 || gj >= W1) continue;
                            for (int c1 = 0; c1 < C1; c1++) {
                                float f = F[(((long)c1 * KS + (KS-1-ky)) * KS + (KS-1-kx)) * C0 + c0];
                                nX[((long)W1 * gi + gj) * C1 + c1] += (double)(f * d);
                            }
                        }
                    }
                }
    }
    for (long k = 0; k < nx; k++) dX[k] = (float)ax[k];
    free(ax);
    if (!train) return;
    for (int c0 = 0; c0 < C0; c0++) {
        double s = 0.0;
        for (long p = 0; p < (long)N * H0 * W0; p++) s += dO[p * C0 + c0];
        dB[c0] += (float)s;
    }
    #pragma omp parallel for collapse(2) schedule(static)
    for (int c1 = 0; c1 < C1; c1++)
        for (int ky = 0; ky < KS; ky++)
            for (int kx = 0; kx < KS; kx++)
                for (int c0 = 0; c0 < C0; c0++) {
                    double s = 0.0;
                    for (int n = 0; n < N; n++) {
                        const float *nI = I + (long)n * H1 * W1 * C1;
                        const float *nO = dO + (long)n * H0 * W0 * C0;
                        for (int i = 0; i < H0; i++) {
                            int gi = i * S + ky - P;
                            if (gi < 0 || gi >= H1) continue;
                            for (int j = 0; j < W0; j++) {
                                int gj = j * S + kx - P;
                                if (gj < 0 || gj >= W1) continue;
                                s += (double)(nI[((long)W1 * gi + gj) * C1 + c1] *
                                              nO[((long)W0 * i + j) * C0 + c0]);
                            }
                        }
                    }
                    dF[(((long)c1 * KS + ky) * KS + kx) * C0 + c0] += (float)(s * dconv_flush_mult(KS, S, ky * KS + kx));
                }
}
/* ---------------------------------------------------------------------------
 * k_pool<KS> — src/nn/nmath.tcu:122-186.  stride == KS, window read WITHOUT a tail
 * guard (:153-160) → callers must pass H1 == H0*KS, W1 == W0*KS for in-bounds reads
 * (the reference reads past the row otherwise; the oracle clamps to the tensor and
 * the tests only use divisible sizes).  avg and upsample-backward: Σ/KS².
 * ------------------------------------------------------------------------- */
void orc_pool(int layer, const float *I, float *O, int N, int H1, int W1, int H0, int W0, int C, int KS)
{
    for (int n = 0; n < N; n++)
        for (int i0 = 0; i0 < H0; i0++)
            for (int j0 = 0; j0 < W0; j0++)
                for (int c = 0; c < C; c++) {
                    const float *ix = I + (long)n * H1 * W1 * C + ((long)(i0 * KS) * W1 + j0 * KS) * C + c;
                    float v = 0.0f;
                    for (int y = 0; y < KS; y++)
                        for (int x = 0; x < KS; x++) {
                            float t = ix[((long)y * W1 + x) * C];
                            int first = (y == 0 && x == 0);
                            switch (layer) {
                            case OL_USAMPLE:
                            case OL_AVGPOOL: v += t; break;
                            case OL_MAXPOOL: v = first ? t : fmaxf(t, v); break;
                            case OL_MINPOOL: v = first ? t : fminf(t, v); break;
                            default: break;
                            }
                        }
                    if (layer == OL_AVGPOOL || layer == OL_USAMPLE) v /= (float)(KS * KS);
                    O[(long)n * H0 * W0 * C + ((long)i0 * W0 + j0) * C + c] = v;
                }
}
/* ---------------------------------------------------------------------------
 * k_dpool<KS> — src/nn/nmath.tcu:475-568.  IN PLACE on the forward input I:
 *   max/min: zero the window, write dO at the FIRST strict max/min in (y,x) scan
 *            order (:535-549; `dx > best` with best initialised to tile[0]);
 *   avg: every cell = dO/KS²;  upsample(-forward): every cell = dO.
 * ------------------------------------------------------------------------- */
void orc_dpool(int layer, float *I, const float *dO, int N, int H1, int W1, int H0, int W0, int C, int KS)
{
    for (int n = 0; n < N; n++)
        for (int i0 = 0; i0 < H0; i0++)
            for (int j0 = 0; j0 < W0; j0++)
                for (int c = 0; c < C; c++) {
                    float *ix = I + (long)n * H1 * W1 * C + ((long)(i0 * KS) * W1 + j0 * KS) * C + c;
                    float d = dO[(long)n * H0 * W0 * C + ((long)i0 * W0 + j0) * C + c];
                    if (layer == OL_AVGPOOL || layer == OL_USAMPLE) {
                        float v = (layer == OL_AVGPOOL) ? d / (float)(KS * KS) : d;
                        for (int y = 0; y < KS; y++)
                            for (int x = 0; x < KS; x++) ix[((long)y * W1 + x) * C] = v;
                    } else {
                        float best = ix[0]; float *argp = ix;
                        for (int y = 0; y < KS; y++)
                            for (int x = 0; x < KS; x++) {
                                float *px = ix + ((long)y * W1 + x) * C;
                                float dx = *px; *px = 0.0f;
                                if (layer == OL_MAXPOOL ? (dx > best) : (dx < best)) { best = dx; argp = px; }
                            }
                        *argp = d;
                    }
                }
}
/* ---------------------------------------------------------------------------
 * k_batchnorm_1/2/3 — src/nn/nmath.cu:177-264 (batch statistics only):
 *   avg = Σx/NHW ; rvar = 1/(sqrt(max(Σx²/NHW - avg², 0)) + 1e-6)   (:232-236, eps OUTSIDE sqrt)
 *   XH = (x-avg)*rvar ; O = XH*gamma + beta                           (:262)
 * ------------------------------------------------------------------------- */
void orc_batchnorm(const float *I, float *O, float *XH, const float *W, const float *B,
                   float *avg, float *rvar, int N, int HW, int C)
{
    const long NHW = (long)N * HW;
    for (int c = 0; c < C; c++) {
        double s = 0.0, q = 0.0;
        for (long p = 0; p < NHW; p++) { float v = I[p * C + c]; s += v; q += (double)(v * v); }
        float fs = (float)s, fq = (float)q;
        float b_avg = fs / (float)NHW;
        float b_var = fq / (float)NHW - b_avg * b_avg;
        avg[c]  = b_avg;
        rvar[c] = 1.0f / (sqrtf(fmaxf(b_var, 0.0f)) + DU_EPS);
    }
    for (long p = 0; p < NHW; p++)
        for (int c = 0; c < C; c++) {
            long k = p * C + c;
            XH[k] = (I[k] - avg[c]) * rvar[c];
            O[k]  = XH[k] * W[c] + B[c];
        }
}
/* ---------------------------------------------------------------------------
 * k_dbatchnorm_1/2/3 — src/nn/nmath.cu:295-414:
 *   s1 = Σdy/NHW ; s2 = Σ(dy·x̂)/NHW ; if train: dβ += s1, dγ += s2  (MEANS, :371-381)
 *   dX = γ·rvar·(dy − s1 − x̂·s2)                                    (:410-413)
 * ------------------------------------------------------------------------- */
void orc_dbatchnorm(const float *dO, const float *XH, float *dX, const float *W,
                    float *dW, float *dB, const float *rvar, float *s1, float *s2,
                    int N, int HW, int C, int train)
{
    const long NHW = (long)N * HW;
    for (int c = 0; c < C; c++) {
        double a = 0.0, b = 0.0;
        for (long p = 0; p < NHW; p++) { a += dO[p * C + c]; b += (double)(dO[p * C + c] * XH[p * C + c]); }
        s1[c] = (float)a / (float)NHW;
        s2[c] = (float)b / (float)NHW;
        if (train) { dB[c] += s1[c]; dW[c] += s2[c]; }
    }
    for (long p = 0; p < NHW; p++)
        for (int c = 0; c < C; c++) {
            long k = p * C + c;
            float g = rvar[c] * W[c];
            dX[k] = g * (dO[k] - s1[c] - XH[k] * s2[c]);
        }
}
/* k_sgd — src/nn/nmath.cu:419-436: dg/Nw where Nw = PARAMETER tensor's N() (gradient.cu:137) */
void orc_sgd(float *G, float *DG, float *M, int Nw, float lr, float b, long n)
{
    for (long j = 0; j < n; j++) {
        float dg = DG[j] / (float)Nw;
        if (fabsf(b) < DU_EPS) G[j] -= lr * dg;
        else { float mi = M[j] = b * M[j] + (1.0f - b) * dg; G[j] -= lr * mi; }
        DG[j] = 0.0f;
    }
}
/* k_adam — src/nn/nmath.cu:438-454: no bias correction, eps outside sqrt, zero dG */
void orc_adam(float *G, float *DG, float *M, float *V, float lr, float b1, float b2, long n)
{
    for (long j = 0; j < n; j++) {
        float dg = DG[j];
        float mi = M[j] = b1 * M[j] + (1.0f - b1) * dg;
        float vi = V[j] = b2 * V[j] + (1.0f - b2) * dg * dg;
        G[j] -= lr * mi / (sqrtf(vi) + DU_EPS);
        DG[j] = 0.0f;
    }
}
/* k_adamw — src/nn/nmath.cu:456-472: g -= lr*(m/(sqrt(v)+eps) - wd*dg) */
void orc_adamw(float *G, float *DG, float *M, float *V, float lr, float b1, float b2, float wd, long n)
{
    for (long j = 0; j < n; j++) {
        float dg = DG[j];
        float mi = M[j] = b1 * M[j] + (1.0f - b1) * dg;
        float vi = V[j] = b2 * V[j] + (1.0f - b2) * dg * dg;
        G[j] -= lr * (mi / (sqrtf(vi) + DU_EPS) - wd * dg);
        DG[j] = 0.0f;
    }
}
/* Model::onehot(Dataset&) — src/nn/loss.cpp:47-72: h[m<E ? m : 0] = 1 */
void orc_onehot(const int *label, float *hot, int N, int E)
{
    memset(hot, 0, (size_t)N * E * sizeof(float));
    for (int n = 0; n < N; n++) { int m = label[n]; hot[(long)n * E + ((m >= 0 && m < E) ? m : 0)] = 1.0f; }
}
/* Model::hit — src/nn/loss.cpp:75-107: Σ_n (int)hot[n, argmax_first(out[n])] */
int orc_hit(const float *out, const float *hot, int N, int E)
{
    int cnt = 0;
    for (int n = 0; n < N; n++) {
        const float *o = out + (long)n * E;
        float m = o[0]; int i = 0;
        for (int e = 1; e < E; e++) if (o[e] > m) { m = o[e]; i = e; }
        cnt += (int)hot[(long)n * E + i];
    }
    return cnt;
}
