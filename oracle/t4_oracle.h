/*
 * t4_oracle.h — CPU restatement of tensorForth's tensor-op hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (tensorforth_b200/,
 * include/) may include, link or call this.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs use it, and only as the
 * checker / the timed CPU baseline.
 *
 * Every function cites the reference file:line (relative to /root/reference)
 * whose arithmetic it restates.  All data is FP32, layout NHWC, Tensor.shape =
 * {H,W,C,N} (src/mu/tensor.h:53,109-112).
 *
 * Parity status: pinned for GEMM / elementwise / linear fwd+bwd / sigmoid /
 * MSE / SGD by the reference's own known-answer scripts (examples/t4_20a.4th,
 * t4_30a.4th, t4_30b.4th, t4_30c.4th — see tests/test_oracle_golden.py).
 * conv2d, pool, softmax+CE, batchnorm, Adam have no numeric golden values in
 * the reference's own tests; those are pinned by outputs of the reference's own
 * kernels run on a B200 through oracle/ref/refkern.cu (tests/golden/*.npz).
 */
#ifndef T4_ORACLE_H
#define T4_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* src/t4math.h:25-56 (enum math_op) — same numeric values */
enum {
    O_ABS = 0, O_NEG, O_EXP, O_LN, O_LOG, O_TANH, O_RELU, O_SIGM, O_SQRT, O_RCP,
    O_SAT, O_IDEN, O_FILL, O_GFILL, O_SCALE, O_POW, O_ADD, O_SUB, O_MUL, O_DIV,
    O_MOD, O_MAX, O_MIN
};
/* src/nn/ntypes.h:16-36 (enum t4_layer) — same numeric values */
enum {
    OL_NONE = 0, OL_CONV, OL_LINEAR, OL_FLATTEN, OL_RELU, OL_TANH, OL_SIGMOID,
    OL_SELU, OL_LEAKYRL, OL_ELU, OL_DROPOUT, OL_SOFTMAX, OL_LOGSMAX, OL_AVGPOOL,
    OL_MAXPOOL, OL_MINPOOL, OL_BATCHNM, OL_USAMPLE, OL_DCONV
};
/* src/nn/ntypes.h:38-43 (enum t4_loss) */
enum { OLOSS_MSE = 0, OLOSS_BCE, OLOSS_CE, OLOSS_NLL };

/* ---- src/t4math.cu ---------------------------------------------------- */
void  orc_gemm(const float *A, const float *B, float *O, float alpha, float beta,
               int tA, int tB, int M, int N, int K, int C);
void  orc_gemm_f64acc(const float *A, const float *B, float *O, float alpha, float beta,
               int M, int N, int K, int C);
void  orc_map(int op, float *A, float v, long n);
void  orc_ts_op(int op, const float *A, float v, float *O, long n);
void  orc_tt_op(int op, const float *A, const float *B, float *O, long n);
void  orc_copy(const float *src, float *dst, long n);
void  orc_transpose(const float *src, float *dst, int H, int W, int C);
void  orc_identity(float *T, int H, int W, int C);
float orc_sum(const float *A, long n);
float orc_nvar(const float *A, float avg, long n);
float orc_max(const float *A, long n, int find_max);
void  orc_dot(const float *A, const float *B, float *O, float alpha, float beta, int K, int C);
float orc_bce_sum(const float *T, const float *O, long n);
/* ---- src/mu/tensor.cu ------------------------------------------------- */
float orc_avg(const float *A, long n);
float orc_std(const float *A, long n);
float orc_norm(const float *A, long n);
float orc_loss(int op, float *out_copy, const float *tgt, long numel, int N);
/* ---- src/nn/nmath.cu, nmath.tcu --------------------------------------- */
void  orc_bias(const float *B, float *O, int N, int E0);
void  orc_dlinear_db(const float *dY, float *dB, int N, int E0);
void  orc_activate(int layer, const float *I, float *O, float *F, float alpha, long n);
void  orc_softmax(const float *I, float *O, int N, int C);
void  orc_logsoftmax(const float *I, float *O, int N, int C);
void  orc_conv2d(const float *I, const float *F, const float *B, float *O,
                 int N, int H1, int W1, int C1, int H0, int W0, int C0, int KS, int S, int P);
void  orc_dconv2d(const float *I, const float *dO, const float *F,
                  float *dX, float *dF, float *dB,
                  int N, int H1, int W1, int C1, int H0, int W0, int C0,
                  int KS, int S, int P, int train);
void  orc_pool(int layer, const float *I, float *O, int N, int H1, int W1, int H0, int W0, int C, int KS);
void  orc_dpool(int layer, float *I, const float *dO, int N, int H1, int W1, int H0, int W0, int C, int KS);
void  orc_batchnorm(const float *I, float *O, float *XH, const float *W, const float *B,
                    float *avg, float *rvar, int N, int HW, int C);
void  orc_dbatchnorm(const float *dO, const float *XH, float *dX, const float *W,
                     float *dW, float *dB, const float *rvar, float *s1, float *s2,
                     int N, int HW, int C, int train);
void  orc_sgd(float *G, float *DG, float *M, int Nw, float lr, float b, long n);
void  orc_adam(float *G, float *DG, float *M, float *V, float lr, float b1, float b2, long n);
void  orc_adamw(float *G, float *DG, float *M, float *V, float lr, float b1, float b2, float wd, long n);
/* ---- src/nn/loss.cpp -------------------------------------------------- */
void  orc_onehot(const int *label, float *hot, int N, int E);
int   orc_hit(const float *out, const float *hot, int N, int E);

#ifdef __cplusplus
}
#endif
#endif
