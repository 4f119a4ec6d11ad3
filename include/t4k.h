/*
 * t4k.h — C-ABI of libt4k.so: the B200-native (sm_100a) replacement of tensorForth's
 * tensor-op hot path.  Plain pointers and sizes, no C++/torch types.
 *
 * The reference has no plugin/FFI layer; its seam is source level: the __KERN__ prototypes
 * of src/t4math.h:138-184 and src/nn/nmath.h:41-112 as launched through the FORK* macros
 * (src/t4base.h:129-159) by src/mu/tensor.cu, src/nn/forward.cu, backprop.cu, gradient.cu.
 * Each entry point below names the reference launch site(s) it replaces (file:line relative
 * to the reference tree).  INTEGRATION.md shows the shim a maintainer adds on the reference
 * side (the bodies of Tensor::xxx / Model::_fxxx/_bxxx re-expressed on these calls).
 *
 * Conventions
 *  - All tensor data is FP32, NHWC; Tensor.shape = {H,W,C,N} (src/mu/tensor.h:53,109-112).
 *  - Every pointer is a DEVICE pointer (cudaMalloc or cudaMallocManaged) owned by the caller
 *    (MMU arena, src/mu/mmu.cu:199-209).  The library never allocates or frees user-visible
 *    memory; its private workspace (reduction partials, split planes, TMA descriptors) is
 *    internal, per device, grown on demand.
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream, which is what
 *    the reference uses: FORK* pass no stream).  Calls are STREAM-ORDERED and ASYNCHRONOUS:
 *    no cudaDeviceSynchronize() inside (the reference syncs after every launch via GPU_CHK,
 *    src/ten4_types.h:192).  Scalar results are written to a device float the caller reads
 *    (reference: Tensor::_tmp = &data[numel], src/mu/tensor.cu:231-233); use t4k_sync() or
 *    a stream-ordered cudaMemcpy at the host-read points.
 *  - Return value: 0 on success; a positive cudaError_t if a launch failed; negative T4K_E*
 *    for argument errors (the reference prints and returns the output untouched,
 *    src/mu/tensor.cu:35-38, src/nn/forward.cu:147-150).  Nothing is printed, nothing
 *    calls cudaDeviceReset().
 *  - There is NO CPU fallback: without a CUDA device every compute call returns an error.
 */
#ifndef T4K_H
#define T4K_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define T4K_VERSION 100                /* 1.00 */
#define T4K_EINVAL  (-1)               /* bad argument / unsupported shape */
#define T4K_ENOSUP  (-2)               /* configuration the reference also rejects (e.g. conv K,S,P) */
#define T4K_ENOMEM  (-3)               /* workspace allocation failed */

typedef void *t4k_stream_t;            /* cudaStream_t */
typedef struct t4k_comm *t4k_comm_t;     /* data-parallel communicator (see "data-parallel extras" below) */

/* math_op — identical numbering to src/t4math.h:25-56 */
enum t4k_math_op {
    T4K_ABS = 0, T4K_NEG, T4K_EXP, T4K_LN, T4K_LOG, T4K_TANH, T4K_RELU, T4K_SIGM, T4K_SQRT,
    T4K_RCP, T4K_SAT, T4K_IDEN, T4K_FILL, T4K_GFILL, T4K_SCALE, T4K_POW,
    T4K_ADD, T4K_SUB, T4K_MUL, T4K_DIV, T4K_MOD, T4K_MAX, T4K_MIN
};
/* t4_layer — identical numbering to src/nn/ntypes.h:16-36 */
enum t4k_layer {
    T4K_L_NONE = 0, T4K_L_CONV, T4K_L_LINEAR, T4K_L_FLATTEN, T4K_L_RELU, T4K_L_TANH,
    T4K_L_SIGMOID, T4K_L_SELU, T4K_L_LEAKYRL, T4K_L_ELU, T4K_L_DROPOUT, T4K_L_SOFTMAX,
    T4K_L_LOGSMAX, T4K_L_AVGPOOL, T4K_L_MAXPOOL, T4K_L_MINPOOL, T4K_L_BATCHNM,
    T4K_L_USAMPLE, T4K_L_DCONV
};
/* t4_loss — src/nn/ntypes.h:38-43 */
enum t4k_loss { T4K_LOSS_MSE = 0, T4K_LOSS_BCE, T4K_LOSS_CE, T4K_LOSS_NLL };
/* rand_opt — src/util.h (UNIFORM, NORMAL) */
enum t4k_rand_opt { T4K_UNIFORM = 0, T4K_NORMAL = 1 };
/* GEMM engine selection for t4k_gemm_ex (0 = automatic) */
enum t4k_gemm_engine { T4K_GEMM_AUTO = 0, T4K_GEMM_SIMT = 1,   /* FP32 FMA (gemm_simt.cu) */
                       T4K_GEMM_TC = 2,                        /* tcgen05 3xTF32, packed operand planes (gemm_tc.cu): large problems */
                       T4K_GEMM_TCF = 3,                       /* tcgen05 3xTF32, split fused into the kernel, one launch (gemm_tcf.cu): layer-sized problems */
                       T4K_GEMM_TC_BF16X3 = 4,                 /* tcgen05 BF16x3 (a = hi + lo in bf16; hi*hi + hi*lo + lo*hi): twice the MMA rate of 3xTF32;
                                                                * measured 4.1e-6 of the result's rms at K=4096 (3xTF32: 1.8e-6; the reference's FP32-FMA
                                                                * accumulation itself: ~3.8e-6).  AUTO takes it for M*N*K >= 2e10 only (T4K_GEMM_BIG=tf32: never) */
                       T4K_GEMM_MMA = 5,                       /* (retired: warp-level mma.sync engine of round 1, measured no faster than FP32 FMA; T4K_ENOSUP) */
                       T4K_GEMM_TL = 6 };                      /* tcgen05 3xTF32 LAYER GEMM (gemm_tl.cu): TMA-fed raw FP32 tiles (any transposition native: MN-major
                                                                * UMMA operands), lo plane derived in shared memory, split-K inside a thread-block cluster reduced over
                                                                * distributed shared memory, fused linear-layer epilogues; one launch.  AUTO takes it for
                                                                * 4e6 <= M*N*K < 2e10 when the operands are 16-byte aligned with row pitches that are multiples of 4 */

/* ---- library / device ------------------------------------------------------------- */
int         t4k_version(void);
const char *t4k_strerror(int rc);
int         t4k_device_count(void);                    /* 0 when no CUDA device/driver */
int         t4k_sm_count(void);                        /* SMs of the current device    */
int         t4k_sync(t4k_stream_t stream);             /* cudaStreamSynchronize        */
long        t4k_launch_count(void);                    /* kernels launched by this library so far */
int         t4k_set_workspace_bank(int bank);          /* 0..7: which set of library workspaces the following calls use — a caller that forks
                                                          * work onto a second stream gives that stream its own bank; returns the previous one */
int         t4k_set_carveout(int pct);                 /* shared-memory split (percent, -1 = leave it to the driver) the SHORT kernels ask for; default 100 so that they can share
                                                          * an SM with the big-shared-memory kernels of another stream (or T4K_CARVEOUT); returns the previous setting */
int         t4k_set_pdl(int on);                       /* programmatic dependent launch for the short kernels (default off, or T4K_PDL=1); returns the previous setting */

/* ---- elementwise: src/t4math.cu:134-234 ------------------------------------------- */
/* k_math via Tensor::map (src/mu/tensor.cu:566-571): in-place A[j] = op(A[j], v) */
int t4k_map(int op, float *A, float v, int64_t n, t4k_stream_t s);
/* k_ts_op via Tensor::ten_op(A,v,O) (src/mu/tensor.cu:17-23): O = A op v, op in ADD..DIV */
int t4k_ts_op(int op, const float *A, float v, float *O, int64_t n, t4k_stream_t s);
/* k_tt_op via Tensor::ten_op(A,B,O) (src/mu/tensor.cu:29-53): O[n] = A[Na==1?0:n] op B[Nb==1?0:n],
 * each slice `hwc` floats, N = max(Na,Nb) slices (the reference loops over n on the host) */
int t4k_tt_op(int op, const float *A, const float *B, float *O, int64_t hwc, int Na, int Nb, t4k_stream_t s);
/* k_copy via Tensor::copy (src/mu/tensor.cu:204-208) */
int t4k_copy(const float *src, float *dst, int64_t n, t4k_stream_t s);
/* k_transpose via Tensor::transpose (src/mu/tensor.cu:210-219): per sample, per channel 2-D transpose */
int t4k_transpose(const float *A, float *T, int N, int H, int W, int C, t4k_stream_t s);
/* k_identity via Tensor::identity (src/mu/tensor.cu:548-555) */
int t4k_identity(float *T, int N, int H, int W, int C, t4k_stream_t s);

/* ---- reductions: src/t4math.cu:23-131,248-365; results OVERWRITE *out (no pre-zero) --- */
int t4k_sum(const float *A, int64_t n, float *out, t4k_stream_t s);                 /* k_sum  tensor.cu:225-236 */
int t4k_nvar(const float *A, float avg, int64_t n, float *out, t4k_stream_t s);     /* k_nvar tensor.cu:244-259 */
int t4k_minmax(const float *A, int64_t n, int find_max, float *out, t4k_stream_t s);/* k_max  tensor.cu:261-277 */
/* mean and the reference's std = sqrt(Σ(x-μ)²)/n (src/mu/tensor.cu:238-251) in one call: out[0]=avg, out[1]=std */
int t4k_avg_std(const float *A, int64_t n, float *out2, t4k_stream_t s);
/* k_dot via Tensor::dot (src/mu/tensor.cu:61-72,279-287): O[n,c] = alpha*Σ_k A[n,k,c]B[n,k,c] + beta*O[n,c] */
int t4k_dot(const float *A, const float *B, float *O, float alpha, float beta,
            int K, int C, int Na, int Nb, t4k_stream_t s);
/* Tensor::loss (src/mu/tensor.cu:289-325, k_bce src/t4math.cu:248-274), NON-destructive on `out`
 * (the reference works on a duplicate, src/nn/loss.cpp:129-132): *loss = loss_kind(out,tgt)/N */
int t4k_loss(int kind, const float *out, const float *tgt, int64_t numel, int N, float *loss, t4k_stream_t s);
/* k_nan_inf via Tensor::has_nan (src/mu/tensor.cu:326-333): *cnt = #NaN + #Inf */
int t4k_nan_inf(const float *A, int64_t n, int *cnt, t4k_stream_t s);

/* ---- GEMM: src/t4math.cu:370-734 via Tensor::mm/linear/gemm1-4 (src/mu/tensor.cu:74-201) --
 * O[b][M,N,C] = alpha * op(A[b]) @ op(B[b]) + beta * O[b], channel interleaved (stride C),
 * op(A) = tA ? A[K,M,C] : A[M,K,C];  op(B) = tB ? B[N,K,C] : B[K,N,C].
 * batch slices are strideA/strideB/strideO floats apart (0 = broadcast, tensor.cu:175-177).
 * FP32 in, FP32 out.  Engine: tcgen05 3xTF32 (error-compensated split, FP32 accumulate in
 * TMEM) for large aligned C==1 problems, FP32-FMA SIMT otherwise. */
int t4k_gemm(const float *A, const float *B, float *O, float alpha, float beta, int tA, int tB,
             int M, int N, int K, int C, int batch, int64_t strideA, int64_t strideB, int64_t strideO,
             t4k_stream_t s);
int t4k_gemm_ex(int engine, const float *A, const float *B, float *O, float alpha, float beta, int tA, int tB,
             int M, int N, int K, int C, int batch, int64_t strideA, int64_t strideB, int64_t strideO,
             t4k_stream_t s);

/* tuning / test hook of the layer GEMM (T4K_GEMM_TL): what 0 = AUTO may take it (default 1; T4K_GEMM_TL=0), 1 = store the masked hi plane
 * explicitly instead of relying on kind::tf32 ignoring the 13 low mantissa bits (default 0), 2 = largest cluster size / split-K factor
 * (default 16).  Returns the previous value. */
int t4k_set_gemm_tl(int what, int value);
/* bring-up aid: when given a device buffer of 16 int64, CTA (0,0,0) of every following layer-GEMM launch stamps clock64() at its phases
 * (0 start, 1 set-up done, 2 first TMA issued, 3 first tile landed, 4 first lo plane done, 5 first MMA issued, 6 last TMA issued, 7 last tile
 * landed, 8 last MMA issued, 9 last accumulator ready, 10 tile parked, 11 cluster barrier passed, 12 first row reduced, 13 epilogue done,
 * 14 exit); NULL switches it off (the default) */
int t4k_gemm_tl_trace(long long *dev16);

/* ---- NN forward: src/nn/forward.cu + src/nn/nmath.cu/.tcu ---------------------------- */
/* k_bias (nmath.cu:27-35, forward.cu:195): Y[n,e] += B[e] */
int t4k_bias(const float *B, float *Y, int N, int E0, t4k_stream_t s);
/* Model::_flinear (forward.cu:158-198): Y[N,E0] = X[N,E1] @ W[E0,E1]^T + B[E0]  (GEMM + fused bias) */
int t4k_linear_fwd(const float *X, const float *W, const float *B, float *Y, int N, int E0, int E1, t4k_stream_t s);
/* _flinear + the following _factivate (forward.cu:158-209) in one pass (the bias, activation and mask ride in the GEMM's split-K finish):
 * Y = X @ W^T + B (the linear layer's output tensor), A = act(Y), F = saved derivative / mask; layer as t4k_activate_fwd */
int t4k_linear_act_fwd(int layer, const float *X, const float *W, const float *B, float *Y, float *A, float *F, float alpha,
                       int N, int E0, int E1, t4k_stream_t s);
/* classifier head, forward: small linear (E0 <= 32, W <= 40 KB) + bias + row softmax in one launch (forward.cu:158-198,231-243):
 * Y = X @ W^T + B, P = softmax(Y).  T4K_ENOSUP when the head is not small (caller: t4k_linear_fwd + t4k_softmax_fwd) */
int t4k_mlp_head_fwd(const float *X, const float *W, const float *B, float *Y, float *P, int N, int E0, int E1, t4k_stream_t s);
/* hidden linear + activation + classifier head (forward.cu:158-243) in two launches (GEMM, then split-K finish + bias + activation + small linear + softmax):
 * X [N,E1] -> Y1 = X @ W1^T + B1 [N,EH], A1 = act(Y1), F1 = saved derivative -> Y2 = A1 @ W2^T + B2 [N,E0], P = softmax(Y2) (+ Pdup).
 * Same tensors, same bits as t4k_linear_act_fwd followed by t4k_mlp_head_fwd.  T4K_ENOSUP when EH > 128, E0 > 32 or the layer is
 * not relu / tanh / sigmoid / selu / leakyrelu / elu: use the two calls. */
int t4k_linear_act_head_fwd(int layer, const float *X, const float *W1, const float *B1, float *Y1, float *A1, float *F1, float alpha,
                            const float *W2, const float *B2, float *Y2, float *P, float *Pdup, int N, int EH, int E1, int E0, t4k_stream_t s);
/* the same (forward.cu:158-198,231-243), with the probabilities also written to Pdup [N,E0] (may be NULL; the reference's Model::loss works on a
 * duplicate too, loss.cpp:129-132): Model::backprop turns P into p - y in place, so a
 * caller that wants the loss kernel to overlap the backward pass (second stream) lets it read the duplicate */
int t4k_mlp_head_fwd_dup(const float *X, const float *W, const float *B, float *Y, float *P, float *Pdup, int N, int E0, int E1, t4k_stream_t s);
/* k_activate (nmath.cu:37-70, forward.cu:201-209): writes O and the saved derivative/mask F.
 * layer in RELU,TANH,SIGMOID,SELU,LEAKYRL,ELU,DROPOUT; for DROPOUT F holds U(0,1] on entry. */
int t4k_activate_fwd(int layer, const float *I, float *O, float *F, float alpha, int64_t n, t4k_stream_t s);
/* k_softmax_small/k_softmax (nmath.cu:74-169, forward.cu:231-243): row softmax, rows of C */
int t4k_softmax_fwd(const float *I, float *O, int N, int C, t4k_stream_t s);
/* Model::_flogsoftmax AS CODED (forward.cu:246-259): O = exp(I) - log10(max(Σ_row exp(I),1e-6)) */
int t4k_logsoftmax_fwd(const float *I, float *O, int N, int C, t4k_stream_t s);
/* k_conv2d<TS,KS,S,P> (nmath.tcu:34-104, forward.cu:126-155).  F is [C1,KS,KS,C0], B is [C0].
 * Writes EVERY element of O (no pre-zero needed, unlike forward.cu:138).
 * (KS,S,P) in {(1,1,0),(3,1,1),(4,2,1),(5,1,2)} as in the reference, else T4K_ENOSUP. */
int t4k_conv2d_fwd(const float *I, const float *F, const float *B, float *O,
                   int N, int H1, int W1, int C1, int H0, int W0, int C0, int KS, int S, int P, t4k_stream_t s);
/* engine selection for conv2d fwd/bwd (tests, benchmarks): T4K_GEMM_AUTO (default: tcgen05 implicit GEMM when the
 * shape is GEMM-sized — stride 1, "same" padding, C1 % 32 == 0, 16 <= C0 <= 128, C0 % 16 == 0 — else CUDA cores),
 * T4K_GEMM_SIMT (never the tensor path), T4K_GEMM_TC (tensor path or T4K_EINVAL) */
int t4k_set_conv_engine(int engine);
/* k_pool<KS> (nmath.tcu:122-186, forward.cu:212-228; also upsample-backward backprop.cu:285-300)
 * layer in AVGPOOL,MAXPOOL,MINPOOL,USAMPLE; KS in {2,3}; stride == KS */
int t4k_pool_fwd(int layer, const float *I, float *O, int N, int H1, int W1, int H0, int W0, int C, int KS, t4k_stream_t s);
/* k_batchnorm_1/2/3 (nmath.cu:177-264, forward.cu:264-309).  scratch3C = mtum[4] layout:
 * [0,C) rvar = 1/(sqrt(max(var,0))+1e-6), [C,2C) mean, [2C,3C) unused in forward. */
int t4k_batchnorm_fwd(const float *I, float *O, float *XH, const float *gamma, const float *beta,
                      float *scratch3C, int N, int HW, int C, t4k_stream_t s);

/* ---- NN backward: src/nn/backprop.cu ------------------------------------------------- */
/* k_dlinear_db (nmath.cu:274-280, backprop.cu:239): dB[e] += Σ_n dY[n,e] */
int t4k_dbias(const float *dY, float *dB, int N, int E0, t4k_stream_t s);
/* Model::_blinear (backprop.cu:194-254): if train { dB += ΣdY; dW += dY^T@X }; dX = dY@W.
 * dX may alias X's buffer?  NO — dW needs X; pass distinct buffers or dX==X (handled: dW first). */
int t4k_linear_bwd(const float *X, const float *W, const float *dY, float *dX, float *dW, float *dB,
                   int N, int E0, int E1, int train, t4k_stream_t s);
/* as t4k_linear_bwd (backprop.cu:194-254); skip_db != 0 leaves dB alone (it was accumulated by t4k_mlp_head_bwd) */
int t4k_linear_bwd_ex(const float *X, const float *W, const float *dY, float *dX, float *dW, float *dB,
                      int N, int E0, int E1, int train, int skip_db, t4k_stream_t s);
/* _blinear followed by the _bactivate of the activation layer in front of it (backprop.cu:194-263) with the mask multiply in the dX
 * GEMM's epilogue: dX = dY @ W (the linear layer's input tensor = the activation's output tensor), dXprev = dX * Fprev (the activation's
 * input tensor); dW/dB as t4k_linear_bwd_ex.  Same tensors written as the two calls. */
int t4k_linear_bwd_act(const float *X, const float *W, const float *dY, float *dX, float *dW, float *dB,
                       const float *Fprev, float *dXprev, int N, int E0, int E1, int train, int skip_db, t4k_stream_t s);
/* dX of the hidden linear layer straight from the classifier head's FORWARD tensors — off the critical path goes t4k_mlp_head_bwd:
 *   dX[N,E1] = A @ W1,  A[n][e] = (Σ_k (P[n][k] - T[n][k]) * W2[k][e]) * F1[n][e]      (W2 [E2,EH], W1 [EH,E1], F1 [N,EH] or NULL)
 * i.e. Model::_bprep, the small linear's dX and the activation backward (backprop.cu:76-140,194-263) are evaluated inside the GEMM's
 * operand producer (same arithmetic, same order as t4k_mlp_head_bwd: the same bits as the dX GEMM run on its stored output).  Reads only
 * tensors the backward pass never overwrites when P is the duplicate of the softmax output (t4k_mlp_head_fwd_dup), so it can run
 * concurrently with t4k_mlp_head_bwd.  T4K_ENOSUP: E2 > 32, EH > 128, EH % 4, or a shape the layer GEMM does not take. */
int t4k_linear_dx_from_head(const float *P, const float *T, const float *W2, const float *F1, const float *W1, float *dX,
                            int N, int E2, int EH, int E1, t4k_stream_t s);
/* the same (backprop.cu:76-140,194-263) for BOTH products of the hidden layer, ONE launch: dX = A @ W1 and dW1 += A^T @ X with A generated K-major for the one and
 * M-major for the other (two problems share the grid of the layer GEMM).  X must not alias dX (Model::backprop passes the flatten layer's
 * duplicate of X).  T4K_ENOSUP: as above, or the two problems do not fit one co-resident wave. */
int t4k_linear_bwd_from_head(const float *P, const float *T, const float *W2, const float *F1, const float *X, const float *W1,
                             float *dX, float *dW1, int N, int E2, int EH, int E1, t4k_stream_t s);
/* THE TRAIN TAIL (inside a fused train step only: the target is known at forward time and the forward values of the tail's layer tensors
 * are never observable): hidden linear + activation + classifier head + softmax AND, on the same rows, Model::_bprep, the head linear's dX
 * and the activation backward (forward.cu:158-243, backprop.cu:76-140,194-263) in ONE launch of the layer GEMM.  On return the layer tensors
 * hold what they hold after t4k_linear_act_head_fwd + t4k_mlp_head_bwd: Y1 = dY1, A1 = dX of the head linear, F1 = mask, Ylin = P = p - y,
 * Pdup = p (for the loss); the head's parameter gradients are left as per-CTA partials in `scratch` (t4k_head_train_scratch_floats floats;
 * 0 = shape not supported) for t4k_head_grad_finish(scratch, *ncta, ...) to add into dW2 / dB2 / dB1 — on any stream, off the critical path. */
int64_t t4k_head_train_scratch_floats(int layer, int N, int EH, int E1, int E0);
int t4k_linear_act_head_train(int layer, const float *X, const float *W1, const float *B1, float *Y1, float *A1, float *F1, float alpha,
                              const float *W2, const float *B2, float *Ylin, float *P, float *Pdup, const float *T,
                              float *scratch, int *ncta, int N, int EH, int E1, int E0, t4k_stream_t s);
int t4k_head_grad_finish(const float *scratch, int ncta, int E0, int EH, float *dW2, float *dB2, float *dB1, t4k_stream_t s);
/* dX = dY @ W and dW += dY^T @ X of Model::_blinear (backprop.cu:239-247) in ONE launch (two problems share the layer GEMM's grid); X must not alias dX.
 * T4K_ENOSUP when a product is outside the layer GEMM's class or the pair does not fit one co-resident wave: use t4k_linear_bwd_ex */
int t4k_linear_bwd_pair(const float *X, const float *W, const float *dY, float *dX, float *dW, int N, int E0, int E1, t4k_stream_t s);
/* classifier head, backward, one launch (backprop.cu:76-140,194-263), E0 <= 32, E1 <= 128 else T4K_ENOSUP:
 *   P <- P - T (Model::_bprep), Ylin <- P - T (softmax backward is a copy), dB += Σ_n (P-T), dW += (P-T)^T @ X2,
 *   X2 <- (P-T) @ W (in place: the small linear's input tensor receives its dX),
 *   if F1: Y1 <- X2 * F1 (backward of the activation in front), dB1 (may be NULL) += Σ_n Y1 (bias gradient of the linear
 *   in front of that activation; without F1 it accumulates Σ_n X2).  dW/dB/dB1 only when train. */
int t4k_mlp_head_bwd(float *P, const float *T, float *Ylin, float *X2, const float *F1, float *Y1, const float *W,
                     float *dW, float *dB, float *dB1, int N, int E0, int E1, int train, t4k_stream_t s);
/* Model::_bactivate (backprop.cu:257-263): dX = dY * F */
int t4k_activate_bwd(const float *dY, const float *F, float *dX, int64_t n, t4k_stream_t s);
/* k_dconv2d (nmath.tcu:211-338, backprop.cu:153-191): dX written in full (no pre-zero needed),
 * uses the reference's 180°-FLIPPED filter taps for dX (nmath.tcu:304); if train: dF += , dB += . */
int t4k_conv2d_bwd(const float *I, const float *dO, const float *F, float *dX, float *dF, float *dB,
                   int N, int H1, int W1, int C1, int H0, int W0, int C0, int KS, int S, int P,
                   int train, t4k_stream_t s);
/* conv-transpose layer (L_DCONV; word `dconv2d`: 4x4, stride 2, padding 1).  The reference wires it as the convolution layer with the two
 * kernels' roles swapped (src/nn/forward.cu:110 -> Model::_bconv, src/nn/backprop.cu:137 -> Model::_fconv; output shape src/nn/model.cpp:129-133):
 *   forward   O [N,H0,W0,C0] = k_dconv2d's input-gradient half applied to I [N,H1,W1,C1] (flipped taps, nmath.tcu:304), + bias per output channel
 *   backward  dX [N,H1,W1,C1] = k_conv2d(dO, F) without bias;  train: dF += k_dconv2d's filter-gradient half on (input, output gradient) = (dO, I),
 *             dB[c0] += sum of dO over the pixels
 * F [C0][K][K][C1]: the filter of the convolution (C0 -> C1, K, S, P) that maps the large image onto the small one, (H0 - K + 2P)/S + 1 == H1.
 * I and dX must not alias (dF reads I). */
int t4k_dconv2d_fwd(const float *I, const float *F, const float *B, float *O,
                    int N, int H1, int W1, int C1, int H0, int W0, int C0, int KS, int S, int P, t4k_stream_t s);
int t4k_dconv2d_bwd(const float *I, const float *dO, const float *F, float *dX, float *dF, float *dB,
                    int N, int H1, int W1, int C1, int H0, int W0, int C0, int KS, int S, int P, int train, t4k_stream_t s);
/* k_dpool<KS> (nmath.tcu:475-568, backprop.cu:266-280; also upsample-forward forward.cu:314-329):
 * IN PLACE on the forward input I: max/min → zero window, dO at first strict max/min. */
int t4k_pool_bwd(int layer, float *I, const float *dO, int N, int H1, int W1, int H0, int W0, int C, int KS, t4k_stream_t s);
/* k_dbatchnorm_1/2/3 (nmath.cu:295-414, backprop.cu:312-370): scratch3C as in forward
 * ([C,2C) and [2C,3C) receive mean(dy), mean(dy*xhat)); if train: dgamma += mean(dy*xhat), dbeta += mean(dy) */
int t4k_batchnorm_bwd(const float *dO, const float *XH, float *dX, const float *gamma,
                      float *dgamma, float *dbeta, float *scratch3C, int N, int HW, int C, int train, t4k_stream_t s);
/* data parallel (SURVEY.md §8e collective 2): the same two layers over a batch SHARDED across the ranks of `comm` — the per-channel sums
 * (Σx, Σx² forward; Σdy, Σdy·x̂ backward; 2C floats) are SUM-all-reduced between the statistics pass and the apply pass, N_global = samples of
 * the whole batch.  dgamma / dbeta receive this rank's share (local sum / global rows): the gradient exchange adds the shares up to the
 * reference's means (nmath.cu:378-381).  `comm`: a communicator of its own, capacity >= 4C (not the gradient arena's). */
int t4k_batchnorm_fwd_dp(t4k_comm_t comm, const float *I, float *O, float *XH, const float *gamma, const float *beta,
                         float *scratch3C, int N, int N_global, int HW, int C, t4k_stream_t s);
int t4k_batchnorm_bwd_dp(t4k_comm_t comm, const float *dO, const float *XH, float *dX, const float *gamma,
                         float *dgamma, float *dbeta, float *scratch3C, int N, int N_global, int HW, int C, int train, t4k_stream_t s);

/* ---- fused CNN block: conv2d → maxpool(2) → relu (→ flatten), the layer group of examples/t4_40a.4th:11-12.
 * One launch each way; writes exactly the layer tensors the per-layer calls write (forward.cu:83-155,201-228;
 * backprop.cu:112-191,257-280).  Eligible: C1 <= 4, C0 <= 16, even H0/W0, sample fits in shared memory; otherwise
 * T4K_ENOSUP and the caller issues the per-layer calls.  flatO may be NULL (no flatten layer).  Icopy (may be NULL or == I)
 * receives a copy of I: the `n0 = input` of Model::forward (forward.cu:36-43) folded into the same launch.
 * backward: dY = gradient at the block output (the flatten output tensor, or actO itself when there is no flatten);
 * actO <- dY, poolO <- dY*actF, convO (forward conv output) <- max-pool routed gradient in place,
 * Iio (conv input) <- dX and dXbuf <- dX (Model::_bconv: `in = dx`), dF/dB += when train. */
int t4k_conv_pool_relu_fwd(const float *I, const float *F, const float *B, float *Icopy, float *convO, float *poolO, float *actO,
                           float *actF, float *flatO, int N, int H1, int W1, int C1, int H0, int W0, int C0, int KS, int S, int P,
                           t4k_stream_t s);
/* the forward block fed from a staged U8 mini-batch: Dataset::_load (src/mu/dataset.cu:124-152: d = ((float)u8 - mean) * scale) +
 * Model::onehot (src/nn/loss.cpp:47-72) + the block in ONE launch.  data = the dataset tensor [N,H1,W1,C1] (first feedN samples
 * rewritten, the rest kept: partial batch as _load), Icopy = the model's input layer, lab32/hot as t4k_dataset_load.  Shapes
 * outside the fused envelope run t4k_dataset_load then t4k_conv_pool_relu_fwd (same results). */
int t4k_conv_pool_relu_fwd_feed(const uint8_t *u8I, const uint8_t *u8L, int feedN, float mean, float scale, int32_t *lab32,
                                float *hot, int E, float *data, const float *F, const float *B, float *Icopy, float *convO,
                                float *poolO, float *actO, float *actF, float *flatO, int N, int H1, int W1, int C1,
                                int H0, int W0, int C0, int KS, int S, int P, t4k_stream_t s);
int t4k_conv_pool_relu_bwd(const float *dY, float *actO, const float *actF, float *poolO, float *convO, float *Iio, float *dXbuf,
                           const float *F, float *dF, float *dB, int N, int H1, int W1, int C1, int H0, int W0, int C0,
                           int KS, int S, int P, int train, t4k_stream_t s);
/* one-shot: the next t4k_conv_pool_relu_bwd[_opt] of this thread records `event` (a cudaEvent_t) between its main kernel and its finish launch —
 * a place to hang side-stream work that may overlap the short finish launch but must not take SMs from the main kernel */
int t4k_conv_pool_relu_bwd_mid_event(void *event);

/* ---- optimizers: src/nn/gradient.cu:133-169 + nmath.cu:419-472 ------------------------ */
int t4k_sgd(float *G, float *DG, float *M, int Nw, float lr, float b, int64_t n, t4k_stream_t s);
int t4k_adam(float *G, float *DG, float *M, float *V, float lr, float b1, float b2, int64_t n, t4k_stream_t s);
int t4k_adamw(float *G, float *DG, float *M, float *V, float lr, float b1, float b2, float wd, int64_t n, t4k_stream_t s);
/* one launch for a whole model: `seg` is a DEVICE array of nseg segments over flat arenas
 * G/DG/M/V (same offsets in each); kind 0=sgd 1=adam 2=adamw.  Replaces the per-tensor launch
 * loop of Model::gradient (gradient.cu:99-121). */
typedef struct { int64_t off; int64_t len; int32_t Nw; int32_t pad; } t4k_seg_t;
int t4k_optim_multi(int kind, float *G, float *DG, float *M, float *V, const t4k_seg_t *seg, int nseg,
                    int64_t total, float lr, float b1, float b2, float wd, t4k_stream_t s);

/* the same over the arena elements [from, to) only (absolute indices: `seg` stays the whole table).  Lets a caller run the optimizer of the
 * layers whose gradients are final early, on a second stream, while backprop still works on the first layers */
int t4k_optim_multi_range(int kind, float *G, float *DG, float *M, float *V, const t4k_seg_t *seg, int nseg,
                          int64_t from, int64_t to, float lr, float b1, float b2, float wd, t4k_stream_t s);
/* optimizer step applied by the kernel that FINISHES a gradient (the conv block's dF/dB reduction): G/M/V are the arena bases, offF/offB
 * the offsets of the filter / bias segments (dF = DG + offF, dB = DG + offB), Nw their t4k_seg_t.Nw; same arithmetic as t4k_optim_multi */
typedef struct { int kind; float lr, b1, b2, wd; float *G, *M, *V; int64_t offF, offB; int32_t NwF, NwB; } t4k_fused_opt_t;
/* t4k_conv_pool_relu_bwd with `opt` (may be NULL = plain): the last launch of the block (per-sample partials -> dF, dB) also applies the
 * optimizer step to the filter and bias and zeroes dF, dB — one launch less at the end of a train step */
int t4k_conv_pool_relu_bwd_opt(const float *dY, float *actO, const float *actF, float *poolO, float *convO, float *Iio, float *dXbuf,
                               const float *F, float *dF, float *dB, int N, int H1, int W1, int C1, int H0, int W0, int C0,
                               int KS, int S, int P, int train, const t4k_fused_opt_t *opt, t4k_stream_t s);

/* ---- data-parallel extras (SURVEY.md §8b "DP extras", §8e) --------------------------------
 * The reference is single-GPU; Model::forward/backprop shard over the batch and the parameter gradients are batch
 * sums (src/nn/backprop.cu:97-103), so a SUM all-reduce of the flat DG arena in front of the optimizer loop of
 * Model::gradient (src/nn/gradient.cu:99-121) is the whole exchange.  One process per GPU; every rank creates a
 * communicator (an exchange block in its own HBM, exported as a 64-byte cudaIpc handle), the host gathers the
 * handles (torch.distributed / MPI / anything) and connects.  The exchange itself is one kernel per call over NVLink
 * peer stores (tensorforth_b200/csrc/comm.cu), CUDA-graph capturable; sums are taken in rank order, so every rank
 * holds bit-identical results.  Every rank must issue the same sequence of calls with the same lengths. */
#define T4K_COMM_HANDLE_BYTES 64
int t4k_comm_create(int rank, int world, int64_t cap_floats, t4k_comm_t *out, void *handle64);
int t4k_comm_connect(t4k_comm_t c, const void *handles /* world x T4K_COMM_HANDLE_BYTES, rank order */);
int t4k_comm_connect_local(t4k_comm_t c, t4k_comm_t *all /* world communicators of THIS process, rank order */);
int t4k_comm_destroy(t4k_comm_t c);
int t4k_comm_status(t4k_comm_t c);      /* 0 healthy; k>0: a wait for rank k-1 timed out (T4K_COMM_TIMEOUT_S seconds, default 60, without progress);
                                         * synchronises the device.  The error is STICKY: the chunk that timed out is not finished (no optimizer
                                         * step on a half-summed gradient) and every later exchange on this communicator is a no-op */
int t4k_comm_poll(t4k_comm_t c);        /* the same word read from mapped host memory, no synchronisation (cheap enough for every step) */
int64_t t4k_comm_capacity(t4k_comm_t c);
int t4k_comm_scalar_mirror(t4k_comm_t c, float *pinned);   /* exchanges launched / captured from now on also store the summed scalars into this page-locked host
                                                             * buffer (unified addressing): the host reads the global loss with no copy node behind the step; NULL = off */
int t4k_shard_info(int64_t n, int world, int rank, int64_t *lo, int64_t *hi);   /* batch shard of a rank: samples [lo, hi) (no device needed) */
/* buf[i] = sum over ranks of buf[i], in place, n <= capacity */
int t4k_allreduce_sum(t4k_comm_t c, float *buf, int64_t n, t4k_stream_t s);
/* t4k_optim_multi on the rank-summed gradient: DG is exchanged, summed and consumed (zeroed) by the same kernel;
 * `scal[0..nscal)` (device, nscal <= 64: loss sums, hit counts) are sum-all-reduced in place in the same exchange.
 * `pushed_from`: value returned by a t4k_dp_push of THIS step (0 or total: nothing was pushed early). */
int t4k_optim_multi_dp(t4k_comm_t c, int kind, float *G, float *DG, float *M, float *V, const t4k_seg_t *seg, int nseg,
                       int64_t total, float lr, float b1, float b2, float wd, float *scal, int nscal, int64_t pushed_from,
                       t4k_stream_t s);
/* the same on the chunks of the arena that START in [from, to) (chunk k starts at k * t4k_comm_chunk_floats(c)): a step may run the exchange +
 * optimizer of the part whose gradients are final early on a side stream, under the rest of backprop, and only the first layers' chunks at the
 * end (chunks carry their own epochs; the scalars ride with chunk 0; every rank splits at the same offsets) */
int t4k_optim_multi_dp_range(t4k_comm_t c, int kind, float *G, float *DG, float *M, float *V, const t4k_seg_t *seg, int nseg,
                             int64_t from, int64_t to, int64_t total, float lr, float b1, float b2, float wd, float *scal, int nscal,
                             int64_t pushed_from, t4k_stream_t s);
int64_t t4k_comm_chunk_floats(t4k_comm_t c);
/* early half of a split exchange: push (and signal) the chunks of DG[0..total) that START at or beyond float `from` —
 * the gradient segments that are already final while backprop still runs (gradients are produced last layer first, and
 * the arena is laid out first layer first).  Returns the float offset of the first pushed chunk (pass it to
 * t4k_optim_multi_dp as `pushed_from`; == total when nothing qualified), negative on error.  Exactly one
 * t4k_optim_multi_dp must follow before the next push; no other exchange on this communicator in between. */
int64_t t4k_dp_push(t4k_comm_t c, const float *DG, int64_t from, int64_t total, t4k_stream_t s);
/* reduce-scatter flavour for more than two ranks: chunk k has ONE owner (rank k % world).  t4k_dp_push_owner sends every chunk that starts at or beyond
 * `from` to its owner only (1x the arena leaves the GPU instead of (world-1)x); t4k_optim_multi_dp_rs — one block per chunk on every rank — lets the owner
 * sum the world pushes in rank order and store the SUM into every rank, after which every rank runs the optimizer on the chunk (state stays replicated).
 * Two NVLink hops instead of one: for the part of the arena whose exchange is off the step's critical path. */
int64_t t4k_dp_push_owner(t4k_comm_t c, const float *DG, int64_t from, int64_t total, t4k_stream_t s);
int t4k_optim_multi_dp_rs(t4k_comm_t c, int kind, float *G, float *DG, float *M, float *V, const t4k_seg_t *seg, int nseg,
                          int64_t from, int64_t total, float lr, float b1, float b2, float wd, int phase, t4k_stream_t s);
/* phase 0: both halves in one launch.  phase 1: the OWNERS' half only (wait for the pushes, sum, send the sums: ceil(chunks / world) blocks per
 * rank — small enough to run next to backprop, right behind the push); phase 2: every rank's half (wait for the sums, optimizer). */
/* the same push with the data moved by the copy engines (peer-to-peer cudaMemcpyAsync) and a one-block kernel raising the flags: no SM is taken
 * from the backward kernels it overlaps.  `step` = exchanges this communicator has completed so far (selects the slot parity the copies are
 * addressed with; a captured step is captured once per parity). */
int64_t t4k_dp_push_dma(t4k_comm_t c, const float *DG, int64_t from, int64_t total, uint32_t step, t4k_stream_t s);

/* ---- RNG: src/util.cu:35-70 via System::rand (src/sys.cpp:77-95) ---------------------- */
/* d[i] = scale * (bias + x), x ~ U(0,1] or N(0,1).  Counter-based Philox4x32-10 keyed by
 * (seed, element index): reproducible and independent of grid size / GPU count (the reference's
 * XORWOW stream is seeded with time(), so bit parity is not defined). */
int t4k_rand_seed(uint64_t seed);
int t4k_rand(float *d, int64_t n, int opt, float bias, float scale, t4k_stream_t s);
/* this rank's shard [before, before + n) of a batch-major tensor of global_n elements, drawn with the counters a single device holding the
 * whole tensor would use (per-rank Philox offset = global element index); every rank advances the stream by global_n */
int t4k_rand_sharded(float *d, int64_t n, int64_t before, int64_t global_n, int opt, float bias, float scale, t4k_stream_t s);
/* dropout forward in ONE launch (Model::_fstep L_DROPOUT, src/nn/forward.cu:98-102 + k_activate, nmath.cu:66-68): draws the mask t4k_rand_sharded(F, n,
 * before, global_n, UNIFORM) would draw (before = 0, global_n = n on one device), F <- (u > rate) ? 1 : 0, O <- I where kept else 0 (no rescaling) */
int t4k_dropout_fwd(const float *I, float *O, float *F, float rate, int64_t n, int64_t before, int64_t global_n, t4k_stream_t s);
int t4k_rand_at(float *d, int64_t n, int opt, float bias, float scale, uint64_t seed, uint64_t offset, t4k_stream_t s);
/* CUDA-graph replays: a captured t4k_rand has its seed and offset baked in; t4k_rand adds a device-side replay epoch (x 2^40) to its
 * counter, and t4k_rand_tick — put once at the head of a captured sequence that draws — advances it, so that every replay draws
 * fresh numbers (dropout masks, latent batches).  Never ticked outside graphs. */
int t4k_rand_tick(t4k_stream_t s);

/* ---- dataset feeding: Dataset::_load (src/mu/dataset.cu:124-152), SURVEY.md §8f row 2 -------------------
 * The reference converts each mini-batch on the host (one float per U8 pixel, then an H2D of 4 bytes per pixel).
 * Here the U8 block is what crosses PCIe; the device does dst[i] = ((float)src[i] - mean) * scale (two roundings,
 * bit-equal to the host loop), widens the U8 labels to int32 and — when `hot` is given — writes their one-hot rows
 * [nlab, E] (Model::onehot(Dataset&), src/nn/loss.cpp:47-72: label >= E -> class 0), all in one launch.
 * mean/scale as Dataset::normalize stores them (scale = 1/given, default mean 0, scale 1/256: dataset.h:35-36). */
int t4k_dataset_load(const uint8_t *src_u8, float *dst, int64_t n, float mean, float scale,
                     const uint8_t *label_u8, int32_t *label_i32, int nlab, float *hot, int E, t4k_stream_t s);

/* ---- loss-side host loops moved on device: src/nn/loss.cpp:47-107 ---------------------- */
/* Model::onehot(Dataset&): hot[n, label<E ? label : 0] = 1, everything else 0; labels are device int32 */
int t4k_onehot(const int32_t *label, float *hot, int N, int E, t4k_stream_t s);
/* Model::hit: *cnt = Σ_n (int)hot[n, argmax_first(out[n])] */
int t4k_hit(const float *out, const float *hot, int N, int E, int *cnt, t4k_stream_t s);
/* fused softmax-output step: p - y in place (Model::_bprep, backprop.cu:97-103) is t4k_tt_op(SUB) */

#ifdef __cplusplus
}
#endif
#endif /* T4K_H */
