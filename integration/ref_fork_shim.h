/*
 * ref_fork_shim.h — the reference's launch seam redirected to libt4k.so.
 *
 * tensorForth launches every kernel through the FORK* macros (src/t4base.h:129-159, src/t4math.h:117-123).  This
 * header is PRE-INCLUDED (nvcc -include) when the reference's own, unmodified src/mu/tensor.cu, src/nn/gradient.cu and
 * src/nn/debug.cu are compiled for the `ten4_b200` binary: it pulls in the reference headers first (their include
 * guards then keep the original macro definitions from coming back), and re-defines FORK / FORK1 / FORK2 / FORK3 /
 * FORK3T so that `FORK(k_xxx, n, args...)` becomes a call of the C-ABI entry point of include/t4k.h that replaces
 * k_xxx.  Argument orders below are the reference's (src/t4math.h:138-184, src/nn/nmath.h:41-112).
 * Kernels outside the hot-path scope (Gauss-Jordan / LU / det) keep the reference's own launch.
 * Host-synchronous semantics are kept exactly as the reference has them: GPU_CHK() after every call
 * (src/ten4_types.h:192) — the VM words read results on the host right after (src/mu/tensor.cu:231-233).
 * The Model path (forward / backprop) does NOT go through here: integration/model_shim.cu.
 */
#pragma once
#include "ten4_types.h"
#include "t4base.h"
#include "t4math.h"
#include "nn/nmath.h"
#include "../include/t4k.h"

#define T4K_CALL(call) { int _rc = (call);                                                        \
    if (_rc > 0) { GPU_ERR((cudaError_t)_rc); }                                                    \
    else if (_rc < 0) { ERROR("%s -> %s\n", #call, t4k_strerror(_rc)); }                           \
    GPU_CHK(); }
/* the reference's launch convention, for the kernels that stay on the reference side */
#define REF_FORK(fn,n,...)    { const dim3 _b(T4_DIM_SQ, 1, 1); const dim3 _g(GRID_BLKS(n), 1, 1); fn<<<_g,_b>>>(__VA_ARGS__, n); GPU_CHK(); }
#define REF_FORK2(fn,_g,n,...) { fn<<<_g,T4_DIM_SQ>>>(__VA_ARGS__, n); GPU_CHK(); }
#define REF_FORK3(fn,h,w,c,...) { const dim3 _b(T4_DIM_SZ, T4_DIM_SZ, 1); const dim3 _g(((w) + _b.x - 1) / _b.x, ((h) + _b.y - 1) / _b.y, c); \
                                  fn<<<_g,_b>>>(__VA_ARGS__, h, w); GPU_CHK(); }
#undef FORK
#undef FORK1
#undef FORK2
#undef FORK3
#undef FORK3T
#define FORK(fn,n,...)        T4SHIM_##fn(n, __VA_ARGS__)
#define FORK1(fn,c,n,...)     T4SHIM1_##fn(c, n, __VA_ARGS__)
#define FORK2(fn,g,n,...)     T4SHIM2_##fn(g, n, __VA_ARGS__)
#define FORK3(fn,h,w,c,...)   T4SHIM3_##fn(h, w, c, __VA_ARGS__)
#define FORK3T(fn,h,w,c,...)  T4SHIM3_##fn(h, w, c, __VA_ARGS__)

/* ---- src/t4math.h:138-150: elementwise, reductions ---- */
#define T4SHIM_k_ts_op(n, op, A, v, O)          T4K_CALL(t4k_ts_op(op, A, v, O, (int64_t)(n), 0))
#define T4SHIM_k_tt_op(n, op, A, B, O)          T4K_CALL(t4k_tt_op(op, A, B, O, (int64_t)(n), 1, 1, 0))
#define T4SHIM_k_math(n, op, A, v)              T4K_CALL(t4k_map(op, A, v, (int64_t)(n), 0))
#define T4SHIM_k_copy(n, S, D)                  T4K_CALL(t4k_copy(S, D, (int64_t)(n), 0))
#define T4SHIM_k_sum(n, S, out)                 T4K_CALL(t4k_sum(S, (int64_t)(n), out, 0))
#define T4SHIM_k_nvar(n, S, avg, out)           T4K_CALL(t4k_nvar(S, avg, (int64_t)(n), out, 0))
#define T4SHIM_k_max(n, S, out, find_max)       T4K_CALL(t4k_minmax(S, (int64_t)(n), (find_max) ? 1 : 0, out, 0))
#define T4SHIM_k_nan_inf(n, S, cnt)             T4K_CALL(t4k_nan_inf(S, (int64_t)(n), cnt, 0))
/* k_bce leaves +Σ[t ln(o+ε) + (1-t) ln(1-o+ε)] in *out (src/t4math.cu:248-274); t4k_loss(BCE, N=1) is its negative */
#define T4SHIM_k_bce(n, T, O, out)              { T4K_CALL(t4k_loss(T4K_LOSS_BCE, O, T, (int64_t)(n), 1, out, 0)); T4K_CALL(t4k_map(T4K_NEG, out, 0.0f, 1, 0)); }
/* ---- BLAS: k_dot (grid (C,1)), the four GEMM variants ---- */
#define T4SHIM1_k_dot(c, n, A, B, O, alpha, beta, K, C)  T4K_CALL(t4k_dot(A, B, O, alpha, beta, K, C, 1, 1, 0))
#define T4SHIM3_k_gemm(h, w, c, A, B, O, alpha, beta, tA, tB, K)                 T4K_CALL(t4k_gemm_ex(T4K_GEMM_SIMT, A, B, O, alpha, beta, tA, tB, h, w, K, c, 1, 0, 0, 0, 0))
#define T4SHIM3_k_gemm_claude(h, w, c, A, B, O, alpha, beta, tA, tB, K)          T4K_CALL(t4k_gemm_ex(T4K_GEMM_SIMT, A, B, O, alpha, beta, tA, tB, h, w, K, c, 1, 0, 0, 0, 0))
#define T4SHIM3_k_gemm_tile_claude(h, w, c, A, B, O, alpha, beta, tA, tB, K)     T4K_CALL(t4k_gemm(A, B, O, alpha, beta, tA, tB, h, w, K, c, 1, 0, 0, 0, 0))
#define T4SHIM3_k_gemm_tile_claude_x2(h, w, c, A, B, O, alpha, beta, tA, tB, K)  T4K_CALL(t4k_gemm(A, B, O, alpha, beta, tA, tB, h, w, K, c, 1, 0, 0, 0, 0))
#define T4SHIM3_k_transpose(h, w, c, S, D)      T4K_CALL(t4k_transpose(S, D, 1, h, w, c, 0))
#define T4SHIM3_k_identity(h, w, c, T)          T4K_CALL(t4k_identity(T, 1, h, w, c, 0))
/* ---- optimizers (src/nn/gradient.cu:133-169; nmath.h:96-110): Nw = g.N() ---- */
#define T4SHIM_k_sgd(n, G, DG, M, Nw, lr, b)             T4K_CALL(t4k_sgd(G, DG, M, Nw, lr, b, (int64_t)(n), 0))
#define T4SHIM_k_adam(n, G, DG, M, V, Nw, lr, b1, b2)    T4K_CALL(t4k_adam(G, DG, M, V, lr, b1, b2, (int64_t)(n), 0))
#define T4SHIM_k_adamw(n, G, DG, M, V, Nw, lr, b1, b2, wd) T4K_CALL(t4k_adamw(G, DG, M, V, lr, b1, b2, wd, (int64_t)(n), 0))
/* ---- out of scope: matrix inversion / LU / det stay on the reference's kernels (src/t4math.cu:742-979) ---- */
#define T4SHIM2_k_find_pivot(g, n, ...)         REF_FORK2(k_find_pivot, g, n, __VA_ARGS__)
#define T4SHIM2_k_logdet(g, n, ...)             REF_FORK2(k_logdet, g, n, __VA_ARGS__)
#define T4SHIM_k_swap_rows(n, ...)              REF_FORK(k_swap_rows, n, __VA_ARGS__)
#define T4SHIM_k_diag(n, ...)                   REF_FORK(k_diag, n, __VA_ARGS__)
#define T4SHIM_k_elim(n, ...)                   REF_FORK(k_elim, n, __VA_ARGS__)
#define T4SHIM_k_lu_col(n, ...)                 REF_FORK(k_lu_col, n, __VA_ARGS__)
#define T4SHIM_k_pivot(n, ...)                  REF_FORK(k_pivot, n, __VA_ARGS__)
#define T4SHIM_k_fsub(n, ...)                   REF_FORK(k_fsub, n, __VA_ARGS__)
#define T4SHIM_k_bsub(n, ...)                   REF_FORK(k_bsub, n, __VA_ARGS__)
#define T4SHIM3_k_lu(h, w, c, ...)              REF_FORK3(k_lu, h, w, c, __VA_ARGS__)
