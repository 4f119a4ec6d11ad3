// optim.cuh — the per-element optimizer step shared by the optimizer kernels (elementwise.cu) and the
// fused all-reduce + optimizer kernel (comm.cu)
#pragma once
#include "common.cuh"
namespace t4k {
// ------------------------------------------------------------------ optimizers (one pass: read g,dg,m,v / write g,dg=0,m,v)
struct OptP { float lr, b1, b2, wd; };
template<int KIND> __device__ __forceinline__ void opt_step(float &g, float &dg, float &m, float &v, float invN, bool mom, OptP p) {
    if (KIND == 0) {                                        // k_sgd (nmath.cu:419-436)
        float d = dg * invN;                                // dg / Nw: Nw is a small power-of-two-free int; see launcher
        if (!mom) g -= p.lr * d;
        else { m = p.b1 * m + (1.0f - p.b1) * d; g -= p.lr * m; }
    } else {                                                // k_adam / k_adamw (nmath.cu:438-472)
        m = p.b1 * m + (1.0f - p.b1) * dg;
        v = p.b2 * v + (1.0f - p.b2) * dg * dg;
        if (KIND == 1) g -= p.lr * m / (__fsqrt_rn(v) + DU_EPS);
        else           g -= p.lr * (m / (__fsqrt_rn(v) + DU_EPS) - p.wd * dg);
    }
    dg = 0.0f;
}
} // namespace t4k
