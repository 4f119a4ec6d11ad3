// act.cuh — activation functions with their saved derivative / mask, shared by k_activate and the fused linear epilogues
#pragma once
#include "common.cuh"
namespace t4k {
// k_activate (src/nn/nmath.cu:37-70): o = act(i), f = saved derivative / mask
#define SELU_L  1.0507
#define SELU_LA 1.7581
template<int L> __device__ __forceinline__ void act(float i, float alpha, float &o, float &f) {
    if (L == T4K_L_RELU)         { if (i > 0.0f) { f = 1.0f; o = i; } else { f = 0.0f; o = 0.0f; } }
    else if (L == T4K_L_TANH)    { o = tanhf(i); f = 1.0f - o * o; }
    else if (L == T4K_L_SIGMOID) { o = 1.0f / (1.0f + expf(-i)); f = o * (1.0f - o); }
    else if (L == T4K_L_SELU)    { if (i > 0.0f) { f = (float)SELU_L; o = i; }                // sic: no lambda on x (nmath.cu:56-58)
                                   else { f = (float)(SELU_LA * (double)__expf(i)); o = (float)((double)f - SELU_LA); } }
    else if (L == T4K_L_LEAKYRL) { if (i > 0.0f) { f = 1.0f; o = i; } else { f = alpha; o = alpha * i; } }
    else if (L == T4K_L_ELU)     { if (i > 0.0f) { f = 1.0f; o = i; } else { f = alpha * __expf(i); o = f - alpha; } }
    else /* DROPOUT */           { if (f > alpha) { f = 1.0f; o = i; } else { f = 0.0f; o = 0.0f; } }  // f holds U(0,1] on entry
}
} // namespace t4k
