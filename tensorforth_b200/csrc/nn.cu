// nn.cu — linear layer forward/backward on top of the GEMM engines
//   Model::_flinear (src/nn/forward.cu:158-198)  : Y = X @ W^T + B      (Tensor::linear tB=true, then k_bias)
//   Model::_blinear (src/nn/backprop.cu:194-254) : dB += ΣdY ; dW += dY^T @ X (beta=1) ; dX = dY @ W
#include "common.cuh"
using namespace t4k;

extern "C" int t4k_linear_fwd(const float *X, const float *W, const float *B, float *Y, int N, int E0, int E1, t4k_stream_t s) {
    if (!X || !W || !B || !Y || N < 1 || E0 < 1 || E1 < 1) return T4K_EINVAL;
    int rc = t4k_gemm(X, W, Y, 1.0f, 0.0f, 0, 1, N, E0, E1, 1, 1, 0, 0, 0, s);
    if (rc) return rc;
    return t4k_bias(B, Y, N, E0, s);
}
extern "C" int t4k_linear_bwd(const float *X, const float *W, const float *dY, float *dX, float *dW, float *dB,
                              int N, int E0, int E1, int train, t4k_stream_t s) {
    if (!X || !W || !dY || !dX || N < 1 || E0 < 1 || E1 < 1) return T4K_EINVAL;
    if (train) {
        if (!dW || !dB) return T4K_EINVAL;
        int rc = t4k_dbias(dY, dB, N, E0, s);                                        // dB[E0] += Σ_n dY
        if (rc) return rc;
        rc = t4k_gemm(dY, X, dW, 1.0f, 1.0f, 1, 0, E0, E1, N, 1, 1, 0, 0, 0, s);      // dW[E0,E1] += dY^T[E0,N] @ X[N,E1]
        if (rc) return rc;
    }
    return t4k_gemm(dY, W, dX, 1.0f, 0.0f, 0, 0, N, E1, E0, 1, 1, 0, 0, 0, s);       // dX[N,E1] = dY[N,E0] @ W[E0,E1]
}
