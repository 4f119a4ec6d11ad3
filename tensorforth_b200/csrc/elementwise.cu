// elementwise.cu — HBM-streaming elementwise kernels (128-bit vectorised, grid sized to the SM count)
//   replaces k_math / k_ts_op / k_tt_op / k_copy / k_transpose / k_identity (src/t4math.cu:134-234),
//   k_bias / k_activate (src/nn/nmath.cu:27-70) and k_sgd / k_adam / k_adamw (src/nn/nmath.cu:419-472).
// Roofline: all HBM-bound; algorithmic bytes per element are listed in DESIGN.md §kernels.
#include "common.cuh"
#include "act.cuh"
#include "optim.cuh"

namespace t4k {

// ------------------------------------------------------------------ element functors
// math identical to the reference's intrinsics (src/t4math.h:60-104, src/t4math.cu:179-199)
template<int OP> __device__ __forceinline__ float map_op(float a, float v, int64_t j, int64_t n) {
    if (OP == T4K_ABS)   return fabsf(a);
    if (OP == T4K_NEG)   return -a;
    if (OP == T4K_EXP)   return __expf(a);
    if (OP == T4K_LN)    return __logf(fmaxf(a, DU_LNX));
    if (OP == T4K_LOG)   return __log10f(fmaxf(a, DU_LNX));
    if (OP == T4K_TANH)  return tanhf(a);
    if (OP == T4K_RELU)  return fmaxf(0.0f, a);
    if (OP == T4K_SIGM)  return 1.0f / (1.0f + expf(-a));
    if (OP == T4K_SQRT)  return __fsqrt_rn(fmaxf(a, 0.0f));
    if (OP == T4K_RCP)   return __frcp_rn(a);
    if (OP == T4K_SAT)   return __saturatef(a);
    if (OP == T4K_FILL)  return v;
    if (OP == T4K_GFILL) return v * (float)j / (float)n;
    if (OP == T4K_SCALE) return a * v;
    if (OP == T4K_POW)   return __powf(a, v);
    if (OP == T4K_ADD)   return a + v;
    if (OP == T4K_SUB)   return a - v;
    if (OP == T4K_MUL)   return a * v;
    if (OP == T4K_DIV)   return a / v;
    return a;
}
template<int OP> __device__ __forceinline__ float bin_op(float a, float b) {
    if (OP == T4K_ADD) return __fadd_rn(a, b);
    if (OP == T4K_SUB) return __fsub_rn(a, b);
    if (OP == T4K_MUL) return __fmul_rn(a, b);
    return __fdiv_rn(a, b);
}

// ------------------------------------------------------------------ k_map: in-place A = op(A, v)
template<int OP, bool VEC>
__global__ void __launch_bounds__(T4K_THREADS) k_map(float *A, float v, int64_t n) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    constexpr bool READS = !(OP == T4K_FILL || OP == T4K_GFILL);
    if (VEC) {
        const int64_t n4 = n >> 2;
        for (int64_t i = tid; i < n4; i += nth) {
            float4 a = READS ? *reinterpret_cast<float4*>(A + 4 * i) : make_float4(0, 0, 0, 0);
            a.x = map_op<OP>(a.x, v, 4 * i + 0, n); a.y = map_op<OP>(a.y, v, 4 * i + 1, n);
            a.z = map_op<OP>(a.z, v, 4 * i + 2, n); a.w = map_op<OP>(a.w, v, 4 * i + 3, n);
            stg4(A + 4 * i, a);
        }
        for (int64_t j = (n4 << 2) + tid; j < n; j += nth) A[j] = map_op<OP>(READS ? A[j] : 0.0f, v, j, n);
    } else {
        for (int64_t j = tid; j < n; j += nth) A[j] = map_op<OP>(READS ? A[j] : 0.0f, v, j, n);
    }
}
template<int OP> static int launch_map(float *A, float v, int64_t n, cudaStream_t st) {
    if (aligned16(A)) k_map<OP, true ><<<stream_grid(n, 4), T4K_THREADS, 0, st>>>(A, v, n);
    else              k_map<OP, false><<<stream_grid(n, 1), T4K_THREADS, 0, st>>>(A, v, n);
    return check_launch();
}

// ------------------------------------------------------------------ k_ts: O = A op v
template<int OP, bool VEC>
__global__ void __launch_bounds__(T4K_THREADS) k_ts(const float *A, float v, float *O, int64_t n) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    if (VEC) {
        const int64_t n4 = n >> 2;
        for (int64_t i = tid; i < n4; i += nth) {
            float4 a = *reinterpret_cast<const float4*>(A + 4 * i);   // A may alias O: plain load
            a.x = bin_op<OP>(a.x, v); a.y = bin_op<OP>(a.y, v); a.z = bin_op<OP>(a.z, v); a.w = bin_op<OP>(a.w, v);
            stg4(O + 4 * i, a);
        }
        for (int64_t j = (n4 << 2) + tid; j < n; j += nth) O[j] = bin_op<OP>(A[j], v);
    } else {
        for (int64_t j = tid; j < n; j += nth) O[j] = bin_op<OP>(A[j], v);
    }
}
// ------------------------------------------------------------------ k_tt: O[n] = A[n|0] op B[n|0]
// one launch covers all N slices (the reference launches per slice, src/mu/tensor.cu:39-46)
template<int OP, bool VEC>
__global__ void __launch_bounds__(T4K_THREADS) k_tt(const float *A, const float *B, float *O,
                                                    int64_t hwc, int64_t total, int bA, int bB) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    if (VEC) {      // hwc % 4 == 0 guaranteed by the launcher when broadcasting
        const int64_t n4 = total >> 2;
        for (int64_t i = tid; i < n4; i += nth) {
            const int64_t j = 4 * i;
            const int64_t ja = bA ? j % hwc : j, jb = bB ? j % hwc : j;
            float4 a = *reinterpret_cast<const float4*>(A + ja);
            float4 b = *reinterpret_cast<const float4*>(B + jb);
            a.x = bin_op<OP>(a.x, b.x); a.y = bin_op<OP>(a.y, b.y); a.z = bin_op<OP>(a.z, b.z); a.w = bin_op<OP>(a.w, b.w);
            stg4(O + j, a);
        }
        for (int64_t j = (n4 << 2) + tid; j < total; j += nth)
            O[j] = bin_op<OP>(A[bA ? j % hwc : j], B[bB ? j % hwc : j]);
    } else {
        for (int64_t j = tid; j < total; j += nth)
            O[j] = bin_op<OP>(A[bA ? j % hwc : j], B[bB ? j % hwc : j]);
    }
}

// ------------------------------------------------------------------ copy
template<bool VEC>
__global__ void __launch_bounds__(T4K_THREADS) k_copy(const float *__restrict__ s, float *__restrict__ d, int64_t n) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    if (VEC) {
        const int64_t n4 = n >> 2;
        int64_t i = tid;
        for (; i + 3 * nth < n4; i += 4 * nth) {            // 4 x 128-bit loads in flight per thread
            float4 a = ldg4(s + 4 * i), b = ldg4(s + 4 * (i + nth)), c = ldg4(s + 4 * (i + 2 * nth)), e = ldg4(s + 4 * (i + 3 * nth));
            stg4(d + 4 * i, a); stg4(d + 4 * (i + nth), b); stg4(d + 4 * (i + 2 * nth), c); stg4(d + 4 * (i + 3 * nth), e);
        }
        for (; i < n4; i += nth) stg4(d + 4 * i, ldg4(s + 4 * i));
        for (int64_t j = (n4 << 2) + tid; j < n; j += nth) d[j] = s[j];
    } else {
        for (int64_t j = tid; j < n; j += nth) d[j] = s[j];
    }
}

// ------------------------------------------------------------------ transpose (index word → bit exact)
// T[n, j, i, c] = A[n, i, j, c]; 32x32 smem tile per channel for C==1, direct for C>1
__global__ void __launch_bounds__(256) k_transpose_c1(const float *__restrict__ A, float *__restrict__ T, int H, int W) {
    __shared__ float tile[32][33];
    const int64_t base = (int64_t)blockIdx.z * H * W;
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += 8) {
        int i = by + r, j = bx + threadIdx.x;
        if (i < H && j < W) tile[r][threadIdx.x] = A[base + (int64_t)i * W + j];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += 8) {
        int j = bx + r, i = by + threadIdx.x;
        if (i < H && j < W) T[base + (int64_t)j * H + i] = tile[threadIdx.x][r];
    }
}
__global__ void __launch_bounds__(T4K_THREADS) k_transpose_c(const float *__restrict__ A, float *__restrict__ T,
                                                             int H, int W, int C, int64_t total) {
    const int64_t hwc = (int64_t)H * W * C;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x) {
        int64_t n = k / hwc, r = k % hwc;            // k indexes the OUTPUT (coalesced writes)
        int c = (int)(r % C); int64_t p = r / C;
        int i = (int)(p % H), j = (int)(p / H);      // T is [W,H,C]
        T[k] = A[n * hwc + ((int64_t)i * W + j) * C + c];
    }
}
__global__ void __launch_bounds__(T4K_THREADS) k_identity(float *T, int H, int W, int C, int64_t total) {
    const int64_t hwc = (int64_t)H * W * C;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x) {
        int64_t p = (k % hwc) / C;
        T[k] = ((int)(p / W) == (int)(p % W)) ? 1.0f : 0.0f;
    }
}

// ------------------------------------------------------------------ bias: Y[n,e] += B[e]
__global__ void __launch_bounds__(T4K_THREADS) k_bias(const float *__restrict__ B, float *Y, int E0, int64_t total) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x)
        Y[k] += __ldg(B + (k % E0));
}

template<int L, bool VEC>
__global__ void __launch_bounds__(T4K_THREADS) k_activate(const float *I, float *O, float *F, float alpha, int64_t n) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    if (VEC) {
        const int64_t n4 = n >> 2;
        for (int64_t k = tid; k < n4; k += nth) {
            float4 i = *reinterpret_cast<const float4*>(I + 4 * k), o, f;
            if (L == T4K_L_DROPOUT) f = *reinterpret_cast<const float4*>(F + 4 * k);
            act<L>(i.x, alpha, o.x, f.x); act<L>(i.y, alpha, o.y, f.y); act<L>(i.z, alpha, o.z, f.z); act<L>(i.w, alpha, o.w, f.w);
            stg4(O + 4 * k, o); stg4(F + 4 * k, f);
        }
        for (int64_t j = (n4 << 2) + tid; j < n; j += nth) { float o, f = (L == T4K_L_DROPOUT) ? F[j] : 0.0f; act<L>(I[j], alpha, o, f); O[j] = o; F[j] = f; }
    } else {
        for (int64_t j = tid; j < n; j += nth) { float o, f = (L == T4K_L_DROPOUT) ? F[j] : 0.0f; act<L>(I[j], alpha, o, f); O[j] = o; F[j] = f; }
    }
}

// ------------------------------------------------------------------ optimizers (one pass: read g,dg,m,v / write g,dg=0,m,v): optim.cuh
// single tensor; SGD uses true division by Nw to match `DG[j] / N` bit for bit
template<int KIND>
__global__ void __launch_bounds__(T4K_THREADS) k_optim(float *G, float *DG, float *M, float *V, int Nw, bool mom, OptP p, int64_t n) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = tid; j < n; j += nth) {
        float g = G[j], dg = DG[j], m = 0.0f, v = 0.0f;
        if (KIND == 0) { dg = dg / (float)Nw; if (mom) m = M[j]; }
        else { m = M[j]; v = V[j]; }
        opt_step<KIND>(g, dg, m, v, 1.0f, mom, p);
        G[j] = g; DG[j] = 0.0f;
        if (KIND == 0) { if (mom) M[j] = m; } else { M[j] = m; V[j] = v; }
    }
}
// whole model in one launch over flat arenas; segment table gives Nw per parameter tensor
template<int KIND>
__global__ void __launch_bounds__(T4K_THREADS) k_optim_multi(float *G, float *DG, float *M, float *V,
                                                             const t4k_seg_t *__restrict__ seg, int nseg,
                                                             int64_t from, int64_t total, bool mom, OptP p) {
    pdl_wait(); pdl_trigger();                  // PDL: nothing global before this line
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = from + tid; j < total; j += nth) {
        float g = G[j], dg = DG[j], m = 0.0f, v = 0.0f;
        if (KIND == 0) {
            int lo = 0, hi = nseg - 1;                     // binary search the owning segment (nseg is tiny)
            while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (seg[mid].off <= j) lo = mid; else hi = mid - 1; }
            dg = dg / (float)seg[lo].Nw;
            if (mom) m = M[j];
        } else { m = M[j]; v = V[j]; }
        opt_step<KIND>(g, dg, m, v, 1.0f, mom, p);
        G[j] = g; DG[j] = 0.0f;
        if (KIND == 0) { if (mom) M[j] = m; } else { M[j] = m; V[j] = v; }
    }
}

} // namespace t4k
using namespace t4k;

// ====================================================================== C ABI
#define MAP_CASE(OPC) case OPC: return launch_map<OPC>(A, v, n, st);
extern "C" int t4k_map(int op, float *A, float v, int64_t n, t4k_stream_t s) {
    if (!A || n < 0) return T4K_EINVAL;
    if (n == 0) return 0;
    cudaStream_t st = STRM(s);
    switch (op) {
        MAP_CASE(T4K_ABS) MAP_CASE(T4K_NEG) MAP_CASE(T4K_EXP) MAP_CASE(T4K_LN) MAP_CASE(T4K_LOG)
        MAP_CASE(T4K_TANH) MAP_CASE(T4K_RELU) MAP_CASE(T4K_SIGM) MAP_CASE(T4K_SQRT) MAP_CASE(T4K_RCP)
        MAP_CASE(T4K_SAT) MAP_CASE(T4K_FILL) MAP_CASE(T4K_GFILL) MAP_CASE(T4K_SCALE) MAP_CASE(T4K_POW)
        MAP_CASE(T4K_ADD) MAP_CASE(T4K_SUB) MAP_CASE(T4K_MUL) MAP_CASE(T4K_DIV)
        default: return T4K_EINVAL;       // reference prints "k_math op=%d not supported" (t4math.cu:199)
    }
}

template<int OP> static int launch_ts(const float *A, float v, float *O, int64_t n, cudaStream_t st) {
    if (aligned16(A) && aligned16(O)) k_ts<OP, true ><<<stream_grid(n, 4), T4K_THREADS, 0, st>>>(A, v, O, n);
    else                              k_ts<OP, false><<<stream_grid(n, 1), T4K_THREADS, 0, st>>>(A, v, O, n);
    return check_launch();
}
extern "C" int t4k_ts_op(int op, const float *A, float v, float *O, int64_t n, t4k_stream_t s) {
    if (!A || !O || n < 0) return T4K_EINVAL;
    if (n == 0) return 0;
    switch (op) {
    case T4K_ADD: return launch_ts<T4K_ADD>(A, v, O, n, STRM(s));
    case T4K_SUB: return launch_ts<T4K_SUB>(A, v, O, n, STRM(s));
    case T4K_MUL: return launch_ts<T4K_MUL>(A, v, O, n, STRM(s));
    case T4K_DIV: return launch_ts<T4K_DIV>(A, v, O, n, STRM(s));
    default: return T4K_EINVAL;
    }
}

template<int OP> static int launch_tt(const float *A, const float *B, float *O, int64_t hwc, int Na, int Nb, cudaStream_t st) {
    const int N = Na > Nb ? Na : Nb;
    const int64_t total = hwc * N;
    const int bA = (Na == 1 && N > 1), bB = (Nb == 1 && N > 1);
    const bool vec = aligned16(A) && aligned16(B) && aligned16(O) && (!(bA || bB) || (hwc & 3) == 0);
    if (vec) k_tt<OP, true ><<<stream_grid(total, 4), T4K_THREADS, 0, st>>>(A, B, O, hwc, total, bA, bB);
    else     k_tt<OP, false><<<stream_grid(total, 1), T4K_THREADS, 0, st>>>(A, B, O, hwc, total, bA, bB);
    return check_launch();
}
extern "C" int t4k_tt_op(int op, const float *A, const float *B, float *O, int64_t hwc, int Na, int Nb, t4k_stream_t s) {
    if (!A || !B || !O || hwc < 0 || Na < 1 || Nb < 1) return T4K_EINVAL;
    if (Na != Nb && Na != 1 && Nb != 1) return T4K_EINVAL;          // tensor.cu:35-38
    if (hwc == 0) return 0;
    switch (op) {
    case T4K_ADD: return launch_tt<T4K_ADD>(A, B, O, hwc, Na, Nb, STRM(s));
    case T4K_SUB: return launch_tt<T4K_SUB>(A, B, O, hwc, Na, Nb, STRM(s));
    case T4K_MUL: return launch_tt<T4K_MUL>(A, B, O, hwc, Na, Nb, STRM(s));
    case T4K_DIV: return launch_tt<T4K_DIV>(A, B, O, hwc, Na, Nb, STRM(s));
    default: return T4K_EINVAL;
    }
}
extern "C" int t4k_activate_bwd(const float *dY, const float *F, float *dX, int64_t n, t4k_stream_t s) {
    return t4k_tt_op(T4K_MUL, dY, F, dX, n, 1, 1, s);
}

extern "C" int t4k_copy(const float *src, float *dst, int64_t n, t4k_stream_t s) {
    if (!src || !dst || n < 0) return T4K_EINVAL;
    if (n == 0 || src == dst) return 0;
    if (aligned16(src) && aligned16(dst)) k_copy<true ><<<stream_grid(n, 16), T4K_THREADS, 0, STRM(s)>>>(src, dst, n);
    else                                  k_copy<false><<<stream_grid(n, 1),  T4K_THREADS, 0, STRM(s)>>>(src, dst, n);
    return check_launch();
}

extern "C" int t4k_transpose(const float *A, float *T, int N, int H, int W, int C, t4k_stream_t s) {
    if (!A || !T || N < 1 || H < 1 || W < 1 || C < 1 || A == T) return T4K_EINVAL;
    if (C == 1 && N <= 65535 && (H + 31) / 32 <= 65535) {
        dim3 g((W + 31) / 32, (H + 31) / 32, N), b(32, 8);
        k_transpose_c1<<<g, b, 0, STRM(s)>>>(A, T, H, W);
    } else {
        int64_t total = (int64_t)N * H * W * C;
        k_transpose_c<<<stream_grid(total), T4K_THREADS, 0, STRM(s)>>>(A, T, H, W, C, total);
    }
    return check_launch();
}
extern "C" int t4k_identity(float *T, int N, int H, int W, int C, t4k_stream_t s) {
    if (!T || N < 1 || H < 1 || W < 1 || C < 1) return T4K_EINVAL;
    int64_t total = (int64_t)N * H * W * C;
    k_identity<<<stream_grid(total), T4K_THREADS, 0, STRM(s)>>>(T, H, W, C, total);
    return check_launch();
}
extern "C" int t4k_bias(const float *B, float *Y, int N, int E0, t4k_stream_t s) {
    if (!B || !Y || N < 1 || E0 < 1) return T4K_EINVAL;
    int64_t total = (int64_t)N * E0;
    k_bias<<<stream_grid(total), T4K_THREADS, 0, STRM(s)>>>(B, Y, E0, total);
    return check_launch();
}

template<int L> static int launch_act(const float *I, float *O, float *F, float alpha, int64_t n, cudaStream_t st) {
    if (aligned16(I) && aligned16(O) && aligned16(F)) k_activate<L, true ><<<stream_grid(n, 4), T4K_THREADS, 0, st>>>(I, O, F, alpha, n);
    else                                              k_activate<L, false><<<stream_grid(n, 1), T4K_THREADS, 0, st>>>(I, O, F, alpha, n);
    return check_launch();
}
extern "C" int t4k_activate_fwd(int layer, const float *I, float *O, float *F, float alpha, int64_t n, t4k_stream_t s) {
    if (!I || !O || !F || n < 0) return T4K_EINVAL;
    if (n == 0) return 0;
    switch (layer) {
    case T4K_L_RELU:    return launch_act<T4K_L_RELU>(I, O, F, alpha, n, STRM(s));
    case T4K_L_TANH:    return launch_act<T4K_L_TANH>(I, O, F, alpha, n, STRM(s));
    case T4K_L_SIGMOID: return launch_act<T4K_L_SIGMOID>(I, O, F, alpha, n, STRM(s));
    case T4K_L_SELU:    return launch_act<T4K_L_SELU>(I, O, F, alpha, n, STRM(s));
    case T4K_L_LEAKYRL: return launch_act<T4K_L_LEAKYRL>(I, O, F, alpha, n, STRM(s));
    case T4K_L_ELU:     return launch_act<T4K_L_ELU>(I, O, F, alpha, n, STRM(s));
    case T4K_L_DROPOUT: return launch_act<T4K_L_DROPOUT>(I, O, F, alpha, n, STRM(s));
    default: return T4K_EINVAL;
    }
}

extern "C" int t4k_sgd(float *G, float *DG, float *M, int Nw, float lr, float b, int64_t n, t4k_stream_t s) {
    if (!G || !DG || n < 0 || Nw < 1) return T4K_EINVAL;
    if (n == 0) return 0;
    const bool mom = !(fabsf(b) < DU_EPS);                      // ZEQ(b), nmath.cu:429
    if (mom && !M) return T4K_EINVAL;
    OptP p{lr, b, 0.0f, 0.0f};
    k_optim<0><<<stream_grid(n), T4K_THREADS, 0, STRM(s)>>>(G, DG, M, nullptr, Nw, mom, p, n);
    return check_launch();
}
extern "C" int t4k_adam(float *G, float *DG, float *M, float *V, float lr, float b1, float b2, int64_t n, t4k_stream_t s) {
    if (!G || !DG || !M || !V || n < 0) return T4K_EINVAL;
    if (n == 0) return 0;
    OptP p{lr, b1, b2, 0.0f};
    k_optim<1><<<stream_grid(n), T4K_THREADS, 0, STRM(s)>>>(G, DG, M, V, 1, true, p, n);
    return check_launch();
}
extern "C" int t4k_adamw(float *G, float *DG, float *M, float *V, float lr, float b1, float b2, float wd, int64_t n, t4k_stream_t s) {
    if (!G || !DG || !M || !V || n < 0) return T4K_EINVAL;
    if (n == 0) return 0;
    OptP p{lr, b1, b2, wd};
    k_optim<2><<<stream_grid(n), T4K_THREADS, 0, STRM(s)>>>(G, DG, M, V, 1, true, p, n);
    return check_launch();
}
extern "C" int t4k_optim_multi(int kind, float *G, float *DG, float *M, float *V, const t4k_seg_t *seg, int nseg,
                               int64_t total, float lr, float b1, float b2, float wd, t4k_stream_t s) {
    return t4k_optim_multi_range(kind, G, DG, M, V, seg, nseg, 0, total, lr, b1, b2, wd, s);
}
extern "C" int t4k_optim_multi_range(int kind, float *G, float *DG, float *M, float *V, const t4k_seg_t *seg, int nseg,
                                     int64_t from, int64_t total, float lr, float b1, float b2, float wd, t4k_stream_t s) {
    if (!G || !DG || !seg || nseg < 1 || from < 0 || total < from) return T4K_EINVAL;
    if (total == from) return 0;
    OptP p{lr, b1, b2, wd};
    const int g = stream_grid(total - from);
    switch (kind) {
    case 0: { const bool mom = !(fabsf(b1) < DU_EPS); if (mom && !M) return T4K_EINVAL;
              launch_pdl(k_optim_multi<0>, dim3(g), dim3(T4K_THREADS), 0, STRM(s), G, DG, M, V, seg, nseg, from, total, mom, p); } break;
    case 1: if (!M || !V) return T4K_EINVAL; launch_pdl(k_optim_multi<1>, dim3(g), dim3(T4K_THREADS), 0, STRM(s), G, DG, M, V, seg, nseg, from, total, true, p); break;
    case 2: if (!M || !V) return T4K_EINVAL; launch_pdl(k_optim_multi<2>, dim3(g), dim3(T4K_THREADS), 0, STRM(s), G, DG, M, V, seg, nseg, from, total, true, p); break;
    default: return T4K_EINVAL;
    }
    return check_launch();
}
