// reduce.cu — reductions, losses, row softmax, column reduce, one-hot / hit
//   replaces k_sum / k_nvar / k_max / k_bce / k_dot / k_nan_inf (src/t4math.cu:23-131,248-365),
//   Tensor::loss (src/mu/tensor.cu:289-325), k_softmax* (src/nn/nmath.cu:74-169),
//   Model::_flogsoftmax (src/nn/forward.cu:246-259), k_dlinear_db (src/nn/nmath.cu:274-280),
//   Model::onehot / Model::hit host loops (src/nn/loss.cpp:47-107).
// All HBM-bound single-pass streams: 128-bit loads, warp-shuffle + smem block reduce, then a
// deterministic "last block finishes" pass over per-block partials (no float atomics, so
// results are run-to-run reproducible, unlike the reference's atomicAdd(sum, v)).
#include "common.cuh"

namespace t4k {

enum { R_SUM = 0, R_NVAR, R_MAX, R_MIN, R_BCE, R_MSE, R_CE, R_NLL, R_DOT };
enum { F_NONE = 0, F_DIV_N, F_SQRT_DIV_N, F_NEG_DIV_N, F_DIV_NEG_N_BCE, F_AXPBY };

#define RMAX_BLOCKS 1024
#define RCTRL       2048          // control word offset inside a reduce slot (see runtime.cu)

template<int KIND> __device__ __forceinline__ float r_init() {
    return KIND == R_MAX ? -FLT_MAX : (KIND == R_MIN ? FLT_MAX : 0.0f);
}
template<int KIND> __device__ __forceinline__ float r_elem(float a, float b, float p0) {
    if (KIND == R_SUM)  return a;
    if (KIND == R_NVAR) { float d = a - p0; return d * d; }
    if (KIND == R_BCE)  return b * __logf(a + DU_EPS) + (1.0f - b) * __logf(1.0f - a + DU_EPS);   // a=out, b=tgt
    if (KIND == R_MSE)  { float d = __fsub_rn(a, b); return __fmul_rn(d, d); }
    if (KIND == R_CE)   return __fmul_rn(__logf(fmaxf(a, DU_LNX)), b);
    if (KIND == R_NLL)  return __fmul_rn(a, b);
    if (KIND == R_DOT)  return a * b;
    return a;
}
template<int KIND> __device__ __forceinline__ float r_comb(float x, float y) {
    if (KIND == R_MAX) return fmaxf(x, y);
    if (KIND == R_MIN) return fminf(x, y);
    return x + y;
}
template<int KIND> __device__ __forceinline__ float r_block(float v, float *sm32) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = r_comb<KIND>(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (lane == 0) sm32[w] = v;
    __syncthreads();
    if (w == 0) {
        v = (lane < (int)(blockDim.x >> 5)) ? sm32[lane] : r_init<KIND>();
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = r_comb<KIND>(v, __shfl_xor_sync(0xffffffffu, v, o));
    }
    __syncthreads();
    return v;
}

// one launch: per-block partial → slot[blockIdx.x]; the last block to arrive combines the
// partials in index order and writes the finished scalar.
template<int KIND, bool TWO, bool VEC>
__global__ void __launch_bounds__(T4K_THREADS)
k_reduce(const float *__restrict__ A, const float *__restrict__ B, float p0, const float *p0_dev,
         int64_t n, float *slot, float *out, int fin, float fa, float fb) {
    __shared__ float sm32[32];
    __shared__ bool  last;
    pdl_wait(); pdl_trigger();                  // PDL: nothing global before this line
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    if (p0_dev) p0 = *p0_dev;
    float v = r_init<KIND>();
    if (VEC) {
        const int64_t n4 = n >> 2;
        for (int64_t i = tid; i < n4; i += nth) {
            float4 a = ldg4(A + 4 * i), b = TWO ? ldg4(B + 4 * i) : make_float4(0, 0, 0, 0);
            v = r_comb<KIND>(v, r_elem<KIND>(a.x, b.x, p0)); v = r_comb<KIND>(v, r_elem<KIND>(a.y, b.y, p0));
            v = r_comb<KIND>(v, r_elem<KIND>(a.z, b.z, p0)); v = r_comb<KIND>(v, r_elem<KIND>(a.w, b.w, p0));
        }
        for (int64_t j = (n4 << 2) + tid; j < n; j += nth) v = r_comb<KIND>(v, r_elem<KIND>(A[j], TWO ? B[j] : 0.0f, p0));
    } else {
        for (int64_t j = tid; j < n; j += nth) v = r_comb<KIND>(v, r_elem<KIND>(A[j], TWO ? B[j] : 0.0f, p0));
    }
    v = r_block<KIND>(v, sm32);
    if (threadIdx.x == 0) {
        slot[blockIdx.x] = v;
        __threadfence();
        unsigned t = atomicAdd(reinterpret_cast<unsigned*>(slot + RCTRL), 1u);
        last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    v = r_init<KIND>();
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) v = r_comb<KIND>(v, __ldcg(slot + i));
    v = r_block<KIND>(v, sm32);
    if (threadIdx.x == 0) {
        float r = v;
        switch (fin) {
        case F_DIV_N:         r = v / fa; break;                          // avg = sum / numel
        case F_SQRT_DIV_N:    r = fa > 0.0f ? __fsqrt_rn(v) / fa : 0.0f; break;   // std (tensor.cu:248)
        case F_NEG_DIV_N:     r = -v / fa; break;                         // loss: z = -Σ; z /= N
        case F_AXPBY:         r = v * fa + out[0] * fb; break;            // k_dot: acc*alpha + O*beta
        default: break;
        }
        out[0] = r;
        *reinterpret_cast<unsigned*>(slot + RCTRL) = 0u;                   // re-arm the slot
    }
}

template<int KIND, bool TWO>
static int launch_reduce(const float *A, const float *B, float p0, const float *p0_dev, int64_t n,
                         float *out, int fin, float fa, float fb, cudaStream_t st) {
    float *slot = reduce_slot(st);
    if (!slot) return T4K_ENOMEM;
    const bool vec = aligned16(A) && (!TWO || aligned16(B));
    int g = stream_grid(n, vec ? 8 : 2);
    if (g > RMAX_BLOCKS) g = RMAX_BLOCKS;
    if (vec) launch_pdl(k_reduce<KIND, TWO, true >, dim3(g), dim3(T4K_THREADS), 0, st, A, B, p0, p0_dev, n, slot, out, fin, fa, fb);
    else     launch_pdl(k_reduce<KIND, TWO, false>, dim3(g), dim3(T4K_THREADS), 0, st, A, B, p0, p0_dev, n, slot, out, fin, fa, fb);
    return check_launch();
}

// ------------------------------------------------------------------ k_dot, general (C > 1 or batch): warp per (n,c)
__global__ void __launch_bounds__(T4K_THREADS)
k_dot_nc(const float *__restrict__ A, const float *__restrict__ B, float *O, float alpha, float beta,
         int K, int C, int N, int64_t sA, int64_t sB) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= N * C) return;
    const int n = w / C, c = w % C;
    const float *a = A + n * sA + c, *b = B + n * sB + c;
    float acc = 0.0f;
    for (int k = lane; k < K; k += 32) acc = fmaf(a[(int64_t)k * C], b[(int64_t)k * C], acc);
    acc = warp_sum(acc);
    if (lane == 0) O[w] = acc * alpha + O[w] * beta;
}

// ------------------------------------------------------------------ NaN / Inf counter
__global__ void __launch_bounds__(T4K_THREADS) k_nan_inf(const float *__restrict__ A, int64_t n, int *cnt) {
    int v = 0;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        float x = A[j];
        v += (isnan(x) || isinf(x)) ? 1 : 0;
    }
    v = __reduce_add_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(cnt, v);
}

// ------------------------------------------------------------------ dB[e] += Σ_n dY[n,e]   (column reduce, coalesced in e)
__global__ void __launch_bounds__(1024) k_dbias(const float *__restrict__ dY, float *dB, int N, int E0) {
    __shared__ float sm[32][33];
    const int e = blockIdx.x * 32 + threadIdx.x;
    float v = 0.0f;
    if (e < E0) for (int n = threadIdx.y; n < N; n += 32) v += dY[(int64_t)n * E0 + e];
    sm[threadIdx.y][threadIdx.x] = v;
    __syncthreads();
    if (threadIdx.y == 0 && e < E0) {
        float s = 0.0f;
        #pragma unroll
        for (int r = 0; r < 32; r++) s += sm[r][threadIdx.x];
        dB[e] += s;
    }
}

// ------------------------------------------------------------------ row softmax / logsoftmax-as-coded: one warp per row
template<bool LOGSM>
__global__ void __launch_bounds__(T4K_THREADS) k_softmax(const float *__restrict__ I, float *O, int N, int C) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= N) return;
    const float *s = I + (int64_t)row * C; float *d = O + (int64_t)row * C;
    if (!LOGSM) {
        float mx = -FLT_MAX;
        for (int c = lane; c < C; c += 32) mx = fmaxf(mx, s[c]);
        mx = warp_max(mx);
        float sm = 0.0f;
        for (int c = lane; c < C; c += 32) { float e = __expf(s[c] - mx); d[c] = e; sm += e; }
        sm = warp_sum(sm);
        for (int c = lane; c < C; c += 32) d[c] = d[c] / sm;
    } else {
        float sm = 0.0f;
        for (int c = lane; c < C; c += 32) { float e = __expf(s[c]); d[c] = e; sm += e; }
        sm = warp_sum(sm);
        const float ls = __log10f(fmaxf(sm, DU_EPS));
        for (int c = lane; c < C; c += 32) d[c] = d[c] - ls;
    }
}

// ------------------------------------------------------------------ one-hot / hit
__global__ void __launch_bounds__(T4K_THREADS) k_onehot(const int32_t *__restrict__ label, float *hot, int E, int64_t total) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t m = (uint32_t)label[k / E];
        const uint32_t h = m < (uint32_t)E ? m : 0u;                 // loss.cpp:66
        hot[k] = ((uint32_t)(k % E) == h) ? 1.0f : 0.0f;
    }
}
__global__ void __launch_bounds__(T4K_THREADS) k_hit(const float *__restrict__ out, const float *__restrict__ hot, int N, int E, int *cnt) {
    __shared__ int sm[T4K_THREADS / 32];
    int v = 0;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        const float *o = out + (int64_t)n * E;
        float m = o[0]; int i = 0;
        for (int e = 1; e < E; e++) { float x = o[e]; if (x > m) { m = x; i = e; } }   // first max wins (loss.cpp:88-95)
        v += (int)hot[(int64_t)n * E + i];
    }
    v = __reduce_add_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) { int s = 0; for (int w = 0; w < T4K_THREADS / 32; w++) s += sm[w]; *cnt = s; }
}

// ------------------------------------------------------------------ Dataset::_load (src/mu/dataset.cu:124-152) on device
// d[i] = ((float)(int)u8[i] - mean) * scale  — two roundings, as the host loop has them (no FMA contraction);
// 16 pixels per thread: one 128-bit load of bytes, four 128-bit stores.  Labels widen U8 -> int32 in the same launch.
__global__ void __launch_bounds__(T4K_THREADS) k_dataset_load(const uint8_t *__restrict__ src, float *dst, int64_t n, float mean, float scale,
                                                              const uint8_t *__restrict__ lab8, int32_t *lab32, int nlab, float *hot, int E, int vec) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = tid; i < nlab; i += nth) lab32[i] = (int32_t)lab8[i];
    if (hot) {                                              // Model::onehot(Dataset&) (loss.cpp:59-68): zeros, then hot[n, m < E ? m : 0] = 1
        for (int64_t i = tid; i < (int64_t)nlab * E; i += nth) {
            const int m = (int)lab8[i / E], e = (int)(i % E);
            hot[i] = (e == (m < E ? m : 0)) ? 1.0f : 0.0f;
        }
    }
    if (vec) {
        const int64_t n16 = n >> 4;
        for (int64_t q = tid; q < n16; q += nth) {
            const uint4 w = __ldg(reinterpret_cast<const uint4*>(src) + q);
            const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
            #pragma unroll
            for (int k = 0; k < 4; k++) {
                float4 o;
                o.x = __fmul_rn(__fsub_rn((float)(int)(ww[k] & 0xffu), mean), scale);
                o.y = __fmul_rn(__fsub_rn((float)(int)((ww[k] >> 8) & 0xffu), mean), scale);
                o.z = __fmul_rn(__fsub_rn((float)(int)((ww[k] >> 16) & 0xffu), mean), scale);
                o.w = __fmul_rn(__fsub_rn((float)(int)(ww[k] >> 24), mean), scale);
                stg4(dst + 16 * q + 4 * k, o);
            }
        }
        for (int64_t i = (n16 << 4) + tid; i < n; i += nth) dst[i] = __fmul_rn(__fsub_rn((float)(int)src[i], mean), scale);
    } else {
        for (int64_t i = tid; i < n; i += nth) dst[i] = __fmul_rn(__fsub_rn((float)(int)src[i], mean), scale);
    }
}

} // namespace t4k
using namespace t4k;

// ====================================================================== C ABI
extern "C" int t4k_sum(const float *A, int64_t n, float *out, t4k_stream_t s) {
    if (!A || !out || n < 0) return T4K_EINVAL;
    return launch_reduce<R_SUM, false>(A, nullptr, 0.0f, nullptr, n, out, F_NONE, 0, 0, STRM(s));
}
extern "C" int t4k_nvar(const float *A, float avg, int64_t n, float *out, t4k_stream_t s) {
    if (!A || !out || n < 0) return T4K_EINVAL;
    return launch_reduce<R_NVAR, false>(A, nullptr, avg, nullptr, n, out, F_NONE, 0, 0, STRM(s));
}
extern "C" int t4k_minmax(const float *A, int64_t n, int find_max, float *out, t4k_stream_t s) {
    if (!A || !out || n < 0) return T4K_EINVAL;
    return find_max ? launch_reduce<R_MAX, false>(A, nullptr, 0.0f, nullptr, n, out, F_NONE, 0, 0, STRM(s))
                    : launch_reduce<R_MIN, false>(A, nullptr, 0.0f, nullptr, n, out, F_NONE, 0, 0, STRM(s));
}
extern "C" int t4k_avg_std(const float *A, int64_t n, float *out2, t4k_stream_t s) {
    if (!A || !out2 || n < 1) return T4K_EINVAL;
    int rc = launch_reduce<R_SUM, false>(A, nullptr, 0.0f, nullptr, n, out2, F_DIV_N, (float)n, 0, STRM(s));
    if (rc) return rc;
    return launch_reduce<R_NVAR, false>(A, nullptr, 0.0f, out2, n, out2 + 1, F_SQRT_DIV_N, (float)n, 0, STRM(s));
}
extern "C" int t4k_dot(const float *A, const float *B, float *O, float alpha, float beta,
                       int K, int C, int Na, int Nb, t4k_stream_t s) {
    if (!A || !B || !O || K < 0 || C < 1 || Na < 1 || Nb < 1) return T4K_EINVAL;
    if (Na != Nb && Na != 1 && Nb != 1) return T4K_EINVAL;
    const int N = Na > Nb ? Na : Nb;
    if (C == 1 && N == 1)
        return launch_reduce<R_DOT, true>(A, B, 0.0f, nullptr, K, O, F_AXPBY, alpha, beta, STRM(s));
    const int64_t sA = (Na == 1 && N > 1) ? 0 : (int64_t)K * C, sB = (Nb == 1 && N > 1) ? 0 : (int64_t)K * C;
    const int warps = N * C;
    k_dot_nc<<<(warps * 32 + T4K_THREADS - 1) / T4K_THREADS, T4K_THREADS, 0, STRM(s)>>>(A, B, O, alpha, beta, K, C, N, sA, sB);
    return check_launch();
}
extern "C" int t4k_loss(int kind, const float *out, const float *tgt, int64_t numel, int N, float *loss, t4k_stream_t s) {
    if (!out || !tgt || !loss || numel < 0 || N < 1) return T4K_EINVAL;
    const float fN = (float)N;
    switch (kind) {
    case T4K_LOSS_MSE: return launch_reduce<R_MSE, true>(out, tgt, 0, nullptr, numel, loss, F_DIV_N,     fN, 0, STRM(s));
    case T4K_LOSS_BCE: return launch_reduce<R_BCE, true>(out, tgt, 0, nullptr, numel, loss, F_NEG_DIV_N, fN, 0, STRM(s));
    case T4K_LOSS_CE:  return launch_reduce<R_CE,  true>(out, tgt, 0, nullptr, numel, loss, F_NEG_DIV_N, fN, 0, STRM(s));
    case T4K_LOSS_NLL: return launch_reduce<R_NLL, true>(out, tgt, 0, nullptr, numel, loss, F_NEG_DIV_N, fN, 0, STRM(s));
    default: return T4K_EINVAL;          // "Model#loss op=%d not supported!" tensor.cu:320
    }
}
extern "C" int t4k_nan_inf(const float *A, int64_t n, int *cnt, t4k_stream_t s) {
    if (!A || !cnt || n < 0) return T4K_EINVAL;
    cudaError_t e = cudaMemsetAsync(cnt, 0, sizeof(int), STRM(s));
    if (e != cudaSuccess) return (int)e;
    if (n == 0) return 0;
    k_nan_inf<<<stream_grid(n, 4), T4K_THREADS, 0, STRM(s)>>>(A, n, cnt);
    return check_launch();
}
extern "C" int t4k_dbias(const float *dY, float *dB, int N, int E0, t4k_stream_t s) {
    if (!dY || !dB || N < 1 || E0 < 1) return T4K_EINVAL;
    dim3 b(32, 32), g((E0 + 31) / 32);
    k_dbias<<<g, b, 0, STRM(s)>>>(dY, dB, N, E0);
    return check_launch();
}
extern "C" int t4k_softmax_fwd(const float *I, float *O, int N, int C, t4k_stream_t s) {
    if (!I || !O || N < 1 || C < 1) return T4K_EINVAL;
    k_softmax<false><<<((int64_t)N * 32 + T4K_THREADS - 1) / T4K_THREADS, T4K_THREADS, 0, STRM(s)>>>(I, O, N, C);
    return check_launch();
}
extern "C" int t4k_logsoftmax_fwd(const float *I, float *O, int N, int C, t4k_stream_t s) {
    if (!I || !O || N < 1 || C < 1) return T4K_EINVAL;
    k_softmax<true><<<((int64_t)N * 32 + T4K_THREADS - 1) / T4K_THREADS, T4K_THREADS, 0, STRM(s)>>>(I, O, N, C);
    return check_launch();
}
extern "C" int t4k_dataset_load(const uint8_t *src, float *dst, int64_t n, float mean, float scale,
                                const uint8_t *lab8, int32_t *lab32, int nlab, float *hot, int E, t4k_stream_t s) {
    if (!src || !dst || n < 0 || nlab < 0 || (nlab && (!lab8 || !lab32)) || (hot && E < 1)) return T4K_EINVAL;
    if (n == 0 && nlab == 0) return 0;
    const int vec = aligned16(src) && aligned16(dst);
    const int64_t work = n / (vec ? 16 : 1) > (int64_t)nlab * (hot ? E : 1) ? n / (vec ? 16 : 1) : (int64_t)nlab * (hot ? E : 1);
    k_dataset_load<<<stream_grid(work > 0 ? work : 1), T4K_THREADS, 0, STRM(s)>>>(src, dst, n, mean, scale, lab8, lab32, nlab, hot, E, vec);
    return check_launch();
}
extern "C" int t4k_onehot(const int32_t *label, float *hot, int N, int E, t4k_stream_t s) {
    if (!label || !hot || N < 1 || E < 1) return T4K_EINVAL;
    int64_t total = (int64_t)N * E;
    k_onehot<<<stream_grid(total), T4K_THREADS, 0, STRM(s)>>>(label, hot, E, total);
    return check_launch();
}
extern "C" int t4k_hit(const float *out, const float *hot, int N, int E, int *cnt, t4k_stream_t s) {
    if (!out || !hot || !cnt || N < 1 || E < 1) return T4K_EINVAL;
    k_hit<<<1, T4K_THREADS, 0, STRM(s)>>>(out, hot, N, E, cnt);
    return check_launch();
}
