// rand.cu — counter-based RNG (Philox4x32-10), full-grid
//   replaces k_rand_init / k_rand (src/util.cu:46-70: 1024 XORWOW states in ONE block, serial over
//   pages) behind System::rand (src/sys.cpp:77-95).  d[i] = scale * (bias + x),
//   x ~ U(0,1] (curand_uniform convention) or N(0,1) (Box-Muller).
// Element i depends only on (seed, offset + i): independent of grid size and of how a tensor is
// sharded across GPUs.  Bit parity with the reference is undefined (it seeds with time()).
#include "common.cuh"

namespace t4k {

static uint64_t g_seed = 0x243F6A8885A308D3ull, g_offset = 0;
// Replay epoch (device memory): a captured CUDA graph bakes seed and offset into its kernel nodes, so every replay would draw the same
// numbers (same dropout masks, same latent batches).  Each draw therefore adds epoch * 2^40 to its counter, and a graph that contains
// draws starts with t4k_rand_tick (epoch += 1).  Eager streams never tick: the host offset alone advances, as before.
static uint64_t *g_epoch_dev[16];
static uint64_t *epoch_ptr() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev >= 16) return nullptr;
    if (!g_epoch_dev[dev]) {
        if (cudaMalloc((void**)&g_epoch_dev[dev], 64) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        cudaMemset(g_epoch_dev[dev], 0, 64);
    }
    return g_epoch_dev[dev];
}
__global__ void k_rand_tick(uint64_t *epoch) { if (threadIdx.x == 0 && blockIdx.x == 0) epoch[0] += 1; }

__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
    #pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += 0x9E3779B9u; key.y += 0xBB67AE85u;
    }
    return ctr;
}
__device__ __forceinline__ float u01(uint32_t x) { return (float)x * 2.3283064365386963e-10f + 1.1641532182693481e-10f; }  // (0,1]

__global__ void __launch_bounds__(T4K_THREADS)
k_rand(float *d, int64_t n, int opt, float bias, float scale, uint64_t seed, uint64_t offset, const uint64_t *__restrict__ epoch) {
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    if (epoch) offset += epoch[0] << 40;
    const int64_t nq = (n + 3) >> 2;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += (int64_t)gridDim.x * blockDim.x) {
        // counter = absolute quad index; unaligned offsets are handled by drawing per element below
        float r[4];
        #pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint64_t e = offset + (uint64_t)(4 * q + k);
            const uint4 x = philox4x32_10(make_uint4((uint32_t)(e >> 2), (uint32_t)(e >> 34), 0u, 0u), key);
            const uint32_t w[4] = {x.x, x.y, x.z, x.w};
            if (opt == T4K_NORMAL) {
                // Box-Muller on the pair (2*(e&1)) of this counter's 4 words: element parity picks cos/sin
                const uint32_t a = w[(e & 2)], b = w[(e & 2) + 1];
                const float u1 = u01(a), u2 = u01(b);
                const float m = sqrtf(-2.0f * logf(u1));
                r[k] = (e & 1) ? m * sinf(6.2831853071795865f * u2) : m * cosf(6.2831853071795865f * u2);
            } else {
                r[k] = u01(w[e & 3]);
            }
        }
        #pragma unroll
        for (int k = 0; k < 4; k++) { const int64_t i = 4 * q + k; if (i < n) d[i] = scale * (bias + r[k]); }
    }
}
// dropout forward in one pass: the mask draw of k_rand (same counters, same U(0,1] conversion: bit-identical masks) and k_activate's L_DROPOUT
// arithmetic (src/nn/nmath.cu:37-70: keep where u > rate, no rescaling) — one launch and 12 bytes per element instead of two launches and 20
__global__ void __launch_bounds__(T4K_THREADS)
k_dropout_fwd(const float *__restrict__ I, float *O, float *F, float rate, int64_t n, uint64_t seed, uint64_t offset, const uint64_t *__restrict__ epoch) {
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    if (epoch) offset += epoch[0] << 40;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t e = offset + (uint64_t)i;
        const uint4 x = philox4x32_10(make_uint4((uint32_t)(e >> 2), (uint32_t)(e >> 34), 0u, 0u), key);
        const uint32_t w[4] = {x.x, x.y, x.z, x.w};
        const float u = 1.0f * (0.0f + u01(w[e & 3]));           // t4k_rand(..., bias 0, scale 1): the same two roundings
        const bool keep = u > rate;
        O[i] = keep ? I[i] : 0.0f; F[i] = keep ? 1.0f : 0.0f;
    }
}
} // namespace t4k
using namespace t4k;

/* Model::_fstep L_DROPOUT (src/nn/forward.cu:98-102 + k_activate): F <- fresh U(0,1] mask draw, thresholded; O = I where kept, else 0.  Draws
 * exactly what t4k_rand_sharded(F, n, before, global_n, UNIFORM) would (before = 0, global_n = n on a single device) and advances the stream alike. */
extern "C" int t4k_dropout_fwd(const float *I, float *O, float *F, float rate, int64_t n, int64_t before, int64_t global_n, t4k_stream_t s) {
    if (!I || !O || !F || n < 0 || before < 0 || global_n < before + n) return T4K_EINVAL;
    if (n > 0) {
        k_dropout_fwd<<<stream_grid(n), T4K_THREADS, 0, STRM(s)>>>(I, O, F, rate, n, g_seed, g_offset + (uint64_t)before, epoch_ptr());
        const int rc = check_launch(); if (rc) return rc;
    }
    g_offset += (uint64_t)((global_n + 3) & ~3ll);
    return 0;
}

extern "C" int t4k_rand_seed(uint64_t seed) {
    g_seed = seed; g_offset = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && dev < 16 && g_epoch_dev[dev]) cudaMemset(g_epoch_dev[dev], 0, 64);   // synchronising; seeding is rare
    else cudaGetLastError();
    return 0;
}
extern "C" int t4k_rand_at(float *d, int64_t n, int opt, float bias, float scale, uint64_t seed, uint64_t offset, t4k_stream_t s) {
    if (!d || n < 0 || (opt != T4K_UNIFORM && opt != T4K_NORMAL)) return T4K_EINVAL;
    if (n == 0) return 0;
    k_rand<<<stream_grid((n + 3) / 4), T4K_THREADS, 0, STRM(s)>>>(d, n, opt, bias, scale, seed, offset, nullptr);
    return check_launch();
}
extern "C" int t4k_rand_tick(t4k_stream_t s) {
    uint64_t *e = epoch_ptr();
    if (!e) return T4K_ENOMEM;
    k_rand_tick<<<1, 32, 0, STRM(s)>>>(e);
    // draws behind a tick are numbered from 0 again: a step's masks depend on (seed, step count, position inside the step) only — not on how many
    // draws the process made before, nor on whether an earlier step ran eagerly or as a captured graph (whose host-side offsets never advance)
    g_offset = 0;
    return check_launch();
}
/* Draw this rank's shard of a batch-major tensor exactly as ONE device holding the whole batch would draw it (per-rank Philox
 * offset = global element index): `n` = this shard's elements, `before` = elements of the shards of the ranks in front of it, `global_n` =
 * elements of the whole tensor.  Element i of the shard gets the counter of global element before + i; the host offset advances by the global
 * length on every rank alike, so successive draws stay aligned across ranks and with a single-device run of the same program. */
extern "C" int t4k_rand_sharded(float *d, int64_t n, int64_t before, int64_t global_n, int opt, float bias, float scale, t4k_stream_t s) {
    if (!d || n < 0 || before < 0 || global_n < before + n || (opt != T4K_UNIFORM && opt != T4K_NORMAL)) return T4K_EINVAL;
    if (n > 0) {
        k_rand<<<stream_grid((n + 3) / 4), T4K_THREADS, 0, STRM(s)>>>(d, n, opt, bias, scale, g_seed, g_offset + (uint64_t)before, epoch_ptr());
        const int rc = check_launch(); if (rc) return rc;
    }
    g_offset += (uint64_t)((global_n + 3) & ~3ll);
    return 0;
}
extern "C" int t4k_rand(float *d, int64_t n, int opt, float bias, float scale, t4k_stream_t s) {
    if (!d || n < 0 || (opt != T4K_UNIFORM && opt != T4K_NORMAL)) return T4K_EINVAL;
    if (n == 0) return 0;
    k_rand<<<stream_grid((n + 3) / 4), T4K_THREADS, 0, STRM(s)>>>(d, n, opt, bias, scale, g_seed, g_offset, epoch_ptr());
    g_offset += (uint64_t)((n + 3) & ~3ll);        // successive calls draw disjoint counters
    return check_launch();
}
