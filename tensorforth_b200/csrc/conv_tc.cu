// conv_tc.cu — stride-1 "same" conv2d (forward and input-gradient) as an implicit GEMM on the 5th-gen
// tensor cores (tcgen05, sm_100a), FP32 in / FP32 out through the 3xTF32 split.
//   replaces k_conv2d<TS,KS,1,P> (src/nn/nmath.tcu:34-104) and the dX part of k_dconv2d<TS,KS,1,P>
//   (src/nn/nmath.tcu:295-324) for channel counts where the problem is a real GEMM (C1 % 32 == 0).
//
//   O[m=(n,y,x), co] = bias[co] + Σ_tap Σ_ci X[n, y+ky-P, x+kx-P, ci] * Wt[tap][co][ci]
//     forward : X = I  (C1 channels), Wt[tap][c0][c1] = F[c1,ky,kx,c0]
//     dgrad   : X = dO (C0 channels), Wt[tap][c1][c0] = F[c1,ky,kx,c0]   — the reference's 180°-flipped
//               taps (nmath.tcu:304) cancel the flip of the analytic gradient, so dX is a plain
//               correlation of dO with the channel-transposed filter.
//
// GEMM view: M = 128 output pixels per tile, N = CO (<=128), K = taps x CI.  N is small, so an
// operand-from-shared-memory MMA would be bound by the A-tile reads (6 KB per 32-cycle MMA at N=64
// = 192 B/clk > the 128 B/clk of shared memory).  Instead the A operand is produced IN REGISTERS and
// handed to the tensor core through TENSOR MEMORY:
//   warp  13    loader      : per k-stage (tap, 64-channel chunk) one elected thread issues TMA:
//                             cp.async.bulk.tensor.2d of the raw FP32 activation tile — X viewed as
//                             [pixels, CI], box 128 pixels x 32 channels, SWIZZLE_128B; a tap is just a
//                             shift of the flattened pixel coordinate, out-of-tensor rows are zero
//                             filled by TMA — plus cp.async.bulk of the pre-split hi/lo weight tile
//                             image (ready-made SWIZZLE_128B K-major UMMA layout); 3-stage mbarrier ring
//   warps 0-7   A producers : each thread owns one pixel row: conflict-free LDS.128 of its row from the
//                             swizzled tile (zero for padding pixels), split hi/lo in registers,
//                             tcgen05.st → TMEM (double buffered)
//   warps 8-11  epilogue    : tcgen05.ld the two accumulators, add (+bias), 128-bit stores
//   warp  12    MMA issuer  : tcgen05.mma kind::tf32, A from TMEM, B (weights) from shared memory;
//                             hi·hi accumulates in one TMEM accumulator, the two cross terms
//                             (lo·hi, hi·lo) in a second one — the tensor core's accumulator add
//                             truncates (round-toward-zero), so keeping the small terms apart and
//                             the chains short keeps FP32-grade accuracy
// Persistent: one CTA per SM, static round-robin over pixel tiles; accumulators double buffered so
// the epilogue of tile i overlaps the MMAs of tile i+1.
// Roofline: tensor pipe (3 MMAs per k-step); HBM traffic = read X once + write O once.
#include "tc_ptx.cuh"
#include <cuda.h>

namespace t4k {

constexpr int CT_THREADS = 448;                 // 14 warps
constexpr int CT_NSTAGE_MAX = 4;

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}

struct ConvTcP {
    const float *Wimg;       // weight tile images [taps*nchunk][2 planes][CK/32][CO][32] (swizzled)
    const float *bias;       // [CO] or nullptr
    float *Y;                // [N,H,W,CO]
    int H, W, CI, CO, KS, P;
    int64_t Mg;              // N*H*W output pixels
    int ntiles;              // ceil(Mg/128)
    int nchunk;              // CI / CK
    int nstage;              // weight ring depth
    int nacc;                // accumulator buffers (2 if CO <= 64 else 1)
};

// ---------------------------------------------------------------- weight images
// Wt[tap][n][k] → image(tap, chunk)[plane][khalf][n][32] with the 16-byte chunk index XOR (n & 7)
// mode 0 (forward): n = c0, k = c1, value F[c1,ky,kx,c0];  mode 1 (dgrad): n = c1, k = c0, same value.
__global__ void __launch_bounds__(256) k_conv_wimg(const float *__restrict__ F, float *__restrict__ img,
                                                   int C1, int C0, int KS, int CK, int mode) {
    const int CI = mode ? C0 : C1, CO = mode ? C1 : C0;
    const int taps = KS * KS;
    const int64_t total = (int64_t)taps * CI * CO;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(t % CI); const int n = (int)((t / CI) % CO); const int tap = (int)(t / ((int64_t)CI * CO));
        const int c1 = mode ? n : k, c0 = mode ? k : n;
        const float x = __ldg(F + ((int64_t)c1 * taps + tap) * C0 + c0);
        uint32_t hi, lo; split_tf32(x, hi, lo);
        const int chunk = k / CK, kk = k % CK, khalf = kk >> 5, k32 = kk & 31;
        const int64_t stage_flts = (int64_t)2 * CK * CO;                       // hi + lo
        const int64_t plane_flts = (int64_t)CK * CO;
        const int64_t o = ((int64_t)tap * (CI / CK) + chunk) * stage_flts + (int64_t)khalf * CO * 32 +
                          (int64_t)n * 32 + ((((k32 >> 2) ^ (n & 7)) << 2) | (k32 & 3));
        img[o] = __uint_as_float(hi);
        img[o + plane_flts] = __uint_as_float(lo);
    }
}

// ---------------------------------------------------------------- the kernel
template<int CK>
__global__ void __launch_bounds__(CT_THREADS, 1) k_conv_tc(ConvTcP p, const __grid_constant__ CUtensorMap xmap) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    constexpr uint32_t A_BYTES = 128u * CK * 4u;                    // raw activation tile: CK/32 swizzled [128 x 32] halves
    const uint32_t W_BYTES = (uint32_t)p.CO * CK * 8u;              // weight image hi + lo
    const uint32_t STAGE_BYTES = A_BYTES + W_BYTES;
    const uint32_t PLANE_BYTES = (uint32_t)p.CO * CK * 4u;
    uint64_t *bars = (uint64_t*)(smem + (size_t)p.nstage * STAGE_BYTES);
    // barrier map: w_full[4] (stage landed: TMA tx) w_empty[4] (8 producer warps + MMA commit) a_full[2] a_empty[2] acc_full[2] acc_empty[2]
    const uint32_t w_full = smem_u32(bars), w_empty = w_full + 8 * CT_NSTAGE_MAX;
    const uint32_t a_full = w_empty + 8 * CT_NSTAGE_MAX, a_empty = a_full + 16;
    const uint32_t acc_full = a_empty + 16, acc_empty = acc_full + 16;
    uint32_t *tmem_slot = (uint32_t*)(bars + 2 * CT_NSTAGE_MAX + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int taps = p.KS * p.KS;
    const int nst = taps * p.nchunk;                       // k stages per tile
    const int CO = p.CO;

    if (warp == 13 && lane == 0) {
        for (int s = 0; s < CT_NSTAGE_MAX; s++) { mbar_init(w_full + 8 * s, 1); mbar_init(w_empty + 8 * s, 9); }
        for (int s = 0; s < 2; s++) {
            mbar_init(a_full + 8 * s, 8);  mbar_init(a_empty + 8 * s, 1);
            mbar_init(acc_full + 8 * s, 1); mbar_init(acc_empty + 8 * s, 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 12) tmem_alloc(smem_u32(tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // TMEM columns: [0, 4*CK) two A buffers {hi[CK], lo[CK]};  [256, 512) accumulators {main[CO], cross[CO]} x nacc
    const uint32_t TM_A = tmem_base, TM_ACC = tmem_base + 256;

    if (warp < 8) {
        // ================= A producers =================
        const int q = warp & 3, h = warp >> 2;             // TMEM lane quarter, channel half of the chunk
        constexpr int CH = CK / 2;                         // channels per thread per stage
        constexpr int NV = CH / 4;                         // 16-byte chunks per thread
        const uint32_t lane_addr = ((uint32_t)(q * 32) << 16);
        const int r = q * 32 + lane;                       // row of the tile = TMEM lane
        // this thread's channels [h*CH, h*CH+CH) of the chunk live in 32-channel half `hf` at 16-byte chunks [cb, cb+NV)
        const int hf = (h * CH) >> 5, cb = ((h * CH) & 31) >> 2;
        const uint8_t *rowp = smem + (size_t)hf * (128 * 128) + (size_t)r * 128;
        int gs = 0;                                        // global stage counter
        for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
            const int64_t m = (int64_t)tile * 128 + r;
            const bool valid = m < p.Mg;
            const int x = (int)(m % p.W); const int y = (int)((m / p.W) % p.H);
            for (int s = 0; s < nst; s++, gs++) {
                const int tap = s / p.nchunk;
                const int dy = tap / p.KS - p.P, dx = tap % p.KS - p.P;
                const bool inb = valid && (unsigned)(y + dy) < (unsigned)p.H && (unsigned)(x + dx) < (unsigned)p.W;
                const int ab = gs & 1, rs = gs % p.nstage;
                mbar_wait(w_full + 8 * rs, (gs / p.nstage) & 1);           // raw tile landed
                float4 v[NV];
                const uint8_t *src = rowp + (size_t)rs * STAGE_BYTES;
                #pragma unroll
                for (int j = 0; j < NV; j++)
                    v[j] = inb ? *reinterpret_cast<const float4*>(src + (((cb + j) ^ (r & 7)) << 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
                mbar_wait(a_empty + 8 * ab, ((gs >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t col = TM_A + lane_addr + (uint32_t)(ab * 2 * CK + h * CH);
                #pragma unroll
                for (int j0 = 0; j0 < NV; j0 += 4) {       // 16 columns per tcgen05.st
                    uint32_t hi[16], lo[16];
                    #pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const float4 w = v[j0 + j];
                        split_tf32(w.x, hi[4 * j + 0], lo[4 * j + 0]); split_tf32(w.y, hi[4 * j + 1], lo[4 * j + 1]);
                        split_tf32(w.z, hi[4 * j + 2], lo[4 * j + 2]); split_tf32(w.w, hi[4 * j + 3], lo[4 * j + 3]);
                    }
                    tmem_st16(col + j0 * 4, hi);
                    tmem_st16(col + CK + j0 * 4, lo);
                }
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { mbar_arrive(a_full + 8 * ab); mbar_arrive(w_empty + 8 * rs); }
            }
        }
    } else if (warp < 12) {
        // ================= epilogue =================
        const int q = warp & 3;
        const uint32_t lane_addr = ((uint32_t)(q * 32) << 16);
        int it = 0;
        for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, it++) {
            const int ab = (p.nacc == 2) ? (it & 1) : 0;
            const int use = (p.nacc == 2) ? (it >> 1) : it;
            mbar_wait(acc_full + 8 * ab, use & 1);
            tc_fence_after();
            const int64_t m = (int64_t)tile * 128 + q * 32 + lane;
            float *o = p.Y + m * CO;
            const uint32_t acc = TM_ACC + lane_addr + (uint32_t)(ab * 2 * CO);
            for (int c = 0; c < CO; c += 16) {
                uint32_t a[16], b[16];
                tmem_ld16(acc + c, a);
                tmem_ld16(acc + CO + c, b);
                tmem_ld_wait();
                if (m < p.Mg) {
                    #pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        float4 r;
                        r.x = __uint_as_float(a[j]) + __uint_as_float(b[j]);
                        r.y = __uint_as_float(a[j + 1]) + __uint_as_float(b[j + 1]);
                        r.z = __uint_as_float(a[j + 2]) + __uint_as_float(b[j + 2]);
                        r.w = __uint_as_float(a[j + 3]) + __uint_as_float(b[j + 3]);
                        if (p.bias) {
                            const float4 bv = ldg4(p.bias + c + j);
                            r.x += bv.x; r.y += bv.y; r.z += bv.z; r.w += bv.w;
                        }
                        stg4(o + c + j, r);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty + 8 * ab);
        }
    } else if (warp == 12) {
        // ================= MMA issuer =================
        const uint32_t idesc = idesc_tf32(128, CO);
        int gs = 0, it = 0;
        for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, it++) {
            const int accb = (p.nacc == 2) ? (it & 1) : 0;
            const int use = (p.nacc == 2) ? (it >> 1) : it;
            mbar_wait(acc_empty + 8 * accb, (use & 1) ^ 1);
            tc_fence_after();
            const uint32_t d_main = TM_ACC + (uint32_t)(accb * 2 * CO), d_cross = d_main + (uint32_t)CO;
            for (int s = 0; s < nst; s++, gs++) {
                const int ab = gs & 1, ws = gs % p.nstage;
                mbar_wait(w_full + 8 * ws, (gs / p.nstage) & 1);
                mbar_wait(a_full + 8 * ab, (gs >> 1) & 1);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t sb = smem_u32(smem + (size_t)ws * STAGE_BYTES) + A_BYTES;
                    const uint32_t a_hi = TM_A + (uint32_t)(ab * 2 * CK), a_lo = a_hi + CK;
                    #pragma unroll
                    for (int ks = 0; ks < CK / 8; ks++) {
                        const uint32_t boff = (uint32_t)(ks >> 2) * (uint32_t)CO * 128u + (uint32_t)(ks & 3) * 32u;
                        const uint64_t b_hi = smem_desc_sw128(sb + boff), b_lo = smem_desc_sw128(sb + PLANE_BYTES + boff);
                        const uint32_t first = (s | ks) ? 1u : 0u;
                        tc_mma_tf32_ts(d_cross, a_lo + ks * 8, b_hi, idesc, first);
                        tc_mma_tf32_ts(d_cross, a_hi + ks * 8, b_lo, idesc, 1u);
                        tc_mma_tf32_ts(d_main,  a_hi + ks * 8, b_hi, idesc, first);
                    }
                }
                __syncwarp();
                if (elect_one()) {
                    tc_commit(a_empty + 8 * ab);
                    tc_commit(w_empty + 8 * ws);
                    if (s == nst - 1) tc_commit(acc_full + 8 * accb);
                }
                __syncwarp();
            }
        }
    } else {
        // ================= loader (TMA) =================
        if (lane == 0) {
            int gs = 0;
            for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
                const int64_t m0 = (int64_t)tile * 128;
                for (int s = 0; s < nst; s++, gs++) {
                    const int rs = gs % p.nstage;
                    const int tap = s / p.nchunk, chunk = s - tap * p.nchunk;
                    const int dy = tap / p.KS - p.P, dx = tap % p.KS - p.P;
                    mbar_wait(w_empty + 8 * rs, ((gs / p.nstage) & 1) ^ 1);
                    mbar_expect_tx(w_full + 8 * rs, STAGE_BYTES);
                    const uint32_t sa = smem_u32(smem + (size_t)rs * STAGE_BYTES);
                    const int prow = (int)(m0 + (int64_t)dy * p.W + dx);    // shifted flattened pixel coordinate (may be < 0: zero filled)
                    #pragma unroll
                    for (int hf = 0; hf < CK / 32; hf++)
                        tma_load_2d(sa + hf * (128 * 128), &xmap, chunk * CK + hf * 32, prow, w_full + 8 * rs);
                    bulk_g2s(sa + A_BYTES, p.Wimg + (size_t)s * (W_BYTES / 4), W_BYTES, w_full + 8 * rs);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 12) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// eligibility of the tensor path: stride 1, "same" padding, square geometry preserved, GEMM-sized channels
bool conv_tc_ok(int H1, int W1, int CI, int H0, int W0, int CO, int KS, int S, int P) {
    return S == 1 && 2 * P == KS - 1 && H0 == H1 && W0 == W1 && (CI % 32) == 0 && (CO % 16) == 0 && CO >= 16 && CO <= 128 && CI >= 32;
}

// X [N,H,W,CI] → Y [N,H,W,CO]; F is the reference filter [C1,KS,KS,C0]; mode 0 forward (CI=C1,CO=C0), 1 dgrad (CI=C0,CO=C1)
int conv_tc(const float *X, const float *F, const float *bias, float *Y, int N, int H, int W, int C1, int C0,
            int KS, int P, int mode, cudaStream_t st) {
    const int CI = mode ? C0 : C1, CO = mode ? C1 : C0;
    const int CK = (CI % 64 == 0) ? 64 : 32;
    const int taps = KS * KS;
    const size_t img_bytes = (size_t)taps * CI * CO * 8;
    float *img = (float*)workspace(img_bytes, 6);
    if (!img) return T4K_ENOMEM;
    {
        const int64_t total = (int64_t)taps * CI * CO;
        k_conv_wimg<<<(int)((total + 255) / 256), 256, 0, st>>>(F, img, C1, C0, KS, CK, mode);
        int rc = check_launch(); if (rc) return rc;
    }
    ConvTcP p{};
    p.Wimg = img; p.bias = bias; p.Y = Y;
    p.H = H; p.W = W; p.CI = CI; p.CO = CO; p.KS = KS; p.P = P;
    p.Mg = (int64_t)N * H * W;
    p.ntiles = (int)((p.Mg + 127) / 128);
    p.nchunk = CI / CK;
    if (p.Mg >= (1LL << 31) - 65536) return T4K_EINVAL;        // TMA coordinates are int32
    const size_t stage = (size_t)128 * CK * 4 + (size_t)CO * CK * 8;
    int nstage = (int)((200 * 1024) / stage);
    if (nstage > 3) nstage = 3;
    if (nstage < 2) return T4K_EINVAL;
    p.nstage = nstage;
    p.nacc = (CO <= 64) ? 2 : 1;
    // X as a 2-D tensor [Mg pixels, CI channels]; box = 128 pixels x 32 channels (one 128-byte swizzle row per pixel)
    static PFN_encodeTiled enc = nullptr;
    if (!enc) {
        void *fp = nullptr; cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qr) != cudaSuccess || !fp) { cudaGetLastError(); return T4K_ENOSUP; }
        enc = (PFN_encodeTiled)fp;
    }
    CUtensorMap xmap;
    const cuuint64_t gdim[2] = {(cuuint64_t)CI, (cuuint64_t)p.Mg};
    const cuuint64_t gstr[1] = {(cuuint64_t)CI * 4};
    const cuuint32_t box[2] = {32, 128}, estr[2] = {1, 1};
    if (enc(&xmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)X, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return T4K_EINVAL;
    const size_t smem = (size_t)nstage * stage + 1024 + 256;
    int grid = sm_count(); if (grid > p.ntiles) grid = p.ntiles;
    static size_t attr32 = 0, attr64 = 0;
    if (CK == 64) {
        if (smem > attr64) { cudaError_t e = cudaFuncSetAttribute(k_conv_tc<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); if (e != cudaSuccess) return (int)e; attr64 = smem; }
        k_conv_tc<64><<<grid, CT_THREADS, smem, st>>>(p, xmap);
    } else {
        if (smem > attr32) { cudaError_t e = cudaFuncSetAttribute(k_conv_tc<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); if (e != cudaSuccess) return (int)e; attr32 = smem; }
        k_conv_tc<32><<<grid, CT_THREADS, smem, st>>>(p, xmap);
    }
    return check_launch();
}

} // namespace t4k
