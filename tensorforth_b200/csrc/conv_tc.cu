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


// ====================================================================== weight gradient on tcgen05
//   dF[c1,tap,c0] += mult(tap) * Σ_pix I[pix + off(tap), c1] * dO[pix, c0]     (k_dconv2d dF part, nmath.tcu:307-336)
//   dB[c0]        += Σ_pix dO[pix, c0]                                          (nmath.tcu:277-283)
// GEMM view per CTA (split-K over contiguous pixel ranges, 64-pixel k-blocks):
//   M-tile j = 128 rows = (128/C1 taps) x C1 input channels, N = C0, K = pixels.
//   A (TMEM)  : thread = (tap, c1) row; LDS.32 of its channel from ONE raw activation "superset" tile that
//               covers the k-block plus the halo of all taps (TMA, 128-byte swizzle), padding by a
//               per-(tap,pixel) validity bit mask; split hi/lo → tcgen05.st
//   B (smem)  : the dO tile transposed to K-major and split hi/lo once per k-block by converter warps
//               (conflict-free LDS.32 → STS.128 into the SWIZZLE_128B UMMA layout), shared by all taps
//   D (TMEM)  : one FP32 accumulator per M-tile (<= 5 x C0 columns); every FB k-blocks the converter
//               warps drain them into the CTA's partial dF in global memory (plain FP32 RN adds) so the
//               tensor core's truncating accumulate never sees chains longer than FB*24 MMAs
//   k_wgrad_tc_fin sums the per-CTA partials in CTA order (deterministic) into dF / dB.
constexpr int WG_KB = 64;                        // pixels per k-block
struct WgP {
    float *part;             // [grid][taps*C1][C0]
    float *partB;            // [grid][2][C0]
    int H, W, C1, C0, KS, P;
    int64_t Mg;
    int nkb;                 // ceil(Mg / 64)
    int SR;                  // rows of the activation superset tile (multiple of 8, <= 192)
    int NMT, TPM, taps;      // M-tiles, taps per M-tile
    int FB;                  // k-blocks between accumulator drains
    uint32_t stage_bytes, off_o, off_bhi, off_blo, off_mask;
};

__global__ void __launch_bounds__(CT_THREADS, 1) k_conv_wgrad_tc(WgP p, const __grid_constant__ CUtensorMap imap,
                                                                  const __grid_constant__ CUtensorMap omap) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = (uint64_t*)(smem + 2 * (size_t)p.stage_bytes);
    // full[2] empty[2] conv_full[2] a_full[2] a_empty[2] acc_full acc_empty
    const uint32_t b_full = smem_u32(bars), b_empty = b_full + 16, b_conv = b_full + 32, a_full = b_full + 48,
                   a_empty = b_full + 64, acc_full = b_full + 80, acc_empty = b_full + 88;
    uint32_t *tmem_slot = (uint32_t*)(bars + 12);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int C0 = p.C0, C1 = p.C1;
    // this CTA's contiguous k-block range
    const int kb0 = (int)((int64_t)p.nkb * blockIdx.x / gridDim.x), kb1 = (int)((int64_t)p.nkb * (blockIdx.x + 1) / gridDim.x);
    const int nkb = kb1 - kb0;
    const int HALO = p.P * p.W + p.P;

    if (warp == 13 && lane == 0) {
        for (int s = 0; s < 2; s++) {
            mbar_init(b_full + 8 * s, 1); mbar_init(b_empty + 8 * s, 9); mbar_init(b_conv + 8 * s, 4);
            mbar_init(a_full + 8 * s, 4); mbar_init(a_empty + 8 * s, 1);
        }
        mbar_init(acc_full, 1); mbar_init(acc_empty, 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 12) tmem_alloc(smem_u32(tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t TM_A = tmem_base, TM_ACC = tmem_base + 128;      // A: 2 buffers x {hi[32], lo[32]}; accumulators: NMT x C0 columns

    if (warp < 8) {
        // ================= A producers: group g fills TMEM buffer g =================
        const int q = warp & 3, g = warp >> 2;
        const uint32_t lane_addr = ((uint32_t)(q * 32) << 16);
        const int L = q * 32 + lane;
        const int tl = L / C1, c1 = L % C1;                 // tap within the M-tile (warp uniform: C1 >= 32), input channel
        const uint32_t coff = (uint32_t)(c1 >> 5) * (uint32_t)p.SR * 128u;   // 32-channel half of the superset tile
        const int ch16 = (c1 & 31) >> 2, cw = (c1 & 3) * 4;
        const int nas = 2 * p.NMT;                          // A stages per k-block
        for (int i = 0; i < nkb; i++) {
            const int st = i & 1;
            mbar_wait(b_full + 8 * st, (i >> 1) & 1);
            mbar_wait(b_conv + 8 * st, (i >> 1) & 1);       // masks written
            const uint8_t *stage = smem + (size_t)st * p.stage_bytes;
            const uint32_t *masks = (const uint32_t*)(stage + p.off_mask);
            for (int a = g; a < nas; a += 2) {              // A stage a = kh * NMT + j
                const int kh = a / p.NMT, j = a - kh * p.NMT;
                const int tap = j * p.TPM + tl;
                const int sidx = i * nas + a;               // global A-stage counter; buffer = sidx & 1 = g
                uint32_t hi[32], lo[32];
                if (tap < p.taps) {
                    const uint32_t mask = masks[tap * 2 + kh];
                    const int row0 = kh * 32 + (tap / p.KS) * p.W + (tap % p.KS);     // superset row of pixel k=0: k + off(tap) + HALO
                    const uint8_t *base = stage + coff + cw;
                    #pragma unroll
                    for (int k = 0; k < 32; k++) {
                        const int r = row0 + k;
                        const float v = ((mask >> k) & 1u) ? *reinterpret_cast<const float*>(base + r * 128 + ((ch16 ^ (r & 7)) << 4)) : 0.0f;
                        split_tf32(v, hi[k], lo[k]);
                    }
                } else {
                    #pragma unroll
                    for (int k = 0; k < 32; k++) { hi[k] = 0u; lo[k] = 0u; }
                }
                mbar_wait(a_empty + 8 * g, ((sidx >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t col = TM_A + lane_addr + (uint32_t)(g * 64);
                tmem_st16(col,      *reinterpret_cast<uint32_t(*)[16]>(&hi[0]));
                tmem_st16(col + 16, *reinterpret_cast<uint32_t(*)[16]>(&hi[16]));
                tmem_st16(col + 32, *reinterpret_cast<uint32_t(*)[16]>(&lo[0]));
                tmem_st16(col + 48, *reinterpret_cast<uint32_t(*)[16]>(&lo[16]));
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(a_full + 8 * g);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(b_empty + 8 * st);   // done reading this k-block's raw tile and masks
        }
    } else if (warp < 12) {
        // ================= converters (dO → K-major hi/lo, masks, dB) + accumulator drain =================
        const int q = warp & 3, t = threadIdx.x - 256;      // t in [0,128)
        const uint32_t lane_addr = ((uint32_t)(q * 32) << 16);
        const int nco = (C0 + 31) >> 5;                     // 32-channel halves of dO
        float dbsum[2] = {0.f, 0.f};
        int nflush = 0;
        auto drain = [&]() {
            mbar_wait(acc_full, nflush & 1);
            tc_fence_after();
            const int L = q * 32 + lane;
            for (int j = 0; j < p.NMT; j++) {
                const bool vrow = (j * p.TPM + L / C1) < p.taps;
                float *dst = p.part + ((size_t)blockIdx.x * p.taps * C1 + (size_t)j * 128 + L) * C0;
                for (int c = 0; c < C0; c += 16) {
                    uint32_t v[16];
                    tmem_ld16(TM_ACC + lane_addr + (uint32_t)(j * C0 + c), v);
                    tmem_ld_wait();
                    if (vrow) {
                        #pragma unroll
                        for (int e = 0; e < 16; e += 4) {
                            float4 r = make_float4(__uint_as_float(v[e]), __uint_as_float(v[e + 1]), __uint_as_float(v[e + 2]), __uint_as_float(v[e + 3]));
                            if (nflush) { const float4 o = *reinterpret_cast<const float4*>(dst + c + e); r.x += o.x; r.y += o.y; r.z += o.z; r.w += o.w; }
                            *reinterpret_cast<float4*>(dst + c + e) = r;
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty);
            nflush++;
        };
        for (int i = 0; i < nkb; i++) {
            const int st = i & 1;
            mbar_wait(b_full + 8 * st, (i >> 1) & 1);
            uint8_t *stage = smem + (size_t)st * p.stage_bytes;
            // --- validity masks: bit k of masks[tap*2 + kh] = pixel (p0 + 32*kh + k) + tap offset lies inside its image
            if (t < 64) {
                const int64_t pix = (int64_t)(kb0 + i) * WG_KB + t;
                const bool vp = pix < p.Mg;
                const int x = (int)(pix % p.W), y = (int)((pix / p.W) % p.H);
                uint32_t *masks = (uint32_t*)(stage + p.off_mask);
                for (int tap = 0; tap < p.taps; tap++) {
                    const int dy = tap / p.KS - p.P, dx = tap % p.KS - p.P;
                    const bool ok = vp && (unsigned)(y + dy) < (unsigned)p.H && (unsigned)(x + dx) < (unsigned)p.W;
                    const uint32_t b = __ballot_sync(0xffffffffu, ok);
                    if (lane == 0) masks[tap * 2 + (t >> 5)] = b;
                }
            }
            // --- B: thread (kh = pixel half, c0) transposes 32 pixels of its channel into the K-major swizzled image
            for (int u = t; u < 2 * C0; u += 128) {
                const int kh = u / C0, c0 = u - kh * C0;
                const uint8_t *src = stage + p.off_o + (size_t)(c0 >> 5) * (WG_KB * 128) + (c0 & 3) * 4;
                const int ch16 = (c0 & 31) >> 2;
                uint8_t *bh = stage + p.off_bhi + (size_t)kh * C0 * 128 + (size_t)c0 * 128;
                uint8_t *bl = stage + p.off_blo + (size_t)kh * C0 * 128 + (size_t)c0 * 128;
                float s = 0.f;
                #pragma unroll
                for (int pc = 0; pc < 8; pc++) {
                    uint32_t hi[4], lo[4];
                    #pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const int r = kh * 32 + pc * 4 + e;
                        const float v = *reinterpret_cast<const float*>(src + r * 128 + ((ch16 ^ (r & 7)) << 4));
                        s += v;
                        split_tf32(v, hi[e], lo[e]);
                    }
                    const int sw = (pc ^ (c0 & 7)) << 4;
                    *reinterpret_cast<uint4*>(bh + sw) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<uint4*>(bl + sw) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                }
                dbsum[(u - t) / 128 & 1] += s;
            }
            (void)nco;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes → visible to the tensor core
            __syncwarp();
            if (lane == 0) mbar_arrive(b_conv + 8 * st);
            if (i > 0 && (i % p.FB) == 0) drain();          // accumulators hold k-blocks [i-FB, i); MMA waits for acc_empty before block i
        }
        if (nkb > 0) drain();
        // dB partials: thread u-slot → (kh, c0)
        for (int u = t, k = 0; u < 2 * C0; u += 128, k++)
            p.partB[((size_t)blockIdx.x * 2 + u / C0) * C0 + (u % C0)] = dbsum[k & 1];
    } else if (warp == 12) {
        // ================= MMA issuer =================
        const uint32_t idesc = idesc_tf32(128, C0);
        const int nas = 2 * p.NMT;
        int nwait = 0;
        for (int i = 0; i < nkb; i++) {
            const int st = i & 1;
            const bool fresh = (i % p.FB) == 0;
            if (fresh && i > 0) { mbar_wait(acc_empty, nwait & 1); nwait++; }
            mbar_wait(b_conv + 8 * st, (i >> 1) & 1);
            tc_fence_after();
            const uint32_t sbh = smem_u32(smem + (size_t)st * p.stage_bytes + p.off_bhi), sbl = smem_u32(smem + (size_t)st * p.stage_bytes + p.off_blo);
            for (int a = 0; a < nas; a++) {
                const int kh = a / p.NMT, j = a - kh * p.NMT;
                const int sidx = i * nas + a, ab = sidx & 1;
                mbar_wait(a_full + 8 * ab, (sidx >> 1) & 1);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t d = TM_ACC + (uint32_t)(j * C0);
                    const uint32_t a_hi = TM_A + (uint32_t)(ab * 64), a_lo = a_hi + 32;
                    #pragma unroll
                    for (int ks = 0; ks < 4; ks++) {
                        const uint32_t boff = (uint32_t)kh * (uint32_t)C0 * 128u + (uint32_t)ks * 32u;
                        const uint64_t b_hi = smem_desc_sw128(sbh + boff), b_lo = smem_desc_sw128(sbl + boff);
                        tc_mma_tf32_ts(d, a_lo + ks * 8, b_hi, idesc, (fresh && kh == 0 && ks == 0) ? 0u : 1u);
                        tc_mma_tf32_ts(d, a_hi + ks * 8, b_lo, idesc, 1u);
                        tc_mma_tf32_ts(d, a_hi + ks * 8, b_hi, idesc, 1u);
                    }
                }
                __syncwarp();
                if (elect_one()) {
                    tc_commit(a_empty + 8 * ab);
                    if (a == nas - 1) {
                        tc_commit(b_empty + 8 * st);
                        if (((i + 1) % p.FB) == 0 || i == nkb - 1) tc_commit(acc_full);
                    }
                }
                __syncwarp();
            }
        }
    } else {
        // ================= loader (TMA) =================
        if (lane == 0) {
            const uint32_t bytes = (uint32_t)(C1 >> 5) * (uint32_t)p.SR * 128u + (uint32_t)(C0 >> 5) * (WG_KB * 128u);
            for (int i = 0; i < nkb; i++) {
                const int st = i & 1;
                mbar_wait(b_empty + 8 * st, ((i >> 1) & 1) ^ 1);
                mbar_expect_tx(b_full + 8 * st, bytes);
                const uint32_t sa = smem_u32(smem + (size_t)st * p.stage_bytes);
                const int p0 = (kb0 + i) * WG_KB;
                for (int hf = 0; hf < (C1 >> 5); hf++) tma_load_2d(sa + hf * p.SR * 128, &imap, hf * 32, p0 - HALO, b_full + 8 * st);
                for (int hf = 0; hf < (C0 >> 5); hf++) tma_load_2d(sa + p.off_o + hf * (WG_KB * 128), &omap, hf * 32, p0, b_full + 8 * st);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 12) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

__host__ __device__ __forceinline__ int dconv_mult(int KS, int S, int tap) {      // see conv.cu: dconv_flush_mult
    const int TS = (16 - KS + S) / S;
    int m = 0;
    for (int ty = 0; ty < 16; ty++) { const int tx = tap - ty * TS; if (tx >= 0 && tx < 16) m++; }
    return m;
}
__global__ void __launch_bounds__(256) k_wgrad_tc_fin(const float *__restrict__ part, const float *__restrict__ partB, float *dF, float *dB,
                                                      int C1, int C0, int taps, int KS, int nparts) {
    const int nF = C1 * taps * C0;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nF) {
        const int c0 = t % C0, tap = (t / C0) % taps, c1 = t / (C0 * taps);
        const size_t src = ((size_t)tap * C1 + c1) * C0 + c0;
        float s = 0.f;
        for (int k = 0; k < nparts; k++) s += part[(size_t)k * nF + src];
        dF[t] += s * (float)dconv_mult(KS, 1, tap);
    } else if (t < nF + C0) {
        const int c0 = t - nF;
        float s = 0.f;
        for (int k = 0; k < 2 * nparts; k++) s += partB[(size_t)k * C0 + c0];
        dB[c0] += s;
    }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);


bool conv_wgrad_tc_ok(int H1, int W1, int C1, int H0, int W0, int C0, int KS, int S, int P) {
    if (!(S == 1 && 2 * P == KS - 1 && H0 == H1 && W0 == W1)) return false;
    if (!(KS == 1 || KS == 3)) return false;
    if (!(C1 == 32 || C1 == 64) || !(C0 == 32 || C0 == 64)) return false;
    return WG_KB + 2 * (P * W1 + P) <= 192;
}

static PFN_encodeTiled tmap_encoder() {
    static PFN_encodeTiled enc = nullptr;
    if (!enc) {
        void *fp = nullptr; cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qr) != cudaSuccess || !fp) { cudaGetLastError(); return nullptr; }
        enc = (PFN_encodeTiled)fp;
    }
    return enc;
}
// 2-D view [rows, C] of an NHWC tensor, box = box_rows x 32 channels, 128-byte swizzle, zero fill out of bounds
static int tmap_rows(CUtensorMap *m, const float *X, int64_t rows, int C, int box_rows) {
    PFN_encodeTiled enc = tmap_encoder();
    if (!enc) return T4K_ENOSUP;
    const cuuint64_t gdim[2] = {(cuuint64_t)C, (cuuint64_t)rows};
    const cuuint64_t gstr[1] = {(cuuint64_t)C * 4};
    const cuuint32_t box[2] = {32, (cuuint32_t)box_rows}, estr[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)X, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS ? 0 : T4K_EINVAL;
}

int conv_wgrad_tc(const float *I, const float *dO, float *dF, float *dB, int N, int H, int W, int C1, int C0,
                  int KS, int P, cudaStream_t st) {
    WgP p{};
    p.H = H; p.W = W; p.C1 = C1; p.C0 = C0; p.KS = KS; p.P = P;
    p.Mg = (int64_t)N * H * W;
    if (p.Mg >= (1LL << 31) - 65536) return T4K_EINVAL;
    p.nkb = (int)((p.Mg + WG_KB - 1) / WG_KB);
    p.taps = KS * KS; p.TPM = 128 / C1; p.NMT = (p.taps + p.TPM - 1) / p.TPM;
    p.SR = (WG_KB + 2 * (P * W + P) + 7) & ~7;
    p.FB = 16;
    uint32_t off = (uint32_t)(C1 >> 5) * p.SR * 128;             p.off_o = off;
    off += (uint32_t)(C0 >> 5) * WG_KB * 128;                    p.off_bhi = off;
    off += 2u * C0 * 128;                                        p.off_blo = off;
    off += 2u * C0 * 128;                                        p.off_mask = off;
    off += 128;
    p.stage_bytes = (off + 1023) & ~1023u;
    int grid = sm_count(); if (grid > p.nkb) grid = p.nkb;
    const size_t nF = (size_t)C1 * p.taps * C0;
    p.part = (float*)workspace((size_t)grid * nF * 4, 4);
    p.partB = (float*)workspace((size_t)grid * 2 * C0 * 4, 5);
    if (!p.part || !p.partB) return T4K_ENOMEM;
    CUtensorMap imap, omap;
    int rc = tmap_rows(&imap, I, p.Mg, C1, p.SR); if (rc) return rc;
    rc = tmap_rows(&omap, dO, p.Mg, C0, WG_KB); if (rc) return rc;
    const size_t smem = 2 * (size_t)p.stage_bytes + 1024 + 256;
    static DevSize attr;
    if (dev_grow(attr, smem)) { cudaError_t e = cudaFuncSetAttribute(k_conv_wgrad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); if (e != cudaSuccess) return (int)e; }
    k_conv_wgrad_tc<<<grid, CT_THREADS, smem, st>>>(p, imap, omap);
    rc = check_launch(); if (rc) return rc;
    const int tot = (int)nF + C0;
    k_wgrad_tc_fin<<<(tot + 255) / 256, 256, 0, st>>>(p.part, p.partB, dF, dB, C1, C0, p.taps, KS, grid);
    return check_launch();
}

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
    CUtensorMap xmap;
    { int rc = tmap_rows(&xmap, X, p.Mg, CI, 128); if (rc) return rc; }
    const size_t smem = (size_t)nstage * stage + 1024 + 256;
    int grid = sm_count(); if (grid > p.ntiles) grid = p.ntiles;
    static DevSize attr32, attr64;
    if (CK == 64) {
        if (dev_grow(attr64, smem)) { cudaError_t e = cudaFuncSetAttribute(k_conv_tc<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); if (e != cudaSuccess) return (int)e; }
        k_conv_tc<64><<<grid, CT_THREADS, smem, st>>>(p, xmap);
    } else {
        if (dev_grow(attr32, smem)) { cudaError_t e = cudaFuncSetAttribute(k_conv_tc<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); if (e != cudaSuccess) return (int)e; }
        k_conv_tc<32><<<grid, CT_THREADS, smem, st>>>(p, xmap);
    }
    return check_launch();
}

} // namespace t4k
