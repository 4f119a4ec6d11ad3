// comm.cu — data-parallel exchange over NVLink / NVSwitch peer memory, fused with the optimizer step
//
// The reference is single-GPU.  Model::forward/backprop shard over the batch; parameter gradients are
// batch SUMS (src/nn/backprop.cu:97-103, nmath.tcu:277,335), so one SUM all-reduce of the flat DG arena between backprop
// and the optimizer (src/nn/gradient.cu:64-126) reproduces the single-GPU step.  That pair — a latency-bound 0.8 MB
// all-reduce followed by a 2.8 us optimizer pass — is ONE kernel here:
//
//   push   every rank stores its chunk of DG into slot[parity][rank] of EVERY peer's exchange block (posted 128-bit
//          stores over NVLink; no round trip), then one `fence.sys` per block + a store of the chunk's epoch flag to each peer
//   wait   spin (acquire loads, local memory) until the chunk's flag from every rank has reached this epoch
//   finish sum the `world` local slots in RANK ORDER (identical bits on every rank, so replicas never drift), run the
//          optimizer step on G/M/V, write DG = 0
//
// Flags are per (rank, chunk): a chunk never waits for more than its own data, there is no grid-wide barrier, and the
// chunk -> slot-range mapping depends only on the communicator's capacity, never on the call's length, so calls of any
// length can be mixed.  Slot reuse is safe without a second barrier: parity p of a chunk is rewritten at that chunk's
// epoch e+2; the writer has by then passed the chunk's wait at e+1, i.e. every peer had launched its e+1 kernel, which
// stream order puts after the end of its epoch-e kernel (the reader of parity p).
// Epoch counters live in device memory, so the kernel is CUDA-graph capturable (Model::step_graph captures the whole
// data-parallel train step: forward + loss + backprop + exchange/optimizer, one graph launch per step).
// One process per GPU: the exchange blocks are cudaMalloc'ed and exported with cudaIpc handles, which the host side
// (tensorforth_b200/dp.py, torch.distributed) gathers.  A wait that sees no progress for T4K_COMM_TIMEOUT_S seconds (default 60: a
// checkpoint save, an evaluation pass or a stalled loader on one rank are legitimate skews) does NOT hang the GPU and does NOT
// apply a half-summed gradient: the chunk's finish is skipped (G / M / V / DG untouched, epoch not advanced) and a STICKY error word
// is raised, in device memory — every later exchange kernel of this communicator returns at once without touching the model —
// and in mapped host memory, which t4k_comm_poll reads without synchronising (Model::_gradient checks it on every optimizer call).
#include "common.cuh"
#include "optim.cuh"
#include <cstring>
#include <cstdlib>
#include <new>

#define COMM_MAXW   8            // ranks (one NVSwitch domain)
#define COMM_MAXB   512          // chunks (= blocks) per call
#define COMM_NSCAL  64           // extra scalars riding in the same exchange (loss sum, hit count …)
#define COMM_FLAGB  ((COMM_MAXW + 1) * COMM_MAXB * 4)       // [rank][chunk] push flags, then [chunk] flags of the owners' broadcasts (flag2)
#define COMM_TIMEOUT_S_DEFAULT 60

struct t4k_comm {
    int rank, world, dev;
    int64_t cap;                 // floats per slot (multiple of 4), scalar tail not included
    int ch4;                     // float4s per chunk
    char *base;                  // this rank's exchange block
    char *peer[COMM_MAXW];       // every rank's block as mapped here (peer[rank] == base)
    bool ipc[COMM_MAXW];
    uint32_t *epoch;             // [COMM_MAXB] per-chunk epoch counters + [COMM_MAXB] sticky error word (local device memory)
    uint32_t *err_host, *err_dev;    // the same error word in mapped pinned host memory (host pointer / device alias): polled without a sync
    float *scal_mirror;              // t4k_comm_scalar_mirror: pinned host floats that receive the summed scalars straight from the exchange kernel
    long long spin_limit;        // clock64 ticks a wait may go without progress
    size_t bytes;
    cudaStream_t cs[COMM_MAXW];  // t4k_dp_push_dma: one copy stream per destination rank (the peer-to-peer copies of a push run side by side)
    cudaEvent_t cfork, cdone[COMM_MAXW];
};

namespace t4k {

struct CommDev {
    char *peer[COMM_MAXW];
    uint32_t *epoch;
    uint32_t *err_host;          // device alias of the mapped host error word
    long long spin_limit;
    int64_t cap;
    int rank, world, ch4;
    float *scal_mirror;          // pinned host memory (unified addressing): the summed scalars are ALSO stored here — no copy node behind the step
};
struct DpOpt { float *G, *M, *V; const t4k_seg_t *seg; int nseg; bool mom; OptP p;
               int b0;                  // first chunk of this launch (push-only launches start past the late chunks)
               int64_t pushed_from; };  // chunks starting at or beyond this float were pushed by an earlier MODE -1 launch of this epoch

__device__ __forceinline__ float *slot_of(const CommDev &c, int where, int par, int r) {
    return reinterpret_cast<float*>(c.peer[where] + COMM_FLAGB) + (int64_t)(par * c.world + r) * (c.cap + COMM_NSCAL);
}
__device__ __forceinline__ uint32_t *flag_of(const CommDev &c, int where, int r, int b) {
    return reinterpret_cast<uint32_t*>(c.peer[where]) + r * COMM_MAXB + b;
}
__device__ __forceinline__ uint32_t *flag2_of(const CommDev &c, int where, int b) {
    return reinterpret_cast<uint32_t*>(c.peer[where]) + COMM_MAXW * COMM_MAXB + b;
}
// landing zone of the owners' broadcasts: the rank-ordered SUM of a chunk, one copy per parity, behind the per-rank slots
__device__ __forceinline__ float *gsum_of(const CommDev &c, int where, int par) {
    return reinterpret_cast<float*>(c.peer[where] + COMM_FLAGB) + (int64_t)2 * c.world * (c.cap + COMM_NSCAL) + (int64_t)par * c.cap;
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) { uint32_t v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }

// MODE 0: buf = Σ_r buf_r (in place).  MODE 1/2/3: sgd / adam / adamw on (G, Σ_r DG_r, M, V), DG = 0.
// MODE -1: push + signal only (no wait, no finish, epoch not advanced): the early half of a split exchange — the gradient
// segments that are final before backprop ends travel while the remaining backward kernels run (Model::step_graph forks
// this launch onto a side stream); the MODE 1/2/3 launch that follows pushes only the chunks below `pushed_from`.
template<int MODE, bool VEC>
__global__ void __launch_bounds__(T4K_THREADS) k_dp_exchange(const __grid_constant__ CommDev c, float *buf, int64_t n, float *scal, int nscal, DpOpt o) {
    __shared__ uint32_t s_ep, s_err;
    pdl_wait(); pdl_trigger();                  // PDL: nothing global before this line
    const int b = blockIdx.x + o.b0, tid = threadIdx.x;
    if (tid == 0) { s_ep = c.epoch[b] + 1; s_err = c.epoch[COMM_MAXB]; }
    __syncthreads();
    if (s_err) return;                          // sticky: an earlier exchange timed out, the replicas are no longer in step — touch nothing
    const uint32_t ep = s_ep;
    const int par = (int)(ep & 1u);
    const int64_t lo = (int64_t)b * c.ch4 * 4;
    const int64_t hi = (lo + (int64_t)c.ch4 * 4 < n) ? lo + (int64_t)c.ch4 * 4 : n;
    // ---- push this rank's chunk to every rank's slot[par][rank] (own slot last: it is the only local store)
    const bool do_push = MODE < 0 || lo < o.pushed_from;
    if (!do_push) { /* pushed and signalled by the early launch */ }
    else if (VEC) {
        for (int64_t i = lo + 4 * tid; i < hi; i += 4 * T4K_THREADS) {
            const float4 v = *reinterpret_cast<const float4*>(buf + i);
            #pragma unroll 1
            for (int k = 1; k <= c.world; k++) {
                const int p = (c.rank + k) % c.world;
                *reinterpret_cast<float4*>(slot_of(c, p, par, c.rank) + i) = v;
            }
        }
    } else {
        for (int64_t i = lo + tid; i < hi; i += T4K_THREADS) {
            const float v = buf[i];
            for (int k = 1; k <= c.world; k++) slot_of(c, (c.rank + k) % c.world, par, c.rank)[i] = v;
        }
    }
    if (MODE >= 0 && b == 0 && tid < nscal) {
        const float v = scal[tid];
        for (int k = 1; k <= c.world; k++) slot_of(c, (c.rank + k) % c.world, par, c.rank)[c.cap + tid] = v;
    }
    __syncthreads();
    // ---- signal every rank, then wait for every rank's signal for this chunk.  ONE fence.sys per block: the barrier orders the
    // block's peer stores before thread 0's fence, and fences are cumulative, so the flag stores that follow publish all of
    // them (measured on 2 B200s, 0.79 MB: a fence in every thread 17.7 us, this 9.9 us; NCCL 14.5 us).
    if (tid == 0 && do_push) {
        __threadfence_system();
        for (int k = 1; k <= c.world; k++) *reinterpret_cast<volatile uint32_t*>(flag_of(c, (c.rank + k) % c.world, c.rank, b)) = ep;
    }
    if (MODE < 0) return;
    if (tid < c.world) {
        const uint32_t *f = flag_of(c, c.rank, tid, b);
        const long long t0 = clock64();
        while ((int32_t)(ld_acquire_sys(f) - ep) < 0) {
            __nanosleep(32);                    // the waiting launch may share its SMs with the rest of backprop (side stream): leave the issue slots to it
            if (clock64() - t0 > c.spin_limit) {
                atomicCAS(c.epoch + COMM_MAXB, 0u, 1u + (uint32_t)tid);           // first failure wins, never cleared
                if (c.err_host) { *reinterpret_cast<volatile uint32_t*>(c.err_host) = 1u + (uint32_t)tid; __threadfence_system(); }
                s_err = 1u + (uint32_t)tid;
                break;
            }
        }
    }
    __syncthreads();
    if (s_err) return;                          // no finish on incomplete data: G / M / V / DG stay as they are, the epoch is not advanced
    // ---- finish: rank-ordered sum of the local slots (+ optimizer)
    const float *mine = slot_of(c, c.rank, par, 0);
    const int64_t sstride = c.cap + COMM_NSCAL;
    if (VEC) {
        for (int64_t i = lo + 4 * tid; i < hi; i += 4 * T4K_THREADS) {
            float4 s = __ldcg(reinterpret_cast<const float4*>(mine + i));
            for (int r = 1; r < c.world; r++) {
                const float4 t = __ldcg(reinterpret_cast<const float4*>(mine + r * sstride + i));
                s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
            }
            if (MODE == 0) { *reinterpret_cast<float4*>(buf + i) = s; }
            else {
                float4 g = *reinterpret_cast<const float4*>(o.G + i), m = make_float4(0, 0, 0, 0), v = m;
                if (MODE == 1) {
                    int l = 0, h = o.nseg - 1;                 // segment of element i (segments are 4-aligned: one lookup per float4)
                    while (l < h) { int mid = (l + h + 1) >> 1; if (o.seg[mid].off <= i) l = mid; else h = mid - 1; }
                    const float nw = (float)o.seg[l].Nw;
                    s.x = s.x / nw; s.y = s.y / nw; s.z = s.z / nw; s.w = s.w / nw;
                    if (o.mom) m = *reinterpret_cast<const float4*>(o.M + i);
                } else { m = *reinterpret_cast<const float4*>(o.M + i); v = *reinterpret_cast<const float4*>(o.V + i); }
                constexpr int K = MODE - 1;
                opt_step<K>(g.x, s.x, m.x, v.x, 1.0f, o.mom, o.p); opt_step<K>(g.y, s.y, m.y, v.y, 1.0f, o.mom, o.p);
                opt_step<K>(g.z, s.z, m.z, v.z, 1.0f, o.mom, o.p); opt_step<K>(g.w, s.w, m.w, v.w, 1.0f, o.mom, o.p);
                *reinterpret_cast<float4*>(o.G + i) = g;
                *reinterpret_cast<float4*>(buf + i) = make_float4(0, 0, 0, 0);
                if (MODE == 1) { if (o.mom) *reinterpret_cast<float4*>(o.M + i) = m; }
                else { *reinterpret_cast<float4*>(o.M + i) = m; *reinterpret_cast<float4*>(o.V + i) = v; }
            }
        }
    } else {                                                   // MODE 0 only (see launcher)
        for (int64_t i = lo + tid; i < hi; i += T4K_THREADS) {
            float s = __ldcg(mine + i);
            for (int r = 1; r < c.world; r++) s += __ldcg(mine + r * sstride + i);
            buf[i] = s;
        }
    }
    if (b == 0 && tid < nscal) {
        float s = __ldcg(mine + c.cap + tid);
        for (int r = 1; r < c.world; r++) s += __ldcg(mine + r * sstride + c.cap + tid);
        scal[tid] = s;
        if (c.scal_mirror) c.scal_mirror[tid] = s;
    }
    if (tid == 0) c.epoch[b] = ep;
}


// ---- reduce-scatter + broadcast of the sums (world > 2).  The all-to-all push above sends every chunk to every rank: (world-1) x the arena
// leaves each GPU through SM stores that share the machine with backprop (8 GPUs: 5.5 MB per rank and step).  Here chunk b has ONE owner,
// rank b % world:
//   k_dp_push_owner   every rank stores its chunk into slot[parity][rank] of the OWNER only and raises the owner's flag: 1 x the arena leaves
//                     the GPU, early, on the side stream;
//   k_dp_exchange_rs  later (rest of the arena: under the first layer's finish launch), one block per chunk on EVERY rank: the owner waits for
//                     the world pushes, sums the slots in RANK ORDER and stores the sum into gsum[parity] of every rank (+ flag2); every rank
//                     — the owner included — then waits for flag2, reads the sum from its own memory and runs the optimizer on the chunk.
// One sum per chunk, computed once: the replicas receive identical bits by construction.  Optimizer state stays replicated (every rank steps
// every chunk), so checkpoints and the single-GPU code paths are unchanged.  Two NVLink hops instead of one: used for the part of the arena
// whose exchange is off the critical path; the first chunk keeps the one-hop all-to-all.
__global__ void __launch_bounds__(T4K_THREADS) k_dp_push_owner(const __grid_constant__ CommDev c, const float *buf, int64_t n, int b0) {
    __shared__ uint32_t s_ep, s_err;
    pdl_wait(); pdl_trigger();
    const int b = blockIdx.x + b0, tid = threadIdx.x;
    if (tid == 0) { s_ep = c.epoch[b] + 1; s_err = c.epoch[COMM_MAXB]; }
    __syncthreads();
    if (s_err) return;
    const uint32_t ep = s_ep;
    const int par = (int)(ep & 1u), own = b % c.world;
    const int64_t lo = (int64_t)b * c.ch4 * 4;
    const int64_t hi = (lo + (int64_t)c.ch4 * 4 < n) ? lo + (int64_t)c.ch4 * 4 : n;
    float *dst = slot_of(c, own, par, c.rank);
    for (int64_t i = lo + 4 * tid; i < hi; i += 4 * T4K_THREADS) *reinterpret_cast<float4*>(dst + i) = *reinterpret_cast<const float4*>(buf + i);
    __syncthreads();
    if (tid == 0) { __threadfence_system(); *reinterpret_cast<volatile uint32_t*>(flag_of(c, own, c.rank, b)) = ep; }
}

// phase 0: both halves in one launch; phase 1: the owners' half only (wait for the pushes, sum, send the sums) — a few blocks per rank, launched
// right behind the push so that the sums are already everywhere when phase 2 (everybody: optimizer on every chunk) runs at the end of backprop
template<int KIND>
__global__ void __launch_bounds__(T4K_THREADS) k_dp_exchange_rs(const __grid_constant__ CommDev c, float *buf, int64_t n, DpOpt o, int phase) {
    __shared__ uint32_t s_ep, s_err;
    pdl_wait(); pdl_trigger();
    const int b = (phase == 1) ? o.b0 + ((c.rank - o.b0 % c.world + c.world) % c.world) + (int)blockIdx.x * c.world     // this rank's blockIdx.x-th own chunk at or past b0
                               : (int)blockIdx.x + o.b0;
    const int tid = threadIdx.x;
    if (phase == 1 && (int64_t)b * c.ch4 * 4 >= n) return;
    if (tid == 0) { s_ep = c.epoch[b] + 1; s_err = c.epoch[COMM_MAXB]; }
    __syncthreads();
    if (s_err) return;
    const uint32_t ep = s_ep;
    const int par = (int)(ep & 1u), own = b % c.world;
    const int64_t lo = (int64_t)b * c.ch4 * 4;
    const int64_t hi = (lo + (int64_t)c.ch4 * 4 < n) ? lo + (int64_t)c.ch4 * 4 : n;
    auto timed_out = [&](int who) {
        atomicCAS(c.epoch + COMM_MAXB, 0u, 1u + (uint32_t)who);
        if (c.err_host) { *reinterpret_cast<volatile uint32_t*>(c.err_host) = 1u + (uint32_t)who; __threadfence_system(); }
        s_err = 1u + (uint32_t)who;
    };
    if (own == c.rank && phase != 2) {
        // ---- owner: the world pushes of this chunk -> rank-ordered sum -> every rank's gsum
        if (tid < c.world) {
            const uint32_t *f = flag_of(c, c.rank, tid, b);
            const long long t0 = clock64();
            while ((int32_t)(ld_acquire_sys(f) - ep) < 0) { __nanosleep(32); if (clock64() - t0 > c.spin_limit) { timed_out(tid); break; } }
        }
        __syncthreads();
        if (s_err) return;
        const float *mine = slot_of(c, c.rank, par, 0);
        const int64_t sstride = c.cap + COMM_NSCAL;
        for (int64_t i = lo + 4 * tid; i < hi; i += 4 * T4K_THREADS) {
            float4 s = __ldcg(reinterpret_cast<const float4*>(mine + i));
            for (int r = 1; r < c.world; r++) {
                const float4 t = __ldcg(reinterpret_cast<const float4*>(mine + r * sstride + i));
                s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
            }
            #pragma unroll 1
            for (int k = 1; k <= c.world; k++) *reinterpret_cast<float4*>(gsum_of(c, (c.rank + k) % c.world, par) + i) = s;
        }
        __syncthreads();
        if (tid == 0) {
            __threadfence_system();
            for (int k = 1; k <= c.world; k++) *reinterpret_cast<volatile uint32_t*>(flag2_of(c, (c.rank + k) % c.world, b)) = ep;
        }
    }
    if (phase == 1) return;
    // ---- every rank: the owner's sum has landed here -> optimizer on the chunk
    if (tid == 0) {
        const uint32_t *f = flag2_of(c, c.rank, b);
        const long long t0 = clock64();
        while ((int32_t)(ld_acquire_sys(f) - ep) < 0) { __nanosleep(32); if (clock64() - t0 > c.spin_limit) { timed_out(own); break; } }
    }
    __syncthreads();
    if (s_err) return;
    const float *sum = gsum_of(c, c.rank, par);
    for (int64_t i = lo + 4 * tid; i < hi; i += 4 * T4K_THREADS) {
        float4 s = __ldcg(reinterpret_cast<const float4*>(sum + i));
        float4 g = *reinterpret_cast<const float4*>(o.G + i), m = make_float4(0, 0, 0, 0), v = m;
        if (KIND == 0) {
            int l = 0, h = o.nseg - 1;
            while (l < h) { int mid = (l + h + 1) >> 1; if (o.seg[mid].off <= i) l = mid; else h = mid - 1; }
            const float nw = (float)o.seg[l].Nw;
            s.x = s.x / nw; s.y = s.y / nw; s.z = s.z / nw; s.w = s.w / nw;
            if (o.mom) m = *reinterpret_cast<const float4*>(o.M + i);
        } else { m = *reinterpret_cast<const float4*>(o.M + i); v = *reinterpret_cast<const float4*>(o.V + i); }
        opt_step<KIND>(g.x, s.x, m.x, v.x, 1.0f, o.mom, o.p); opt_step<KIND>(g.y, s.y, m.y, v.y, 1.0f, o.mom, o.p);
        opt_step<KIND>(g.z, s.z, m.z, v.z, 1.0f, o.mom, o.p); opt_step<KIND>(g.w, s.w, m.w, v.w, 1.0f, o.mom, o.p);
        *reinterpret_cast<float4*>(o.G + i) = g;
        *reinterpret_cast<float4*>(buf + i) = make_float4(0, 0, 0, 0);
        if (KIND == 0) { if (o.mom) *reinterpret_cast<float4*>(o.M + i) = m; }
        else { *reinterpret_cast<float4*>(o.M + i) = m; *reinterpret_cast<float4*>(o.V + i) = v; }
    }
    if (tid == 0) c.epoch[b] = ep;
}

// signal half of a push whose DATA travelled by copy engine (t4k_dp_push_dma): the chunks' epoch flags, stored to every rank (this one
// included) once the copies of this stream have completed.  `par` is the slot parity the host addressed the copies with: it must be the
// parity of the epoch the chunks are about to complete — a disagreement (an exchange the host did not count) raises the sticky error.
__global__ void __launch_bounds__(T4K_THREADS) k_dp_signal(const __grid_constant__ CommDev c, int b0, int b1, int par) {
    pdl_wait(); pdl_trigger();
    if (c.epoch[COMM_MAXB]) return;
    __threadfence_system();
    for (int b = b0 + threadIdx.x; b < b1; b += blockDim.x) {
        const uint32_t ep = c.epoch[b] + 1;
        if ((int)(ep & 1u) != par) {
            atomicCAS(c.epoch + COMM_MAXB, 0u, 1u + (uint32_t)c.rank);
            if (c.err_host) { *reinterpret_cast<volatile uint32_t*>(c.err_host) = 1u + (uint32_t)c.rank; __threadfence_system(); }
            continue;
        }
        for (int k = 1; k <= c.world; k++) *reinterpret_cast<volatile uint32_t*>(flag_of(c, (c.rank + k) % c.world, c.rank, b)) = ep;
    }
}

static CommDev devview(const t4k_comm *c) {
    CommDev d;
    for (int i = 0; i < COMM_MAXW; i++) d.peer[i] = c->peer[i];
    d.epoch = c->epoch; d.err_host = c->err_dev; d.spin_limit = c->spin_limit; d.cap = c->cap; d.rank = c->rank; d.world = c->world; d.ch4 = c->ch4; d.scal_mirror = c->scal_mirror;
    return d;
}
// An SM changes its L1 / shared-memory split only when it is empty.  Exchange kernels WAIT — resident on every SM — and would pin the split
// they were launched with (they use no shared memory: the smallest) for as long as they wait: a kernel of another stream that needs a larger
// carve-out (the fused conv blocks, the layer GEMM) could not start next to them.  One GPU per rank never runs anything next to its own
// exchange; several ranks of one process on one device (the single-GPU data-parallel tests) do.  The exchange reads through L2 (__ldcg,
// peer stores), so asking for the largest shared-memory split costs it nothing.
static void exchange_carveout() {
    static bool done[16];
    const int dev = cur_device();
    if (dev < 0 || dev >= 16 || done[dev]) return;
    const void *k[] = {(const void*)k_dp_push_owner, (const void*)k_dp_exchange_rs<0>, (const void*)k_dp_exchange_rs<1>, (const void*)k_dp_exchange_rs<2>, (const void*)k_dp_signal, (const void*)k_dp_exchange<-1, true>, (const void*)k_dp_exchange<0, true>, (const void*)k_dp_exchange<0, false>,
                       (const void*)k_dp_exchange<1, true>, (const void*)k_dp_exchange<2, true>, (const void*)k_dp_exchange<3, true>};
    for (const void *f : k) if (cudaFuncSetAttribute(f, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared) != cudaSuccess) cudaGetLastError();
    done[dev] = true;
}
static bool ready(const t4k_comm *c) {
    if (!c || !c->base) return false;
    for (int i = 0; i < c->world; i++) if (!c->peer[i]) return false;
    return true;
}

} // namespace t4k
using namespace t4k;

extern "C" {

int t4k_comm_create(int rank, int world, int64_t cap_floats, t4k_comm_t *out, void *handle64) {
    if (!out || world < 1 || world > COMM_MAXW || rank < 0 || rank >= world || cap_floats < 1) return T4K_EINVAL;
    t4k_comm *c = new (std::nothrow) t4k_comm();
    if (!c) return T4K_ENOMEM;
    memset(c, 0, sizeof(*c));
    c->rank = rank; c->world = world;
    c->cap = (cap_floats + 3) & ~(int64_t)3;
    // chunking: a multiple of 32 float4s per chunk, about 2 chunks per SM, at most COMM_MAXB.  Peer-store throughput per SM is
    // limited, so many small chunks win (0.79 MB, 2 GPUs: 296 chunks 9.9 us, 64: 18.6, 16: 53); T4K_COMM_CHUNKS overrides for tuning
    const int64_t cap4 = c->cap / 4;
    int64_t nb = cap4 / 256; int64_t want = 2 * (int64_t)sm_count();
    if (const char *e = getenv("T4K_COMM_CHUNKS")) want = atoi(e) > 0 ? atoi(e) : want;
    if (nb > want) nb = want; if (nb < 1) nb = 1; if (nb > COMM_MAXB) nb = COMM_MAXB;
    int64_t ch4 = (cap4 + nb - 1) / nb; ch4 = (ch4 + 31) & ~(int64_t)31;
    while ((cap4 + ch4 - 1) / ch4 > COMM_MAXB) ch4 += 32;
    c->ch4 = (int)ch4;
    if (cudaGetDevice(&c->dev) != cudaSuccess) { cudaGetLastError(); delete c; return T4K_EINVAL; }
    exchange_carveout();
    c->bytes = (size_t)COMM_FLAGB + (size_t)2 * world * (size_t)(c->cap + COMM_NSCAL) * 4 + (size_t)2 * c->cap * 4;
    cudaError_t e = cudaMalloc((void**)&c->base, c->bytes);
    if (e != cudaSuccess) { cudaGetLastError(); delete c; return T4K_ENOMEM; }
    e = cudaMalloc((void**)&c->epoch, (COMM_MAXB + 8) * 4);
    if (e != cudaSuccess) { cudaGetLastError(); cudaFree(c->base); delete c; return T4K_ENOMEM; }
    cudaMemset(c->base, 0, c->bytes);
    cudaMemset(c->epoch, 0, (COMM_MAXB + 8) * 4);
    if (cudaHostAlloc((void**)&c->err_host, 64, cudaHostAllocMapped) == cudaSuccess && c->err_host) {
        *c->err_host = 0;
        if (cudaHostGetDevicePointer((void**)&c->err_dev, c->err_host, 0) != cudaSuccess) { cudaGetLastError(); c->err_dev = nullptr; }
    } else { cudaGetLastError(); c->err_host = c->err_dev = nullptr; }
    {
        int khz = 0; double secs = COMM_TIMEOUT_S_DEFAULT;
        if (const char *e = getenv("T4K_COMM_TIMEOUT_S")) { const double v = atof(e); if (v > 0) secs = v; }
        if (cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, c->dev) != cudaSuccess || khz <= 0) { cudaGetLastError(); khz = 2000000; }
        c->spin_limit = (long long)(secs * 1e3 * (double)khz);
    }
    cudaDeviceSynchronize();
    c->peer[rank] = c->base;
    if (handle64) {
        cudaIpcMemHandle_t h;
        static_assert(sizeof(h) <= T4K_COMM_HANDLE_BYTES, "handle size");
        memset(handle64, 0, T4K_COMM_HANDLE_BYTES);
        e = cudaIpcGetMemHandle(&h, c->base);
        if (e != cudaSuccess) { cudaGetLastError(); if (world > 1) { cudaFree(c->base); cudaFree(c->epoch); delete c; return (int)e; } }
        else memcpy(handle64, &h, sizeof(h));
    }
    *out = c;
    return 0;
}

int t4k_comm_connect(t4k_comm_t c, const void *handles) {
    if (!c || (!handles && c->world > 1)) return T4K_EINVAL;
    for (int r = 0; r < c->world; r++) {
        if (r == c->rank || c->peer[r]) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char*)handles + (size_t)r * T4K_COMM_HANDLE_BYTES, sizeof(h));
        void *p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) { cudaGetLastError(); return (int)e; }
        c->peer[r] = (char*)p; c->ipc[r] = true;
    }
    return 0;
}

/* same-process wiring (tests; several ranks of one process, on one or several devices with peer access enabled) */
int t4k_comm_connect_local(t4k_comm_t c, t4k_comm_t *all) {
    if (!c || !all) return T4K_EINVAL;
    for (int r = 0; r < c->world; r++) {
        if (!all[r] || all[r]->world != c->world || all[r]->rank != r || all[r]->cap != c->cap) return T4K_EINVAL;
        c->peer[r] = all[r]->base;
    }
    return 0;
}

int t4k_comm_destroy(t4k_comm_t c) {
    if (!c) return 0;
    cudaDeviceSynchronize();
    for (int r = 0; r < c->world; r++) if (c->ipc[r] && c->peer[r]) cudaIpcCloseMemHandle(c->peer[r]);
    if (c->base) cudaFree(c->base);
    if (c->epoch) cudaFree(c->epoch);
    if (c->err_host) cudaFreeHost(c->err_host);
    if (c->cfork) {
        cudaEventDestroy(c->cfork);
        for (int p = 0; p < c->world; p++) { if (c->cs[p]) cudaStreamDestroy(c->cs[p]); if (c->cdone[p]) cudaEventDestroy(c->cdone[p]); }
    }
    cudaGetLastError();
    delete c;
    return 0;
}

/* 0 = healthy; k>0 = a wait for rank k-1 timed out (synchronises the device) */
int t4k_comm_status(t4k_comm_t c) {
    if (!c) return T4K_EINVAL;
    uint32_t w = 0;
    cudaError_t e = cudaMemcpy(&w, c->epoch + COMM_MAXB, 4, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return (int)e;
    return (int)w;
}

/* the sticky error word without synchronising anything (mapped host memory written by the kernel that timed out) */
int t4k_comm_poll(t4k_comm_t c) {
    if (!c) return T4K_EINVAL;
    return c->err_host ? (int)*reinterpret_cast<volatile uint32_t*>(c->err_host) : 0;
}

int64_t t4k_comm_capacity(t4k_comm_t c) { return c ? c->cap : 0; }

/* the exchanges launched (or captured) from now on also store the summed scalars into `pinned` — page-locked host memory, reachable from the device
 * under unified addressing — so that a training loop reads the global loss on the host without a copy node behind the step (NULL: off) */
int t4k_comm_scalar_mirror(t4k_comm_t c, float *pinned) { if (!c) return T4K_EINVAL; c->scal_mirror = pinned; return 0; }

/* samples [lo, hi) of a batch of n owned by `rank` of `world`: contiguous, sizes differ by at most one, first ranks larger (host-only) */
int t4k_shard_info(int64_t n, int world, int rank, int64_t *lo, int64_t *hi) {
    if (n < 0 || world < 1 || rank < 0 || rank >= world || !lo || !hi) return T4K_EINVAL;
    const int64_t q = n / world, r = n % world;
    *lo = rank * q + (rank < r ? rank : r);
    *hi = *lo + q + (rank < r ? 1 : 0);
    return 0;
}

int t4k_allreduce_sum(t4k_comm_t c, float *buf, int64_t n, t4k_stream_t s) {
    if (!ready(c) || !buf || n < 0 || n > c->cap) return T4K_EINVAL;
    if (n == 0) return 0;
    const int grid = (int)((((n + 3) / 4) + c->ch4 - 1) / c->ch4);
    DpOpt o{}; o.pushed_from = c->cap + 1;
    if ((n & 3) == 0 && aligned16(buf)) launch_pdl(k_dp_exchange<0, true>, dim3(grid), dim3(T4K_THREADS), 0, STRM(s), devview(c), buf, n, nullptr, 0, o);
    else                                launch_pdl(k_dp_exchange<0, false>, dim3(grid), dim3(T4K_THREADS), 0, STRM(s), devview(c), buf, n, nullptr, 0, o);
    return check_launch();
}

int64_t t4k_dp_push(t4k_comm_t c, const float *DG, int64_t from, int64_t total, t4k_stream_t s) {
    if (!ready(c) || !DG || from < 0 || total < 0 || total > c->cap || (total & 3) || !aligned16(DG)) return T4K_EINVAL;
    const int64_t chf = (int64_t)c->ch4 * 4;
    const int64_t b0 = (from + chf - 1) / chf, nb = (total + chf - 1) / chf;
    if (b0 >= nb) return total;                                   // nothing starts at or beyond `from`
    DpOpt o{}; o.b0 = (int)b0; o.pushed_from = 0;
    launch_pdl(k_dp_exchange<-1, true>, dim3((unsigned)(nb - b0)), dim3(T4K_THREADS), 0, STRM(s), devview(c), const_cast<float*>(DG), total, nullptr, 0, o);
    const int rc = check_launch();
    return rc ? (rc > 0 ? -(int64_t)rc - 1000 : rc) : b0 * chf;
}

int64_t t4k_comm_chunk_floats(t4k_comm_t c) { return c ? (int64_t)c->ch4 * 4 : 0; }

static int optim_dp_launch(t4k_comm_t c, int kind, float *G, float *DG, float *M, float *V, const t4k_seg_t *seg, int nseg,
                           int64_t total, float lr, float b1, float b2, float wd, float *scal, int nscal, int64_t pushed_from, int b0, int grid, t4k_stream_t s);


/* reduce-scatter flavour of the early push (world > 2): every chunk that STARTS at or beyond `from` goes to its owner (rank chunk % world) only.
 * Must be followed, for the same chunks, by t4k_optim_multi_dp_rs — not by the all-to-all exchange.  Returns the first pushed float offset. */
int64_t t4k_dp_push_owner(t4k_comm_t c, const float *DG, int64_t from, int64_t total, t4k_stream_t s) {
    if (!ready(c) || !DG || from < 0 || total < 0 || total > c->cap || (total & 3) || !aligned16(DG)) return T4K_EINVAL;
    const int64_t chf = (int64_t)c->ch4 * 4;
    const int64_t b0 = (from + chf - 1) / chf, nb = (total + chf - 1) / chf;
    if (b0 >= nb) return total;
    launch_pdl(k_dp_push_owner, dim3((unsigned)(nb - b0)), dim3(T4K_THREADS), 0, STRM(s), devview(c), DG, total, (int)b0);
    const int rc = check_launch();
    return rc ? (rc > 0 ? -(int64_t)rc - 1000 : rc) : b0 * chf;
}
/* the owners sum, send the sums to every rank, every rank runs the optimizer: the chunks that start in [from, total), all pushed by t4k_dp_push_owner.
 * phase 0: one launch; phase 1 then phase 2: the owners' half (a few blocks per rank: may run next to backprop) and the optimizer half apart */
int t4k_optim_multi_dp_rs(t4k_comm_t c, int kind, float *G, float *DG, float *M, float *V, const t4k_seg_t *seg, int nseg,
                          int64_t from, int64_t total, float lr, float b1, float b2, float wd, int phase, t4k_stream_t s) {
    if (phase < 0 || phase > 2) return T4K_EINVAL;
    if (!ready(c) || !G || !DG || !seg || nseg < 1 || total < 0 || total > c->cap || (total & 3) || from < 0 || from > total || !aligned16(DG) || !aligned16(G)) return T4K_EINVAL;
    const int64_t chf = (int64_t)c->ch4 * 4;
    const int64_t b0 = (from + chf - 1) / chf, nb = (total + chf - 1) / chf;
    if (b0 >= nb) return 0;
    DpOpt o{G, M, V, seg, nseg, true, OptP{lr, b1, b2, wd}, (int)b0, 0};
    CommDev d = devview(c);
    const dim3 grid(phase == 1 ? (unsigned)((nb - b0 + c->world - 1) / c->world) : (unsigned)(nb - b0));
    switch (kind) {
    case 0: o.mom = !(fabsf(b1) < DU_EPS); if (o.mom && !M) return T4K_EINVAL;
            launch_pdl(k_dp_exchange_rs<0>, grid, dim3(T4K_THREADS), 0, STRM(s), d, DG, total, o, phase); break;
    case 1: if (!M || !V || !aligned16(M) || !aligned16(V)) return T4K_EINVAL;
            launch_pdl(k_dp_exchange_rs<1>, grid, dim3(T4K_THREADS), 0, STRM(s), d, DG, total, o, phase); break;
    case 2: if (!M || !V || !aligned16(M) || !aligned16(V)) return T4K_EINVAL;
            launch_pdl(k_dp_exchange_rs<2>, grid, dim3(T4K_THREADS), 0, STRM(s), d, DG, total, o, phase); break;
    default: return T4K_EINVAL;
    }
    return check_launch();
}

/* t4k_dp_push with the data moved by the COPY ENGINES (one peer-to-peer cudaMemcpyAsync per rank, own slot included) instead of SM stores: the
 * push shares the machine with the rest of backprop without taking an SM from it; a one-block kernel then raises the chunks' flags.  Copy nodes
 * carry fixed addresses, so the caller states the slot parity: `step` = number of exchanges this communicator has COMPLETED over these chunks
 * (every exchange of a model's arena covers all of them, so the model counts its optimizer calls); a captured step is captured once per parity. */
int64_t t4k_dp_push_dma(t4k_comm_t c, const float *DG, int64_t from, int64_t total, uint32_t step, t4k_stream_t s) {
    if (!ready(c) || !DG || from < 0 || total < 0 || total > c->cap || (total & 3) || !aligned16(DG)) return T4K_EINVAL;
    const int64_t chf = (int64_t)c->ch4 * 4;
    const int64_t b0 = (from + chf - 1) / chf, nb = (total + chf - 1) / chf;
    if (b0 >= nb) return total;
    const int par = (int)((step + 1u) & 1u);
    const int64_t off = b0 * chf, sstride = c->cap + COMM_NSCAL;
    // more than one remote destination: one copy stream per destination, forked off `s` and joined back in front of the signal kernel (under
    // stream capture: parallel copy nodes) — seven 0.8 MB copies in a row on ONE stream outlast the backward kernels they hide under
    // (8 GPUs: 110.7 us per step against 96.1 with the push kernel)
    const bool fan = c->world > 2;
    if (fan && !c->cfork) {
        if (cudaEventCreateWithFlags(&c->cfork, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return T4K_ENOMEM; }
        for (int p = 0; p < c->world; p++)
            if (cudaStreamCreateWithFlags(&c->cs[p], cudaStreamNonBlocking) != cudaSuccess || cudaEventCreateWithFlags(&c->cdone[p], cudaEventDisableTiming) != cudaSuccess) {
                cudaGetLastError(); return T4K_ENOMEM;
            }
    }
    if (fan) cudaEventRecord(c->cfork, STRM(s));
    for (int k = 1; k <= c->world; k++) {
        const int p = (c->rank + k) % c->world;
        float *dst = reinterpret_cast<float*>(c->peer[p] + COMM_FLAGB) + (int64_t)(par * c->world + c->rank) * sstride + off;
        cudaStream_t cst = fan ? c->cs[p] : STRM(s);
        if (fan) cudaStreamWaitEvent(cst, c->cfork, 0);
        cudaError_t e = cudaMemcpyAsync(dst, DG + off, (size_t)(total - off) * sizeof(float), cudaMemcpyDeviceToDevice, cst);
        if (e != cudaSuccess) { cudaGetLastError(); return -(int64_t)e - 1000; }
        if (fan) { cudaEventRecord(c->cdone[p], cst); cudaStreamWaitEvent(STRM(s), c->cdone[p], 0); }
    }
    launch_pdl(k_dp_signal, dim3(1), dim3(T4K_THREADS), 0, STRM(s), devview(c), (int)b0, (int)nb, par);
    const int rc = check_launch();
    return rc ? (rc > 0 ? -(int64_t)rc - 1000 : rc) : off;
}

int t4k_optim_multi_dp(t4k_comm_t c, int kind, float *G, float *DG, float *M, float *V, const t4k_seg_t *seg, int nseg,
                       int64_t total, float lr, float b1, float b2, float wd, float *scal, int nscal, int64_t pushed_from, t4k_stream_t s) {
    if (!ready(c) || total < 0 || total > c->cap) return T4K_EINVAL;
    if (total == 0) return 0;
    return optim_dp_launch(c, kind, G, DG, M, V, seg, nseg, total, lr, b1, b2, wd, scal, nscal, pushed_from, 0, (int)(((total / 4) + c->ch4 - 1) / c->ch4), s);
}

/* the same exchange + optimizer on the chunks that START in [from, to) only (chunk = t4k_comm_chunk_floats floats, chunk k starts at k * that).
 * A step may issue it twice on disjoint ranges — the part of the arena whose gradients are final early on a side stream, under the rest of
 * backprop, and the first layers' chunks at the end — instead of one launch over everything: chunks carry their own epochs.  The scalars ride
 * with chunk 0.  Every rank must split at the same offsets.  pushed_from as in t4k_optim_multi_dp (chunks at or beyond it were pushed early). */
int t4k_optim_multi_dp_range(t4k_comm_t c, int kind, float *G, float *DG, float *M, float *V, const t4k_seg_t *seg, int nseg,
                             int64_t from, int64_t to, int64_t total, float lr, float b1, float b2, float wd, float *scal, int nscal,
                             int64_t pushed_from, t4k_stream_t s) {
    if (!ready(c) || total < 0 || total > c->cap || from < 0 || to < from || to > total) return T4K_EINVAL;
    const int64_t chf = (int64_t)c->ch4 * 4;
    const int64_t nb = (total + chf - 1) / chf;
    int64_t c0 = (from + chf - 1) / chf, c1 = (to >= total) ? nb : (to + chf - 1) / chf;
    if (c1 > nb) c1 = nb;
    if (c0 >= c1) return 0;
    return optim_dp_launch(c, kind, G, DG, M, V, seg, nseg, total, lr, b1, b2, wd, c0 == 0 ? scal : nullptr, c0 == 0 ? nscal : 0, pushed_from, (int)c0, (int)(c1 - c0), s);
}

static int optim_dp_launch(t4k_comm_t c, int kind, float *G, float *DG, float *M, float *V, const t4k_seg_t *seg, int nseg,
                           int64_t total, float lr, float b1, float b2, float wd, float *scal, int nscal, int64_t pushed_from, int b0, int grid, t4k_stream_t s) {
    if (!G || !DG || !seg || nseg < 1 || (total & 3) || !aligned16(DG) || !aligned16(G) ||
        nscal < 0 || nscal > COMM_NSCAL || (nscal && !scal)) return T4K_EINVAL;
    DpOpt o{G, M, V, seg, nseg, true, OptP{lr, b1, b2, wd}, b0, (pushed_from > 0 && pushed_from <= total) ? pushed_from : total + 1};
    CommDev d = devview(c);
    switch (kind) {
    case 0: o.mom = !(fabsf(b1) < DU_EPS); if (o.mom && !M) return T4K_EINVAL;
            launch_pdl(k_dp_exchange<1, true>, dim3(grid), dim3(T4K_THREADS), 0, STRM(s), d, DG, total, scal, nscal, o); break;
    case 1: if (!M || !V || !aligned16(M) || !aligned16(V)) return T4K_EINVAL;
            launch_pdl(k_dp_exchange<2, true>, dim3(grid), dim3(T4K_THREADS), 0, STRM(s), d, DG, total, scal, nscal, o); break;
    case 2: if (!M || !V || !aligned16(M) || !aligned16(V)) return T4K_EINVAL;
            launch_pdl(k_dp_exchange<3, true>, dim3(grid), dim3(T4K_THREADS), 0, STRM(s), d, DG, total, scal, nscal, o); break;
    default: return T4K_EINVAL;
    }
    return check_launch();
}

} // extern "C"
