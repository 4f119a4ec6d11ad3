// t4host.cpp — bodies of t4::Tensor / t4::Model on top of the kernel C-ABI (include/t4k.h), plus the
// flat C interface of include/t4host.h.  Reference citations are file:line in the reference tree.
#include "t4host.hpp"
#include "../../../include/t4host.h"
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <fstream>
#include <sstream>
#include <string>
#include <cstdarg>
#include <cstring>
#include <cfloat>

namespace t4 {

// =============================================================================== Runtime
// per host thread: the lane the thread works on (every thread starts on lane 0, see Runtime::use_lane)
static thread_local cudaStream_t g_stream = nullptr, g_stream2 = nullptr;      // library stream + side stream (forked work inside a step)
static thread_local cudaEvent_t  g_fork = nullptr, g_join = nullptr, g_head = nullptr, g_mid = nullptr, g_push = nullptr;
static bool  g_init = false;
static int   g_device = 0;
// Lanes: independent (stream, side stream, events, workspace banks) sets of ONE process.  Lane 0 is the default and the only one a
// training process uses; the data-parallel tests run `world` ranks of one process on one device, each on its own lane, so that the ranks'
// kernels can wait on one another exactly as they do across GPUs (Runtime::use_lane).
#define T4_MAX_LANES 4
static struct Lane { cudaStream_t s = nullptr, s2 = nullptr; cudaEvent_t fork = nullptr, join = nullptr, head = nullptr, mid = nullptr, push = nullptr; } g_lanes[T4_MAX_LANES];
static thread_local int g_lane = 0;
#define WS_BANK_MAIN (2 * g_lane)
#define WS_BANK_SIDE (2 * g_lane + 1)
static char  g_err[512] = "";

int Runtime::init(int device) {
    if (g_init) return 0;
    if (cudaSetDevice(device) != cudaSuccess) { error("cudaSetDevice(%d) failed: no CUDA device (there is no CPU fallback)", device); cudaGetLastError(); return T4K_EINVAL; }
    if (cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking) != cudaSuccess) { error("cudaStreamCreate failed"); return T4K_EINVAL; }
    if (cudaStreamCreateWithFlags(&g_stream2, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreateWithFlags(&g_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&g_join, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&g_head, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&g_mid, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&g_push, cudaEventDisableTiming) != cudaSuccess) { error("cudaStreamCreate failed"); return T4K_EINVAL; }
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t keep = UINT64_MAX;                     // keep freed blocks cached: alloc/free in `for @ drop next` loops stay cheap
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    g_lanes[0] = Lane{g_stream, g_stream2, g_fork, g_join, g_head, g_mid, g_push};
    g_device = device;
    g_init = true;
    return 0;
}
int Runtime::use_lane(int k) {
    if (k < 0 || k >= T4_MAX_LANES || (!g_init && init(0))) return T4K_EINVAL;
    Lane &l = g_lanes[k];
    cudaSetDevice(g_device);                              // the current device is per host thread as well
    if (!l.s) {
        if (cudaStreamCreateWithFlags(&l.s, cudaStreamNonBlocking) != cudaSuccess || cudaStreamCreateWithFlags(&l.s2, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&l.fork, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&l.join, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&l.head, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&l.mid, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&l.push, cudaEventDisableTiming) != cudaSuccess) {
            cudaGetLastError(); error("lane %d: stream creation failed", k); return T4K_EINVAL;
        }
    }
    g_lane = k; g_stream = l.s; g_stream2 = l.s2; g_fork = l.fork; g_join = l.join; g_head = l.head; g_mid = l.mid; g_push = l.push;
    t4k_set_workspace_bank(WS_BANK_MAIN);
    return 0;
}
void *Runtime::stream() {
    if (!g_stream && g_init) use_lane(0);                 // a thread other than the one that initialised the runtime: lane 0 until it picks another
    return (void*)g_stream;
}
int   Runtime::sync()   { return (int)cudaStreamSynchronize((cudaStream_t)stream()); }
void  Runtime::error(const char *fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
}
const char *Runtime::last_error() { return g_err; }
void *Runtime::alloc(size_t bytes) {
    if (!g_init && init(0)) return nullptr;
    void *p = nullptr;
    bytes = (bytes + 255) & ~(size_t)255;
    if (cudaMallocAsync(&p, bytes, (cudaStream_t)stream()) != cudaSuccess) { cudaGetLastError(); error("device allocation of %zu bytes failed", bytes); return nullptr; }
    return p;
}
void Runtime::free(void *p) { if (p) cudaFreeAsync(p, (cudaStream_t)stream()); }

#define ST         (Runtime::stream())
#define KCHK(call) do { int _rc = (call); if (_rc) Runtime::error("%s -> %d (%s)", #call, _rc, t4k_strerror(_rc)); } while (0)
// how the early half of the data-parallel exchange travels (T4K_DP_EARLY): "dma" copy engines + the exchange/optimizer of the rest of the arena
// under the first layer's finish launch, "sm" the push kernel of comm.cu, "range" the whole exchange + optimizer of the finished part at once
// (measured slower: its waiting CTAs sit on SMs the conv block's backward needs).  Default by world size, from the measured steps (MNIST CNN,
// N=512 per GPU, us per step dma / sm): 2 GPUs 79.5 / 82.0, 8 GPUs 111.1 / 96.1 — seven peer-to-peer copies per rank, even on seven streams,
// outlast the 25 us of backward they hide under, while the push kernel spreads them over the SMs' store paths.
static int opt_late_on() {
    static int v = -1;
    if (v < 0) { const char *e = getenv("T4K_OPT_LATE"); v = (e && e[0] == '1') ? 1 : 0; }
    return v;
}
static int dp_mirror_on() {                                // T4K_DP_MIRROR=0: the loss read-back of a data-parallel step is a copy node behind the exchange, as before
    static int v = -1;
    if (v < 0) { const char *e = getenv("T4K_DP_MIRROR"); v = (e && e[0] == '0') ? 0 : 1; }
    return v;
}
static int g_dp_rest = -1;
static int dp_rest_on() {                                  // T4K_DP_REST=0: the end of the step exchanges the whole arena in one launch, as in round 1
    int &v = g_dp_rest;
    if (v < 0) { const char *e = getenv("T4K_DP_REST"); v = (e && e[0] == '0') ? 0 : 1; }
    return v;
}
static int dp_rs_on() {                                    // T4K_DP_RS=0: all-to-all pushes at every world size
    static int v = -1;
    if (v < 0) { const char *e = getenv("T4K_DP_RS"); v = (e && e[0] == '0') ? 0 : 1; }
    return v;
}
static int dp_rs_phased() {                                // T4K_DP_RS_PHASE=1: the owners' half of the reduce-scatter exchange right behind the push (two launches)
    static int v = -1;
    if (v < 0) { const char *e = getenv("T4K_DP_RS_PHASE"); v = (e && e[0] == '1') ? 1 : 0; }
    return v;
}
static int g_dp_early = -1;
static int dp_early_mode(int world = 2) {
    int &m = g_dp_early;
    if (m < 0) { const char *e = getenv("T4K_DP_EARLY"); m = !e ? 3 : (!strcmp(e, "sm") ? 0 : (!strcmp(e, "range") ? 2 : (!strcmp(e, "dma") ? 1 : 3))); }
    return m == 3 ? (world <= 2 ? 1 : 0) : m;
}
static inline DU SCALAR(DU v) { uint32_t u; memcpy(&u, &v, 4); u &= ~1u; memcpy(&v, &u, 4); return v; }   // src/t4base.h:33 (object tag bit cleared)

// =============================================================================== Tensor
Tensor &Tensor::reset(void *mem, U64 sz, t4_layer fn) {                 // tensor.cu:461-484
    numel = sz; rank = 1; train = false; err = false; iparm = 0; xparm = 0;
    const U64 GB = 1ull << 30;
    const U16 s[4] = {1, 1, 1, 1};
    const U32 h[4] = {(U32)(sz > GB ? (sz >> 30) : sz), (U32)(sz > GB ? GB : 1), 1, 1};
    data = (DU*)mem; grad_fn = fn;
    memcpy(stride, s, sizeof(s)); memcpy(shape, h, sizeof(h));
    for (int i = 0; i < 5; i++) grad[i] = mtum[i] = nullptr;
    _tmp = data ? &data[numel] : nullptr;
    return *this;
}
static Tensor &alloc_tensor(U64 sz) {
    Tensor *t = new Tensor();
    void *mem = Runtime::alloc((sz + 4) * sizeof(DU));                   // numel + scratch (MMU::talloc: numel+1, mmu.cu:202)
    t->reset(mem, sz);
    return *t;
}
Tensor &Tensor::create(U64 sz)                    { return alloc_tensor(sz); }
Tensor &Tensor::create(U32 h, U32 w)              { return alloc_tensor((U64)h * w).reshape(h, w); }
Tensor &Tensor::create(U32 n, U32 h, U32 w, U32 c){ return alloc_tensor((U64)n * h * w * c).reshape(n, h, w, c); }
Tensor &Tensor::create_like(Tensor &t)            { return create(t.N(), t.H(), t.W(), t.C()); }
Tensor &Tensor::copy_of(Tensor &t) {                                     // MMU::copy, mmu.cu:274-295
    Tensor &o = alloc_tensor(t.numel);
    o.rank = t.rank; memcpy(o.shape, t.shape, sizeof(o.shape)); memcpy(o.stride, t.stride, sizeof(o.stride));
    o.xparm = t.xparm; o.iparm = t.iparm;
    copy(t, o);
    return o;
}
void Tensor::destroy(Tensor &t) {
    if (t.owns) Runtime::free(t.data);
    delete &t;
}
bool Tensor::is_same_shape(Tensor &t) { return memcmp(shape, t.shape, sizeof(shape)) == 0; }

Tensor &Tensor::reshape(U64 sz) {                                        // tensor.cu:487-497
    if (sz == numel) { DU *d = data; t4_layer fn = grad_fn; Tensor *g[5], *m[5]; memcpy(g, grad, sizeof(g)); memcpy(m, mtum, sizeof(m));
                       reset(d, numel, fn); memcpy(grad, g, sizeof(g)); memcpy(mtum, m, sizeof(m)); }
    else Runtime::error("  tensor#reshape sz != numel (%ld != %ld)\n", (long)sz, (long)numel);
    return *this;
}
Tensor &Tensor::reshape(U32 h, U32 w) {                                  // tensor.cu:499-514
    if ((U64)h * w == numel) { rank = 2; const U16 s[4] = {1, 1, 1, 1}; const U32 t[4] = {h, w, 1, 1}; memcpy(stride, s, sizeof(s)); memcpy(shape, t, sizeof(t)); }
    else Runtime::error("  tensor#reshape sz != numel (%ld != %ld)\n", (long)((U64)h * w), (long)numel);
    return *this;
}
Tensor &Tensor::reshape(U32 n, U32 h, U32 w, U32 c) {                    // tensor.cu:516-531
    if ((U64)n * h * w * c == numel) { rank = 4; const U16 s[4] = {1, 1, 1, 1}; const U32 t[4] = {h, w, c, n}; memcpy(stride, s, sizeof(s)); memcpy(shape, t, sizeof(t)); }
    else Runtime::error("  tensor#reshape sz != numel (%ld != %ld)\n", (long)((U64)n * h * w * c), (long)numel);
    return *this;
}
Tensor &Tensor::identity() { KCHK(t4k_identity(data, N(), H(), W(), C(), ST)); return *this; }      // tensor.cu:548-555
Tensor &Tensor::zeros()    { cudaMemsetAsync(data, 0, sizeof(DU) * numel, (cudaStream_t)ST); return *this; }   // tensor.cu:557-562
Tensor &Tensor::map(math_op op, DU v) { KCHK(t4k_map(op, data, v, numel, ST)); return *this; }       // tensor.cu:564-571
Tensor &Tensor::normalize(DU avg, DU std) {                              // tensor.cu:573-578
    KCHK(t4k_ts_op(T4K_SUB, data, avg, data, numel, ST)); KCHK(t4k_ts_op(T4K_DIV, data, std, data, numel, ST)); return *this;
}
Tensor &Tensor::ten_op(math_op op, Tensor &A, DU v, Tensor &O) {         // tensor.cu:17-23
    KCHK(t4k_ts_op(op, A.data, v, O.data, A.numel, ST)); return O;
}
Tensor &Tensor::ten_op(math_op op, Tensor &A, Tensor &B, Tensor &O) {    // tensor.cu:29-53
    const U32 Na = A.N(), Nb = B.N();
    if (A.HWC() != B.HWC() || (Na == 1 ? B.numel : A.numel) != O.numel) {
        Runtime::error("  tensor#ten_op A.HWC(%ld)!=B.HWC(%ld) or N, C diff\n", (long)A.HWC(), (long)B.HWC());
        return O;
    }
    KCHK(t4k_tt_op(op, A.data, B.data, O.data, A.HWC(), Na, Nb, ST));
    return O;
}
Tensor &Tensor::dot(Tensor &A, Tensor &B, Tensor &O, DU alpha, DU beta) { // tensor.cu:61-72
    KCHK(t4k_dot(A.data, B.data, O.data, alpha, beta, A.W(), A.C(), A.N(), B.N(), ST)); return O;
}
Tensor &Tensor::mm(Tensor &A, Tensor &B, Tensor &O, bool inc, bool tA, bool tB) {   // tensor.cu:74-77
    return gemm3(A, B, O, 1.0f, inc ? 1.0f : 0.0f, tA, tB);
}
Tensor &Tensor::linear(Tensor &A, Tensor &B, Tensor &O, int H, int W, int K, DU alpha, DU beta, bool tA, bool tB) { // tensor.cu:80-87
    KCHK(t4k_gemm(A.data, B.data, O.data, alpha, beta, tA, tB, H, W, K, 1, 1, 0, 0, 0, ST)); return O;
}
Tensor &Tensor::_gemm(int engine, Tensor &A, Tensor &B, Tensor &O, DU alpha, DU beta, bool tA, bool tB, const char *nm) {
    U32 H  = tA ? A.W() : A.H(), W  = tB ? B.H() : B.W();               // tensor.cu:162-180
    U32 Ka = tA ? A.H() : A.W(), Kb = tB ? B.W() : B.H();
    U32 Na = A.N(), Nb = B.N(), C = B.C();
    U32 N  = std::max(Na, Nb);
    if (Ka != Kb || N != O.N() || C != O.C()) { Runtime::error("  tensor#%s ka(%d)!=kb(%d) or N, C diff\n", nm, Ka, Kb); return O; }
    KCHK(t4k_gemm_ex(engine, A.data, B.data, O.data, alpha, beta, tA, tB, H, W, Ka, C, N,
                     Na == 1 && N > 1 ? 0 : (int64_t)A.HWC(), Nb == 1 && N > 1 ? 0 : (int64_t)B.HWC(), (int64_t)O.HWC(), ST));
    return O;
}
// gemm1/gemm2 are the reference's naive / 16x16-tiled kernels (double accumulator); all four words map
// onto the same engines here and stay as aliases
Tensor &Tensor::gemm1(Tensor &A, Tensor &B, Tensor &O, DU a, DU b, bool tA, bool tB) { return A._gemm(T4K_GEMM_SIMT, A, B, O, a, b, tA, tB, "gemm1"); }
Tensor &Tensor::gemm2(Tensor &A, Tensor &B, Tensor &O, DU a, DU b, bool tA, bool tB) { return A._gemm(T4K_GEMM_SIMT, A, B, O, a, b, tA, tB, "gemm2"); }
Tensor &Tensor::gemm3(Tensor &A, Tensor &B, Tensor &O, DU a, DU b, bool tA, bool tB) { return A._gemm(T4K_GEMM_AUTO, A, B, O, a, b, tA, tB, "gemm3"); }
Tensor &Tensor::gemm4(Tensor &A, Tensor &B, Tensor &O, DU a, DU b, bool tA, bool tB) { return A._gemm(T4K_GEMM_AUTO, A, B, O, a, b, tA, tB, "gemm4"); }
Tensor &Tensor::copy(Tensor &A, Tensor &O) { KCHK(t4k_copy(A.data, O.data, A.numel, ST)); return O; }          // tensor.cu:204-208
Tensor &Tensor::transpose(Tensor &A, Tensor &T) { KCHK(t4k_transpose(A.data, T.data, A.N(), A.H(), A.W(), A.C(), ST)); return T; } // :210-219

static DU read_scalar(DU *dev) { DU v = 0; cudaMemcpyAsync(&v, dev, sizeof(DU), cudaMemcpyDeviceToHost, (cudaStream_t)ST); Runtime::sync(); return v; }
DU Tensor::sum()  { KCHK(t4k_sum(data, numel, _tmp, ST)); return SCALAR(read_scalar(_tmp)); }                  // tensor.cu:225-236
DU Tensor::avg()  { DU v = sum() / numel; return SCALAR(v); }                                                  // :238-242
DU Tensor::std()  { KCHK(t4k_avg_std(data, numel, _tmp, ST)); return SCALAR(read_scalar(_tmp + 1)); }          // :244-251 (sqrt(Σ(x-μ)²)/n)
DU Tensor::norm() { KCHK(t4k_nvar(data, 0.0f, numel, _tmp, ST)); return SCALAR(sqrtf(read_scalar(_tmp))); }    // :253-259
DU Tensor::max()  { KCHK(t4k_minmax(data, numel, 1, _tmp, ST)); return SCALAR(read_scalar(_tmp)); }            // :261-268
DU Tensor::min()  { KCHK(t4k_minmax(data, numel, 0, _tmp, ST)); return SCALAR(read_scalar(_tmp)); }            // :269-277
DU Tensor::dot(Tensor &B) {                                                                                   // :279-287
    if (rank == 1 && B.rank == 1 && numel == B.numel) KCHK(t4k_dot(data, B.data, _tmp, 1.0f, 0.0f, (int)numel, 1, 1, 1, ST));
    else Runtime::error("A.dot(B) dim? %ld != %ld)\n", (long)numel, (long)B.numel);
    return SCALAR(read_scalar(_tmp));
}
DU Tensor::loss(t4_loss op, Tensor &tgt) {                                                                    // :289-325
    KCHK(t4k_loss(op, data, tgt.data, numel, N(), _tmp, ST));
    return SCALAR(read_scalar(_tmp));
}
U32 Tensor::has_nan() {
    KCHK(t4k_nan_inf(data, numel, (int*)_tmp, ST));
    int cnt = 0; cudaMemcpyAsync(&cnt, _tmp, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)ST); Runtime::sync(); return cnt;
}
int Tensor::h2d(const DU *h, U64 n) { return (int)cudaMemcpyAsync(data, h, sizeof(DU) * (n ? n : numel), cudaMemcpyHostToDevice, (cudaStream_t)ST); }
int Tensor::d2h(DU *h, U64 n) { cudaMemcpyAsync(h, data, sizeof(DU) * (n ? n : numel), cudaMemcpyDeviceToHost, (cudaStream_t)ST); return Runtime::sync(); }

// =============================================================================== Dataset
static cudaStream_t g_copy = nullptr;                                     // H2D feeder stream
Dataset &Dataset::create(U32 n, U32 h, U32 w, U32 c) {
    Dataset *d = new Dataset();
    const U64 sz = (U64)n * h * w * c;
    d->reset(Runtime::alloc((sz + 4) * sizeof(DU)), sz);
    d->reshape(n, h, w, c);
    d->label = (int32_t*)Runtime::alloc(((size_t)n + 4) * sizeof(int32_t));
    if (!g_copy) cudaStreamCreateWithFlags(&g_copy, cudaStreamNonBlocking);
    for (int b = 0; b < 2; b++) {
        d->_simg[b] = (uint8_t*)Runtime::alloc(sz + 16); d->_slab[b] = (uint8_t*)Runtime::alloc((size_t)n + 16);
        cudaEvent_t e; cudaEventCreateWithFlags(&e, cudaEventDisableTiming); d->_staged[b] = e;
        cudaEventCreateWithFlags(&e, cudaEventDisableTiming); d->_consumed[b] = e;
    }
    cudaMemsetAsync(d->data, 0, sz * sizeof(DU), (cudaStream_t)ST);
    cudaMemsetAsync(d->label, 0, (size_t)n * sizeof(int32_t), (cudaStream_t)ST);
    Runtime::sync();                                                       // staging buffers come from the library stream's pool
    return *d;
}
void Dataset::destroy(Dataset &d) {
    Runtime::sync(); if (g_copy) cudaStreamSynchronize(g_copy);
    for (int b = 0; b < 2; b++) { Runtime::free(d._simg[b]); Runtime::free(d._slab[b]); cudaEventDestroy((cudaEvent_t)d._staged[b]); cudaEventDestroy((cudaEvent_t)d._consumed[b]); }
    Runtime::free(d.label); Runtime::free(d.data);
    delete &d;
}
void Dataset::normalize(DU mean, DU scale) {                               // dataset.cu:33-41
    _mean = mean;
    if (fabsf(scale) < DU_EPS_H) { Runtime::error("scale == 0?\n"); _scale = 1.0f; }
    else _scale = 1.0f / scale;
}
int Dataset::stage(const uint8_t *img_host, const uint8_t *lab_host, int n) {
    if (!img_host || !lab_host || n < 1 || n > (int)N()) { Runtime::error("dataset#stage n=%d of %d\n", n, (int)N()); return T4K_EINVAL; }
    if (_head - _tail >= 2) { Runtime::error("dataset#stage: two batches already staged, commit one first\n"); return T4K_EINVAL; }
    const int b = _head & 1;
    if (_head >= 2) cudaStreamWaitEvent(g_copy, (cudaEvent_t)_consumed[b], 0);          // the batch that used this buffer has been normalised
    cudaMemcpyAsync(_simg[b], img_host, (size_t)n * HWC(), cudaMemcpyHostToDevice, g_copy);
    cudaMemcpyAsync(_slab[b], lab_host, (size_t)n, cudaMemcpyHostToDevice, g_copy);
    cudaEventRecord((cudaEvent_t)_staged[b], g_copy);
    _sn[b] = n; _head++;
    return 0;
}
int Dataset::commit_begin(const uint8_t **simg, const uint8_t **slab, int *n) {
    if (_head == _tail) { Runtime::error("dataset#commit: nothing staged\n"); return T4K_EINVAL; }
    const int b = _tail & 1;
    cudaStreamWaitEvent((cudaStream_t)ST, (cudaEvent_t)_staged[b], 0);
    *simg = _simg[b]; *slab = _slab[b]; *n = _sn[b];
    return 0;
}
int Dataset::commit_launch(const uint8_t *simg, const uint8_t *slab, int n, DU *hot, int E) {
    // partial batch: the tail keeps the previous values, as _load does
    const int rc = t4k_dataset_load(simg, data, (int64_t)n * HWC(), _mean, _scale, slab, label, n, hot, E, ST);
    KCHK(rc);
    return rc;
}
void Dataset::commit_end() {
    const int b = _tail & 1;
    cudaEventRecord((cudaEvent_t)_consumed[b], (cudaStream_t)ST);
    batch_sz = _sn[b]; _tail++; batch_id++;
}
int Dataset::commit(DU *hot, int E) {
    const uint8_t *si, *sl; int n;
    int rc = commit_begin(&si, &sl, &n);
    if (rc) return rc;
    rc = commit_launch(si, sl, n, hot, E);
    commit_end();
    return rc;
}

// =============================================================================== Model
Model::Model(U32 n, U32 h, U32 w, U32 c) { _layers.push_back(&Tensor::create(n, h, w, c)); }
Model::~Model() {
    Runtime::sync();
    for (Tensor *t : _layers) {
        for (int i = 0; i < 5; i++) { if (t->grad[i]) Tensor::destroy(*t->grad[i]); if (t->mtum[i]) Tensor::destroy(*t->mtum[i]); }
        Tensor::destroy(*t);
    }
    if (_own_hot && _hot) Tensor::destroy(*_hot);
    Runtime::free(_G); Runtime::free(_DG); Runtime::free(_M); Runtime::free(_V); Runtime::free(_seg_dev); Runtime::free(_cnt_dev); Runtime::free(_pdup); Runtime::free(_hscratch);
    _drop_graphs();
    if (_loss_pin) { cudaFreeHost(_loss_pin); for (int b = 0; b < 2; b++) cudaEventDestroy((cudaEvent_t)_loss_ev[b]); }
}
Tensor &Model::operator[](S32 i) { return *_layers[(i < 0) ? (S32)_layers.size() + i : i]; }     // model.cpp:47-49
int Model::batch_size() { return _layers.empty() ? 1 : (int)_layers[0]->N(); }
void Model::_RAND(Tensor &t, DU scale) {                                  // model.cpp:74-79: [-scale, scale)
    KCHK(t4k_rand(t.data, t.numel, T4K_UNIFORM, -0.5f, scale * 2.0f, ST));
}
// ---- layer factory (model.cpp:83-117)
Model &Model::add(t4_layer fn, U32 n, DU bias, U16 *opt) {
    Tensor &in = (*this)[-1];
    if (in.grad_fn != T4K_L_NONE) return *this;
    for (int i = 0; i < 5; i++) in.grad[i] = in.mtum[i] = nullptr;
    U16 dflt[4] = {3, 1, 0, 1};
    size_t before = _layers.size();
    switch (fn) {
    case T4K_L_CONV:    _iconv(in, n, bias, opt ? opt : dflt); break;
    case T4K_L_DCONV: { U16 d4[4] = {4, 2, 0, 1}; _iconv(in, n, bias, opt ? opt : d4, true); } break;   // word `dconv2d` = _conv(4, true, 2), netvm.cpp:315
    case T4K_L_LINEAR:  _ilinear(in, n, bias);                 break;
    case T4K_L_FLATTEN: _iflatten(in);                         break;
    case T4K_L_RELU: case T4K_L_TANH: case T4K_L_SIGMOID: case T4K_L_SELU:
    case T4K_L_LEAKYRL: case T4K_L_ELU: case T4K_L_DROPOUT: _iactivate(in, bias); break;
    case T4K_L_SOFTMAX: case T4K_L_LOGSMAX: _isoftmax(in);     break;
    case T4K_L_AVGPOOL: case T4K_L_MAXPOOL: case T4K_L_MINPOOL: _ipool(in, (U16)n); break;
    case T4K_L_BATCHNM: _ibatchnorm(in, bias);                 break;
    case T4K_L_USAMPLE: _iup(in, (U16)n, bias);                break;
    default: Runtime::error("Model#add layer %d not supported\n", fn); err = true; return *this;
    }
    if (_layers.size() > before) in.grad_fn = fn;
    return *this;
}
void Model::_iconv(Tensor &in, U32 C0, DU bias, U16 *opt, bool txn) {     // model.cpp:122-180
    U32 N1 = in.N(), H1 = in.H(), W1 = in.W(), C1 = in.C();
    U16 Kx = opt[0], Ky = opt[0], S = opt[1];
    U16 P  = (Kx > 1 && opt[2]) ? opt[2] : (Kx - 1) / 2;
    U16 H0, W0;
    if (txn) {                                                            // transposed: output padding + output size (model.cpp:129-133)
        const U16 P0 = (H1 + P * 2 - Kx) % S;
        H0 = (H1 - 1) * S - P * 2 + Kx + P0;
        W0 = (W1 - 1) * S - P * 2 + Ky + P0;
    } else {
        H0 = (H1 - Kx + P * 2) / S + 1;
        W0 = (H1 - Ky + P * 2) / S + 1;                                   // sic: W0 from H1 (model.cpp:137)
    }
    if ((!txn && Kx != 1 && Kx != 3 && Kx != 5) || (txn && Kx != 4)) {
        Runtime::error("nn#iconv %s f=[%d,%d]? 1x1, 3x3, 4x4, and 5x5 supported only.\n", txn ? "dconv2d" : "conv2d", Kx, Ky); err = true; return;
    }
    in.stride[0] = in.stride[1] = S; in.stride[2] = in.stride[3] = P; in.xparm = bias;
    // conv: f [C1][K][K][C0].  conv-transpose: the same number of elements, held as the filter [C0][K][K][C1] of the convolution (C0 -> C1) whose
    // kernels the layer runs with swapped roles (t4k_dconv2d_*); Kaiming range from the layer's input channels either way (model.cpp:160)
    Tensor *f = in.grad[0] = txn ? &Tensor::create(C0, Kx, Ky, C1) : &Tensor::create(C1, Kx, Ky, C0);
    Tensor *b = in.grad[1] = &Tensor::create((U64)C0);
    in.grad[2] = txn ? &Tensor::create(C0, Kx, Ky, C1).zeros() : &Tensor::create(C1, Kx, Ky, C0).zeros();
    in.grad[3] = &Tensor::create((U64)C0).zeros();
    in.grad[4] = &Tensor::create(N1, H1, in.W(), C1).zeros();
    DU k = sqrtf(6.0f / (Kx * Ky * C1));
    _RAND(*f, k); _RAND(*b, bias);
    _layers.push_back(&Tensor::create(N1, H0, W0, C0));
}
void Model::_ilinear(Tensor &in, U32 E0, DU bias) {                       // model.cpp:183-226
    U32 N1 = in.N(); U64 E1 = in.HWC();
    Tensor *w = in.grad[0] = &Tensor::create(1, E0, (U32)E1, 1);
    Tensor *b = in.grad[1] = &Tensor::create((U64)E0);
    in.grad[2] = &Tensor::create(1, E0, (U32)E1, 1).zeros();
    in.grad[3] = &Tensor::create((U64)E0).zeros();
    in.xparm = bias;
    DU k = sqrtf(1.0f / (E0 + E1));
    _RAND(*w, k); _RAND(*b, bias);
    _layers.push_back(&Tensor::create(N1, 1, E0, 1));
}
void Model::_iflatten(Tensor &in) { _layers.push_back(&Tensor::create(in.N(), 1, (U32)in.HWC(), 1)); }          // model.cpp:227-233
void Model::_isoftmax(Tensor &in) { in.grad[4] = &Tensor::create(1, in.H(), in.W(), in.C()); _layers.push_back(&Tensor::create_like(in)); } // :237-245
void Model::_iactivate(Tensor &in, DU alpha) { in.grad[4] = &Tensor::create_like(in); in.xparm = alpha; _layers.push_back(&Tensor::create_like(in)); } // :247-256
void Model::_ipool(Tensor &in, U16 k) {                                   // model.cpp:260-274
    if (k != 2 && k != 3) { Runtime::error("nn#ipool k=%dx%d? 2x2 and 3x3 supported only\n", k, k); err = true; return; }
    U32 H0 = (in.H() + k - 1) / k, W0 = (in.W() + k - 1) / k;
    U16 s[4] = {k, 1, 1, 0}; memcpy(in.stride, s, sizeof(s));
    _layers.push_back(&Tensor::create(in.N(), H0, W0, in.C()));
}
void Model::_ibatchnorm(Tensor &in, DU m) {                               // model.cpp:276-292
    const U32 C = in.C();
    in.grad[0] = &Tensor::create((U64)C).map(T4K_FILL, 1.0f);
    in.grad[2] = &Tensor::create((U64)C).zeros();     // reference leaves d_gamma/d_beta uninitialised (model.cpp:281,283); zero is the value a fresh arena has
    in.grad[1] = &Tensor::create((U64)C).zeros();
    in.grad[3] = &Tensor::create((U64)C).zeros();
    in.grad[4] = &Tensor::create_like(in);
    in.mtum[4] = &Tensor::create((U64)C * 3);
    in.xparm = m;
    _layers.push_back(&Tensor::create_like(in));
}
void Model::_iup(Tensor &in, U16 k, DU method) {                          // model.cpp:294-310
    if (k != 2 && k != 3) { Runtime::error("nn#iup k=%dx%d? only 2x2 and 3x3 supported\n", k, k); err = true; return; }
    in.iparm = (U32)method;
    U16 s[4] = {k, 1, 1, 1}; memcpy(in.stride, s, sizeof(s));
    _layers.push_back(&Tensor::create(in.N(), in.H() * k, in.W() * k, in.C()));
}
// ---- forward (forward.cu:29-113)
Model &Model::forward(Tensor &input) {
    Tensor &n0 = (*this)[0];
    if (input.numel != n0.numel) {
        Runtime::error("nn#forward dataset wrong shape[%d,%d,%d,%d] != model input[%d,%d,%d,%d]\n",
                       input.N(), input.H(), input.W(), input.C(), n0.N(), n0.H(), n0.W(), n0.C());
        _feed = nullptr;
        return *this;
    }
    size_t i0 = 0;
    if (input.data != n0.data) {
        i0 = (size_t)_ffused(0, input.data);               // first block fused: the `n0 = input` copy (and a pending dataset feed) ride in the same launch
        if (_feed) _feed_fallback();                       // not taken by a fused block: the plain load, before the copy below reads the tensor
        if (!i0) n0 = input;
    }
    else if (_feed) _feed_fallback();
    for (size_t i = i0; i + 1 < _layers.size(); ) {
        int adv = _ffused(i);                              // conv → maxpool(2) → relu (→ flatten) in one launch
        if (!adv) adv = _ffused_linear(i);                 // linear → activation | linear → softmax
        if (adv) { i += adv; continue; }
        _fstep(*_layers[i], *_layers[i + 1]); i++;
    }
    return *this;
}
void Model::_feed_fallback() {
    const StepExtra &x = *_feed; _feed = nullptr;
    x.ds->commit_launch(x.simg, x.slab, x.n, _feed_hot, (int)(*this)[-1].HWC());
}
// The canonical CNN block of the reference's examples ("conv2d 2 maxpool relu [flatten]", t4_40a.4th:11-12):
// same layer tensors written as the per-layer path (forward.cu:83-113), one kernel.  Returns layers consumed.
int Model::_ffused(size_t i, const DU *src) {
    const size_t n = _layers.size();
    if (!fuse || i + 3 >= n) return 0;
    Tensor &in = *_layers[i], &co = *_layers[i + 1], &po = *_layers[i + 2], &ao = *_layers[i + 3];
    if (in.grad_fn != T4K_L_CONV || co.grad_fn != T4K_L_MAXPOOL || co.stride[0] != 2 || po.grad_fn != T4K_L_RELU) return 0;
    Tensor *fl = (ao.grad_fn == T4K_L_FLATTEN && i + 4 < n) ? _layers[i + 4] : nullptr;
    Tensor &f = *in.grad[0], &b = *in.grad[1];
    if (src && _feed) {                                    // step_graph with a Dataset: U8 -> normalise -> one-hot inside this launch
        const StepExtra &x = *_feed; _feed = nullptr;      // the entry point always performs the load (fused, or as its own launch in front)
        int rc = t4k_conv_pool_relu_fwd_feed(x.simg, x.slab, x.n, x.ds->_mean, x.ds->_scale, x.ds->label, _feed_hot, (int)(*this)[-1].HWC(),
                                             (DU*)src, f.data, b.data, in.data, co.data, po.data, ao.data, po.grad[4]->data, fl ? fl->data : nullptr,
                                             co.N(), in.H(), in.W(), in.C(), co.H(), co.W(), co.C(), f.H(), in.stride[0], in.stride[2], ST);
        if (rc == T4K_ENOSUP) return 0;
        KCHK(rc);
        return fl ? 4 : 3;
    }
    int rc = t4k_conv_pool_relu_fwd(src ? src : in.data, f.data, b.data, src ? in.data : nullptr, co.data, po.data, ao.data, po.grad[4]->data, fl ? fl->data : nullptr,
                                    co.N(), in.H(), in.W(), in.C(), co.H(), co.W(), co.C(), f.H(), in.stride[0], in.stride[2], ST);
    if (rc == T4K_ENOSUP) return 0;
    KCHK(rc);
    return fl ? 4 : 3;
}
// linear → activation: bias + activation (+ mask) ride in the GEMM's split-K finish; linear → softmax (small head): one launch.
// Same layer tensors as _flinear + _factivate / _fsoftmax (forward.cu:158-243).
static bool mask_act(t4_layer fn) { return fn == T4K_L_RELU || fn == T4K_L_TANH || fn == T4K_L_SELU || fn == T4K_L_LEAKYRL || fn == T4K_L_ELU; }
int Model::_ffused_linear(size_t i) {
    const size_t n = _layers.size();
    if (!fuse || i + 2 >= n) return 0;
    Tensor &in = *_layers[i], &lo = *_layers[i + 1], &ao = *_layers[i + 2];
    if (in.grad_fn != T4K_L_LINEAR) return 0;
    const int N = (int)lo.N(), E0 = (int)lo.HWC(), E1 = (int)in.HWC();
    const t4_layer fn = lo.grad_fn;
    int rc;
    // linear -> activation -> small linear -> softmax at the end of the model: the split-K finish of the first linear and the whole head
    // are one launch (the activations of a row never leave the warp that owns it)
    if ((mask_act(fn) || fn == T4K_L_SIGMOID) && i + 5 == n && ao.grad_fn == T4K_L_LINEAR && _layers[i + 3]->grad_fn == T4K_L_SOFTMAX) {
        Tensor &l2o = *_layers[i + 3], &po = *_layers[i + 4];
        DU *dup = (_want_pdup && _pdup) ? _pdup : nullptr;
        _tail_done = false;
        if (_fwd_tgt && _hscratch && train && mask_act(fn) && _fwd_tgt->numel == po.numel) {
            // fused train step: forward tail + the head's backward on the same rows, one launch (the layer tensors get their backward values)
            rc = t4k_linear_act_head_train(fn, in.data, in.grad[0]->data, in.grad[1]->data, lo.data, ao.data, lo.grad[4]->data, lo.xparm,
                                           ao.grad[0]->data, ao.grad[1]->data, l2o.data, po.data, dup, _fwd_tgt->data, _hscratch, &_hncta,
                                           N, E0, E1, (int)l2o.HWC(), ST);
            if (rc == 0) { _tail_done = true; _pdup_valid = dup != nullptr; return 4; }
            if (rc != T4K_ENOSUP) KCHK(rc);
        }
        rc = t4k_linear_act_head_fwd(fn, in.data, in.grad[0]->data, in.grad[1]->data, lo.data, ao.data, lo.grad[4]->data, lo.xparm,
                                     ao.grad[0]->data, ao.grad[1]->data, l2o.data, po.data, dup, N, E0, E1, (int)l2o.HWC(), ST);
        if (rc != T4K_ENOSUP) { KCHK(rc); _pdup_valid = (rc == 0 && dup != nullptr); return 4; }
    }
    if (fn == T4K_L_SOFTMAX) {
        DU *dup = nullptr;
        if (_want_pdup && _pdup && i + 3 == n) {           // the model's output layer, inside step_graph: keep a copy of p for the loss
            dup = _pdup;                                   // allocated by _step_graph, outside any stream capture
        }
        rc = t4k_mlp_head_fwd_dup(in.data, in.grad[0]->data, in.grad[1]->data, lo.data, ao.data, dup, N, E0, E1, ST);
        _pdup_valid = (rc == 0 && dup != nullptr);
    }
    else if (mask_act(fn) || fn == T4K_L_SIGMOID)
        rc = t4k_linear_act_fwd(fn, in.data, in.grad[0]->data, in.grad[1]->data, lo.data, ao.data, lo.grad[4]->data, lo.xparm, N, E0, E1, ST);
    else return 0;
    if (rc == T4K_ENOSUP) return 0;
    KCHK(rc);
    return 2;
}
// tail of a classifier, backward: [linear →] activation → small linear → softmax with target y.  One launch does
// _bprep (p - y), the softmax pass-through, the small linear's dB/dW/dX, the activation backward and the dB of the
// linear in front (backprop.cu:76-140,194-263).  Returns the index of the next layer to process (-1: none left) or -2: not fused.
int Model::_bfused_head(Tensor &tgt, bool *skip_db) {
    const int n = (int)_layers.size();
    *skip_db = false;
    if (!fuse || n < 4) return -2;
    Tensor &P = *_layers[n - 1], &yl = *_layers[n - 2], &x2 = *_layers[n - 3];
    if (yl.grad_fn != T4K_L_SOFTMAX || x2.grad_fn != T4K_L_LINEAR) return -2;
    const int N = (int)P.N(), E0 = (int)P.HWC(), E1 = (int)x2.HWC();
    if (E0 > 32 || E1 > 128) return -2;
    Tensor *act = (n >= 5 && (mask_act(_layers[n - 4]->grad_fn) || _layers[n - 4]->grad_fn == T4K_L_DROPOUT)) ? _layers[n - 4] : nullptr;
    const int prev = act ? n - 5 : n - 4;
    Tensor *lin1 = (prev >= 0 && _layers[prev]->grad_fn == T4K_L_LINEAR && train) ? _layers[prev] : nullptr;
    if (_tail_done) {
        // the forward tail kernel of this step already wrote p - y, the head linear's dX and the activation backward (Model::forward, train tail);
        // what is left are the head's parameter gradients: per-CTA partials -> dW2, dB2, dB1 on the side stream, joined at the end of the step
        _tail_done = false;
        cudaStream_t st = (cudaStream_t)ST;
        cudaEventRecord(g_fork, st); cudaStreamWaitEvent(g_stream2, g_fork, 0);
        KCHK(t4k_head_grad_finish(_hscratch, _hncta, E0, E1, x2.grad[2]->data, x2.grad[3]->data, lin1 ? lin1->grad[3]->data : nullptr, (t4k_stream_t)g_stream2));
        cudaEventRecord(g_join, g_stream2);
        _side_join = true;
        *skip_db = lin1 != nullptr;
        return prev;
    }
    // Inside step_graph (a duplicate of the softmax output exists) with the hidden linear's dW on the side stream (a flatten in front of it
    // holds a second copy of X): the head kernel is NOT launched here.  Model::_blinear puts it on the side stream in front of the dW GEMM,
    // and the dX GEMM of the hidden layer — the critical path — evaluates p - y, the small linear's dX and the activation backward in its
    // operand producer from tensors nobody overwrites (t4k_linear_dx_from_head).
    if (lin1 && _pdup_valid && _pdup && prev > 0 && _layers[prev - 1]->grad_fn == T4K_L_FLATTEN && _layers[prev - 1]->data != lin1->data &&
        _layers[prev - 1]->numel == lin1->numel && (E1 & 3) == 0 && (lin1->HWC() & 3) == 0) {
        _hp.on = true; _hp.P = P.data; _hp.T = tgt.data; _hp.Ylin = yl.data; _hp.X2 = x2.data;
        _hp.F1 = act ? act->grad[4]->data : nullptr; _hp.Y1 = act ? act->data : nullptr; _hp.W2 = x2.grad[0]->data;
        _hp.dW2 = x2.grad[2]->data; _hp.dB2 = x2.grad[3]->data; _hp.dB1 = lin1->grad[3]->data; _hp.N = N; _hp.E0 = E0; _hp.E1 = E1;
        *skip_db = true;
        return prev;
    }
    int rc = t4k_mlp_head_bwd(P.data, tgt.data, yl.data, x2.data, act ? act->grad[4]->data : nullptr, act ? act->data : nullptr,
                              x2.grad[0]->data, x2.grad[2]->data, x2.grad[3]->data, lin1 ? lin1->grad[3]->data : nullptr,
                              N, E0, E1, train, ST);
    if (rc == T4K_ENOSUP) return -2;
    KCHK(rc);
    *skip_db = lin1 != nullptr;
    return prev;
}
int Model::_bfused(int i) {                                // i = index of the block's LAST layer (relu or flatten); returns layers consumed
    if (!fuse) return 0;
    const bool flat = _layers[i]->grad_fn == T4K_L_FLATTEN;
    const int ir = flat ? i - 1 : i;                       // relu layer index
    if (ir < 2) return 0;
    Tensor &in = *_layers[ir - 2], &co = *_layers[ir - 1], &po = *_layers[ir], &ao = *_layers[ir + 1];
    if (po.grad_fn != T4K_L_RELU || co.grad_fn != T4K_L_MAXPOOL || co.stride[0] != 2 || in.grad_fn != T4K_L_CONV) return 0;
    Tensor &dy = *_layers[i + 1];                          // gradient arriving at the block output
    Tensor &f = *in.grad[0], &df = *in.grad[2], &db = *in.grad[3], &dx = *in.grad[4];
    // _skip_flat_copy: the flatten backward (ao = dy) was issued on the side stream behind the dW GEMM that still reads ao
    // (see backprop): passing dy as the copy's destination makes the kernel skip it
    DU *flat_dst = (flat && _skip_flat_copy) ? dy.data : ao.data;
    int rc = T4K_ENOSUP;
    if (_oe.on && _oe.rest && !_oe.first && train && df.data == _DG && db.data > _DG && db.data < _DG + _first_end) {
        // the block of the FIRST parameter layer, the rest of the arena already stepped on the side stream: its finish launch steps its own segments
        t4k_fused_opt_t fo{_oe.kind, _oe.lr, _oe.b1, _oe.b2, _oe.wd, _G, _M, _V, 0, (int64_t)(db.data - _DG), (int32_t)f.N(), (int32_t)in.grad[1]->N()};
        if (_oe.late) t4k_conv_pool_relu_bwd_mid_event((void*)g_mid);
        rc = t4k_conv_pool_relu_bwd_opt(dy.data, flat_dst, po.grad[4]->data, po.data, co.data, in.data, dx.data, f.data, df.data, db.data,
                                        in.N(), in.H(), in.W(), in.C(), co.H(), co.W(), co.C(), f.H(), in.stride[0], in.stride[2], train, &fo, ST);
        t4k_conv_pool_relu_bwd_mid_event(nullptr);
        if (rc == 0) _oe.first = true;
        if (rc == 0 && _oe.late) {                          // the rest of the arena: stepped on the side stream from the end of the block's main kernel
            _oe.late = false;
            cudaStreamWaitEvent(g_stream2, g_mid, 0);
            KCHK(t4k_optim_multi_range(_oe.kind, _G, _DG, _M, _V, (const t4k_seg_t*)_seg_dev, _nseg, _first_end, (int64_t)_total, _oe.lr, _oe.b1, _oe.b2, _oe.wd,
                                       (t4k_stream_t)g_stream2));
            cudaEventRecord(g_join, g_stream2);
            _side_join = true;
        }
    }
    // data parallel, the block of the FIRST parameter layer, the rest of the arena already pushed to the peers (copy engines, _dp_push): the
    // exchange + optimizer of that rest runs on the side stream from the moment the block's main kernel is done (it must not take SMs from it:
    // one exact wave), under the block's finish launch — the end of the step then only exchanges the first chunk
    const bool dp_rest = _comm && _dpo.on && !_dpo.rest && dp_early_mode(_dp_world) <= 1 && dp_rest_on() && _dp_pushed_from > 0 && _dp_pushed_from < (int64_t)_total && train && df.data == _DG;
    if (rc == T4K_ENOSUP) {
        if (dp_rest) t4k_conv_pool_relu_bwd_mid_event((void*)g_mid);
        rc = t4k_conv_pool_relu_bwd(dy.data, flat_dst, po.grad[4]->data, po.data, co.data, in.data, dx.data, f.data, df.data, db.data,
                                    in.N(), in.H(), in.W(), in.C(), co.H(), co.W(), co.C(), f.H(), in.stride[0], in.stride[2], train, ST);
        t4k_conv_pool_relu_bwd_mid_event(nullptr);
        if (dp_rest && rc == 0) {
            cudaStreamWaitEvent(g_stream2, g_mid, 0);         // behind the block's main kernel — and behind the push itself (same side stream)
            const int rr = _dpo.rs ? t4k_optim_multi_dp_rs((t4k_comm_t)_comm, _dpo.kind, _G, _DG, _M, _V, (const t4k_seg_t*)_seg_dev, _nseg, _dp_pushed_from, (int64_t)_total,
                                                           _dpo.lr, _dpo.b1, _dpo.b2, _dpo.wd, dp_rs_phased() ? 2 : 0, (t4k_stream_t)g_stream2)
                                   : t4k_optim_multi_dp_range((t4k_comm_t)_comm, _dpo.kind, _G, _DG, _M, _V, (const t4k_seg_t*)_seg_dev, _nseg, _dp_pushed_from, (int64_t)_total,
                                                              (int64_t)_total, _dpo.lr, _dpo.b1, _dpo.b2, _dpo.wd, nullptr, 0, _dp_pushed_from, (t4k_stream_t)g_stream2);
            cudaEventRecord(g_join, g_stream2);
            _side_join = true;
            if (rr == 0) _dpo.rest = true; else Runtime::error("t4k_optim_multi_dp_range -> %d", rr);
        }
    }
    if (rc == T4K_ENOSUP) return 0;
    KCHK(rc);
    _skip_flat_copy = false;
    return flat ? 4 : 3;
}
void Model::_fstep(Tensor &in, Tensor &out) {                             // forward.cu:83-113
    t4_layer fn = in.grad_fn;
    switch (fn) {
    case T4K_L_CONV:    _fconv(in, out);   break;
    case T4K_L_DCONV:   _fdconv(in, out);  break;
    case T4K_L_LINEAR:  _flinear(in, out); break;
    case T4K_L_FLATTEN: out = in;          break;
    case T4K_L_RELU: case T4K_L_TANH: case T4K_L_SIGMOID: case T4K_L_SELU:
    case T4K_L_LEAKYRL: case T4K_L_ELU: _factivate(in, out, fn); break;
    case T4K_L_DROPOUT: {
        // fresh mask every forward (forward.cu:98-102), drawn and applied in one launch; data parallel: the mask of THIS shard of the global batch (rand.cu)
        Tensor &t = *in.grad[4];
        const int64_t n = (int64_t)t.numel;
        KCHK(t4k_dropout_fwd(in.data, out.data, t.data, in.xparm, n, _dp_world > 1 ? (int64_t)_dp_rank * n : 0, _dp_world > 1 ? (int64_t)_dp_world * n : n, ST));
    } break;
    case T4K_L_SOFTMAX: _fsoftmax(in, out);    break;
    case T4K_L_LOGSMAX: _flogsoftmax(in, out); break;
    case T4K_L_AVGPOOL: case T4K_L_MAXPOOL: case T4K_L_MINPOOL: _fpool(in, out, fn); break;
    case T4K_L_BATCHNM: _fbatchnorm(in, out);  break;
    case T4K_L_USAMPLE: _fupsample(in, out);   break;
    default: Runtime::error("nn#fstep layer=%d not supported\n", fn);
    }
}
int Model::_fconv(Tensor &in, Tensor &out) {                              // forward.cu:126-155
    Tensor &f = *in.grad[0], &b = *in.grad[1];
    int rc = t4k_conv2d_fwd(in.data, f.data, b.data, out.data, out.N(), in.H(), in.W(), in.C(), out.H(), out.W(), out.C(),
                            f.H(), in.stride[0], in.stride[2], ST);
    if (rc == T4K_ENOSUP) { Runtime::error("nn#fconv kernel_size=%d stride=%d padding=%d not supported\n", f.H(), in.stride[0], in.stride[2]); return -1; }
    KCHK(rc);
    return 0;
}
int Model::_fdconv(Tensor &in, Tensor &out) {                             // forward.cu:110 (the convolution's kernels, roles swapped)
    Tensor &f = *in.grad[0], &b = *in.grad[1];
    int rc = t4k_dconv2d_fwd(in.data, f.data, b.data, out.data, out.N(), in.H(), in.W(), in.C(), out.H(), out.W(), out.C(),
                             f.H(), in.stride[0], in.stride[2], ST);
    if (rc == T4K_ENOSUP) { Runtime::error("nn#fconv kernel_size=%d stride=%d padding=%d not supported\n", f.H(), in.stride[0], in.stride[2]); return -1; }
    KCHK(rc);
    return 0;
}
int Model::_bdconv(Tensor &in, Tensor &out) {                             // backprop.cu:137
    Tensor &f = *in.grad[0], &df = *in.grad[2], &db = *in.grad[3], &dx = *in.grad[4];
    int rc = t4k_dconv2d_bwd(in.data, out.data, f.data, dx.data, df.data, db.data, in.N(), in.H(), in.W(), in.C(),
                             out.H(), out.W(), out.C(), f.H(), in.stride[0], in.stride[2], train, ST);
    if (rc == T4K_ENOSUP) { Runtime::error("nn#bconv kernel_size=%d stride=%d padding=%d not supported\n", f.H(), in.stride[0], in.stride[2]); return -1; }
    KCHK(rc);
    in = dx;                                                               // x = dX (overwrite), as Model::_bconv
    return 0;
}
int Model::_flinear(Tensor &in, Tensor &out) {                            // forward.cu:158-198
    KCHK(t4k_linear_fwd(in.data, in.grad[0]->data, in.grad[1]->data, out.data, out.N(), (int)out.HWC(), (int)in.HWC(), ST));
    return 0;
}
int Model::_factivate(Tensor &in, Tensor &out, t4_layer fn) {             // forward.cu:201-209
    KCHK(t4k_activate_fwd(fn, in.data, out.data, in.grad[4]->data, in.xparm, in.numel, ST)); return 0;
}
int Model::_fpool(Tensor &in, Tensor &out, t4_layer fn) {                 // forward.cu:212-228
    int rc = t4k_pool_fwd(fn, in.data, out.data, out.N(), in.H(), in.W(), out.H(), out.W(), out.C(), in.stride[0], ST);
    if (rc == T4K_ENOSUP) { Runtime::error("nn#fpool kernel_size=%d not supported\n", in.stride[0]); return -1; }
    KCHK(rc); return 0;
}
int Model::_fsoftmax(Tensor &in, Tensor &out)    { KCHK(t4k_softmax_fwd(in.data, out.data, in.N(), (int)in.HWC(), ST)); return 0; }     // forward.cu:231-243
int Model::_flogsoftmax(Tensor &in, Tensor &out) { KCHK(t4k_logsoftmax_fwd(in.data, out.data, in.N(), (int)in.HWC(), ST)); return 0; }  // forward.cu:246-259
int Model::_fbatchnorm(Tensor &in, Tensor &out) {                         // forward.cu:264-309
    if (_comm_stat && _dp_world > 1) {                                     // statistics of the GLOBAL batch (SURVEY §8e collective 2)
        KCHK(t4k_batchnorm_fwd_dp((t4k_comm_t)_comm_stat, in.data, out.data, in.grad[4]->data, in.grad[0]->data, in.grad[1]->data, in.mtum[4]->data,
                                  out.N(), out.N() * _dp_world, out.H() * out.W(), out.C(), ST));
        return 0;
    }
    KCHK(t4k_batchnorm_fwd(in.data, out.data, in.grad[4]->data, in.grad[0]->data, in.grad[1]->data, in.mtum[4]->data,
                           out.N(), out.H() * out.W(), out.C(), ST));
    return 0;
}
int Model::_fupsample(Tensor &in, Tensor &out) {                          // forward.cu:314-329 (nearest: every cell of the KxK block = in)
    KCHK(t4k_pool_bwd(T4K_L_USAMPLE, out.data, in.data, in.N(), out.H(), out.W(), in.H(), in.W(), in.C(), in.stride[0], ST)); return 0;
}
// ---- backprop (backprop.cu:40-140)
Model &Model::backprop() {
    if (_hot) return backprop(*_hot);
    Runtime::error("nn#backprop missing onehot vector?\n");
    return *this;
}
Model &Model::backprop(Tensor &tgt) {
    int i = (int)_layers.size() - 2, j = 0;
    bool skip_db = false;
    Tensor &out = (*this)[-1];
    const int nxt = (out.numel == tgt.numel) ? _bfused_head(tgt, &skip_db) : -2;    // -2: head not fused
    if (nxt == -2) { if (_bprep(tgt)) return *this; }
    else { i = nxt; j = 1; }
    for (; i >= 0; j++) {
        const t4_layer fn = _layers[i]->grad_fn;
        if (_dp_early && _dp_pushed_from < 0 && i < _second_layer) _dp_push();     // every gradient but the first parameter layer's is final
        if (_oe.on && !_oe.rest && i < _second_layer && (int64_t)_total > _first_end) {
            // T4K_OPT_LATE=1: do not launch the rest-of-arena optimizer next to the conv block's main kernel (which fills the machine in exactly one
            // wave: every CTA of another kernel resident on an SM displaces one of its CTAs) but from the block's mid event, under its finish launch
            if (opt_late_on() && j > 0 && (fn == T4K_L_FLATTEN || fn == T4K_L_RELU)) { _oe.rest = true; _oe.late = true; }
            else _opt_push();
        }
        const int adv = (j > 0 && (fn == T4K_L_FLATTEN || fn == T4K_L_RELU)) ? _bfused(i) : 0;
        if (adv) { i -= adv; continue; }
        if (_skip_flat_copy) {                              // the fused block did not take the flatten: rejoin before the per-layer path touches it
            cudaStreamWaitEvent((cudaStream_t)ST, g_join, 0); _skip_flat_copy = false; _side_join = false;
        }
        if (fn == T4K_L_LINEAR && j > 0) {
            // a flatten in front of the linear layer leaves a second copy of X in its own input tensor (flatten is a copy and its
            // backward has not run yet): dW can read that copy while dX overwrites X in place — two independent GEMMs, forked
            Tensor *xdup = (fuse && i > 0 && _layers[i - 1]->grad_fn == T4K_L_FLATTEN && _layers[i - 1]->data != _layers[i]->data &&
                            _layers[i - 1]->numel == _layers[i]->numel) ? _layers[i - 1] : nullptr;
            // when the conv->pool->relu->flatten block kernel comes next (it is the only other writer of the duplicate), the side
            // branch also takes the flatten backward (duplicate = dX) and is joined at the end of backprop: dW runs under dX + conv block
            const bool defer = xdup && i >= 4 && _layers[i - 2]->grad_fn == T4K_L_RELU && _layers[i - 3]->grad_fn == T4K_L_MAXPOOL &&
                               _layers[i - 4]->grad_fn == T4K_L_CONV;
            // an activation / dropout in front of the linear layer: its backward (dX * saved mask) rides in the dX GEMM's epilogue
            Tensor *pa = (fuse && !xdup && i > 0 && (mask_act(_layers[i - 1]->grad_fn) || _layers[i - 1]->grad_fn == T4K_L_DROPOUT) &&
                          _layers[i - 1]->numel == _layers[i]->numel && _layers[i - 1]->grad[4]) ? _layers[i - 1] : nullptr;
            if (pa) { _blinear_act(*_layers[i], *_layers[i + 1], *pa, skip_db); skip_db = false; i -= 2; j++; continue; }
            _blinear(*_layers[i], *_layers[i + 1], skip_db, xdup, defer); skip_db = false;
        }
        else _bstep(*_layers[i], *_layers[i + 1], j == 0);
        i--;
    }
    if (_side_join) {                                                                               // side-stream branch of this backprop
        if (_comm && _dpo.rest && dp_early_mode(_dp_world) <= 1)
            // data parallel: the exchange + optimizer of the rest of the arena is still running there and nothing at the end of the step reads what
            // it writes — the optimizer call waits for the side stream's EARLIER work only (loss for the scalars), the step's end joins the rest
            cudaStreamWaitEvent((cudaStream_t)ST, g_push, 0);
        else { cudaStreamWaitEvent((cudaStream_t)ST, g_join, 0); _side_join = false; }
    }
    return *this;
}
void Model::_opt_push() {
    // the optimizer of every parameter layer but the first, on the side stream (behind the dW GEMMs that live there, and behind everything the
    // library stream has issued so far); joined at the end of backprop / the step
    cudaStream_t st = (cudaStream_t)ST;
    cudaEventRecord(g_fork, st); cudaStreamWaitEvent(g_stream2, g_fork, 0);
    KCHK(t4k_optim_multi_range(_oe.kind, _G, _DG, _M, _V, (const t4k_seg_t*)_seg_dev, _nseg, _first_end, (int64_t)_total, _oe.lr, _oe.b1, _oe.b2, _oe.wd,
                               (t4k_stream_t)g_stream2));
    cudaEventRecord(g_join, g_stream2);
    _side_join = true; _oe.rest = true;
}
void Model::_dp_push() {
    // Split exchange: fork a side stream off the library stream once every gradient but the first parameter layer's is final.
    cudaStream_t st = (cudaStream_t)ST;
    cudaEventRecord(g_fork, st); cudaStreamWaitEvent(g_stream2, g_fork, 0);
    const int64_t chf = t4k_comm_chunk_floats((t4k_comm_t)_comm);
    const int64_t split = chf > 0 ? ((_first_end + chf - 1) / chf) * chf : 0;          // first chunk boundary at or past the first layer's segments
    if (dp_early_mode(_dp_world) == 2 && _dpo.on && split > 0 && split < (int64_t)_total) {
        // (a) the whole exchange + optimizer of the chunks past the first layer runs THERE, under the first layer's backward (the peers reach
        // this point at the same place of their step): what stays on the critical path at the end of the step is the first chunk's exchange alone
        const int rc = t4k_optim_multi_dp_range((t4k_comm_t)_comm, _dpo.kind, _G, _DG, _M, _V, (const t4k_seg_t*)_seg_dev, _nseg, split, (int64_t)_total, (int64_t)_total,
                                                _dpo.lr, _dpo.b1, _dpo.b2, _dpo.wd, nullptr, 0, 0, (t4k_stream_t)g_stream2);
        cudaEventRecord(g_join, g_stream2);
        _side_join = true;
        if (rc == 0) { _dp_pushed_from = split; _dpo.rest = true; return; }
        Runtime::error("t4k_optim_multi_dp_range -> %d", rc);
    }
    // (b) push the finished part of the gradient arena to the peers while the remaining backward kernels run — by copy engine (no SM taken from
    // the conv block's backward, which fills the machine in exactly one wave) or, T4K_DP_EARLY=sm, by the MODE -1 kernel of comm.cu;
    // Model::_gradient joins it in front of the optimizer.
    if (dp_early_mode(_dp_world) == 1) {
        const int64_t r = t4k_dp_push_dma((t4k_comm_t)_comm, _DG, _first_end, (int64_t)_total, _dp_step, (t4k_stream_t)g_stream2);
        cudaEventRecord(g_join, g_stream2);
        cudaEventRecord(g_push, g_stream2);                 // everything the side stream holds up to here (loss, head gradients, the push)
        _dp_join = true;
        if (r < 0) { Runtime::error("t4k_dp_push_dma -> %ld", (long)r); _dp_pushed_from = (int64_t)_total; }
        else _dp_pushed_from = r;
        return;
    }
    // more than two ranks, optimizer arguments known (inside step_graph, not the first call): reduce-scatter flavour — every chunk goes to ONE owner
    // (1x the arena leaves the GPU instead of (world-1)x, next to the conv block's backward), the owners' sums come back in the rest exchange
    const bool rs = _dp_world > 4 && _dpo.on && dp_rest_on() && dp_rs_on();     // measured: 8 GPUs 89.0 vs 94.9 us per step, 4 GPUs 87.3 vs 84.4 (the all-to-all wins there)
    const int64_t r = rs ? t4k_dp_push_owner((t4k_comm_t)_comm, _DG, _first_end, (int64_t)_total, (t4k_stream_t)g_stream2)
                         : t4k_dp_push((t4k_comm_t)_comm, _DG, _first_end, (int64_t)_total, (t4k_stream_t)g_stream2);
    if (rs && dp_rs_phased() && r >= 0 && r < (int64_t)_total)   // the owners' half right behind the push: a handful of blocks per rank, they fit next to the conv block's backward
        KCHK(t4k_optim_multi_dp_rs((t4k_comm_t)_comm, _dpo.kind, _G, _DG, _M, _V, (const t4k_seg_t*)_seg_dev, _nseg, r, (int64_t)_total,
                                   _dpo.lr, _dpo.b1, _dpo.b2, _dpo.wd, 1, (t4k_stream_t)g_stream2));
    cudaEventRecord(g_join, g_stream2);
    cudaEventRecord(g_push, g_stream2);
    _dp_join = true;
    if (r < 0) { Runtime::error("t4k_dp_push -> %ld", (long)r); _dp_pushed_from = (int64_t)_total; }
    else { _dp_pushed_from = r; _dpo.rs = rs && r < (int64_t)_total; }
}
int Model::_bprep(Tensor &tgt) {                                          // backprop.cu:76-109
    Tensor &out = (*this)[-1];
    if (out.numel != tgt.numel) {
        Runtime::error("Model#bprep: Onehot wrong shape[%d,%d,%d,%d] != [%d,%d,%d,%d], numel=%ld,%ld ",
                       tgt.N(), tgt.H(), tgt.W(), tgt.C(), out.N(), out.H(), out.W(), out.C(), (long)tgt.numel, (long)out.numel);
        return 1;
    }
    switch ((*this)[-2].grad_fn) {
    case T4K_L_LINEAR: case T4K_L_SIGMOID: case T4K_L_SOFTMAX: case T4K_L_LOGSMAX:
        KCHK(t4k_tt_op(T4K_SUB, out.data, tgt.data, out.data, out.numel, 1, 1, ST)); break;       // out -= tgt  (p - y, not divided by N)
    default: out = tgt; break;
    }
    return 0;
}
void Model::_bstep(Tensor &in, Tensor &out, bool last_layer) {            // backprop.cu:112-140
    t4_layer fn = in.grad_fn;
    switch (fn) {
    case T4K_L_CONV:    _bconv(in, out); break;
    case T4K_L_DCONV:   _bdconv(in, out); break;
    case T4K_L_LINEAR:  if (last_layer) in = out; else _blinear(in, out); break;
    case T4K_L_FLATTEN: in = out; break;
    case T4K_L_RELU: case T4K_L_TANH: case T4K_L_SELU: case T4K_L_LEAKYRL: case T4K_L_ELU:
    case T4K_L_DROPOUT: _bactivate(in, out); break;
    case T4K_L_SIGMOID: case T4K_L_SOFTMAX: case T4K_L_LOGSMAX: in = out; break;   // pass-through, also for hidden sigmoids (backprop.cu:129-131)
    case T4K_L_MAXPOOL: case T4K_L_AVGPOOL: case T4K_L_MINPOOL: _bpool(in, out, fn); break;
    case T4K_L_BATCHNM: _bbatchnorm(in, out); break;
    case T4K_L_USAMPLE: _bupsample(in, out, fn); break;
    default: Runtime::error("nn#bstep layer=%d not supported\n", fn);
    }
}
int Model::_bconv(Tensor &in, Tensor &out) {                              // backprop.cu:153-191
    Tensor &f = *in.grad[0], &df = *in.grad[2], &db = *in.grad[3], &dx = *in.grad[4];
    int rc = t4k_conv2d_bwd(in.data, out.data, f.data, dx.data, df.data, db.data, in.N(), in.H(), in.W(), in.C(),
                            out.H(), out.W(), out.C(), f.H(), in.stride[0], in.stride[2], train, ST);
    if (rc == T4K_ENOSUP) { Runtime::error("nn#bconv kernel_size=%d stride=%d padding=%d not supported\n", f.H(), in.stride[0], in.stride[2]); return -1; }
    KCHK(rc);
    in = dx;                                                               // x = dX (overwrite)
    return 0;
}
int Model::_blinear(Tensor &in, Tensor &out, bool skip_db, Tensor *xdup, bool defer) { // backprop.cu:194-254
    Tensor &w = *in.grad[0], &dw = *in.grad[2], &db = *in.grad[3];
    if (xdup && train) {
        // dW[E0,E1] += dY^T @ X (X read from its duplicate) on the side stream  ||  dX = dY @ W in place on the library stream
        const int N = (int)in.N(), E0 = (int)out.HWC(), E1 = (int)in.HWC();
        cudaStream_t st = (cudaStream_t)ST;
        if (!skip_db) KCHK(t4k_dbias(out.data, db.data, N, E0, ST));
        if (!_hp.on) {
            // dX and dW in ONE launch of the layer GEMM (X read from its duplicate): the conv block that follows writes the flatten backward itself
            const int rcp = t4k_linear_bwd_pair(xdup->data, w.data, out.data, in.data, dw.data, N, E0, E1, ST);
            if (rcp == 0) return 0;
            if (rcp != T4K_ENOSUP) KCHK(rcp);
        }
        cudaEventRecord(g_fork, st); cudaStreamWaitEvent(g_stream2, g_fork, 0);
        const bool hp = _hp.on; _hp.on = false;
        if (hp) {
            // the deferred head backward (see _bfused_head) goes to the side stream: parameter gradients of the head, dB of this layer and the
            // layer tensors' gradient values — nothing the library stream waits for before the end of the step
            KCHK(t4k_mlp_head_bwd(_hp.P, _hp.T, _hp.Ylin, _hp.X2, _hp.F1, _hp.Y1, _hp.W2, _hp.dW2, _hp.dB2, _hp.dB1, _hp.N, _hp.E0, _hp.E1, train, (t4k_stream_t)g_stream2));
            cudaEventRecord(g_head, g_stream2);
            cudaEventRecord(g_join, g_stream2);
            // dX and dW of this layer in ONE launch on the library stream, their dY operand generated from the head's forward tensors
            const int rcg = t4k_linear_bwd_from_head(_pdup, _hp.T, _hp.W2, _hp.F1, xdup->data, w.data, in.data, dw.data, N, _hp.E0, E0, E1, ST);
            if (rcg == 0) { _side_join = true; return 0; }                 // the conv block that follows writes the flatten backward itself
            if (rcg != T4K_ENOSUP) KCHK(rcg);
        }
        t4k_set_workspace_bank(WS_BANK_SIDE);
        KCHK(t4k_gemm(out.data, xdup->data, dw.data, 1.0f, 1.0f, 1, 0, E0, E1, N, 1, 1, 0, 0, 0, (t4k_stream_t)g_stream2));
        t4k_set_workspace_bank(WS_BANK_MAIN);
        cudaEventRecord(g_join, g_stream2);
        int rcx = T4K_ENOSUP;
        if (hp) rcx = t4k_linear_dx_from_head(_pdup, _hp.T, _hp.W2, _hp.F1, w.data, in.data, N, _hp.E0, E0, E1, ST);
        if (rcx == T4K_ENOSUP) {
            if (hp) cudaStreamWaitEvent(st, g_head, 0);      // not taken: dY comes from the head kernel after all
            KCHK(t4k_gemm(out.data, w.data, in.data, 1.0f, 0.0f, 0, 0, N, E1, E0, 1, 1, 0, 0, 0, ST));
        } else KCHK(rcx);
        if (!defer) { cudaStreamWaitEvent(st, g_join, 0); return 0; }
        // deferred join: the flatten backward (duplicate <- dX) follows dW on the side stream, once dX is there
        cudaEventRecord(g_fork, st); cudaStreamWaitEvent(g_stream2, g_fork, 0);
        KCHK(t4k_copy(in.data, xdup->data, in.numel, (t4k_stream_t)g_stream2));
        cudaEventRecord(g_join, g_stream2);
        _side_join = true; _skip_flat_copy = true;
        return 0;
    }
    // dX overwrites the layer input in place; dW needs X first → t4k_linear_bwd orders dW before dX,
    // and dX = dY@W does not read X, so in.data may be both X and dX.
    KCHK(t4k_linear_bwd_ex(in.data, w.data, out.data, in.data, dw.data, db.data, in.N(), (int)out.HWC(), (int)in.HWC(), train, skip_db, ST));
    return 0;
}
int Model::_blinear_act(Tensor &in, Tensor &out, Tensor &act_in, bool skip_db) {     // backprop.cu:194-263
    Tensor &w = *in.grad[0], &dw = *in.grad[2], &db = *in.grad[3];
    KCHK(t4k_linear_bwd_act(in.data, w.data, out.data, in.data, dw.data, db.data, act_in.grad[4]->data, act_in.data,
                            in.N(), (int)out.HWC(), (int)in.HWC(), train, skip_db, ST));
    return 0;
}
int Model::_bactivate(Tensor &in, Tensor &out) { KCHK(t4k_activate_bwd(out.data, in.grad[4]->data, in.data, in.numel, ST)); return 0; } // backprop.cu:257-263
int Model::_bpool(Tensor &in, Tensor &out, t4_layer fn) {                 // backprop.cu:266-280
    int rc = t4k_pool_bwd(fn, in.data, out.data, out.N(), in.H(), in.W(), out.H(), out.W(), out.C(), in.stride[0], ST);
    if (rc == T4K_ENOSUP) { Runtime::error("nn#bpool kernel_size=%d not supported\n", in.stride[0]); return -1; }
    KCHK(rc); return 0;
}
int Model::_bupsample(Tensor &in, Tensor &out, t4_layer fn) {             // backprop.cu:285-300 (k_pool USAMPLE = Σ/K²)
    (void)fn;
    KCHK(t4k_pool_fwd(T4K_L_USAMPLE, out.data, in.data, in.N(), out.H(), out.W(), in.H(), in.W(), in.C(), in.stride[0], ST)); return 0;
}
int Model::_bbatchnorm(Tensor &in, Tensor &out) {                         // backprop.cu:312-370
    if (_comm_stat && _dp_world > 1) {
        KCHK(t4k_batchnorm_bwd_dp((t4k_comm_t)_comm_stat, out.data, in.grad[4]->data, in.data, in.grad[0]->data, in.grad[2]->data, in.grad[3]->data,
                                  in.mtum[4]->data, in.N(), in.N() * _dp_world, in.H() * in.W(), in.C(), train, ST));
        return 0;
    }
    KCHK(t4k_batchnorm_bwd(out.data, in.grad[4]->data, in.data, in.grad[0]->data, in.grad[2]->data, in.grad[3]->data,
                           in.mtum[4]->data, in.N(), in.H() * in.W(), in.C(), train, ST));
    return 0;
}
// ---- loss / onehot / hit (loss.cpp:16-136)
Tensor &Model::onehot() {
    if (_hot) return *_hot;
    Runtime::error("Model.onehot not provided by dataset, use nn.onehot= to setup!\n");
    return (*this)[-1];
}
Tensor &Model::onehot(Tensor &t) {                                        // loss.cpp:26-43
    Tensor &out = (*this)[-1];
    if (_hot) { if (_own_hot) Tensor::destroy(*_hot); _hot = nullptr; }
    else if (t.N() != out.N() || (U32)t.HWC() != (U32)out.HWC()) { Runtime::error("Model.onehot dimension is not [%d,1,%d,1]\n", out.N(), (U32)out.HWC()); return t; }
    _hot = &t; _own_hot = false;
    _hit = hit(true);
    return *_hot;
}
Model &Model::forward(Dataset &ds) {                                      // forward.cu:29-78: a Dataset input also refreshes onehot and hit (:72-75)
    forward((Tensor&)ds);
    onehot_labels(ds.label);
    Tensor &out = (*this)[-1];
    if (!_cnt_dev) _cnt_dev = (int*)Runtime::alloc(256);
    KCHK(t4k_hit(out.data, _hot->data, out.N(), (int)out.HWC(), _cnt_dev, ST));       // stays on the device until `nn.hit` asks
    _hit_dev = true;
    return *this;
}
int Model::step_graph(Dataset &ds, t4_loss lop, DU *loss_dev, t4_optimizer op, DU lr, DU b1, DU b2, DU wd) {
    // one iteration of `ds for forward loss.ce backprop nn.adam next` (the loop fetches the next mini-batch, src/vm/eforth.cpp:614-634):
    // normalise the staged U8 batch into the dataset tensor + one-hot its labels (one launch, inside the captured step), then the step
    Tensor &out = (*this)[-1];
    if (!_hot) { _hot = &Tensor::create(out.N(), 1, (U32)out.HWC(), 1); _own_hot = true; _hot->zeros(); }
    StepExtra x; x.ds = &ds;
    int rc = ds.commit_begin(&x.simg, &x.slab, &x.n);
    if (rc) return rc;
    rc = _step_graph((Tensor&)ds, *_hot, lop, loss_dev, op, lr, b1, b2, wd, x);
    ds.commit_end();
    return rc;
}
int Model::train_step(Dataset &ds, t4_loss lop, DU *loss_dev, t4_optimizer op, DU lr, DU b1, DU b2, DU wd, DU *prev_loss) {
    if (!loss_dev) return T4K_EINVAL;
    if (!_loss_pin) {
        if (cudaHostAlloc((void**)&_loss_pin, 64, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return T4K_ENOMEM; }
        for (int b = 0; b < 2; b++) { cudaEvent_t e; cudaEventCreateWithFlags(&e, cudaEventDisableTiming); _loss_ev[b] = e; }
    }
    Tensor &out = (*this)[-1];
    if (!_hot) { _hot = &Tensor::create(out.N(), 1, (U32)out.HWC(), 1); _own_hot = true; _hot->zeros(); }
    const unsigned b = _tstep & 1;
    StepExtra x; x.ds = &ds; x.loss_pin = &_loss_pin[b];                  // the loss D2H rides in the captured step too
    int rc = ds.commit_begin(&x.simg, &x.slab, &x.n);
    if (rc) return rc;
    rc = _step_graph((Tensor&)ds, *_hot, lop, loss_dev, op, lr, b1, b2, wd, x);
    ds.commit_end();
    if (rc) return rc;
    cudaEventRecord((cudaEvent_t)_loss_ev[b], (cudaStream_t)ST);
    if (prev_loss) {
        if (_tstep >= 1) { cudaEventSynchronize((cudaEvent_t)_loss_ev[b ^ 1]); *prev_loss = _loss_pin[b ^ 1]; }
        else *prev_loss = NAN;
    }
    _tstep++;
    return 0;
}
int Model::train_flush(DU *last_loss) {
    if (!_tstep || !_loss_pin) return T4K_EINVAL;
    const unsigned b = (_tstep - 1) & 1;
    cudaError_t e = cudaEventSynchronize((cudaEvent_t)_loss_ev[b]);
    if (last_loss) *last_loss = _loss_pin[b];
    return (int)e;
}
Tensor &Model::onehot_labels(const int32_t *labels_dev) {                 // loss.cpp:47-72, on device
    Tensor &out = (*this)[-1];
    if (!_hot) { _hot = &Tensor::create(out.N(), 1, (U32)out.HWC(), 1); _own_hot = true; }
    KCHK(t4k_onehot(labels_dev, _hot->data, out.N(), (int)out.HWC(), ST));
    return *_hot;
}
int Model::hit(bool recalc) {                                             // loss.cpp:75-107
    if (!recalc && _hit_dev) {                                            // counted on the device by forward(Dataset&)
        int cnt = 0; cudaMemcpyAsync(&cnt, _cnt_dev, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)ST); Runtime::sync();
        _hit_dev = false; return _hit = cnt;
    }
    if (!recalc) return _hit;
    if (!_hot) return 0;
    Tensor &out = (*this)[-1];
    if (!_cnt_dev) _cnt_dev = (int*)Runtime::alloc(256);
    KCHK(t4k_hit(out.data, _hot->data, out.N(), (int)out.HWC(), _cnt_dev, ST));
    int cnt = 0; cudaMemcpyAsync(&cnt, _cnt_dev, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)ST); Runtime::sync();
    return _hit = cnt;
}
DU Model::loss(t4_loss op) { return _hot ? loss(op, *_hot) : 0.0f; }
DU Model::loss(t4_loss op, Tensor &tgt) {                                 // loss.cpp:119-136 (non-destructive: no _loss duplicate needed)
    Tensor &out = (*this)[-1];
    if (out.numel != tgt.numel) {
        Runtime::error("nn::loss model output shape[%d,%d,%d,%d] != tgt[%d,%d,%d,%d]\n",
                       out.N(), out.H(), out.W(), out.C(), tgt.N(), tgt.H(), tgt.W(), tgt.C());
        return 0.0f;
    }
    return out.loss(op, tgt);
}
int Model::loss_async(t4_loss op, Tensor &tgt, DU *loss_dev) {
    Tensor &out = (*this)[-1];
    if (out.numel != tgt.numel) return T4K_EINVAL;
    return t4k_loss(op, out.data, tgt.data, out.numel, out.N(), loss_dev, ST);
}
// ---- optimizers (gradient.cu:20-169)
Model &Model::grad_alloc(t4_optimizer op) {
    // Reference: per-tensor m / v tensors (gradient.cu:20-59).  Here: every (w,dw),(b,db) pair moves into flat
    // arenas G / DG (+ M, V, zero filled) at identical offsets, the Tensor objects become views; one
    // t4k_optim_multi launch then updates the whole model and DG is one contiguous all-reduce payload.
    struct Seg { Tensor *g, *dg; int Nw; int layer; };
    std::vector<Seg> segs;
    for (size_t i = 0; i + 1 < _layers.size(); i++) {
        Tensor &in = *_layers[i];
        // Nw = parameter tensor's N() (gradient.cu:137); conv-transpose: the reference's filter is T4(C1,K,K,C0), N() = the layer's input channels
        if (in.grad[0] && in.grad[2]) segs.push_back({in.grad[0], in.grad[2], (int)(in.grad_fn == T4K_L_DCONV ? in.C() : in.grad[0]->N()), (int)i});
        if (in.grad[1] && in.grad[3]) segs.push_back({in.grad[1], in.grad[3], (int)in.grad[1]->N(), (int)i});
    }
    _arena_opt = op;
    if (segs.empty()) return *this;
    std::vector<t4k_seg_t> table;
    U64 off = 0;
    _second_layer = (int)_layers.size(); _first_end = 0;
    for (auto &s : segs) {
        U64 len = (s.g->numel + 3) & ~3ull; table.push_back({(int64_t)off, (int64_t)len, s.Nw, 0}); off += len;
        if (s.layer == segs[0].layer) _first_end = (int64_t)off;
        else if (s.layer < _second_layer) _second_layer = s.layer;
    }
    _total = off; _nseg = (int)segs.size();
    _G = (DU*)Runtime::alloc(_total * 4); _DG = (DU*)Runtime::alloc(_total * 4);
    _M = (DU*)Runtime::alloc(_total * 4); _V  = (DU*)Runtime::alloc(_total * 4);
    _seg_dev = Runtime::alloc(4096 + 256 + sizeof(t4k_seg_t) * table.size());
    cudaStream_t st = (cudaStream_t)ST;
    cudaMemsetAsync(_G, 0, _total * 4, st); cudaMemsetAsync(_DG, 0, _total * 4, st);
    cudaMemsetAsync(_M, 0, _total * 4, st); cudaMemsetAsync(_V, 0, _total * 4, st);
    cudaMemcpyAsync(_seg_dev, table.data(), sizeof(t4k_seg_t) * table.size(), cudaMemcpyHostToDevice, st);
    for (size_t k = 0; k < segs.size(); k++) {
        Tensor *g = segs[k].g, *dg = segs[k].dg;
        KCHK(t4k_copy(g->data, _G + table[k].off, g->numel, ST)); KCHK(t4k_copy(dg->data, _DG + table[k].off, dg->numel, ST));
        if (g->owns) Runtime::free(g->data);
        if (dg->owns) Runtime::free(dg->data);
        g->data = _G + table[k].off;  g->owns = false;  g->_tmp = (DU*)_seg_dev + 1024;     // views share a scratch past the segment table
        dg->data = _DG + table[k].off; dg->owns = false; dg->_tmp = (DU*)_seg_dev + 1024;
    }
    Runtime::sync();                      // `table` is host memory read by the async copy
    return *this;
}
Model &Model::_gradient(t4_optimizer op, DU lr, DU b1, DU b2, DU wd) {     // gradient.cu:64-126
    if (_iter++ == 0 && epoch == 0 && !_G) grad_alloc(op);
    if (!train || !_G) return *this;
    const int kind = (op == OPTI_SGD || op == OPTI_SGDM) ? 0 : (op == OPTI_ADAM ? 1 : 2);
    if (_comm && _dpo.rs && !_dpo.rest) {
        // chunks went to their owners (reduce-scatter push) but no fused conv block picked the rest exchange up: it runs here, in front of the first chunks'
        if (_dp_join) { cudaStreamWaitEvent((cudaStream_t)ST, g_join, 0); _dp_join = false; }
        KCHK(t4k_optim_multi_dp_rs((t4k_comm_t)_comm, kind, _G, _DG, _M, _V, (const t4k_seg_t*)_seg_dev, _nseg, _dp_pushed_from, (int64_t)_total, lr, b1, b2, wd, dp_rs_phased() ? 2 : 0, ST));
        _dpo.rest = true;
    }
    if (_comm && _dpo.rest) {
        // the side stream exchanged and stepped everything from `_dp_pushed_from` on (Model::_dp_push): the first chunks are what is left
        _dp_join = false;                                                   // ordered by backprop's wait on the push event; the rest exchange is joined at the end of the step
        KCHK(t4k_optim_multi_dp_range((t4k_comm_t)_comm, kind, _G, _DG, _M, _V, (const t4k_seg_t*)_seg_dev, _nseg, 0, _dp_pushed_from, (int64_t)_total, lr, b1, b2, wd,
                                      _dp_scal, _dp_nscal, _dp_pushed_from, ST));
        _dp_pushed_from = -1; _dpo = DpOptEarly(); _dp_step++;
    }
    else if (_comm) {
        if (_dp_join) { cudaStreamWaitEvent((cudaStream_t)ST, g_join, 0); _dp_join = false; }     // the early push of this step (side stream)
        KCHK(t4k_optim_multi_dp((t4k_comm_t)_comm, kind, _G, _DG, _M, _V, (const t4k_seg_t*)_seg_dev, _nseg, (int64_t)_total, lr, b1, b2, wd,
                                _dp_scal, _dp_nscal, _dp_pushed_from > 0 ? _dp_pushed_from : (int64_t)_total, ST));
        _dp_pushed_from = -1; _dp_step++;
    }
    else if (_oe.on && _oe.rest && _oe.late) {                              // deferred and never picked up by a fused first block: the whole arena here
        KCHK(t4k_optim_multi(kind, _G, _DG, _M, _V, (const t4k_seg_t*)_seg_dev, _nseg, (int64_t)_total, lr, b1, b2, wd, ST));
    }
    else if (_oe.on && _oe.rest) {                                          // the side stream stepped [first_end, total) during backprop (_opt_push)
        if (!_oe.first) KCHK(t4k_optim_multi_range(kind, _G, _DG, _M, _V, (const t4k_seg_t*)_seg_dev, _nseg, 0, _first_end, lr, b1, b2, wd, ST));
    }
    else       KCHK(t4k_optim_multi(kind, _G, _DG, _M, _V, (const t4k_seg_t*)_seg_dev, _nseg, (int64_t)_total, lr, b1, b2, wd, ST));
    _oe = OptEarly();
    return *this;
}
Model &Model::sgd(DU lr, DU b) {                                           // gradient.cu:133-143: momentum forced to 0 on the first call
    DU beta = _iter ? b : 0.0f;
    return _gradient(fabsf(b) < DU_EPS_H ? OPTI_SGD : OPTI_SGDM, lr, beta, 0.0f, 0.0f);
}
Model &Model::adam(DU lr, DU b1, DU b2)         { return _gradient(OPTI_ADAM, lr, b1, b2, 0.0f); }     // gradient.cu:145-157 (no bias correction)
Model &Model::adamw(DU lr, DU wd, DU b1, DU b2) { return _gradient(OPTI_ADAMW, lr, b1, b2, wd); }      // gradient.cu:159-169
// ---- persistence: the reference's model file (src/io/aio_model.cpp)
const char *Model::nname(int fn) {
    static const char *name[] = { "output ", "conv2d ", "linear ", "flatten", "relu   ", "tanh   ", "sigmoid", "selu   ", "leakyrl", "elu    ",
                                  "dropout", "softmax", "logsmax", "avgpool", "maxpool", "minpool", "batchnm", "upsampl", "dconv2d" };
    return (fn >= 0 && fn < 19) ? name[fn] : "unknown";
}
static std::string layer_parm(Tensor &in, Tensor &out) {                  // AIO::_parm, aio_model.cpp:102-141 (ostream << float formatting)
    std::ostringstream o;
    const int S = in.stride[0]; const DU p = in.xparm;
    switch (in.grad_fn) {
    case T4K_L_CONV: case T4K_L_DCONV: o << "bias=" << p << ", C=" << out.C() << ", K=" << in.grad[0]->H() << ", S=" << S << ", P=" << in.stride[2]; break;
    case T4K_L_LINEAR:  o << "bias=" << p << ", H=" << in.grad[0]->H(); break;
    case T4K_L_SELU: case T4K_L_LEAKYRL: case T4K_L_ELU: o << "bias=" << p; break;
    case T4K_L_DROPOUT: o << "rate=" << p * 100.0 << '%'; break;
    case T4K_L_AVGPOOL: case T4K_L_MAXPOOL: case T4K_L_MINPOOL: o << S << "x" << S; break;
    case T4K_L_BATCHNM: o << "mtum=" << p; break;
    case T4K_L_USAMPLE: { const char *nm[] = { "nearest", "linear", "bilinear", "cubic" }; o << S << "x" << S << " " << nm[in.iparm & 3]; } break;
    default: break;
    }
    return o.str();
}
// Optimizer state for resume (the reference's file has none: a reloaded model restarts Adam from m = v = 0).  It FOLLOWS the
// reference's closing "---" line, where the reference's reader has already stopped (AIO::_nload_param reads one section per parametrised layer),
// so a file with state still loads in the reference:
//   "\ optimizer state iter=<calls> epoch=<n> floats=<arena length>\n--- m.arena\n" <raw FP32> "\n--- v.arena\n" <raw FP32> "\n---\n"
// m / v are the flat moment arenas (grad_alloc: every parameter segment at its offset, 4-float aligned).
static const char *OPT_TAG = "\\ optimizer state";
int Model::save(const char *fname, bool opt_state) {                       // AIO::nsave + _nsave_model + _nsave_param
    std::ofstream fs(fname, std::ios_base::binary);
    if (!fs.is_open()) { Runtime::error("} => failed to open for output\n"); return 1; }
    fs << "\\ tensorForth v4.0 model\n";
    const int n = (int)_layers.size();
    for (int i = 0; i < n - 1; i++) fs << layer_parm(*_layers[i], *_layers[i + 1]) << nname(_layers[i]->grad_fn) << std::endl;
    std::vector<DU> h;
    auto dump = [&](char pn, const char *nm, Tensor &t) {
        fs << "\n--- " << pn << "." << nm << std::endl;
        h.resize(t.numel); t.d2h(h.data());
        fs.write((const char*)h.data(), t.numel * sizeof(DU));
    };
    for (int i = 0; i < n - 1; i++) {
        Tensor &in = *_layers[i];
        switch (in.grad_fn) {
        case T4K_L_CONV: case T4K_L_DCONV: case T4K_L_LINEAR: dump('w', nname(in.grad_fn), *in.grad[0]); dump('b', nname(in.grad_fn), *in.grad[1]); break;
        case T4K_L_BATCHNM: dump('w', nname(in.grad_fn), *in.grad[0]); break;
        default: break;
        }
    }
    fs << "\n---" << std::endl;
    if (opt_state && _G && _M && _V) {
        h.resize(_total);
        fs << OPT_TAG << " iter=" << _iter << " epoch=" << epoch << " floats=" << _total << "\n--- m.arena\n";
        Runtime::sync();
        cudaMemcpy(h.data(), _M, _total * sizeof(DU), cudaMemcpyDeviceToHost); fs.write((const char*)h.data(), _total * sizeof(DU));
        fs << "\n--- v.arena\n";
        cudaMemcpy(h.data(), _V, _total * sizeof(DU), cudaMemcpyDeviceToHost); fs.write((const char*)h.data(), _total * sizeof(DU));
        fs << "\n---" << std::endl;
    }
    return fs.good() ? 0 : 1;
}
int Model::load(const char *fname) {                                       // AIO::nload (parameter path) + _nload_param
    std::ifstream fs(fname, std::ios_base::binary);
    if (!fs.is_open()) { Runtime::error("} => failed to open for input\n"); return 1; }
    std::string line;
    while (std::getline(fs, line) && line.length()) {}                     // skip the model section (up to the blank line)
    std::vector<DU> h;
    int err = 0;
    auto read = [&](Tensor &t) {
        while (std::getline(fs, line) && !line.length()) {}                // skip blank lines
        if (line.size() < 3 || line[0] != '-' || line[1] != '-' || line[2] != '-') { Runtime::error(" model format error"); err = 1; return; }
        h.resize(t.numel);
        fs.read((char*)h.data(), t.numel * sizeof(DU));
        if ((U64)fs.gcount() != t.numel * sizeof(DU)) { Runtime::error(" model format error"); err = 1; return; }
        t.h2d(h.data()); Runtime::sync();                                  // `h` is reused by the next section
    };
    const int n = (int)_layers.size();
    for (int i = 0; i < n - 1 && !err; i++) {
        Tensor &in = *_layers[i];
        switch (in.grad_fn) {
        case T4K_L_CONV: case T4K_L_DCONV: case T4K_L_LINEAR: read(*in.grad[0]); if (!err) read(*in.grad[1]); break;
        case T4K_L_BATCHNM: read(*in.grad[0]); break;
        default: break;
        }
    }
    if (err) return err;
    // optional optimizer state behind the reference's sections (see save)
    while (std::getline(fs, line)) {
        if (line.compare(0, strlen(OPT_TAG), OPT_TAG) != 0) continue;
        long it = 0, ep = 0; unsigned long long n = 0;
        if (sscanf(line.c_str() + strlen(OPT_TAG), " iter=%ld epoch=%ld floats=%llu", &it, &ep, &n) != 3) { Runtime::error(" model format error (optimizer state)"); return 1; }
        if (!_G) grad_alloc(OPTI_ADAM);
        if (n != _total) { Runtime::error(" optimizer state of %llu floats does not fit this model (%llu)", n, (unsigned long long)_total); return 1; }
        h.resize(_total);
        for (DU *dst : {_M, _V}) {
            if (!std::getline(fs, line) || line.compare(0, 3, "---") != 0) { Runtime::error(" model format error (optimizer state)"); return 1; }
            fs.read((char*)h.data(), _total * sizeof(DU));
            if ((U64)fs.gcount() != _total * sizeof(DU)) { Runtime::error(" model format error (optimizer state)"); return 1; }
            cudaMemcpy(dst, h.data(), _total * sizeof(DU), cudaMemcpyHostToDevice);
            std::getline(fs, line);                                        // the newline in front of the next "---"
        }
        _iter = (int)it; epoch = (int)ep;
        _drop_graphs();                                                    // captured steps baked the first-call SGD momentum / arena pointers of the old state
        break;
    }
    return 0;
}
int Model::arena(DU **G, DU **DG, int64_t *total) {
    if (!_G) grad_alloc(OPTI_ADAM);
    if (G) *G = _G;
    if (DG) *DG = _DG;
    if (total) *total = (int64_t)_total;
    return _G ? 0 : T4K_EINVAL;
}
int Model::bn_channels() {
    int c = 0;
    for (Tensor *t : _layers) if (t->grad_fn == T4K_L_BATCHNM) c = std::max(c, (int)t->C());
    return c;
}
int Model::dp_shard(int rank, int world, void *comm_stat) {
    if (world < 1 || rank < 0 || rank >= world) return T4K_EINVAL;
    const int C = bn_channels();
    if (comm_stat && t4k_comm_capacity((t4k_comm_t)comm_stat) < 4 * (int64_t)C) return T4K_EINVAL;
    _dp_rank = rank; _dp_world = world; _comm_stat = world > 1 ? comm_stat : nullptr;
    Runtime::sync(); _drop_graphs();
    return 0;
}
int Model::dp_attach(void *comm, DU *scal, int nscal) {
    if (!_G) grad_alloc(OPTI_ADAM);
    if (comm && (!_G || t4k_comm_capacity((t4k_comm_t)comm) < (int64_t)_total || nscal < 0 || nscal > 64)) return T4K_EINVAL;
    if (comm && bn_channels() && !(_comm_stat && _dp_world > 1)) {
        // per-shard batch statistics would silently train another model than the single-device step (nmath.cu:177-264,295-414 over the whole batch)
        Runtime::error("nn#dp_attach: the model has batchnorm layers; give it a statistics communicator first (Model::dp_shard)\n");
        return T4K_ENOSUP;
    }
    _comm = comm; _dp_scal = scal; _dp_nscal = comm ? nscal : 0; _dp_step = 0;     // a fresh communicator: no exchange completed yet
    Runtime::sync(); _drop_graphs();                          // the captured optimizer node changes
    return 0;
}
// ---- one train step as a CUDA graph: forward + loss + backprop + optimizer
int Model::step_graph(Tensor &input, Tensor &tgt, t4_loss lop, DU *loss_dev, t4_optimizer op, DU lr, DU b1, DU b2, DU wd) {
    return _step_graph(input, tgt, lop, loss_dev, op, lr, b1, b2, wd, StepExtra());
}
void Model::_drop_graphs() {
    for (auto &g : _graphs) if (g.exec) { cudaGraphExecDestroy((cudaGraphExec_t)g.exec); g.exec = nullptr; }
}
int Model::_step_graph(Tensor &input, Tensor &tgt, t4_loss lop, DU *loss_dev, t4_optimizer op, DU lr, DU b1, DU b2, DU wd, const StepExtra &x) {
    bool has_dropout = false;
    for (Tensor *t : _layers) if (t->grad_fn == T4K_L_DROPOUT) has_dropout = true;
    auto run = [&]() {
        if (has_dropout) t4k_rand_tick(ST);                    // replayed graphs draw a fresh dropout mask every step (rand.cu)
        if (x.ds) { _feed = &x; _feed_hot = tgt.data; }        // dataset feeding (normalise + one-hot): rides in the first fused block, else its own launch (forward)
        _want_pdup = loss_dev != nullptr; _pdup_valid = false;
        _fwd_tgt = (train && fuse && tgt.numel == (*this)[-1].numel) ? &tgt : nullptr;     // backprop(tgt) follows at once: the forward tail may do the head's backward
        forward(input);
        _fwd_tgt = nullptr;
        _want_pdup = false;
        // the loss does not feed the backward pass: when the head kernel left a duplicate of the softmax output (backprop overwrites
        // the original with p - y), the loss kernel — and its read-back — run on the side stream under the backward kernels
        const bool side_loss = loss_dev && _pdup_valid && tgt.numel == (*this)[-1].numel;
        if (side_loss) {
            cudaEventRecord(g_fork, (cudaStream_t)ST); cudaStreamWaitEvent(g_stream2, g_fork, 0);
            KCHK(t4k_loss(lop, _pdup, tgt.data, (int64_t)tgt.numel, (int)(*this)[-1].N(), loss_dev, (t4k_stream_t)g_stream2));
            if (x.loss_pin && !_comm) cudaMemcpyAsync(x.loss_pin, loss_dev, sizeof(DU), cudaMemcpyDeviceToHost, g_stream2);
            cudaEventRecord(g_join, g_stream2);
            _side_join = true;
        }
        else if (loss_dev) loss_async(lop, tgt, loss_dev);
        // loss read-back: the 4-byte D2H takes the copy engine ~6 us — on a branch of its own (side stream) it overlaps backprop
        // instead of delaying the next step.  With a communicator attached the loss is only final after the exchange (the ranks'
        // loss sums ride in it), so there it stays at the end of the step.
        const bool early_loss = x.loss_pin && loss_dev && !_comm && !side_loss;
        if (early_loss) {
            cudaEventRecord(g_fork, (cudaStream_t)ST); cudaStreamWaitEvent(g_stream2, g_fork, 0);
            cudaMemcpyAsync(x.loss_pin, loss_dev, sizeof(DU), cudaMemcpyDeviceToHost, g_stream2);
            cudaEventRecord(g_join, g_stream2);
        }
        _dp_early = _comm && (int)op >= 0 && train && _G;  // forward -> backprop -> optimizer is one unit here: the exchange may start early
        _dpo = DpOptEarly();
        if (_dp_early && _iter > 0 && fuse) {              // ... and so may the optimizer of everything but the first chunks (see _dp_push); not on the first call (SGD forces its momentum to 0 there)
            _dpo.on = true; _dpo.lr = lr; _dpo.b2 = b2; _dpo.wd = 0.0f;
            if (op == OPTI_SGD || op == OPTI_SGDM) { _dpo.kind = 0; _dpo.b1 = fabsf(b1) < DU_EPS_H ? 0.0f : b1; _dpo.b2 = 0.0f; }
            else if (op == OPTI_ADAM) { _dpo.kind = 1; _dpo.b1 = b1; }
            else { _dpo.kind = 2; _dpo.b1 = b1; _dpo.wd = wd; }
        }
        _oe = OptEarly();
        if (!_comm && (int)op >= 0 && train && _G && _iter > 0 && fuse) {     // single GPU: the optimizer may start early too (see OptEarly)
            _oe.on = true; _oe.lr = lr; _oe.b2 = b2; _oe.wd = 0.0f;
            if (op == OPTI_SGD || op == OPTI_SGDM) { _oe.kind = 0; _oe.b1 = fabsf(b1) < DU_EPS_H ? 0.0f : b1; _oe.b2 = 0.0f; }
            else if (op == OPTI_ADAM) { _oe.kind = 1; _oe.b1 = b1; }
            else { _oe.kind = 2; _oe.b1 = b1; _oe.wd = wd; }
        }
        backprop(tgt);
        _dp_early = false;
        // data parallel with a pinned loss slot: the exchange kernel itself stores the summed loss there (no copy node behind the step)
        const bool mirror = _comm && (int)op >= 0 && x.loss_pin && loss_dev && _dp_scal == loss_dev && _dp_nscal >= 1 && dp_mirror_on();
        if (_comm) t4k_comm_scalar_mirror((t4k_comm_t)_comm, mirror ? x.loss_pin : nullptr);
        if ((int)op >= 0) {                                // op < 0 — data parallel over NCCL: the caller all-reduces DG, then calls the optimizer
            switch (op) { case OPTI_SGD: case OPTI_SGDM: sgd(lr, b1); break; case OPTI_ADAM: adam(lr, b1, b2); break; default: adamw(lr, wd, b1, b2); }
        }
        if (_side_join) { cudaStreamWaitEvent((cudaStream_t)ST, g_join, 0); _side_join = false; }   // side-stream work of this step (backprop joins its own)
        if (early_loss) cudaStreamWaitEvent((cudaStream_t)ST, g_join, 0);
        else if (x.loss_pin && loss_dev && !(side_loss && !_comm) && !mirror) cudaMemcpyAsync(x.loss_pin, loss_dev, sizeof(DU), cudaMemcpyDeviceToHost, (cudaStream_t)ST);
        if (_comm && mirror) t4k_comm_scalar_mirror((t4k_comm_t)_comm, nullptr);
    };
    if (loss_dev && !_pdup) _pdup = (DU*)Runtime::alloc((size_t)(*this)[-1].numel * sizeof(DU) + 64);   // not inside the capture below
    if (!_hscratch && fuse && train) {                    // train tail: linear -> mask activation -> small linear -> softmax at the end of the model
        const int n = (int)_layers.size();
        if (n >= 5 && _layers[n - 5]->grad_fn == T4K_L_LINEAR && mask_act(_layers[n - 4]->grad_fn) && _layers[n - 3]->grad_fn == T4K_L_LINEAR &&
            _layers[n - 2]->grad_fn == T4K_L_SOFTMAX) {
            const int64_t nf = t4k_head_train_scratch_floats(_layers[n - 4]->grad_fn, (int)_layers[n - 4]->N(), (int)_layers[n - 4]->HWC(),
                                                             (int)_layers[n - 5]->HWC(), (int)_layers[n - 2]->HWC());
            if (nf > 0) _hscratch = (DU*)Runtime::alloc((size_t)nf * sizeof(DU) + 64);
        }
    }
    U64 key[13] = {0}; float f4[4] = {lr, b1, b2, wd};
    key[12] = _comm ? (U64)(_dp_step & 1u) : 0;           // data parallel: the early push's copy nodes address the exchange slots of one parity (t4k_dp_push_dma)
    key[0] = (U64)input.data; key[1] = (U64)tgt.data; key[2] = (U64)lop; key[3] = (U64)loss_dev; key[4] = (U64)op;
    memcpy(&key[5], f4, 16); key[7] = (U64)train; key[8] = (U64)x.simg; key[9] = (U64)x.slab; key[10] = (U64)x.n; key[11] = (U64)x.loss_pin;
    // SGD's first call forces momentum 0 (host state) and the first optimizer call builds the arenas: run those eagerly
    if (!_G || (_iter == 0 && (int)op >= 0)) { run(); return 0; }
    cudaStream_t st = (cudaStream_t)ST;
    StepGraph *slot = nullptr, *lru = nullptr;        // cached graph for this key, else an empty slot, else the least recently used
    for (auto &g : _graphs) {
        if (g.exec && memcmp(key, g.key, sizeof(key)) == 0) { slot = &g; break; }
        if (!lru || (!g.exec && lru->exec) || (!!g.exec == !!lru->exec && g.used < lru->used)) lru = &g;
    }
    if (!slot) {
        slot = lru;
        if (slot->exec) { cudaGraphExecDestroy((cudaGraphExec_t)slot->exec); slot->exec = nullptr; }
        cudaGraph_t graph = nullptr;
        if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); run(); return 0; }
        const int it = _iter; const uint32_t ds = _dp_step;
        run();
        cudaError_t e = cudaStreamEndCapture(st, &graph);
        if (e != cudaSuccess || !graph) { cudaGetLastError(); _iter = it; _dp_step = ds; Runtime::error("graph capture failed: %s", cudaGetErrorString(e)); run(); return 0; }
        cudaGraphExec_t ex = nullptr;
        e = cudaGraphInstantiate(&ex, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) { cudaGetLastError(); _iter = it; _dp_step = ds; run(); return 0; }
        slot->exec = ex; memcpy(slot->key, key, sizeof(key));
        _iter = it; _dp_step = ds;                    // the captured run did not execute; the launch below is the step
    }
    slot->used = ++_graph_clock;
    if ((int)op >= 0) { _iter++; if (_comm) _dp_step++; }
    return (int)cudaGraphLaunch((cudaGraphExec_t)slot->exec, st);
}

} // namespace t4

// =============================================================================== flat C interface
using namespace t4;
#define TT(x) (*(Tensor*)(x))
#define MM(x) (*(Model*)(x))
extern "C" {
const char *t4h_last_error(void) { return Runtime::last_error(); }
int   t4h_init(int device) { return Runtime::init(device); }
void *t4h_stream(void) { return Runtime::stream(); }
void *t4h_side_stream(void) { Runtime::stream(); return (void*)g_stream2; }
int   t4h_sync(void) { return Runtime::sync(); }
long  t4h_launch_count(void) { return t4k_launch_count(); }

t4h_tensor t4h_tensor_new(int rank, uint32_t n, uint32_t h, uint32_t w, uint32_t c) {
    if (Runtime::init(0)) return nullptr;
    Tensor *t = rank == 1 ? &Tensor::create((U64)h) : rank == 2 ? &Tensor::create(h, w) : &Tensor::create(n, h, w, c);
    if (!t->data) { delete t; return nullptr; }
    return (t4h_tensor)t;
}
void  t4h_tensor_free(t4h_tensor t) { if (t) Tensor::destroy(TT(t)); }
float *t4h_tensor_data(t4h_tensor t) { return TT(t).data; }
int   t4h_tensor_shape(t4h_tensor t, uint32_t nhwc[4]) { nhwc[0] = TT(t).N(); nhwc[1] = TT(t).H(); nhwc[2] = TT(t).W(); nhwc[3] = TT(t).C(); return 0; }
int64_t t4h_tensor_numel(t4h_tensor t) { return (int64_t)TT(t).numel; }
int   t4h_tensor_rank(t4h_tensor t) { return (int)TT(t).rank; }
int   t4h_tensor_h2d(t4h_tensor t, const float *src, int64_t n) { return TT(t).h2d(src, n); }
int   t4h_tensor_d2h(t4h_tensor t, float *dst, int64_t n) { return TT(t).d2h(dst, n); }
int   t4h_tensor_reshape(t4h_tensor t, int rank, uint32_t n, uint32_t h, uint32_t w, uint32_t c) {
    Tensor &x = TT(t); const U64 want = rank == 1 ? h : rank == 2 ? (U64)h * w : (U64)n * h * w * c;
    if (want != x.numel) { if (rank == 1) x.reshape((U64)h); else if (rank == 2) x.reshape(h, w); else x.reshape(n, h, w, c); return T4K_EINVAL; }
    if (rank == 1) x.reshape((U64)h); else if (rank == 2) x.reshape(h, w); else x.reshape(n, h, w, c);
    return 0;
}
t4h_tensor t4h_tensor_copy(t4h_tensor t) { return (t4h_tensor)&Tensor::copy_of(TT(t)); }
int   t4h_tensor_map(t4h_tensor t, int op, float v) { return t4k_map(op, TT(t).data, v, TT(t).numel, Runtime::stream()); }
int   t4h_tensor_identity(t4h_tensor t) { TT(t).identity(); return 0; }
int   t4h_tensor_rand(t4h_tensor t, int opt) { return t4k_rand(TT(t).data, TT(t).numel, opt, 0.0f, 1.0f, Runtime::stream()); }
int   t4h_ten_op_s(int op, t4h_tensor A, float v, t4h_tensor O) { return t4k_ts_op(op, TT(A).data, v, TT(O).data, TT(A).numel, Runtime::stream()); }
int   t4h_ten_op_t(int op, t4h_tensor A, t4h_tensor B, t4h_tensor O) {
    Tensor &a = TT(A), &b = TT(B), &o = TT(O);
    if (a.HWC() != b.HWC() || (a.N() == 1 ? b.numel : a.numel) != o.numel) { Tensor::ten_op((math_op)op, a, b, o); return T4K_EINVAL; }
    return t4k_tt_op(op, a.data, b.data, o.data, a.HWC(), a.N(), b.N(), Runtime::stream());
}
int   t4h_mm(t4h_tensor A, t4h_tensor B, t4h_tensor O, int inc, int tA, int tB) { Tensor::mm(TT(A), TT(B), TT(O), inc, tA, tB); return 0; }
int   t4h_gemm(int variant, t4h_tensor A, t4h_tensor B, t4h_tensor O, float alpha, float beta, int tA, int tB) {
    switch (variant) { case 1: Tensor::gemm1(TT(A), TT(B), TT(O), alpha, beta, tA, tB); break; case 2: Tensor::gemm2(TT(A), TT(B), TT(O), alpha, beta, tA, tB); break;
                       case 4: Tensor::gemm4(TT(A), TT(B), TT(O), alpha, beta, tA, tB); break; default: Tensor::gemm3(TT(A), TT(B), TT(O), alpha, beta, tA, tB); }
    return 0;
}
t4h_tensor t4h_matmul(t4h_tensor A_, t4h_tensor B_) {                     // TensorVM::_tdot, tenvm.cpp:329-367
    Tensor &A = TT(A_), &B = TT(B_);
    U32 Na = A.N(), Ha = A.H(), Wa = A.W(), Ca = A.C(), Nb = B.N(), Hb = B.H(), Wb = B.W(), Cb = B.C();
    if (B.rank == 1 && A.rank != 1 && Wa == B.numel) { Tensor &C = Tensor::create((U64)Ha); Tensor::mm(A, B, C); return (t4h_tensor)&C; }
    if (A.rank == 2 && B.rank == 2 && Wa == Hb)       { Tensor &C = Tensor::create(Ha, Wb); Tensor::mm(A, B, C); return (t4h_tensor)&C; }
    if ((Na == 1 || Nb == 1) && Na != Nb && Ca == Cb && Wa == Hb) {
        Tensor &C = Tensor::create(std::max(Na, Nb), Ha, Wb, Ca); Tensor::mm(A, B, C); return (t4h_tensor)&C;
    }
    Runtime::error("A.W != B.H dim?");
    return nullptr;
}
t4h_tensor t4h_transpose(t4h_tensor A_) {                                 // blas1(T_XPOS), tenvm.cpp:128-130
    Tensor &A = TT(A_); Tensor &T = Tensor::copy_of(A);
    if (A.rank == 2) T.reshape(A.W(), A.H()); else T.reshape(A.N(), A.W(), A.H(), A.C());
    Tensor::transpose(A, T); return (t4h_tensor)&T;
}
float t4h_tensor_sum(t4h_tensor t)  { return TT(t).sum(); }
float t4h_tensor_avg(t4h_tensor t)  { return TT(t).avg(); }
float t4h_tensor_std(t4h_tensor t)  { return TT(t).std(); }
float t4h_tensor_norm(t4h_tensor t) { return TT(t).norm(); }
float t4h_tensor_max(t4h_tensor t)  { return TT(t).max(); }
float t4h_tensor_min(t4h_tensor t)  { return TT(t).min(); }
float t4h_tensor_dot(t4h_tensor A, t4h_tensor B) { return TT(A).dot(TT(B)); }
float t4h_tensor_loss(t4h_tensor o, int op, t4h_tensor tgt) { return TT(o).loss((t4_loss)op, TT(tgt)); }

t4h_model t4h_model_new(uint32_t n, uint32_t h, uint32_t w, uint32_t c) { if (Runtime::init(0)) return nullptr; return (t4h_model) new Model(n, h, w, c); }
void  t4h_model_free(t4h_model m) { delete (Model*)m; }
int   t4h_model_add(t4h_model m, int layer, uint32_t n, float bias, const uint16_t *opt) {
    U16 o[4]; if (opt) memcpy(o, opt, sizeof(o));
    const U64 before = MM(m).numel();
    MM(m).add((t4_layer)layer, n, bias, opt ? o : nullptr);
    return MM(m).numel() > before ? 0 : T4K_ENOSUP;
}
int   t4h_model_numel(t4h_model m) { return (int)MM(m).numel(); }
t4h_tensor t4h_model_layer(t4h_model m, int i) {
    const int n = (int)MM(m).numel(); const int k = i < 0 ? n + i : i;
    return (k < 0 || k >= n) ? nullptr : (t4h_tensor)&MM(m)[i];
}
t4h_tensor t4h_model_param(t4h_model m, int i, int which) {              // NetVM::_get_parm, netvm.cpp:153-166
    t4h_tensor L = t4h_model_layer(m, i);
    if (!L || which < 0 || which > 4) return nullptr;
    Tensor &t = TT(L);
    Tensor *p = which ? t.grad[which] : (t.grad[0] ? t.grad[0] : t.grad[4]);
    return (t4h_tensor)p;
}
int   t4h_model_set_param(t4h_model m, int i, int which, t4h_tensor src) {   // NetVM::_set_parm, netvm.cpp:171-193
    Tensor *p = (Tensor*)t4h_model_param(m, i, which);
    if (!p || !src || TT(src).numel != p->numel) { Runtime::error("Tensor and model parameter is not the same shape"); return T4K_EINVAL; }
    Tensor::copy(TT(src), *p);
    return 0;
}
int   t4h_model_train(t4h_model m, int on) { MM(m).train = on != 0; return 0; }
int   t4h_model_fuse(t4h_model m, int on) { MM(m).fuse = on != 0; return 0; }
int   t4h_model_forward(t4h_model m, t4h_tensor input) {
    Tensor &n0 = MM(m)[0];
    if (TT(input).numel != n0.numel) { MM(m).forward(TT(input)); return T4K_EINVAL; }
    MM(m).forward(TT(input)); return 0;
}
int   t4h_model_backprop(t4h_model m, t4h_tensor tgt) {
    if (!tgt) { MM(m).backprop(); return 0; }
    if (TT(tgt).numel != MM(m)[-1].numel) { MM(m).backprop(TT(tgt)); return T4K_EINVAL; }
    MM(m).backprop(TT(tgt)); return 0;
}
float t4h_model_loss(t4h_model m, int op, t4h_tensor tgt) { return tgt ? MM(m).loss((t4_loss)op, TT(tgt)) : MM(m).loss((t4_loss)op); }
int   t4h_model_loss_async(t4h_model m, int op, t4h_tensor tgt, float *loss_dev) { return MM(m).loss_async((t4_loss)op, TT(tgt), loss_dev); }
t4h_dataset t4h_dataset_create(int n, int h, int w, int c) { return (t4h_dataset)&Dataset::create(n, h, w, c); }
void  t4h_dataset_destroy(t4h_dataset d) { if (d) Dataset::destroy(*(Dataset*)d); }
void  t4h_dataset_normalize(t4h_dataset d, float mean, float scale) { ((Dataset*)d)->normalize(mean, scale); }
int   t4h_dataset_stage(t4h_dataset d, const uint8_t *img_host, const uint8_t *lab_host, int n) { return ((Dataset*)d)->stage(img_host, lab_host, n); }
int   t4h_dataset_commit(t4h_dataset d) { return ((Dataset*)d)->commit(); }
t4h_tensor t4h_dataset_tensor(t4h_dataset d) { return (t4h_tensor)(Tensor*)(Dataset*)d; }
const int32_t *t4h_dataset_labels(t4h_dataset d) { return ((Dataset*)d)->label; }
int   t4h_model_forward_ds(t4h_model m, t4h_dataset d) { MM(m).forward(*(Dataset*)d); return 0; }
int   t4h_model_step_graph_ds(t4h_model m, t4h_dataset d, int lop, float *loss_dev, int optimizer, float lr, float b1, float b2, float wd) {
    return MM(m).step_graph(*(Dataset*)d, (t4_loss)lop, loss_dev, (t4_optimizer)optimizer, lr, b1, b2, wd);
}
int   t4h_model_train_step_ds(t4h_model m, t4h_dataset d, int lop, float *loss_dev, int optimizer, float lr, float b1, float b2, float wd, float *prev_loss) {
    return MM(m).train_step(*(Dataset*)d, (t4_loss)lop, loss_dev, (t4_optimizer)optimizer, lr, b1, b2, wd, prev_loss);
}
int   t4h_model_train_flush(t4h_model m, float *last_loss) { return MM(m).train_flush(last_loss); }
int   t4h_model_onehot_labels(t4h_model m, const int32_t *labels_dev) { MM(m).onehot_labels(labels_dev); return 0; }
int   t4h_model_onehot_set(t4h_model m, t4h_tensor hot) { MM(m).onehot(TT(hot)); return 0; }
int   t4h_model_hit(t4h_model m, int recalc) { return MM(m).hit(recalc != 0); }
int   t4h_model_sgd(t4h_model m, float lr, float b) { MM(m).sgd(lr, b); return 0; }
int   t4h_model_adam(t4h_model m, float lr, float b1, float b2) { MM(m).adam(lr, b1, b2); return 0; }
int   t4h_model_adamw(t4h_model m, float lr, float wd, float b1, float b2) { MM(m).adamw(lr, wd, b1, b2); return 0; }
/* ---- generic capture: whatever the library enqueues on its stream between begin and end becomes a replayable CUDA graph (several
 * models, tensor words, draws ...).  No host reads in between (they synchronise); warm the sequence up once first so that every
 * workspace and arena exists.  The graph starts with t4k_rand_tick: replays draw fresh random numbers. */
int   t4h_capture_begin(void) {
    cudaStream_t st = (cudaStream_t)Runtime::stream();
    cudaError_t e = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
    if (e != cudaSuccess) { cudaGetLastError(); return (int)e; }
    return t4k_rand_tick((t4k_stream_t)st);
}
void *t4h_capture_end(void) {
    cudaGraph_t g = nullptr;
    if (cudaStreamEndCapture((cudaStream_t)Runtime::stream(), &g) != cudaSuccess || !g) { cudaGetLastError(); Runtime::error("graph capture failed"); return nullptr; }
    cudaGraphExec_t ex = nullptr;
    cudaError_t e = cudaGraphInstantiate(&ex, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) { cudaGetLastError(); Runtime::error("graph instantiation failed: %s", cudaGetErrorString(e)); return nullptr; }
    return (void*)ex;
}
int   t4h_graph_launch(void *g) { return g ? (int)cudaGraphLaunch((cudaGraphExec_t)g, (cudaStream_t)Runtime::stream()) : T4K_EINVAL; }
void  t4h_graph_free(void *g) { if (g) { Runtime::sync(); cudaGraphExecDestroy((cudaGraphExec_t)g); } }
int   t4h_model_save(t4h_model m, const char *fname) { return MM(m).save(fname); }
int   t4h_model_save_state(t4h_model m, const char *fname) { return MM(m).save(fname, true); }
int   t4h_model_load(t4h_model m, const char *fname) { return MM(m).load(fname); }
int   t4h_model_arena(t4h_model m, float **G, float **DG, int64_t *total) { return MM(m).arena(G, DG, total); }
int   t4h_model_dp_attach(t4h_model m, void *comm, float *scal, int nscal) { return MM(m).dp_attach(comm, scal, nscal); }
int   t4h_model_dp_shard(t4h_model m, int rank, int world, void *comm_stat) { return MM(m).dp_shard(rank, world, comm_stat); }
int   t4h_model_bn_channels(t4h_model m) { return MM(m).bn_channels(); }
int   t4h_use_lane(int lane) { return Runtime::use_lane(lane); }
int   t4h_set_dp_rest(int on) { const int was = dp_rest_on(); g_dp_rest = on ? 1 : 0; return was; }
int   t4h_set_dp_early(int mode) { const int was = g_dp_early; g_dp_early = (mode < 0 || mode > 3) ? 3 : mode; return was; }
int   t4h_tensor_rand_sharded(t4h_tensor t, int opt, int rank, int world) {
    if (world < 1 || rank < 0 || rank >= world) return T4K_EINVAL;
    const int64_t n = (int64_t)TT(t).numel;
    return t4k_rand_sharded(TT(t).data, n, rank * n, world * n, opt, 0.0f, 1.0f, Runtime::stream());
}
int   t4h_model_step_graph(t4h_model m, t4h_tensor input, t4h_tensor tgt, int lop, float *loss_dev, int optimizer, float lr, float b1, float b2, float wd) {
    return MM(m).step_graph(TT(input), TT(tgt), (t4_loss)lop, loss_dev, (t4_optimizer)optimizer, lr, b1, b2, wd);
}
} // extern "C"
