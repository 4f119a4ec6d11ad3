// runtime.cu — library state: device attributes, workspace, launch accounting, error strings
#include "common.cuh"
#include <mutex>
#include <vector>
#include <cstdio>
#include <cstdlib>

namespace t4k {

long g_launches = 0;
int  g_carve = []{ const char *e = getenv("T4K_CARVEOUT"); const int v = e ? atoi(e) : 100; return v > 100 ? 100 : v; }();   // common.cuh: shared-memory split asked for by the short kernels
int  g_pdl = []{ const char *e = getenv("T4K_PDL"); return (e && e[0] == '1') ? 1 : 0; }();   // opt-in: no net gain measured on the MNIST step (common.cuh)

int cur_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return -1; }
    return dev;
}
int sm_count() {
    static int n[16];
    const int dev = cur_device();
    if (dev < 0 || dev >= 16) return T4K_SMS;
    if (n[dev] == 0) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) { cudaGetLastError(); return T4K_SMS; }
        n[dev] = v;
    }
    return n[dev];
}

int check_launch() {
    ++g_launches;
    cudaError_t e = cudaGetLastError();
    return (int)e;
}

// Library-owned scratch, one growing buffer per (device, slot).  Ownership rule of the boundary
// (SURVEY.md §8b): user-visible tensors belong to the caller's arena, workspace lives here.
// Regrowth RETIRES the old block instead of freeing it: its address may be baked into cached CUDA graphs (Model::step_graph,
// th.Graph) that a later, larger call must not invalidate — those graphs keep replaying on the block they were captured with (it is
// large enough for them).  Growth is geometric, so the retired blocks of a slot add up to less than 4x its final size; they are
// released at process exit.  A regrowth attempted INSIDE a stream capture fails (cudaMalloc is not capturable): the call returns
// T4K_ENOMEM, the capture is abandoned by the caller (Model::_step_graph falls back to the eager step, which sizes the workspace).
#define MAX_DEV  16
#define MAX_SLOT 64          // 8 slots x 8 banks (odd banks: work forked onto a side stream; banks 2k, 2k+1: lane k of the host runtime, see t4k_set_workspace_bank)
static thread_local int g_ws_bank = 0;       // per host thread (the host runtime gives every lane, and each lane's side stream, its own bank)
static void  *g_ws[MAX_DEV][MAX_SLOT];
static size_t g_ws_sz[MAX_DEV][MAX_SLOT];
static std::mutex g_mu;
static std::vector<void*> g_retired;                  // outgrown blocks: possibly referenced by cached graphs, never freed while the process runs

void *workspace(size_t bytes, int slot) {
    int dev = 0;
    slot += 8 * g_ws_bank;
    if (cudaGetDevice(&dev) != cudaSuccess || dev >= MAX_DEV || slot >= MAX_SLOT) return nullptr;
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_ws_sz[dev][slot] < bytes) {
        size_t want = bytes + (bytes >> 2);
        want = (want + 255) & ~(size_t)255;
        void *p = nullptr;
        if (cudaMalloc(&p, want) != cudaSuccess) { cudaGetLastError(); return nullptr; }      // the old block (if any) stays valid for its users
        if (g_ws[dev][slot]) g_retired.push_back(g_ws[dev][slot]);
        g_ws[dev][slot] = p; g_ws_sz[dev][slot] = want;
    }
    return g_ws[dev][slot];
}

// Reduction scratch: ring of 64 slots x (1024 partial floats + 8 control words), control words
// are zero on entry and re-zeroed by the finishing block, so no memset per call.
#define RSLOTS      64
#define RSLOT_FLTS  (2048 + 8)
static float *g_red[MAX_DEV];
static unsigned g_red_next[MAX_DEV];

float *reduce_slot(cudaStream_t) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev >= MAX_DEV) return nullptr;
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_red[dev]) {
        void *p = nullptr;
        size_t sz = (size_t)RSLOTS * RSLOT_FLTS * sizeof(float);
        if (cudaMalloc(&p, sz) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        cudaMemset(p, 0, sz);
        g_red[dev] = (float*)p;
    }
    unsigned k = g_red_next[dev]++ % RSLOTS;
    return g_red[dev] + (size_t)k * RSLOT_FLTS;
}

} // namespace t4k

extern "C" {

int t4k_version(void) { return T4K_VERSION; }

const char *t4k_strerror(int rc) {
    switch (rc) {
    case 0:          return "ok";
    case T4K_EINVAL: return "t4k: invalid argument / unsupported shape";
    case T4K_ENOSUP: return "t4k: configuration not supported (as in the reference)";
    case T4K_ENOMEM: return "t4k: workspace allocation failed";
    default:         return rc > 0 ? cudaGetErrorString((cudaError_t)rc) : "t4k: unknown error";
    }
}

int t4k_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int t4k_sm_count(void) { return t4k::sm_count(); }

int t4k_sync(t4k_stream_t s) { return (int)cudaStreamSynchronize((cudaStream_t)s); }

long t4k_launch_count(void) { return t4k::g_launches; }

int t4k_set_workspace_bank(int bank) { int was = t4k::g_ws_bank; t4k::g_ws_bank = bank & 7; return was; }

int t4k_set_carveout(int pct) { int was = t4k::g_carve; t4k::g_carve = pct > 100 ? 100 : pct; return was; }

int t4k_set_pdl(int on) { int was = t4k::g_pdl; t4k::g_pdl = on ? 1 : 0; return was; }

} // extern "C"
