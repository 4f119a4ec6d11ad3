/*
 * arena_shim.h — SURVEY.md §8f row 3: the object store of the reference's VM, re-sized and re-homed for a 180 GB device.
 *
 * PRE-INCLUDED (nvcc -include) when the reference's own, unmodified src/mu/mmu.cu and src/ten4.cu are compiled for `ten4_b200`.  The reference
 * keeps every tensor of the VM in ONE cudaMallocManaged block of 2 GiB (src/ten4_config.h:67, src/mu/mmu.cu:45) carved up by a TLSF allocator
 * whose block headers live INSIDE that block (32-bit sizes and offsets, written by the host: src/mu/tlsf.h:19-31) — a BASELINE config 5 tensor
 * (8192x56x56x64 FP32 = 6.6 GB) cannot exist in it, and every allocation makes the host touch pages that the kernels then pull back.
 * Here:
 *   * T4_OSTORE_SZ is a run-time size: T4_OSTORE_GB gibibytes (environment), default 3/4 of the device's free memory — managed memory is
 *     populated on first touch, so reserving the range costs nothing;
 *   * MM_ALLOC also advises the driver that the range lives on the device (cudaMemAdviseSetPreferredLocation): pages the host touched for a
 *     `.` or a `t!` return to HBM with the next kernel and stay there;
 *   * the allocator itself is integration/arena_shim.cpp (compiled INSTEAD of src/mu/tlsf.cpp): same TLSF class interface, 64-bit sizes,
 *     bookkeeping in host memory (the store is never touched by the host on alloc/free), 256-byte aligned blocks (TMA / 128-bit loads).
 * Nothing of the reference is copied: the macros below are re-definitions of two of its configuration names.
 */
#pragma once
#include "ten4_config.h"
#include "ten4_types.h"
#include <cuda_runtime.h>
#include <cstdlib>

static inline long &t4b_ostore_ref() { static long sz = 0; return sz; }
static inline long t4b_ostore_bytes() {
    long &sz = t4b_ostore_ref();
    if (sz) return sz;
    if (const char *e = getenv("T4_OSTORE_GB")) { const double g = atof(e); if (g > 0) sz = (long)(g * 1073741824.0); }
    if (!sz) {
        size_t fr = 0, tot = 0;
        if (cudaMemGetInfo(&fr, &tot) == cudaSuccess && fr > 0) sz = (long)(fr / 4 * 3);
        else { cudaGetLastError(); sz = 2048L * 1024 * 1024; }
    }
    sz &= ~0xFFFFFL;                                  /* whole MiB */
    return sz;
}
static inline cudaError_t t4b_mm_alloc(void **p, size_t bytes) {
    /* a managed range needs backing the HOST can also provide (the box's RAM + its limits decide): on refusal the store is halved until the
     * driver accepts it — down to the reference's own 2 GiB — and T4_OSTORE_SZ reports what was obtained */
    cudaError_t e = cudaMallocManaged(p, bytes);
    while (e != cudaSuccess && bytes > (2048UL << 20)) {
        cudaGetLastError();
        bytes = (bytes / 2) & ~(size_t)0xFFFFF;
        e = cudaMallocManaged(p, bytes);
    }
    if (e != cudaSuccess) return e;
    if ((long)bytes < t4b_ostore_ref() || !t4b_ostore_ref()) t4b_ostore_ref() = (long)bytes;
    int dev = 0;
    cudaGetDevice(&dev);
    /* advice is best effort: a platform without concurrent managed access refuses it, the block still works as the reference's does */
    if (cudaMemAdvise(*p, bytes, cudaMemAdviseSetPreferredLocation, dev) != cudaSuccess) cudaGetLastError();
    return cudaSuccess;
}
template<typename T> static inline cudaError_t t4b_mm_alloc(T **p, size_t bytes) { return t4b_mm_alloc((void**)p, bytes); }

#undef  T4_OSTORE_SZ
#define T4_OSTORE_SZ   (t4b_ostore_bytes())
#undef  MM_ALLOC
#define MM_ALLOC(...)  GPU_ERR(t4b_mm_alloc(__VA_ARGS__))
