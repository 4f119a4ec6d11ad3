// arena_shim.cpp — the object-store allocator of `ten4_b200` (see arena_shim.h): the reference's TLSF class INTERFACE (src/mu/tlsf.h:80-126,
// compiled against that header, unmodified) with another implementation behind it.  The reference threads its free lists through 8/16-byte
// headers inside the managed block (32-bit sizes: 2 GiB blocks at most, 4 GiB heaps; the host dirties a page of the store per call); here the
// store is opaque device-preferred memory and the bookkeeping is two ordered maps in host memory:
//   free blocks by offset (coalescing with both neighbours on free) and by size (best fit, lowest offset first), used blocks by offset.
// Every block is a multiple of 256 bytes at a 256-byte aligned address: tensors are TMA- and 128-bit-load-able (the reference's are 8-byte aligned).
#include <map>
#include <set>
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>
#include "mu/tlsf.h"

#if T4_DO_OBJ
namespace t4::mu {

namespace {
constexpr U64 ALIGN = 256;
struct Arena {
    std::map<U64, U64> free_by_off;                      // offset -> size
    std::set<std::pair<U64, U64>> free_by_size;          // (size, offset)
    std::map<U64, U64> used;                             // offset -> size
    U64 peak = 0, in_use = 0;
    void add_free(U64 off, U64 sz) { free_by_off[off] = sz; free_by_size.insert({sz, off}); }
    void del_free(std::map<U64, U64>::iterator it) { free_by_size.erase({it->second, it->first}); free_by_off.erase(it); }
} g_a;
}

__HOST__ TLSF &TLSF::get_instance() { static TLSF t; return t; }

__HOST__ void TLSF::init(U8 *mem, U64 sz, U64 off) {
    TRACE("\\ TLSF: ostore=%p, alloc=0x%lx (arena_shim: host-side bookkeeping, 64-bit, %lu-byte blocks)\n", mem, sz, (unsigned long)ALIGN);
    U64 base = ((U64)(uintptr_t)(mem + off) + ALIGN - 1) & ~(ALIGN - 1);
    _heap    = (U8*)(uintptr_t)base;
    _heap_sz = (sz - (base - (U64)(uintptr_t)mem)) & ~(ALIGN - 1);
    g_a = Arena();
    g_a.add_free(0, _heap_sz);
}

__HOST__ void *TLSF::malloc(U64 sz) {
    // one float of slack behind every block: the reference's reductions park their result in data[numel] (`_tmp`), and MMU::copy sizes the
    // copy's block to numel floats exactly (src/mu/mmu.cu:288-289) — in its own TLSF that float lands in the 8-byte alignment padding / header gap
    const U64 need = (sz + sizeof(DU) + ALIGN - 1) & ~(ALIGN - 1);
    std::lock_guard<std::mutex> lk(_mutex);
    auto it = g_a.free_by_size.lower_bound({need, 0});   // smallest block that fits, lowest offset among equals
    if (it == g_a.free_by_size.end()) { ERROR("TLSF::malloc(0x%lx) out of object store (%lu MiB in use)\n", (unsigned long)sz, (unsigned long)(g_a.in_use >> 20)); return NIL; }
    const U64 bsz = it->first, off = it->second;
    g_a.free_by_size.erase(it); g_a.free_by_off.erase(off);
    if (bsz > need) g_a.add_free(off + need, bsz - need);
    g_a.used[off] = need;
    g_a.in_use += need; if (g_a.in_use > g_a.peak) g_a.peak = g_a.in_use;
    return _heap + off;
}

__HOST__ void TLSF::free(void *ptr) {
    if (!ptr) return;
    std::lock_guard<std::mutex> lk(_mutex);
    U64 off = (U64)((U8*)ptr - _heap);
    auto u = g_a.used.find(off);
    if (u == g_a.used.end()) { ERROR("TLSF::free(%p) not an allocated block\n", ptr); return; }
    U64 sz = u->second;
    g_a.used.erase(u); g_a.in_use -= sz;
    auto nx = g_a.free_by_off.lower_bound(off);
    if (nx != g_a.free_by_off.end() && nx->first == off + sz) { sz += nx->second; g_a.del_free(nx); }
    auto pv = g_a.free_by_off.lower_bound(off);
    if (pv != g_a.free_by_off.begin()) {
        --pv;
        if (pv->first + pv->second == off) { off = pv->first; sz += pv->second; g_a.del_free(pv); }
    }
    g_a.add_free(off, sz);
}

__HOST__ void *TLSF::realloc(void *p0, U64 sz) {
    if (!p0) return malloc(sz);
    U64 old = 0;
    {
        std::lock_guard<std::mutex> lk(_mutex);
        auto u = g_a.used.find((U64)((U8*)p0 - _heap));
        if (u == g_a.used.end()) { ERROR("TLSF::realloc(%p) not an allocated block\n", p0); return NIL; }
        old = u->second;
    }
    if (sz <= old) return p0;
    void *p1 = malloc(sz);
    if (!p1) return NIL;
    cudaMemcpy(p1, p0, old, cudaMemcpyDefault);          // device-side copy: the host does not pull the pages
    free(p0);
    return p1;
}

__HOST__ int  TLSF::_mmu_ok() { return 1; }
__HOST__ void TLSF::_show_stat() {
#if T4_VERBOSE > 1                                       // as the reference: the allocator reports only in verbose builds (src/mu/tlsf.cpp:414-415)
    std::lock_guard<std::mutex> lk(_mutex);
    U64 fr = 0, big = 0;
    for (auto &b : g_a.free_by_off) { fr += b.second; if (b.second > big) big = b.second; }
    printf("\\ object store %p: %lu MiB, used %lu MiB in %zu blocks (peak %lu MiB), free %lu MiB in %zu blocks (largest %lu MiB)\n",
           (void*)_heap, (unsigned long)(_heap_sz >> 20), (unsigned long)(g_a.in_use >> 20), g_a.used.size(), (unsigned long)(g_a.peak >> 20),
           (unsigned long)(fr >> 20), g_a.free_by_off.size(), (unsigned long)(big >> 20));
#endif
}
__HOST__ void TLSF::_dump_freelist() {
#if MM_DEBUG
    for (auto &b : g_a.free_by_off) printf("\\   free %12lx +%lx\n", (unsigned long)b.first, (unsigned long)b.second);
#endif
}

} // namespace t4::mu
#endif // T4_DO_OBJ
