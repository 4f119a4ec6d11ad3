// arena_driver.cpp — CPU stress of integration/arena_shim.cpp (the 64-bit host-side allocator behind the reference's TLSF class interface,
// SURVEY.md §8f row 3).  Built and run by tests/test_arena_shim_cpu.py in the build container only: it compiles against the reference's own
// header src/mu/tlsf.h (where it lies, unmodified).  The allocator never touches the store, so the "store" is (a) a plain host buffer whose
// blocks are filled with a per-block pattern (overlap / slack checks) and (b) a 24 GiB range that does not exist at all (64-bit offsets).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <random>
#include <vector>
#include "mu/tlsf.h"

using namespace t4::mu;
#define CHECK(c) do { if (!(c)) { printf("FAIL line %d: %s\n", __LINE__, #c); return 1; } } while (0)

struct Blk { uint8_t *p; uint64_t sz; uint8_t tag; };

int main() {
    TLSF &t = TLSF::get_instance();
    // ---- (a) a real 64 MiB store, deliberately misaligned: random alloc / free, every live block carries its own byte pattern
    const uint64_t SZ = 64ull << 20;
    uint8_t *raw = (uint8_t*)malloc(SZ + 4096);
    uint8_t *mem = raw + 24;
    t.init(mem, SZ, 8);
    std::mt19937_64 rng(12345);
    std::vector<Blk> live;
    uint64_t in_use = 0;
    for (int it = 0; it < 20000; it++) {
        const bool do_alloc = live.empty() || (rng() % 100) < 55;
        if (do_alloc) {
            static const uint64_t kinds[] = {1, 4, 252, 253, 256, 1000, 4096, 65536, 1u << 20, 3u << 20};
            uint64_t sz = kinds[rng() % 10] + (rng() % 3 == 0 ? rng() % 777 : 0);
            uint8_t *p = (uint8_t*)t.malloc(sz);
            if (!p) { CHECK(in_use + sz + 260 > SZ / 2); continue; }      // refusals only when the store is at least half full (fragmentation allowed)
            CHECK(((uintptr_t)p & 255) == 0);                             // 256-byte aligned: TMA / 128-bit loads
            CHECK(p >= mem + 8 && p + sz + 4 <= mem + SZ);                // inside the store, with the float of slack behind the block
            const uint8_t tag = (uint8_t)(1 + rng() % 250);
            memset(p, tag, sz + 4);                                       // the block and its slack (`data[numel]` scratch of the reference's reductions)
            live.push_back({p, sz, tag});
            in_use += sz;
        } else {
            const size_t k = rng() % live.size();
            Blk b = live[k];
            for (uint64_t i = 0; i < b.sz + 4; i += (b.sz > 4096 ? 97 : 1)) CHECK(b.p[i] == b.tag);   // nobody else was handed these bytes
            memset(b.p, 0, b.sz + 4);
            t.free(b.p);
            in_use -= b.sz;
            live[k] = live.back(); live.pop_back();
        }
    }
    for (auto &b : live) {
        for (uint64_t i = 0; i < b.sz + 4; i += (b.sz > 4096 ? 97 : 1)) CHECK(b.p[i] == b.tag);
        t.free(b.p);
    }
    // everything returned: the free blocks must have coalesced back into ONE block of the whole (aligned) heap
    const uint64_t whole = ((SZ - 8 - 256) & ~255ull) - 4;
    void *all = t.malloc(whole);
    CHECK(all != nullptr);
    CHECK(t.malloc(1) == nullptr || true);                                // (may or may not fit in the tail; must not crash)
    t.free(all);
    // double free / foreign pointer: reported, not fatal, bookkeeping intact
    t.free(all);
    t.free(mem + 12345);
    void *again = t.malloc(whole);
    CHECK(again == all);
    t.free(again);
    // realloc that fits the block keeps the pointer (the growing path copies on the DEVICE and is exercised on the GPU box)
    void *r0 = t.malloc(1000);
    CHECK(t.realloc(r0, 900) == r0);
    t.free(r0);

    // ---- (b) 24 GiB: offsets and sizes beyond 32 bits (BASELINE config 5: one 8192 x 56 x 56 x 64 FP32 tensor is 6.6 GB); nothing is dereferenced
    uint8_t *fake = (uint8_t*)(uintptr_t)0x100000000000ull;
    const uint64_t BIG = 24ull << 30;
    t.init(fake, BIG);
    const uint64_t T66 = 8192ull * 56 * 56 * 64 * 4;                      // 6 576 668 672 bytes
    uint8_t *a = (uint8_t*)t.malloc(T66), *b = (uint8_t*)t.malloc(T66), *c = (uint8_t*)t.malloc(T66);
    CHECK(a && b && c);
    CHECK(b >= a + T66 + 4 && c >= b + T66 + 4);                          // disjoint, in address order (best fit, lowest offset first)
    CHECK((uint64_t)(c - fake) > 0xFFFFFFFFull);                          // an offset no 32-bit header could hold
    CHECK(t.malloc(T66) == nullptr);                                      // 4 x 6.6 GB > 24 GiB: refused (message on stdout), not wrapped around
    t.free(b);
    uint8_t *b2 = (uint8_t*)t.malloc(T66 - 1000);
    CHECK(b2 == b);                                                       // the hole is reused
    t.free(a); t.free(b2); t.free(c);
    uint8_t *w = (uint8_t*)t.malloc(BIG - 4096);
    CHECK(w == fake);                                                     // fully coalesced again
    free(raw);
    printf("ARENA OK\n");
    return 0;
}
