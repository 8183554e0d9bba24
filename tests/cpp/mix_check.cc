// Host-side check of the key mixing used by the partial-sort counting (gossamer_b200/csrc/common.cuh):
// key_unmix(key_mix(x)) == x for 64- and 128-bit keys, and single-base substitutions anywhere in a window
// change the low 32 bits of the mixed key (the property the group reduce relies on for its speed, never for
// its correctness).  Built and run by tests/test_abi_and_host.py with the host compiler only.
#include <cstdio>
#include <cstdlib>
#include "../../gossamer_b200/csrc/keys.h"

using namespace gsb;

static u64 rnd(u64& s) { s += 0x9E3779B97F4A7C15ull; u64 z = s; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }

int main() {
    u64 s = 12345;
    const u64 edge[] = {0ull, 1ull, ~0ull, 1ull << 63, 0x5555555555555555ull, 0xAAAAAAAAAAAAAAAAull};
    for (u64 x : edge) if (key_unmix(key_mix(x)) != x) { printf("u64 inverse fails for %llx\n", x); return 1; }
    u64 same_low32 = 0, trials = 0;
    for (int i = 0; i < 2000000; ++i) {
        const u64 x = rnd(s);
        if (key_unmix(key_mix(x)) != x) { printf("u64 inverse fails for %llx\n", x); return 1; }
        Key128 k; k.lo = x; k.hi = rnd(s) >> 16;
        const Key128 m = key_mix(k), b = key_unmix(m);
        if (b.lo != k.lo || b.hi != k.hi || m.hi != k.hi) { printf("Key128 inverse fails\n"); return 1; }
        if (i < 20000) {
            for (int base = 0; base < 32; ++base)
                for (u64 pat = 1; pat < 4; ++pat) {
                    const u64 y = x ^ (pat << (2 * base));
                    ++trials;
                    if ((u32)key_mix(y) == (u32)key_mix(x)) ++same_low32;
                }
        }
    }
    // chance level is trials / 2^32 (< 1 expected collision)
    if (same_low32 > 4) { printf("substitution variants collide in the low 32 bits: %llu of %llu\n", same_low32, trials); return 1; }
    printf("ok %llu %llu\n", same_low32, trials);
    return 0;
}
