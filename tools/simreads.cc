// simreads.cc -- seeded synthetic read simulator for the benchmark and the parity tests
// (SURVEY.md section 8d): random genome, uniform read starts, random strand, iid substitution
// errors, FASTQ output `@r<idx>\n<seq>\n+\n<L x 'I'>\n`.  PRNG: xoshiro256** seeded through
// splitmix64.  Not part of the product; no reference code involved.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {
struct Rng {
    uint64_t s[4];
    static uint64_t splitmix(uint64_t& x) {
        uint64_t z = (x += 0x9E3779B97F4A7C15ULL);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        return z ^ (z >> 31);
    }
    explicit Rng(uint64_t seed) { for (int i = 0; i < 4; ++i) s[i] = splitmix(seed); }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next() {
        uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
        return r;
    }
    uint64_t below(uint64_t n) { return (uint64_t)(((unsigned __int128)next() * n) >> 64); }
    double unit() { return (next() >> 11) * (1.0 / 9007199254740992.0); }
};
}  // namespace

extern "C" {

// bytes one FASTQ record of read `idx` takes
static uint64_t record_bytes(uint64_t idx, uint64_t L) {
    char tmp[32];
    int d = snprintf(tmp, sizeof tmp, "%llu", (unsigned long long)idx);
    return 2 + (uint64_t)d + 1 + L + 1 + 2 + L + 1;          // "@r" idx \n seq \n "+\n" qual \n
}

uint64_t sim_fastq_bytes(uint64_t n_reads, uint64_t L, uint64_t first_idx) {
    uint64_t total = 0;
    for (uint64_t i = 0; i < n_reads; ++i) total += record_bytes(first_idx + i, L);
    return total;
}

void sim_genome(uint64_t G, uint64_t seed, char* out) {
    Rng rng(seed);
    for (uint64_t i = 0; i < G; i += 32) {
        uint64_t r = rng.next();
        for (uint64_t j = i; j < G && j < i + 32; ++j) { out[j] = "ACGT"[r & 3]; r >>= 2; }
    }
}

// Writes n_reads FASTQ records into out (sized by sim_fastq_bytes); returns bytes written.
uint64_t sim_reads_fastq(const char* genome, uint64_t G, uint64_t L, uint64_t n_reads, double err, uint64_t seed,
                         uint64_t first_idx, char* out) {
    Rng rng(seed);
    char* p = out;
    static const char comp[256] = {0};
    (void)comp;
    for (uint64_t i = 0; i < n_reads; ++i) {
        uint64_t start = rng.below(G - L + 1);
        bool flip = rng.next() >> 63;
        p += sprintf(p, "@r%llu\n", (unsigned long long)(first_idx + i));
        for (uint64_t j = 0; j < L; ++j) {
            char b;
            if (!flip) b = genome[start + j];
            else {
                char g = genome[start + L - 1 - j];
                b = g == 'A' ? 'T' : g == 'C' ? 'G' : g == 'G' ? 'C' : 'A';
            }
            if (err > 0 && rng.unit() < err) {
                int code = b == 'A' ? 0 : b == 'C' ? 1 : b == 'G' ? 2 : 3;
                b = "ACGT"[(code + 1 + (int)rng.below(3)) & 3];
            }
            *p++ = b;
        }
        *p++ = '\n'; *p++ = '+'; *p++ = '\n';
        memset(p, 'I', L); p += L;
        *p++ = '\n';
    }
    return (uint64_t)(p - out);
}

}  // extern "C"

#ifdef SIMREADS_MAIN
int main(int argc, char** argv) {
    if (argc < 7) { fprintf(stderr, "usage: simreads G L n_reads err seed_genome seed_reads > reads.fq\n"); return 1; }
    uint64_t G = strtoull(argv[1], 0, 10), L = strtoull(argv[2], 0, 10), n = strtoull(argv[3], 0, 10);
    double e = atof(argv[4]);
    uint64_t sg = strtoull(argv[5], 0, 10), sr = strtoull(argv[6], 0, 10);
    std::vector<char> genome(G);
    sim_genome(G, sg, genome.data());
    std::vector<char> out(sim_fastq_bytes(n, L, 0));
    uint64_t w = sim_reads_fastq(genome.data(), G, L, n, e, sr, 0, out.data());
    fwrite(out.data(), 1, w, stdout);
    return 0;
}
#endif
