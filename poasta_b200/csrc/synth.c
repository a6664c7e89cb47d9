/* Seeded synthetic inputs for BASELINE.json's configs.
 * PRNG: xoshiro256** seeded through splitmix64. Alphabet ACGT uniform. Each root base is kept with
 * probability 1 - rate, else substituted / preceded by an inserted base / deleted with probability 1/3
 * each. Built as poasta_b200/libsynth.so; used by bench.py and the tests (not by the aligner). */
#include <stdint.h>
#include <stdio.h>
#include <string.h>

typedef struct { uint64_t s[4]; } rng_t;

static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }

static void rng_seed(rng_t *r, uint64_t seed) {
    uint64_t x = seed;
    for (int i = 0; i < 4; ++i) {
        x += 0x9E3779B97F4A7C15ull;
        uint64_t z = x;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        r->s[i] = z ^ (z >> 31);
    }
}

static uint64_t rng_next(rng_t *r) {
    uint64_t *s = r->s;
    uint64_t result = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0];
    s[3] ^= s[1];
    s[1] ^= s[2];
    s[0] ^= s[3];
    s[2] ^= t;
    s[3] = rotl(s[3], 45);
    return result;
}

static double rng_uniform(rng_t *r) { return (double)(rng_next(r) >> 11) * (1.0 / 9007199254740992.0); }

static const char ACGT[4] = {'A', 'C', 'G', 'T'};

static size_t mutate(rng_t *r, const char *src, size_t n, double rate, char *dst) {
    size_t k = 0;
    for (size_t i = 0; i < n; ++i) {
        if (rng_uniform(r) < rate) {
            uint64_t kind = rng_next(r) % 3;
            if (kind == 0) {
                char c;
                do c = ACGT[rng_next(r) & 3]; while (c == src[i]);
                dst[k++] = c;
            } else if (kind == 1) {
                dst[k++] = ACGT[rng_next(r) & 3];
                dst[k++] = src[i];
            }
        } else {
            dst[k++] = src[i];
        }
    }
    return k;
}

/* Writes one family as FASTA text (records ">{prefix}_s{k}") into out (capacity cap); returns bytes written
 * or 0 if cap is too small. Members are independent mutations of one random root (star phylogeny). */
size_t synth_family(uint64_t seed, uint32_t n_seqs, uint32_t root_len, double rate, const char *prefix, char *out, size_t cap) {
    rng_t r;
    rng_seed(&r, seed);
    char root[65536], buf[2 * 65536];
    if (root_len > sizeof root) return 0;
    for (uint32_t i = 0; i < root_len; ++i) root[i] = ACGT[rng_next(&r) & 3];
    size_t pos = 0;
    for (uint32_t k = 0; k < n_seqs; ++k) {
        size_t n = mutate(&r, root, root_len, rate, buf);
        if (pos + n + strlen(prefix) + 32 > cap) return 0;
        pos += (size_t)sprintf(out + pos, ">%s_s%u\n", prefix, k);
        memcpy(out + pos, buf, n);
        pos += n;
        out[pos++] = '\n';
    }
    return pos;
}

/* Config 5: haplotypes (mutations of one root) and reads sampled from them. */
size_t synth_haplotypes(uint64_t seed, uint32_t n_haps, uint32_t root_len, double rate, char *out, size_t cap) {
    return synth_family(seed, n_haps, root_len, rate, "hap", out, cap);
}

size_t synth_read(uint64_t seed, const char *hap, uint32_t hap_len, uint32_t read_len, double rate, char *out, size_t cap) {
    rng_t r;
    rng_seed(&r, seed);
    uint32_t span = hap_len > read_len ? hap_len - read_len : 1;
    uint32_t start = (uint32_t)(rng_next(&r) % span);
    uint32_t n = hap_len - start < read_len ? hap_len - start : read_len;
    if (cap < 2 * (size_t)n + 2) return 0;
    return mutate(&r, hap + start, n, rate, out);
}
