/*
 * standin_sketcher.c — TEST INFRASTRUCTURE, not part of the engine.
 *
 * A stand-in for the step BEFORE the distance path (pp_sketchlib.constructDatabase, reached from
 * PopPUNK/sketchlib.py:348-435), so that BASELINE config 1 — the reference's own smoke test on the 29 assemblies of
 * test/example_set.tar.bz2 (test/run_test.py:20-21, test/references.txt) — has real-genome input.  pp-sketchlib's
 * source is not available here, so its hash values cannot be reproduced: the sketches this tool writes have the
 * reference's SCHEMA (sketchsize64*14 uint64 words per k, bindash bit-sliced, see include/ppb.h) and the published
 * construction (BinDash one-permutation b-bit MinHash, citation.py:35-38: canonical k-mers, one hash, S equal-width
 * bins over the hash range, the minimum per bin, empty bins filled from the next non-empty bin, the low 14 bits
 * kept), but are NOT bit-compatible with a real PopPUNK database.  They pin nothing about pp-sketchlib; they give
 * the parity tests and tools/perf_shapes.py a realistic Jaccard spectrum from real genomes.
 *
 * usage: standin_sketcher <fasta> <sketchsize64> <k1,k2,...>   -> sketchsize64*14 uint64 words per k on stdout (binary)
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define BBITS 14

static inline uint64_t mix64(uint64_t x) { /* splitmix64 finaliser */
    x += 0x9e3779b97f4a7c15ull;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}

static int code_of(int c) {
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return -1;
    }
}

int main(int argc, char **argv) {
    if (argc != 4) {
        fprintf(stderr, "usage: %s fasta sketchsize64 k1,k2,...\n", argv[0]);
        return 2;
    }
    const int ss64 = atoi(argv[2]);
    const uint64_t S = 64ull * (uint64_t)ss64;
    int ks[64], nk = 0;
    for (char *tok = strtok(argv[3], ","); tok && nk < 64; tok = strtok(NULL, ",")) ks[nk++] = atoi(tok);
    FILE *f = fopen(argv[1], "rb");
    if (!f) {
        perror(argv[1]);
        return 1;
    }
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    char *buf = (char *)malloc((size_t)sz + 1);
    if (fread(buf, 1, (size_t)sz, f) != (size_t)sz) return 1;
    fclose(f);
    uint64_t *mins = (uint64_t *)malloc(sizeof(uint64_t) * S);
    uint64_t *words = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)ss64 * BBITS);
    for (int t = 0; t < nk; t++) {
        const int k = ks[t];
        if (k < 3 || k > 31) return 2;
        const uint64_t mask = (1ull << (2 * k)) - 1;
        for (uint64_t b = 0; b < S; b++) mins[b] = UINT64_MAX;
        uint64_t fw = 0, rv = 0;
        int run = 0, in_header = 0;
        for (long p = 0; p < sz; p++) {
            const char c = buf[p];
            if (c == '>') { in_header = 1; run = 0; continue; }
            if (in_header) { if (c == '\n') in_header = 0; continue; }
            if (c == '\n' || c == '\r') continue;
            const int code = code_of(c);
            if (code < 0) { run = 0; continue; }
            fw = ((fw << 2) | (uint64_t)code) & mask;
            rv = (rv >> 2) | ((uint64_t)(3 - code) << (2 * (k - 1)));
            if (++run >= k) {
                const uint64_t h = mix64(fw < rv ? fw : rv); /* canonical k-mer */
                const uint64_t bin = (uint64_t)(((unsigned __int128)h * S) >> 64);
                if (h < mins[bin]) mins[bin] = h;
            }
        }
        /* densification: an empty bin takes the value of the next non-empty bin (cyclically) */
        uint64_t first = S;
        for (uint64_t b = 0; b < S; b++) if (mins[b] != UINT64_MAX) { first = b; break; }
        if (first == S) { fprintf(stderr, "no k-mers of length %d in %s\n", k, argv[1]); return 1; }
        uint64_t next = mins[first];
        for (uint64_t d = 0; d < S; d++) {
            const uint64_t b = (first + S - d) % S;   /* walk backwards from `first` so `next` is the next non-empty */
            if (mins[b] != UINT64_MAX) next = mins[b]; else mins[b] = next;
        }
        /* bit-slice: word [s*14 + b] holds bit b of the signatures of bins 64s .. 64s+63 */
        memset(words, 0, sizeof(uint64_t) * (size_t)ss64 * BBITS);
        for (uint64_t bin = 0; bin < S; bin++) {
            const uint64_t sig = mins[bin] & ((1u << BBITS) - 1);
            for (int b = 0; b < BBITS; b++)
                if ((sig >> b) & 1) words[(bin / 64) * BBITS + b] |= 1ull << (bin % 64);
        }
        fwrite(words, sizeof(uint64_t), (size_t)ss64 * BBITS, stdout);
    }
    return 0;
}
