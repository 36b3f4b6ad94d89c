// rng_glibc.cuh -- exact model of glibc's srand()/rand() (TYPE_3 additive feedback generator),
// the RNG the reference actually uses (stochasticSpike.c:948 srand, :297 and :334 rand; SURVEY.md D1
// and App. C -- NOT drand48).
//
//   r[0] = seed (0 -> 1), r[i] = 16807 * r[i-1] mod 2147483647 (Schrage, int32)   for i = 1..30
//   r[31..33] = r[0..2];  r[i] = r[i-31] + r[i-3] (mod 2^32) for i >= 34;  rand() #k = r[344+k] >> 1
//
// The recurrence holds from i = 34 on, so w[j] = r[3+j] satisfies w[j] = w[j-3] + w[j-31] for all
// j >= 31 and is a linear recurrence with characteristic polynomial P(x) = x^31 - x^28 - 1 over
// Z/2^32.  Skip-ahead: w[n+t] = sum_j c_j w[j+t] with c = x^n mod P, so any block of the stream can
// be generated independently from the 61-word seed window -- that is what lets the GPU produce the
// whole stream in parallel while staying bit-identical to libc.
#pragma once
#include <stdint.h>

#define GLIBC_RAND_MAX   2147483647
#define GLIBC_CUT4       2147483644u      // (RAND_MAX / 4) * 4: randomNum(0,3) rejects r >= cutoff (stochasticSpike.c:292-299)
#define GLIBC_DEG        31

#ifdef __CUDACC__
#define RNG_HD __host__ __device__
#else
#define RNG_HD
#endif

// w[0..60] = r[3..63]
RNG_HD inline void glibc_seed_window(unsigned seed, uint32_t w[61])
{
    int32_t r[64];
    int32_t word = (int32_t)(seed ? seed : 1u);
    r[0] = word;
    for (int i = 1; i < 31; i++) {
        int32_t hi = word / 127773, lo = word % 127773;
        word = 16807 * lo - 2836 * hi;
        if (word < 0) word += 2147483647;
        r[i] = word;
    }
    r[31] = r[0]; r[32] = r[1]; r[33] = r[2];
    for (int i = 34; i < 64; i++) r[i] = (int32_t)((uint32_t)r[i - 31] + (uint32_t)r[i - 3]);
    for (int j = 0; j < 61; j++) w[j] = (uint32_t)r[3 + j];
}

// c = a * b mod P, coefficients mod 2^32
RNG_HD inline void glibc_poly_mulmod(const uint32_t a[31], const uint32_t b[31], uint32_t c[31])
{
    uint32_t t[61];
    for (int i = 0; i < 61; i++) t[i] = 0;
    for (int i = 0; i < 31; i++) {
        uint32_t ai = a[i];
        if (!ai) continue;
        for (int j = 0; j < 31; j++) t[i + j] += ai * b[j];
    }
    for (int d = 60; d >= 31; d--) {           // x^d = x^(d-3) + x^(d-31)
        uint32_t v = t[d];
        t[d - 3] += v;
        t[d - 31] += v;
    }
    for (int i = 0; i < 31; i++) c[i] = t[i];
}

// c = x^n mod P
RNG_HD inline void glibc_poly_xpow(uint64_t n, uint32_t c[31])
{
    uint32_t base[31], acc[31], tmp[31];
    for (int i = 0; i < 31; i++) { base[i] = 0; acc[i] = 0; }
    base[1] = 1; acc[0] = 1;
    while (n) {
        if (n & 1) { glibc_poly_mulmod(acc, base, tmp); for (int i = 0; i < 31; i++) acc[i] = tmp[i]; }
        n >>= 1;
        if (n) { glibc_poly_mulmod(base, base, tmp); for (int i = 0; i < 31; i++) base[i] = tmp[i]; }
    }
    for (int i = 0; i < 31; i++) c[i] = acc[i];
}

// The 31 words that precede rand() #k0: hist[t] = r[344 + k0 - 31 + t] = w[310 + k0 + t], t = 0..30,
// given c = x^(310 + k0) mod P and the seed window.
RNG_HD inline void glibc_history(const uint32_t c[31], const uint32_t w[61], uint32_t hist[31])
{
    for (int t = 0; t < 31; t++) {
        uint32_t s = 0;
        for (int j = 0; j < 31; j++) s += c[j] * w[j + t];
        hist[t] = s;
    }
}

// One rand() value from its class: index into "GCAT" (stochasticSpike.c:340) or 4 = rejected by randomNum
RNG_HD inline int glibc_class4(uint32_t r) { return r >= GLIBC_CUT4 ? 4 : (int)(r & 3u); }
