/* tools/check_division.c -- exhaustive CPU check of the kernels' division (csrc/mpi_common.cuh: div_by_rcp).
 *
 *   q = a / b  is computed as  y = RN(1/b);  q0 = RN(a*y);  r = fma(-q0, b, a);  q = fma(r, y, q0)
 *
 * and must equal the IEEE quotient RN(a/b) (the reference divides with ATen's fp32 `div`).  Scaling a or b by a power of
 * two scales every intermediate exactly (no overflow/underflow inside the kernels' guarded exponent range 2^+-40), so it is
 * enough to check mantissas: ALL 2^23 divisor mantissas b in [1,2) against a structured set of dividend mantissas a in [1,2)
 * (the extremes, values adjacent to b and to 2/b-type rounding boundaries, and a pseudo-random fill): 2^23 x 432 = 3.6e9 cases.
 * Build and run (about a minute on 8 cores):  gcc -O2 -ffp-contract=off -o /tmp/check_division tools/check_division.c -lm -lpthread
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define NA 432
static float A[NA];
static inline float from_bits(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

typedef struct { uint32_t lo, hi; unsigned long long bad, n; } job_t;

static void* run(void* arg) {
    job_t* j = (job_t*)arg;
    for (uint32_t m = j->lo; m < j->hi; ++m) {
        const float b = from_bits(0x3f800000u | m);
        const float y = 1.0f / b;                       /* RN(1/b), what __frcp_rn returns */
        for (int k = 0; k < NA + 6; ++k) {
            float a;
            if (k < NA) a = A[k];
            else {                                      /* dividends tied to this divisor */
                const uint32_t bb = bits(b);
                const uint32_t near[6] = {bb, bb + 1, bb - 1, bb ^ 0x400000u, (bb + 0x155555u) & 0x3fffffffu, 0x3fffffffu - (bb & 0x7fffffu)};
                a = from_bits(0x3f800000u | (near[k - NA] & 0x7fffffu));
            }
            const float q0 = a * y;
            const float r = fmaf(-q0, b, a);
            const float q = fmaf(r, y, q0);
            j->n++;
            if (bits(q) != bits(a / b)) {
                if (j->bad < 5) fprintf(stderr, "MISMATCH a=%a b=%a fast=%a ieee=%a\n", a, b, q, a / b);
                j->bad++;
            }
        }
    }
    return NULL;
}

int main(void) {
    uint32_t s = 12345u;
    int n = 0;
    const uint32_t fixed[] = {0, 1, 2, 3, 0x7fffff, 0x7ffffe, 0x7ffffd, 0x400000, 0x3fffff, 0x400001, 0x200000, 0x600000,
                              0x555555, 0x2aaaaa, 0x555556, 0x2aaaab, 0x333333, 0x4ccccd, 0x100000, 0x700000};
    for (unsigned i = 0; i < sizeof(fixed) / 4; ++i) A[n++] = from_bits(0x3f800000u | fixed[i]);
    for (int k = 0; k < 23; ++k) { A[n++] = from_bits(0x3f800000u | (1u << k)); A[n++] = from_bits(0x3f800000u | (0x7fffffu ^ (1u << k))); }
    while (n < NA) { s = s * 1664525u + 1013904223u; A[n++] = from_bits(0x3f800000u | (s >> 9)); }
    enum { T = 8 };
    pthread_t th[T];
    job_t jobs[T];
    for (int t = 0; t < T; ++t) {
        jobs[t].lo = (uint32_t)((1ull << 23) * t / T); jobs[t].hi = (uint32_t)((1ull << 23) * (t + 1) / T); jobs[t].bad = jobs[t].n = 0;
        pthread_create(&th[t], NULL, run, &jobs[t]);
    }
    unsigned long long bad = 0, tot = 0;
    for (int t = 0; t < T; ++t) { pthread_join(th[t], NULL); bad += jobs[t].bad; tot += jobs[t].n; }
    printf("div_by_rcp vs IEEE division: %llu cases (all 2^23 divisor mantissas x %d dividends), %llu mismatches\n", tot, NA + 6, bad);
    return bad != 0;
}
