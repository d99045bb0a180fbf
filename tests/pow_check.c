/* Host check of csrc/ssb_pow.cuh against the host libm (tests/test_pow_exact.py builds and runs
 * it): prints the number of inputs whose result differs from pow() in any bit. */
#include <stdio.h>
#include <stdlib.h>
#include "../sugarscape_b200/csrc/ssb_pow.cuh"

static uint64_t s = 88172645463325252ull;
static uint64_t rnd(void) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }

int main(int argc, char **argv) {
    long n = argc > 1 ? atol(argv[1]) : 1000000, bad = 0, inexact = 0, total = 0;
    /* welfare-like inputs: small integers and halves ** (m / M) */
    for (int m = 1; m <= 12; m++)
        for (int M = m; M <= 24; M++)
            for (int xi = 1; xi <= 2000; xi++) {
                double x = xi / 2.0, y = (double)m / (double)M;
                int exact;
                double a = ssb_py_pow(x, y, &exact), b = pow(x, y);
                total++; inexact += !exact;
                if (memcmp(&a, &b, 8)) { if (bad++ < 5) printf("DIFF pow(%a, %a): %a vs %a\n", x, y, a, b); }
            }
    for (long k = 0; k < n; k++) {
        double x = (double)(rnd() % 100000000) / (double)(1 + rnd() % 4096);
        double y = (double)(1 + rnd() % 64) / (double)(1 + rnd() % 64);
        if (k & 1) { x = ldexp((double)(rnd() >> 11), -53 + (int)(rnd() % 40) - 10); y = ldexp((double)(rnd() >> 11), -53 + (int)(rnd() % 8) - 4); }
        if (x <= 0) continue;
        int exact;
        double a = ssb_py_pow(x, y, &exact), b = pow(x, y);
        total++; inexact += !exact;
        if (memcmp(&a, &b, 8)) { if (bad++ < 5) printf("DIFF pow(%a, %a): %a vs %a\n", x, y, a, b); }
    }
    printf("checked=%ld bad=%ld fallback=%ld\n", total, bad, inexact);
    return bad != 0;
}
