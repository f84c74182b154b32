/* Exhaustive-ish check of the GAE kernels' table division (di_hpc_b200/csrc/gae.cu: div_by_table):
 *     r = RN(1/d);  q0 = RN(x*r);  q = fma(fma(-d, q0, x), r, q0)        must equal RN(x/d)
 * for every denominator d_t the GAE recurrence d_t = 1 + lambda*d_{t+1} produces (origin/gae.py:34, evaluated in
 * double and rounded to fp32 like the library does) and a stream of numerators x.  Uses the hardware fmaf, i.e. the
 * same single-rounding FMA the GPU executes.  Prints "<cases> <mismatches>".  Test infrastructure only. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

static uint64_t s = 0x9E3779B97F4A7C15ull;
static uint32_t rnd(void) {
    s ^= s << 13;
    s ^= s >> 7;
    s ^= s << 17;
    return (uint32_t)(s >> 32);
}

int main(void) {
    const double lambdas[] = {0.97, 0.95, 0.9, 0.8, 0.5, 1.0, 0.999, 0.3};
    const int T = 2048;
    unsigned long long cases = 0, bad = 0;
    for (unsigned li = 0; li < sizeof(lambdas) / sizeof(lambdas[0]); ++li) {
        double den = 0.0;
        for (int t = T - 1; t >= 0; --t) {
            den = 1.0 + lambdas[li] * den;
            const float d = (float)den;
            const float r = 1.0f / d;
            for (int k = 0; k < 4096; ++k) {
                float x;
                uint32_t bits = rnd();
                if (k & 1) {  /* any finite normal-range float, both signs */
                    bits = (bits & 0x807FFFFFu) | ((uint32_t)(64 + (rnd() % 128)) << 23);
                    memcpy(&x, &bits, 4);
                } else {      /* the magnitudes GAE actually sees: |g| up to a few d */
                    x = ((float)(int32_t)bits / 2147483648.0f) * 8.0f * d;
                }
                const float q0 = x * r;
                const float q = fmaf(fmaf(-d, q0, x), r, q0);
                const float want = x / d;
                ++cases;
                if (!(q == want)) ++bad;
            }
        }
    }
    printf("%llu %llu\n", cases, bad);
    return 0;
}
