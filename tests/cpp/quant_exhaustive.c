/* Exhaustive check of k_fdct's FMA-pipe quantiser (csrc/enc_fdct.cu stage B, constants from csrc/api.cu make_quant)
 * against jcdctmgr.c's integer form  q = sign(c) * ((|c| + d/2) / d),  d = 8 * qtbl,  for every divisor qtbl = 1..255
 * and every |c| <= 2^18 (the transform's outputs stay below 2^15). Host arithmetic is IEEE single precision with a
 * fused multiply-add, the same operations the kernel issues (FADD, FFMA). Build: gcc -O2 -ffp-contract=off ... -lm */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
int main(void) {
    long bad = 0, n = 0;
    for (int q = 1; q <= 255; q++) {
        const int d = 8 * q;
        /* the table value carries a factor 4096 (exact): the quotient then sits at the 2^12 weight of 1.5 * 2^35, and the
         * same value decides "non-zero" in one saturating FMA: |c| / d * 4096 - 2047 is >= 1 from |c| = d/2 on and < 0 below */
        const float finv = (float)((1.0 / (double)d) * (1.0 + 1.0 / 1048576.0)) * 4096.0f;
        for (int c = -(1 << 18); c <= (1 << 18); c++) {
            const int a = c < 0 ? -c : c;
            int z = (a + d / 2) / d;
            if (c < 0) z = -z;
            const float xf = u2f(0x4B400000u + (uint32_t)c) - 12582912.0f; /* exact int -> float */
            const float r = fmaf(xf, finv, 51539607552.0f);   /* 1.5 * 2^35 */
            const int got = (int)(f2u(r) - 0x51400000u);
            float sat = fmaf(fabsf(xf), finv, -2047.0f);
            if (sat > 1.0f) sat = 1.0f;
            if (!(sat > 0.0f)) sat = 0.0f;
            n++;
            if (got != z || (sat != 0.0f) != (z != 0) || (sat != 0.0f && sat != 1.0f) ||
                (z >= -32768 && z <= 32767 && (int16_t)(f2u(r) & 0xFFFFu) != (int16_t)z)) {
                if (bad < 10) printf("q=%d c=%d want %d got %d\n", q, c, z, got);
                bad++;
            }
        }
    }
    printf("checked %ld bad %ld\n", n, bad);
    return bad != 0;
}
