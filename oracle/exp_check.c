/* TEST INFRASTRUCTURE.  glibc's double exp (sysdeps/ieee754/dbl-64/e_exp.c, glibc >= 2.28; the FMA build the x86-64 ifunc selects on
 * every CPU with FMA: contraction pattern read off the disassembly of libm 2.39) restated with explicit fma()/mul/add and compared bit
 * for bit with the host libm's exp.  The CUDA region kernel (csrc/bw_ops.cu glibc_exp) performs exactly these operations; the
 * reference's logistic_activate / softmax call exp(double) (src/core/yolo_math.cpp:19,234).  Prints the number of mismatching
 * inputs (must be 0).  Usage: exp_check [n] */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../yolo-fpga-accelerator_b200/csrc/glibc_exp_data.h"
static const uint64_t T[256] = {Y2_GLIBC_EXP_TAB};
static double asd(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }
static uint64_t asu(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }
static double restated_exp(double x)
{
    uint32_t abstop = (uint32_t)(asu(x) >> 52) & 0x7ff;
    if (abstop - 0x3c9 >= 0x3f) {                        /* |x| < 2^-54 or |x| >= 512 or NaN */
        if (abstop - 0x3c9 >= 0x80000000u) return 1.0 + x;
        if (abstop >= 0x409) {                           /* |x| >= 1024 */
            if (asu(x) == asu(-INFINITY)) return 0.0;
            if (abstop >= 0x7ff) return 1.0 + x;
            return (asu(x) >> 63) ? 0x1p-767 * 0x1p-767 : 0x1p769 * 0x1p769;
        }
        abstop = 0;
    }
    double kd = fma(Y2_GLIBC_EXP_INVLN2N, x, Y2_GLIBC_EXP_SHIFT);
    uint64_t ki = asu(kd);
    kd -= Y2_GLIBC_EXP_SHIFT;
    double r = fma(kd, Y2_GLIBC_EXP_NEGLN2LON, fma(kd, Y2_GLIBC_EXP_NEGLN2HIN, x));
    uint64_t idx = 2 * (ki % 128), top = ki << 45;
    double tail = asd(T[idx]);
    uint64_t sbits = T[idx + 1] + top;
    double r2 = r * r;
    double tmp = fma(r2 * r2, fma(r, Y2_GLIBC_EXP_C5, Y2_GLIBC_EXP_C4), fma(fma(r, Y2_GLIBC_EXP_C3, Y2_GLIBC_EXP_C2), r2, tail + r));
    if (abstop == 0) {
        if ((ki & 0x80000000u) == 0) {                   /* k > 0: the exponent of scale may have overflowed */
            sbits -= 1009ull << 52;
            double scale = asd(sbits);
            return 0x1p1009 * fma(scale, tmp, scale);
        }
        sbits += 1022ull << 52;                          /* k < 0: care in the subnormal range */
        double scale = asd(sbits), st = scale * tmp, y = scale + st;
        if (y < 1.0) {
            double lo = scale - y + st, hi = 1.0 + y;
            lo = 1.0 - hi + y + lo;
            y = (hi + lo) - 1.0;
            if (y == 0.0) y = 0.0;
        }
        return 0x1p-1022 * y;
    }
    double scale = asd(sbits);
    return fma(scale, tmp, scale);
}
int main(int argc, char **argv)
{
    long n = argc > 1 ? atol(argv[1]) : (1L << 22), bad = 0;
    uint64_t seed = 88172645463325252ull;
    for (long i = 0; i < n; ++i) {
        seed ^= seed << 13; seed ^= seed >> 7; seed ^= seed << 17;
        double x;
        switch (i & 3) {
        case 0: x = ((double)(int64_t)seed) / 9223372036854775808.0 * 40.0; break;              /* [-40, 40): the region head's range */
        case 1: x = (double)(float)(((double)(int64_t)seed) / 9223372036854775808.0 * 90.0); break;   /* float-valued arguments */
        case 2: x = ((double)(int64_t)seed) / 9223372036854775808.0 * 1100.0; break;            /* incl. the scaled special cases and over/underflow */
        default: x = asd(seed); break;                                                          /* any bit pattern */
        }
        volatile double a = exp(x);
        double b = restated_exp(x);
        if (asu(a) != asu(b) && !(a != a && b != b)) {
            if (++bad <= 5) fprintf(stderr, "x=%a libm=%a restated=%a\n", x, (double)a, b);
        }
    }
    printf("%ld\n", bad);
    return bad != 0;
}
