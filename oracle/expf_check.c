/* TEST INFRASTRUCTURE.  glibc's expf (sysdeps/ieee754/flt-32/e_expf.c, glibc >= 2.27: 32-entry table of 2^(i/32) + cubic in double
 * arithmetic) restated with plain double operations and compared bit for bit with the host libm's expf.  The CUDA detection kernel
 * (csrc/bw_ops.cu glibc_expf) performs exactly these operations; the reference's get_region_box calls std::exp(float)
 * (src/core/yolo_region.cpp:23-24), i.e. this function.  Prints the number of mismatching inputs (must be 0). */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
static double T[32];
static double asd(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }
static uint64_t asu(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }
static float restated_expf(float x)
{
    const double InvLn2N = 0x1.71547652b82fep+0 * 32, SHIFT = 0x1.8p+52;
    const double C0 = 0x1.c6af84b912394p-5 / 32 / 32 / 32, C1 = 0x1.ebfce50fac4f3p-3 / 32 / 32, C2 = 0x1.62e42ff0c52d6p-1 / 32;
    double xd = x, z = InvLn2N * xd, kd = z + SHIFT;
    uint64_t ki = asu(kd);
    kd -= SHIFT;
    double r = z - kd;
    uint64_t t = asu(T[ki % 32]) - ((ki % 32) << 47);
    t += ki << 47;
    double s = asd(t), zz = C0 * r + C1, r2 = r * r, y = C2 * r + 1;
    y = zz * r2 + y;
    y = y * s;
    return (float)y;
}
int main(int argc, char **argv)
{
    long n = argc > 1 ? atol(argv[1]) : (1L << 22), bad = 0;
    for (int i = 0; i < 32; ++i) T[i] = (double)exp2l((long double)i / 32);
    uint32_t seed = 12345;
    for (long i = 0; i < n; ++i) {
        seed = seed * 1664525u + 1013904223u;
        float x = ((int32_t)seed) / (float)(1u << 31) * 20.0f;      /* [-20, 20): far beyond any box w/h logit */
        volatile float a = expf(x);
        float b = restated_expf(x);
        if (memcmp((void *)&a, &b, 4)) ++bad;
    }
    printf("%ld\n", bad);
    return bad != 0;
}
