/* vb_detmath.h -- reproducible single-precision elementary functions.
 *
 * Why this exists: `flatten` decides HOW MANY lines a curve becomes from
 * `ceil(n_frac * scale_multiplier)` (vello_shaders/shader/flatten.wgsl:444), and n_frac is a
 * chain of atan2/sin/cos/asin/acos/pow. WGSL leaves the precision of those builtins
 * implementation-defined and the reference's own CPU twin (Rust std -> libm) is not
 * bit-identical to its GPU path either (SURVEY.md section 7 "flatten parity"). To be able to
 * assert BIT-EXACT parity of every downstream integer (line counts, tile and segment indices)
 * between the CUDA kernels and the CPU oracle, both sides evaluate the same polynomial
 * kernels built only from IEEE-754 exactly-rounded operations: + - * / sqrt, fmaf, rintf/floorf
 * and integer bit manipulation. No libm / libdevice transcendental is called.
 *
 * Accuracy (measured in tests/test_detmath.py against float64 libm): <= 2 ulp on the argument
 * ranges the pipeline uses (angles |x| < 1e4 for sin/cos). That is inside what WGSL allows
 * for every builtin used (e.g. atan2: 4096 ulp, sin/cos: abs error 2^-11).
 *
 * Compile the including translation unit with contraction OFF (gcc -ffp-contract=off,
 * nvcc -fmad=false); every fused operation here is an explicit fmaf().
 *
 * Plain C99 + CUDA; no dependencies.
 */
#ifndef VB_DETMATH_H
#define VB_DETMATH_H

#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define VB_HD __host__ __device__ __forceinline__
#else
#define VB_HD static inline
#endif

VB_HD uint32_t vb_f2u(float f) {
#ifdef __CUDA_ARCH__
    return __float_as_uint(f);
#else
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
#endif
}
VB_HD float vb_u2f(uint32_t u) {
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}

#define VB_PI_F 3.14159274101257324f
#define VB_PIO2_F 1.57079637050628662f
#define VB_PIO4_F 0.785398185253143311f

/* sin and cos together: Cody-Waite reduction by pi/2 (two FMA steps), Cephes minimax kernels on
 * [-pi/4, pi/4]. */
VB_HD void vb_sincosf(float x, float *sn, float *cs) {
    float n = rintf(x * 0.636619746685028076f);
    float r = fmaf(n, -1.57079637050628662f, x);
    r = fmaf(n, 4.37113900018624283e-8f, r);
    float z = r * r;
    float ps = fmaf(fmaf(-1.9515295891e-4f, z, 8.3321608736e-3f), z, -1.6666654611e-1f);
    float s = fmaf(ps * z, r, r);
    float pc = fmaf(fmaf(2.443315711809948e-5f, z, -1.388731625493765e-3f), z, 4.166664568298827e-2f);
    float c = fmaf(pc * z, z, fmaf(-0.5f, z, 1.0f));
    int q = ((int)n) & 3;
    float so = (q & 1) ? c : s;
    float co = (q & 1) ? s : c;
    if (q == 2 || q == 3) so = -so;
    if (q == 1 || q == 2) co = -co;
    *sn = so;
    *cs = co;
}
VB_HD float vb_sinf(float x) {
    float s, c;
    vb_sincosf(x, &s, &c);
    return s;
}
VB_HD float vb_cosf(float x) {
    float s, c;
    vb_sincosf(x, &s, &c);
    return c;
}

/* atan on x >= 0 (Cephes atanf). */
VB_HD float vb_atan_pos(float x) {
    float y;
    if (x > 2.414213562373095f) {
        y = VB_PIO2_F;
        x = -1.0f / x;
    } else if (x > 0.4142135623730950f) {
        y = VB_PIO4_F;
        x = (x - 1.0f) / (x + 1.0f);
    } else {
        y = 0.0f;
    }
    float z = x * x;
    float p = fmaf(fmaf(fmaf(8.05374449538e-2f, z, -1.38776856032e-1f), z, 1.99777106478e-1f), z, -3.33329491539e-1f);
    return y + fmaf(p * z, x, x);
}
VB_HD float vb_atan2f(float y, float x) {
    float ax = fabsf(x), ay = fabsf(y);
    float r;
    if (ax == 0.0f) {
        r = (ay == 0.0f) ? 0.0f : VB_PIO2_F;
    } else {
        r = vb_atan_pos(ay / ax);
    }
    if (x < 0.0f) r = VB_PI_F - r;
    return (y < 0.0f) ? -r : r;
}

/* asin / acos (Cephes asinf); arguments are clamped to [-1, 1]. */
VB_HD float vb_asin_core(float a, int *flag) { /* a in [0,1] */
    float z, s;
    if (a > 0.5f) {
        z = 0.5f * (1.0f - a);
        s = sqrtf(z);
        *flag = 1;
    } else {
        s = a;
        z = a * a;
        *flag = 0;
    }
    float p = fmaf(fmaf(fmaf(fmaf(4.2163199048e-2f, z, 2.4181311049e-2f), z, 4.5470025998e-2f), z, 7.4953002686e-2f), z,
                   1.6666752422e-1f);
    return fmaf(p * z, s, s);
}
VB_HD float vb_asinf(float x) {
    float a = fminf(fabsf(x), 1.0f);
    int flag;
    float r = vb_asin_core(a, &flag);
    if (flag) r = VB_PIO2_F - (r + r);
    return (x < 0.0f) ? -r : r;
}
VB_HD float vb_acosf(float x) {
    if (x < -1.0f) x = -1.0f;
    if (x > 1.0f) x = 1.0f;
    int flag;
    if (x > 0.5f) {
        float r = vb_asin_core(x, &flag); /* flag==1: r = asin(sqrt((1-x)/2)) */
        return r + r;
    }
    if (x < -0.5f) {
        float r = vb_asin_core(-x, &flag);
        return VB_PI_F - (r + r);
    }
    float r = vb_asin_core(fabsf(x), &flag); /* flag==0 */
    return (x < 0.0f) ? VB_PIO2_F + r : VB_PIO2_F - r;
}

/* cbrt for finite x; Kahan's bit-hack seed, two Newton steps and one FMA-residual correction. */
VB_HD float vb_cbrtf(float x) {
    float a = fabsf(x);
    if (a == 0.0f || !(a < INFINITY)) return x;
    float scale = 1.0f;
    if (a < 1.17549435e-38f) { /* denormal: scale by 2^24, undo by 2^-8 */
        a *= 16777216.0f;
        scale = 0.00390625f;
    }
    float t = vb_u2f(vb_f2u(a) / 3u + 709958130u);
    t = (t + t + a / (t * t)) * 0.333333343f;
    t = (t + t + a / (t * t)) * 0.333333343f;
    float t2 = t * t;
    float e = fmaf(-t2, t, a);
    t = t + e / (3.0f * t2);
    t *= scale;
    return (x < 0.0f) ? -t : t;
}
/* |x|^(2/3), the only fractional power flatten needs (flatten.wgsl:456, :278). */
VB_HD float vb_pow_2_3(float ax) {
    float c = vb_cbrtf(ax);
    return c * c;
}

/* 2^n * m for integer-valued n, split so that denormal results round once. */
VB_HD float vb_ldexpf_i(float m, int n) {
    if (n > 127) {
        m *= 1.70141183e38f; /* 2^127 */
        n -= 127;
        if (n > 127) n = 127;
    } else if (n < -126) {
        m *= 1.17549435e-38f; /* 2^-126 */
        n += 126;
        if (n < -126) n = -126;
    }
    return m * vb_u2f((uint32_t)(n + 127) << 23);
}
VB_HD float vb_expf(float x) {
    if (x > 88.72283905f) return INFINITY;
    if (x < -103.972f) return 0.0f;
    if (x != x) return x;
    float n = rintf(x * 1.44269504088896341f);
    float r = fmaf(n, -0.693145751953125f, x);
    r = fmaf(n, -1.428606765330187045e-06f, r);
    float p = fmaf(fmaf(fmaf(fmaf(fmaf(1.9875691500e-4f, r, 1.3981999507e-3f), r, 8.3334519073e-3f), r, 4.1665795894e-2f), r,
                        1.6666665459e-1f), r, 5.0000001201e-1f);
    float e = fmaf(p * r, r, r) + 1.0f;
    return vb_ldexpf_i(e, (int)n);
}
/* natural log for x > 0 (Cephes logf). */
VB_HD float vb_logf(float x) {
    if (!(x > 0.0f)) return (x == 0.0f) ? -INFINITY : NAN;
    if (!(x < INFINITY)) return x;
    int e = 0;
    uint32_t u = vb_f2u(x);
    if (u < 0x00800000u) { /* denormal */
        x *= 8388608.0f;
        u = vb_f2u(x);
        e = -23;
    }
    e += (int)(u >> 23) - 126;
    float m = vb_u2f((u & 0x007fffffu) | 0x3f000000u); /* [0.5, 1) */
    if (m < 0.707106781186547524f) {
        e -= 1;
        m = m + m - 1.0f;
    } else {
        m = m - 1.0f;
    }
    float z = m * m;
    float p = fmaf(fmaf(fmaf(fmaf(fmaf(fmaf(fmaf(fmaf(7.0376836292e-2f, m, -1.1514610310e-1f), m, 1.1676998740e-1f), m,
                                                  -1.2420140846e-1f), m, 1.4249322787e-1f), m, -1.6668057665e-1f), m,
                             2.0000714765e-1f), m, -2.4999993993e-1f), m, 3.3333331174e-1f);
    float y = m * z * p;
    float fe = (float)e;
    y = fmaf(fe, -2.12194440e-4f, y);
    y = fmaf(-0.5f, z, y);
    float r = m + y;
    return fmaf(fe, 0.693359375f, r);
}
/* x^y for x >= 0 (the only domain fine's blurred rounded rect uses, fine.wgsl:1199,1233). */
VB_HD float vb_powf_pos(float x, float y) {
    if (y == 0.0f) return 1.0f;
    if (x == 0.0f) return (y > 0.0f) ? 0.0f : INFINITY;
    return vb_expf(y * vb_logf(x));
}

#endif /* VB_DETMATH_H */
