// k_flatten.cu -- curve flattening + stroke expansion to LineSoup, per-path bounding boxes.
//
// Reference: vello_shaders/shader/flatten.wgsl (Euler-spiral flatten :326-481, arcs :494-519,
// caps :521-545, joins :547-631, segment decode :683-766, main :831-923) and its CPU twin
// vello_shaders/src/cpu/{flatten,euler}.rs. Also folds in bbox_clear.wgsl.
//
// B200 design (differs from the WGSL on purpose):
//  * The WGSL bump-allocates every line with a global atomicAdd, so the line order is a race.
//    Here each thread first COUNTS the lines of its tag, the CTA scans the counts with warp
//    shuffles, resolves its base by decoupled look-back over CTAs, and then EMITS: line order is
//    deterministic (tag order, then emission order), identical to the serial CPU shader, and there
//    is no per-line atomic. Consecutive threads write consecutive 24 B records.
//  * Transcendentals come from vb_detmath.h (IEEE-only) and the TU is compiled with -fmad=false,
//    so `lines` is bit-identical to the oracle's, and so is everything downstream.
// Algorithmic bytes: 1 B tag + 20/4 B monoid + <= 32 B coords per segment, 24 B per line out.
#include <cuda_fp16.h>

#include "vb_detmath.h"
#include "vb_device.cuh"

#define FL_THREADS 256

struct fv2 { float x, y; };
__device__ __forceinline__ fv2 F2(float x, float y) { fv2 r; r.x = x; r.y = y; return r; }
__device__ __forceinline__ fv2 operator+(fv2 a, fv2 b) { return F2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ fv2 operator-(fv2 a, fv2 b) { return F2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ fv2 operator*(fv2 a, float s) { return F2(a.x * s, a.y * s); }
__device__ __forceinline__ float fdot(fv2 a, fv2 b) { return a.x * b.x + a.y * b.y; }
__device__ __forceinline__ float flen(fv2 a) { return sqrtf(a.x * a.x + a.y * a.y); }
__device__ __forceinline__ fv2 fnorm(fv2 a) { float l = flen(a); return F2(a.x / l, a.y / l); }
__device__ __forceinline__ bool feq(fv2 a, fv2 b) { return a.x == b.x && a.y == b.y; }

struct FXform { float m0, m1, m2, m3, tx, ty; };
__device__ __forceinline__ fv2 fx_apply(const FXform &t, fv2 p) { // flatten.wgsl:668-672 (explicit fma)
    return F2(fmaf(t.m0, p.x, fmaf(t.m2, p.y, t.tx)), fmaf(t.m1, p.x, fmaf(t.m3, p.y, t.ty)));
}

// MODE 0: count lines only. MODE 1: count, accumulate the bbox and stash the first FL_CACHE lines of this thread
// in shared memory (most tags produce <= FL_CACHE lines, so the geometry is computed once). MODE 2: emit to global.
#define FL_CACHE 6
template <int MODE>
struct Flat {
    VbLineSoup *lines;
    uint32_t lines_size;
    uint32_t ix; // next line slot (MODE 2) / running count (MODE 0, 1)
    float bx0, by0, bx1, by1;
    float4 *cache; // MODE 1: &cache[0][threadIdx.x], stride FL_THREADS
    __device__ __forceinline__ void write_line(uint32_t path_ix, fv2 p0, fv2 p1) {
        if (MODE != 0) {
            bx0 = fminf(bx0, fminf(p0.x, p1.x));
            by0 = fminf(by0, fminf(p0.y, p1.y));
            bx1 = fmaxf(bx1, fmaxf(p0.x, p1.x));
            by1 = fmaxf(by1, fmaxf(p0.y, p1.y));
        }
        if (MODE == 1) {
            if (ix < FL_CACHE) cache[ix * FL_THREADS] = make_float4(p0.x, p0.y, p1.x, p1.y);
        }
        if (MODE == 2) {
            if (ix < lines_size) {
                uint2 *dst = reinterpret_cast<uint2 *>(lines + ix);
                dst[0] = make_uint2(path_ix, 0u);
                dst[1] = make_uint2(__float_as_uint(p0.x), __float_as_uint(p0.y));
                dst[2] = make_uint2(__float_as_uint(p1.x), __float_as_uint(p1.y));
            }
        }
        ix++;
    }
    __device__ __forceinline__ void line_xf(uint32_t path_ix, fv2 p0, fv2 p1, const FXform &t) {
        if (MODE != 0) write_line(path_ix, fx_apply(t, p0), fx_apply(t, p1));
        else ix++;
    }
};

#define DERIV_THRESH 1e-6f
#define DERIV_THRESH_SQUARED (DERIV_THRESH * DERIV_THRESH)
#define DERIV_EPS 1e-6f
#define SUBDIV_LIMIT (1.0f / 65536.0f)
#define K1_THRESH 1e-3f
#define DIST_THRESH 1e-3f
#define TANGENT_THRESH 1e-6f

struct CubicParams { float th0, th1, chord_len, err; };
struct EulerParams { float th0, k0, k1, ch; };

__device__ CubicParams cubic_from_points_derivs(fv2 p0, fv2 p1, fv2 q0, fv2 q1, float dt) { // flatten.wgsl:94-133
    CubicParams r;
    fv2 chord = p1 - p0;
    float chord_squared = fdot(chord, chord);
    float chord_len = sqrtf(chord_squared);
    if (chord_squared < DERIV_THRESH_SQUARED) {
        float chord_err = sqrtf((9.f / 32.0f) * (fdot(q0, q0) + fdot(q1, q1))) * dt;
        r.th0 = 0.f; r.th1 = 0.f; r.chord_len = DERIV_THRESH; r.err = chord_err;
        return r;
    }
    float scale = dt / chord_squared;
    fv2 h0 = F2(q0.x * chord.x + q0.y * chord.y, q0.y * chord.x - q0.x * chord.y);
    float th0 = vb_atan2f(h0.y, h0.x);
    float d0 = flen(h0) * scale;
    fv2 h1 = F2(q1.x * chord.x + q1.y * chord.y, q1.x * chord.y - q1.y * chord.x);
    float th1 = vb_atan2f(h1.y, h1.x);
    float d1 = flen(h1) * scale;
    float s0, cth0, s1, cth1;
    vb_sincosf(th0, &s0, &cth0);
    vb_sincosf(th1, &s1, &cth1);
    float err = 2.0f;
    if (cth0 * cth1 >= 0.0f) {
        float e0 = (2.f / 3.f) / fmaxf(1.0f + cth0, 1e-9f);
        float e1 = (2.f / 3.f) / fmaxf(1.0f + cth1, 1e-9f);
        float s01 = cth0 * s1 + cth1 * s0;
        float amin = 0.15f * (2.f * e0 * s0 + 2.f * e1 * s1 - e0 * e1 * s01);
        float a = 0.15f * (2.f * d0 * s0 + 2.f * d1 * s1 - d0 * d1 * s01);
        float aerr = fabsf(a - amin);
        float symm = fabsf(th0 + th1);
        float asymm = fabsf(th0 - th1);
        float dist = flen(F2(d0 - e0, d1 - e1));
        float symm2 = symm * symm;
        float ctr = (4.625e-6f * symm * symm2 + 7.5e-3f * asymm) * symm2;
        float halo = (5e-3f * symm + 7e-2f * asymm) * dist;
        err = ctr + 1.55f * aerr + halo;
    }
    err *= chord_len;
    r.th0 = th0; r.th1 = th1; r.chord_len = chord_len; r.err = err;
    return r;
}

__device__ EulerParams es_params_from_angles(float th0, float th1) { // flatten.wgsl:135-161
    float k0 = th0 + th1;
    float dth = th1 - th0;
    float d2 = dth * dth;
    float k2 = k0 * k0;
    float a = 6.0f;
    a -= d2 * (1.f / 70.f);
    a -= (d2 * d2) * (1.f / 10780.f);
    a += (d2 * d2 * d2) * 2.769178184818219e-07f;
    float b = -0.1f + d2 * (1.f / 4200.f) + d2 * d2 * 1.6959677820260655e-05f;
    float cc = -1.f / 1400.f + d2 * 6.84915970574303e-05f - k2 * 7.936475029053326e-06f;
    a += (b + cc * k2) * k2;
    float k1 = dth * a;
    float ch = 1.0f;
    ch -= d2 * (1.f / 40.f);
    ch += (d2 * d2) * 0.00034226190482569864f;
    ch -= (d2 * d2 * d2) * 1.9349474568904524e-06f;
    float b_ = -1.f / 24.f + d2 * 0.0024702380951963226f - d2 * d2 * 3.7297408997537985e-05f;
    float c_ = 1.f / 1920.f - d2 * 4.87350869747975e-05f - k2 * 3.1001936068463107e-06f;
    ch += (b_ + c_ * k2) * k2;
    EulerParams r = {th0, k0, k1, ch};
    return r;
}
__device__ __forceinline__ float es_eval_th(const EulerParams &p, float t) { return (p.k0 + 0.5f * p.k1 * (t - 1.0f)) * t - p.th0; }

__device__ fv2 integ_euler_10(float k0, float k1) { // flatten.wgsl:168-202
    float t1_1 = k0;
    float t1_2 = 0.5f * k1;
    float t2_2 = t1_1 * t1_1;
    float t2_3 = 2.f * (t1_1 * t1_2);
    float t2_4 = t1_2 * t1_2;
    float t3_4 = t2_2 * t1_2 + t2_3 * t1_1;
    float t3_6 = t2_4 * t1_2;
    float t4_4 = t2_2 * t2_2;
    float t4_5 = 2.f * (t2_2 * t2_3);
    float t4_6 = 2.f * (t2_2 * t2_4) + t2_3 * t2_3;
    float t4_7 = 2.f * (t2_3 * t2_4);
    float t4_8 = t2_4 * t2_4;
    float t5_6 = t4_4 * t1_2 + t4_5 * t1_1;
    float t5_8 = t4_6 * t1_2 + t4_7 * t1_1;
    float t6_6 = t4_4 * t2_2;
    float t6_7 = t4_4 * t2_3 + t4_5 * t2_2;
    float t6_8 = t4_4 * t2_4 + t4_5 * t2_3 + t4_6 * t2_2;
    float t7_8 = t6_6 * t1_2 + t6_7 * t1_1;
    float t8_8 = t6_6 * t2_2;
    float u = 1.f;
    u -= (1.f / 24.f) * t2_2 + (1.f / 160.f) * t2_4;
    u += (1.f / 1920.f) * t4_4 + (1.f / 10752.f) * t4_6 + (1.f / 55296.f) * t4_8;
    u -= (1.f / 322560.f) * t6_6 + (1.f / 1658880.f) * t6_8;
    u += (1.f / 92897280.f) * t8_8;
    float v = (1.f / 12.f) * t1_2;
    v -= (1.f / 480.f) * t3_4 + (1.f / 2688.f) * t3_6;
    v += (1.f / 53760.f) * t5_6 + (1.f / 276480.f) * t5_8;
    v -= (1.f / 11612160.f) * t7_8;
    return F2(u, v);
}
__device__ fv2 es_seg_eval_with_offset(fv2 p0, fv2 p1, const EulerParams &p, float t, float normalized_offset) {
    // es_params_eval_with_offset + es_params_eval (flatten.wgsl:204-231)
    float th = es_eval_th(p, t);
    float sth, cth;
    vb_sincosf(th, &sth, &cth);
    fv2 v = F2(normalized_offset * sth, normalized_offset * cth);
    float thm = es_eval_th(p, t * 0.5f);
    fv2 uv = integ_euler_10((p.k0 + p.k1 * (0.5f * t - 0.5f)) * t, p.k1 * t * t);
    float scale = t / p.ch;
    float sm, cm;
    vb_sincosf(thm, &sm, &cm);
    float s = scale * sm;
    float c = scale * cm;
    fv2 e = F2(uv.x * c - uv.y * s, -uv.y * c - uv.x * s);
    fv2 xy = e + v;
    fv2 chord = p1 - p0;
    return F2(p0.x + (chord.x * xy.x - chord.y * xy.y), p0.y + (chord.x * xy.y + chord.y * xy.x));
}
__device__ __forceinline__ float pow_1_5_signed(float x) { return x * sqrtf(fabsf(x)); }

#define BREAK1 0.8f
#define BREAK2 1.25f
#define BREAK3 2.1f
#define SIN_SCALE 1.0976991822760038f
#define QUAD_A1 0.6406f
#define QUAD_B1 (-0.81f)
#define QUAD_C1 0.9148117935952064f
#define QUAD_A2 0.5f
#define QUAD_B2 (-0.156f)
#define QUAD_C2 0.16145779359520596f
#define FRAC_PI_4 0.7853981633974483f
#define CBRT_9_8 1.040041911525952f

__device__ float espc_int_approx(float x) { // flatten.wgsl:246-259
    float y = fabsf(x);
    float a;
    if (y < BREAK1) {
        a = vb_sinf(SIN_SCALE * y) * (1.0f / SIN_SCALE);
    } else if (y < BREAK2) {
        a = (sqrtf(8.0f) / 3.0f) * pow_1_5_signed(y - 1.0f) + FRAC_PI_4;
    } else {
        float qa = y < BREAK3 ? QUAD_A1 : QUAD_A2;
        float qb = y < BREAK3 ? QUAD_B1 : QUAD_B2;
        float qc = y < BREAK3 ? QUAD_C1 : QUAD_C2;
        a = (qa * y + qb) * y + qc;
    }
    return a * vb_signf(x);
}
__device__ float espc_int_inv_approx(float x) { // flatten.wgsl:261-275
    float y = fabsf(x);
    float a;
    if (y < 0.7010707591262915f) {
        a = vb_asinf(y * SIN_SCALE) * (1.0f / SIN_SCALE);
    } else if (y < 0.903249293595206f) {
        float b = y - FRAC_PI_4;
        float u = vb_pow_2_3(fabsf(b)) * vb_signf(b);
        a = u * CBRT_9_8 + 1.0f;
    } else {
        const float W1 = 0.5f * QUAD_B1 / QUAD_A1, V1 = 1.0f / QUAD_A1, U1 = W1 * W1 - QUAD_C1 / QUAD_A1;
        const float W2 = 0.5f * QUAD_B2 / QUAD_A2, V2_ = 1.0f / QUAD_A2, U2 = W2 * W2 - QUAD_C2 / QUAD_A2;
        bool first = y < 2.038857793595206f;
        float u = first ? U1 : U2, v = first ? V1 : V2_, w = first ? W1 : W2;
        a = sqrtf(u + v * y) - w;
    }
    return a * vb_signf(x);
}

struct PointDeriv { fv2 p, q; };
__device__ PointDeriv eval_cubic_and_deriv(fv2 p0, fv2 p1, fv2 p2, fv2 p3, float t) { // flatten.wgsl:282-290
    float m = 1.0f - t;
    float mm = m * m;
    float mt = m * t;
    float tt = t * t;
    PointDeriv r;
    float a = mm * m, b = 3.0f * mm, c = 3.0f * mt;
    r.p.x = p0.x * a + ((p1.x * b + p2.x * c) + p3.x * tt) * t;
    r.p.y = p0.y * a + ((p1.y * b + p2.y * c) + p3.y * tt) * t;
    float d = 2.0f * mt;
    r.q.x = ((p1.x - p0.x) * mm + (p2.x - p1.x) * d) + (p3.x - p2.x) * tt;
    r.q.y = ((p1.y - p0.y) * mm + (p2.y - p1.y) * d) + (p3.y - p2.y) * tt;
    return r;
}
__device__ fv2 cubic_start_tangent(fv2 p0, fv2 p1, fv2 p2, fv2 p3) {
    const float EPS = 1e-12f;
    fv2 d01 = p1 - p0, d02 = p2 - p0, d03 = p3 - p0;
    if (fdot(d01, d01) > EPS) return d01;
    if (fdot(d02, d02) > EPS) return d02;
    return d03;
}
__device__ fv2 cubic_end_tangent(fv2 p0, fv2 p1, fv2 p2, fv2 p3) {
    const float EPS = 1e-12f;
    fv2 d23 = p3 - p2, d13 = p3 - p1, d03 = p3 - p0;
    if (fdot(d23, d23) > EPS) return d23;
    if (fdot(d13, d13) > EPS) return d13;
    return d03;
}

struct CubicPoints { fv2 p0, p1, p2, p3; };

template <int EMIT>
__device__ void flatten_euler(Flat<EMIT> &f, const CubicPoints &cubic, uint32_t path_ix, const FXform &local_to_device,
                              float offset, fv2 start_p, fv2 end_p) { // flatten.wgsl:326-481
    fv2 p0, p1, p2, p3;
    float scale;
    FXform transform;
    fv2 t_start = start_p, t_end = end_p;
    if (offset == 0.f) {
        p0 = fx_apply(local_to_device, cubic.p0);
        p1 = fx_apply(local_to_device, cubic.p1);
        p2 = fx_apply(local_to_device, cubic.p2);
        p3 = fx_apply(local_to_device, cubic.p3);
        scale = 1.f;
        transform.m0 = 1.f; transform.m1 = 0.f; transform.m2 = 0.f; transform.m3 = 1.f; transform.tx = 0.f; transform.ty = 0.f;
        t_start = p0;
        t_end = p3;
    } else {
        p0 = cubic.p0; p1 = cubic.p1; p2 = cubic.p2; p3 = cubic.p3;
        transform = local_to_device;
        scale = 0.5f * (flen(F2(transform.m0 + transform.m3, transform.m1 - transform.m2)) +
                        flen(F2(transform.m0 - transform.m3, transform.m1 + transform.m2)));
    }
    if (feq(p0, p1) && feq(p0, p2) && feq(p0, p3)) return;
    const float tol = 0.25f;
    uint32_t t0_u = 0u;
    float dt = 1.0f;
    fv2 last_p = p0;
    fv2 last_q = p1 - p0;
    if (fdot(last_q, last_q) < DERIV_THRESH_SQUARED) last_q = eval_cubic_and_deriv(p0, p1, p2, p3, DERIV_EPS).q;
    float last_t = 0.0f;
    fv2 lp0 = t_start;
    for (;;) {
        float t0 = (float)t0_u * dt;
        if (t0 == 1.0f) break;
        float t1 = t0 + dt;
        fv2 this_p0 = last_p;
        fv2 this_q0 = last_q;
        PointDeriv this_pq1 = eval_cubic_and_deriv(p0, p1, p2, p3, t1);
        if (fdot(this_pq1.q, this_pq1.q) < DERIV_THRESH_SQUARED) {
            PointDeriv new_pq1 = eval_cubic_and_deriv(p0, p1, p2, p3, t1 - DERIV_EPS);
            this_pq1.q = new_pq1.q;
            if (t1 < 1.0f) {
                this_pq1.p = new_pq1.p;
                t1 = t1 - DERIV_EPS;
            }
        }
        float actual_dt = t1 - last_t;
        CubicParams cp = cubic_from_points_derivs(this_p0, this_pq1.p, this_q0, this_pq1.q, actual_dt);
        if (cp.err * scale <= tol || dt <= SUBDIV_LIMIT) {
            EulerParams ep = es_params_from_angles(cp.th0, cp.th1);
            float k0 = ep.k0 - 0.5f * ep.k1;
            float k1 = ep.k1;
            float normalized_offset = offset / cp.chord_len;
            float dist_scaled = normalized_offset * ep.ch;
            float scale_multiplier = sqrtf(0.125f * scale * cp.chord_len / (ep.ch * tol));
            float a = 0.0f, b = 0.0f, integral = 0.0f, int0 = 0.0f, n_frac;
            int robust = 0;
            if (fabsf(k1) < K1_THRESH) {
                float k = ep.k0;
                n_frac = sqrtf(fabsf(k * (k * dist_scaled + 1.0f)));
                robust = 1;
            } else if (fabsf(dist_scaled) < DIST_THRESH) {
                a = k1;
                b = k0;
                int0 = pow_1_5_signed(b);
                float int1 = pow_1_5_signed(a + b);
                integral = int1 - int0;
                n_frac = (2.f / 3.f) * integral / a;
                robust = 2;
            } else {
                a = -2.0f * dist_scaled * k1;
                b = -1.0f - 2.0f * dist_scaled * k0;
                int0 = espc_int_approx(b);
                float int1 = espc_int_approx(a + b);
                integral = int1 - int0;
                float k_peak = k0 - k1 * b / a;
                float integrand_peak = sqrtf(fabsf(k_peak * (k_peak * dist_scaled + 1.0f)));
                n_frac = integral * integrand_peak / a;
            }
            float n = vb_clampf(ceilf(n_frac * scale_multiplier), 1.0f, 100.0f);
            uint32_t n_u = vb_f2u_sat(n);
            if (EMIT != 0) {
                for (uint32_t i = 0u; i < n_u; i++) {
                    fv2 lp1;
                    if (i + 1u == n_u && t1 == 1.0f) {
                        lp1 = t_end;
                    } else {
                        float t = (float)(i + 1u) / n;
                        float s = t;
                        if (robust != 1) {
                            float u = integral * t + int0;
                            float inv;
                            if (robust == 2) inv = vb_pow_2_3(fabsf(u)) * vb_signf(u);
                            else inv = espc_int_inv_approx(u);
                            s = (inv - b) / a;
                        }
                        lp1 = es_seg_eval_with_offset(this_p0, this_pq1.p, ep, s, normalized_offset);
                    }
                    fv2 l0 = offset >= 0.f ? lp0 : lp1;
                    fv2 l1 = offset >= 0.f ? lp1 : lp0;
                    f.line_xf(path_ix, l0, l1, transform);
                    lp0 = lp1;
                }
            } else {
                f.ix += n_u;
            }
            last_p = this_pq1.p;
            last_q = this_pq1.q;
            last_t = t1;
            t0_u += 1u;
            uint32_t shift = (uint32_t)(__ffs((int)t0_u) - 1);
            t0_u >>= shift;
            dt *= (float)(1u << shift);
        } else {
            t0_u = t0_u * 2u;
            dt *= 0.5f;
        }
    }
}

template <int EMIT>
__device__ void flatten_arc(Flat<EMIT> &f, uint32_t path_ix, fv2 begin, fv2 end, fv2 center, float angle, const FXform &t) {
    fv2 p0 = fx_apply(t, begin);
    fv2 r = begin - center;
    const float MIN_THETA = 0.0001f;
    const float tol = 0.25f;
    float radius = fmaxf(tol, flen(p0 - fx_apply(t, center)));
    float theta = fmaxf(MIN_THETA, 2.f * vb_acosf(1.f - tol / radius));
    uint32_t n_lines = max(1u, vb_f2u_sat(ceilf(angle / theta)));
    if (EMIT == 0) { f.ix += n_lines; return; }
    float s, c;
    vb_sincosf(theta, &s, &c);
    for (uint32_t i = 0u; i + 1u < n_lines; i++) {
        r = F2(c * r.x + s * r.y, -s * r.x + c * r.y);
        fv2 p1 = fx_apply(t, center + r);
        f.write_line(path_ix, p0, p1);
        p0 = p1;
    }
    fv2 p1 = fx_apply(t, end);
    f.write_line(path_ix, p0, p1);
}

#define STYLE_FLAGS_STYLE 0x80000000u
#define STYLE_FLAGS_FILL 0x40000000u
#define STYLE_MITER_LIMIT_MASK 0xFFFFu
#define STYLE_FLAGS_START_CAP_MASK 0x0C000000u
#define STYLE_FLAGS_END_CAP_MASK 0x03000000u
#define STYLE_FLAGS_CAP_SQUARE 0x01000000u
#define STYLE_FLAGS_CAP_ROUND 0x02000000u
#define STYLE_FLAGS_JOIN_MASK 0x30000000u
#define STYLE_FLAGS_JOIN_BEVEL 0u
#define STYLE_FLAGS_JOIN_MITER 0x10000000u
#define STYLE_FLAGS_JOIN_ROUND 0x20000000u

template <int EMIT>
__device__ void draw_cap(Flat<EMIT> &f, uint32_t path_ix, uint32_t cap_style, fv2 point, fv2 cap0, fv2 cap1, fv2 offset_tangent,
                         const FXform &t) { // flatten.wgsl:521-545 (slot order of the WGSL)
    if (cap_style == STYLE_FLAGS_CAP_ROUND) {
        flatten_arc<EMIT>(f, path_ix, cap0, cap1, point, 3.1415927f, t);
        return;
    }
    fv2 start = cap0, end = cap1;
    if (cap_style == STYLE_FLAGS_CAP_SQUARE) {
        fv2 v = offset_tangent;
        fv2 p0 = start + v;
        fv2 p1 = end + v;
        f.line_xf(path_ix, p0, p1, t);
        f.line_xf(path_ix, start, p0, t);
        f.line_xf(path_ix, p1, end, t);
        return;
    }
    f.line_xf(path_ix, start, end, t);
}

__device__ __forceinline__ float f16_bits_to_f32(uint32_t h) { return __half2float(__ushort_as_half((unsigned short)(h & 0xffffu))); }

template <int EMIT>
__device__ void draw_join(Flat<EMIT> &f, uint32_t path_ix, uint32_t style_flags, fv2 p0, fv2 tan_prev, fv2 tan_next, fv2 n_prev,
                          fv2 n_next, const FXform &t) { // flatten.wgsl:547-631
    fv2 front0 = p0 + n_prev;
    fv2 front1 = p0 + n_next;
    fv2 back0 = p0 - n_next;
    fv2 back1 = p0 - n_prev;
    float cr = tan_prev.x * tan_next.y - tan_prev.y * tan_next.x;
    float d = fdot(tan_prev, tan_next);
    switch (style_flags & STYLE_FLAGS_JOIN_MASK) {
    case STYLE_FLAGS_JOIN_BEVEL:
        f.line_xf(path_ix, front0, front1, t);
        f.line_xf(path_ix, back0, back1, t);
        break;
    case STYLE_FLAGS_JOIN_MITER: {
        float hyp = flen(F2(cr, d));
        float miter_limit = f16_bits_to_f32(style_flags & STYLE_MITER_LIMIT_MASK);
        if (2.f * hyp < (hyp + d) * miter_limit * miter_limit && fabsf(cr) > TANGENT_THRESH * TANGENT_THRESH) {
            bool is_backside = cr > 0.f;
            fv2 fp_last = is_backside ? back1 : front0;
            fv2 fp_this = is_backside ? back0 : front1;
            fv2 p = is_backside ? back0 : front0;
            fv2 v = fp_this - fp_last;
            float h = (tan_prev.x * v.y - tan_prev.y * v.x) / cr;
            fv2 miter_pt = fp_this - tan_next * h;
            f.line_xf(path_ix, p, miter_pt, t);
            if (is_backside) back0 = miter_pt; else front0 = miter_pt;
        }
        f.line_xf(path_ix, front0, front1, t);
        f.line_xf(path_ix, back0, back1, t);
        break;
    }
    case STYLE_FLAGS_JOIN_ROUND: {
        fv2 arc0, arc1, other0, other1;
        if (cr > 0.f) { arc0 = back0; arc1 = back1; other0 = front0; other1 = front1; }
        else { arc0 = front0; arc1 = front1; other0 = back0; other1 = back1; }
        flatten_arc<EMIT>(f, path_ix, arc0, arc1, p0, fabsf(vb_atan2f(cr, d)), t);
        f.line_xf(path_ix, other0, other1, t);
        break;
    }
    default: break;
    }
}

struct PathTagData { uint32_t tag_byte; uint32_t trans_ix, pathseg_offset, style_ix, path_ix; };

__device__ __forceinline__ void fl_reduce_prefix(uint32_t w, uint32_t &trans, uint32_t &off, uint32_t &style, uint32_t &path) {
    // reduce_tag (shared/pathtag.wgsl:58-71) of the bytes below this thread's byte
    uint32_t point_count = w & 0x3030303u;
    trans = __popc(w & (0x20u * 0x1010101u));
    uint32_t n_points = point_count + ((w >> 2) & 0x1010101u);
    uint32_t a = n_points + (n_points & (((w >> 3) & 0x1010101u) * 15u));
    a += a >> 8;
    a += a >> 16;
    off = a & 0xffu;
    path = __popc(w & (0x10u * 0x1010101u));
    style = __popc(w & (0x40u * 0x1010101u)) * 2u;
}

__device__ PathTagData compute_tag_monoid(const VbConfig &cfg, const uint32_t *__restrict__ scene,
                                          const VbTagMonoid *__restrict__ tag_monoids, uint32_t ix) { // flatten.wgsl:683-699
    PathTagData r;
    uint32_t wi = ix >> 2;
    if (wi >= cfg.n_tag_words) { // one past the padded stream: an all-zero tag
        r.tag_byte = 0; r.trans_ix = 0; r.pathseg_offset = 0; r.style_ix = 0; r.path_ix = 0;
        return r;
    }
    uint32_t tag_word = __ldg(scene + cfg.layout.path_tag_base + wi);
    uint32_t shift = (ix & 3u) * 8u;
    uint32_t tr, of, st, pa;
    fl_reduce_prefix(tag_word & ((1u << shift) - 1u), tr, of, st, pa);
    VbTagMonoid base = tag_monoids[wi];
    r.tag_byte = (tag_word >> shift) & 0xffu;
    r.trans_ix = base.trans_ix + tr - 1u;
    r.pathseg_offset = base.pathseg_offset + of;
    r.style_ix = base.style_ix + st - 2u;
    r.path_ix = base.path_ix + pa;
    return r;
}

__device__ __forceinline__ fv2 read_f32_point(const VbConfig &cfg, const uint32_t *__restrict__ scene, uint32_t ix) {
    uint32_t b = cfg.layout.path_data_base + ix;
    return F2(__uint_as_float(vb_scene(scene, cfg, b)), __uint_as_float(vb_scene(scene, cfg, b + 1)));
}
__device__ __forceinline__ fv2 read_i16_point(const VbConfig &cfg, const uint32_t *__restrict__ scene, uint32_t ix) {
    uint32_t raw = vb_scene(scene, cfg, cfg.layout.path_data_base + ix);
    return F2((float)(((int32_t)(raw << 16)) >> 16), (float)(((int32_t)raw) >> 16));
}

__device__ CubicPoints read_path_segment(const VbConfig &cfg, const uint32_t *__restrict__ scene, const PathTagData &tag,
                                         bool is_stroke) { // flatten.wgsl:708-766
    fv2 p0, p1, p2 = F2(0, 0), p3 = F2(0, 0);
    uint32_t seg_type = tag.tag_byte & 3u;
    uint32_t off = tag.pathseg_offset;
    bool is_stroke_cap_marker = is_stroke && (tag.tag_byte & 4u) != 0u;
    bool is_open = seg_type == 2u;
    if (tag.tag_byte & 8u) {
        p0 = read_f32_point(cfg, scene, off);
        p1 = read_f32_point(cfg, scene, off + 2u);
        if (seg_type >= 2u) {
            p2 = read_f32_point(cfg, scene, off + 4u);
            if (seg_type == 3u) p3 = read_f32_point(cfg, scene, off + 6u);
        }
    } else {
        p0 = read_i16_point(cfg, scene, off);
        p1 = read_i16_point(cfg, scene, off + 1u);
        if (seg_type >= 2u) {
            p2 = read_i16_point(cfg, scene, off + 2u);
            if (seg_type == 3u) p3 = read_i16_point(cfg, scene, off + 3u);
        }
    }
    if (is_stroke_cap_marker && is_open) {
        p0 = p1;
        p1 = p2;
        seg_type = 1u;
    }
    const float third = 1.0f / 3.0f;
    if (seg_type == 1u) {
        p3 = p1;
        p2 = p3 + (p0 - p3) * third;
        p1 = p0 + (p3 - p0) * third;
    } else if (seg_type == 2u) {
        p3 = p2;
        p2 = p1 + (p2 - p1) * third;
        p1 = p1 + (p0 - p1) * third;
    }
    CubicPoints r = {p0, p1, p2, p3};
    return r;
}

// Everything one tag byte produces. EMIT=false only counts lines.
template <int EMIT>
__device__ void flatten_tag(Flat<EMIT> &f, const VbConfig &cfg, const uint32_t *__restrict__ scene,
                            const VbTagMonoid *__restrict__ tag_monoids, const PathTagData &tag, uint32_t ix, uint32_t style_flags) {
    uint32_t seg_type = tag.tag_byte & 3u;
    if (seg_type == 0u) return;
    const uint32_t path_ix = tag.path_ix;
    bool is_stroke = (style_flags & STYLE_FLAGS_STYLE) != 0u;
    FXform transform;
    {
        uint32_t b = cfg.layout.transform_base + tag.trans_ix * 6u;
        transform.m0 = __uint_as_float(vb_scene(scene, cfg, b));
        transform.m1 = __uint_as_float(vb_scene(scene, cfg, b + 1));
        transform.m2 = __uint_as_float(vb_scene(scene, cfg, b + 2));
        transform.m3 = __uint_as_float(vb_scene(scene, cfg, b + 3));
        transform.tx = __uint_as_float(vb_scene(scene, cfg, b + 4));
        transform.ty = __uint_as_float(vb_scene(scene, cfg, b + 5));
    }
    CubicPoints pts = read_path_segment(cfg, scene, tag, is_stroke);
    if (is_stroke) {
        float linewidth = __uint_as_float(vb_scene(scene, cfg, cfg.layout.style_base + tag.style_ix + 1u));
        float offset = 0.5f * linewidth;
        bool is_open = seg_type != 1u;
        bool is_stroke_cap_marker = (tag.tag_byte & 4u) != 0u;
        if (is_stroke_cap_marker) {
            if (is_open) {
                fv2 tangent = pts.p3 - pts.p0;
                fv2 offset_tangent = fnorm(tangent) * offset;
                fv2 n = F2(-offset_tangent.y, offset_tangent.x);
                draw_cap<EMIT>(f, path_ix, (style_flags & STYLE_FLAGS_START_CAP_MASK) >> 2, pts.p0, pts.p0 - n, pts.p0 + n,
                               F2(-offset_tangent.x, -offset_tangent.y), transform);
            }
        } else {
            PathTagData ntag = compute_tag_monoid(cfg, scene, tag_monoids, ix + 1u);
            CubicPoints npts = read_path_segment(cfg, scene, ntag, true);
            bool n_is_closed = (ntag.tag_byte & 3u) == 1u;
            bool n_is_marker = (ntag.tag_byte & 4u) != 0u;
            bool do_join = !n_is_marker || n_is_closed;
            fv2 n_tangent = npts.p3 - npts.p0;
            if (!n_is_marker) n_tangent = cubic_start_tangent(npts.p0, npts.p1, npts.p2, npts.p3);
            fv2 tan_start = cubic_start_tangent(pts.p0, pts.p1, pts.p2, pts.p3);
            if (fdot(tan_start, tan_start) < TANGENT_THRESH * TANGENT_THRESH) tan_start = F2(TANGENT_THRESH, 0.f);
            fv2 tan_prev = cubic_end_tangent(pts.p0, pts.p1, pts.p2, pts.p3);
            if (fdot(tan_prev, tan_prev) < TANGENT_THRESH * TANGENT_THRESH) tan_prev = F2(TANGENT_THRESH, 0.f);
            fv2 tan_next = n_tangent;
            if (fdot(tan_next, tan_next) < TANGENT_THRESH * TANGENT_THRESH) tan_next = F2(TANGENT_THRESH, 0.f);
            fv2 n_start = fnorm(F2(-tan_start.y, tan_start.x)) * offset;
            fv2 offset_tangent = fnorm(tan_prev) * offset;
            fv2 n_prev = F2(-offset_tangent.y, offset_tangent.x);
            fv2 tnn = fnorm(tan_next) * offset;
            fv2 n_next = F2(-tnn.y, tnn.x);
            flatten_euler<EMIT>(f, pts, path_ix, transform, offset, pts.p0 + n_start, pts.p3 + n_prev);
            flatten_euler<EMIT>(f, pts, path_ix, transform, -offset, pts.p0 - n_start, pts.p3 - n_prev);
            if (do_join) {
                draw_join<EMIT>(f, path_ix, style_flags, pts.p3, tan_prev, tan_next, n_prev, n_next, transform);
            } else {
                draw_cap<EMIT>(f, path_ix, style_flags & STYLE_FLAGS_END_CAP_MASK, pts.p3, pts.p3 + n_prev, pts.p3 - n_prev,
                               offset_tangent, transform);
            }
        }
    } else {
        // Fast path for a line-to in a fill (the bulk of map-like scenes). For a degree-raised line the general
        // algorithm provably accepts the whole range at the first step and emits exactly ONE line (p0', p3'):
        //  * err = O(angle^2) * chord, and the tangent angles of a degree-raised line are pure rounding noise
        //    (~3 ulp(coord)/chord); for chords shorter than that noise err <= 2 * chord <= 0.024 px  -> accepted;
        //  * n = ceil(n_frac * sqrt(chord/2)) with n_frac <= sqrt(|k0|) ~ sqrt(3 ulp/chord) -> n_frac*mult <= sqrt(1.5 ulp) < 1;
        //  * the single line runs from t_start = p0' to t_end = p3' (t1 == 1 exactly).
        // The bounds need ulp(coord) <= 2^-8, hence the |coord| < 65536 guard; anything else takes the general path.
        // tests/test_gpu_parity.py compares `lines` bit-for-bit with the oracle, which has no such shortcut.
        if (seg_type == 1u) {
            const fv2 q0 = fx_apply(transform, pts.p0), q1 = fx_apply(transform, pts.p1);
            const fv2 q2 = fx_apply(transform, pts.p2), q3 = fx_apply(transform, pts.p3);
            const float lim = 65536.0f;
            if (fabsf(q0.x) < lim && fabsf(q0.y) < lim && fabsf(q3.x) < lim && fabsf(q3.y) < lim && fabsf(q1.x) < lim && fabsf(q1.y) < lim &&
                fabsf(q2.x) < lim && fabsf(q2.y) < lim) {
                if (feq(q0, q1) && feq(q0, q2) && feq(q0, q3)) return;
                if (EMIT != 0) f.write_line(path_ix, q0, q3);
                else f.ix++;
                return;
            }
        }
        flatten_euler<EMIT>(f, pts, path_ix, transform, 0.f, pts.p0, pts.p3);
    }
}

// bbox_clear.wgsl: path bboxes start at (+INT_MAX, -INT_MAX)
__global__ void k_bbox_clear(uint32_t n_paths, VbPathBbox *path_bboxes) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_paths) {
        VbPathBbox b;
        b.x0 = 0x7fffffff; b.y0 = 0x7fffffff; b.x1 = (int32_t)0x80000000; b.y1 = (int32_t)0x80000000;
        b.draw_flags = 0; b.trans_ix = 0;
        path_bboxes[i] = b;
    }
}

// Three small kernels instead of one look-back pass. flatten's per-thread work varies by 10-100x (a fill line vs a
// stroked curve with round joins); with a single-pass look-back every warp that has finished counting sits on its
// registers until ALL earlier partitions have published, so the slowest tag in flight gates the whole machine
// (ncu r1: 18 % issue utilisation, stalls = barrier / look-back spin). Instead:
//   A  k_flatten        : each thread flattens its tag ONCE, lines go to a scratch arena at an offset taken with one
//                         atomicAdd per warp (lane order inside the warp = tag order); per-warp {count, scratch offset}
//   B  k_flatten_scan   : exclusive scan of the per-warp counts (one CTA; ~50k values)
//   C  k_flatten_reorder: per warp, block-copy scratch[src .. src+count) -> lines[dst ..): final order = tag order,
//                         deterministic and identical to the serial CPU shader, independent of the atomics' order.
#define FL_MAX_CACHE 12 // lines a thread keeps in registers/local before it learns its offset; beyond that: 2nd pass
__global__ void __launch_bounds__(FL_THREADS)
k_flatten(VbConfig cfg, const uint32_t *__restrict__ scene, const VbTagMonoid *__restrict__ tag_monoids,
          VbPathBbox *path_bboxes, VbBump *bump, VbLineSoup *scratch, uint32_t *part_count, uint32_t *part_src, uint32_t *scratch_ctr,
          uint32_t n_parts) {
    const uint32_t lane = vb_lane();
    const uint32_t part = blockIdx.x * (FL_THREADS / 32) + (threadIdx.x >> 5);
    if (part >= n_parts) return;
    const uint32_t ix = part * 32u + lane;
    const uint32_t n_tags = cfg.n_tag_words * 4u;
    const uint32_t n_paths = cfg.layout.n_paths;
    PathTagData tag;
    tag.tag_byte = 0; tag.trans_ix = 0; tag.pathseg_offset = 0; tag.style_ix = 0; tag.path_ix = 0;
    uint32_t style_flags = 0;
    if (ix < n_tags) {
        tag = compute_tag_monoid(cfg, scene, tag_monoids, ix);
        style_flags = vb_scene(scene, cfg, cfg.layout.style_base + tag.style_ix);
        if ((tag.tag_byte & 0x10u) != 0u && tag.path_ix < n_paths) {
            path_bboxes[tag.path_ix].draw_flags = (style_flags & STYLE_FLAGS_FILL) == 0u ? 0u : 1u;
            path_bboxes[tag.path_ix].trans_ix = tag.trans_ix;
        }
    }
    // pass 1: count (cheap for the common tags: the geometry of the first FL_CACHE lines is kept in shared memory)
    __shared__ float4 sh_cache[FL_CACHE][FL_THREADS];
    Flat<1> fc;
    fc.lines = nullptr; fc.lines_size = 0; fc.ix = 0;
    fc.bx0 = 1e31f; fc.by0 = 1e31f; fc.bx1 = -1e31f; fc.by1 = -1e31f;
    fc.cache = &sh_cache[0][threadIdx.x];
    flatten_tag<1>(fc, cfg, scene, tag_monoids, tag, ix, style_flags);
    __syncwarp();
    const uint32_t incl = vb_warp_incl_scan(fc.ix);
    const uint32_t total = __shfl_sync(VB_FULL, incl, 31);
    uint32_t base = 0u;
    if (lane == 31u && total != 0u) base = atomicAdd(scratch_ctr, total);
    base = __shfl_sync(VB_FULL, base, 31);
    if (lane == 0u) {
        part_count[part] = total;
        part_src[part] = base;
    }
    if (fc.ix != 0u) {
        const uint32_t out0 = base + incl - fc.ix;
        float bx0 = fc.bx0, by0 = fc.by0, bx1 = fc.bx1, by1 = fc.by1;
        if (fc.ix <= FL_CACHE) {
            for (uint32_t k = 0; k < fc.ix; k++) {
                const uint32_t o = out0 + k;
                if (o < cfg.lines_size) {
                    const float4 l = sh_cache[k][threadIdx.x];
                    uint2 *dst = reinterpret_cast<uint2 *>(scratch + o);
                    dst[0] = make_uint2(tag.path_ix, 0u);
                    dst[1] = make_uint2(__float_as_uint(l.x), __float_as_uint(l.y));
                    dst[2] = make_uint2(__float_as_uint(l.z), __float_as_uint(l.w));
                }
            }
        } else {
            Flat<2> fe;
            fe.lines = scratch; fe.lines_size = cfg.lines_size; fe.ix = out0;
            fe.bx0 = 1e31f; fe.by0 = 1e31f; fe.bx1 = -1e31f; fe.by1 = -1e31f;
            fe.cache = nullptr;
            flatten_tag<2>(fe, cfg, scene, tag_monoids, tag, ix, style_flags);
            bx0 = fe.bx0; by0 = fe.by0; bx1 = fe.bx1; by1 = fe.by1;
        }
        if ((bx1 > bx0 || by1 > by0) && tag.path_ix < n_paths) {
            VbPathBbox *o = path_bboxes + tag.path_ix;
            atomicMin(&o->x0, vb_f2i_sat(floorf(bx0)));
            atomicMin(&o->y0, vb_f2i_sat(floorf(by0)));
            atomicMax(&o->x1, vb_f2i_sat(ceilf(bx1)));
            atomicMax(&o->y1, vb_f2i_sat(ceilf(by1)));
        }
    }
}

// B: exclusive scan of part_count -> destination offsets; publishes bump.lines. One CTA, 8 values per thread per pass.
#define FS_THREADS 1024
#define FS_PER_THREAD 8
__global__ void __launch_bounds__(FS_THREADS)
k_flatten_scan(VbConfig cfg, uint32_t n_parts, const uint32_t *__restrict__ part_count, uint32_t *part_dst, VbBump *bump) {
    __shared__ uint32_t sh_scan[FS_THREADS / 32 + 2];
    uint32_t carry = 0u;
    for (uint32_t base = 0u; base < n_parts; base += FS_THREADS * FS_PER_THREAD) {
        const uint32_t i0 = base + threadIdx.x * FS_PER_THREAD;
        uint32_t v[FS_PER_THREAD], sum = 0u;
#pragma unroll
        for (int k = 0; k < FS_PER_THREAD; k++) {
            v[k] = i0 + k < n_parts ? part_count[i0 + k] : 0u;
            sum += v[k];
        }
        uint32_t total;
        uint32_t run = carry + vb_block_excl_scan(sum, sh_scan, &total);
#pragma unroll
        for (int k = 0; k < FS_PER_THREAD; k++) {
            if (i0 + k < n_parts) part_dst[i0 + k] = run;
            run += v[k];
        }
        carry += total;
    }
    if (threadIdx.x == 0) {
        bump->lines = carry;
        if (carry > cfg.lines_size) atomicOr(&bump->failed, VB_STAGE_FLATTEN);
    }
}

// C: per partition, copy its block of lines from the scratch arena to its final position (8-byte words, coalesced).
#define FR_THREADS 256
__global__ void __launch_bounds__(FR_THREADS)
k_flatten_reorder(VbConfig cfg, uint32_t n_parts, const uint32_t *__restrict__ part_count, const uint32_t *__restrict__ part_src,
                  const uint32_t *__restrict__ part_dst, const VbLineSoup *__restrict__ scratch, VbLineSoup *lines) {
    const uint32_t lane = vb_lane();
    const uint32_t warps = gridDim.x * (FR_THREADS / 32);
    for (uint32_t part = blockIdx.x * (FR_THREADS / 32) + (threadIdx.x >> 5); part < n_parts; part += warps) {
        const uint32_t n = part_count[part];
        if (n == 0u) continue;
        const uint32_t src = part_src[part], dst = part_dst[part];
        if (src + n > cfg.lines_size || dst + n > cfg.lines_size) continue; // overflowed frame: will be re-run
        const uint2 *s = reinterpret_cast<const uint2 *>(scratch + src);
        uint2 *d = reinterpret_cast<uint2 *>(lines + dst);
        for (uint32_t k = lane; k < n * 3u; k += 32u) d[k] = __ldg(s + k);
    }
}

extern "C" void vb_launch_flatten(const VbConfig *cfg, const uint32_t *scene, const VbTagMonoid *tag_monoids,
                                  VbPathBbox *path_bboxes, VbBump *bump, VbLineSoup *lines, VbLineSoup *scratch, uint32_t *part_mem /* 3*n_parts */,
                                  uint32_t *scratch_ctr, uint32_t n_parts, cudaStream_t st) {
    uint32_t n_paths = cfg->layout.n_paths;
    if (n_paths) k_bbox_clear<<<(n_paths + 255) / 256, 256, 0, st>>>(n_paths, path_bboxes);
    if (n_parts) {
        uint32_t *part_count = part_mem, *part_src = part_mem + n_parts, *part_dst = part_mem + 2 * (size_t)n_parts;
        const uint32_t warps_per_cta = FL_THREADS / 32;
        k_flatten<<<(n_parts + warps_per_cta - 1) / warps_per_cta, FL_THREADS, 0, st>>>(*cfg, scene, tag_monoids, path_bboxes, bump, scratch,
                                                                                       part_count, part_src, scratch_ctr, n_parts);
        k_flatten_scan<<<1, FS_THREADS, 0, st>>>(*cfg, n_parts, part_count, part_dst, bump);
        uint32_t blocks = (n_parts + (FR_THREADS / 32) - 1) / (FR_THREADS / 32);
        if (blocks > 148u * 8u) blocks = 148u * 8u;
        k_flatten_reorder<<<blocks, FR_THREADS, 0, st>>>(*cfg, n_parts, part_count, part_src, part_dst, scratch, lines);
    }
}
extern "C" uint32_t vb_flatten_parts(uint32_t n_tag_words) { return (n_tag_words * 4u + 31u) / 32u; }
