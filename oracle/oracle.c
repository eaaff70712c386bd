/* oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE ONLY; see oracle.h for scope and the
 * "parity unpinned" statement).  Plain C11 + OpenMP, compiled with -ffp-contract=off.
 */
#include "oracle.h"
#include "tableaus_gen.h"
#include <math.h>
#include <omp.h>
#include <string.h>
#include <stddef.h>

/* ---------------------------------------------------------------- fast log2 / exp2
 * The PI controller needs EEst^beta1 / qold^beta2.  Upstream evaluates both powers with an
 * approximate Float32 routine (SURVEY.md 7.3 "fastpower", [UPSTREAM-RECALLED] FastPower.jl);
 * this restatement keeps that character -- Float32, ~2e-6 accurate -- but spells both halves
 * out in IEEE-exact primitive operations (integer exponent extraction, FMA-Horner polynomials)
 * so that the CUDA kernels reproduce them bit for bit:
 *   log2(x) = e + t*P5(t),  t = mantissa-1 in [0,1)     (max abs error 2.2e-6)
 *   2^y     = 2^rint(y) * Q6(y - rint(y)),  Q6 = degree-6 Taylor of exp(f ln 2), |f|<=1/2
 * The controller works in the log domain (one log2 of EEst^2, one exp2 per step); see
 * oracle_impl.inc for the exact expression tree.
 */
float orc_fastlog2(float x) {
    uint32_t ix;
    memcpy(&ix, &x, 4);
    const int e = (int)(ix >> 23) - 127;
    const uint32_t im = (ix & 0x007fffffu) | 0x3f800000u;
    float m;
    memcpy(&m, &im, 4);
    const float t = m - 1.0f;
    float p = -0.02645725943148136f;
    p = fmaf(p, t, 0.12345092743635178f);
    p = fmaf(p, t, -0.27953752875328064f);
    p = fmaf(p, t, 0.45827049016952515f);
    p = fmaf(p, t, -0.7182818651199341f);
    p = fmaf(p, t, 1.442553162574768f);
    return fmaf(t, p, (float)e);
}
float orc_fastexp2(float y) {
    y = fminf(fmaxf(y, -125.0f), 125.0f);
    const float fi = rintf(y);
    const float f = y - fi;
    float p = 1.5403530e-4f;
    p = fmaf(p, f, 1.3333558e-3f);
    p = fmaf(p, f, 9.6181291e-3f);
    p = fmaf(p, f, 5.5504109e-2f);
    p = fmaf(p, f, 2.4022651e-1f);
    p = fmaf(p, f, 6.9314718e-1f);
    p = fmaf(p, f, 1.0f);
    uint32_t ip;
    memcpy(&ip, &p, 4);
    ip += (uint32_t)((int32_t)fi << 23);
    memcpy(&p, &ip, 4);
    return p;
}
/* Float32 reciprocal used by the error norm (restates b2_rcp_nr, kernels/b2_common.cuh): x0 = bits(0x7EF311C7 - bits(s))
 * (relative error <= 5.1 %), two Newton steps x <- fma(x, fma(-s, x, 1), x) -> relative error <= 6.6e-6.  EEst only
 * steers the step size; the spec'd arithmetic (IEEE division) is orc_opts.spec_arith. */
float orc_rcp_nr(float s) {
    uint32_t is;
    memcpy(&is, &s, 4);
    const uint32_t ix = 0x7EF311C7u - is;
    float x;
    memcpy(&x, &ix, 4);
    x = fmaf(x, fmaf(-s, x, 1.0f), x);
    x = fmaf(x, fmaf(-s, x, 1.0f), x);
    return x;
}

/* theta = (tau - tprev) / dt of a saveat point (restates b2_theta): Float32 multiplies by a three-step Newton
 * reciprocal (relative error <= 9e-8, under one ulp of theta), Float64 divides. */
static float theta_f32(float num, float dt) {
    uint32_t is;
    memcpy(&is, &dt, 4);
    const uint32_t ix = 0x7EF311C7u - is;
    float x;
    memcpy(&x, &ix, 4);
    x = fmaf(x, fmaf(-dt, x, 1.0f), x);
    x = fmaf(x, fmaf(-dt, x, 1.0f), x);
    x = fmaf(x, fmaf(-dt, x, 1.0f), x);
    return num * x;
}
static double theta_f64(double num, double dt) { return num / dt; }

/* PI controller in the log domain, Float32, division-free (restates b2_pi_controller, kernels/b2_control.cuh):
 *   l = log2(EEst) = 0.5*log2(EEst^2), lq = log2(qold)
 *   accept (EEst <= 1): dt_next  = dt * clamp(gamma * 2^(beta2*lq - beta1*l), qmin, qmax)   [= dt / q of A.5]
 *   reject            : dt_retry = dt * max(gamma * 2^(-beta1*l), qmin)                      [= dt / min(1/qmin, q11/gamma)]
 *   EEst == 0         : dt_next  = dt * qmax */
typedef struct { float qmin, qmax, gam, beta1, beta2, lqinit; } orc_ctl;
typedef struct { int ok, isn, zero; float l, qi; } orc_decision;
static orc_decision orc_pi_controller(float EE2, float lq, const orc_ctl* c) {
    orc_decision d;
    d.isn = EE2 != EE2;
    d.ok = EE2 <= 1.0f;
    d.zero = EE2 == 0.0f;
    d.l = 0.5f * orc_fastlog2(EE2);
    const float nbl = -c->beta1 * d.l;
    float qi = orc_fastexp2(d.ok ? fmaf(c->beta2, lq, nbl) : nbl);
    qi = fmaxf(c->qmin, qi * c->gam);
    qi = d.ok ? fminf(c->qmax, qi) : qi;
    d.qi = d.zero ? c->qmax : qi;
    return d;
}
static float orc_ctl_lq_next(const orc_decision* d, const orc_ctl* c) { return fmaxf(d->zero ? c->lqinit : d->l, c->lqinit); }

float orc_fastpow(float x, float y) {
    if (!(x > 0.0f)) return 0.0f;
    return orc_fastexp2(y * orc_fastlog2(x));
}

/* ---------------------------------------------------------------- Philox4x32-10 (B.9) */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
/* Every trajectory consumes ONE continuous stream of standard normals; normal number q of trajectory i is
 * element q % K of the Philox block with counter = (i_lo, i_hi, b_lo, b_hi), b = q / K, key = (seed_lo, seed_hi),
 * K = normals per Philox call (4 in f32, 2 in f64).  A step that needs m normals takes the next m of the
 * stream, so no generated normal is thrown away (SURVEY 7.3 "Integer Philox competes with FMA").
 * f32: four u32 -> two Box-Muller pairs -> 4 normals; uniforms ((x>>8)+0.5)*2^-24 in (0,1).
 * f64: two 53-bit uniforms ((x>>11)+0.5)*2^-53 -> one pair -> 2 normals. */
#define TWO_PI 6.283185307179586476925
void orc_normals_f32(uint64_t seed, uint64_t traj, uint64_t block, float z[4]) {
    const uint32_t ctr[4] = {(uint32_t)traj, (uint32_t)(traj >> 32), (uint32_t)block, (uint32_t)(block >> 32)};
    const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t r[4];
    orc_philox4x32_10(ctr, key, r);
    for (int h = 0; h < 2; h++) {
        const float u1 = ((float)(r[2 * h] >> 8) + 0.5f) * 5.9604644775390625e-8f;
        const float u2 = ((float)(r[2 * h + 1] >> 8) + 0.5f) * 5.9604644775390625e-8f;
        const float rad = sqrtf(-2.0f * logf(u1));
        /* cos / sin of the angle 2 pi u2 correctly rounded to Float32 (evaluated in double): the kernels use
         * sincospif(2 u2), which is accurate to ~1 ulp of the same true value */
        const double ang = TWO_PI * (double)u2;
        z[2 * h] = rad * (float)cos(ang);
        z[2 * h + 1] = rad * (float)sin(ang);
    }
}
void orc_normals_f64(uint64_t seed, uint64_t traj, uint64_t block, double z[2]) {
    const uint32_t ctr[4] = {(uint32_t)traj, (uint32_t)(traj >> 32), (uint32_t)block, (uint32_t)(block >> 32)};
    const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t r[4];
    orc_philox4x32_10(ctr, key, r);
    const uint64_t a = ((uint64_t)r[0] << 32) | r[1], b = ((uint64_t)r[2] << 32) | r[3];
    const double u1 = ((double)(a >> 11) + 0.5) * 1.1102230246251565404e-16;
    const double u2 = ((double)(b >> 11) + 0.5) * 1.1102230246251565404e-16;
    const double rad = sqrt(-2.0 * log(u1));
    const double ang = TWO_PI * u2;
    z[0] = rad * cos(ang);
    z[1] = rad * sin(ang);
}
int orc_max_threads(void) { return omp_get_max_threads(); }

/* ---------------------------------------------------------------- f64 instantiation */
#define REAL double
#define SUF(x) x##_f64
#define FMA fma
#define FABS fabs
#define FMAX fmax
#define FMIN fmin
#define SQRT sqrt
#define POW pow
#define REAL_EPS 2.220446049250313e-16
#define NORMALS_PER_CALL 2
#include "oracle_impl.inc"
#undef REAL
#undef SUF
#undef FMA
#undef FABS
#undef FMAX
#undef FMIN
#undef SQRT
#undef POW
#undef REAL_EPS
#undef NORMALS_PER_CALL

/* ---------------------------------------------------------------- f32 instantiation */
#define REAL float
#define SUF(x) x##_f32
#define FMA fmaf
#define FABS fabsf
#define FMAX fmaxf
#define FMIN fminf
#define SQRT sqrtf
#define POW powf
#define REAL_EPS 1.1920928955078125e-7f
#define NORMALS_PER_CALL 4
#include "oracle_impl.inc"
