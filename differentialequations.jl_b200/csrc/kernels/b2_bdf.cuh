// b2_bdf.cuh -- FBDF: variable-order (1..5), variable-step BDF in fixed-leading-coefficient form with a per-thread
// register LU and the analytic Jacobian b2_jac.  Reference name: FBDF /root/reference/test/qa/qa.jl:57 (re-exported from
// OrdinaryDiffEq).  Restated from the published algorithm [UPSTREAM-RECALLED structure: the last
// k+1 solution values kept at their own times, re-sampled on the equidistant grid of the current dt by Lagrange
// interpolation, constant-step BDF-k coefficients, divided-difference error and order estimates]; the written contract
// is oracle/oracle_impl.inc (fbdf_step / fbdf_accept / fbdf_reject / fbdf_push), mirrored here operation for operation.
//
//   history   times x_j (x_0 = t, newest first) and the Newton divided differences dd[l] = [x_0..x_l]u -- in registers
//   predictor the interpolation polynomial of (x_0..x_k) at t + dt, Horner on the Newton form (first step: u0 = uprev)
//   corrector z + tmp = beta_k dt f(z, t + dt), tmp = a_1 uprev + sum_s a_{s+1} P(t - s dt); simplified Newton, W = I - beta dt J(uprev, t)
//             factored once per step, convergence-rate test (kappa = 1/100, <= 10 iterations)
//   estimates terk_m = m! dt^m [t+dt, x_0..x_{m-1}]u (one new diagonal of the Newton table per step: k+1 divisions), lte = -(1/(k+1) + sum_j (a_j/beta) r_j) terk_{k+1}
//   control   the driver's OWN_CONTROL path (b2_ode_driver.cuh): accept iff ||lte|| <= 1, order +-1 from the scaled
//             norms T_m, dt_new = dt / q with q = (2 T_k/(k+1))^(1/(k+1)) and a steady band, Newton failure: dt/2
// Saveat uses the cubic Hermite interpolant on (uprev, f(uprev)), (u, f(u)) -- upstream's default for multistep methods.
#pragma once
#include "b2_common.cuh"
#include "b2_rosenbrock.cuh"   // B2LU

// [k][0] = beta_k, [k][j] = a_j: u_{n+1} + sum_j a_j u_{n+1-j} = beta dt f(u_{n+1}).  Read with a per-lane order k:
// plain global constants (a __constant__ bank would serialise the lanes of a warp that sit at different orders).
static __device__ const double B2_BDFC[6][6] = {
    {0, 0, 0, 0, 0, 0},
    {1.0, -1.0, 0, 0, 0, 0},
    {2.0 / 3.0, -4.0 / 3.0, 1.0 / 3.0, 0, 0, 0},
    {6.0 / 11.0, -18.0 / 11.0, 9.0 / 11.0, -2.0 / 11.0, 0, 0},
    {12.0 / 25.0, -48.0 / 25.0, 36.0 / 25.0, -16.0 / 25.0, 3.0 / 25.0, 0},
    {60.0 / 137.0, -300.0 / 137.0, 300.0 / 137.0, -200.0 / 137.0, 75.0 / 137.0, -12.0 / 137.0}};
// a_j / beta (the oracle divides the table entries at run time: the same IEEE quotient, folded here at compile time)
static __device__ const double B2_BDFR[6][6] = {
    {0, 0, 0, 0, 0, 0},
    {0, (-1.0) / (1.0), 0, 0, 0, 0},
    {0, (-4.0 / 3.0) / (2.0 / 3.0), (1.0 / 3.0) / (2.0 / 3.0), 0, 0, 0},
    {0, (-18.0 / 11.0) / (6.0 / 11.0), (9.0 / 11.0) / (6.0 / 11.0), (-2.0 / 11.0) / (6.0 / 11.0), 0, 0},
    {0, (-48.0 / 25.0) / (12.0 / 25.0), (36.0 / 25.0) / (12.0 / 25.0), (-16.0 / 25.0) / (12.0 / 25.0), (3.0 / 25.0) / (12.0 / 25.0), 0},
    {0, (-300.0 / 137.0) / (60.0 / 137.0), (300.0 / 137.0) / (60.0 / 137.0), (-200.0 / 137.0) / (60.0 / 137.0), (75.0 / 137.0) / (60.0 / 137.0),
     (-12.0 / 137.0) / (60.0 / 137.0)}};
// 1/(m+1) and log2(m+1) as Float32 constants, m = 0..6
static __device__ const float B2_INVP1[7] = {1.0f, 0.5f, 0.333333343f, 0.25f, 0.2f, 0.166666672f, 0.142857149f};
static __device__ const float B2_LG2P1[7] = {0.0f, 1.0f, 1.5849625f, 2.0f, 2.32192802f, 2.5849625f, 2.80735493f};

// The order k is a run-time value per lane, but every array below is indexed with compile-time constants only (loops
// over the maximal range with the inactive levels masked), so the divided-difference table stays in registers.
struct B2Fbdf {
    static constexpr int ORDER = 1;          // initial-dt exponent and controller defaults (the order itself is run-time)
    static constexpr int DEG = 0;
    static constexpr bool OWN_CONTROL = true;
    __device__ __forceinline__ void bind(real*) {}
    __device__ __forceinline__ void poly_coeffs(int, real (&)[1]) const {}

    real f0[B2_N], fnew[B2_N];
    int k, ncons, consfail, iters, nlfails, Lv;
    float eta_old;
    real x[7];            // times of the history, newest first (x[0] = t)
    real dd[7][B2_N];     // Newton divided differences of the history: dd[l] = [x_0 .. x_l]u
    real inv[7];          // 1 / (t + dt - x_{l-1}) of the last step attempt: accepted() re-forms the new diagonal from them
    float T2[8];          // squared scaled norms of terk_m

    // first step, and after every callback that modified u (upstream: u_modified -> reinitFBDF!): order 1, empty history
    __device__ __forceinline__ void start(const real (&u)[B2_N], const real (&p)[B2_NPA], real t) {
        b2_rhs(f0, u, p, t);
        k = 1;
        ncons = consfail = iters = nlfails = Lv = 0;
        eta_old = 1.0f;
#pragma unroll
        for (int l = 0; l < 7; l++) {
            x[l] = 0;
            inv[l] = 0;
#pragma unroll
            for (int i = 0; i < B2_N; i++) dd[l][i] = 0;
        }
    }
    __device__ __forceinline__ real fsal0(int i) const { return f0[i]; }
    __device__ __forceinline__ float Tat(int m) const {   // T2[m] for a run-time m without indexing memory
        float v = T2[0];
#pragma unroll
        for (int j = 1; j < 8; j++) v = (m == j) ? T2[j] : v;
        return v;
    }

    // the interpolation polynomial of the history (x_0..x_k) at xe, Newton form / Horner: P = dd[k];
    // P = fma(P, xe - x_l, dd[l]) for l = k-1..0.  INVARIANT: dd[l] = 0 for l > k (kept by accepted() / drop_levels(); the
    // levels above the order are never read by the algorithm), so the fixed-length loop needs no masks:
    // fma(0, dx, 0) = 0 and fma(0, dx, dd[k]) = dd[k] exactly -- the variable-length loop's bits.
    __device__ __forceinline__ void poly(real xe, real (&out)[B2_N]) const {
#pragma unroll
        for (int i = 0; i < B2_N; i++) out[i] = 0;
#pragma unroll
        for (int l = 5; l >= 0; l--) {
            const real dx = xe - x[l];
#pragma unroll
            for (int i = 0; i < B2_N; i++) out[i] = b2_fma(out[i], dx, dd[l][i]);
        }
    }
    __device__ __forceinline__ void drop_levels() {   // after the order went down
#pragma unroll
        for (int l = 1; l <= 6; l++) {
#pragma unroll
            for (int i = 0; i < B2_N; i++) dd[l][i] = (l > k) ? (real)0 : dd[l][i];
        }
    }

    // one step attempt; returns true when the Newton iteration failed (un / ut are then meaningless)
    __device__ __forceinline__ bool step(const real (&up)[B2_N], const real (&p)[B2_NPA], real t, real dt, real (&un)[B2_N],
                                         real (&ut)[B2_N], const B2Args& a, int& nf) {
        // the lanes that step enter together (the driver's __syncwarp() + `if (do_step)`); they leave the Newton loop after
        // different iteration counts and are brought back together behind it (ncu: without this the ~1500 instructions
        // behind the loop ran 1.7x per warp-iteration at 16/32 lanes)
        const unsigned am = __activemask();
        const real tdt = t + dt;
        const real beta = (real)B2_BDFC[k][0];
        real z[B2_N], tmp[B2_N];
        if (iters == 0) {
            x[0] = t;
#pragma unroll
            for (int i = 0; i < B2_N; i++) dd[0][i] = up[i];
        }
        // predictor
        if (iters >= 1) {
            poly(tdt, z);
        } else {
#pragma unroll
            for (int i = 0; i < B2_N; i++) z[i] = up[i];
        }
        // tmp = a_1 uprev + sum_s a_{s+1} P(t - s dt): the BDF-k combination of the history re-sampled on the grid of dt
        {
            const real a1 = (real)B2_BDFC[k][1];
#pragma unroll
            for (int i = 0; i < B2_N; i++) tmp[i] = a1 * dd[0][i];
        }
#pragma unroll
        for (int s = 1; s <= 4; s++) {
            if (s <= k - 1) {
                real P[B2_N];
                poly(t - (real)s * dt, P);
                const real as = (real)B2_BDFC[k][s + 1];
#pragma unroll
                for (int i = 0; i < B2_N; i++) tmp[i] = b2_fma(as, P[i], tmp[i]);
            }
        }
        // W = I - beta dt J(uprev, t)
        const real bdt = beta * dt;
        B2LU lu;
        {
            real J[B2_N * B2_N];
            b2_jac(J, up, p, t);
#pragma unroll
            for (int i = 0; i < B2_N; i++)
#pragma unroll
#if B2_HAS_MASS
                // constant mass matrix (M u' = f, possibly singular: index-1 DAE): M (z + tmp) = beta dt f(z), W = M - beta dt J
                for (int j = 0; j < B2_N; j++) lu.A[i][j] = b2_fma(-bdt, J[i * B2_N + j], (real)B2_MASS_[i * B2_N + j]);
#else
                for (int j = 0; j < B2_N; j++) lu.A[i][j] = b2_fma(-bdt, J[i * B2_N + j], (i == j) ? (real)1 : (real)0);
#endif
        }
        lu.factor();
        const float inv_n = __fdiv_rn(1.0f, (float)B2_N), kappa = 0.01f;
        float eta = b2_fastexp2(__fmul_rn(0.8f, b2_fastlog2(fmaxf(eta_old, 1.1920929e-7f))));
        float ndz_prev = 0.0f;
        bool conv = false;
        for (int it = 1; it <= 10; it++) {
            real fz[B2_N], dz[B2_N];
            b2_rhs(fz, z, p, tdt);
            nf++;
#if B2_HAS_MASS
            {
                real v[B2_N];
#pragma unroll
                for (int i = 0; i < B2_N; i++) v[i] = z[i] + tmp[i];
#pragma unroll
                for (int i = 0; i < B2_N; i++) {   // (M v)_i over the non-zero entries of row i in index order
                    real acc = 0;
                    bool first = true;
#pragma unroll
                    for (int j = 0; j < B2_N; j++) {
                        if (B2_MASS_[i * B2_N + j] != 0.0) {
                            acc = first ? (real)B2_MASS_[i * B2_N + j] * v[j] : b2_fma((real)B2_MASS_[i * B2_N + j], v[j], acc);
                            first = false;
                        }
                    }
                    dz[i] = b2_fma(bdt, fz[i], -acc);
                }
            }
#else
#pragma unroll
            for (int i = 0; i < B2_N; i++) dz[i] = b2_fma(bdt, fz[i], -(z[i] + tmp[i]));
#endif
            lu.solve(dz);
            float acc = 0.0f;
#pragma unroll
            for (int i = 0; i < B2_N; i++) {
                const real sk = b2_fma(b2_max(b2_abs(up[i]), b2_abs(z[i])), B2_RTOL(a, i), B2_ATOL(a, i));
                const float r = __fmul_rn((float)dz[i], b2_rcp_nr((float)sk));
                acc = __fmaf_rn(r, r, acc);
            }
            const float ndz = __fsqrt_rn(__fmul_rn(acc, inv_n));
#pragma unroll
            for (int i = 0; i < B2_N; i++) z[i] = z[i] + dz[i];
            if (!(ndz == ndz)) break;                 // NaN: failure
            if (it == 1) {
                if (ndz < 1e-5f) {
                    conv = true;
                    break;
                }
            } else {
                const float theta = __fdiv_rn(ndz, ndz_prev);
                if (theta > 2.0f) break;              // diverging
                if (theta < 1.0f) {
                    float pw = 1.0f;
                    for (int m = 0; m < 10 - it; m++) pw = __fmul_rn(pw, theta);
                    const float om = __fsub_rn(1.0f, theta);
                    if (__fdiv_rn(__fmul_rn(ndz, pw), om) > kappa) break;   // will not get there in the remaining iterations
                    eta = __fdiv_rn(theta, om);
                } else {
                    eta = 1e30f;
                }
            }
            if (__fmul_rn(eta, ndz) < kappa) {
                conv = true;
                break;
            }
            ndz_prev = ndz;
        }
        __syncwarp(am);
        if (!conv) {
            nlfails++;
            return true;
        }
        nlfails = 0;
        eta_old = eta;
#pragma unroll
        for (int i = 0; i < B2_N; i++) un[i] = z[i];
        // divided differences through the new point, one new diagonal of the Newton table:
        //   nd[0] = u, nd[l] = (nd[l-1] - dd[l-1]) / (t + dt - x_{l-1});  terk_l = l! dt^l nd[l]
        // Only the running level is kept; accepted() re-forms the diagonal from the stored reciprocals (same operations).
        float rsk[B2_N];
#pragma unroll
        for (int i = 0; i < B2_N; i++) {
            const real sk = b2_fma(b2_max(b2_abs(up[i]), b2_abs(un[i])), B2_RTOL(a, i), B2_ATOL(a, i));
            rsk[i] = b2_rcp_nr((float)sk);
        }
        Lv = (k + 1 < iters + 1) ? k + 1 : iters + 1;   // levels the history supports (k+1 unless first step)
        real fac = 1, terkp1[B2_N], cur[B2_N];
#pragma unroll
        for (int m = 0; m < 8; m++) T2[m] = 0.0f;
        {
            float acc = 0.0f;
#pragma unroll
            for (int i = 0; i < B2_N; i++) {
                cur[i] = un[i];
                terkp1[i] = 0;
                const float r = __fmul_rn((float)un[i], rsk[i]);
                acc = __fmaf_rn(r, r, acc);
            }
            T2[0] = __fmul_rn(acc, inv_n);
        }
#pragma unroll
        for (int l = 1; l <= 6; l++) {
            if (l <= Lv) {
                const real iv = (real)1 / (tdt - x[l - 1]);
                inv[l] = iv;
                fac = fac * ((real)l * dt);
                float acc = 0.0f;
#pragma unroll
                for (int i = 0; i < B2_N; i++) {
                    cur[i] = (cur[i] - dd[l - 1][i]) * iv;
                    const real v = fac * cur[i];
                    if (l == k + 1) terkp1[i] = v;
                    const float r = __fmul_rn((float)v, rsk[i]);
                    acc = __fmaf_rn(r, r, acc);
                }
                T2[l] = __fmul_rn(acc, inv_n);
            }
        }
        if (Lv < k + 1) {
            // first step (one history point): the error estimate is the change of the solution itself
#pragma unroll
            for (int i = 0; i < B2_N; i++) ut[i] = un[i] - up[i];
        } else {
            real cl = (real)1 / (real)(k + 1);
            const real inv_fac = (real)1 / fac;
#pragma unroll
            for (int j = 2; j <= 5; j++) {
                if (j <= k) {
                    const real xj = t - (real)(j - 1) * dt;
                    real num = 1;
#pragma unroll
                    for (int m = 0; m <= 5; m++)
                        if (m <= k) num = num * (xj - x[m]);
                    cl = b2_fma((real)B2_BDFR[k][j], num * inv_fac, cl);
                }
            }
#pragma unroll
            for (int i = 0; i < B2_N; i++) ut[i] = cl * terkp1[i];
        }
        {   // the order-(k+1) estimate is only trusted after k+2 steps at order k (value selects, not conditional stores:
            // the compiler turns those into a run-time index and the whole table moves to local memory)
            const bool drop = !(ncons > k + 1 && k < 5);
#pragma unroll
            for (int m = 2; m <= 6; m++) T2[m] = (drop && m == k + 1) ? 0.0f : T2[m];
        }
        return false;
    }

    // accepted step: order selection and the next dt (as a multiplier of dt)
    __device__ __forceinline__ float accept(float qmin, float qmax) {
        // bit m of dec: T_{m-1} > T_m.  Order kn keeps / raises when T_{kn-2} > T_{kn-1} > T_kn > T_{kn+1}: bits kn-1, kn, kn+1
        unsigned dec = 0;
#pragma unroll
        for (int m = 1; m <= 7; m++) dec |= (T2[m - 1] > T2[m]) ? (1u << m) : 0u;
        int kn = k;
        if (kn < 5 && ncons >= kn + 2 &&
            ((kn == 1 && (dec & 4u)) || (kn == 2 && (dec & 12u) == 12u) || (kn > 2 && ((dec >> (kn - 1)) & 7u) == 7u))) {
            kn++;
        } else {
            while (kn > 2 && ((dec >> (kn - 1)) & 7u) != 7u) kn--;
        }
        const float terk2 = Tat(kn);
        if (kn != k) ncons = 0;
        k = kn;
        float qi;
        if (terk2 == 0.0f) {
            qi = qmax;
        } else {
            // log2 q = (1 + log2(T_k) - log2(k+1)) / (k+1)
            const float lq = __fmul_rn(__fsub_rn(__fadd_rn(1.0f, __fmul_rn(0.5f, b2_fastlog2(terk2))), B2_LG2P1[kn]), B2_INVP1[kn]);
            qi = (lq >= 0.0f && lq <= 1.0f) ? 1.0f : b2_fastexp2(-lq);   // steady band: 1 <= q <= 2 keeps dt
            qi = fminf(qmax, fmaxf(qmin, qi));
        }
        consfail = 0;
        ncons++;
        iters++;
        return qi;
    }
    // rejected step (||lte|| > 1): new dt multiplier, possibly one order down
    __device__ __forceinline__ float reject(float EE2) {
        const int k0 = k;
        consfail++;
        ncons = 0;
        const float half = consfail > 1 ? 0.5f : 1.0f;
        const float lz = 0.26303440f;   // log2(1.2)
        float l = -__fadd_rn(lz, __fmul_rn(__fmul_rn(0.5f, b2_fastlog2(EE2)), B2_INVP1[k0]));
        if (k0 > 1) {
            const float Tm = Tat(k0 - 1);
            const float lm = Tm > 0.0f ? -__fadd_rn(lz, __fmul_rn(__fmul_rn(0.5f, b2_fastlog2(Tm)), B2_INVP1[k0 - 1])) : 0.0f;
            if (lm > l) {
                l = lm;
                k = k0 - 1;
                drop_levels();
            }
        }
        return __fmul_rn(half, b2_fastexp2(fminf(l, 0.0f)));
    }
    // Newton failure: the driver halves dt; one order down after three failures in a row
    __device__ __forceinline__ void newton_fail() {
        if (k > 1 && nlfails >= 3) {
            k--;
            drop_levels();
        }
        consfail++;
        ncons = 0;
    }
    __device__ __forceinline__ void fixed_accept() { iters++; }   // fixed step: the order stays 1 (backward Euler)

    // f(u_new) (Hermite end slope) and the history push: the new diagonal of the divided-difference table, re-formed in
    // place from the step's reciprocals, becomes the table
    __device__ __forceinline__ void accepted(const real (&u)[B2_N], const real (&p)[B2_NPA], real tnew, int& nf) {
        b2_rhs(fnew, u, p, tnew);
        nf += 1;
#pragma unroll
        for (int j = 6; j >= 1; j--) x[j] = x[j - 1];
        x[0] = tnew;
        real prev[B2_N];
#pragma unroll
        for (int i = 0; i < B2_N; i++) prev[i] = u[i];
        // (monotone conditions `l <= Lv` compile to predicated register moves; a chain of `l == Lv` stores would be merged
        // into one run-time-indexed store and move the table to local memory)
#pragma unroll
        for (int l = 1; l <= 6; l++) {
            if (l <= Lv) {
#pragma unroll
                for (int i = 0; i < B2_N; i++) {
                    const real c = (prev[i] - dd[l - 1][i]) * inv[l];
                    dd[l - 1][i] = prev[i];
                    prev[i] = c;
                }
            }
        }
        // level Lv receives the last value; k is already the next step's order: levels above it are zero (poly's invariant)
#pragma unroll
        for (int l = 1; l <= 6; l++) {
#pragma unroll
            for (int i = 0; i < B2_N; i++) dd[l][i] = (l > k) ? (real)0 : ((l == Lv) ? prev[i] : dd[l][i]);
        }
    }
    __device__ __forceinline__ void prepare_dense(const real (&)[B2_N], const real (&)[B2_NPA], real, real, int&) {}
    // cubic Hermite on (up, f0), (un, fnew): the default dense output of multistep methods
    __device__ __forceinline__ void interp(const real (&up)[B2_N], const real (&un)[B2_N], real th, real dt,
                                           real (&out)[B2_N]) const {
        const real om = (real)1 - th;
#pragma unroll
        for (int i = 0; i < B2_N; i++) {
            const real du = un[i] - up[i];
            real inner = ((real)1 - (real)2 * th) * du;
            inner = b2_fma((th - (real)1) * dt, f0[i], inner);
            inner = b2_fma(th * dt, fnew[i], inner);
            real v = om * up[i];
            v = b2_fma(th, un[i], v);
            out[i] = b2_fma(th * (th - (real)1), inner, v);
        }
    }
    __device__ __forceinline__ void advance() {
#pragma unroll
        for (int i = 0; i < B2_N; i++) f0[i] = fnew[i];
    }
};
