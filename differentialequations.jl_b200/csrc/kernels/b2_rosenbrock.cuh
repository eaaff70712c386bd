// b2_rosenbrock.cuh -- Rosenbrock23 and Rodas4/Rodas5/Rodas5P steppers with a per-thread,
// register-resident LU (partial pivoting, fully unrolled) and the analytic Jacobian b2_jac.
// Reference names: Rosenbrock23 /root/reference/test/qa/qa.jl:98, Rodas5P qa.jl:97 (used at
// src/DifferentialEquations.jl:28, test/core.jl:47); step forms SURVEY.md A.10, B.2, B.4-B.7.
// Replaces upstream's LinearSolve LU + ForwardDiff Jacobian.
#pragma once
#include "b2_common.cuh"
#include "tableaus_gen.cuh"

// Small systems (n <= 10): fully unrolled, A / dinv / piv are registers (spilling to constant-offset local memory beyond n ~ 6).  Larger systems: the O(n^3) unrolled code is megabytes of
// SASS (n = 16: a 4 MB cubin, minutes of JIT, instruction-cache bound) -- the loops stay rolled, the matrix lives in the
// thread's local memory (L1 / L2 resident) and factor / solve are out-of-line functions.  Same operations in the same order:
// bit-identical either way (B2_LU_ROLLED can be forced with B200ENS_DEFINES for the A/B test).
#ifndef B2_LU_ROLLED
#define B2_LU_ROLLED (B2_N > 10)   // measured crossover (tools/exp_stiff_mid.py): n = 10 unrolled 22.7 vs rolled 28.1 ms, n = 12: 59.5 vs 51.2 ms
#endif
#if B2_LU_ROLLED
#define B2_LU_UNROLL _Pragma("unroll 1")
#define B2_LU_INLINE __noinline__
#else
#define B2_LU_UNROLL _Pragma("unroll")
#define B2_LU_INLINE __forceinline__
#endif
// W (n x n) factored once per step and reused for every stage right-hand side.
struct B2LU {
    real A[B2_N][B2_N];
    real dinv[B2_N];
    int piv[B2_N];

#if !B2_LU_ROLLED
    __device__ B2_LU_INLINE void factor() {
B2_LU_UNROLL
        for (int k = 0; k < B2_N; k++) {
            int pr = k;
            real best = b2_abs(A[k][k]);
B2_LU_UNROLL
            for (int i = k + 1; i < B2_N; i++) {
                const real v = b2_abs(A[i][k]);
                if (v > best) {
                    best = v;
                    pr = i;
                }
            }
            piv[k] = pr;
B2_LU_UNROLL
            for (int i = k + 1; i < B2_N; i++) {
                const bool sw = pr == i;
B2_LU_UNROLL
                for (int j = 0; j < B2_N; j++) {
                    const real x = A[k][j], y = A[i][j];
                    A[k][j] = sw ? y : x;
                    A[i][j] = sw ? x : y;
                }
            }
            dinv[k] = (real)1 / A[k][k];
B2_LU_UNROLL
            for (int i = k + 1; i < B2_N; i++) {
                const real l = A[i][k] * dinv[k];
                A[i][k] = l;
B2_LU_UNROLL
                for (int j = k + 1; j < B2_N; j++) A[i][j] = b2_fma(-l, A[k][j], A[i][j]);
            }
        }
    }
    __device__ B2_LU_INLINE void solve(real (&b)[B2_N]) const {
B2_LU_UNROLL
        for (int k = 0; k < B2_N; k++) {
B2_LU_UNROLL
            for (int i = k + 1; i < B2_N; i++) {
                const bool sw = piv[k] == i;
                const real x = b[k], y = b[i];
                b[k] = sw ? y : x;
                b[i] = sw ? x : y;
            }
        }
B2_LU_UNROLL
        for (int i = 1; i < B2_N; i++) {
            real s = b[i];
B2_LU_UNROLL
            for (int j = 0; j < i; j++) s = b2_fma(-A[i][j], b[j], s);
            b[i] = s;
        }
B2_LU_UNROLL
        for (int i = B2_N - 1; i >= 0; i--) {
            real s = b[i];
B2_LU_UNROLL
            for (int j = i + 1; j < B2_N; j++) s = b2_fma(-A[i][j], b[j], s);
            b[i] = s * dinv[i];
        }
    }
#else
    // Rolled variant: the matrix is local memory, every access is a load / store with L1 / L2 latency.  The same
    // operations on the same values (bit-identical), organised for memory: rows are swapped only when the pivot moved
    // (the select form above touches 2 n^2 entries per column), and the inner loops fetch four entries of both rows
    // before the dependent multiply-adds so that the loads overlap.
    __device__ B2_LU_INLINE void factor() {
        for (int k = 0; k < B2_N; k++) {
            int pr = k;
            real best = b2_abs(A[k][k]);
            for (int i = k + 1; i < B2_N; i++) {
                const real v = b2_abs(A[i][k]);
                if (v > best) {
                    best = v;
                    pr = i;
                }
            }
            piv[k] = pr;
            if (pr != k) {
                for (int j = 0; j < B2_N; j++) {
                    const real x = A[k][j], y = A[pr][j];
                    A[k][j] = y;
                    A[pr][j] = x;
                }
            }
            const real di = (real)1 / A[k][k];
            dinv[k] = di;
            for (int i = k + 1; i < B2_N; i++) {
                const real l = A[i][k] * di;
                A[i][k] = l;
                const real nl = -l;
                int j = k + 1;
                for (; j + 3 < B2_N; j += 4) {
                    const real p0 = A[k][j], p1 = A[k][j + 1], p2 = A[k][j + 2], p3 = A[k][j + 3];
                    const real a0 = A[i][j], a1 = A[i][j + 1], a2 = A[i][j + 2], a3 = A[i][j + 3];
                    A[i][j] = b2_fma(nl, p0, a0);
                    A[i][j + 1] = b2_fma(nl, p1, a1);
                    A[i][j + 2] = b2_fma(nl, p2, a2);
                    A[i][j + 3] = b2_fma(nl, p3, a3);
                }
                for (; j < B2_N; j++) A[i][j] = b2_fma(nl, A[k][j], A[i][j]);
            }
        }
    }
    __device__ B2_LU_INLINE void solve(real (&b)[B2_N]) const {
        for (int k = 0; k < B2_N; k++) {
            const int pr = piv[k];
            if (pr != k) {
                const real x = b[k];
                b[k] = b[pr];
                b[pr] = x;
            }
        }
        for (int i = 1; i < B2_N; i++) {
            real s = b[i];
            int j = 0;
            for (; j + 3 < i; j += 4) {
                const real a0 = A[i][j], a1 = A[i][j + 1], a2 = A[i][j + 2], a3 = A[i][j + 3];
                const real y0 = b[j], y1 = b[j + 1], y2 = b[j + 2], y3 = b[j + 3];
                s = b2_fma(-a0, y0, s);
                s = b2_fma(-a1, y1, s);
                s = b2_fma(-a2, y2, s);
                s = b2_fma(-a3, y3, s);
            }
            for (; j < i; j++) s = b2_fma(-A[i][j], b[j], s);
            b[i] = s;
        }
        for (int i = B2_N - 1; i >= 0; i--) {
            real s = b[i];
            int j = i + 1;
            for (; j + 3 < B2_N; j += 4) {
                const real a0 = A[i][j], a1 = A[i][j + 1], a2 = A[i][j + 2], a3 = A[i][j + 3];
                const real y0 = b[j], y1 = b[j + 1], y2 = b[j + 2], y3 = b[j + 3];
                s = b2_fma(-a0, y0, s);
                s = b2_fma(-a1, y1, s);
                s = b2_fma(-a2, y2, s);
                s = b2_fma(-a3, y3, s);
            }
            for (; j < B2_N; j++) s = b2_fma(-A[i][j], b[j], s);
            b[i] = s * dinv[i];
        }
    }
#endif
};

#define B2_ROS23_D 0.29289321881345247560   /* 1/(2+sqrt 2) */
#define B2_ROS23_E32 7.41421356237309504880 /* 6+sqrt 2 */

struct B2Ros23 {
    static constexpr int ORDER = 2;
    __device__ __forceinline__ void bind(real*) {}
    static constexpr int DEG = 0;  // no coefficient form: the event search uses interp() directly
    __device__ __forceinline__ void poly_coeffs(int, real (&)[1]) const {}
    real f0[B2_N], k1[B2_N], k2[B2_N], f2[B2_N];

    __device__ __forceinline__ void start(const real (&u)[B2_N], const real (&p)[B2_NPA], real t) {
        b2_rhs(f0, u, p, t);
    }
    __device__ __forceinline__ real fsal0(int i) const { return f0[i]; }
    __device__ __forceinline__ void step(const real (&up)[B2_N], const real (&p)[B2_NPA], real t, real dt,
                                         real (&u)[B2_N], real (&ut)[B2_N], bool, int& nf) {
        B2LU lu;
        real J[B2_N * B2_N], tmp[B2_N], rhs[B2_N], f1[B2_N];
        const real dtd = dt * (real)B2_ROS23_D;
        b2_jac(J, up, p, t);
#pragma unroll
        for (int i = 0; i < B2_N; i++)
#pragma unroll
            for (int j = 0; j < B2_N; j++) lu.A[i][j] = (i == j ? (real)1 : (real)0) - dtd * J[i * B2_N + j];
        lu.factor();
#if B2_HAS_TGRAD
        real dT[B2_N];
        b2_tgrad(dT, up, p, t);
#pragma unroll
        for (int i = 0; i < B2_N; i++) rhs[i] = b2_fma(dtd, dT[i], f0[i]);
#else
#pragma unroll
        for (int i = 0; i < B2_N; i++) rhs[i] = f0[i];
#endif
        lu.solve(rhs);
#pragma unroll
        for (int i = 0; i < B2_N; i++) {
            k1[i] = rhs[i];
            tmp[i] = b2_fma((real)0.5 * dt, rhs[i], up[i]);
        }
        b2_rhs(f1, tmp, p, t + (real)0.5 * dt);
#pragma unroll
        for (int i = 0; i < B2_N; i++) rhs[i] = f1[i] - k1[i];
        lu.solve(rhs);
#pragma unroll
        for (int i = 0; i < B2_N; i++) {
            k2[i] = rhs[i] + k1[i];
            u[i] = b2_fma(dt, k2[i], up[i]);
        }
        b2_rhs(f2, u, p, t + dt);
        nf += 2;
#pragma unroll
        for (int i = 0; i < B2_N; i++) {
            real r = f2[i];
            r = b2_fma(-(real)B2_ROS23_E32, k2[i] - f1[i], r);
            r = b2_fma(-(real)2, k1[i] - f0[i], r);
#if B2_HAS_TGRAD
            r = b2_fma(dtd, dT[i], r);
#endif
            rhs[i] = r;
        }
        lu.solve(rhs);
#pragma unroll
        for (int i = 0; i < B2_N; i++) {
            const real e = b2_fma(-(real)2, k2[i], k1[i]) + rhs[i];
            ut[i] = (dt * (real)(1.0 / 6.0)) * e;
        }
    }
    __device__ __forceinline__ void accepted(const real (&)[B2_N], const real (&)[B2_NPA], real, int&) {}
    __device__ __forceinline__ void prepare_dense(const real (&)[B2_N], const real (&)[B2_NPA], real, real, int&) {}
    __device__ __forceinline__ void interp(const real (&up)[B2_N], const real (&)[B2_N], real th, real dt,
                                           real (&out)[B2_N]) const {
        const real den = (real)(1.0 / (1.0 - 2.0 * B2_ROS23_D));
        const real c1 = th * ((real)1 - th) * den;
        const real c2 = th * (th - (real)(2.0 * B2_ROS23_D)) * den;
#pragma unroll
        for (int i = 0; i < B2_N; i++) {
            real s = c1 * k1[i];
            s = b2_fma(c2, k2[i], s);
            out[i] = b2_fma(dt, s, up[i]);
        }
    }
    __device__ __forceinline__ void advance() {
#pragma unroll
        for (int i = 0; i < B2_N; i++) f0[i] = f2[i];
    }
};

// ---- Rodas family: tableau in __constant__ memory (indexed with compile-time offsets after
// unrolling, so every coefficient is a constant-bank operand of its FMA).
#if B2_ALG == 4
#define RT(x) ((real)(B2T_RODAS5_##x))
#define B2_RODAS_S 8
#define B2_RODAS_NEXP 6
#elif B2_ALG == 5
#define RT(x) ((real)(B2T_RODAS5P_##x))
#define B2_RODAS_S 8
#define B2_RODAS_NEXP 6
#elif B2_ALG == 8
#define RT(x) ((real)(B2T_RODAS4_##x))
#define B2_RODAS_S 6
#define B2_RODAS_NEXP 5
#endif

#ifdef B2_RODAS_S
#if B2_RODAS_S == 8
__constant__ real B2_RODAS_A[8][8] = {{0}, {RT(a21)}, {RT(a31), RT(a32)}, {RT(a41), RT(a42), RT(a43)},
                                      {RT(a51), RT(a52), RT(a53), RT(a54)},
                                      {RT(a61), RT(a62), RT(a63), RT(a64), RT(a65)}, {0}, {0}};
__constant__ real B2_RODAS_C[8][8] = {{0}, {RT(C21)}, {RT(C31), RT(C32)}, {RT(C41), RT(C42), RT(C43)},
                                      {RT(C51), RT(C52), RT(C53), RT(C54)},
                                      {RT(C61), RT(C62), RT(C63), RT(C64), RT(C65)},
                                      {RT(C71), RT(C72), RT(C73), RT(C74), RT(C75), RT(C76)},
                                      {RT(C81), RT(C82), RT(C83), RT(C84), RT(C85), RT(C86), RT(C87)}};
__constant__ real B2_RODAS_c[8] = {0, RT(c2), RT(c3), RT(c4), RT(c5), 1, 1, 1};
__constant__ real B2_RODAS_d[8] = {RT(d1), RT(d2), RT(d3), RT(d4), RT(d5), 0, 0, 0};
// dense-output weights in the transformed stage variables (order 4, derived: tools/derive_rodas_dense.py)
#define B2_RODAS_HDEG 4
__constant__ real B2_RODAS_H[8][4] = {{RT(H11), RT(H12), RT(H13), RT(H14)}, {RT(H21), RT(H22), RT(H23), RT(H24)},
                                      {RT(H31), RT(H32), RT(H33), RT(H34)}, {RT(H41), RT(H42), RT(H43), RT(H44)},
                                      {RT(H51), RT(H52), RT(H53), RT(H54)}, {RT(H61), RT(H62), RT(H63), RT(H64)},
                                      {RT(H71), RT(H72), RT(H73), RT(H74)}, {RT(H81), RT(H82), RT(H83), RT(H84)}};
#else
__constant__ real B2_RODAS_A[6][6] = {{0}, {RT(a21)}, {RT(a31), RT(a32)}, {RT(a41), RT(a42), RT(a43)},
                                      {RT(a51), RT(a52), RT(a53), RT(a54)}, {0}};
__constant__ real B2_RODAS_C[6][6] = {{0}, {RT(C21)}, {RT(C31), RT(C32)}, {RT(C41), RT(C42), RT(C43)},
                                      {RT(C51), RT(C52), RT(C53), RT(C54)},
                                      {RT(C61), RT(C62), RT(C63), RT(C64), RT(C65)}};
__constant__ real B2_RODAS_c[6] = {0, RT(c2), RT(c3), RT(c4), 1, 1};
__constant__ real B2_RODAS_d[6] = {RT(d1), RT(d2), RT(d3), RT(d4), 0, 0};
#define B2_RODAS_HDEG 3   // Rodas4: order-3 dense output
__constant__ real B2_RODAS_H[6][4] = {{RT(H11), RT(H12), RT(H13), 0}, {RT(H21), RT(H22), RT(H23), 0}, {RT(H31), RT(H32), RT(H33), 0},
                                      {RT(H41), RT(H42), RT(H43), 0}, {RT(H51), RT(H52), RT(H53), 0}, {RT(H61), RT(H62), RT(H63), 0}};
#endif

struct B2Rodas {
    static constexpr int ORDER = (B2_RODAS_S == 8) ? 5 : 4;
    __device__ __forceinline__ void bind(real*) {}
    static constexpr int DEG = 0;  // no coefficient form: the event search uses interp() directly
    __device__ __forceinline__ void poly_coeffs(int, real (&)[1]) const {}
    real f0[B2_N], fnew[B2_N];
    real k[B2_RODAS_S][B2_N];   // stage increments of the last step: the dense output is a weighted sum of them

    __device__ __forceinline__ void start(const real (&u)[B2_N], const real (&p)[B2_NPA], real t) {
        b2_rhs(f0, u, p, t);
    }
    __device__ __forceinline__ real fsal0(int i) const { return f0[i]; }
    __device__ __forceinline__ void step(const real (&up)[B2_N], const real (&p)[B2_NPA], real t, real dt,
                                         real (&u)[B2_N], real (&ut)[B2_N], bool, int& nf) {
        B2LU lu;
        real J[B2_N * B2_N], U[B2_N], rhs[B2_N], fU[B2_N];
        const real dtgi = (real)1 / (dt * RT(gamma));
        const real dtinv = (real)1 / dt;
        b2_jac(J, up, p, t);
#pragma unroll
        for (int i = 0; i < B2_N; i++)
#pragma unroll
#if B2_HAS_MASS
            // M u' = f (constant mass matrix, possibly singular: index-1 DAE): W = M/(gamma dt) - J
            for (int j = 0; j < B2_N; j++) lu.A[i][j] = (real)B2_MASS_[i * B2_N + j] * dtgi - J[i * B2_N + j];
#else
            for (int j = 0; j < B2_N; j++) lu.A[i][j] = (i == j ? dtgi : (real)0) - J[i * B2_N + j];
#endif
        lu.factor();
#if B2_HAS_TGRAD
        real dT[B2_N];
        b2_tgrad(dT, up, p, t);
#endif
#pragma unroll
        for (int st = 0; st < B2_RODAS_S; st++) {
            if (st == 0) {
#pragma unroll
                for (int i = 0; i < B2_N; i++) {
                    U[i] = up[i];
                    fU[i] = f0[i];
                }
            } else {
                if (st < B2_RODAS_NEXP) {
#pragma unroll
                    for (int i = 0; i < B2_N; i++) {
                        real v = up[i];
#pragma unroll
                        for (int j = 0; j < st; j++) v = b2_fma(B2_RODAS_A[st][j], k[j][i], v);
                        U[i] = v;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < B2_N; i++) U[i] = U[i] + k[st - 1][i];
                }
                b2_rhs(fU, U, p, t + B2_RODAS_c[st] * dt);
                nf += 1;
            }
#if B2_HAS_MASS
            real cacc[B2_N];   // sum_j C[st][j] k_j, multiplied by M below (non-zero entries of a row in index order)
            if (st > 0) {
#pragma unroll
                for (int i = 0; i < B2_N; i++) {
                    real acc = B2_RODAS_C[st][0] * k[0][i];
#pragma unroll
                    for (int j = 1; j < st; j++) acc = b2_fma(B2_RODAS_C[st][j], k[j][i], acc);
                    cacc[i] = acc;
                }
            }
#endif
#pragma unroll
            for (int i = 0; i < B2_N; i++) {
                real r = fU[i];
                if (st > 0) {
#if B2_HAS_MASS
                    real acc = 0;
                    bool first = true;
#pragma unroll
                    for (int j = 0; j < B2_N; j++) {
                        if (B2_MASS_[i * B2_N + j] != 0.0) {
                            acc = first ? (real)B2_MASS_[i * B2_N + j] * cacc[j] : b2_fma((real)B2_MASS_[i * B2_N + j], cacc[j], acc);
                            first = false;
                        }
                    }
#else
                    real acc = B2_RODAS_C[st][0] * k[0][i];
#pragma unroll
                    for (int j = 1; j < st; j++) acc = b2_fma(B2_RODAS_C[st][j], k[j][i], acc);
#endif
                    r = b2_fma(acc, dtinv, r);
                }
#if B2_HAS_TGRAD
                if (st < B2_RODAS_NEXP - 1) r = b2_fma(dt * B2_RODAS_d[st], dT[i], r);
#endif
                rhs[i] = r;
            }
            lu.solve(rhs);
#pragma unroll
            for (int i = 0; i < B2_N; i++) k[st][i] = rhs[i];
        }
#pragma unroll
        for (int i = 0; i < B2_N; i++) {
            u[i] = U[i] + k[B2_RODAS_S - 1][i];
            ut[i] = k[B2_RODAS_S - 1][i];
        }
    }
    // f(u_new): next step's stage-1 slope (and the Hermite end slope)
    __device__ __forceinline__ void accepted(const real (&u)[B2_N], const real (&p)[B2_NPA], real tnew, int& nf) {
        b2_rhs(fnew, u, p, tnew);
        nf += 1;
    }
    __device__ __forceinline__ void prepare_dense(const real (&)[B2_N], const real (&)[B2_NPA], real, real, int&) {}
    // Dense output in the transformed stage variables: out = up + sum_i w_i k_i, w_i = theta * Horner(theta; h_i1..h_iD).
    // Order 4 (Rodas5 / Rodas5P) / 3 (Rodas4), derived from the Rosenbrock order conditions (tools/derive_rodas_dense.py);
    // documented deviation: upstream's own coefficients were not recoverable.  On Robertson it is as accurate
    // as saving at tstops and an order of magnitude better than the cubic Hermite it replaced (profiles/README.md).
    __device__ __forceinline__ void interp(const real (&up)[B2_N], const real (&)[B2_N], real th, real,
                                           real (&out)[B2_N]) const {
        real w[B2_RODAS_S];
#pragma unroll
        for (int st = 0; st < B2_RODAS_S; st++) {
            real pv = B2_RODAS_H[st][B2_RODAS_HDEG - 1];
#pragma unroll
            for (int q = B2_RODAS_HDEG - 2; q >= 0; q--) pv = b2_fma(th, pv, B2_RODAS_H[st][q]);
            w[st] = th * pv;
        }
#pragma unroll
        for (int i = 0; i < B2_N; i++) {
            real v = up[i];
#pragma unroll
            for (int st = 0; st < B2_RODAS_S; st++) v = b2_fma(w[st], k[st][i], v);
            out[i] = v;
        }
    }
    __device__ __forceinline__ void advance() {
#pragma unroll
        for (int i = 0; i < B2_N; i++) f0[i] = fnew[i];
    }
};
#undef RT
#endif
