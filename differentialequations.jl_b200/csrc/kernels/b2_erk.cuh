// b2_erk.cuh -- explicit Runge-Kutta steppers: Tsit5 (5(4), FSAL, free 4th-order interpolant)
// and Vern7 (7(6), 10 stages + 6 lazy stages for an order-6 interpolant).
// Reference names: Tsit5 /root/reference/test/qa/qa.jl:119, Vern7 qa.jl:126 (bodies upstream,
// OrdinaryDiffEq 7); tables SURVEY.md B.1 / B.3, literals -> FFMA immediates (f32) or
// constant-bank operands (f64) after full unrolling.
#pragma once
#include "b2_common.cuh"
#include "tableaus_gen.cuh"

#define TS(x) ((real)(B2T_TSIT5_##x))

struct B2Tsit5 {
    static constexpr int ORDER = 5;
    real k1[B2_N], k2[B2_N], k3[B2_N], k4[B2_N], k5[B2_N], k6[B2_N], k7[B2_N];

    __device__ __forceinline__ void start(const real (&u)[B2_N], const real (&p)[B2_NPA], real t) {
        b2_rhs(k1, u, p, t);
    }
    // one step attempt from (up, t) with k1 = f(up, t); writes the proposal u and dt*error estimate
    __device__ __forceinline__ void step(const real (&up)[B2_N], const real (&p)[B2_NPA], real t, real dt,
                                         real (&u)[B2_N], real (&ut)[B2_N], bool adaptive, int& nf) {
        real tmp[B2_N];
#pragma unroll
        for (int i = 0; i < B2_N; i++) tmp[i] = b2_fma(dt, TS(a21) * k1[i], up[i]);
        b2_rhs(k2, tmp, p, t + TS(c2) * dt);
#pragma unroll
        for (int i = 0; i < B2_N; i++) {
            real s = TS(a31) * k1[i];
            s = b2_fma(TS(a32), k2[i], s);
            tmp[i] = b2_fma(dt, s, up[i]);
        }
        b2_rhs(k3, tmp, p, t + TS(c3) * dt);
#pragma unroll
        for (int i = 0; i < B2_N; i++) {
            real s = TS(a41) * k1[i];
            s = b2_fma(TS(a42), k2[i], s);
            s = b2_fma(TS(a43), k3[i], s);
            tmp[i] = b2_fma(dt, s, up[i]);
        }
        b2_rhs(k4, tmp, p, t + TS(c4) * dt);
#pragma unroll
        for (int i = 0; i < B2_N; i++) {
            real s = TS(a51) * k1[i];
            s = b2_fma(TS(a52), k2[i], s);
            s = b2_fma(TS(a53), k3[i], s);
            s = b2_fma(TS(a54), k4[i], s);
            tmp[i] = b2_fma(dt, s, up[i]);
        }
        b2_rhs(k5, tmp, p, t + TS(c5) * dt);
#pragma unroll
        for (int i = 0; i < B2_N; i++) {
            real s = TS(a61) * k1[i];
            s = b2_fma(TS(a62), k2[i], s);
            s = b2_fma(TS(a63), k3[i], s);
            s = b2_fma(TS(a64), k4[i], s);
            s = b2_fma(TS(a65), k5[i], s);
            tmp[i] = b2_fma(dt, s, up[i]);
        }
        b2_rhs(k6, tmp, p, t + dt);
#pragma unroll
        for (int i = 0; i < B2_N; i++) {
            real s = TS(a71) * k1[i];
            s = b2_fma(TS(a72), k2[i], s);
            s = b2_fma(TS(a73), k3[i], s);
            s = b2_fma(TS(a74), k4[i], s);
            s = b2_fma(TS(a75), k5[i], s);
            s = b2_fma(TS(a76), k6[i], s);
            u[i] = b2_fma(dt, s, up[i]);
        }
        b2_rhs(k7, u, p, t + dt);
        nf += 6;
        if (adaptive) {
#pragma unroll
            for (int i = 0; i < B2_N; i++) {
                real s = TS(btilde1) * k1[i];
                s = b2_fma(TS(btilde2), k2[i], s);
                s = b2_fma(TS(btilde3), k3[i], s);
                s = b2_fma(TS(btilde4), k4[i], s);
                s = b2_fma(TS(btilde5), k5[i], s);
                s = b2_fma(TS(btilde6), k6[i], s);
                s = b2_fma(TS(btilde7), k7[i], s);
                ut[i] = dt * s;
            }
        }
    }
    // called once per accepted step before interpolation / FSAL hand-over (no-op here)
    __device__ __forceinline__ void accepted(const real (&)[B2_N], const real (&)[B2_NPA], real, int&) {}
    __device__ __forceinline__ void prepare_dense(const real (&)[B2_N], const real (&)[B2_NPA], real, real, int&) {}
    // u(t + th*dt) = up + dt * sum_i b_i(th) k_i
    __device__ __forceinline__ void interp(const real (&up)[B2_N], const real (&)[B2_N], real th, real dt,
                                           real (&out)[B2_N]) const {
#define TSB(i, r1) (th * b2_fma(th, b2_fma(th, b2_fma(th, TS(r##i##4), TS(r##i##3)), TS(r##i##2)), (real)(r1)))
        const real b1 = TSB(1, B2T_TSIT5_r11), b2 = TSB(2, 0.0), b3 = TSB(3, 0.0), b4 = TSB(4, 0.0),
                   b5 = TSB(5, 0.0), b6 = TSB(6, 0.0), b7 = TSB(7, 0.0);
#undef TSB
#pragma unroll
        for (int i = 0; i < B2_N; i++) {
            real s = b1 * k1[i];
            s = b2_fma(b2, k2[i], s);
            s = b2_fma(b3, k3[i], s);
            s = b2_fma(b4, k4[i], s);
            s = b2_fma(b5, k5[i], s);
            s = b2_fma(b6, k6[i], s);
            s = b2_fma(b7, k7[i], s);
            out[i] = b2_fma(dt, s, up[i]);
        }
    }
    // FSAL: k7 = f(u_new) becomes the next step's k1
    __device__ __forceinline__ void advance() {
#pragma unroll
        for (int i = 0; i < B2_N; i++) k1[i] = k7[i];
    }
};
#undef TS

#define V7(x) ((real)(B2T_VERN7_##x))
#define V7X(x) ((real)(B2T_VERN7_EXTRA_##x))

struct B2Vern7 {
    static constexpr int ORDER = 7;
    // k2 and k3 only feed stages 3/4, so they share scratch; k10 is the error-only stage.
    real k1[B2_N], k4[B2_N], k5[B2_N], k6[B2_N], k7[B2_N], k8[B2_N], k9[B2_N], k11[B2_N];
    real k12[B2_N], k13[B2_N], k14[B2_N], k15[B2_N], k16[B2_N];
    bool have_extra;

    __device__ __forceinline__ void start(const real (&u)[B2_N], const real (&p)[B2_NPA], real t) {
        b2_rhs(k1, u, p, t);
        have_extra = false;
    }
    __device__ __forceinline__ void step(const real (&up)[B2_N], const real (&p)[B2_NPA], real t, real dt,
                                         real (&u)[B2_N], real (&ut)[B2_N], bool adaptive, int& nf) {
        real tmp[B2_N], k2[B2_N], k3[B2_N];
#pragma unroll
        for (int i = 0; i < B2_N; i++) tmp[i] = b2_fma(dt, V7(a0201) * k1[i], up[i]);
        b2_rhs(k2, tmp, p, t + V7(c2) * dt);
#pragma unroll
        for (int i = 0; i < B2_N; i++) {
            real s = V7(a0301) * k1[i];
            s = b2_fma(V7(a0302), k2[i], s);
            tmp[i] = b2_fma(dt, s, up[i]);
        }
        b2_rhs(k3, tmp, p, t + V7(c3) * dt);
#pragma unroll
        for (int i = 0; i < B2_N; i++) {
            real s = V7(a0401) * k1[i];
            s = b2_fma(V7(a0403), k3[i], s);
            tmp[i] = b2_fma(dt, s, up[i]);
        }
        b2_rhs(k4, tmp, p, t + V7(c4) * dt);
#pragma unroll
        for (int i = 0; i < B2_N; i++) {
            real s = V7(a0501) * k1[i];
            s = b2_fma(V7(a0503), k3[i], s);
            s = b2_fma(V7(a0504), k4[i], s);
            tmp[i] = b2_fma(dt, s, up[i]);
        }
        b2_rhs(k5, tmp, p, t + V7(c5) * dt);
#pragma unroll
        for (int i = 0; i < B2_N; i++) {
            real s = V7(a0601) * k1[i];
            s = b2_fma(V7(a0603), k3[i], s);
            s = b2_fma(V7(a0604), k4[i], s);
            s = b2_fma(V7(a0605), k5[i], s);
            tmp[i] = b2_fma(dt, s, up[i]);
        }
        b2_rhs(k6, tmp, p, t + V7(c6) * dt);
#pragma unroll
        for (int i = 0; i < B2_N; i++) {
            real s = V7(a0701) * k1[i];
            s = b2_fma(V7(a0703), k3[i], s);
            s = b2_fma(V7(a0704), k4[i], s);
            s = b2_fma(V7(a0705), k5[i], s);
            s = b2_fma(V7(a0706), k6[i], s);
            tmp[i] = b2_fma(dt, s, up[i]);
        }
        b2_rhs(k7, tmp, p, t + V7(c7) * dt);
#pragma unroll
        for (int i = 0; i < B2_N; i++) {
            real s = V7(a0801) * k1[i];
            s = b2_fma(V7(a0803), k3[i], s);
            s = b2_fma(V7(a0804), k4[i], s);
            s = b2_fma(V7(a0805), k5[i], s);
            s = b2_fma(V7(a0806), k6[i], s);
            s = b2_fma(V7(a0807), k7[i], s);
            tmp[i] = b2_fma(dt, s, up[i]);
        }
        b2_rhs(k8, tmp, p, t + V7(c8) * dt);
#pragma unroll
        for (int i = 0; i < B2_N; i++) {
            real s = V7(a0901) * k1[i];
            s = b2_fma(V7(a0903), k3[i], s);
            s = b2_fma(V7(a0904), k4[i], s);
            s = b2_fma(V7(a0905), k5[i], s);
            s = b2_fma(V7(a0906), k6[i], s);
            s = b2_fma(V7(a0907), k7[i], s);
            s = b2_fma(V7(a0908), k8[i], s);
            tmp[i] = b2_fma(dt, s, up[i]);
        }
        b2_rhs(k9, tmp, p, t + dt);
        real k10[B2_N];
        if (adaptive) {
#pragma unroll
            for (int i = 0; i < B2_N; i++) {
                real s = V7(a1001) * k1[i];
                s = b2_fma(V7(a1003), k3[i], s);
                s = b2_fma(V7(a1004), k4[i], s);
                s = b2_fma(V7(a1005), k5[i], s);
                s = b2_fma(V7(a1006), k6[i], s);
                s = b2_fma(V7(a1007), k7[i], s);
                tmp[i] = b2_fma(dt, s, up[i]);
            }
            b2_rhs(k10, tmp, p, t + dt);
        }
#pragma unroll
        for (int i = 0; i < B2_N; i++) {
            real s = V7(b1) * k1[i];
            s = b2_fma(V7(b4), k4[i], s);
            s = b2_fma(V7(b5), k5[i], s);
            s = b2_fma(V7(b6), k6[i], s);
            s = b2_fma(V7(b7), k7[i], s);
            s = b2_fma(V7(b8), k8[i], s);
            s = b2_fma(V7(b9), k9[i], s);
            u[i] = b2_fma(dt, s, up[i]);
        }
        if (adaptive) {
#pragma unroll
            for (int i = 0; i < B2_N; i++) {
                real s = V7(btilde1) * k1[i];
                s = b2_fma(V7(btilde4), k4[i], s);
                s = b2_fma(V7(btilde5), k5[i], s);
                s = b2_fma(V7(btilde6), k6[i], s);
                s = b2_fma(V7(btilde7), k7[i], s);
                s = b2_fma(V7(btilde8), k8[i], s);
                s = b2_fma(V7(btilde9), k9[i], s);
                s = b2_fma(V7(btilde10), k10[i], s);
                ut[i] = dt * s;
            }
        }
        nf += adaptive ? 9 : 8;
        have_extra = false;
    }
    // k11 = f(u_new): dense-output stage 11 and the next step's k1 (Vern7 is not FSAL in the step itself)
    __device__ __forceinline__ void accepted(const real (&u)[B2_N], const real (&p)[B2_NPA], real tnew, int& nf) {
        b2_rhs(k11, u, p, tnew);
        nf += 1;
    }
    // lazy stages 12..16, only on steps that are interpolated (saveat / event search)
    __device__ __forceinline__ void prepare_dense(const real (&up)[B2_N], const real (&p)[B2_NPA], real t, real dt,
                                                  int& nf) {
        if (have_extra) return;
        real tmp[B2_N];
#define V7ROW(r, KLAST)                                                    \
    _Pragma("unroll") for (int i = 0; i < B2_N; i++) {                     \
        real s = V7X(a##r##01) * k1[i];                                    \
        s = b2_fma(V7X(a##r##04), k4[i], s);                               \
        s = b2_fma(V7X(a##r##05), k5[i], s);                               \
        s = b2_fma(V7X(a##r##06), k6[i], s);                               \
        s = b2_fma(V7X(a##r##07), k7[i], s);                               \
        s = b2_fma(V7X(a##r##08), k8[i], s);                               \
        s = b2_fma(V7X(a##r##09), k9[i], s);                               \
        s = b2_fma(V7X(a##r##11), k11[i], s);                              \
        KLAST tmp[i] = b2_fma(dt, s, up[i]);                               \
    }
        V7ROW(12, )
        b2_rhs(k12, tmp, p, t + V7X(c12) * dt);
        V7ROW(13, s = b2_fma(V7X(a1312), k12[i], s);)
        b2_rhs(k13, tmp, p, t + V7X(c13) * dt);
        V7ROW(14, s = b2_fma(V7X(a1412), k12[i], s); s = b2_fma(V7X(a1413), k13[i], s);)
        b2_rhs(k14, tmp, p, t + V7X(c14) * dt);
        V7ROW(15, s = b2_fma(V7X(a1512), k12[i], s); s = b2_fma(V7X(a1513), k13[i], s);)
        b2_rhs(k15, tmp, p, t + V7X(c15) * dt);
        V7ROW(16, s = b2_fma(V7X(a1612), k12[i], s); s = b2_fma(V7X(a1613), k13[i], s);)
        b2_rhs(k16, tmp, p, t + V7X(c16) * dt);
#undef V7ROW
        nf += 5;
        have_extra = true;
    }
    __device__ __forceinline__ void interp(const real (&up)[B2_N], const real (&)[B2_N], real th, real dt,
                                           real (&out)[B2_N]) const {
#define V7B(ss)                                                                                                   \
    (th * b2_fma(th, b2_fma(th, b2_fma(th, b2_fma(th, b2_fma(th, V7(R##ss##_6), V7(R##ss##_5)), V7(R##ss##_4)), \
                                       V7(R##ss##_3)), V7(R##ss##_2)), V7(R##ss##_1)))
        const real b01 = V7B(01), b04 = V7B(04), b05 = V7B(05), b06 = V7B(06), b07 = V7B(07), b08 = V7B(08),
                   b09 = V7B(09), b11 = V7B(11), b12 = V7B(12), b13 = V7B(13), b14 = V7B(14), b15 = V7B(15),
                   b16 = V7B(16);
#undef V7B
#pragma unroll
        for (int i = 0; i < B2_N; i++) {
            real s = b01 * k1[i];
            s = b2_fma(b04, k4[i], s);
            s = b2_fma(b05, k5[i], s);
            s = b2_fma(b06, k6[i], s);
            s = b2_fma(b07, k7[i], s);
            s = b2_fma(b08, k8[i], s);
            s = b2_fma(b09, k9[i], s);
            s = b2_fma(b11, k11[i], s);
            s = b2_fma(b12, k12[i], s);
            s = b2_fma(b13, k13[i], s);
            s = b2_fma(b14, k14[i], s);
            s = b2_fma(b15, k15[i], s);
            s = b2_fma(b16, k16[i], s);
            out[i] = b2_fma(dt, s, up[i]);
        }
    }
    __device__ __forceinline__ void advance() {
#pragma unroll
        for (int i = 0; i < B2_N; i++) k1[i] = k11[i];
    }
};
#undef V7
#undef V7X
