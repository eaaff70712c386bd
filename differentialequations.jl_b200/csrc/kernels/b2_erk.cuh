// b2_erk.cuh -- explicit Runge-Kutta steppers: Tsit5 (5(4), FSAL, free 4th-order interpolant)
// and Vern7 (7(6), 10 stages + 6 lazy stages for an order-6 interpolant).
// Reference names: Tsit5 /root/reference/test/qa/qa.jl:119, Vern7 qa.jl:126 (bodies upstream,
// OrdinaryDiffEq 7); tables SURVEY.md B.1 / B.3, literals -> FFMA immediates (f32) or
// constant-bank operands (f64) after full unrolling.
#pragma once
#include "b2_common.cuh"
#include "tableaus_gen.cuh"

// ---- stage-vector storage.  Small systems keep every k-vector in registers.  When n_state x stages
// no longer fits the register file (ptxas spills: config 5, n=16 f64 Vern7 spilled 4 KB/thread and became
// DRAM-bound on local-memory traffic, profiles/r1_net16_*), the model is recompiled with B2_KSMEM=1 and the
// k-vectors live in shared memory as [vector][component][thread] (bank-conflict-free, no spills).
// B2_NV: components of the state a THREAD holds.  One trajectory per thread: all of them.  Split mode (B2_SPLIT,
// b2_ode_driver_split.cuh): the four warps of a CTA share 32 trajectories, warp g owns components [g*NL, (g+1)*NL),
// and every RHS evaluation goes through b2_split_rhs (publish the stage argument in shared memory, barrier, evaluate
// only the owned outputs).
#if B2_SPLIT
#define B2_NV B2_NL
#define B2_RHS_REG(out, x, t_) b2_split_rhs(*this, out, x, p, t_)
#else
#define B2_NV B2_N
#define B2_RHS_REG(out, x, t_) b2_rhs(out, x, p, t_)
#endif

#if B2_KSMEM
#define B2_KSTRIDE B2_BLOCK
#define B2_KDECL(name) real* name
#define B2_KBIND(name, slot) name = kbase + (size_t)(slot) * B2_N * B2_BLOCK
// RHS into a shared-memory vector: evaluate into registers, then store
#define B2_RHS_TO(k, x, t_)                                  \
    do {                                                     \
        real f_[B2_N];                                       \
        b2_rhs(f_, x, p, t_);                                \
        _Pragma("unroll") for (int i_ = 0; i_ < B2_N; i_++) KV(k, i_) = f_[i_]; \
    } while (0)
#else
#define B2_KSTRIDE 1
#define B2_KDECL(name) real name[B2_NV]
#define B2_KBIND(name, slot)
#define B2_RHS_TO(k, x, t_) B2_RHS_REG(k, x, t_)
#endif
#define KV(name, i) name[(i) * B2_KSTRIDE]
// In shared-memory mode ptxas otherwise hoists the loads of all components of a stage to the top of the unrolled
// loop (hundreds of live registers -> spills); a compiler-level memory barrier per component keeps live ranges short.
#if B2_KSMEM
#define B2_KBAR asm volatile("" ::: "memory")
#else
#define B2_KBAR
#endif

// ---- chunked loops over the components a thread holds: a chunk is a component PAIR (FFMA2/FMUL2, B2_PACK2) with a
// scalar tail for odd n, or one scalar everywhere else.  V = chunk type, i = first component of the chunk.
#if B2_PACK2
#define B2_CHUNKS(BODY)                                                                       \
    _Pragma("unroll") for (int i = 0; i + 1 < B2_NV; i += 2) { typedef b2p V; BODY }           \
    if (B2_NV & 1) { constexpr int i = B2_NV - 1; typedef float V; BODY }
#define KL(name) b2_ld<V>(name, i)
#define UL(name) b2_ld<V>(name, i)
#define CB(x) b2_bc<V>(x)
#define VST(name, val) b2_st(name, i, val)
#else
#define B2_CHUNKS(BODY) _Pragma("unroll") for (int i = 0; i < B2_NV; i++) { typedef real V; BODY B2_KBAR; }
#define KL(name) KV(name, i)
#define UL(name) name[i]
#define CB(x) (x)
#define VST(name, val) name[i] = (val)
#endif

// Coefficients are literals: 32-bit FFMA / FFMA2 immediates in Float32; in Float64 ptxas materialises each with two
// moves (UMOV / IMAD.MOV).  Declaring them `__constant__ double` instead (constant-bank loads, LDCU.128 per pair) was
// measured in round 2: -5 % static instructions, but no faster (Tsit5 f64 1.443 -> 1.447 ms, Vern7 split 32.0 -> 33.3 ms
// per 200k: more spills, constant-cache latency on the critical path), so the literals stay.
#define TS(x) ((real)(B2T_TSIT5_##x))

struct B2Tsit5 {
    static constexpr int ORDER = 5;
    static constexpr int NVEC = 7;   // k-vectors held per thread (shared-memory slots when B2_KSMEM)
    B2_KDECL(k1); B2_KDECL(k2); B2_KDECL(k3); B2_KDECL(k4); B2_KDECL(k5); B2_KDECL(k6); B2_KDECL(k7);
    __device__ __forceinline__ void bind(real* kbase) {
        (void)kbase;
        B2_KBIND(k1, 0); B2_KBIND(k2, 1); B2_KBIND(k3, 2); B2_KBIND(k4, 3); B2_KBIND(k5, 4); B2_KBIND(k6, 5); B2_KBIND(k7, 6);
    }

    __device__ __forceinline__ void start(const real (&u)[B2_NV], const real (&p)[B2_NPA], real t) {
        B2_RHS_TO(k1, u, t);
    }
    __device__ __forceinline__ real fsal0(int i) const { return KV(k1, i); }  // f(u, t) of the current state
#if B2_SPLIT
    B2Xchg xc;   // shared-memory exchange context of the 4-warp group (b2_split.cuh)
    __device__ __forceinline__ void reset_dense() {}
    __device__ __forceinline__ void rhs(real (&f)[B2_NV], const real (&x)[B2_NV], const real (&p)[B2_NPA], real t) { B2_RHS_REG(f, x, t); }
    __device__ __forceinline__ void set_k1(const real (&f)[B2_NV]) {
#pragma unroll
        for (int i = 0; i < B2_NV; i++) k1[i] = f[i];
    }
#endif
    // one step attempt from (up, t) with k1 = f(up, t); writes the proposal u and dt*error estimate
    __device__ __forceinline__ void step(const real (&up)[B2_NV], const real (&p)[B2_NPA], real t, real dt,
                                         real (&u)[B2_NV], real (&ut)[B2_NV], bool adaptive, int& nf) {
        real tmp[B2_NV];
        B2_CHUNKS(VST(tmp, b2_fma(CB(dt), CB(TS(a21)) * KL(k1), UL(up)));)
        B2_RHS_TO(k2, tmp, t + TS(c2) * dt);
        B2_CHUNKS(V s = CB(TS(a31)) * KL(k1);
                  s = b2_fma(CB(TS(a32)), KL(k2), s);
                  VST(tmp, b2_fma(CB(dt), s, UL(up)));)
        B2_RHS_TO(k3, tmp, t + TS(c3) * dt);
        B2_CHUNKS(V s = CB(TS(a41)) * KL(k1);
                  s = b2_fma(CB(TS(a42)), KL(k2), s);
                  s = b2_fma(CB(TS(a43)), KL(k3), s);
                  VST(tmp, b2_fma(CB(dt), s, UL(up)));)
        B2_RHS_TO(k4, tmp, t + TS(c4) * dt);
        B2_CHUNKS(V s = CB(TS(a51)) * KL(k1);
                  s = b2_fma(CB(TS(a52)), KL(k2), s);
                  s = b2_fma(CB(TS(a53)), KL(k3), s);
                  s = b2_fma(CB(TS(a54)), KL(k4), s);
                  VST(tmp, b2_fma(CB(dt), s, UL(up)));)
        B2_RHS_TO(k5, tmp, t + TS(c5) * dt);
        B2_CHUNKS(V s = CB(TS(a61)) * KL(k1);
                  s = b2_fma(CB(TS(a62)), KL(k2), s);
                  s = b2_fma(CB(TS(a63)), KL(k3), s);
                  s = b2_fma(CB(TS(a64)), KL(k4), s);
                  s = b2_fma(CB(TS(a65)), KL(k5), s);
                  VST(tmp, b2_fma(CB(dt), s, UL(up)));)
        B2_RHS_TO(k6, tmp, t + dt);
        B2_CHUNKS(V s = CB(TS(a71)) * KL(k1);
                  s = b2_fma(CB(TS(a72)), KL(k2), s);
                  s = b2_fma(CB(TS(a73)), KL(k3), s);
                  s = b2_fma(CB(TS(a74)), KL(k4), s);
                  s = b2_fma(CB(TS(a75)), KL(k5), s);
                  s = b2_fma(CB(TS(a76)), KL(k6), s);
                  VST(u, b2_fma(CB(dt), s, UL(up)));)
        B2_RHS_TO(k7, u, t + dt);
        nf += 6;
        if (adaptive) {
            B2_CHUNKS(V s = CB(TS(btilde1)) * KL(k1);
                      s = b2_fma(CB(TS(btilde2)), KL(k2), s);
                      s = b2_fma(CB(TS(btilde3)), KL(k3), s);
                      s = b2_fma(CB(TS(btilde4)), KL(k4), s);
                      s = b2_fma(CB(TS(btilde5)), KL(k5), s);
                      s = b2_fma(CB(TS(btilde6)), KL(k6), s);
                      s = b2_fma(CB(TS(btilde7)), KL(k7), s);
                      VST(ut, CB(dt) * s);)
        }
    }
    // called once per accepted step before interpolation / FSAL hand-over (no-op here)
    __device__ __forceinline__ void accepted(const real (&)[B2_NV], const real (&)[B2_NPA], real, int&) {}
    __device__ __forceinline__ void prepare_dense(const real (&)[B2_NV], const real (&)[B2_NPA], real, real, int&) {}
    // u(t + th*dt) = up + dt * sum_i b_i(th) k_i with b_1 = th*q_1(th), b_i = th^2*q_i(th) (i >= 2: r_i1 = 0), evaluated as
    //   up + (dt*th) * ( q_1*k_1 + th * sum_{i>=2} q_i*k_i ),   q_i = r_i2 + th*(r_i3 + th*r_i4)
    // -- the common factors th, th^2 are applied once per component instead of once per stage (15 + 9 per chunk issue
    // slots instead of 28 + 8; the saveat block runs at ~5/32 lanes).  Same tree as the oracle's tsit5_interp.
    __device__ __forceinline__ void interp(const real (&up)[B2_NV], const real (&)[B2_NV], real th, real dt,
                                           real (&out)[B2_NV]) const {
#define TSQ(i) b2_fma(th, b2_fma(th, TS(r##i##4), TS(r##i##3)), TS(r##i##2))
        const real q1 = b2_fma(th, TSQ(1), TS(r11));
        const real q2 = TSQ(2), q3 = TSQ(3), q4 = TSQ(4), q5 = TSQ(5), q6 = TSQ(6), q7 = TSQ(7);
#undef TSQ
        const real dth = dt * th;
        B2_CHUNKS(V s = CB(q2) * KL(k2);
                  s = b2_fma(CB(q3), KL(k3), s);
                  s = b2_fma(CB(q4), KL(k4), s);
                  s = b2_fma(CB(q5), KL(k5), s);
                  s = b2_fma(CB(q6), KL(k6), s);
                  s = b2_fma(CB(q7), KL(k7), s);
                  s = b2_fma(CB(th), s, CB(q1) * KL(k1));
                  VST(out, b2_fma(CB(dth), s, UL(up)));)
    }
    // Coefficient form of the interpolant for ONE component (used by the event search, which evaluates the dense
    // output many times per step): u_i(t + th*dt) = up_i + dt * th*(C1 + th*(C2 + th*(C3 + th*C4))), C_j = sum_s r_sj k_s[i]
    static constexpr int DEG = 4;
    __device__ __forceinline__ void poly_coeffs(int i, real (&c)[DEG]) const {
        c[0] = TS(r11) * KV(k1, i);
#define TSC(j)                                          \
    {                                                   \
        real s = TS(r1##j) * KV(k1, i);                 \
        s = b2_fma(TS(r2##j), KV(k2, i), s);            \
        s = b2_fma(TS(r3##j), KV(k3, i), s);            \
        s = b2_fma(TS(r4##j), KV(k4, i), s);            \
        s = b2_fma(TS(r5##j), KV(k5, i), s);            \
        s = b2_fma(TS(r6##j), KV(k6, i), s);            \
        s = b2_fma(TS(r7##j), KV(k7, i), s);            \
        c[j - 1] = s;                                   \
    }
        TSC(2) TSC(3) TSC(4)
#undef TSC
    }
    // FSAL: k7 = f(u_new) becomes the next step's k1
    __device__ __forceinline__ void advance() {
#if B2_KSMEM
        real* t_ = k1;
        k1 = k7;
        k7 = t_;
#else
#pragma unroll
        for (int i = 0; i < B2_NV; i++) k1[i] = k7[i];
#endif
    }
};
#undef TS

#define V7(x) ((real)(B2T_VERN7_##x))
#define V7X(x) ((real)(B2T_VERN7_EXTRA_##x))

struct B2Vern7 {
    static constexpr int ORDER = 7;
    // k2 (feeds stage 3 only) and k10 (error-only stage) are step-local register scratch.
    static constexpr int NVEC = 14;
    B2_KDECL(k1); B2_KDECL(k3); B2_KDECL(k4); B2_KDECL(k5); B2_KDECL(k6); B2_KDECL(k7); B2_KDECL(k8); B2_KDECL(k9);
    B2_KDECL(k11); B2_KDECL(k12); B2_KDECL(k13); B2_KDECL(k14); B2_KDECL(k15); B2_KDECL(k16);
    __device__ __forceinline__ void bind(real* kbase) {
        (void)kbase;
        B2_KBIND(k1, 0); B2_KBIND(k3, 1); B2_KBIND(k4, 2); B2_KBIND(k5, 3); B2_KBIND(k6, 4); B2_KBIND(k7, 5); B2_KBIND(k8, 6);
        B2_KBIND(k9, 7); B2_KBIND(k11, 8);
        B2_KBIND(k12, 9); B2_KBIND(k13, 10); B2_KBIND(k14, 11); B2_KBIND(k15, 12); B2_KBIND(k16, 13);
    }
    bool have_extra;

    __device__ __forceinline__ void start(const real (&u)[B2_NV], const real (&p)[B2_NPA], real t) {
        B2_RHS_TO(k1, u, t);
        have_extra = false;
    }
    __device__ __forceinline__ real fsal0(int i) const { return KV(k1, i); }
#if B2_SPLIT
    B2Xchg xc;
    __device__ __forceinline__ void reset_dense() { have_extra = false; }
    __device__ __forceinline__ void rhs(real (&f)[B2_NV], const real (&x)[B2_NV], const real (&p)[B2_NPA], real t) { B2_RHS_REG(f, x, t); }
    __device__ __forceinline__ void set_k1(const real (&f)[B2_NV]) {
#pragma unroll
        for (int i = 0; i < B2_NV; i++) k1[i] = f[i];
    }
#endif
    __device__ __forceinline__ void step(const real (&up)[B2_NV], const real (&p)[B2_NPA], real t, real dt,
                                         real (&u)[B2_NV], real (&ut)[B2_NV], bool adaptive, int& nf) {
        real tmp[B2_NV], q2[B2_NV];
#pragma unroll
        for (int i = 0; i < B2_NV; i++) tmp[i] = b2_fma(dt, V7(a0201) * KV(k1, i), up[i]);
        B2_RHS_REG(q2, tmp, t + V7(c2) * dt);
#pragma unroll
        for (int i = 0; i < B2_NV; i++) {
            real s = V7(a0301) * KV(k1, i);
            s = b2_fma(V7(a0302), q2[i], s);
            tmp[i] = b2_fma(dt, s, up[i]);
            B2_KBAR;
        }
        B2_RHS_TO(k3, tmp, t + V7(c3) * dt);
#pragma unroll
        for (int i = 0; i < B2_NV; i++) {
            real s = V7(a0401) * KV(k1, i);
            s = b2_fma(V7(a0403), KV(k3, i), s);
            tmp[i] = b2_fma(dt, s, up[i]);
            B2_KBAR;
        }
        B2_RHS_TO(k4, tmp, t + V7(c4) * dt);
#pragma unroll
        for (int i = 0; i < B2_NV; i++) {
            real s = V7(a0501) * KV(k1, i);
            s = b2_fma(V7(a0503), KV(k3, i), s);
            s = b2_fma(V7(a0504), KV(k4, i), s);
            tmp[i] = b2_fma(dt, s, up[i]);
            B2_KBAR;
        }
        B2_RHS_TO(k5, tmp, t + V7(c5) * dt);
#pragma unroll
        for (int i = 0; i < B2_NV; i++) {
            real s = V7(a0601) * KV(k1, i);
            s = b2_fma(V7(a0603), KV(k3, i), s);
            s = b2_fma(V7(a0604), KV(k4, i), s);
            s = b2_fma(V7(a0605), KV(k5, i), s);
            tmp[i] = b2_fma(dt, s, up[i]);
            B2_KBAR;
        }
        B2_RHS_TO(k6, tmp, t + V7(c6) * dt);
#pragma unroll
        for (int i = 0; i < B2_NV; i++) {
            real s = V7(a0701) * KV(k1, i);
            s = b2_fma(V7(a0703), KV(k3, i), s);
            s = b2_fma(V7(a0704), KV(k4, i), s);
            s = b2_fma(V7(a0705), KV(k5, i), s);
            s = b2_fma(V7(a0706), KV(k6, i), s);
            tmp[i] = b2_fma(dt, s, up[i]);
            B2_KBAR;
        }
        B2_RHS_TO(k7, tmp, t + V7(c7) * dt);
#pragma unroll
        for (int i = 0; i < B2_NV; i++) {
            real s = V7(a0801) * KV(k1, i);
            s = b2_fma(V7(a0803), KV(k3, i), s);
            s = b2_fma(V7(a0804), KV(k4, i), s);
            s = b2_fma(V7(a0805), KV(k5, i), s);
            s = b2_fma(V7(a0806), KV(k6, i), s);
            s = b2_fma(V7(a0807), KV(k7, i), s);
            tmp[i] = b2_fma(dt, s, up[i]);
            B2_KBAR;
        }
        B2_RHS_TO(k8, tmp, t + V7(c8) * dt);
#pragma unroll
        for (int i = 0; i < B2_NV; i++) {
            real s = V7(a0901) * KV(k1, i);
            s = b2_fma(V7(a0903), KV(k3, i), s);
            s = b2_fma(V7(a0904), KV(k4, i), s);
            s = b2_fma(V7(a0905), KV(k5, i), s);
            s = b2_fma(V7(a0906), KV(k6, i), s);
            s = b2_fma(V7(a0907), KV(k7, i), s);
            s = b2_fma(V7(a0908), KV(k8, i), s);
            tmp[i] = b2_fma(dt, s, up[i]);
            B2_KBAR;
        }
        B2_RHS_TO(k9, tmp, t + dt);
        real q10[B2_NV];
        if (adaptive) {
#pragma unroll
            for (int i = 0; i < B2_NV; i++) {
                real s = V7(a1001) * KV(k1, i);
                s = b2_fma(V7(a1003), KV(k3, i), s);
                s = b2_fma(V7(a1004), KV(k4, i), s);
                s = b2_fma(V7(a1005), KV(k5, i), s);
                s = b2_fma(V7(a1006), KV(k6, i), s);
                s = b2_fma(V7(a1007), KV(k7, i), s);
                tmp[i] = b2_fma(dt, s, up[i]);
                B2_KBAR;
            }
            B2_RHS_REG(q10, tmp, t + dt);
        }
#pragma unroll
        for (int i = 0; i < B2_NV; i++) {
            real s = V7(b1) * KV(k1, i);
            s = b2_fma(V7(b4), KV(k4, i), s);
            s = b2_fma(V7(b5), KV(k5, i), s);
            s = b2_fma(V7(b6), KV(k6, i), s);
            s = b2_fma(V7(b7), KV(k7, i), s);
            s = b2_fma(V7(b8), KV(k8, i), s);
            s = b2_fma(V7(b9), KV(k9, i), s);
            u[i] = b2_fma(dt, s, up[i]);
            B2_KBAR;
        }
        if (adaptive) {
#pragma unroll
            for (int i = 0; i < B2_NV; i++) {
                real s = V7(btilde1) * KV(k1, i);
                s = b2_fma(V7(btilde4), KV(k4, i), s);
                s = b2_fma(V7(btilde5), KV(k5, i), s);
                s = b2_fma(V7(btilde6), KV(k6, i), s);
                s = b2_fma(V7(btilde7), KV(k7, i), s);
                s = b2_fma(V7(btilde8), KV(k8, i), s);
                s = b2_fma(V7(btilde9), KV(k9, i), s);
                s = b2_fma(V7(btilde10), q10[i], s);
                ut[i] = dt * s;
                B2_KBAR;
            }
        }
        nf += adaptive ? 9 : 8;
        have_extra = false;
    }
    // k11 = f(u_new): dense-output stage 11 and the next step's k1 (Vern7 is not FSAL in the step itself)
    __device__ __forceinline__ void accepted(const real (&u)[B2_NV], const real (&p)[B2_NPA], real tnew, int& nf) {
        B2_RHS_TO(k11, u, tnew);
        nf += 1;
    }
    // lazy stages 12..16, only on steps that are interpolated (saveat / event search)
    __device__ __forceinline__ void prepare_dense(const real (&up)[B2_NV], const real (&p)[B2_NPA], real t, real dt,
                                                  int& nf) {
        if (have_extra) return;
        real tmp[B2_NV];
#define V7ROW(r, KLAST)                                                    \
    _Pragma("unroll") for (int i = 0; i < B2_NV; i++) {                     \
        real s = V7X(a##r##01) * KV(k1, i);                                    \
        s = b2_fma(V7X(a##r##04), KV(k4, i), s);                               \
        s = b2_fma(V7X(a##r##05), KV(k5, i), s);                               \
        s = b2_fma(V7X(a##r##06), KV(k6, i), s);                               \
        s = b2_fma(V7X(a##r##07), KV(k7, i), s);                               \
        s = b2_fma(V7X(a##r##08), KV(k8, i), s);                               \
        s = b2_fma(V7X(a##r##09), KV(k9, i), s);                               \
        s = b2_fma(V7X(a##r##11), KV(k11, i), s);                              \
        KLAST tmp[i] = b2_fma(dt, s, up[i]);                               \
        B2_KBAR;                                                           \
    }
        V7ROW(12, )
        B2_RHS_TO(k12, tmp, t + V7X(c12) * dt);
        V7ROW(13, s = b2_fma(V7X(a1312), KV(k12, i), s);)
        B2_RHS_TO(k13, tmp, t + V7X(c13) * dt);
        V7ROW(14, s = b2_fma(V7X(a1412), KV(k12, i), s); s = b2_fma(V7X(a1413), KV(k13, i), s);)
        B2_RHS_TO(k14, tmp, t + V7X(c14) * dt);
        V7ROW(15, s = b2_fma(V7X(a1512), KV(k12, i), s); s = b2_fma(V7X(a1513), KV(k13, i), s);)
        B2_RHS_TO(k15, tmp, t + V7X(c15) * dt);
        V7ROW(16, s = b2_fma(V7X(a1612), KV(k12, i), s); s = b2_fma(V7X(a1613), KV(k13, i), s);)
        B2_RHS_TO(k16, tmp, t + V7X(c16) * dt);
#undef V7ROW
        nf += 5;
        have_extra = true;
    }
    __device__ __forceinline__ void interp(const real (&up)[B2_NV], const real (&)[B2_NV], real th, real dt,
                                           real (&out)[B2_NV]) const {
#define V7B(ss)                                                                                                   \
    (th * b2_fma(th, b2_fma(th, b2_fma(th, b2_fma(th, b2_fma(th, V7(R##ss##_6), V7(R##ss##_5)), V7(R##ss##_4)), \
                                       V7(R##ss##_3)), V7(R##ss##_2)), V7(R##ss##_1)))
        const real b01 = V7B(01), b04 = V7B(04), b05 = V7B(05), b06 = V7B(06), b07 = V7B(07), b08 = V7B(08),
                   b09 = V7B(09), b11 = V7B(11), b12 = V7B(12), b13 = V7B(13), b14 = V7B(14), b15 = V7B(15),
                   b16 = V7B(16);
#undef V7B
#pragma unroll
        for (int i = 0; i < B2_NV; i++) {
            real s = b01 * KV(k1, i);
            s = b2_fma(b04, KV(k4, i), s);
            s = b2_fma(b05, KV(k5, i), s);
            s = b2_fma(b06, KV(k6, i), s);
            s = b2_fma(b07, KV(k7, i), s);
            s = b2_fma(b08, KV(k8, i), s);
            s = b2_fma(b09, KV(k9, i), s);
            s = b2_fma(b11, KV(k11, i), s);
            s = b2_fma(b12, KV(k12, i), s);
            s = b2_fma(b13, KV(k13, i), s);
            s = b2_fma(b14, KV(k14, i), s);
            s = b2_fma(b15, KV(k15, i), s);
            s = b2_fma(b16, KV(k16, i), s);
            out[i] = b2_fma(dt, s, up[i]);
            B2_KBAR;
        }
    }
    // coefficient form of the order-6 interpolant for ONE component (event search): C_j = sum_s R_s,j k_s[i]
    static constexpr int DEG = 6;
    __device__ __forceinline__ void poly_coeffs(int i, real (&c)[DEG]) const {
#define V7C(j)                                           \
    {                                                    \
        real s = V7(R01_##j) * KV(k1, i);                \
        s = b2_fma(V7(R04_##j), KV(k4, i), s);           \
        s = b2_fma(V7(R05_##j), KV(k5, i), s);           \
        s = b2_fma(V7(R06_##j), KV(k6, i), s);           \
        s = b2_fma(V7(R07_##j), KV(k7, i), s);           \
        s = b2_fma(V7(R08_##j), KV(k8, i), s);           \
        s = b2_fma(V7(R09_##j), KV(k9, i), s);           \
        s = b2_fma(V7(R11_##j), KV(k11, i), s);          \
        s = b2_fma(V7(R12_##j), KV(k12, i), s);          \
        s = b2_fma(V7(R13_##j), KV(k13, i), s);          \
        s = b2_fma(V7(R14_##j), KV(k14, i), s);          \
        s = b2_fma(V7(R15_##j), KV(k15, i), s);          \
        s = b2_fma(V7(R16_##j), KV(k16, i), s);          \
        c[j - 1] = s;                                    \
    }
        V7C(1) V7C(2) V7C(3) V7C(4) V7C(5) V7C(6)
#undef V7C
    }
    __device__ __forceinline__ void advance() {
#if B2_KSMEM
        real* t_ = k1;
        k1 = k11;
        k11 = t_;
#else
#pragma unroll
        for (int i = 0; i < B2_NV; i++) k1[i] = k11[i];
#endif
    }
};
#undef V7
#undef V7X
