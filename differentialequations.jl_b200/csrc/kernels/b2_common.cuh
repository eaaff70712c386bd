// b2_common.cuh -- shared device code of the B200 ensemble kernels (sm_100a, NVRTC-compiled).
//
// One trajectory per thread, state in registers.  The generated prelude defines
//   B2_F64 (0/1), B2_NSTATE, B2_NPARAM, B2_ALG, B2_HAS_JAC/TGRAD/NOISE/EVENT
// and the model's device functions b2_rhs / b2_jac / ... before including this file.
//
// Arithmetic contract (mirrors, but does not share code with, oracle/oracle_impl.inc):
// every multiply-add that is meant to fuse is an explicit b2_fma(); NVRTC runs with
// --fmad=false so nothing else contracts.  Linear combinations accumulate left to right:
//   s = c1*x1; s = fma(c2,x2,s); ...   stage argument U = fma(dt, s, uprev).
#pragma once

// sreal: the scalar storage type of u, p, t in global memory; real: the type the steppers compute in (the same).
#if B2_F64
typedef double sreal;
#define B2_EPS 2.220446049250313e-16
#else
typedef float sreal;
#define B2_EPS 1.1920928955078125e-7f
#endif
typedef sreal real;

#ifndef B2_COND_MASK
#define B2_COND_MASK 0xffffffffu   // components of u the ContinuousCallback condition reads (bit i = u[i])
#endif
#ifndef B2_HAS_DEVENT
#define B2_HAS_DEVENT 0
#endif
// ContinuousCallback direction (an upcrossing triggers affect!, a downcrossing affect_neg!).  The condition
// source redefines B2_EVENT_DIR: 0 both directions, +1 upcrossings only (affect_neg! = nothing), -1 downcrossings only
// (affect! = nothing); the affect source sets B2_HAS_AFFECT_NEG 1 when it also defines b2_affect_neg (a downcrossing of
// a scalar callback then runs it instead of b2_affect).  The direction is the sign of the condition at the step start.
#define B2_EVENT_DIR 0
#define B2_HAS_AFFECT_NEG 0
#ifndef B2_HAS_MASS
#define B2_HAS_MASS 0   // 1: the model source defines the constant mass matrix B2_MASS_[n*n] (Rodas family: M u' = f)
#endif
#ifndef B2_KSMEM
#define B2_KSMEM 0   // 1: ERK stage vectors live in shared memory (large n_state), see b2_erk.cuh
#endif
#define B2_N B2_NSTATE
// solve(...; save_idxs = [...]): only these components of the state are saved, in this order.  The generated prelude
// then defines B2_NOUT and `static constexpr int B2_SAVE_IDXS_[B2_NOUT]`; out_u rows have B2_NOUT entries.
#ifdef B2_NOUT
#define B2_SIDX(k) (B2_SAVE_IDXS_[k])
#define B2_HAS_SAVE_IDXS 1
#else
#define B2_NOUT B2_NSTATE
#define B2_SIDX(k) (k)
#define B2_HAS_SAVE_IDXS 0
#endif
#define B2_NPA (B2_NPARAM > 0 ? B2_NPARAM : 1)
#define B2_FULL 0xffffffffu

// retcodes (include/b200ens.h enum b200ens_retcode); 0 doubles as "still running"
#define B2_RC_SUCCESS 1
#define B2_RC_TERMINATED 2
#define B2_RC_MAXITERS 3
#define B2_RC_DTLESSTHANMIN 4
#define B2_RC_UNSTABLE 5
#define B2_RC_DTNAN 6
#define B2_RC_FAILURE 7

struct B2Stats {
    int naccept, nreject, nf, nevents;
};

// Kernel argument block; one layout for f32 and f64 (scalars travel as double).
struct B2Args {
    const void* u0;       // [N][n_state]
    const void* p;        // [N][n_param]
    const void* saveat;   // [n_save]
    const void* dW;       // injected increments or null
    void* out_u;          // [N][n_save][n_state]
    int* retcode;         // [N]
    B2Stats* stats;       // [N] or null
    unsigned long long* work_counter;
    long long N;
    long long maxiters;
    long long nsteps_noise;
    unsigned long long seed;
    unsigned long long traj_offset;
    double t0, t1, dt, abstol, reltol, dtmin, dtmax, qmin, qmax, gamma, beta1, beta2, qoldinit;
    int n_save, adaptive, refill_threshold, stage_stride;
    int noise_injected, event_terminate, interp_points, save_tstops;
    // the same scalars pre-converted to float by the host (f32 kernels read these: no F2F in the loop)
    float f_t0, f_t1, f_dt, f_abstol, f_reltol, f_dtmin, f_dtmax, f_qmin, f_qmax, f_gamma, f_beta1, f_beta2, f_qoldinit;
    int pad0_;
    const unsigned* perm;  // queue position -> trajectory index (expected-work order, b2_work.cuh) or null = identity
    // per-component tolerances (solve(...; abstol = [..], reltol = [..]); scalars are broadcast by the host).  Read with
    // compile-time indices from the constant bank: no registers, FMA constant operands.
    double tol_a[32], tol_r[32];
    float f_tol_a[32], f_tol_r[32];
    // save_everystep: out_u is [N][n_save = capacity][n_state], every_t [N][capacity] receives the step times; slot 0 is
    // (t0, u0), slot k the state after the k-th accepted step; slots past the capacity are dropped (stats.naccept tells)
    void* every_t;
    int save_every, pad1_;
    // fused ensemble moments (b200ens_solve_moments, rows of >= 1024 values): when mom_sum is set the ODE kernels ADD
    // every saved value and its square to mom_sum / mom_sq [n_save][n_state] (double, global reductions) instead of
    // storing it to out_u, which is then never allocated; a trajectory that ends in a failure bumps mom_fail (the host
    // recomputes such a chunk through out_u, so only successful trajectories ever count).
    double* mom_sum;
    double* mom_sq;
    unsigned long long* mom_fail;
    // solve(...; tstops = [...]): times the integrator must hit exactly (handle_tstop!), ascending, of the
    // state type; read by the generic entries only
    const void* tstops;
    int n_tstops, pad2_;
};

#if B2_F64
#define B2_ARG(a, name) ((a).name)
#define B2_ATOL(a, i) ((a).tol_a[i])
#define B2_RTOL(a, i) ((a).tol_r[i])
#else
#define B2_ARG(a, name) ((a).f_##name)
#define B2_ATOL(a, i) ((a).f_tol_a[i])
#define B2_RTOL(a, i) ((a).f_tol_r[i])
#endif

__device__ __forceinline__ float b2_fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ double b2_fma(double a, double b, double c) { return __fma_rn(a, b, c); }
__device__ __forceinline__ float b2_abs(float a) { return fabsf(a); }
__device__ __forceinline__ double b2_abs(double a) { return fabs(a); }
__device__ __forceinline__ float b2_max(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ double b2_max(double a, double b) { return fmax(a, b); }
__device__ __forceinline__ float b2_min(float a, float b) { return fminf(a, b); }
__device__ __forceinline__ double b2_min(double a, double b) { return fmin(a, b); }
__device__ __forceinline__ float b2_sqrt(float a) { return __fsqrt_rn(a); }
__device__ __forceinline__ double b2_sqrt(double a) { return __dsqrt_rn(a); }
__device__ __forceinline__ bool b2_isnan(sreal a) { return a != a; }

// ---- packed component pairs (Float32, one trajectory per thread): Blackwell's FFMA2 / FMUL2 / FADD2 apply the IEEE
// round-to-nearest FP32 operation to both halves of a 64-bit register pair and take a 32-bit immediate or a scalar
// register broadcast as the other operand, so the linear combinations of two state COMPONENTS of one trajectory cost
// one issue slot instead of two -- same bits, no extra registers (profiles/README.md, round 2).  B2_PACK2 = 1 turns the
// chunked loops of b2_erk.cuh into pair loops (+ a scalar tail for odd n); everywhere else a chunk is one scalar.
#ifndef B2_PACK2
#define B2_PACK2 (!B2_F64 && !B2_KSMEM)
#endif
#if B2_PACK2
struct b2p {
    float2 v;
};
__device__ __forceinline__ b2p operator*(b2p a, b2p b) { b2p r; r.v = __fmul2_rn(a.v, b.v); return r; }
__device__ __forceinline__ b2p b2_fma(b2p a, b2p b, b2p c) { b2p r; r.v = __ffma2_rn(a.v, b.v, c.v); return r; }
template <class V> __device__ __forceinline__ V b2_bc(float x);
template <> __device__ __forceinline__ float b2_bc<float>(float x) { return x; }
template <> __device__ __forceinline__ b2p b2_bc<b2p>(float x) { b2p r; r.v = make_float2(x, x); return r; }
template <class V> __device__ __forceinline__ V b2_ld(const float* a, int i);
template <> __device__ __forceinline__ float b2_ld<float>(const float* a, int i) { return a[i]; }
template <> __device__ __forceinline__ b2p b2_ld<b2p>(const float* a, int i) { b2p r; r.v = make_float2(a[i], a[i + 1]); return r; }
__device__ __forceinline__ void b2_st(float* a, int i, float v) { a[i] = v; }
__device__ __forceinline__ void b2_st(float* a, int i, b2p v) { a[i] = v.v.x; a[i + 1] = v.v.y; }
#endif

// ---- deterministic float log2 / exp2 for the PI controller (same primitive sequence as the
// oracle's orc_fastlog2 / orc_fastexp2; restates upstream's approximate FastPower):
//   log2(x) = e + t*P5(t), t = mantissa-1;   2^y = 2^rint(y) * Q6(y - rint(y))
__device__ __forceinline__ float b2_fastlog2(float x) {
    const unsigned ix = __float_as_uint(x);
    const int e = (int)(ix >> 23) - 127;
    const float m = __uint_as_float((ix & 0x007fffffu) | 0x3f800000u);
    const float t = __fsub_rn(m, 1.0f);
    float p = -0.02645725943148136f;
    p = __fmaf_rn(p, t, 0.12345092743635178f);
    p = __fmaf_rn(p, t, -0.27953752875328064f);
    p = __fmaf_rn(p, t, 0.45827049016952515f);
    p = __fmaf_rn(p, t, -0.7182818651199341f);
    p = __fmaf_rn(p, t, 1.442553162574768f);
    return __fmaf_rn(t, p, (float)e);
}
__device__ __forceinline__ float b2_fastexp2(float y) {
    y = fminf(fmaxf(y, -125.0f), 125.0f);
    const float fi = rintf(y);
    const float f = __fsub_rn(y, fi);
    float p = 1.5403530e-4f;
    p = __fmaf_rn(p, f, 1.3333558e-3f);
    p = __fmaf_rn(p, f, 9.6181291e-3f);
    p = __fmaf_rn(p, f, 5.5504109e-2f);
    p = __fmaf_rn(p, f, 2.4022651e-1f);
    p = __fmaf_rn(p, f, 6.9314718e-1f);
    p = __fmaf_rn(p, f, 1.0f);
    return __uint_as_float(__float_as_uint(p) + (unsigned)((int)fi << 23));
}

// ---- deterministic Float32 reciprocal for the error norm: exponent-flip initial guess (relative error <= 5.1 %)
// + two Newton steps x <- x + x*(1 - s*x) written as FMAs -> relative error <= 6.6e-6, seven issue slots, no MUFU, no
// slow-path branch (an IEEE division is 13-15 slots with its FCHK / BSSY / BSYNC scaffolding, profiles/README.md).
// Same primitive sequence as the oracle's orc_rcp_nr.  s = +0 -> +inf, s = +inf -> -inf (an overflowed proposal is
// REJECTED through r = utilde * -inf, not accepted with a zero estimate), NaN -> NaN.
__device__ __forceinline__ float b2_rcp_nr(float s) {
    float x = __uint_as_float(0x7EF311C7u - __float_as_uint(s));
    x = __fmaf_rn(x, __fmaf_rn(-s, x, 1.0f), x);
    x = __fmaf_rn(x, __fmaf_rn(-s, x, 1.0f), x);
    return x;
}

// theta = (tau - tprev) / dt of a saveat point.  Float32: multiply by a three-step Newton reciprocal (relative error
// <= 9e-8, under one ulp of theta; 8 issue slots against 13-15 for the IEEE division with its slow-path scaffolding --
// the saveat block runs at ~5/32 lanes in 9 of 10 iterations, profiles/README.md).  Float64: IEEE division.
__device__ __forceinline__ float b2_theta(float num, float dt) {
    float x = __uint_as_float(0x7EF311C7u - __float_as_uint(dt));
    x = __fmaf_rn(x, __fmaf_rn(-dt, x, 1.0f), x);
    x = __fmaf_rn(x, __fmaf_rn(-dt, x, 1.0f), x);
    x = __fmaf_rn(x, __fmaf_rn(-dt, x, 1.0f), x);
    return __fmul_rn(num, x);
}
__device__ __forceinline__ double b2_theta(double num, double dt) { return num / dt; }

// ---- ITP root-find helper: pw = eps * 2^(k+1), eps = 2*B2_EPS, k = number of halvings of wd until wd <= 2*eps
// (the oracle computes it with that loop; ~50 iterations in Float64).  All quantities are powers of two, so the
// closed form from the exponent bits is exact: with wd = m * 2^e, m in [1,2): k = e - tau + (m > 1), 2*eps = 2^tau.
__device__ __forceinline__ float b2_itp_pw(float wd) {
    const int tau = -21;   // 2*eps = 2 * 2 * 2^-23
    int k = 0;
    if (wd > 4.76837158203125e-07f) {   // 2^-21
        const unsigned b = __float_as_uint(wd);
        k = (int)(b >> 23) - 127 - tau + ((b & 0x007fffffu) ? 1 : 0);
    }
    return __uint_as_float((unsigned)(tau + k + 127) << 23);
}
__device__ __forceinline__ double b2_itp_pw(double wd) {
    const int tau = -50;   // 2*eps = 2 * 2 * 2^-52
    int k = 0;
    if (wd > 8.8817841970012523e-16) {   // 2^-50
        const unsigned long long b = (unsigned long long)__double_as_longlong(wd);
        k = (int)(b >> 52) - 1023 - tau + ((b & 0x000fffffffffffffull) ? 1 : 0);
    }
    return __longlong_as_double((long long)(tau + k + 1023) << 52);
}

// ---- per-lane output sink: shared-memory staging (flushed coalesced by the whole warp
// when the lane retires) or direct global stores for outputs too large to stage.
struct B2Sink {
    sreal* stage;       // this lane's staging row (null in direct mode)
    sreal* gout;       // out_u as real*
    long long base;    // idx * n_save * B2_NOUT
    const B2Args* margs;   // fused-moments mode: the kernel's argument block (generic entries; nullptr in the specialised
                           // ones folds everything away).  The accumulator pointers are read from the constant bank
                           // where they are used instead of living in registers for the whole loop.
    __device__ __forceinline__ bool moments() const { return margs != nullptr && margs->mom_sum != nullptr; }
    __device__ __forceinline__ void put(int si, const sreal (&v)[B2_N]) const {
        if (moments()) {
            double* const ms = margs->mom_sum + si * B2_NOUT;
            double* const mq = margs->mom_sq + si * B2_NOUT;
#pragma unroll
            for (int k = 0; k < B2_NOUT; k++) {
                const double x = (double)v[B2_SIDX(k)];
                atomicAdd(ms + k, x);
                atomicAdd(mq + k, x * x);
            }
        } else if (stage) {
#pragma unroll
            for (int k = 0; k < B2_NOUT; k++) stage[si * B2_NOUT + k] = v[B2_SIDX(k)];
        } else {
#pragma unroll
            for (int k = 0; k < B2_NOUT; k++) gout[base + (long long)si * B2_NOUT + k] = v[B2_SIDX(k)];
        }
    }
    __device__ __forceinline__ void fill(int si, int n_save, sreal v) const {
        if (moments()) return;   // failures are counted, not accumulated (B2Args.mom_fail)
        for (; si < n_save; si++) {
#pragma unroll
            for (int k = 0; k < B2_NOUT; k++) {
                if (stage) stage[si * B2_NOUT + k] = v;
                else gout[base + (long long)si * B2_NOUT + k] = v;
            }
        }
    }
};

// Warp-cooperative flush of the staged outputs of every lane in `dirty_mask`: all 32 lanes
// copy one retired lane's row at a time -> coalesced 128-byte global stores.
__device__ __forceinline__ void b2_flush(unsigned dirty_mask, const sreal* warp_stage, int stride, sreal* gout,
                                         long long idx, int out_per_traj, unsigned lane) {
    __syncwarp();
    while (dirty_mask) {
        const int L = __ffs(dirty_mask) - 1;
        dirty_mask &= dirty_mask - 1;
        const long long iL = __shfl_sync(B2_FULL, idx, L);
        const sreal* src = warp_stage + (size_t)L * stride;
        sreal* dst = gout + iL * (long long)out_per_traj;
        for (int j = lane; j < out_per_traj; j += 32) dst[j] = src[j];
    }
    __syncwarp();
}

// Warp-aggregated work fetch: one atomicAdd per warp hands consecutive trajectory indices
// to the idle lanes (in lane order).  Returns this lane's index or -1.
__device__ __forceinline__ long long b2_fetch(unsigned idle_mask, unsigned long long* counter, long long N,
                                              unsigned lane, bool& exhausted) {
    const int cnt = __popc(idle_mask);
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(counter, (unsigned long long)cnt);
    base = __shfl_sync(B2_FULL, base, 0);
    if ((long long)base + cnt >= N) exhausted = true;
    if (!((idle_mask >> lane) & 1u)) return -1;
    const long long my = (long long)base + __popc(idle_mask & ((1u << lane) - 1u));
    return my < N ? my : -1;
}
