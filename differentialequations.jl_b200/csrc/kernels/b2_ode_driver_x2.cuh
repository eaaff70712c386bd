// b2_ode_driver_x2.cuh -- the persistent ODE ensemble kernel with TWO trajectories per thread in packed
// FP32 (Blackwell FFMA2 / FADD2 / FMUL2, sm_100+ only).
//
// Why: the scalar Tsit5 kernel is issue-bound (ncu: 86% of issue slots used, FMA pipe only 47% busy --
// profiles/r1_tsit5_f32_ncu_full.summary.txt).  Packing two independent trajectories into the halves of a
// 64-bit register pair lets every stage combination, RHS operation, error-estimate and interpolant FMA be ONE
// issue slot for two trajectories, moving the bound from the scheduler to the FMA pipe itself.  Each packed
// operation is the exact per-half IEEE operation, so results are bit-identical to the scalar kernel and to
// the CPU oracle (same parity tests).
//
// Control (step-size controller, accept/reject, saveat bookkeeping, retire/refill) stays scalar and is done
// per half: each thread owns two independent "slots" h = 0, 1 with their own t, dt, controller memory,
// trajectory index and counters.  Same semantics as b2_ode_driver.cuh; no callbacks,
// direct global stores only.
#pragma once
#include "b2_common.cuh"

#if B2_X2

// packed versions of the deterministic log2 / exp2 (identical per-half arithmetic to b2_fastlog2/b2_fastexp2)
__device__ __forceinline__ float2 b2_fastlog2_x2(float2 x) {
    const unsigned ix = __float_as_uint(x.x), iy = __float_as_uint(x.y);
    const float2 e = make_float2((float)((int)(ix >> 23) - 127), (float)((int)(iy >> 23) - 127));
    const float2 m = make_float2(__uint_as_float((ix & 0x007fffffu) | 0x3f800000u), __uint_as_float((iy & 0x007fffffu) | 0x3f800000u));
    const float2 t = __fadd2_rn(m, make_float2(-1.0f, -1.0f));
    float2 p = make_float2(-0.02645725943148136f, -0.02645725943148136f);
    p = __ffma2_rn(p, t, make_float2(0.12345092743635178f, 0.12345092743635178f));
    p = __ffma2_rn(p, t, make_float2(-0.27953752875328064f, -0.27953752875328064f));
    p = __ffma2_rn(p, t, make_float2(0.45827049016952515f, 0.45827049016952515f));
    p = __ffma2_rn(p, t, make_float2(-0.7182818651199341f, -0.7182818651199341f));
    p = __ffma2_rn(p, t, make_float2(1.442553162574768f, 1.442553162574768f));
    return __ffma2_rn(t, p, e);
}
__device__ __forceinline__ float2 b2_fastexp2_x2(float2 y) {
    y.x = fminf(fmaxf(y.x, -125.0f), 125.0f);
    y.y = fminf(fmaxf(y.y, -125.0f), 125.0f);
    const float2 fi = make_float2(rintf(y.x), rintf(y.y));
    const float2 f = __fadd2_rn(y, make_float2(-fi.x, -fi.y));
    float2 p = make_float2(1.5403530e-4f, 1.5403530e-4f);
    p = __ffma2_rn(p, f, make_float2(1.3333558e-3f, 1.3333558e-3f));
    p = __ffma2_rn(p, f, make_float2(9.6181291e-3f, 9.6181291e-3f));
    p = __ffma2_rn(p, f, make_float2(5.5504109e-2f, 5.5504109e-2f));
    p = __ffma2_rn(p, f, make_float2(2.4022651e-1f, 2.4022651e-1f));
    p = __ffma2_rn(p, f, make_float2(6.9314718e-1f, 6.9314718e-1f));
    p = __ffma2_rn(p, f, make_float2(1.0f, 1.0f));
    return make_float2(__uint_as_float(__float_as_uint(p.x) + (unsigned)((int)fi.x << 23)),
                       __uint_as_float(__float_as_uint(p.y) + (unsigned)((int)fi.y << 23)));
}

template <class Alg>
__device__ __forceinline__ void b2_ode_driver_x2(const B2Args& a) {
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    float* const gout = reinterpret_cast<float*>(a.out_u);
    const float* const gu0 = reinterpret_cast<const float*>(a.u0);
    const float* const gp = reinterpret_cast<const float*>(a.p);
    const float* const gsave = reinterpret_cast<const float*>(a.saveat);
    const int n_save = a.n_save;
    const long long out_per_traj = (long long)n_save * B2_N;

    const float t0 = a.f_t0, t1 = a.f_t1, dt_user = a.f_dt;
    const float qmax = a.f_qmax, qmin = a.f_qmin, gam = a.f_gamma;
    const float inv_qmax = 1.0f / qmax, inv_qmin = 1.0f / qmin, inv_gam = 1.0f / gam;
    const real inv_n = real(1.0f / (float)B2_N);
    const float qoldinit = a.f_qoldinit, dtmax = a.f_dtmax, dtmin = a.f_dtmin;
    const float beta1 = a.f_beta1, beta2 = a.f_beta2;
    const float lqinit = b2_fastlog2(qoldinit);
    const bool adaptive = a.adaptive != 0;
    const bool save_tstops = a.save_tstops != 0;
    const int maxiters = a.maxiters > 0x7fffffffLL ? 0x7fffffff : (int)a.maxiters;

    Alg alg;
    alg.bind(nullptr);
    real u[B2_N], p[B2_NPA];
    float t[2] = {t0, t0}, dt[2] = {dt_user, dt_user}, lq[2] = {lqinit, lqinit};
    long long idx[2] = {-1, -1};
    int iter[2] = {0, 0}, si[2] = {0, 0}, naccept[2] = {0, 0}, nreject[2] = {0, 0}, nf[2] = {0, 0};
    bool active[2] = {false, false};
    bool exhausted = false;
#pragma unroll
    for (int i = 0; i < B2_N; i++) u[i] = real(0.0f);
#pragma unroll
    for (int i = 0; i < B2_NPA; i++) p[i] = real(0.0f);
#pragma unroll
    for (int i = 0; i < B2_N; i++) alg.k1[i] = real(0.0f);

    for (;;) {
        // ---------------- phase 0: retire / refill over the 64 slots of the warp
        const unsigned idle0 = __ballot_sync(B2_FULL, !active[0]);
        const unsigned idle1 = __ballot_sync(B2_FULL, !active[1]);
        const int n_idle = __popc(idle0) + __popc(idle1);
        if (n_idle) {
            const bool all_idle = n_idle == 64;
            if (exhausted ? all_idle : (n_idle >= a.refill_threshold || all_idle)) {
                if (!exhausted) {
                    unsigned long long base = 0;
                    if (lane == 0) base = atomicAdd(a.work_counter, (unsigned long long)n_idle);
                    base = __shfl_sync(B2_FULL, base, 0);
                    if ((long long)base + n_idle >= a.N) exhausted = true;
                    const long long my[2] = {(long long)base + __popc(idle0 & lt_mask),
                                             (long long)base + __popc(idle0) + __popc(idle1 & lt_mask)};
                    bool fresh[2] = {false, false};
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        if (!active[h] && my[h] < a.N) {
                            idx[h] = my[h];
                            float v[B2_N];
#pragma unroll
                            for (int i = 0; i < B2_N; i++) {
                                v[i] = gu0[idx[h] * B2_N + i];
                                b2_set(u[i], h, v[i]);
                            }
#pragma unroll
                            for (int i = 0; i < B2_NPARAM; i++) b2_set(p[i], h, gp[idx[h] * B2_NPARAM + i]);
                            t[h] = t0;
                            dt[h] = dt_user;
                            lq[h] = lqinit;
                            iter[h] = 0;
                            si[h] = 0;
                            naccept[h] = nreject[h] = 0;
                            nf[h] = 1;
                            // the first saved value is u0 itself (test/core.jl:34)
                            while (si[h] < n_save && __ldg(gsave + si[h]) <= t0) {
#pragma unroll
                                for (int i = 0; i < B2_N; i++) gout[idx[h] * out_per_traj + (long long)si[h] * B2_N + i] = v[i];
                                si[h]++;
                            }
                            active[h] = true;
                            fresh[h] = true;
                        }
                    }
                    if (fresh[0] || fresh[1]) alg.start_masked(u, p, real(t[0], t[1]), fresh[0], fresh[1]);
                }
                if (__ballot_sync(B2_FULL, active[0] || active[1]) == 0u) break;
            }
        }

        // ---------------- phase 1: pre-step control per slot (loopheader!/check_error!)
        int rc[2] = {0, 0};
        bool do_step[2] = {false, false}, accepted[2] = {false, false};
        float tprev[2] = {t[0], t[1]}, tnew[2] = {t[0], t[1]}, dts[2], dtnew[2], tstop[2] = {t1, t1};
#pragma unroll
        for (int h = 0; h < 2; h++) {
            if (active[h]) {
                iter[h]++;
                if (!adaptive) dt[h] = dt_user;
                if (save_tstops && si[h] < n_save) {
                    const float s = __ldg(gsave + si[h]);
                    if (s < t1) tstop[h] = s;
                }
                const bool clipped = dt[h] > tstop[h] - t[h];
                if (clipped) dt[h] = tstop[h] - t[h];
                const bool toosmall = dt[h] <= fmaxf(dtmin, (float)B2_EPS * fabsf(t[h]));
                if (iter[h] > maxiters) rc[h] = B2_RC_MAXITERS;
                else if (dt[h] != dt[h]) rc[h] = B2_RC_DTNAN;
                else if (adaptive & !clipped & toosmall) rc[h] = B2_RC_DTLESSTHANMIN;
                do_step[h] = rc[h] == 0;
            }
            dts[h] = dt[h];
            dtnew[h] = dt[h];
        }
        __syncwarp();  // the whole warp enters the packed stepper together

        // ---------------- one packed step attempt for both slots + per-slot controller
        real un[B2_N], ut[B2_N];
        if (do_step[0] || do_step[1]) {
            int nfd = 0;
            alg.step(u, p, real(t[0], t[1]), real(dts[0], dts[1]), un, ut, adaptive, nfd);
            if (adaptive) {
                // error norm (A.4), packed; accept iff EEst^2 <= 1; PI controller (A.5) in the log domain
                real acc = real(0.0f);
#pragma unroll
                for (int i = 0; i < B2_N; i++) {
                    const real sk = b2_fma(b2_max(b2_abs(u[i]), b2_abs(un[i])), real(a.f_tol_r[i]), real(a.f_tol_a[i]));
                    const real r = ut[i] / sk;
                    acc = b2_fma(r, r, acc);
                }
                const real EE2 = acc * inv_n;
                const float2 l2 = __fmul2_rn(make_float2(0.5f, 0.5f), b2_fastlog2_x2(EE2.v));
                const float2 ex = __ffma2_rn(make_float2(-beta2, -beta2), make_float2(lq[0], lq[1]),
                                             __fmul2_rn(make_float2(beta1, beta1), l2));
                const float2 qraw = b2_fastexp2_x2(ex);
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    if (do_step[h]) {
                        nf[h] += nfd;
                        const float ee = b2_get(EE2, h);
                        if (ee != ee) {
                            rc[h] = B2_RC_DTNAN;  // upstream: NaN EEst -> NaN dt -> ReturnCode.DtNaN
                        } else {
                            float q, l = lqinit;
                            if (ee == 0.0f) {
                                q = inv_qmax;
                            } else {
                                l = h ? l2.y : l2.x;
                                q = fmaxf(inv_qmax, fminf(inv_qmin, __fmul_rn(h ? qraw.y : qraw.x, inv_gam)));
                            }
                            if (!(ee <= 1.0f)) {
                                nreject[h]++;
                                const float q11 = b2_fastexp2(__fmul_rn(beta1, l));
                                dt[h] = __fmul_rn(dt[h], __fdiv_rn(1.0f, fminf(inv_qmin, __fmul_rn(q11, inv_gam))));
                            } else {
                                accepted[h] = true;
                                lq[h] = fmaxf(l, lqinit);
                                dtnew[h] = __fmul_rn(dt[h], __fdiv_rn(1.0f, q));
                            }
                        }
                    }
                }
            } else {
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    if (do_step[h]) {
                        nf[h] += nfd;
                        bool bad = false;
#pragma unroll
                        for (int i = 0; i < B2_N; i++) bad |= b2_get(un[i], h) != b2_get(un[i], h);
                        if (bad) rc[h] = B2_RC_UNSTABLE;
                        else accepted[h] = true;
                    }
                }
            }
#pragma unroll
            for (int h = 0; h < 2; h++) {
                if (accepted[h]) {
                    naccept[h]++;
                    tnew[h] = __fadd_rn(t[h], dts[h]);
                    if (fabsf(tnew[h] - tstop[h]) < 100.0f * (float)B2_EPS * fmaxf(fabsf(tnew[h]), fabsf(tstop[h]))) tnew[h] = tstop[h];
                }
            }
        }

        // ---------------- phase 2: saveat through the dense output (A.6), warp-convergent, packed
        for (;;) {
            float tau[2] = {0.0f, 0.0f};
            bool need[2] = {false, false};
#pragma unroll
            for (int h = 0; h < 2; h++) {
                if (accepted[h] && si[h] < n_save) {
                    tau[h] = __ldg(gsave + si[h]);
                    need[h] = tau[h] <= tnew[h];
                }
            }
            if (!__any_sync(B2_FULL, need[0] || need[1])) break;
            const real th(__fdiv_rn(tau[0] - tprev[0], dts[0]), __fdiv_rn(tau[1] - tprev[1], dts[1]));
            real w[B2_N];
            alg.interp(u, un, th, real(dts[0], dts[1]), w);
#pragma unroll
            for (int h = 0; h < 2; h++) {
                if (need[h]) {
                    const bool at_end = tau[h] == tnew[h];
                    float* dst = gout + idx[h] * out_per_traj + (long long)si[h] * B2_N;
#pragma unroll
                    for (int i = 0; i < B2_N; i++) dst[i] = at_end ? b2_get(un[i], h) : b2_get(w[i], h);
                    si[h]++;
                }
            }
        }

        // ---------------- phase 3: commit accepted slots (FSAL hand-over, next dt)
        if (accepted[0] || accepted[1]) {
#pragma unroll
            for (int i = 0; i < B2_N; i++) u[i] = b2_blend(accepted[0], accepted[1], un[i], u[i]);
            alg.advance_masked(accepted[0], accepted[1]);
#pragma unroll
            for (int h = 0; h < 2; h++) {
                if (accepted[h]) {
                    t[h] = tnew[h];
                    if (adaptive) dt[h] = fminf(dtmax, dtnew[h]);
                    if (rc[h] == 0 && !(t[h] < t1)) rc[h] = B2_RC_SUCCESS;
                }
            }
        }

        // ---------------- phase 4: retire finished / failed slots
#pragma unroll
        for (int h = 0; h < 2; h++) {
            if (rc[h] != 0) {
                if (rc[h] != B2_RC_SUCCESS) {
                    for (; si[h] < n_save; si[h]++) {
#pragma unroll
                        for (int i = 0; i < B2_N; i++)
                            gout[idx[h] * out_per_traj + (long long)si[h] * B2_N + i] = __int_as_float(0x7fc00000);
                    }
                }
                a.retcode[idx[h]] = rc[h];
                if (a.stats) {
                    B2Stats s;
                    s.naccept = naccept[h];
                    s.nreject = nreject[h];
                    s.nf = nf[h];
                    s.nevents = 0;
                    a.stats[idx[h]] = s;
                }
                active[h] = false;
            }
        }
    }
}
#endif  // B2_X2
