// b2_ode_driver.cuh -- the persistent one-trajectory-per-thread ODE ensemble kernel body.
//
// Replaces, for the whole ensemble at once, the per-trajectory integrator loop that
// EnsembleThreads runs on the CPU (SciMLBase.__solve/solve_batch/batch_func ->
// OrdinaryDiffEq.solve!: loopheader!/perform_step!/loopfooter!/savevalues!/handle_callbacks!;
// call sites /root/reference/test/core.jl:14,32,47,72,93; semantics SURVEY.md A.1-A.8).
//
// Structure: every warp is persistent.  Each outer iteration a lane performs ONE step
// attempt of its trajectory.  Lanes whose trajectory finished are parked; when at least
// `refill_threshold` lanes of a warp are parked (warp vote), the warp flushes the parked
// lanes' staged saveat outputs with coalesced stores and refills them from a global atomic
// work counter, so SIMT lanes stay busy although step counts differ 10x between
// trajectories (0.54 warp efficiency without refill).
#pragma once
#include "b2_common.cuh"
#include "b2_control.cuh"

// Steppers with their own step-size / order control (FBDF, b2_bdf.cuh) declare `static constexpr bool OWN_CONTROL = true`:
// their step() reports a failed Newton iteration and the accept / reject decisions call back into the stepper.
template <class A, class = void>
struct b2_own_control {
    static constexpr bool value = false;
};
template <class A>
struct b2_own_control<A, decltype((void)A::OWN_CONTROL)> {
    static constexpr bool value = A::OWN_CONTROL;
};

// ADAPT / TSTOPS: 0 or 1 = compile-time specialisation of the two solve options that sit in the per-iteration
// control path, -1 = read them from the argument block.  AUTODT = 0 compiles the automatic-initial-step block out
// (the specialised entry is only launched with a caller-supplied dt; keeps its register pressure down).
template <class Alg, int ADAPT = -1, int TSTOPS = -1, int AUTODT = 1, int STAGED = -1>
__device__ __forceinline__ void b2_ode_driver(const B2Args& a) {
    extern __shared__ __align__(16) unsigned char b2_smem[];
    const unsigned lane = threadIdx.x & 31u;
    const int warp_in_block = threadIdx.x >> 5;
    // STAGED = 0: output staging compiled out (direct global stores, the default and the faster mode, profiles/)
    const int stride = STAGED == 0 ? 0 : a.stage_stride;
    real* const warp_stage = reinterpret_cast<real*>(b2_smem) + (size_t)warp_in_block * 32 * stride;
    real* const gout = reinterpret_cast<real*>(a.out_u);
    const real* const gu0 = reinterpret_cast<const real*>(a.u0);
    const real* const gp = reinterpret_cast<const real*>(a.p);
    const real* const gsave = reinterpret_cast<const real*>(a.saveat);
    const int n_save = a.n_save;
    const int out_per_traj = n_save * B2_NOUT;

    const real t0 = B2_ARG(a, t0), t1 = B2_ARG(a, t1), dt_user = B2_ARG(a, dt);
    const B2Ctl ctl = b2_ctl_init(a);   // the PI controller works in Float32 (b2_control.cuh)
    const float lqinit = ctl.lqinit;
    const float inv_n = __fdiv_rn(1.0f, (float)B2_N);
    const real dtmax = B2_ARG(a, dtmax), dtmin = B2_ARG(a, dtmin);
    constexpr bool EVERY = ADAPT < 0;   // save_everystep exists in the generic entry only (specialised entries: saveat)
    const bool adaptive = ADAPT < 0 ? (a.adaptive != 0) : (ADAPT != 0);
    const bool save_tstops = TSTOPS < 0 ? (a.save_tstops != 0) : (TSTOPS != 0);
    // solve(...; tstops): generic entry only (the specialised entries are not launched with tstops and fold this away)
    const int n_ts = ADAPT < 0 ? a.n_tstops : 0;
    const real* const gts = reinterpret_cast<const real*>(a.tstops);

    Alg alg;
#if B2_KSMEM
    // shared-memory stage vectors sit behind the (optional) output staging area: [vector][component][thread]
    alg.bind(reinterpret_cast<real*>(b2_smem + (size_t)stride * B2_BLOCK * sizeof(real)) + threadIdx.x);
#else
    alg.bind(nullptr);
#endif
    real u[B2_N], p[B2_NPA];
    real t = t0, dt = dt_user;
    float lq = lqinit;
    long long idx = -1;
    int iter = 0;   // step attempts of this lane's trajectory (maxiters is clamped to 2^31 - 1: 32-bit compare per iteration)
    const int maxit = a.maxiters > 0x7fffffffLL ? 0x7fffffff : (int)a.maxiters;
    int si = 0, naccept = 0, nreject = 0, nf = 0, nevents = 0;
    int ti = 0;   // next user tstop of this lane
    // next save time of this lane (cached: the saveat check runs twice per iteration); +inf when none is left
    real tau_next = (real)__int_as_float(0x7f800000);
    bool active = false, dirty = false, exhausted = false;
#if B2_HAS_EVENT
    bool just_fired = false;
    int ev_last = 0;   // VectorContinuousCallback: index of the function that fired the last event
    const int ip = a.interp_points;
#endif
    B2Sink sink;
    sink.stage = stride ? warp_stage + (size_t)lane * stride : nullptr;
    sink.gout = gout;
    sink.base = 0;
    // fused ensemble moments: only the generic entry (ADAPT < 0) reads the pointers, the specialised entries fold them away
    sink.margs = ADAPT < 0 ? &a : nullptr;

    for (;;) {
        // ---------------- phase 0: retire / refill (warp-uniform control flow)
        const unsigned idle = __ballot_sync(B2_FULL, !active);
        if (idle) {
            const bool all_idle = idle == B2_FULL;
            if (exhausted ? all_idle : (__popc(idle) >= a.refill_threshold || all_idle)) {
                if (stride) {
                    const unsigned dmask = __ballot_sync(B2_FULL, dirty);
                    b2_flush(dmask, warp_stage, stride, gout, idx, out_per_traj, lane);
                    dirty = false;
                }
                if (!exhausted) {
                    const long long my = b2_fetch(idle, a.work_counter, a.N, lane, exhausted);
                    if (!active && my >= 0) {
                        idx = a.perm ? (long long)__ldg(a.perm + my) : my;
                        sink.base = idx * (long long)out_per_traj;
#pragma unroll
                        for (int i = 0; i < B2_N; i++) u[i] = gu0[idx * B2_N + i];
#pragma unroll
                        for (int i = 0; i < B2_NPARAM; i++) p[i] = gp[idx * B2_NPARAM + i];
                        t = t0;
                        dt = dt_user;
                        lq = lqinit;
                        iter = 0;
                        si = 0;
                        naccept = nreject = nevents = 0;
                        ti = 0;
                        while (ti < n_ts && __ldg(gts + ti) <= t0) ti++;
                        // the first saved value is u0 itself (test/core.jl:34)
                        if (EVERY && a.save_every) {
                            // save_everystep: slot 0 = (t0, u0); `si` counts the slots written, no saveat grid
                            if (n_save > 0) {
                                sink.put(0, u);
                                reinterpret_cast<real*>(a.every_t)[idx * (long long)n_save] = t0;
                            }
                            si = 1;
                            tau_next = (real)__int_as_float(0x7f800000);
                        } else {
                            while (si < n_save && __ldg(gsave + si) <= t0) {
                                sink.put(si, u);
                                si++;
                            }
                            tau_next = si < n_save ? __ldg(gsave + si) : (real)__int_as_float(0x7f800000);
                        }
                        alg.start(u, p, t);
                        nf = 1;
                        if (AUTODT && adaptive && !(dt_user > (real)0)) {
                            // automatic initial step (Hairer-Norsett-Wanner as in OrdinaryDiffEq's initdt)
                            real a0 = 0, a1 = 0, a2 = 0, u1[B2_N], f1[B2_N];
#pragma unroll
                            for (int i = 0; i < B2_N; i++) {
                                const real sk = b2_fma(b2_abs(u[i]), B2_RTOL(a, i), B2_ATOL(a, i));
                                const real r0 = u[i] / sk, r1 = alg.fsal0(i) / sk;
                                a0 = b2_fma(r0, r0, a0);
                                a1 = b2_fma(r1, r1, a1);
                            }
                            const real d0 = b2_sqrt(a0 / (real)B2_N), d1 = b2_sqrt(a1 / (real)B2_N);
                            real dt0 = (d0 < (real)1e-5 || d1 < (real)1e-5) ? (real)1e-6 : (real)0.01 * (d0 / d1);
                            dt0 = b2_min(dt0, dtmax);
#pragma unroll
                            for (int i = 0; i < B2_N; i++) u1[i] = b2_fma(dt0, alg.fsal0(i), u[i]);
                            b2_rhs(f1, u1, p, t0 + dt0);
                            nf++;
#pragma unroll
                            for (int i = 0; i < B2_N; i++) {
                                const real sk = b2_fma(b2_abs(u[i]), B2_RTOL(a, i), B2_ATOL(a, i));
                                const real r2 = (f1[i] - alg.fsal0(i)) / sk;
                                a2 = b2_fma(r2, r2, a2);
                            }
                            const real d2 = b2_sqrt(a2 / (real)B2_N) / dt0;
                            const real dmx = b2_max(d1, d2);
                            real dt1;
                            if (dmx <= (real)1e-15) dt1 = b2_max((real)1e-6, dt0 * (real)1e-3);
                            else dt1 = (real)b2_fastexp2(__fmul_rn(-__fadd_rn(6.6438562f, b2_fastlog2((float)dmx)),
                                                                  __fdiv_rn(1.0f, (float)Alg::ORDER)));
                            dt = b2_min(b2_min((real)100 * dt0, dt1), dtmax);
                        }
                        active = true;
#if B2_HAS_EVENT
                        just_fired = false;
#endif
                    }
                }
                if (__ballot_sync(B2_FULL, active) == 0u) break;
            }
        }

        // ---------------- phase 1: one step attempt per active lane
        // (loopheader! / check_error! / perform_step! / stepsize controller)
        int rc = 0;
        bool accepted = false, fired = false;
        const real tprev = t;
        real tnew = t, dts = dt, dtnew = dt;
        real un[B2_N], ut[B2_N];
#if B2_HAS_EVENT
        real th_end = 1;
        int ev_idx = 0;   // which event function fired (VectorContinuousCallback)
#endif
        bool do_step = false;
        real tstop = t1;
        if (active) {
            iter++;
            if (!adaptive) dt = dt_user;
            if (save_tstops && tau_next < t1) tstop = tau_next;
            if (ti < n_ts) tstop = b2_min(tstop, __ldg(gts + ti));
            const bool clipped = dt > tstop - t;
            if (clipped) dt = tstop - t;
            const bool toosmall = dt <= b2_max(dtmin, (real)B2_EPS * b2_abs(t));
            if (iter > maxit) rc = B2_RC_MAXITERS;
            else if (b2_isnan(dt)) rc = B2_RC_DTNAN;
            else if (adaptive & !clipped & toosmall) rc = B2_RC_DTLESSTHANMIN;
            do_step = rc == 0;
        }
        // Everything above is cheap per-lane control; the barrier makes the whole warp enter the
        // stepper together.  Without it ptxas merges the `clipped` branch straight into the stepper
        // and the two lane groups run the ~390-instruction body separately (ncu: body executed 1.36x
        // per iteration at 19/32 threads).
        __syncwarp();
        {
            {
                if (do_step) {
                    bool nfail = false;   // OWN_CONTROL steppers: the corrector's Newton iteration did not converge
                    if constexpr (b2_own_control<Alg>::value) nfail = alg.step(u, p, t, dt, un, ut, a, nf);
                    else alg.step(u, p, t, dt, un, ut, adaptive, nf);
                    accepted = true;
                    dts = dt;
                    dtnew = dt;
                    // error norm (A.4): scale in the working precision, ratio / square / sum in Float32 (EEst only
                    // steers the step size; keeps Float64 kernels free of IEEE double divisions).  Accept iff
                    // EEst^2 <= 1.  The division is the Newton reciprocal b2_rcp_nr (6.6e-6 accurate).
                    auto err_norm2 = [&]() -> float {
                        float acc = 0.0f;
#pragma unroll
                        for (int i = 0; i < B2_N; i++) {
                            const real sk = b2_fma(b2_max(b2_abs(u[i]), b2_abs(un[i])), B2_RTOL(a, i), B2_ATOL(a, i));
#ifdef B2_NORM_DIV   // experiments only (B200ENS_DEFINES): the IEEE division the Newton reciprocal replaced; NOT the oracle's bits
                            const float r = __fdiv_rn((float)ut[i], (float)sk);
#else
                            const float r = __fmul_rn((float)ut[i], b2_rcp_nr((float)sk));
#endif
                            acc = __fmaf_rn(r, r, acc);
                        }
#if B2_F64
                        // The Float32 ratio overflows to inf/inf = NaN when a wildly unstable attempt produces |u| beyond
                        // the Float32 range although the Float64 ratio is an ordinary number (Robertson, Rodas5P, first
                        // step of the automatic dt: |u_new| ~ 1e185, EEst ~ 1e3 -> must be REJECTED, not DtNaN).  Rare
                        // slow path: form the ratios in the working precision.  Genuine NaNs stay NaN.
                        if (acc != acc) {
                            acc = 0.0f;
#pragma unroll
                            for (int i = 0; i < B2_N; i++) {
                                const real sk = b2_fma(b2_max(b2_abs(u[i]), b2_abs(un[i])), B2_RTOL(a, i), B2_ATOL(a, i));
                                const float r = (float)(ut[i] / sk);
                                acc = __fmaf_rn(r, r, acc);
                            }
                        }
#endif
                        return __fmul_rn(acc, inv_n);
                    };
                    if (b2_own_control<Alg>::value) {
                        if constexpr (b2_own_control<Alg>::value) {
                            if (nfail) {
                                // retry with dt/2 (a fixed-step run cannot: Failure)
                                accepted = false;
                                if (adaptive) {
                                    nreject++;
                                    alg.newton_fail();
                                    dt = dt * (real)0.5;
                                } else {
                                    rc = B2_RC_FAILURE;
                                }
                            } else if (adaptive) {
                                const float EE2 = err_norm2();
                                if (EE2 != EE2) {
                                    rc = B2_RC_DTNAN;
                                    accepted = false;
                                } else if (!(EE2 <= 1.0f)) {
                                    accepted = false;
                                    nreject++;
                                    dt = dt * (real)alg.reject(EE2);
                                } else {
                                    dtnew = dt * (real)alg.accept(ctl.qmin, ctl.qmax);
                                }
                            } else {
                                alg.fixed_accept();
                                bool bad = false;
#pragma unroll
                                for (int i = 0; i < B2_N; i++) bad |= b2_isnan(un[i]);
                                if (bad) {
                                    rc = B2_RC_UNSTABLE;
                                    accepted = false;
                                }
                            }
                        }
                    } else if (adaptive) {
                        const float EE2 = err_norm2();
                        // PI controller (A.5), log domain, branch-free accept/reject (b2_control.cuh)
                        const B2Decision d = b2_pi_controller(EE2, lq, ctl);
                        const real dtq = dt * (real)d.qi;
                        const bool ok = d.ok, isn = d.isn;
                        accepted = ok;
                        if (isn) rc = B2_RC_DTNAN;
                        nreject += (!ok && !isn) ? 1 : 0;
                        lq = ok ? b2_ctl_lq_next(d, ctl) : lq;
                        dtnew = ok ? dtq : dtnew;
                        dt = (!ok && !isn) ? dtq : dt;
                    } else {
                        bool bad = false;
#pragma unroll
                        for (int i = 0; i < B2_N; i++) bad |= b2_isnan(un[i]);
                        if (bad) {
                            rc = B2_RC_UNSTABLE;
                            accepted = false;
                        }
                    }
                    if (accepted) {
                        naccept++;
                        tnew = t + dts;
                        if (b2_abs(tnew - tstop) < (real)100 * (real)B2_EPS * b2_max(b2_abs(tnew), b2_abs(tstop)))
                            tnew = tstop;
                        alg.accepted(un, p, tnew, nf);
#if B2_HAS_EVENT
                        // ---------------- ContinuousCallback (A.8): sign change over interp_points samples
                        // of the dense output, then bisection on theta keeping the LEFT side of the root
                        real w[B2_N];
                        alg.prepare_dense(u, p, tprev, dts, nf);
                        // The event search evaluates the dense output 10-25 times per step.  For the ERK steppers it
                        // uses the coefficient form of the interpolant, built once per step and only for the
                        // components the condition reads (B2_COND_MASK, from the symbolic condition): Horner in theta
                        // instead of re-weighting all stage vectors at every probe.
                        constexpr int PD = Alg::DEG > 0 ? Alg::DEG : 1;
                        real cc[B2_N][PD];
                        if (Alg::DEG > 0) {
#pragma unroll
                            for (int i = 0; i < B2_N; i++)
                                if ((B2_COND_MASK >> i) & 1u) alg.poly_coeffs(i, cc[i]);
                        }
                        auto fill_w = [&](real th) {
                            if (Alg::DEG > 0) {
#pragma unroll
                                for (int i = 0; i < B2_N; i++) {
                                    if ((B2_COND_MASK >> i) & 1u) {
                                        real pv = cc[i][PD - 1];
#pragma unroll
                                        for (int j = PD - 2; j >= 0; j--) pv = b2_fma(th, pv, cc[i][j]);
                                        w[i] = b2_fma(dts, th * pv, u[i]);
                                    } else {
                                        w[i] = u[i];  // not read by the condition
                                    }
                                }
                            } else {
                                alg.interp(u, un, th, dts, w);
                            }
                        };
#ifdef B2_NCOND
                        // ---- VectorContinuousCallback (qa.jl:124), search shared with the split kernel (b2_control.cuh)
                        fired = b2_vevent_search(
                            ip, just_fired, ev_last, [&](real* g) { b2_vcondition(g, u, p, tprev); },
                            [&](real th, real* g) {
                                fill_w(th);
                                b2_vcondition(g, w, p, b2_fma(th, dts, tprev));
                            },
                            [&](real* g) { b2_vcondition(g, un, p, tnew); }, th_end, ev_idx);
#else
                        // ---- scalar ContinuousCallback: sign change over interp_points samples, ITP root-find keeping the
                        // LEFT side of the root (b2_control.cuh)
                        fired = b2_event_search(
                            ip, just_fired, [&]() -> real { return b2_condition(u, p, tprev); },
                            [&](real th) -> real {
                                fill_w(th);
                                return b2_condition(w, p, b2_fma(th, dts, tprev));
                            },
                            [&]() -> real { return b2_condition(un, p, tnew); }, th_end, ev_idx);   // ev_idx: 1 = downcrossing
#endif   // B2_NCOND
                        if (fired) tnew = b2_fma(th_end, dts, tprev);
#endif
                    }
                }
            }
        }

        // ---------------- phase 2: saveat through the dense output (A.6), WARP-CONVERGENT.
        // Lanes cross their saveat points on different iterations; letting each lane run the
        // interpolant on its own serialises ~90 instructions per save at 1/32 lane efficiency
        // (measured with ncu: 60% of all issue slots, 17.7 active threads per instruction).
        // Instead the whole warp evaluates the interpolant whenever ANY lane needs a save, each
        // lane with its own theta, and only the lanes that need it store.
        for (;;) {
            const real tau = tau_next;
            const bool need = accepted && tau <= tnew;
            if (!__any_sync(B2_FULL, need)) break;
            const bool at_end = tau == tnew && !fired;  // the step lands exactly on the save point: store u_new
            if (need && !at_end) alg.prepare_dense(u, p, tprev, dts, nf);
            real w[B2_N];
            alg.interp(u, un, b2_theta(tau - tprev, dts), dts, w);
            if (need) {
                if (at_end) sink.put(si, un);
                else sink.put(si, w);
                si++;
                tau_next = si < n_save ? __ldg(gsave + si) : (real)__int_as_float(0x7f800000);
            }
        }

        // ---------------- phase 3: commit the accepted step (loopfooter!: FSAL hand-over, next dt)
        if (accepted) {
            t = tnew;
            while (ti < n_ts && __ldg(gts + ti) <= t) ti++;
#if B2_HAS_EVENT
            if (fired) {
                real w[B2_N];
                alg.interp(u, un, th_end, dts, w);
#ifdef B2_NCOND
                b2_vaffect(w, p, t, ev_idx);
                if (((B2_VTERM_MASK >> ev_idx) & 1u) || (a.event_terminate & 1)) rc = B2_RC_TERMINATED;   // this index's affect! called terminate!
#else
#if B2_HAS_AFFECT_NEG   // a downcrossing runs affect_neg! (event_terminate bit 2: it calls terminate!)
                if (ev_idx == 1) {
                    b2_affect_neg(w, p, t);
                    if (a.event_terminate & 4) rc = B2_RC_TERMINATED;
                } else {
                    b2_affect(w, p, t);
                    if (a.event_terminate & 1) rc = B2_RC_TERMINATED;
                }
#else
                b2_affect(w, p, t);
                if (a.event_terminate & 1) rc = B2_RC_TERMINATED;
#endif
#endif
#pragma unroll
                for (int i = 0; i < B2_N; i++) u[i] = w[i];
                nevents++;
                alg.start(u, p, t);
                nf++;
                just_fired = true;
            } else
#endif
            {
#pragma unroll
                for (int i = 0; i < B2_N; i++) u[i] = un[i];
                alg.advance();
#if B2_HAS_EVENT
                just_fired = false;
#endif
            }
#if B2_HAS_DEVENT
            // DiscreteCallback (test/core.jl:76-77): condition(u,t,integrator)::Bool tested on every accepted
            // state; affect! modifies u, so the FSAL derivative is re-evaluated
            if (rc == 0 && b2_dcondition(u, p, t)) {
                b2_daffect(u, p, t);
                nevents++;
                alg.start(u, p, t);
                nf++;
                if (a.event_terminate & 2) rc = B2_RC_TERMINATED;
            }
#endif
            if (adaptive) dt = b2_min(dtmax, dtnew);
            if (rc == 0 && !(t < t1)) rc = B2_RC_SUCCESS;
            if (EVERY && a.save_every) {   // the state after this accepted step (after callbacks), with its time
                if (si < n_save) {
                    sink.put(si, u);
                    reinterpret_cast<real*>(a.every_t)[idx * (long long)n_save + si] = t;
                }
                si++;
            }
        }

        // ---------------- phase 4: retire finished / failed lanes
        if (rc != 0) {
            if (EVERY && a.save_every) {   // unused slots: NaN in both arrays
                for (int k = si; k < n_save; k++) reinterpret_cast<real*>(a.every_t)[idx * (long long)n_save + k] = (real)__int_as_float(0x7fc00000);
                sink.fill(si < n_save ? si : n_save, n_save, (real)__int_as_float(0x7fc00000));
            } else if (rc == B2_RC_TERMINATED) {
                for (; si < n_save; si++) sink.put(si, u);
            } else if (rc != B2_RC_SUCCESS) {
                sink.fill(si, n_save, (real)__int_as_float(0x7fc00000));
            }
            a.retcode[idx] = rc;
            if (sink.moments() && rc != B2_RC_SUCCESS && rc != B2_RC_TERMINATED) atomicAdd(a.mom_fail, 1ull);
            if (a.stats) {
                B2Stats s;
                s.naccept = naccept;
                s.nreject = nreject;
                s.nf = nf;
                s.nevents = nevents;
                a.stats[idx] = s;
            }
            active = false;
            dirty = stride != 0;
        }
    }
}
