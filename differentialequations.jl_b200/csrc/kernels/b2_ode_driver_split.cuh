// b2_ode_driver_split.cuh -- the persistent ODE ensemble kernel for LARGE systems: one trajectory per LANE of a
// 4-warp CTA, the state and stage vectors split by component over the four warps (b2_split.cuh).
//
// Same semantics, same per-component expression trees and therefore the same bits as b2_ode_driver.cuh (SURVEY.md
// A.1-A.8: loopheader!/perform_step!/loopfooter!/savevalues!/ContinuousCallback); what changes is who holds what:
//   * lane l of every warp of the CTA works on the SAME trajectory slot l; warp g holds components [g*NL,(g+1)*NL);
//   * t, dt, controller memory, saveat index, counters, return code are REPLICATED in the four warps and evolve
//     identically (deterministic arithmetic on identical inputs), so every branch that guards a barrier is CTA-uniform;
//   * anything that contains a barrier (RHS evaluations inside the stepper, the error-norm gather, the event-data
//     gather) is executed by all 128 threads unconditionally; per-lane predicates only select what is COMMITTED;
//   * warp 0 alone talks to the work queue and writes retcodes / stats.
// Supported: Tsit5 / Vern7, adaptive or fixed dt, automatic initial dt, saveat through the dense output or as tstops,
// scalar ContinuousCallback, expected-work ordering, per-component tolerances.  Not here (the host keeps the
// one-thread kernel for these): DiscreteCallback, VectorContinuousCallback, save_everystep, output staging.
#pragma once
#include "b2_common.cuh"
#include "b2_control.cuh"
#include "b2_split.cuh"

// owned block of a full-length register array, by warp role (compile-time indices in every case: no local memory)
#define B2_OWNED(dst, full, g_)                                                                                   \
    switch (g_) {                                                                                                 \
    case 0: _Pragma("unroll") for (int j = 0; j < B2_NL; j++) dst[j] = (0 * B2_NL + j < B2_N) ? full[(0 * B2_NL + j < B2_N) ? 0 * B2_NL + j : 0] : (real)0; break; \
    case 1: _Pragma("unroll") for (int j = 0; j < B2_NL; j++) dst[j] = (1 * B2_NL + j < B2_N) ? full[(1 * B2_NL + j < B2_N) ? 1 * B2_NL + j : 0] : (real)0; break; \
    case 2: _Pragma("unroll") for (int j = 0; j < B2_NL; j++) dst[j] = (2 * B2_NL + j < B2_N) ? full[(2 * B2_NL + j < B2_N) ? 2 * B2_NL + j : 0] : (real)0; break; \
    default: _Pragma("unroll") for (int j = 0; j < B2_NL; j++) dst[j] = (3 * B2_NL + j < B2_N) ? full[(3 * B2_NL + j < B2_N) ? 3 * B2_NL + j : 0] : (real)0; break; \
    }

#if B2_HAS_EVENT
#define B2_EV_M (b2_popc_c(B2_COND_MASK & (B2_N >= 32 ? 0xffffffffu : ((1u << B2_N) - 1u))))
#else
#define B2_EV_M 1
#endif

// ADAPT / TSTOPS / AUTODT: compile-time specialisation as in b2_ode_driver.cuh (-1 / 1 = decided at run time)
template <class Alg, int ADAPT = -1, int TSTOPS = -1, int AUTODT = 1>
__device__ __forceinline__ void b2_ode_driver_split(const B2Args& a) {
    __shared__ __align__(16) real s_xchg[2 * B2_NP * 32];
    __shared__ long long s_my[32];
    __shared__ __align__(16) real s_par[B2_NPA * 32];   // the trajectories' parameters, read by the out-of-line RHS
    __shared__ int s_flag[4];
    const unsigned lane = threadIdx.x & 31u;
    // Warp role = which quarter of the components this warp owns (role 0 also talks to the work queue and runs the event
    // search).  Warp w of every CTA sits on scheduler w % 4, so with role = warp index one scheduler of the SM would run
    // the heaviest role of EVERY resident CTA while the other three wait at the barriers; rotating the assignment with the
    // CTA index spreads the roles over the schedulers (B2_ROLE_ROTATE=0: experiments).
#ifndef B2_ROLE_ROTATE
#define B2_ROLE_ROTATE 1
#endif
    const int g = B2_ROLE_ROTATE ? (int)(((threadIdx.x >> 5) + blockIdx.x) & 3u) : (int)(threadIdx.x >> 5);
    const int c0 = g * B2_NL;
    real* const gout = reinterpret_cast<real*>(a.out_u);
    const real* const gu0 = reinterpret_cast<const real*>(a.u0);
    const real* const gp = reinterpret_cast<const real*>(a.p);
    const real* const gsave = reinterpret_cast<const real*>(a.saveat);
    const int n_save = a.n_save;
    const long long out_per_traj = (long long)n_save * B2_NOUT;
    // fused ensemble moments (B2Args.mom_sum): generic entry only; a saved value is ADDED to the per-(save point,
    // component) sums instead of being stored
    // (the pointers are read from the kernel-argument constant bank where they are used, not held in registers)
#define msum (ADAPT < 0 ? a.mom_sum : (double*)nullptr)
    auto put_at = [&](long long obase_, int si_, int pos, real v) {   // pos: position inside the output row
        if (msum) {
            const double x = (double)v;
            atomicAdd(a.mom_sum + (long long)si_ * B2_NOUT + pos, x);
            atomicAdd(a.mom_sq + (long long)si_ * B2_NOUT + pos, x * x);
        } else {
            gout[obase_ + (long long)si_ * B2_NOUT + pos] = v;
        }
    };
    // component c0 + j of this warp's block (j is a compile-time index at every call site)
    auto put_owned = [&](long long obase_, int si_, int j, real v) {
#if B2_HAS_SAVE_IDXS
#pragma unroll
        for (int k = 0; k < B2_NOUT; k++)
            if (B2_SIDX(k) == c0 + j) put_at(obase_, si_, k, v);   // save_idxs: where (and whether) the component is saved
#else
        put_at(obase_, si_, c0 + j, v);
#endif
    };

    const real t0 = B2_ARG(a, t0), t1 = B2_ARG(a, t1), dt_user = B2_ARG(a, dt);
    const B2Ctl ctl = b2_ctl_init(a);   // the PI controller works in Float32 (b2_control.cuh)
    const float lqinit = ctl.lqinit;
    const float inv_n = __fdiv_rn(1.0f, (float)B2_N);
    const real dtmax = B2_ARG(a, dtmax), dtmin = B2_ARG(a, dtmin);
    const bool adaptive = ADAPT < 0 ? (a.adaptive != 0) : (ADAPT != 0);
    const bool save_tstops = TSTOPS < 0 ? (a.save_tstops != 0) : (TSTOPS != 0);
    const int n_ts = ADAPT < 0 ? a.n_tstops : 0;   // solve(...; tstops): generic entry only
    const real* const gts = reinterpret_cast<const real*>(a.tstops);
    const real INF = (real)__int_as_float(0x7f800000);
    // tolerances of the owned components: read from the kernel-argument constant bank with a warp-uniform index where
    // they are used (not held in 4 * NL registers); padded components take the last real one
#define B2_TOLIDX(j) ((c0 + (j) < B2_N) ? c0 + (j) : B2_N - 1)
#define atol_(j) B2_ATOL(a, B2_TOLIDX(j))
#define rtol_(j) B2_RTOL(a, B2_TOLIDX(j))

    Alg alg;
    alg.bind(nullptr);
    alg.xc.base = s_xchg;
    alg.xc.lane = (int)lane;
    alg.xc.phase = 0;
    alg.xc.g = g;
    alg.xc.pcol = s_par + lane;
    real u[B2_NL], p[B2_NPA];   // p: placeholder for the stepper interface (the split RHS reads shared memory); never loaded
#pragma unroll
    for (int j = 0; j < B2_NL; j++) u[j] = 0;
#pragma unroll
    for (int j = 0; j < B2_NPA; j++) p[j] = 0;
    // the parameters of this lane's trajectory, fetched from shared memory where a callback function needs them
    auto load_params = [&](real (&pe)[B2_NPA]) {
#pragma unroll
        for (int i = 0; i < B2_NPA; i++) pe[i] = i < B2_NPARAM ? s_par[i * 32 + lane] : (real)0;
    };
    {
        real z[B2_NL];
#pragma unroll
        for (int j = 0; j < B2_NL; j++) z[j] = 0;
        alg.set_k1(z);
    }
    real t = t0, dt = dt_user, tau_next = INF;
    float lq = lqinit;
    long long idx = -1, obase = 0;
    int iter = 0;
    const int maxit = a.maxiters > 0x7fffffffLL ? 0x7fffffff : (int)a.maxiters;
    int si = 0, naccept = 0, nreject = 0, nf = 0, nevents = 0;
    int ti = 0;   // next user tstop of this lane (replicated in the four warps)
    bool active = false, exhausted = false;
#if B2_HAS_EVENT
    __shared__ __align__(16) real s_ev[B2_EV_M * (Alg::DEG + 2) * 32];
    __shared__ real s_evres[32];
    __shared__ int s_evidx[32];
    bool just_fired = false;
    int ev_last = 0;   // VectorContinuousCallback: index of the function that fired the last event
    const int ip = a.interp_points;
#endif

    for (;;) {
        // ---------------- phase 0: retire / refill (CTA-uniform: the four warps hold identical `active` masks)
        const unsigned idle = __ballot_sync(B2_FULL, !active);
        if (idle) {
            const bool all_idle = idle == B2_FULL;
            if (exhausted ? all_idle : (__popc(idle) >= a.refill_threshold || all_idle)) {
                if (!exhausted) {
                    if (g == 0) {
                        bool ex = false;
                        const long long mine = b2_fetch(idle, a.work_counter, a.N, lane, ex);
                        s_my[lane] = mine;
                        if (lane == 0) s_flag[0] = ex ? 1 : 0;
                    }
                    __syncthreads();
                    const long long my = s_my[lane];
                    exhausted = s_flag[0] != 0;
                    const bool fresh = !active && my >= 0;
                    if (fresh) {
                        idx = a.perm ? (long long)__ldg(a.perm + my) : my;
                        obase = idx * out_per_traj;
#pragma unroll
                        for (int j = 0; j < B2_NL; j++) u[j] = (c0 + j < B2_N) ? gu0[idx * B2_N + c0 + j] : (real)0;
                        // the trajectory's parameters live in shared memory only (read by the out-of-line RHS and, on demand,
                        // by the callback functions): keeping a register copy would cost 2 * n_param registers for the
                        // whole loop in a kernel that sits at its register budget
                        if (g == 0) {
#pragma unroll
                            for (int i = 0; i < B2_NPARAM; i++) s_par[i * 32 + lane] = gp[idx * B2_NPARAM + i];
                        }
                        t = t0;
                        dt = dt_user;
                        lq = lqinit;
                        iter = 0;
                        si = 0;
                        naccept = nreject = nevents = 0;
                        ti = 0;
                        while (ti < n_ts && __ldg(gts + ti) <= t0) ti++;
                        // the first saved value is u0 itself (test/core.jl:34)
                        while (si < n_save && __ldg(gsave + si) <= t0) {
#pragma unroll
                            for (int j = 0; j < B2_NL; j++)
                                if (c0 + j < B2_N) put_owned(obase, si, j, u[j]);
                            si++;
                        }
                        tau_next = si < n_save ? __ldg(gsave + si) : INF;
                    }
                    // f(u0, t0): evaluated by the whole CTA (barrier inside), committed by the fresh lanes only
                    real f0[B2_NL];
                    alg.rhs(f0, u, p, t);
                    if (fresh) {
                        alg.set_k1(f0);
                        nf = 1;
                        active = true;
#if B2_HAS_EVENT
                        just_fired = false;
#endif
                    }
                    if (AUTODT && adaptive && !(dt_user > (real)0)) {
                        // automatic initial step (same formula and summation order as b2_ode_driver.cuh);
                        // whole CTA, committed by the fresh lanes
                        real r0[B2_NL], r1[B2_NL], r2[B2_NL], u1[B2_NL], f1[B2_NL];
#pragma unroll
                        for (int j = 0; j < B2_NL; j++) {
                            const real sk = b2_fma(b2_abs(u[j]), rtol_(j), atol_(j));
                            r0[j] = u[j] / sk;
                            r1[j] = f0[j] / sk;
                        }
                        real a0 = 0, a1 = 0, a2 = 0;
                        {
                            const real* b0 = b2_split_publish(alg.xc, r0);
#pragma unroll
                            for (int i = 0; i < B2_N; i++) a0 = b2_fma(b0[i * 32], b0[i * 32], a0);
                            const real* b1 = b2_split_publish(alg.xc, r1);
#pragma unroll
                            for (int i = 0; i < B2_N; i++) a1 = b2_fma(b1[i * 32], b1[i * 32], a1);
                        }
                        const real d0 = b2_sqrt(a0 / (real)B2_N), d1 = b2_sqrt(a1 / (real)B2_N);
                        real dt0 = (d0 < (real)1e-5 || d1 < (real)1e-5) ? (real)1e-6 : (real)0.01 * (d0 / d1);
                        dt0 = b2_min(dt0, dtmax);
#pragma unroll
                        for (int j = 0; j < B2_NL; j++) u1[j] = b2_fma(dt0, f0[j], u[j]);
                        alg.rhs(f1, u1, p, t0 + dt0);
#pragma unroll
                        for (int j = 0; j < B2_NL; j++) {
                            const real sk = b2_fma(b2_abs(u[j]), rtol_(j), atol_(j));
                            r2[j] = (f1[j] - f0[j]) / sk;
                        }
                        {
                            const real* b2p = b2_split_publish(alg.xc, r2);
#pragma unroll
                            for (int i = 0; i < B2_N; i++) a2 = b2_fma(b2p[i * 32], b2p[i * 32], a2);
                        }
                        const real d2 = b2_sqrt(a2 / (real)B2_N) / dt0;
                        const real dmx = b2_max(d1, d2);
                        real dt1;
                        if (dmx <= (real)1e-15) dt1 = b2_max((real)1e-6, dt0 * (real)1e-3);
                        else dt1 = (real)b2_fastexp2(__fmul_rn(-__fadd_rn(6.6438562f, b2_fastlog2((float)dmx)),
                                                              __fdiv_rn(1.0f, (float)Alg::ORDER)));
                        if (fresh) {
                            dt = b2_min(b2_min((real)100 * dt0, dt1), dtmax);
                            nf++;
                        }
                    }
                }
                if (__ballot_sync(B2_FULL, active) == 0u) break;
            }
        }

        // ---------------- phase 1: pre-step control (replicated), then ONE step attempt by the whole CTA
        int rc = 0;
        bool accepted = false, fired = false;
        const real tprev = t;
        real tnew = t, dts = dt, dtnew = dt;
        real un[B2_NL], ut[B2_NL];
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
        {
            int nfd = 0;
            alg.step(u, p, t, dt, un, ut, adaptive, nfd);
            if (do_step) nf += nfd;
        }
        dts = dt;
        dtnew = dt;
        if (adaptive) {
            // error norm (A.4): every warp publishes the ratios of its components and sums ALL n squares in index order
            float r[B2_NL];
#pragma unroll
            for (int j = 0; j < B2_NL; j++) {
                const real sk = b2_fma(b2_max(b2_abs(u[j]), b2_abs(un[j])), rtol_(j), atol_(j));
                r[j] = __fmul_rn((float)ut[j], b2_rcp_nr((float)sk));
            }
            const float* rb = b2_split_publish(alg.xc, r);
            float acc = 0.0f;
#pragma unroll
            for (int i = 0; i < B2_N; i++) {
                const float ri = rb[i * 32];
                acc = __fmaf_rn(ri, ri, acc);
            }
#if B2_F64
            // Float32 overflow of a finite Float64 ratio (see b2_ode_driver.cuh).  The slow path contains an exchange, so
            // the whole CTA takes it when any lane needs it (every warp holds the same acc per lane).
            if (__syncthreads_or(acc != acc)) {
                float r2[B2_NL];
#pragma unroll
                for (int j = 0; j < B2_NL; j++) {
                    const real sk = b2_fma(b2_max(b2_abs(u[j]), b2_abs(un[j])), rtol_(j), atol_(j));
                    r2[j] = (float)(ut[j] / sk);
                }
                const float* rb2 = b2_split_publish(alg.xc, r2);
                float acc2 = 0.0f;
#pragma unroll
                for (int i = 0; i < B2_N; i++) {
                    const float ri = rb2[i * 32];
                    acc2 = __fmaf_rn(ri, ri, acc2);
                }
                if (acc != acc) acc = acc2;
            }
#endif
            const float EE2 = __fmul_rn(acc, inv_n);
            const B2Decision d = b2_pi_controller(EE2, lq, ctl);   // shared with the one-thread driver
            const real dtq = dt * (real)d.qi;
            const bool ok = d.ok, isn = d.isn;
            if (do_step) {
                accepted = ok;
                if (isn) rc = B2_RC_DTNAN;
                nreject += (!ok && !isn) ? 1 : 0;
                lq = ok ? b2_ctl_lq_next(d, ctl) : lq;
                dtnew = ok ? dtq : dtnew;
                dt = (!ok && !isn) ? dtq : dt;
            }
        } else if (do_step) {
            // fixed step: NaN anywhere in the new state -> Unstable.  Each warp sees its own block only, so publish a flag.
            accepted = true;
        }
        if (!adaptive) {
            float bad[B2_NL];
#pragma unroll
            for (int j = 0; j < B2_NL; j++) bad[j] = b2_isnan(un[j]) ? 1.0f : 0.0f;
            const float* bb = b2_split_publish(alg.xc, bad);
            float any_bad = 0.0f;
#pragma unroll
            for (int i = 0; i < B2_N; i++) any_bad += bb[i * 32];
            if (do_step && any_bad != 0.0f) {
                rc = B2_RC_UNSTABLE;
                accepted = false;
            }
        }
        if (accepted) {
            naccept++;
            tnew = t + dts;
            if (b2_abs(tnew - tstop) < (real)100 * (real)B2_EPS * b2_max(b2_abs(tnew), b2_abs(tstop))) tnew = tstop;
        }
        {
            // Vern7: k11 = f(u_new) (dense-output stage and next step's k1); Tsit5: nothing
            int nfa = 0;
            alg.accepted(un, p, tnew, nfa);
            if (accepted) nf += nfa;
        }
        bool dense_ready = false;   // CTA-uniform: the lazy dense-output stages of this step exist
        bool lane_extra = false;    // this lane has paid for them in its nf statistic
        int nfe = 0;
#if B2_HAS_EVENT
        {
            // ---------------- ContinuousCallback (A.8).  The search itself is replicated per lane (no barriers): it
            // runs on the coefficient form of the interpolant of the components the condition reads, which their
            // owner warps publish once per step together with u and u_new of those components.
            alg.reset_dense();
            alg.prepare_dense(u, p, tprev, dts, nfe);
            dense_ready = true;
            if (accepted) {
                nf += nfe;
                lane_extra = true;
            }
            constexpr int PD = Alg::DEG;
            constexpr int EW = PD + 2;
            {
                int m = 0;
#pragma unroll
                for (int c = 0; c < B2_N; c++) {
                    if ((B2_COND_MASK >> c) & 1u) {
                        if (g == c / B2_NL) {   // warp-uniform
                            real cc[PD];
                            alg.poly_coeffs(c % B2_NL, cc);
                            real* e = s_ev + (size_t)m * EW * 32 + lane;
                            e[0] = u[c % B2_NL];
                            e[32] = un[c % B2_NL];
#pragma unroll
                            for (int k = 0; k < PD; k++) e[(2 + k) * 32] = cc[k];
                        }
                        m++;
                    }
                }
            }
            __syncthreads();
            real eu[B2_EV_M], eun[B2_EV_M], ecc[B2_EV_M][PD];
#pragma unroll
            for (int m = 0; m < B2_EV_M; m++) {
                const real* e = s_ev + (size_t)m * EW * 32 + lane;
                eu[m] = e[0];
                eun[m] = e[32];
#pragma unroll
                for (int k = 0; k < PD; k++) ecc[m][k] = e[(2 + k) * 32];
            }
            real w[B2_N], pe[B2_NPA];
            load_params(pe);
            auto scatter = [&](const real (&v)[B2_EV_M]) {
                int m = 0;
#pragma unroll
                for (int c = 0; c < B2_N; c++) {
                    w[c] = 0;
                    if ((B2_COND_MASK >> c) & 1u) {
                        w[c] = v[m];
                        m++;
                    }
                }
            };
#ifndef B2_NCOND
            auto cond_at = [&](real th) -> real {
                real v[B2_EV_M];
#pragma unroll
                for (int m = 0; m < B2_EV_M; m++) {
                    real pv = ecc[m][PD - 1];
#pragma unroll
                    for (int j = PD - 2; j >= 0; j--) pv = b2_fma(th, pv, ecc[m][j]);
                    v[m] = b2_fma(dts, th * pv, eu[m]);
                }
                scatter(v);
                return b2_condition(w, pe, b2_fma(th, dts, tprev));
            };
#endif
            // The search is sequential per lane and only ~1 lane in 10 fires on a given step: it would cost every warp the
            // full ~20 root-find iterations at 3/32 lane efficiency (ncu: 27 % of all issued instructions when
            // replicated).  Warp 0 searches alone and publishes theta of the event (or -1) for the other three.
            if (g == 0 && accepted) {
#ifdef B2_NCOND
                // VectorContinuousCallback (qa.jl:124): the search of b2_control.cuh, shared with the one-thread driver
                auto vcond_at = [&](real th, real* gv) {
                    real v[B2_EV_M];
#pragma unroll
                    for (int m = 0; m < B2_EV_M; m++) {
                        real pv = ecc[m][PD - 1];
#pragma unroll
                        for (int j = PD - 2; j >= 0; j--) pv = b2_fma(th, pv, ecc[m][j]);
                        v[m] = b2_fma(dts, th * pv, eu[m]);
                    }
                    scatter(v);
                    b2_vcondition(gv, w, pe, b2_fma(th, dts, tprev));
                };
                fired = b2_vevent_search(
                    ip, just_fired, ev_last,
                    [&](real* gv) {
                        scatter(eu);
                        b2_vcondition(gv, w, pe, tprev);
                    },
                    vcond_at,
                    [&](real* gv) {
                        scatter(eun);
                        b2_vcondition(gv, w, pe, tnew);
                    },
                    th_end, ev_idx);
#else
                fired = b2_event_search(
                    ip, just_fired,
                    [&]() -> real {
                        scatter(eu);
                        return b2_condition(w, pe, tprev);
                    },
                    cond_at,
                    [&]() -> real {
                        scatter(eun);
                        return b2_condition(w, pe, tnew);
                    },
                    th_end, ev_idx);   // ev_idx: 1 = downcrossing
#endif
            }
            if (g == 0) s_evidx[lane] = ev_idx;
            if (g == 0) s_evres[lane] = fired ? th_end : (real)-1;
            __syncthreads();
            {
                const real th_pub = s_evres[lane];
                fired = accepted && th_pub >= (real)0;
                if (fired) {
                    th_end = th_pub;
                    ev_idx = s_evidx[lane];
                    ev_last = ev_idx;
                    tnew = b2_fma(th_end, dts, tprev);
                }
            }
        }
#endif

        // ---------------- phase 2: saveat through the dense output (A.6); the loop condition is CTA-uniform
        for (;;) {
            const real tau = tau_next;
            const bool need = accepted && tau <= tnew;
            if (!__any_sync(B2_FULL, need)) break;
            const bool at_end = tau == tnew && !fired;  // the step lands exactly on the save point: store u_new
            if (!dense_ready) {   // lazy dense-output stages, whole CTA; a lane pays for them when it first interpolates
                alg.reset_dense();
                alg.prepare_dense(u, p, tprev, dts, nfe);
                dense_ready = true;
            }
            if (need && !at_end && !lane_extra) {
                nf += nfe;
                lane_extra = true;
            }
            real w[B2_NL];
            alg.interp(u, un, b2_theta(tau - tprev, dts), dts, w);
            if (need) {
#pragma unroll
                for (int j = 0; j < B2_NL; j++)
                    if (c0 + j < B2_N) put_owned(obase, si, j, at_end ? un[j] : w[j]);
                si++;
                tau_next = si < n_save ? __ldg(gsave + si) : INF;
            }
        }

        // ---------------- phase 3: commit the accepted step
#if B2_HAS_EVENT
        if (__any_sync(B2_FULL, fired)) {   // CTA-uniform
            // state at the event time: every warp interpolates its block, the full vector is gathered, affect! is
            // replicated, f(u_after) is evaluated by the whole CTA and committed by the lanes that fired
            real wl[B2_NL];
            alg.interp(u, un, th_end, dts, wl);
            const real* fb = b2_split_publish(alg.xc, wl);
            real W[B2_N];
#pragma unroll
            for (int i = 0; i < B2_N; i++) W[i] = fb[i * 32];
            real pa[B2_NPA];
            load_params(pa);
#ifdef B2_NCOND
            b2_vaffect(W, pa, tnew, ev_idx);   // per lane: the index of the function that fired
#elif B2_HAS_AFFECT_NEG
            if (ev_idx == 1) b2_affect_neg(W, pa, tnew);   // per lane: a downcrossing runs affect_neg!
            else b2_affect(W, pa, tnew);
#else
            b2_affect(W, pa, tnew);
#endif
            real wn[B2_NL], fnew[B2_NL];
            B2_OWNED(wn, W, g)
            alg.rhs(fnew, wn, p, tnew);
            if (fired) {
#pragma unroll
                for (int j = 0; j < B2_NL; j++) u[j] = wn[j];
                alg.set_k1(fnew);
                nevents++;
                nf++;
                just_fired = true;
                if (a.event_terminate & ((B2_HAS_AFFECT_NEG && ev_idx == 1) ? 4 : 1)) rc = B2_RC_TERMINATED;
#ifdef B2_NCOND
                if ((B2_VTERM_MASK >> ev_idx) & 1u) rc = B2_RC_TERMINATED;   // this index's affect! called terminate!
#endif
            }
        }
#endif
        if (accepted) {
            t = tnew;
            while (ti < n_ts && __ldg(gts + ti) <= t) ti++;
            if (!fired) {
#pragma unroll
                for (int j = 0; j < B2_NL; j++) u[j] = un[j];
                alg.advance();
#if B2_HAS_EVENT
                just_fired = false;
#endif
            }
            if (adaptive) dt = b2_min(dtmax, dtnew);
            if (rc == 0 && !(t < t1)) rc = B2_RC_SUCCESS;
        }

        // ---------------- phase 4: retire finished / failed lanes
        if (rc != 0) {
            if (rc != B2_RC_SUCCESS && !(msum && rc != B2_RC_TERMINATED)) {
                for (; si < n_save; si++) {
#pragma unroll
                    for (int j = 0; j < B2_NL; j++)
                        if (c0 + j < B2_N) put_owned(obase, si, j, rc == B2_RC_TERMINATED ? u[j] : (real)__int_as_float(0x7fc00000));
                }
            }
            if (g == 0) {
                a.retcode[idx] = rc;
                if (msum && rc != B2_RC_SUCCESS && rc != B2_RC_TERMINATED) atomicAdd(a.mom_fail, 1ull);
                if (a.stats) {
                    B2Stats s;
                    s.naccept = naccept;
                    s.nreject = nreject;
                    s.nf = nf;
                    s.nevents = nevents;
                    a.stats[idx] = s;
                }
            }
            active = false;
        }
    }
}
#undef msum
#undef atol_
#undef rtol_
#undef B2_TOLIDX
