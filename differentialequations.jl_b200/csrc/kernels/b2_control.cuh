// b2_control.cuh -- the pieces of the adaptive integrator loop that every driver shares (one-thread, split, adaptive
// SDE): the PI step-size controller, and the ContinuousCallback / VectorContinuousCallback event search with its ITP
// root-find.  One copy, so a semantic fix lands once (+ once in the oracle, oracle/oracle_impl.inc).
// Upstream names: OrdinaryDiffEq stepsize_controller! / step_accept_controller! / step_reject_controller! (PIController)
// and DiffEqBase find_callback_time / find_root (ContinuousCallback, /root/reference/test/qa/qa.jl:26,124;
// call sites test/core.jl:69-78); semantics SURVEY.md A.5 / A.8.
#pragma once
#include "b2_common.cuh"

// ---- PI controller (A.5) in the log domain, Float32, branch-free, division-free.
//   l = log2(EEst) = 0.5*log2(EEst^2), lq = log2(qold)
//   accept (EEst <= 1): dt_next = dt * clamp(gamma * 2^(beta2*lq - beta1*l), qmin, qmax)   [= dt / q, q = EEst^b1/qold^b2/gamma]
//   reject            : dt_retry = dt * max(gamma * 2^(-beta1*l), qmin)                      [= dt / min(1/qmin, q11/gamma)]
//   EEst == 0         : dt_next = dt * qmax
// Rejections are rare per lane (3 %) but some lane of a warp rejects in 37 % of the iterations, so both cases go
// through the same instructions and only the exponent argument and the upper clamp are selected (profiles/README.md).
struct B2Ctl {
    float qmin, qmax, gam, beta1, beta2, lqinit;
};
struct B2Decision {
    bool ok, isn, zero;   // accept, EEst is NaN (upstream: NaN dt -> ReturnCode.DtNaN), EEst == 0
    float l;              // log2(EEst)
    float qi;             // dt multiplier (1/q)
};
__device__ __forceinline__ B2Ctl b2_ctl_init(const B2Args& a) {
    B2Ctl c;
    c.qmin = a.f_qmin;
    c.qmax = a.f_qmax;
    c.gam = a.f_gamma;
    c.beta1 = a.f_beta1;
    c.beta2 = a.f_beta2;
    c.lqinit = b2_fastlog2(a.f_qoldinit);
    return c;
}
__device__ __forceinline__ B2Decision b2_pi_controller(const float EE2, const float lq, const B2Ctl& c) {
    B2Decision d;
    d.isn = EE2 != EE2;
    d.ok = EE2 <= 1.0f;
    d.zero = EE2 == 0.0f;
    d.l = __fmul_rn(0.5f, b2_fastlog2(EE2));
    const float nbl = __fmul_rn(-c.beta1, d.l);
    float qi = b2_fastexp2(d.ok ? __fmaf_rn(c.beta2, lq, nbl) : nbl);
    qi = fmaxf(c.qmin, __fmul_rn(qi, c.gam));
    qi = d.ok ? fminf(c.qmax, qi) : qi;
    d.qi = d.zero ? c.qmax : qi;
    return d;
}
// controller memory after an accepted step: qold = max(EEst, qoldinit) in the log domain
__device__ __forceinline__ float b2_ctl_lq_next(const B2Decision& d, const B2Ctl& c) {
    return fmaxf(d.zero ? c.lqinit : d.l, c.lqinit);
}

// ---- ITP bracketing root-find on theta in (lo, hi) (kappa1 = 0.2/(b-a), kappa2 = 2, n0 = 1) down to a bracket of
// 4 eps; keeps and returns the LEFT end (LeftRootFind: the condition has not changed sign yet at the event time).
// cond_at(theta) evaluates the event function on the dense output; gprev is its sign reference at the step start.
// The truncation offset has a FLOOR of 1.5 eps: plain ITP reaches a bracket of ~1e-10 in ~7 evaluations and then
// degenerates into bisection (its offset kappa1*w^2 falls below an ulp, the regula-falsi point lands ON the root and
// only one side of the bracket moves): 17 evaluations on average, 25 for the slowest of four lanes.  With the floor the
// probe sits 1.5 eps beside the estimated root, both sides close and the search ends after 9-10 evaluations -- the
// root-find was 27 % of all instructions of config 5 at 3/32 lanes, on every CTA's critical path (profiles/README.md).
// Still worst case bisection + 1 evaluations.  Same evaluation sequence as the oracle.
template <class CondAt>
__device__ __forceinline__ real b2_itp_left(real lo, real hi, real glo, real ghi, const real gprev, CondAt&& cond_at) {
    const real eps = (real)2 * (real)B2_EPS;
    const real k1 = (real)0.2 / (hi - lo);
    real pw = b2_itp_pw(hi - lo);   // eps * 2^(halvings + 1), closed form
    for (int it = 0; it < 100 && hi - lo > (real)2 * eps; it++) {
        const real xh = (real)0.5 * (lo + hi);
        const real r = pw - (real)0.5 * (hi - lo);
        pw *= (real)0.5;
        const real delta = b2_max(k1 * (hi - lo) * (hi - lo), (real)0.75 * eps);
        const real xf = (ghi * lo - glo * hi) / (ghi - glo);
        const real sg = (xh - xf) >= 0 ? (real)1 : (real)-1;
        const real xt = (delta <= b2_abs(xh - xf)) ? xf + sg * delta : xh;
        real x = (b2_abs(xt - xh) <= r) ? xt : xh - sg * r;
        if (!(x > lo && x < hi)) x = xh;
        if (!(x > lo && x < hi)) break;
        const real g = cond_at(x);
        if ((gprev < 0 && g >= 0) || (gprev > 0 && g <= 0)) {
            hi = x;
            ghi = g;
        } else {
            lo = x;
            glo = g;
        }
    }
    return lo;
}
// a sign change relative to the sign at the step start, in an enabled direction (B2_EVENT_DIR, b2_common.cuh)
__device__ __forceinline__ bool b2_sign_change(real gprev, real g) {
    return (B2_EVENT_DIR >= 0 && gprev < 0 && g >= 0) || (B2_EVENT_DIR <= 0 && gprev > 0 && g <= 0);
}

// ---- scalar ContinuousCallback (A.8): sign change over interp_points samples of the dense output, then the ITP
// root-find.  cond_start(): the event function at (u, tprev); cond_at(theta): on the interpolant; cond_end(): at
// (u_new, tnew).  After an event at the end of the previous step the reference sign is taken at theta = 0.01
// (repeat_nudge).  Returns true and the event's theta when the event fires inside this step.
// `down` receives 1 when the event is a downcrossing (condition positive at the step start), else 0.
template <class CondStart, class CondAt, class CondEnd>
__device__ __forceinline__ bool b2_event_search(const int ip, const bool just_fired, CondStart&& cond_start, CondAt&& cond_at,
                                                CondEnd&& cond_end, real& th_end, int& down) {
    real gprev, lo = 0, hi = 0, glo, ghi = 0;
    bool fired = false;
    if (just_fired) {
        gprev = cond_at((real)0.01);
        lo = (real)0.01;
    } else {
        gprev = cond_start();
    }
    glo = gprev;
#ifndef B2_EVENT_BATCH
#define B2_EVENT_BATCH 1
#endif
    // The samples are consumed in order, up to the first sign change.  B2_EVENT_BATCH > 1 EVALUATES them that many at a
    // time (independent divisions and Horner chains instead of ~200 cycles of latency per sample; samples past the
    // first sign change are computed and dropped, the event functions are pure, so every returned bit is unchanged).
    // Measured on config 5 (split kernel, the other three warps wait for this search): 35.3 ms (1) / 36.4 (5) / 39.0 (10)
    // per 200k trajectories -- the extra live values spill in a kernel that sits at its register budget -- so 1 stays.
    for (int m0 = 1; m0 <= ip && !fired; m0 += B2_EVENT_BATCH) {
        real thb[B2_EVENT_BATCH], gb[B2_EVENT_BATCH];
#pragma unroll
        for (int c = 0; c < B2_EVENT_BATCH; c++) {
            const int mm = m0 + c;
            thb[c] = (mm >= ip) ? (real)1 : (real)mm / (real)ip;
        }
#pragma unroll
        for (int c = 0; c < B2_EVENT_BATCH; c++) {
            const int mm = m0 + c;
            gb[c] = (mm < ip) ? cond_at(thb[c]) : (real)0;
        }
#pragma unroll
        for (int c = 0; c < B2_EVENT_BATCH; c++) {
            const int mm = m0 + c;
            if (mm <= ip && !fired) {
                const real g = (mm == ip) ? cond_end() : gb[c];
                if (b2_sign_change(gprev, g)) {
                    fired = true;
                    hi = thb[c];
                    ghi = g;
                } else {
                    lo = thb[c];
                    glo = g;
                }
            }
        }
    }
    if (fired) {
        th_end = b2_itp_left(lo, hi, glo, ghi, gprev, cond_at);
        down = gprev < 0 ? 0 : 1;
    }
    return fired;
}

#ifdef B2_NCOND
// ---- VectorContinuousCallback (qa.jl:124): B2_NCOND event functions.  The first of the interp_points sub-intervals in
// which ANY of them changes sign is searched; every function that changed sign there gets its own ITP root-find on that
// bracket; the earliest root fires (lower index on ties) and its index goes to affect!(integrator, idx).  After an
// event only the function that fired (ev_last) takes its reference sign at theta = 0.01, the others keep their sign at
// the step start, so a crossing right after the event (a corner) is not lost.
// vcond_start(g[]), vcond_at(theta, g[]), vcond_end(g[]) fill all B2_NCOND values.
template <class VStart, class VAt, class VEnd>
__device__ __forceinline__ bool b2_vevent_search(const int ip, const bool just_fired, int& ev_last, VStart&& vcond_start,
                                                 VAt&& vcond_at, VEnd&& vcond_end, real& th_end, int& ev_idx) {
    real gp_[B2_NCOND], gl_[B2_NCOND], gh_[B2_NCOND], gv_[B2_NCOND], lo_[B2_NCOND];
    unsigned chg = 0;
    bool fired = false;
    real hi = 0;
    vcond_start(gp_);
#pragma unroll
    for (int k = 0; k < B2_NCOND; k++) lo_[k] = 0;
    if (just_fired) {
        vcond_at((real)0.01, gv_);
#pragma unroll
        for (int k = 0; k < B2_NCOND; k++)
            if (k == ev_last) {
                gp_[k] = gv_[k];
                lo_[k] = (real)0.01;
            }
    }
#pragma unroll
    for (int k = 0; k < B2_NCOND; k++) gl_[k] = gp_[k];
    for (int mm = 1; mm <= ip && !fired; mm++) {
        const real th = (mm == ip) ? (real)1 : (real)mm / (real)ip;
        if (mm == ip) vcond_end(gv_);
        else vcond_at(th, gv_);
        chg = 0;
#pragma unroll
        for (int k = 0; k < B2_NCOND; k++)
            if (b2_sign_change(gp_[k], gv_[k])) chg |= 1u << k;
        if (chg) {
            fired = true;
            hi = th;
#pragma unroll
            for (int k = 0; k < B2_NCOND; k++) gh_[k] = gv_[k];
        } else {
#pragma unroll
            for (int k = 0; k < B2_NCOND; k++) {
                gl_[k] = gv_[k];
                lo_[k] = th;
            }
        }
    }
    if (fired) {
        real best = (real)2;
        int bidx = 0;
#pragma unroll
        for (int k = 0; k < B2_NCOND; k++) {
            if ((chg >> k) & 1u) {
                const real lo_k = b2_itp_left(lo_[k], hi, gl_[k], gh_[k], gp_[k], [&](real x) -> real {
                    real gx_[B2_NCOND];
                    vcond_at(x, gx_);
                    return gx_[k];
                });
                if (lo_k < best) {   // earliest event wins; the lower index on ties
                    best = lo_k;
                    bidx = k;
                }
            }
        }
        th_end = best;
        ev_idx = bidx;
        ev_last = bidx;
    }
    return fired;
}
#endif
