// b2_sde_adaptive.cuh -- ADAPTIVE SDE ensemble kernel (SRIW1, SOSRA): embedded error estimate, the PI controller of
// the ODE path with the strong order 3/2, and rejection sampling with memory (RSwM1, Rackauckas & Nie 2017) so that a
// rejected step keeps the part of the Brownian path it already sampled.  Compiled only into models built with
// B200ENS_MODEL_SDE_ADAPTIVE (B2_SDE_ADAPT); same expression tree as oracle/oracle_impl.inc solve_sde_adaptive.
// EXPERIMENTAL in round 1: SRIW1 parity with the oracle was measured on a B200, the SOSRA case has not run yet (the
// host API keeps it behind B200ENS_EXPERIMENTAL_SDE_ADAPTIVE=1).  Reference names: SDEProblem /root/reference/test/qa/qa.jl:103; the
// adaptive loop lives in StochasticDiffEq, outside the dep closure.
#pragma once
#include "b2_sde.cuh"
#include "b2_control.cuh"

#ifndef B2_RSWM_DEPTH
#define B2_RSWM_DEPTH 48   // remembered pieces per trajectory (one per consecutive rejection); overflow -> Failure
#endif
#define B2_RC_FAILURE_ 7

// Persistent warps with LANE REFILL, like the ODE driver (b2_ode_driver.cuh): step counts differ per path (GBM at tol
// 1e-3: 335 accepted steps on average, +-30 %; more with rejections), so every outer iteration a lane makes ONE step
// attempt of its path, finished lanes park, and when `refill_threshold` lanes of a warp are parked the warp hands them
// new paths from the global counter.  The remembered-increment stack lives in the lane's local memory and is simply
// reset (sp = 0) when the lane takes a new path.  Outputs go straight to global memory (no staging).
template <int ALG>
__device__ __forceinline__ void b2_sde_adaptive_driver(const B2Args& a) {
    static_assert(ALG == 7 || ALG == 9, "adaptive SDE stepping needs an embedded estimate: SOSRA or SRIW1");
    const unsigned lane = threadIdx.x & 31u;
    real* const gout = reinterpret_cast<real*>(a.out_u);
    const real* const gu0 = reinterpret_cast<const real*>(a.u0);
    const real* const gp = reinterpret_cast<const real*>(a.p);
    const real* const gsave = reinterpret_cast<const real*>(a.saveat);
    const int n_save = a.n_save;
    const int out_per_traj = n_save * B2_NOUT;
    const real t0 = B2_ARG(a, t0), t1 = B2_ARG(a, t1), dt_user = B2_ARG(a, dt);
    const real dtmax = B2_ARG(a, dtmax), dtmin = B2_ARG(a, dtmin);
    const real delta = ALG == 9 ? (real)(1.0 / 6.0) : (real)1;
    const B2Ctl ctl = b2_ctl_init(a);   // the PI controller of the ODE drivers (b2_control.cuh), strong order 3/2 exponents
    const float lqinit = ctl.lqinit;

    B2Sink sink;
    sink.stage = nullptr;
    sink.gout = gout;
    sink.margs = nullptr;   // fused moments: ODE kernels only
    sink.base = 0;
    // ---- per-lane path state
    real u[B2_N], p[B2_NPA];
    real t = t0, dt = dt_user, h = (real)0, hreq = (real)0;   // hreq: the step the controller asked for, before a remembered piece cut it
    float lq = lqinit;
    int si = 0, sp = 0, naccept = 0, nreject = 0;
    bool have = false;   // (h, dW, dZ) already hold the cut increments of a rejected step
    long long iter = 0, idx = -1;
    unsigned long long traj = 0;
    real zbuf[B2_NORMALS_PER_CALL];
    int zavail = 0;
    unsigned long long zblock = 0;
    // the remembered pieces of the Brownian path beyond t (local memory: indexed dynamically)
    real sk_len[B2_RSWM_DEPTH], sk_W[B2_RSWM_DEPTH][B2_N], sk_Z[B2_RSWM_DEPTH][B2_N];
    real dW[B2_N], dZ[B2_N], z[2 * B2_N];
    // noise_injected: the caller supplies the STANDARD NORMALS of every trajectory, [N][nsteps_noise], consumed in
    // order in place of the Philox stream -- pathwise parity with the oracle including every accept / reject
    const real* zinj = nullptr;
    long long zpos = 0;
    bool starved = false, active = false, exhausted = false;
    auto draw = [&]() {   // the next 2n normals of this trajectory's stream, in order
#pragma unroll
        for (int j = 0; j < 2 * B2_N; j++) {
            if (zinj) {
                if (zpos >= a.nsteps_noise) {
                    starved = true;
                    z[j] = (real)0;
                } else {
                    z[j] = zinj[zpos++];
                }
                continue;
            }
            if (zavail == 0) {
                b2_normals(a.seed, traj, zblock, zbuf);
                zblock++;
                zavail = B2_NORMALS_PER_CALL;
            }
            const int pos = B2_NORMALS_PER_CALL - zavail;
#if B2_F64
            z[j] = pos == 0 ? zbuf[0] : zbuf[1];
#else
            z[j] = pos == 0 ? zbuf[0] : pos == 1 ? zbuf[1] : pos == 2 ? zbuf[2] : zbuf[3];
#endif
            zavail--;
        }
    };

    for (;;) {
        // ---------------- retire / refill (warp-uniform control flow)
        const unsigned idle = __ballot_sync(B2_FULL, !active);
        if (idle) {
            const bool all_idle = idle == B2_FULL;
            if (exhausted ? all_idle : (__popc(idle) >= a.refill_threshold || all_idle)) {
                if (!exhausted) {
                    const long long my = b2_fetch(idle, a.work_counter, a.N, lane, exhausted);
                    if (!active && my >= 0) {
                        idx = my;
                        sink.base = idx * (long long)out_per_traj;
#pragma unroll
                        for (int i = 0; i < B2_N; i++) u[i] = gu0[idx * B2_N + i];
#pragma unroll
                        for (int i = 0; i < B2_NPARAM; i++) p[i] = gp[idx * B2_NPARAM + i];
                        t = t0;
                        dt = dt_user;
                        h = hreq = (real)0;
                        lq = lqinit;
                        si = sp = naccept = nreject = 0;
                        have = false;
                        iter = 0;
                        traj = a.traj_offset + (unsigned long long)idx;
                        zavail = 0;
                        zblock = 0;
                        zinj = a.noise_injected ? reinterpret_cast<const real*>(a.dW) + (size_t)idx * (size_t)a.nsteps_noise : nullptr;
                        zpos = 0;
                        starved = false;
                        while (si < n_save && __ldg(gsave + si) <= t0) {
                            sink.put(si, u);
                            si++;
                        }
                        active = true;
                    }
                }
                if (__ballot_sync(B2_FULL, active) == 0u) break;
            }
        }

        // ---------------- one step attempt per active lane (the body of the oracle's while (t < t1) loop; `continue` there
        // = fall through to the next outer iteration here)
        int rc = 0;
        if (active) do {
            iter++;
            if (iter > a.maxiters) {
                rc = B2_RC_MAXITERS;
                break;
            }
            if (dt != dt) {
                rc = B2_RC_DTNAN;
                break;
            }
            if (!have) {
                h = b2_min(dt, dtmax);
                bool clipped = false;
                if (h > t1 - t) {
                    h = t1 - t;
                    clipped = true;
                }
                if (!clipped && h <= b2_max(dtmin, (real)B2_EPS * b2_abs(t))) {
                    rc = B2_RC_DTLESSTHANMIN;
                    break;
                }
                hreq = h;
                if (sp == 0) {   // nothing remembered beyond t: fresh increments
                    draw();
                    const real sq = b2_sqrt(h);
#pragma unroll
                    for (int i = 0; i < B2_N; i++) {
                        dW[i] = sq * z[i];
                        dZ[i] = sq * z[B2_N + i];
                    }
                } else if (sk_len[sp - 1] <= h) {   // the remembered piece is the step
                    sp--;
                    h = sk_len[sp];
#pragma unroll
                    for (int i = 0; i < B2_N; i++) {
                        dW[i] = sk_W[sp][i];
                        dZ[i] = sk_Z[sp][i];
                    }
                } else {   // the step ends inside the remembered piece: Brownian bridge at q = h / L
                    const real L = sk_len[sp - 1], q = h / L, sd = b2_sqrt(((real)1 - q) * h);
                    draw();
#pragma unroll
                    for (int i = 0; i < B2_N; i++) {
                        dW[i] = b2_fma(sd, z[i], q * sk_W[sp - 1][i]);
                        dZ[i] = b2_fma(sd, z[B2_N + i], q * sk_Z[sp - 1][i]);
                        sk_W[sp - 1][i] -= dW[i];
                        sk_Z[sp - 1][i] -= dZ[i];
                    }
                    sk_len[sp - 1] = L - h;
                }
            }
            have = false;
            if (starved) {   // the injected stream of normals ran out
                rc = B2_RC_FAILURE_;
                break;
            }
            real up[B2_N], E1[B2_N], E2[B2_N];
#pragma unroll
            for (int i = 0; i < B2_N; i++) up[i] = u[i];
            b2_sde_step<ALG, true>(u, up, dW, dZ, p, t, h, E1, E2);
            float acc = 0.0f;
#pragma unroll
            for (int i = 0; i < B2_N; i++) {
                const real sk = b2_fma(b2_max(b2_abs(up[i]), b2_abs(u[i])), B2_RTOL(a, i), B2_ATOL(a, i));
                const float r = __fmul_rn((float)b2_fma(delta, E1[i], E2[i]), b2_rcp_nr((float)sk));
                acc = __fmaf_rn(r, r, acc);
            }
            const float EE2 = __fmul_rn(acc, __fdiv_rn(1.0f, (float)B2_N));
            const B2Decision d = b2_pi_controller(EE2, lq, ctl);
            if (d.isn) {
                rc = B2_RC_DTNAN;
                break;
            }
            if (!d.ok) {   // reject: keep the first part of the increments, remember the rest
                nreject++;
                const real qr = (real)d.qi;   // h' / h = max(qmin, gamma * EEst^-beta1), in [qmin, gamma]
                if (sp >= B2_RSWM_DEPTH) {
                    rc = B2_RC_FAILURE_;
                    break;
                }
                const real hn = qr * h, sd = b2_sqrt(((real)1 - qr) * hn);
                draw();
#pragma unroll
                for (int i = 0; i < B2_N; i++) {
                    const real w = b2_fma(sd, z[i], qr * dW[i]), zz = b2_fma(sd, z[B2_N + i], qr * dZ[i]);
                    sk_W[sp][i] = dW[i] - w;
                    sk_Z[sp][i] = dZ[i] - zz;
                    dW[i] = w;
                    dZ[i] = zz;
                    u[i] = up[i];
                }
                sk_len[sp++] = h - hn;
                h = hn;
                dt = hn;
                hreq = hn;
                have = true;
                if (starved) {
                    rc = B2_RC_FAILURE_;
                    break;
                }
                if (hn <= b2_max(dtmin, (real)B2_EPS * b2_abs(t))) {
                    rc = B2_RC_DTLESSTHANMIN;
                    break;
                }
                break;   // retry with the cut increments in the next outer iteration
            }
            lq = b2_ctl_lq_next(d, ctl);
            dt = h * (real)d.qi;
            // a step cut short to END ON a remembered time point does not shrink the proposal below what the controller
            // had asked for (the leftover of a rejected step can be arbitrarily short; see the oracle)
            if (h < hreq) dt = b2_max(dt, hreq);
            naccept++;
            const real tprev = t;
            real tnew = t + h;
            if (b2_abs(tnew - t1) < (real)100 * (real)B2_EPS * b2_max(b2_abs(tnew), b2_abs(t1))) tnew = t1;
            while (si < n_save) {   // linear interpolation between accepted steps
                const real tau = __ldg(gsave + si);
                if (!(tau <= tnew)) break;
                if (tau == tnew) {
                    sink.put(si, u);
                } else {
                    const real th = (tau - tprev) / h;
                    real w[B2_N];
#pragma unroll
                    for (int i = 0; i < B2_N; i++) w[i] = b2_fma(th, u[i] - up[i], up[i]);
                    sink.put(si, w);
                }
                si++;
            }
            t = tnew;
            if (!(t < t1)) rc = B2_RC_SUCCESS;
        } while (false);

        // ---------------- retire finished / failed lanes
        if (rc != 0) {
            if (rc != B2_RC_SUCCESS) sink.fill(si, n_save, (real)__int_as_float(0x7fc00000));
            a.retcode[idx] = rc;
            if (a.stats) {
                B2Stats s;
                s.naccept = naccept;
                s.nreject = nreject;
                s.nf = (naccept + nreject) * (ALG == 9 ? 2 : 3);
                s.nevents = 0;
                a.stats[idx] = s;
            }
            active = false;
        }
    }
}
