// cats_kernels.cu -- the batched CATS economy period kernel for sm_100a and its C-ABI.
//
// One thread block = one economy.  The economy blob is pulled from HBM into shared memory with a
// single TMA bulk copy (cp.async.bulk + mbarrier), advanced n_periods periods entirely on chip,
// and pushed back with a bulk store.  Phase order is the order of Cats.step
// (/root/reference/rmabm/environments/cats.py:355-419); every rule is docs/MODEL.md.
//
// Parallel structure inside a block
//   - per-firm and per-household rules: one thread per agent (strided loops)
//   - fp64 aggregates: det_sum, one warp per sum (fixed order => batch-size invariant, bit-exact)
//   - matching markets (job search, capital goods, consumption goods): a warp prepares 32 agents
//     at a time in parallel (visiting order, Philox candidate draws, price sort in registers)
//     and commits them in visiting order, so outcomes equal the sequential reference semantics.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>

#include "../../include/cats_b200.h"
#include "cats_device.cuh"
#include "cats_layout.h"

#ifndef CATS_GOODS_BACKOFF
#define CATS_GOODS_BACKOFF 1
#endif
#ifndef CATS_GOODS_FASTPATH
#define CATS_GOODS_FASTPATH 1
#endif

namespace cats {

struct KArgs {
    Layout L;
    TapeOff T;
    unsigned char *blobs;
    long long B;
    const double *params;
    long long params_stride;
    uint32_t k0, k1;
    unsigned long long economy_base;
    int n_periods;
    unsigned flags;
    int n_agents;
    const double *actions;
    const unsigned char *mask;
    double bankruptcy_reward;
    double *obs;
    double *reward;
    unsigned char *terminated;
    unsigned *status;
    double *stats;
    const unsigned char *tape;
    int stop_phase;
    unsigned char *debug;
    long long *cycles;  // debug: [16] per-phase clock64 totals of block 0
};

// per-block view of the shared-memory working set
struct Ctx {
    unsigned char *sm;
    unsigned char *blob;  // this economy's blob in global memory (C-firm rows, PA, perm live there)
    const KArgs *a;
    double *scal, *par;
    int *Leff, *Ld, *vac, *Oc;
    double *PA, *perm;  // household tail: GLOBAL memory (blob), not shared
    double *red;
    int *misc;
    unsigned char *bad;
    int tid, lane, warp, nwarps, nt;
    long long b;  // economy index inside this launch
    int last;     // last phase id to execute (stop_phase or 99)
    unsigned goods_epoch;  // goods markets run so far by this block (mbarrier phase bookkeeping)
    Draws d;
    __device__ __forceinline__ double *cf(int field) const { return reinterpret_cast<double *>(blob + a->L.o_c + field * a->L.rowF); }  // GLOBAL
    __device__ __forceinline__ double *kf(int field) const { return reinterpret_cast<double *>(sm + a->L.o_k + field * a->L.rowN); }
    __device__ __forceinline__ double *at(int off) const { return reinterpret_cast<double *>(sm + off); }
};

// misc int slots / red double slots
enum { M_NSEEK = 0, M_STATUS = 1, M_NS_C = 2, M_NS_K = 3, M_NBAD = 4, M_LSUM = 5, M_POOL = 6, M_WCNT = 8 /* .. +32 */ };
enum { R_SUM = 0, R_MEAN = 32 };

// ---------------------------------------------------------------------------------------------
// a8 firms_get_credit! for one firm (cats.py:376-377)
__device__ __forceinline__ void credit_one(const Ctx &c, double need, double b1, double b2, double &deb,
                                           double A, double &interest_r, double &Ftot, double liquidity,
                                           int &Ld, int Leff, int &vac)
{
    const double theta = c.par[P_theta], mu = c.par[P_mu], r_f = c.par[P_r_f];
    const double wb = c.scal[S_wb];
    double B = need > 0.0 ? need : 0.0;
    double credit = 0.0;
    if (B > 0.0) {
        double d0 = deb;
        double lev = (d0 + B) / ((A + B) + d0);
        double x = b1 + b2 * lev;
        double pr = 1.0 / (1.0 + det_exp(-x));
        double y = 1.0 + 1.0 / pr;
        double pw = det_exp(y * c.par[P_log1mtheta]);
        double Xi = (1.0 - pw) / theta;
        double rate = mu * ((1.0 + r_f / theta) / Xi - theta);
        double maxL = (c.par[P_phi] * c.scal[S_bank_E] - pr * d0) / pr;
        if (!(maxL > 0.0)) maxL = 0.0;
        credit = B < maxL ? B : maxL;
        double d1 = d0 + credit;
        deb = d1;
        if (d1 > 0.0) interest_r = (d0 * interest_r + rate * credit) / d1;
    }
    Ftot = credit;
    double afford = floor((credit + liquidity) / wb);
    double ld = (double)Ld;
    if (afford < ld) ld = afford;
    int Lc = d2count(ld);
    if (Lc < 1) Lc = 1;
    Ld = Lc;
    vac = Lc - Leff;
}

// a3-a8: every per-firm decision of the period in one pass (phases 3..10).  No cross-firm
// dependence: the bank rations against its start-of-period equity.
__device__ void phase_firm_decisions(Ctx &c)
{
    const Layout &L = c.a->L;
    const int F = L.F, N = L.N, last = c.last;
    const double alpha = c.par[P_alpha], q_adj = c.par[P_q_adj], p_adj = c.par[P_p_adj];
    const double wb = c.scal[S_wb];
    for (int j = c.tid; j < F + N; j += c.nt) {
        if (j < F) {
            const int f = j;
            double *De = c.cf(C_De), *P = c.cf(C_P);
            const double Yprev = c.cf(C_Y_prev)[f], Yd = c.cf(C_Yd)[f];
            double p = P[f], de = De[f];
            const double prevDe = de, prevP = p;
            // phase 3: a3 firms_decide_price_quantity! (cats.py:363)
            {
                double stock = Yprev - Yd, avg = c.scal[S_price];
                double u = c.d.uniform(PH_PRICE_C, f);
                if (stock <= 0.0) {
                    if (p >= avg) de = Yprev + (-stock) * q_adj;
                    else { de = Yprev; p = p * (1.0 + u * p_adj); }
                } else {
                    if (p > avg) { de = Yprev; p = p * (1.0 - u * p_adj); }
                    else de = Yprev - stock * q_adj;
                }
                if (de < alpha) de = alpha;
            }
            // phase 5: a5 firms_decide_investment! (cats.py:366)
            double kd = 0.0;
            if (last >= 5) {
                const double Iprob = c.par[P_Iprob], barX = c.par[P_barX], eta = c.par[P_eta];
                double u = c.d.uniform(PH_INVEST, f);
                if (u < Iprob) {
                    double K = c.cf(C_K)[f];
                    double Kdes = c.cf(C_barYK)[f] / barX;
                    double depn = ((barX * K) * eta) / Iprob;
                    kd = (Kdes - K) + depn;
                    if (!(kd > 0.0)) kd = 0.0;
                }
                c.at(L.o_cKdem)[f] = kd;
            }
            // phase 6: a6 _implement_action (cats.py:434-464 / 673-704)
            if (last >= 6 && f < c.a->n_agents && !(c.a->flags & CATS_FLAG_BURNIN) && c.a->actions) {
                const long long ai = c.b * c.a->n_agents + f;
                const bool real = c.a->mask ? (c.a->mask[ai] != 0) : true;
                if (real) {
                    double a0 = c.a->actions[2 * ai], a1 = c.a->actions[2 * ai + 1];
                    double base = (c.a->flags & CATS_FLAG_AVG_PRICE) ? c.scal[S_price] : prevP;
                    if (c.a->flags & CATS_FLAG_LOG) {
                        de = det_exp(det_log(prevDe + 1e-8) + a0);
                        p = det_exp(det_log(base + 1e-8) + a1);
                    } else {
                        de = prevDe * a0;
                        p = base * a1;
                    }
                    if (de < alpha) de = alpha;
                }
            }
            De[f] = de; P[f] = p;
            // phase 7: a7 firms_decide_labour! (cats.py:373)
            if (last >= 7) {
                double aa = ceil(de / alpha);
                double bb = ceil((c.cf(C_K)[f] * c.par[P_k]) / alpha);
                int Ld = d2count(aa < bb ? aa : bb);
                int vac = 0;
                // phase 9: a8 firms_get_credit! (cats.py:376)
                if (last >= 9) {
                    double liq = c.cf(C_liquidity)[f];
                    double need = (wb * (double)Ld + kd * c.scal[S_price_k]) - liq;
                    double deb = c.cf(C_deb)[f], ir = c.cf(C_interest_r)[f], Ftot;
                    credit_one(c, need, c.par[P_b1], c.par[P_b2], deb, c.cf(C_A)[f], ir, Ftot, liq, Ld,
                               c.Leff[f], vac);
                    c.cf(C_deb)[f] = deb; c.cf(C_interest_r)[f] = ir; c.at(L.o_cFtot)[f] = Ftot;
                }
                c.Ld[f] = Ld; c.vac[f] = vac;
            }
        } else {
            const int i = j - F;
            double *De = c.kf(K_De), *P = c.kf(K_P);
            const double Yprev = c.kf(K_Y_prev)[i], Yd = c.kf(K_Yd)[i], avail = c.kf(K_Yavail)[i];
            double p = P[i], de = De[i];
            // phase 4: a4 firms_decide_price_quantity! for K-firms (cats.py:364)
            if (last >= 4) {
                double stock = avail - Yd, avg = c.scal[S_price_k];
                double u = c.d.uniform(PH_PRICE_K, i);
                if (stock <= 0.0) {
                    if (p >= avg) de = Yprev + (-stock) * q_adj;
                    else { de = Yprev; p = p * (1.0 + u * p_adj); }
                } else {
                    if (p > avg) { de = Yprev; p = p * (1.0 - u * p_adj); }
                    else de = Yprev - stock * q_adj;
                }
                if (de < alpha) de = alpha;
                De[i] = de; P[i] = p;
            }
            // phase 8: a7 firms_decide_labour! for K-firms (cats.py:374)
            if (last >= 8) {
                int Ld = d2count(ceil(de / alpha));
                int vac = 0;
                // phase 10: a8 firms_get_credit! for K-firms (cats.py:377)
                if (last >= 10) {
                    double liq = c.kf(K_liquidity)[i];
                    double need = wb * (double)Ld - liq;
                    double deb = c.kf(K_deb)[i], ir = c.kf(K_interest_r)[i], Ftot;
                    credit_one(c, need, c.par[P_b_k1], c.par[P_b_k2], deb, c.kf(K_A)[i], ir, Ftot, liq, Ld,
                               c.Leff[F + i], vac);
                    c.kf(K_deb)[i] = deb; c.kf(K_interest_r)[i] = ir; c.at(L.o_kFtot)[i] = Ftot;
                }
                c.Ld[F + i] = Ld; c.vac[F + i] = vac;
            }
        }
    }
    __syncthreads();
    // bank loan stock: += det_sum(new credit), C-firms then K-firms
    if (last >= 9) {
        for (int q = c.warp; q < 2; q += c.nwarps) {
            const double *x = q == 0 ? c.at(L.o_cFtot) : c.at(L.o_kFtot);
            double s = warp_det_sum(q == 0 ? F : N, c.lane, [&](int i) { return x[i]; });
            if (c.lane == 0) c.red[R_SUM + q] = s;
        }
        __syncthreads();
        if (c.tid == 0) {
            double loans = c.scal[S_bank_loans] + c.red[R_SUM + 0];
            if (last >= 10) loans = loans + c.red[R_SUM + 1];
            c.scal[S_bank_loans] = loans;
        }
    }
}

// a9 firms_fire_workers! (cats.py:379-380).  A firm with s = Leff - Ld > 0 dismisses the s
// employees with the smallest (fire_key, worker index).  Step 1 gathers the employees of every
// over-staffed firm into a pool (one pass over the workers); step 2: one warp per over-staffed
// firm does s arg-min selections over the pool.  The result does not depend on the pool order.
__device__ void phase_fire(Ctx &c, int limit)
{
    const Layout &L = c.a->L;
    const int W = L.W;
    int *pool_w = reinterpret_cast<int *>(c.sm + L.o_scr);
    unsigned *pool_k = reinterpret_cast<unsigned *>(c.sm + L.o_scr) + W;
    if (c.tid == 0) c.misc[M_POOL] = 0;
    __syncthreads();
    for (int w = c.tid; w < W; w += c.nt) {
        int f = c.Oc[w];
        if (f >= 0 && f < limit && c.vac[f] < 0) {
            int i = atomicAdd(&c.misc[M_POOL], 1);
            pool_w[i] = w; pool_k[i] = c.d.fire_key(w);
        }
    }
    __syncthreads();
    const int n = c.misc[M_POOL];
    if (n == 0) return;
    for (int f = c.warp; f < limit; f += c.nwarps) {
        int s = -c.vac[f];
        if (s <= 0) continue;
        for (int r = 0; r < s; ++r) {
            unsigned bk = 0xffffffffu; int bw = 0x7fffffff;
            for (int i = c.lane; i < n; i += 32) {
                int w = pool_w[i];
                if (c.Oc[w] == f) {
                    unsigned key = pool_k[i];
                    if (key < bk || (key == bk && w < bw)) { bk = key; bw = w; }
                }
            }
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
                unsigned ok = __shfl_xor_sync(FULL, bk, off);
                int ow = __shfl_xor_sync(FULL, bw, off);
                if (ok < bk || (ok == bk && ow < bw)) { bk = ok; bw = ow; }
            }
            if (bw == 0x7fffffff) break;
            if (c.lane == 0) c.Oc[bw] = -1;
            __syncwarp();
        }
        if (c.lane == 0) { c.Leff[f] = c.Ld[f]; c.vac[f] = 0; }
    }
}

// a10 workers_search_job! (cats.py:382).  The unemployed are compacted in visiting order; warp 0
// then takes 32 of them at a time.  Every lane picks the first of its z_e candidates that still
// has a vacancy; the longest prefix of lanes whose picks cannot have been pre-empted by a lower
// lane is committed at once (lane j is safe iff fewer than vac[f] lower lanes chose the same f),
// the rest retry against the updated vacancies.  Outcome == the sequential visiting-order rule.
template <int ZM>
__device__ void phase_job_search(Ctx &c)
{
    const Layout &L = c.a->L;
    const int W = L.W, nf = L.F + L.N;
    const int z = c.d.z_e < nf ? c.d.z_e : nf;
    Order ord; ord.init(c.d, MKT_JOB, W);
    int *seekers = reinterpret_cast<int *>(c.sm + L.o_scr);
    int total = 0;
    for (int base = 0; base < W; base += c.nt) {
        int pos = base + c.tid, w = -1; bool un = false;
        if (pos < W) { w = ord.at(pos); un = c.Oc[w] < 0; }
        unsigned bal = __ballot_sync(FULL, un);
        if (c.lane == 0) c.misc[M_WCNT + c.warp] = __popc(bal);
        __syncthreads();
        int before = 0, all = 0;
        for (int q = 0; q < c.nwarps; ++q) { int n = c.misc[M_WCNT + q]; if (q < c.warp) before += n; all += n; }
        if (un) seekers[total + before + __popc(bal & ((1u << c.lane) - 1u))] = w;
        total += all;
        __syncthreads();
    }
    if (c.warp != 0) return;
    const unsigned lt = (1u << c.lane) - 1u;
    for (int base = 0; base < total; base += 32) {
        const int idx = base + c.lane;
        int w = -1;
        int cand[ZM];
        if (idx < total) { w = seekers[idx]; c.d.candidates<ZM>(PH_JOB, w, nf, z, cand); }
        unsigned pending = __ballot_sync(FULL, idx < total);
        while (pending) {
            const bool mine = (pending >> c.lane) & 1u;
            int pick = -1;
            if (mine) {
#pragma unroll
                for (int k = 0; k < ZM; ++k)
                    if (k < z && pick < 0 && c.vac[cand[k]] > 0) pick = cand[k];
            }
            const bool has = mine && pick >= 0;
            const unsigned peers = __match_any_sync(FULL, has ? pick : -1 - c.lane);
            const bool ok = !has || __popc(peers & lt) < c.vac[pick];
            const unsigned badm = __ballot_sync(FULL, mine && !ok);
            const unsigned commit = badm ? (pending & ((1u << (__ffs(badm) - 1)) - 1u)) : pending;
            __syncwarp();
            if (has && ((commit >> c.lane) & 1u)) {
                c.Oc[w] = pick;
                const unsigned grp = peers & commit;
                if ((grp & lt) == 0) { int n = __popc(grp); c.vac[pick] -= n; c.Leff[pick] += n; }
            }
            pending &= ~commit;
            __syncwarp();
        }
    }
}

// a11 firms_produce! (cats.py:384-385)
__device__ void phase_produce(Ctx &c)
{
    const Layout &L = c.a->L;
    const int F = L.F, N = L.N, last = c.last;
    const double alpha = c.par[P_alpha], k = c.par[P_k], eta = c.par[P_eta], theta = c.par[P_theta];
    const double wb = c.scal[S_wb], pk = c.scal[S_price_k];
    for (int j = c.tid; j < F + N; j += c.nt) {
        if (j < F) {
            const int f = j;
            double Leff = (double)c.Leff[f];
            double cap_l = Leff * alpha, cap_k = c.cf(C_K)[f] * k;
            double Yp = cap_l < cap_k ? cap_l : cap_k;
            double De = c.cf(C_De)[f];
            double Y = De < Yp ? De : Yp;
            c.cf(C_Y_prev)[f] = Y; c.at(L.o_cYstock)[f] = Y;
            double deb = c.cf(C_deb)[f];
            double interests = c.cf(C_interest_r)[f] * deb;
            double wages = wb * Leff;
            c.cf(C_interests)[f] = interests; c.cf(C_wages)[f] = wages;
            double Pl = 0.0;
            if (Y > 0.0) Pl = (1.01 * (((wages + interests) + ((Y / k) * eta) * pk) + theta * deb)) / Y;
            if (c.cf(C_P)[f] < Pl) c.cf(C_P)[f] = Pl;
            c.cf(C_Q)[f] = 0.0; c.cf(C_Yd)[f] = 0.0; c.cf(C_investment)[f] = 0.0; c.at(L.o_cvalinv)[f] = 0.0;
        } else if (last >= 15) {
            const int i = j - F;
            double Leff = (double)c.Leff[F + i];
            double want = c.kf(K_De)[i];
            double cap_l = Leff * alpha;
            double Y = want < cap_l ? want : cap_l;
            c.kf(K_Y_prev)[i] = Y;
            double avail = Y + c.kf(K_inv)[i];
            c.at(L.o_kYstock)[i] = avail; c.kf(K_Yavail)[i] = avail;
            double deb = c.kf(K_deb)[i];
            double interests = c.kf(K_interest_r)[i] * deb;
            double wages = wb * Leff;
            c.kf(K_interests)[i] = interests; c.kf(K_wages)[i] = wages;
            double Pl = 0.0;
            if (Y > 0.0) Pl = (1.01 * ((wages + interests) + theta * deb)) / Y;
            if (c.kf(K_P)[i] < Pl) c.kf(K_P)[i] = Pl;
            c.kf(K_Q)[i] = 0.0; c.kf(K_Yd)[i] = 0.0;
        }
    }
}

// a12 capital_goods_market! (cats.py:387)
template <int ZM>
__device__ void phase_capital_market(Ctx &c)
{
    if (c.warp != 0) return;
    const Layout &L = c.a->L;
    const int F = L.F, N = L.N;
    const int z = c.d.z_k < N ? c.d.z_k : N;
    Order ord; ord.init(c.d, MKT_CAP, F);
    double *Pk = c.kf(K_P), *Ydk = c.kf(K_Yd), *Qk = c.kf(K_Q), *Ysk = c.at(L.o_kYstock);
    double *Kdem = c.at(L.o_cKdem), *Ftot = c.at(L.o_cFtot), *liq = c.cf(C_liquidity);
    double *inv = c.cf(C_investment), *valinv = c.at(L.o_cvalinv), *wages = c.cf(C_wages);
    for (int base = 0; base < F; base += 32) {
        int pos = base + c.lane, f = -1; bool act = false; double kd = 0.0;
        int cand[ZM]; double pr[ZM];
        if (pos < F) { f = ord.at(pos); kd = Kdem[f]; act = kd > 0.0; }
        if (act) {
            c.d.candidates<ZM>(PH_CAP, f, N, z, cand);
#pragma unroll
            for (int k = 0; k < ZM; ++k) pr[k] = k < z ? Pk[cand[k]] : 0.0;
            sort_by_price<ZM>(cand, pr, z);
        }
        unsigned m = __ballot_sync(FULL, act);
        while (m) {
            int l = __ffs(m) - 1; m &= m - 1;
            if (c.lane == l) {
                double cb = (Ftot[f] + liq[f]) - wages[f];
                double lq = liq[f], iv = inv[f], vi = valinv[f];
#pragma unroll
                for (int k = 0; k < ZM; ++k) {
                    if (k < z && cb > 0.0 && kd > 0.0) {
                        int best = cand[k];
                        double pk = pr[k];
                        double afford = cb / pk;
                        double want = afford < kd ? afford : kd;
                        Ydk[best] = Ydk[best] + want;
                        double ys = Ysk[best];
                        if (ys > 0.0) {
                            double full = kd * pk;
                            double spend = cb < full ? cb : full;
                            double q = spend / pk;
                            if (ys > q) {
                                Ysk[best] = ys - q;
                                Qk[best] = Qk[best] + q;
                                kd = kd - q; cb = cb - spend;
                                lq = lq - spend; iv = iv + q; vi = vi + spend;
                            } else {
                                double cost = ys * pk;
                                kd = kd - ys; cb = cb - cost;
                                lq = lq - cost;
                                Qk[best] = Qk[best] + ys;
                                iv = iv + ys; vi = vi + cost;
                                Ysk[best] = 0.0;
                            }
                        }
                    }
                }
                liq[f] = lq; inv[f] = iv; valinv[f] = vi; Kdem[f] = kd;
            }
            __syncwarp();
        }
    }
}

// a13 gov_adjusts_subsidy! (cats.py:389) and the government side of a14 workers_get_paid!
// (cats.py:391): subsidy level, wage tax and subsidy totals.  The employed head-count is the sum
// of Leff (== #{Oc >= 0}: invariant kept by firing, job search and bankruptcy).
__device__ void phase_gov_income(Ctx &c)
{
    const Layout &L = c.a->L;
    if (c.warp == 0) {
        int n = warp_int_sum(L.F + L.N, c.lane, [&](int i) { return c.Leff[i]; });
        if (c.lane == 0) {
            double sub = c.par[P_subsidy] * c.scal[S_wb];
            if (c.par[P_maastricht] > 0.0 && c.scal[S_deficitGDP] > c.par[P_target_deficit]) {
                double scale = 1.0 - c.par[P_maastricht] * (c.scal[S_deficitGDP] - c.par[P_target_deficit]);
                if (!(scale > 0.0)) scale = 0.0;
                sub = sub * scale;
            }
            c.scal[S_gov_subsidy_level] = sub;
            if (c.last >= 18) {
                const double wb = c.scal[S_wb], tax = c.par[P_tax_rate];
                c.scal[S_gov_TA] = c.scal[S_gov_TA] + (wb * tax) * (double)n;
                c.scal[S_gov_G] = c.scal[S_gov_G] + sub * (double)(L.W - n);
            }
        }
    }
}

// income and consumption budget of one household (a14 workers_get_paid!, a15
// households_find_cons_budget!, cats.py:391-393); pa/pm are updated in place, returns the budget
__device__ __forceinline__ double household_budget(const Ctx &c, int h, double &pa, double &pm, int upto)
{
    const Layout &L = c.a->L;
    const double wb = c.scal[S_wb], sub = c.scal[S_gov_subsidy_level], price = c.scal[S_price];
    double inc;
    if (h < L.W) {
        inc = c.Oc[h] >= 0 ? wb * (1.0 - c.par[P_tax_rate]) : sub;
        pa = pa + inc;
    } else {
        inc = h < L.W + L.F ? c.cf(C_div_income)[h - L.W] : c.kf(K_div_income)[h - L.W - L.F];
    }
    if (upto < 19) return 0.0;
    const double xi = c.par[P_xi], chi = c.par[P_chi];
    double ir = inc / price;
    pm = pm * xi + (1.0 - xi) * ir;
    double target = pm + (chi * pa) / price;
    double b = target * price;
    if (b > pa) b = pa;
    if (!(b > 0.0)) b = 0.0;
    pa = pa - b;
    return b;
}

// debug path only (stop_phase 18 / 19): the household rules without the market
__device__ void phase_households_only(Ctx &c)
{
    for (int h = c.tid; h < c.a->L.H; h += c.nt) {
        double pa = c.PA[h], pm = c.perm[h];
        household_budget(c, h, pa, pm, c.last);
        c.PA[h] = pa; c.perm[h] = pm;
    }
}

// One consumption-goods seller as the market sees it.  The SoA rows (P, Yd, Q, stock) are packed
// into this 32-byte record when the market opens and unpacked when it closes, so that a visit is
// one shared-memory round trip instead of three dependent ones.
struct __align__(16) MktRec { double ys, yd, q, p; };

// a14 + a15 + a16 consumption_goods_market! (cats.py:391-394), fused and warp-specialised.
//
// Households are visited in the market's random order, 32 ("a chunk") at a time.
//   producers (warps 1..): for one chunk each, every lane pulls its household's wealth and
//       permanent income from global memory, computes income and budget (a14, a15), draws the z_c
//       candidate firms (Philox), sorts them by price and pre-divides the demand b/P it would
//       express at each -- then publishes the chunk in a shared-memory slot (mbarrier "full").
//   consumer (warp 0): takes the chunks in order and lets the lanes commit one after the other in
//       visiting order against the sellers' records -- the sequential-market rule of the reference.
//       All candidate records of a buyer are pulled up front, so a turn is one smem round trip.
// The serial chain is what bounds an economy-period; the producers keep everything that does not
// depend on market state off it.
struct GoodsSlot {            // one chunk of 32 prepared buyers, SoA over lanes
    double b[32];             // budget
    double pa[32];            // wealth after the budget was set aside
    double dk[MAX_Z][32];     // demand b / P at the k-th cheapest candidate
    int off[MAX_Z][32];       // shared-memory byte offset of the k-th candidate's MktRec (dummy record if k >= z)
    int h[32];                // household id (-1: no household in this lane)
};
static_assert(sizeof(GoodsSlot) == 32 * (8 + 8 + 8 * MAX_Z + 4 * MAX_Z + 4), "GoodsSlot must match cats_layout.h o_ring sizing");
constexpr int GOODS_SLOTS = 2;

template <int ZM>
__device__ void phase_goods_market(Ctx &c)
{
    const Layout &L = c.a->L;
    const int F = L.F, H = L.H;
    const int z = c.d.z_c < F ? c.d.z_c : F;
    MktRec *mkt = reinterpret_cast<MktRec *>(c.sm + L.o_scr);   // F records + 1 dummy
    {
        const double *P = c.cf(C_P), *Ys = c.at(L.o_cYstock);
        for (int f = c.tid; f <= F; f += c.nt) {
            MktRec r; r.ys = f < F ? Ys[f] : 0.0; r.yd = 0.0; r.q = 0.0; r.p = f < F ? P[f] : 1.0; mkt[f] = r;
        }
    }
    __syncthreads();
    const int nchunks = (H + 31) / 32;
    const int nprod = c.nwarps - 1;
    unsigned char *slots = c.sm + L.o_ring;
    uint64_t *bars = reinterpret_cast<uint64_t *>(c.sm + L.o_misc + 128);  // full[s] = bars[2s], empty[s] = bars[2s+1]
    if (c.warp > 0) {
        // ---------------- producers: chunk ch goes to slot ch % GOODS_SLOTS, made by warp 1 + ch % nprod
        Order ord; ord.init(c.d, MKT_GOODS, H);
        for (int ch = c.warp - 1; ch < nchunks; ch += nprod) {
            const int s = ch % GOODS_SLOTS;
            GoodsSlot *slot = reinterpret_cast<GoodsSlot *>(slots + (size_t)s * sizeof(GoodsSlot));
            const int pos = ch * 32 + c.lane;
            int h = -1; double pa = 0.0, pm = 0.0, b = 0.0;
            int cand[ZM]; double pr[ZM], dk[ZM];
#pragma unroll
            for (int k = 0; k < ZM; ++k) { dk[k] = 0.0; cand[k] = F; }
            if (pos < H) {
                h = ord.at(pos);
                pa = c.PA[h]; pm = c.perm[h];
                b = household_budget(c, h, pa, pm, 99);
                c.perm[h] = pm;
            }
            if (b > 0.0) {
                c.d.candidates<ZM>(PH_GOODS, h, F, z, cand);
#pragma unroll
                for (int k = 0; k < ZM; ++k) pr[k] = k < z ? mkt[cand[k]].p : 1.0;
                sort_by_price<ZM>(cand, pr, z);
#pragma unroll
                for (int k = 0; k < ZM; ++k) { dk[k] = k < z ? b / pr[k] : 0.0; if (k >= z) cand[k] = F; }
            }
            // wait until the consumer has drained the previous occupant of this slot
            const unsigned use = c.goods_epoch * (unsigned)((nchunks + GOODS_SLOTS - 1 - s) / GOODS_SLOTS) + (unsigned)(ch / GOODS_SLOTS);  // n-th use of the slot since kernel start
            mbar_wait_backoff(&bars[2 * s + 1], (use & 1u) ^ 1u);
            slot->b[c.lane] = b; slot->pa[c.lane] = pa; slot->h[c.lane] = h;
#pragma unroll
            for (int k = 0; k < ZM; ++k) { slot->dk[k][c.lane] = dk[k]; slot->off[k][c.lane] = cand[k] * (int)sizeof(MktRec); }
            __syncwarp();
            if (c.lane == 0) mbar_arrive(&bars[2 * s]);                  // publish
        }
    } else {
        // ---------------- consumer ----------------
        unsigned char *mbase = reinterpret_cast<unsigned char *>(mkt);
        for (int ch = 0; ch < nchunks; ++ch) {
            const int s = ch % GOODS_SLOTS;
            GoodsSlot *slot = reinterpret_cast<GoodsSlot *>(slots + (size_t)s * sizeof(GoodsSlot));
            mbar_wait(&bars[2 * s], (c.goods_epoch * (unsigned)((nchunks + GOODS_SLOTS - 1 - s) / GOODS_SLOTS) + (unsigned)(ch / GOODS_SLOTS)) & 1u);
            double b = slot->b[c.lane];
            const double pa = slot->pa[c.lane];
            const int h = slot->h[c.lane];
            double dk[ZM]; int off[ZM];
#pragma unroll
            for (int k = 0; k < ZM; ++k) { dk[k] = slot->dk[k][c.lane]; off[k] = slot->off[k][c.lane]; }
            __syncwarp();
            if (c.lane == 0) mbar_arrive(&bars[2 * s + 1]);              // slot may be refilled
            unsigned m = __ballot_sync(FULL, b > 0.0);
            while (m) {
                const int l = __ffs(m) - 1; m &= m - 1;
                if (c.lane == l) {
                    // pull every candidate's (stock, demand) pair up front: one latency per turn
                    double2 rec[ZM];
#pragma unroll
                    for (int k = 0; k < ZM; ++k) rec[k] = *reinterpret_cast<const double2 *>(mbase + off[k]);
#pragma unroll
                    for (int k = 0; k < ZM; ++k) {
                        if (b > 0.0) {
                            MktRec *r = reinterpret_cast<MktRec *>(mbase + off[k]);
                            const double d = dk[k];
                            double ys = rec[k].x;
                            const double yd = rec[k].y + d;
                            if (ys > 0.0) {
                                if (ys > d) {
                                    ys = ys - d;
                                    r->q = r->q + d;
                                    b = 0.0;
                                } else {
                                    b = b - ys * r->p;
                                    r->q = r->q + ys;
                                    ys = 0.0;
#pragma unroll
                                    for (int kk = k + 1; kk < ZM; ++kk)      // rare: partial fill
                                        dk[kk] = b / reinterpret_cast<const MktRec *>(mbase + off[kk])->p;
                                }
                            }
                            *reinterpret_cast<double2 *>(r) = make_double2(ys, yd);
                        }
                    }
                }
                __syncwarp();
            }
            if (h >= 0) c.PA[h] = pa + b;
        }
    }
    __syncthreads();
    c.goods_epoch += 1u;
    {
        double *Yd = c.cf(C_Yd), *Q = c.cf(C_Q), *Ys = c.at(L.o_cYstock);
        for (int f = c.tid; f < F; f += c.nt) { MktRec r = mkt[f]; Ys[f] = r.ys; Yd[f] = r.yd; Q[f] = r.q; }
    }
}

// a17 firms_accounting! (cats.py:397-398) + a18 reward (cats.py:493-556)
__device__ void phase_accounting(Ctx &c)
{
    const Layout &L = c.a->L;
    const int W = L.W, F = L.F, N = L.N, last = c.last;
    const double delta = c.par[P_delta], eta = c.par[P_eta], k = c.par[P_k], theta = c.par[P_theta];
    const double divs = c.par[P_div], taxd = c.par[P_tax_rate_d];
    double *inst_c = c.at(L.o_cFtot), *dtax_c = c.at(L.o_cKdem);    // rows reused once consumed
    double *inst_k = c.at(L.o_kFtot), *dtax_k = c.at(L.o_kYstock);
    for (int j = c.tid; j < F + N; j += c.nt) {
        if (j < F) {
            const int f = j;
            double Yp = c.cf(C_Y_prev)[f];
            c.cf(C_barYK)[f] = delta * c.cf(C_barYK)[f] + (1.0 - delta) * (Yp / k);
            double dep = (eta * Yp) / k;
            double Kold = c.cf(C_K)[f];
            c.cf(C_K)[f] = (Kold - dep) + c.cf(C_investment)[f];
            double dv = 0.0;
            double cv = c.cf(C_capital_value)[f];
            if (Kold != 0.0) dv = (dep * cv) / Kold;
            c.cf(C_capital_value)[f] = (cv - dv) + c.at(L.o_cvalinv)[f];
            double RIC = c.cf(C_P)[f] * c.cf(C_Q)[f];
            double wages = c.cf(C_wages)[f], interests = c.cf(C_interests)[f];
            double deb = c.cf(C_deb)[f];
            double in = theta * deb;
            double liq = ((((c.cf(C_liquidity)[f] + RIC) + c.at(L.o_cFtot)[f]) - wages) - interests) - in;
            c.cf(C_deb)[f] = deb - in;
            double pi = ((RIC - wages) - interests) - dv;
            double A = c.cf(C_A)[f] + pi;
            double net = 0.0, dt = 0.0;
            if (pi > 0.0) {
                double dvd = divs * pi;
                liq = liq - dvd; A = A - dvd;
                net = dvd * (1.0 - taxd);
                dt = dvd - net;
                c.PA[W + f] = c.PA[W + f] + net;
            }
            c.cf(C_div_income)[f] = net;
            c.cf(C_liquidity)[f] = liq; c.cf(C_A)[f] = A;
            inst_c[f] = in; dtax_c[f] = dt;
        } else if (last >= 22) {
            const int i = j - F;
            double RIC = c.kf(K_P)[i] * c.kf(K_Q)[i];
            c.kf(K_inv)[i] = (1.0 - c.par[P_inventory_depreciation]) * c.at(L.o_kYstock)[i];
            double wages = c.kf(K_wages)[i], interests = c.kf(K_interests)[i];
            double deb = c.kf(K_deb)[i];
            double in = theta * deb;
            double liq = ((((c.kf(K_liquidity)[i] + RIC) + c.at(L.o_kFtot)[i]) - wages) - interests) - in;
            c.kf(K_deb)[i] = deb - in;
            double pi = (RIC - wages) - interests;
            double A = c.kf(K_A)[i] + pi;
            double net = 0.0, dt = 0.0;
            if (pi > 0.0) {
                double dvd = divs * pi;
                liq = liq - dvd; A = A - dvd;
                net = dvd * (1.0 - taxd);
                dt = dvd - net;
                c.PA[W + F + i] = c.PA[W + F + i] + net;
            }
            c.kf(K_div_income)[i] = net;
            c.kf(K_liquidity)[i] = liq; c.kf(K_A)[i] = A;
            inst_k[i] = in; dtax_k[i] = dt;
        }
    }
    __syncthreads();
    // six det_sums: installments, interests, dividend tax for C then K
    for (int q = c.warp; q < 6; q += c.nwarps) {
        const bool isk = q >= 3; const int n = isk ? N : F;
        const double *x = q == 0 ? inst_c : q == 1 ? c.cf(C_interests) : q == 2 ? dtax_c
                        : q == 3 ? inst_k : q == 4 ? c.kf(K_interests) : dtax_k;
        double s = warp_det_sum(n, c.lane, [&](int i) { return x[i]; });
        if (c.lane == 0) c.red[R_SUM + q] = s;
    }
    __syncthreads();
    if (c.tid == 0) {
        double loans = c.scal[S_bank_loans] - c.red[R_SUM + 0];
        double ii = c.scal[S_bank_interest_income] + c.red[R_SUM + 1];
        double ta = c.scal[S_gov_TA] + c.red[R_SUM + 2];
        if (last >= 22) {
            loans = loans - c.red[R_SUM + 3];
            ii = ii + c.red[R_SUM + 4];
            ta = ta + c.red[R_SUM + 5];
        }
        c.scal[S_bank_loans] = loans; c.scal[S_bank_interest_income] = ii; c.scal[S_gov_TA] = ta;
    }
}

__device__ void phase_reward(Ctx &c)
{
    const KArgs &a = *c.a;
    if (!a.reward || a.n_agents <= 0) return;
    const int F = a.L.F, nA = a.n_agents;
    const double *P = c.cf(C_P), *Q = c.cf(C_Q);
    if (a.flags & CATS_FLAG_REWARD_RMS) {
        // cats.py:548-549: Python sum() over the RL firms, then over the others, left to right
        if (c.tid == 0) {
            double tot_rl = 0.0, tot_other = 0.0;
            for (int f = 0; f < nA; ++f) tot_rl = tot_rl + P[f] * Q[f];
            for (int f = nA; f < F; ++f) tot_other = tot_other + P[f] * Q[f];
            c.red[R_SUM + 8] = tot_rl + tot_other;
        }
        __syncthreads();
        double total = c.red[R_SUM + 8];
        for (int i = c.tid; i < nA; i += c.nt) a.reward[c.b * nA + i] = (P[i] * Q[i]) / total;
        return;
    }
    const double eta = c.par[P_eta], k = c.par[P_k];
    for (int i = c.tid; i < nA; i += c.nt) {
        double dep = (eta * c.cf(C_Y_prev)[i]) / k;
        double den = (c.cf(C_K)[i] + dep) - c.cf(C_investment)[i];
        double dv;
        if (den == 0.0) { dv = 0.0; atomicOr(&c.misc[M_STATUS], (int)CATS_STATUS_DEPVAL_DIV0); }
        else dv = (dep * c.cf(C_capital_value)[i]) / den;
        double RIC = P[i] * Q[i];
        double profits = ((RIC - c.cf(C_wages)[i]) - c.cf(C_interests)[i]) - dv;
        profits = (profits * c.scal[S_init_price]) / c.scal[S_price];
        if (profits != profits) { profits = 0.0; atomicOr(&c.misc[M_STATUS], (int)CATS_STATUS_PROFIT_NAN); }
        if (c.cf(C_A)[i] <= 0.0) profits = a.bankruptcy_reward;
        a.reward[c.b * nA + i] = profits;
    }
}

// a19 update_tracking_variables! (cats.py:401)
__device__ void phase_tracking(Ctx &c)
{
    const Layout &L = c.a->L;
    const int F = L.F, N = L.N;
    const double *cP = c.cf(C_P), *cQ = c.cf(C_Q), *kP = c.kf(K_P), *kQ = c.kf(K_Q);
    for (int q = c.warp; q < 9; q += c.nwarps) {
        if (q == 8) {
            int n = warp_int_sum(F + N, c.lane, [&](int i) { return c.Leff[i]; });
            if (c.lane == 0) c.misc[M_LSUM] = n;
            continue;
        }
        double s;
        switch (q) {
        case 0: s = warp_det_sum(F, c.lane, [&](int i) { return cQ[i]; }); break;
        case 1: s = warp_det_sum(F, c.lane, [&](int i) { return cP[i] * cQ[i]; }); break;
        case 2: s = warp_det_sum(F, c.lane, [&](int i) { return cP[i]; }); break;
        case 3: s = warp_det_sum(N, c.lane, [&](int i) { return kQ[i]; }); break;
        case 4: s = warp_det_sum(N, c.lane, [&](int i) { return kP[i] * kQ[i]; }); break;
        case 5: { const double *x = c.cf(C_Y_prev); s = warp_det_sum(F, c.lane, [&](int i) { return x[i]; }); } break;
        case 6: { const double *x = c.kf(K_Y_prev); s = warp_det_sum(N, c.lane, [&](int i) { return x[i]; }); } break;
        default: { const double *x = c.cf(C_K); s = warp_det_sum(F, c.lane, [&](int i) { return x[i]; }); } break;
        }
        if (c.lane == 0) c.red[R_SUM + q] = s;
    }
    __syncthreads();
    if (c.tid == 0) {
        const double *r = c.red + R_SUM;
        double *s = c.scal;
        double sQ = r[0], sPQ = r[1], sQk = r[3], sPQk = r[4], sY = r[5], sYk = r[6];
        double price_prev = s[S_price];
        double price = sQ > 0.0 ? sPQ / sQ : r[2] / (double)F;
        double price_k = sQk > 0.0 ? sPQk / sQk : s[S_price_k];
        double Yreal = sY * s[S_init_price] + sYk * s[S_init_price_k];
        double Ynom = sY * price + sYk * price_k;
        s[S_price] = price; s[S_price_k] = price_k;
        s[S_Y_real] = Yreal; s[S_Y_nominal_tot] = Ynom;
        s[S_gdp_deflator] = Yreal > 0.0 ? Ynom / Yreal : 0.0;
        s[S_inflationRate] = price / price_prev - 1.0;
        s[S_Un] = 1.0 - (double)c.misc[M_LSUM] / (double)L.W;
        s[S_consumption] = sPQ;
        s[S_totK] = r[7];
    }
}

// a20 firms_go_bankrupt! (cats.py:404) + a21 _get_terminated (cats.py:483-486)
__device__ void phase_bankrupt(Ctx &c)
{
    const Layout &L = c.a->L;
    const int W = L.W, F = L.F, N = L.N;
    const double thr = c.par[P_E_threshold_scale] * c.scal[S_price];
    unsigned char *bad = c.bad;
    for (int j = c.tid; j < F + N; j += c.nt) {
        double A = j < F ? c.cf(C_A)[j] : c.kf(K_A)[j - F];
        double liq = j < F ? c.cf(C_liquidity)[j] : c.kf(K_liquidity)[j - F];
        bad[j] = ((A <= thr) || (liq < 0.0)) ? 1 : 0;
    }
    __syncthreads();
    // 11 masked det_sums + 2 survivor counts
    for (int q = c.warp; q < 13; q += c.nwarps) {
        if (q >= 11) {
            const int isk = q - 11;
            int n = warp_int_sum(isk ? N : F, c.lane, [&](int i) { return bad[isk ? F + i : i] ? 0 : 1; });
            if (c.lane == 0) c.misc[isk ? M_NS_K : M_NS_C] = n;
            continue;
        }
        double s;
        if (q < 6) {
            const double *A = c.cf(C_A), *liq = c.cf(C_liquidity), *deb = c.cf(C_deb);
            const double *x = q == 0 ? c.cf(C_K) : q == 1 ? liq : q == 2 ? c.cf(C_P) : c.cf(C_Y_prev);
            s = warp_det_sum(F, c.lane, [&](int i) {
                const bool b = bad[i] != 0;
                if (q < 4) return b ? 0.0 : x[i];
                if (q == 5) return b ? deb[i] : 0.0;
                double lq = liq[i], a = A[i];
                double expo = deb[i] + (lq < 0.0 ? -lq : 0.0);
                double loss = a < 0.0 ? -a : 0.0;
                if (loss > expo) loss = expo;
                return b ? loss : 0.0;
            });
        } else {
            const int qq = q - 6;  // 0 liq, 1 P, 2 Y_prev, 3 loss, 4 deb
            const double *A = c.kf(K_A), *liq = c.kf(K_liquidity), *deb = c.kf(K_deb);
            const double *x = qq == 0 ? liq : qq == 1 ? c.kf(K_P) : c.kf(K_Y_prev);
            s = warp_det_sum(N, c.lane, [&](int i) {
                const bool b = bad[F + i] != 0;
                if (qq < 3) return b ? 0.0 : x[i];
                if (qq == 4) return b ? deb[i] : 0.0;
                double lq = liq[i], a = A[i];
                double expo = deb[i] + (lq < 0.0 ? -lq : 0.0);
                double loss = a < 0.0 ? -a : 0.0;
                if (loss > expo) loss = expo;
                return b ? loss : 0.0;
            });
        }
        if (c.lane == 0) c.red[R_SUM + q] = s;
    }
    __syncthreads();
    if (c.tid == 0) {
        const double *r = c.red + R_SUM;
        double *mean = c.red + R_MEAN;
        int nsC = c.misc[M_NS_C], nsK = c.misc[M_NS_K];
        double mK = 10.0, mL = 0.0, mP = c.scal[S_price], mY = c.par[P_alpha];
        if (nsC > 0) { mK = r[0] / (double)nsC; mL = r[1] / (double)nsC; mP = r[2] / (double)nsC; mY = r[3] / (double)nsC; }
        mean[0] = mK; mean[1] = mL; mean[2] = mP; mean[3] = mY;
        double bd = c.scal[S_bank_bad_debt] + r[4];
        double loans = c.scal[S_bank_loans] - r[5];
        double mLk = 0.0, mPk = c.scal[S_price_k], mYk = c.par[P_alpha];
        if (nsK > 0) { mLk = r[6] / (double)nsK; mPk = r[7] / (double)nsK; mYk = r[8] / (double)nsK; }
        mean[4] = mLk; mean[5] = mPk; mean[6] = mYk;
        bd = bd + r[9];
        loans = loans - r[10];
        c.scal[S_bank_bad_debt] = bd; c.scal[S_bank_loans] = loans;
        c.scal[S_bankruptcy_rate] = (double)((F - nsC) + (N - nsK)) / (double)(F + N);
    }
    __syncthreads();
    const double *mean = c.red + R_MEAN;
    for (int j = c.tid; j < F + N; j += c.nt) {
        if (!bad[j]) continue;
        if (j < F) {
            const int f = j, ow = W + f;
            double inj = mean[1] > 0.0 ? mean[1] : 0.0;
            double pa = c.PA[ow];
            if (inj > pa) inj = pa;
            if (!(inj > 0.0)) inj = 0.0;
            c.PA[ow] = pa - inj;
            double mK = mean[0], mY = mean[3];
            double cv = mK * c.scal[S_price_k];
            c.cf(C_liquidity)[f] = inj; c.cf(C_K)[f] = mK; c.cf(C_capital_value)[f] = cv;
            c.cf(C_A)[f] = inj + cv; c.cf(C_deb)[f] = 0.0; c.cf(C_P)[f] = mean[2];
            c.cf(C_Y_prev)[f] = mY; c.cf(C_Yd)[f] = mY; c.cf(C_De)[f] = mY;
            c.cf(C_barYK)[f] = mY / c.par[P_k]; c.cf(C_interest_r)[f] = c.par[P_r_f];
            c.cf(C_Q)[f] = 0.0; c.cf(C_investment)[f] = 0.0; c.cf(C_wages)[f] = 0.0;
            c.cf(C_interests)[f] = 0.0; c.cf(C_div_income)[f] = 0.0;
            c.Leff[f] = 0;
        } else {
            const int i = j - F, ow = W + F + i;
            double inj = mean[4] > 0.0 ? mean[4] : 0.0;
            double pa = c.PA[ow];
            if (inj > pa) inj = pa;
            if (!(inj > 0.0)) inj = 0.0;
            c.PA[ow] = pa - inj;
            double mY = mean[6];
            c.kf(K_liquidity)[i] = inj; c.kf(K_A)[i] = inj; c.kf(K_deb)[i] = 0.0;
            c.kf(K_P)[i] = mean[5]; c.kf(K_Y_prev)[i] = mY; c.kf(K_Yd)[i] = mY; c.kf(K_Yavail)[i] = mY;
            c.kf(K_De)[i] = mY; c.kf(K_inv)[i] = 0.0; c.kf(K_interest_r)[i] = c.par[P_r_f];
            c.kf(K_Q)[i] = 0.0; c.kf(K_wages)[i] = 0.0; c.kf(K_interests)[i] = 0.0;
            c.kf(K_div_income)[i] = 0.0;
            c.Leff[F + i] = 0;
        }
    }
    for (int w = c.tid; w < W; w += c.nt) { int o = c.Oc[w]; if (o >= 0 && bad[o]) c.Oc[w] = -1; }
    __syncthreads();
    if (c.warp == 0) {
        const double *A = c.cf(C_A);
        double s = warp_det_sum(F, c.lane, [&](int i) { return A[i]; });
        if (c.lane == 0) c.scal[S_terminated] = s == 0.0 ? 1.0 : 0.0;
    }
}

// a22 bank_accounting!, a23 gov_accounting!, a24 adjust_wages!, a25 timestep (cats.py:407-413)
__device__ void phase_closing(Ctx &c)
{
    const Layout &L = c.a->L;
    const int last = c.last;
    if (c.tid == 0) {
        double *s = c.scal;
        double gross = (s[S_bank_interest_income] + s[S_gov_r_b] * s[S_gov_bonds]) - s[S_bank_bad_debt];
        double dB = 0.0;
        if (gross > 0.0 && s[S_bank_E] > 0.0) dB = c.par[P_div] * gross;
        s[S_bank_E] = s[S_bank_E] + (gross - dB);
        s[S_dividendsB] = dB; s[S_profitsB] = gross;
        c.red[R_SUM + 0] = dB > 0.0 ? dB / (double)(L.F + L.N) : 0.0;
        c.misc[M_NBAD] = dB > 0.0 ? 1 : 0;
        if (last >= 27) {
            s[S_gov_G] = s[S_gov_G] + s[S_gov_r_b] * s[S_gov_bonds];
            double GB = s[S_gov_TA] - s[S_gov_G];
            s[S_GB] = GB;
            s[S_gov_bonds] = s[S_gov_bonds] - GB;
            s[S_deficitGDP] = s[S_Y_nominal_tot] > 0.0 ? (-GB) / s[S_Y_nominal_tot] : 0.0;
        }
        if (last >= 28) {
            double Un = s[S_Un], ut = c.par[P_u_target];
            if (Un < ut) s[S_wb] = s[S_wb] * (1.0 + c.par[P_wage_update_up] * (ut - Un));
            else s[S_wb] = s[S_wb] * (1.0 - c.par[P_wage_update_down] * (Un - ut));
        }
        if (last >= 29) s[S_timestep] = s[S_timestep] + 1.0;
    }
    __syncthreads();
    if (c.misc[M_NBAD]) {
        double share = c.red[R_SUM + 0];
        for (int h = L.W + c.tid; h < L.H; h += c.nt) c.PA[h] = c.PA[h] + share;
    }
}

// ---------------------------------------------------------------------------------------------
// one period; phase ids as docs/MODEL.md section 6 (stop_phase halts after the given id)
template <int ZM>
__device__ void run_period(Ctx &c)
{
    const KArgs &a = *c.a;
    const int F = a.L.F, N = a.L.N, last = c.last;
    long long t0 = 0;
    if (a.cycles && c.tid == 0 && blockIdx.x == 0) t0 = clock64();
#define TICK(slot) do { if (a.cycles && c.tid == 0 && blockIdx.x == 0) { long long t1 = clock64(); a.cycles[slot] += t1 - t0; t0 = t1; } } while (0)
    // phases 1-2: reset_gov_and_banking_variables!, set_bond_interest_rate! (cats.py:355-356)
    if (c.tid == 0) {
        c.scal[S_gov_TA] = 0.0; c.scal[S_gov_G] = 0.0;
        c.scal[S_bank_interest_income] = 0.0; c.scal[S_bank_bad_debt] = 0.0;
        if (last >= 2) c.scal[S_gov_r_b] = c.par[P_interest_rate];
        c.misc[M_STATUS] = 0;
    }
    __syncthreads();
    if (last < 3) return;
    phase_firm_decisions(c);  // 3..10
    __syncthreads();
    TICK(0);
    if (last < 11) return;
    phase_fire(c, last >= 12 ? F + N : F);  // 11, 12
    __syncthreads();
    TICK(1);
    if (last < 13) return;
    phase_job_search<ZM>(c);  // 13
    __syncthreads();
    TICK(2);
    if (last < 14) return;
    phase_produce(c);         // 14, 15
    __syncthreads();
    TICK(3);
    if (last < 16) return;
    phase_capital_market<ZM>(c);  // 16
    __syncthreads();
    TICK(4);
    if (last < 17) return;
    phase_gov_income(c);      // 17 (+ government side of 18)
    __syncthreads();
    TICK(5);
    if (last < 18) return;
    if (last < 20) { phase_households_only(c); __syncthreads(); return; }
    phase_goods_market<ZM>(c);  // 18, 19, 20
    __syncthreads();
    TICK(6);
    if (last < 21) return;
    phase_accounting(c);      // 21, 22
    __syncthreads();
    TICK(7);
    if (last < 23) return;
    phase_reward(c);          // 23
    __syncthreads();
    TICK(8);
    if (last < 24) return;
    phase_tracking(c);        // 24
    __syncthreads();
    TICK(9);
    if (last < 25) return;
    phase_bankrupt(c);        // 25
    __syncthreads();
    TICK(10);
    if (last < 26) return;
    phase_closing(c);         // 26..29
    __syncthreads();
    TICK(11);
#undef TICK
}

template <int NT, int ZM, int MINB>
__global__ void __launch_bounds__(NT, MINB) cats_period_kernel(const __grid_constant__ KArgs a)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const Layout &L = a.L;
    Ctx c;
    c.sm = smem; c.a = &a;
    c.scal = reinterpret_cast<double *>(smem + L.o_scal);
    c.par = reinterpret_cast<double *>(smem + L.o_par);
    c.Leff = reinterpret_cast<int *>(smem + L.o_Leff);
    c.Ld = reinterpret_cast<int *>(smem + L.o_Ld);
    c.vac = reinterpret_cast<int *>(smem + L.o_vac);
    c.Oc = reinterpret_cast<int *>(smem + L.o_Oc);
    c.red = reinterpret_cast<double *>(smem + L.o_red);
    c.misc = reinterpret_cast<int *>(smem + L.o_misc);
    c.bad = smem + L.o_bad;
    c.tid = threadIdx.x; c.lane = threadIdx.x & 31; c.warp = threadIdx.x >> 5;
    c.nt = NT; c.nwarps = NT / 32;
    c.last = a.stop_phase > 0 ? a.stop_phase : 99;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + L.o_misc + 192);

    c.goods_epoch = 0u;
    if (c.tid == 0) {
        mbar_init(bar, 1);
        uint64_t *ring = reinterpret_cast<uint64_t *>(smem + L.o_misc + 128);
        for (int i = 0; i < 2 * GOODS_SLOTS; ++i) mbar_init(&ring[i], 1);
        fence_mbar_init();
    }
    __syncthreads();
    uint32_t parity = 0;

    for (long long b = blockIdx.x; b < a.B; b += gridDim.x) {
        c.b = b;
        unsigned char *blob = a.blobs + (size_t)b * (size_t)L.blob;
        c.blob = blob;
        c.PA = reinterpret_cast<double *>(blob + L.o_PA);
        c.perm = reinterpret_cast<double *>(blob + L.o_perm);
        // HBM -> smem: one TMA bulk copy of the resident prefix (firm tables, employment links)
        if (c.tid == 0) {
            fence_proxy_async();
            mbar_expect_tx(bar, (uint32_t)L.resident);
            bulk_g2s(smem, blob, (uint32_t)L.resident, bar);
        }
        // parameters (shared row or per-economy row) while the copy is in flight
        const double *pr = a.params + (a.params_stride ? (size_t)b * (size_t)a.params_stride : 0);
        for (int i = c.tid; i < N_PARAMS; i += NT) c.par[i] = pr[i];
        if (c.tid == 0) c.par[P_log1mtheta] = det_log(1.0 - pr[P_theta]);
        // zero the transients so that debug images are deterministic
        for (int i = L.resident / 4 + c.tid; i < L.workset / 4; i += NT) reinterpret_cast<int *>(smem)[i] = 0;
        mbar_wait(bar, parity);
        parity ^= 1u;
        __syncthreads();

        c.d.k0 = a.k0; c.d.k1 = a.k1;
        c.d.economy = (uint32_t)(a.economy_base + (unsigned long long)b);
        c.d.t = a.T;
        c.d.z_c = (int)c.par[P_z_c]; c.d.z_k = (int)c.par[P_z_k]; c.d.z_e = (int)c.par[P_z_e];

        for (int t = 0; t < a.n_periods; ++t) {
            c.d.period = (uint32_t)c.scal[S_timestep];
            c.d.tape = a.tape ? a.tape + ((size_t)b * (size_t)a.n_periods + (size_t)t) * (size_t)a.T.bytes : nullptr;
            run_period<ZM>(c);
            __syncthreads();
            if (a.stats && c.tid < N_STATS) {
                const double *s = c.scal;
                double v;
                switch (c.tid) {
                case 0: v = s[S_Y_real]; break;       case 1: v = s[S_wb]; break;
                case 2: v = s[S_Un]; break;           case 3: v = s[S_bankruptcy_rate]; break;
                case 4: v = s[S_price]; break;        case 5: v = s[S_Y_nominal_tot]; break;
                case 6: v = s[S_gdp_deflator]; break; case 7: v = s[S_inflationRate]; break;
                case 8: v = s[S_consumption]; break;  case 9: v = s[S_totK]; break;
                case 10: v = s[S_dividendsB]; break;  case 11: v = s[S_profitsB]; break;
                case 12: v = s[S_GB]; break;          case 13: v = s[S_deficitGDP]; break;
                case 14: v = s[S_gov_bonds]; break;   default: v = s[S_price_k]; break;
                }
                a.stats[((size_t)b * (size_t)a.n_periods + (size_t)t) * N_STATS + c.tid] = v;
            }
            if (c.tid == 0 && c.misc[M_STATUS])
                c.scal[S_status] = (double)((unsigned)c.scal[S_status] | (unsigned)c.misc[M_STATUS]);
        }
        __syncthreads();
        // a26 _get_obs (cats.py:466-481 / 655-671), a21 terminated, status
        if (a.obs) {
            for (int i = c.tid; i < a.n_agents; i += NT) {
                double yp = c.cf(C_Y_prev)[i], yd = c.cf(C_Yd)[i], p = c.cf(C_P)[i], ap = c.scal[S_price];
                double o0, o1;
                if (a.flags & CATS_FLAG_LOG) {
                    o0 = det_log(yp + 1e-8) - det_log(yd + 1e-8);
                    o1 = det_log(p + 1e-8) - det_log(ap + 1e-8);
                } else { o0 = yp - yd; o1 = p - ap; }
                a.obs[2 * (b * a.n_agents + i)] = o0;
                a.obs[2 * (b * a.n_agents + i) + 1] = o1;
            }
        }
        if (c.tid == 0) {
            if (a.terminated) a.terminated[b] = c.scal[S_terminated] != 0.0 ? 1 : 0;
            if (a.status) a.status[b] = (unsigned)c.scal[S_status];
        }
        if (a.debug) {
            const int4 *src = reinterpret_cast<const int4 *>(smem);
            int4 *dst = reinterpret_cast<int4 *>(a.debug + (size_t)b * (size_t)L.workset);
            for (int i = c.tid; i < L.workset / 16; i += NT) dst[i] = src[i];
        }
        // smem -> HBM: bulk store of the resident prefix
        fence_proxy_async();
        __syncthreads();
        if (c.tid == 0) {
            bulk_s2g(blob, smem, (uint32_t)L.resident);
            bulk_commit();
            bulk_wait_all();
        }
        __syncthreads();
    }
}

// a0: initialise_model (cats.py:224) -- deterministic initial state, docs/MODEL.md section 5
__global__ void cats_init_kernel(unsigned char *blobs, long long B, Layout L, const double *params,
                                 long long params_stride)
{
    const long long b = blockIdx.x;
    if (b >= B) return;
    unsigned char *blob = blobs + (size_t)b * (size_t)L.blob;
    const double *par = params + (params_stride ? (size_t)b * (size_t)params_stride : 0);
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < L.blob / 8; i += nt) reinterpret_cast<double *>(blob)[i] = 0.0;
    __syncthreads();
    const double P0 = 3.0, PK0 = 3.0, K0 = 10.0, LIQ0 = 10.0, YC0 = 5.0, YK0 = 3.0, PA0 = 2.0;
    double *scal = reinterpret_cast<double *>(blob + L.o_scal);
    if (tid == 0) {
        scal[S_price] = P0; scal[S_price_k] = PK0; scal[S_init_price] = P0; scal[S_init_price_k] = PK0;
        scal[S_wb] = par[P_wb]; scal[S_bank_E] = 3000.0; scal[S_gdp_deflator] = 1.0; scal[S_Un] = 1.0;
        scal[S_gov_r_b] = par[P_interest_rate];
    }
    auto cf = [&](int f) { return reinterpret_cast<double *>(blob + L.o_c + f * L.rowF); };
    auto kf = [&](int f) { return reinterpret_cast<double *>(blob + L.o_k + f * L.rowN); };
    for (int f = tid; f < L.F; f += nt) {
        cf(C_De)[f] = 1.0; cf(C_P)[f] = P0; cf(C_Y_prev)[f] = YC0; cf(C_Yd)[f] = YC0;
        cf(C_K)[f] = K0; cf(C_capital_value)[f] = K0 * PK0; cf(C_liquidity)[f] = LIQ0;
        cf(C_A)[f] = LIQ0 + K0 * PK0; cf(C_barYK)[f] = YC0 / par[P_k]; cf(C_interest_r)[f] = par[P_r_f];
    }
    for (int i = tid; i < L.N; i += nt) {
        kf(K_De)[i] = 1.0; kf(K_P)[i] = PK0; kf(K_Y_prev)[i] = YK0; kf(K_Yd)[i] = YK0;
        kf(K_A)[i] = LIQ0; kf(K_liquidity)[i] = LIQ0; kf(K_interest_r)[i] = par[P_r_f];
        kf(K_Yavail)[i] = YK0;
    }
    double *PA = reinterpret_cast<double *>(blob + L.o_PA), *perm = reinterpret_cast<double *>(blob + L.o_perm);
    for (int h = tid; h < L.H; h += nt) { PA[h] = PA0; perm[h] = 1.0 / P0; }
    int *Oc = reinterpret_cast<int *>(blob + L.o_Oc);
    for (int w = tid; w < L.W; w += nt) Oc[w] = -1;
}

}  // namespace cats

// =============================================================================================
// C-ABI
using namespace cats;

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};
static long long *g_cycles = nullptr;  // debug: set by cats_debug_phase_cycles()
constexpr int kThreads = 64;    // threads per economy block: one market consumer warp + one producer warp
constexpr int kMinBlocks = 8;   // __launch_bounds__ hint: register budget for >= 8 resident economies / SM
typedef void (*KernelFn)(const KArgs);
static KernelFn select_kernel(int max_z)
{
    if (max_z > 0 && max_z <= 5) return cats_period_kernel<kThreads, 5, kMinBlocks>;
    return cats_period_kernel<kThreads, MAX_Z, kMinBlocks>;
}

static int fail(int code, const char *fmt, const char *detail)
{
    snprintf(g_err, sizeof g_err, fmt, detail);
    return code;
}

static int check_sizes(int W, int F, int N)
{
    if (W < 1 || F < 1 || N < 1) return fail(CATS_EINVAL, "W, F, N must be >= 1%s", "");
    if ((long long)W + F + N > (1 << 24) || F + N > 65535) return fail(CATS_EINVAL, "economy too large%s", "");
    return CATS_OK;
}

struct FieldDesc { const char *name; int kind; int idx; };  // kind: 0 scal,1 c f64,2 k f64,3 cLeff,4 kLeff,5 PA,6 perm,7 Oc
static const FieldDesc kFields[] = {
    {"scal", 0, 0},
    {"c_De", 1, C_De}, {"c_P", 1, C_P}, {"c_Y_prev", 1, C_Y_prev}, {"c_Yd", 1, C_Yd}, {"c_Q", 1, C_Q},
    {"c_A", 1, C_A}, {"c_K", 1, C_K}, {"c_investment", 1, C_investment}, {"c_capital_value", 1, C_capital_value},
    {"c_wages", 1, C_wages}, {"c_interests", 1, C_interests}, {"c_deb", 1, C_deb}, {"c_liquidity", 1, C_liquidity},
    {"c_barYK", 1, C_barYK}, {"c_interest_r", 1, C_interest_r}, {"c_div_income", 1, C_div_income},
    {"k_De", 2, K_De}, {"k_P", 2, K_P}, {"k_Y_prev", 2, K_Y_prev}, {"k_Yd", 2, K_Yd}, {"k_Q", 2, K_Q},
    {"k_A", 2, K_A}, {"k_wages", 2, K_wages}, {"k_interests", 2, K_interests}, {"k_deb", 2, K_deb},
    {"k_liquidity", 2, K_liquidity}, {"k_interest_r", 2, K_interest_r}, {"k_div_income", 2, K_div_income},
    {"k_inv", 2, K_inv}, {"k_Yavail", 2, K_Yavail},
    {"c_Leff", 3, 0}, {"k_Leff", 4, 0}, {"PA", 5, 0}, {"perm", 6, 0}, {"Oc", 7, 0},
};
constexpr int kNumFields = (int)(sizeof(kFields) / sizeof(kFields[0]));

extern "C" {

int cats_abi_version(void) { return CATS_ABI_VERSION; }
const char *cats_last_error(void) { return g_err; }
uint64_t cats_launch_count(void) { return g_launches.load(); }
void cats_debug_phase_cycles(void *d_cycles16) { g_cycles = (long long *)d_cycles16; }

size_t cats_blob_bytes(int W, int F, int N)
{
    if (check_sizes(W, F, N)) return 0;
    return (size_t)make_layout(W, F, N).blob;
}

size_t cats_workset_bytes(int W, int F, int N)
{
    if (check_sizes(W, F, N)) return 0;
    return (size_t)make_layout(W, F, N).workset;
}

const char *cats_field_name(int index) { return (index >= 0 && index < kNumFields) ? kFields[index].name : nullptr; }

int cats_field_info(int W, int F, int N, const char *name, uint64_t *byte_offset, int32_t *count, int32_t *dtype)
{
    if (int rc = check_sizes(W, F, N)) return rc;
    if (!name || !byte_offset || !count || !dtype) return fail(CATS_EINVAL, "null argument%s", "");
    Layout L = make_layout(W, F, N);
    for (int i = 0; i < kNumFields; ++i) {
        if (strcmp(name, kFields[i].name)) continue;
        const FieldDesc &d = kFields[i];
        switch (d.kind) {
        case 0: *byte_offset = L.o_scal; *count = N_SCAL; *dtype = 0; break;
        case 1: *byte_offset = (uint64_t)L.o_c + (uint64_t)d.idx * L.rowF; *count = F; *dtype = 0; break;
        case 2: *byte_offset = (uint64_t)L.o_k + (uint64_t)d.idx * L.rowN; *count = N; *dtype = 0; break;
        case 3: *byte_offset = L.o_Leff; *count = F; *dtype = 1; break;
        case 4: *byte_offset = (uint64_t)L.o_Leff + 4ull * F; *count = N; *dtype = 1; break;
        case 5: *byte_offset = L.o_PA; *count = L.H; *dtype = 0; break;
        case 6: *byte_offset = L.o_perm; *count = L.H; *dtype = 0; break;
        default: *byte_offset = L.o_Oc; *count = W; *dtype = 1; break;
        }
        return CATS_OK;
    }
    return fail(CATS_EINVAL, "unknown field '%s'", name);
}

int cats_workset_field_info(int W, int F, int N, const char *name, uint64_t *byte_offset, int32_t *count,
                            int32_t *dtype)
{
    if (int rc = check_sizes(W, F, N)) return rc;
    if (!name || !byte_offset || !count || !dtype) return fail(CATS_EINVAL, "null argument%s", "");
    Layout L = make_layout(W, F, N);
    struct T { const char *n; uint64_t off; int cnt; int dt; } t[] = {
        {"c_Ld", (uint64_t)L.o_Ld, F, 1}, {"k_Ld", (uint64_t)L.o_Ld + 4ull * F, N, 1},
        {"c_vac", (uint64_t)L.o_vac, F, 1}, {"k_vac", (uint64_t)L.o_vac + 4ull * F, N, 1},
        {"c_Ftot", (uint64_t)L.o_cFtot, F, 0}, {"c_Kdem", (uint64_t)L.o_cKdem, F, 0},
        {"c_Ystock", (uint64_t)L.o_cYstock, F, 0}, {"c_valinv", (uint64_t)L.o_cvalinv, F, 0},
        {"k_Ftot", (uint64_t)L.o_kFtot, N, 0}, {"k_Ystock", (uint64_t)L.o_kYstock, N, 0},
    };
    for (auto &e : t)
        if (!strcmp(name, e.n)) { *byte_offset = e.off; *count = e.cnt; *dtype = e.dt; return CATS_OK; }
    return fail(CATS_EINVAL, "unknown transient field '%s'", name);
}

size_t cats_tape_bytes(int W, int F, int N, int z_c, int z_k, int z_e)
{
    return (size_t)make_tape(W, F, N, z_c, z_k, z_e).bytes;
}

int cats_tape_offsets(int W, int F, int N, int z_c, int z_k, int z_e, uint64_t out11[11])
{
    if (!out11) return fail(CATS_EINVAL, "null argument%s", "");
    TapeOff t = make_tape(W, F, N, z_c, z_k, z_e);
    const uint32_t v[11] = {t.u_price_c, t.u_price_k, t.u_invest, t.fire_key, t.job_order, t.job_cand,
                            t.cap_order, t.cap_cand, t.goods_order, t.goods_cand, t.bytes};
    for (int i = 0; i < 11; ++i) out11[i] = v[i];
    return CATS_OK;
}

static int configure_kernel(const Layout &L, KernelFn fn, int *blocks_per_sm, int *num_sms)
{
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return fail(CATS_ECUDA, "cudaGetDevice: %s", cudaGetErrorString(e));
    int max_optin = 0;
    cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (L.smem > max_optin) return fail(CATS_ESMEM, "economy working set exceeds shared memory per block%s", "");
    e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, L.smem);
    if (e != cudaSuccess) return fail(CATS_ECUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, kThreads, L.smem);
    if (e != cudaSuccess) return fail(CATS_ECUDA, "occupancy query: %s", cudaGetErrorString(e));
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (occ < 1) return fail(CATS_ESMEM, "kernel cannot be resident with this working set%s", "");
    *blocks_per_sm = occ; *num_sms = sms;
    return CATS_OK;
}

int cats_launch_info(int W, int F, int N, int32_t *threads, int32_t *smem_bytes, int32_t *blocks_per_sm,
                     int32_t *num_sms)
{
    if (int rc = check_sizes(W, F, N)) return rc;
    Layout L = make_layout(W, F, N);
    int occ = 0, sms = 0;
    if (int rc = configure_kernel(L, select_kernel(5), &occ, &sms)) return rc;
    if (threads) *threads = kThreads;
    if (smem_bytes) *smem_bytes = L.smem;
    if (blocks_per_sm) *blocks_per_sm = occ;
    if (num_sms) *num_sms = sms;
    return CATS_OK;
}

int cats_init(void *d_blobs, int64_t B, int W, int F, int N, const double *d_params, int64_t params_stride,
              void *stream)
{
    if (int rc = check_sizes(W, F, N)) return rc;
    if (!d_blobs || !d_params || B < 0) return fail(CATS_EINVAL, "cats_init: null pointer or negative B%s", "");
    if (params_stride != 0 && params_stride < CATS_N_PARAMS) return fail(CATS_EINVAL, "params_stride must be 0 or >= 34%s", "");
    if (B == 0) return CATS_OK;
    Layout L = make_layout(W, F, N);
    cats_init_kernel<<<(unsigned)B, 256, 0, (cudaStream_t)stream>>>((unsigned char *)d_blobs, B, L, d_params,
                                                                   params_stride);
    g_launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(CATS_ECUDA, "cats_init launch: %s", cudaGetErrorString(e));
    return CATS_OK;
}

int cats_step(const cats_step_args *s)
{
    if (!s) return fail(CATS_EINVAL, "null args%s", "");
    if (int rc = check_sizes(s->W, s->F, s->N)) return rc;
    if (!s->d_blobs || !s->d_params || s->B < 0) return fail(CATS_EINVAL, "cats_step: null pointer or negative B%s", "");
    if (s->params_stride != 0 && s->params_stride < CATS_N_PARAMS) return fail(CATS_EINVAL, "params_stride must be 0 or >= 34%s", "");
    if (s->n_periods < 1) return fail(CATS_EINVAL, "n_periods must be >= 1%s", "");
    if (s->max_z < 0 || s->max_z > CATS_MAX_Z) return fail(CATS_EINVAL, "max_z must be in [0, 8]%s", "");
    if (s->n_agents < 0 || s->n_agents > s->F) return fail(CATS_EINVAL, "n_agents must be in [0, F]%s", "");
    if (s->n_periods > 1 && s->n_agents > 0 && !(s->flags & CATS_FLAG_BURNIN) && s->d_actions)
        return fail(CATS_EINVAL, "multi-period launches take no actions (use the BURNIN flag)%s", "");
    if (s->stop_phase && s->n_periods != 1) return fail(CATS_EINVAL, "stop_phase needs n_periods == 1%s", "");
    if (((uintptr_t)s->d_blobs & 127u) != 0) return fail(CATS_EINVAL, "d_blobs must be 128-byte aligned%s", "");
    if (s->B == 0) return CATS_OK;

    KArgs a;
    a.L = make_layout(s->W, s->F, s->N);
    a.T = TapeOff{};
    a.blobs = (unsigned char *)s->d_blobs; a.B = s->B;
    a.params = s->d_params; a.params_stride = s->params_stride;
    a.k0 = (uint32_t)s->seed; a.k1 = (uint32_t)(s->seed >> 32);
    a.economy_base = s->economy_base;
    a.n_periods = s->n_periods; a.flags = s->flags; a.n_agents = s->n_agents;
    a.actions = s->d_actions; a.mask = s->d_action_mask;
    a.bankruptcy_reward = s->bankruptcy_reward;
    a.obs = s->d_obs; a.reward = s->d_reward; a.terminated = s->d_terminated; a.status = s->d_status;
    a.stats = s->d_stats; a.tape = s->d_tape; a.stop_phase = s->stop_phase;
    a.debug = (unsigned char *)s->d_debug;
    a.cycles = g_cycles;
    if (s->d_tape) {
        a.T = make_tape(s->W, s->F, s->N, s->tape_z[0], s->tape_z[1], s->tape_z[2]);
    }
    int occ = 0, sms = 0;
    KernelFn fn = select_kernel(s->max_z);
    if (int rc = configure_kernel(a.L, fn, &occ, &sms)) return rc;
    long long grid = (long long)occ * sms;
    if (grid > s->B) grid = s->B;
    fn<<<(unsigned)grid, kThreads, a.L.smem, (cudaStream_t)s->stream>>>(a);
    g_launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(CATS_ECUDA, "cats_step launch: %s", cudaGetErrorString(e));
    return CATS_OK;
}

size_t cats_host_scratch_bytes(int64_t B, int n_agents, int n_periods)
{
    size_t nA = (size_t)(n_agents > 0 ? n_agents : 0), b = (size_t)(B > 0 ? B : 0);
    size_t o = 0;
    o += (b * nA * 2 * 8 + 255) & ~(size_t)255;  // actions
    o += (b * nA + 255) & ~(size_t)255;          // mask
    o += (b * nA * 2 * 8 + 255) & ~(size_t)255;  // obs
    o += (b * nA * 8 + 255) & ~(size_t)255;      // reward
    o += (b + 255) & ~(size_t)255;               // terminated
    o += (b * 4 + 255) & ~(size_t)255;           // status
    o += (b * (size_t)(n_periods > 0 ? n_periods : 1) * N_STATS * 8 + 255) & ~(size_t)255;  // stats
    return o;
}

int cats_step_host(const cats_step_args *h, void *d_scratch)
{
    if (!h || !d_scratch) return fail(CATS_EINVAL, "null args%s", "");
    cudaStream_t st = (cudaStream_t)h->stream;
    const size_t nA = (size_t)(h->n_agents > 0 ? h->n_agents : 0), b = (size_t)(h->B > 0 ? h->B : 0);
    const size_t np = (size_t)(h->n_periods > 0 ? h->n_periods : 1);
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    unsigned char *p = (unsigned char *)d_scratch;
    double *d_act = (double *)p; p += up(b * nA * 16);
    unsigned char *d_msk = p; p += up(b * nA);
    double *d_obs = (double *)p; p += up(b * nA * 16);
    double *d_rew = (double *)p; p += up(b * nA * 8);
    unsigned char *d_term = p; p += up(b);
    unsigned *d_stat = (unsigned *)p; p += up(b * 4);
    double *d_stats = (double *)p;
    cats_step_args s = *h;
    cudaError_t e = cudaSuccess;
    if (h->d_actions && nA) { e = cudaMemcpyAsync(d_act, h->d_actions, b * nA * 16, cudaMemcpyHostToDevice, st); s.d_actions = d_act; }
    if (e == cudaSuccess && h->d_action_mask && nA) { e = cudaMemcpyAsync(d_msk, h->d_action_mask, b * nA, cudaMemcpyHostToDevice, st); s.d_action_mask = d_msk; }
    if (e != cudaSuccess) return fail(CATS_ECUDA, "H2D copy: %s", cudaGetErrorString(e));
    s.d_obs = (h->d_obs && nA) ? d_obs : nullptr;
    s.d_reward = (h->d_reward && nA) ? d_rew : nullptr;
    s.d_terminated = h->d_terminated ? d_term : nullptr;
    s.d_status = h->d_status ? d_stat : nullptr;
    s.d_stats = h->d_stats ? d_stats : nullptr;
    if (int rc = cats_step(&s)) return rc;
    if (s.d_obs) e = cudaMemcpyAsync(h->d_obs, d_obs, b * nA * 16, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && s.d_reward) e = cudaMemcpyAsync(h->d_reward, d_rew, b * nA * 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && s.d_terminated) e = cudaMemcpyAsync(h->d_terminated, d_term, b, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && s.d_status) e = cudaMemcpyAsync(h->d_status, d_stat, b * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && s.d_stats) e = cudaMemcpyAsync(h->d_stats, d_stats, b * np * N_STATS * 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return fail(CATS_ECUDA, "D2H copy / sync: %s", cudaGetErrorString(e));
    return CATS_OK;
}

}  // extern "C"
