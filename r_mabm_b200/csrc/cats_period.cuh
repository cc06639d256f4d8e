// cats_period.cuh -- the batched CATS economy period kernel for sm_100a.
//
// One WARP = one economy; G economies share a block (448 threads at the default size, 28 economies
// resident per SM) only so that they start together and walk the period code in step (see
// cats_period_kernel).  The resident prefix of the economy blob is pulled from HBM into shared memory
// with a single TMA bulk copy (cp.async.bulk + mbarrier), the economy is advanced n_periods periods,
// and the prefix is pushed back with a bulk store; the tail of the blob (employment links, C-firm
// rows, household records) is streamed through L2.  Phase order is the order of Cats.step
// (/root/reference/rmabm/environments/cats.py:355-419); every rule is docs/MODEL.md.
//
// Parallel structure inside a warp
//   - per-firm rules: lane l owns firms l, l+32, l+64, ... (coalesced rows; the fixed-order fp64
//     aggregates "det_sum" become a register partial per lane + one xor butterfly, no memory)
//   - matching markets (job search, capital goods, consumption goods): 32 agents in visiting order
//     are prepared in parallel (visiting order, Philox candidate draws, price sort in registers) and
//     committed in rounds: every lane walks its candidate list against the current seller table,
//     lanes that meet at one seller replay each other's demands in lane order (shuffles), and every
//     lane whose outcome cannot depend on an uncommitted lower lane is committed at once -- also past
//     a seller that sells out (rule R2, see phase_goods_market); the outcome is exactly the
//     sequential visiting-order market of the reference, including the order of every fp64 add.
//   - the warps of a block exchange no data; block barriers only re-align them (once per period and
//     once before the closing phases).
//
// Compiled in three modes (CtxT<MODE>): 0 = every feature, 1 = production (Philox only, no debug
// stop), 2 = production with the default-size layout as compile-time constants.  DESIGN.md 2.2b.
#ifndef CATS_PERIOD_CUH
#define CATS_PERIOD_CUH
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>

#include "../../include/cats_b200.h"
#include "cats_device.cuh"
#include "cats_layout.h"

namespace cats {

struct KArgs {
    Layout L;
    TapeOff T;
    OrderDims od[3];   // visiting-order network dimensions of the job, capital and goods markets
    unsigned char *blobs;
    long long B;
    const double *params;
    long long params_stride;
    uint32_t rk[20];   // Philox round keys: {seed_lo + r * 0x9E3779B9, seed_hi + r * 0xBB67AE85}, r = 0..9
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
    int tape_z[3];
    int stop_phase;
    unsigned char *debug;
    long long *cycles;  // debug: [16] per-phase clock64 totals of block 0
    int slice;          // bytes of shared memory per economy (Layout::smem rounded up to 128)
    int bars;           // 1: the block re-aligns before the accounting phases (0 under stop_phase)
    const unsigned char *active;  // [B] or nullptr: economies with a zero entry sit this launch out
};

// per-warp view of the working set.  Only two base pointers are kept: every section address is
// base + a constant-bank offset, recomputed where it is used (registers are what bounds residency).
// `c.last`: the debug stop (stop_phase) as a value that is the constant 99 where the kernel has no debug stop
template <bool ST> struct LastPhase;
template <> struct LastPhase<false> {
    int v;
    __device__ __forceinline__ operator int() const { return v; }
    __device__ __forceinline__ void operator=(int x) { v = x; }
};
template <> struct LastPhase<true> {
    __device__ __forceinline__ constexpr operator int() const { return 99; }
    __device__ __forceinline__ void operator=(int) {}
};
// MODE 0: every feature (replay tape, debug stop after a phase, debug image of the working set).
// MODE 1: production -- draws come from Philox only, no debug stop, no debug image.
// MODE 2: production for the reference's default configuration -- the layout (W=1000, F=100, N=20) and
//         the candidate-list lengths (z_c=5, z_k=2, z_e=5) are compile-time constants (DefaultLayout) too.
// STRICT: the plain arithmetic of docs/MODEL.md 4.6 (CATS_FLAG_STRICT): purchases and real values divide by the
//         price instead of multiplying by a once-rounded reciprocal, and the demand registered at a
//         consumption-goods seller is an fp64 sum in visiting order instead of an order-free fixed-point one.
template <int MODE, bool STRICT_ = false>
struct CtxT {
    static constexpr bool kStatic = MODE == 2;   // layout resolved at compile time
    static constexpr bool kProd = MODE >= 1;     // Philox only, no debug stop / image
    static constexpr bool kStrict = STRICT_;
    static constexpr bool ST = kStatic;
    __device__ __forceinline__ decltype(auto) layout() const
    {
        if constexpr (ST) return DefaultLayout{}; else return (a->L);
    }
    unsigned char *sm;    // shared-memory working set
    unsigned char *blob;  // this economy's blob in global memory (Oc, C-firm rows, cold K rows, PA, perm live there)
    const KArgs *a;
    int lane;
    long long b;          // economy index inside this launch
    LastPhase<kProd> last;   // last phase id to execute (stop_phase or 99); a constant in the ST instantiation
    DrawsT<!kProd> d;   // production instantiations draw from Philox only
    __device__ __forceinline__ unsigned lt() const { return (1u << lane) - 1u; }   // lanes below this one
    __device__ __forceinline__ double *scal() const { return reinterpret_cast<double *>(sm + layout().o_scal); }
    __device__ __forceinline__ double *par() const { return reinterpret_cast<double *>(sm + layout().s_par); }
    __device__ __forceinline__ int *Leff() const { return reinterpret_cast<int *>(sm + layout().o_Leff); }
    __device__ __forceinline__ int *vac() const { return reinterpret_cast<int *>(sm + layout().s_vac); }   // Ld == Leff + vac
    __device__ __forceinline__ unsigned *emp() const { return reinterpret_cast<unsigned *>(sm + layout().s_emp); }
    __device__ __forceinline__ int *misc() const { return reinterpret_cast<int *>(sm + layout().s_misc); }
    __device__ __forceinline__ OrderKeys *okeys() const { return reinterpret_cast<OrderKeys *>(sm + layout().s_misc + 16); }
    __device__ __forceinline__ volatile double *gk() const { return reinterpret_cast<volatile double *>(sm + layout().s_misc + 56); }
    __device__ __forceinline__ int *Oc() const { return reinterpret_cast<int *>(blob + layout().o_Oc); }          // GLOBAL
    // households: {wealth, permanent income} records, GLOBAL
    __device__ __forceinline__ double2 *hh() const { return reinterpret_cast<double2 *>(blob + layout().o_PA); }
    __device__ __forceinline__ double &PA(int h) const { return hh()[h].x; }
    __device__ __forceinline__ double &perm(int h) const { return hh()[h].y; }
    // candidate-list lengths: constants in the default-configuration instantiation (the host checks them)
    __device__ __forceinline__ int z_c() const { if constexpr (kStatic) return DefaultLayout::kZc; else return (int)par()[P_z_c]; }
    __device__ __forceinline__ int z_k() const { if constexpr (kStatic) return DefaultLayout::kZk; else return (int)par()[P_z_k]; }
    __device__ __forceinline__ int z_e() const { if constexpr (kStatic) return DefaultLayout::kZe; else return (int)par()[P_z_e]; }
    __device__ __forceinline__ double *cf(int field) const { return reinterpret_cast<double *>(blob + layout().o_c + field * layout().rowF); }  // GLOBAL
    // K-firm rows: hot ones (P, Yd, Q) are SHARED, the others GLOBAL; `field` is a compile-time constant
    __device__ __forceinline__ double *kf(int field) const
    {
        return field < NK_HOT ? reinterpret_cast<double *>(sm + layout().o_kh + field * layout().rowN)
                              : reinterpret_cast<double *>(blob + layout().o_kc + (field - NK_HOT) * layout().rowN);
    }
    __device__ __forceinline__ double *at(int off) const { return reinterpret_cast<double *>(sm + off); }
    __device__ __forceinline__ int *iat(int off) const { return reinterpret_cast<int *>(sm + off); }
};

enum { M_STATUS = 0, M_EPISODE = 1 };   // M_EPISODE: the economy's episode counter << 12 (Philox counter word 1, above phase | block)
constexpr int JOB_SORT_MAX = 128;   // job search: up to this many seekers are sorted by position (see phase_job_search)

__device__ __forceinline__ double bfly(double a)
{
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) a = a + __shfl_xor_sync(FULL, a, off);
    return a;
}
__device__ __forceinline__ int bfly_i(int a)
{
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) a += __shfl_xor_sync(FULL, a, off);
    return a;
}

// ---------------------------------------------------------------------------------------------
// a8 firms_get_credit! for one firm (cats.py:376-377)
template <class C>
__device__ __forceinline__ void credit_one(const C &c, double need, double b1, double b2, double &deb,
                                           double A, double &interest_r, double &Ftot, double liquidity,
                                           int &Ld, int Leff, int &vac)
{
    const double theta = c.par()[P_theta], mu = c.par()[P_mu], r_f = c.par()[P_r_f];
    const double wb = c.scal()[S_wb];
    double B = need > 0.0 ? need : 0.0;
    double credit = 0.0;
    if (B > 0.0) {
        double d0 = deb;
        double lev = (d0 + B) / ((A + B) + d0);
        double x = b1 + b2 * lev;
        double pr = 1.0 / (1.0 + det_exp(-x));
        double y = 1.0 + 1.0 / pr;
        double pw = det_exp(y * c.par()[P_log1mtheta]);
        double Xi = (1.0 - pw) / theta;
        double rate = mu * ((1.0 + r_f / theta) / Xi - theta);
        double maxL = (c.par()[P_phi] * c.scal()[S_bank_E] - pr * d0) / pr;
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
// The per-firm loops over C-firms [f_lo, f_hi) and K-firms [k_lo, k_hi): the whole population for the one-warp
// kernel, one warp's 32-aligned share in the multi-warp kernel (which takes the sums from shared memory instead).
template <class C>
__device__ __forceinline__ void firm_decisions_loops(C &c, int f_lo, int f_hi, int k_lo, int k_hi, double &accC, double &accK)
{
    const auto &L = c.layout();
    const int F = L.F, last = c.last;
    const double alpha = c.par()[P_alpha], q_adj = c.par()[P_q_adj], p_adj = c.par()[P_p_adj];
    const double wb = c.scal()[S_wb];
    for (int f = f_lo + c.lane; f < f_hi; f += 32) {
        double *De = c.cf(C_De), *P = c.cf(C_P);
        // every row value this firm needs, loaded before the first store (the rows share one base
        // pointer: the compiler will not move a load over a store on its own)
        const double Yprev = c.cf(C_Y_prev)[f], Yd = c.cf(C_Yd)[f];
        double p = P[f], de = De[f];
        const double K_f = c.cf(C_K)[f], barYK_f = c.cf(C_barYK)[f], liq_f = c.cf(C_liquidity)[f];
        const double deb_f = c.cf(C_deb)[f], ir_f = c.cf(C_interest_r)[f], A_f = c.cf(C_A)[f];
        const double prevDe = de, prevP = p;
        // phase 3: a3 firms_decide_price_quantity! (cats.py:363)
        {
            double stock = Yprev - Yd, avg = c.scal()[S_price];
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
            const double Iprob = c.par()[P_Iprob], barX = c.par()[P_barX], eta = c.par()[P_eta];
            double u = c.d.uniform(PH_INVEST, f);
            if (u < Iprob) {
                double K = K_f;
                double Kdes = barYK_f / barX;
                double depn = ((barX * K) * eta) / Iprob;
                kd = (Kdes - K) + depn;
                if (!(kd > 0.0)) kd = 0.0;
            }
            c.at(L.s_cKdem)[f] = kd;
        }
        // phase 6: a6 _implement_action (cats.py:434-464 / 673-704)
        if (last >= 6 && f < c.a->n_agents && !(c.a->flags & CATS_FLAG_BURNIN) && c.a->actions) {
            const long long ai = c.b * c.a->n_agents + f;
            const bool real = c.a->mask ? (c.a->mask[ai] != 0) : true;
            if (real) {
                double a0 = c.a->actions[2 * ai], a1 = c.a->actions[2 * ai + 1];
                double base = (c.a->flags & CATS_FLAG_AVG_PRICE) ? c.scal()[S_price] : prevP;
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
        // phase 7: a7 firms_decide_labour! (cats.py:373)
        if (last >= 7) {
            double aa = ceil(de / alpha);
            double bb = ceil((K_f * c.par()[P_k]) / alpha);
            int Ld = d2count(aa < bb ? aa : bb);
            int vac = Ld - c.Leff()[f];
            // phase 9: a8 firms_get_credit! (cats.py:376)
            if (last >= 9) {
                double liq = liq_f;
                double need = (wb * (double)Ld + kd * c.scal()[S_price_k]) - liq;
                double deb = deb_f, ir = ir_f, Ftot;
                credit_one(c, need, c.par()[P_b1], c.par()[P_b2], deb, A_f, ir, Ftot, liq, Ld,
                           c.Leff()[f], vac);
                c.cf(C_deb)[f] = deb; c.cf(C_interest_r)[f] = ir; c.at(L.s_cFtot)[f] = Ftot;
                accC = accC + Ftot;
            }
            c.vac()[f] = vac;
        }
        De[f] = de; P[f] = p;
    }
    for (int i = k_lo + c.lane; i < k_hi; i += 32) {
        double *De = c.kf(K_De), *P = c.kf(K_P);
        const double Yprev = c.kf(K_Y_prev)[i], Yd = c.kf(K_Yd)[i], avail = c.kf(K_Yavail)[i];
        double p = P[i], de = De[i];
        // phase 4: a4 firms_decide_price_quantity! for K-firms (cats.py:364)
        if (last >= 4) {
            double stock = avail - Yd, avg = c.scal()[S_price_k];
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
            int vac = Ld - c.Leff()[F + i];
            // phase 10: a8 firms_get_credit! for K-firms (cats.py:377)
            if (last >= 10) {
                double liq = c.kf(K_liquidity)[i];
                double need = wb * (double)Ld - liq;
                double deb = c.kf(K_deb)[i], ir = c.kf(K_interest_r)[i], Ftot;
                credit_one(c, need, c.par()[P_b_k1], c.par()[P_b_k2], deb, c.kf(K_A)[i], ir, Ftot, liq, Ld,
                           c.Leff()[F + i], vac);
                c.kf(K_deb)[i] = deb; c.kf(K_interest_r)[i] = ir; c.at(L.s_kFtot)[i] = Ftot;
                accK = accK + Ftot;
            }
            c.vac()[F + i] = vac;
        }
    }
}

// a3-a8: every per-firm decision of the period in one pass (phases 3..10).  No cross-firm
// dependence: the bank rations against its start-of-period equity.
template <class C>
__device__ void phase_firm_decisions(C &c)
{
    const auto &L = c.layout();
    const int last = c.last;
    double accC = 0.0, accK = 0.0;  // det_sum partials of new credit
    firm_decisions_loops(c, 0, L.F, 0, L.N, accC, accK);
    // bank loan stock: += det_sum(new credit), C-firms then K-firms
    if (last >= 9) {
        accC = bfly(accC); accK = bfly(accK);
        if (c.lane == 0) {
            double loans = c.scal()[S_bank_loans] + accC;
            if (last >= 10) loans = loans + accK;
            c.scal()[S_bank_loans] = loans;
        }
    }
    __syncwarp();
}

// a9 firms_fire_workers! (cats.py:379-380).  A firm with s = Leff - Ld > 0 dismisses the s
// employees with the smallest (fire_key, worker index).  The employees of the over-staffed firms
// are bucketed by firm (count, prefix, scatter) and every pooled worker ranks itself inside its
// firm's bucket; rank < s means dismissed.  The result does not depend on the bucket order.
// One batch of phase_fire: consecutive firms [f0, f1) whose pooled employees fit a pool of `cap` entries;
// writes each firm's bucket start into ic[f] (count | start << 16) and returns the entries pooled.  One warp.
template <class C>
__device__ __forceinline__ int fire_batch(C &c, int *ic, int f0, int limit, int cap, int &f1_out)
{
    const int lane = c.lane;
    int base = 0, f1 = f0;
    bool stop = false;
    while (!stop && f1 < limit) {
        const int f = f1 + lane;
        const int n = f < limit ? ic[f] : 0;
        int incl = n;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) { int t = __shfl_up_sync(FULL, incl, off); if (lane >= off) incl += t; }
        const unsigned fm = __ballot_sync(FULL, base + incl <= cap);
        const int nfit = fm == FULL ? 32 : __ffs(~fm) - 1;
        if (f < limit && lane < nfit) ic[f] = n | ((base + incl - n) << 16);
        const int add = __shfl_sync(FULL, incl, nfit > 0 ? nfit - 1 : 0);
        if (nfit > 0) base += add;
        f1 += nfit;
        if (nfit < 32) stop = true;
    }
    if (f1 > limit) f1 = limit;
    f1_out = f1;
    return base;
}

// the slow path of phase_fire for ONE firm whose employees do not fit the pool: s arg-min selections over all
// workers.  One warp.
template <class C>
__device__ __forceinline__ void fire_oversize(C &c, int f)
{
    const int W = c.layout().W, lane = c.lane;
    const int s = -c.vac()[f];
#ifdef CATS_EMU_COUNT
    if (lane == 0) simt::g_count[2]++;
#endif
    for (int r = 0; r < s; ++r) {
        unsigned bk = 0xffffffffu; int bw = 0x7fffffff;
        for (int w = lane; w < W; w += 32) {
            if (c.Oc()[w] == f) {
                unsigned key = c.d.fire_key(w);
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
        if (lane == 0) c.Oc()[bw] = -1;
        __syncwarp();
    }
    if (lane == 0) { c.Leff()[f] = c.Leff()[f] + c.vac()[f]; c.vac()[f] = 0; }
    __syncwarp();
}

template <class C>
__device__ void phase_fire(C &c, int limit)
{
    const auto &L = c.layout();
    const int W = L.W, lane = c.lane;
    int *ic = c.iat(L.s_ic);    // per firm: pooled head-count (low 16 bits) | pool cursor (high 16 bits)
    // pooled head-count of an over-staffed firm == its Leff (#{w : Oc[w] == f} == Leff[f] is an invariant)
    bool over = false;
    for (int f = lane; f < limit; f += 32) { const bool o = c.vac()[f] < 0; over |= o; ic[f] = o ? c.Leff()[f] : 0; }
    if (!__any_sync(FULL, over)) return;
    __syncwarp();
    const int cap = L.U_bytes / 8;
    int *pool_w = c.iat(L.s_U);
    unsigned *pool_k = reinterpret_cast<unsigned *>(c.sm + L.s_U) + cap;
    int f0 = 0;
    while (f0 < limit) {
        // batch [f0, f1): consecutive firms whose pooled employees fit the pool
        int f1;
        const int total = fire_batch(c, ic, f0, limit, cap, f1);
        __syncwarp();
        if (f1 == f0) {
            // one firm larger than the pool: s arg-min selections over all workers (slow, rare)
            fire_oversize(c, f0);
            f0 = f0 + 1;
            continue;
        }
        if (total > 0) {
            // the pass over the employment links: four workers per lane and load (the row is padded to 16 bytes), two
            // loads in flight -- the pass waits on L2 a quarter as often as a link per lane would.  A pool entry is
            // worker | employer << 16, so the ranking below need not read the link again.
            const int4 *Oc4 = reinterpret_cast<const int4 *>(c.Oc());
            const int nq = (W + 3) >> 2;
            auto pool = [&](int f, int w) {
                if (f >= f0 && f < f1 && w < W && c.vac()[f] < 0) {
                    const int slot = atomicAdd(&ic[f], 0x10000) >> 16;
                    pool_w[slot] = w | (f << 16);
                }
            };
            const int4 none = make_int4(-1, -1, -1, -1);
            int4 q0 = lane < nq ? Oc4[lane] : none, q1 = lane + 32 < nq ? Oc4[lane + 32] : none;
#pragma unroll 1
            for (int q = lane; q < nq; q += 32) {
                const int4 v = q0;
                q0 = q1;
                q1 = q + 64 < nq ? Oc4[q + 64] : none;
                const int w = q << 2;
                pool(v.x, w); pool(v.y, w + 1); pool(v.z, w + 2); pool(v.w, w + 3);
            }
            __syncwarp();
            // the keys in a dense pass over the pool: a draw per POOLED worker (an eighth of the workers at the default
            // size), not a Philox block in every iteration of the pass over all of them
            for (int i = lane; i < total; i += 32) pool_k[i] = c.d.fire_key(pool_w[i] & 0xffff);
            __syncwarp();
            for (int i = lane; i < total; i += 32) {
                const int e = pool_w[i]; const unsigned key = pool_k[i];
                const int w = e & 0xffff, f = e >> 16;
                const int v = ic[f];
                const int end = v >> 16, beg = end - (v & 0xffff);
                int rank = 0;
                for (int j = beg; j < end; ++j) {
                    const unsigned kj = pool_k[j]; const int wj = pool_w[j] & 0xffff;
                    rank += (kj < key || (kj == key && wj < w)) ? 1 : 0;
                }
                if (rank < -c.vac()[f]) c.Oc()[w] = -1;
            }
            __syncwarp();
        }
        for (int f = f0 + lane; f < f1; f += 32)
            if (c.vac()[f] < 0) { c.Leff()[f] = c.Leff()[f] + c.vac()[f]; c.vac()[f] = 0; }
        __syncwarp();
        f0 = f1;
    }
}

// a10 workers_search_job! (cats.py:382).  The unemployed are compacted in visiting order, then
// taken 32 at a time.  Every lane picks the first of its z_e candidates that still has a vacancy;
// the longest prefix of lanes whose picks cannot have been pre-empted by a lower lane is committed
// at once (lane j is safe iff fewer than vac[f] lower lanes chose the same f), the rest retry
// against the updated vacancies.  Outcome == the sequential visiting-order rule.
template <int ZM, class C>
__device__ __forceinline__ void job_match(C &c, const unsigned short *seekers, int total, bool listed);

template <int ZM, class C>
__device__ void phase_job_search(C &c)
{
    const auto &L = c.layout();
    const int W = L.W, nf = L.F + L.N, lane = c.lane;
    const int z = c.z_e() < nf ? c.z_e() : nf;
    OrderT<C::kStatic ? DefaultLayout::kW : 0, !C::kProd> ord; ord.init(c.d, MKT_JOB, c.a->od[MKT_JOB], c.okeys(), lane);
    unsigned short *seekers = reinterpret_cast<unsigned short *>(c.sm + L.s_U);
    int total = 0;
    // Few seekers (the usual case: unemployment of a few percent): list them with one coalesced pass
    // over the employment links, ask the visiting order for each one's POSITION (the inverse
    // bijection) and sort the handful by position -- instead of evaluating the order at all W
    // positions and gathering W links in that order.  Same visiting order, by construction.
    bool listed = false;
    const int cap3 = L.U_bytes / 6;   // three uint16 arrays: seekers in visiting order | by index | positions
    if (!ord.from_tape()) {
        unsigned short *byidx = seekers + cap3, *posv = seekers + 2 * cap3;
        int U = 0;
        for (int base = 0; base < W; base += 128) {
            int oc[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) { const int w = base + 32 * q + lane; oc[q] = w < W ? c.Oc()[w] : 0; }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const bool un = oc[q] < 0;
                const unsigned bal = __ballot_sync(FULL, un);
                const int at = U + __popc(bal & c.lt());
                if (un && at < cap3) byidx[at] = (unsigned short)(base + 32 * q + lane);
                if (lane == 0 && base + 32 * q < W) c.emp()[(base >> 5) + q] = ~bal;   // employed bits (hires are added below)
                U += __popc(bal);
            }
        }
        // Rank by position.  With room for a W-bit bitmap of the positions and its word prefix (the firing / capital
        // scratch, idle here) a seeker's rank is the number of bits below its own: O(U + W / 32), any U that fits the
        // list; otherwise all pairs, up to JOB_SORT_MAX seekers.
        const int nwords = (W + 31) >> 5;
        const bool bitmap_fits = 2 * nwords <= L.F + L.N;
        if (U <= cap3 && (bitmap_fits || U <= JOB_SORT_MAX)) {
            __syncwarp();
            if (bitmap_fits) {
                unsigned *bitmap = reinterpret_cast<unsigned *>(c.iat(L.s_ic));
                int *prefix = c.iat(L.s_ic) + nwords;
                for (int i = lane; i < nwords; i += 32) bitmap[i] = 0u;
                __syncwarp();
                for (int i = lane; i < U; i += 32) {
                    const int pos = ord.position_of((int)byidx[i]);
                    posv[i] = (unsigned short)pos;
                    atomicOr(&bitmap[pos >> 5], 1u << (pos & 31));
                }
                __syncwarp();
                int run = 0;
                for (int base = 0; base < nwords; base += 32) {
                    const int i = base + lane;
                    const int n = i < nwords ? __popc(bitmap[i]) : 0;
                    int incl = n;
#pragma unroll
                    for (int off = 1; off < 32; off <<= 1) { const int t = __shfl_up_sync(FULL, incl, off); if (lane >= off) incl += t; }
                    if (i < nwords) prefix[i] = run + incl - n;
                    run += __shfl_sync(FULL, incl, 31);
                }
                __syncwarp();
                for (int i = lane; i < U; i += 32) {
                    const int pos = (int)posv[i];
                    seekers[prefix[pos >> 5] + __popc(bitmap[pos >> 5] & ((1u << (pos & 31)) - 1u))] = byidx[i];
                }
            } else {
                for (int i = lane; i < U; i += 32) posv[i] = (unsigned short)ord.position_of((int)byidx[i]);
                __syncwarp();
                for (int i = lane; i < U; i += 32) {
                    const unsigned mine = posv[i];
                    int r = 0;
                    for (int j = 0; j < U; ++j) r += posv[j] < mine ? 1 : 0;   // positions are distinct
                    seekers[r] = byidx[i];
                }
            }
            total = U;
            listed = true;
        }
        __syncwarp();
    }
    for (int base = 0; !listed && base < W; base += 128) {
        int w[4]; int oc[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int pos = base + 32 * q + lane;
            w[q] = pos < W ? ord.at(pos) : -1;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) oc[q] = w[q] >= 0 ? c.Oc()[w[q]] : 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const bool un = oc[q] < 0;
            const unsigned bal = __ballot_sync(FULL, un);
            if (un) seekers[total + __popc(bal & c.lt())] = (unsigned short)w[q];
            total += __popc(bal);
        }
    }
    __syncwarp();
    job_match<ZM>(c, seekers, total, listed);
    if (!listed) {   // the bits the goods market reads instead of gathering employment links
        for (int base = 0; base < W; base += 32) {
            const int w = base + lane;
            const unsigned bal = __ballot_sync(FULL, w < W && c.Oc()[w] >= 0);
            if (lane == 0) c.emp()[base >> 5] = bal;
        }
    }
    __syncwarp();
}

// The matching rounds of the job search over `total` seekers listed in visiting order.  One warp.
template <int ZM, class C>
__device__ __forceinline__ void job_match(C &c, const unsigned short *seekers, int total, bool listed)
{
    const auto &L = c.layout();
    const int nf = L.F + L.N, lane = c.lane;
    const int z = c.z_e() < nf ? c.z_e() : nf;
    const unsigned lt = c.lt();
    for (int base = 0; base < total; base += 32) {
        const int idx = base + lane;
        int w = -1;
        int cand[ZM];
        if (idx < total) { w = (int)seekers[idx]; c.d.template candidates<ZM>(PH_JOB, w, nf, z, cand); }
        unsigned pending = __ballot_sync(FULL, idx < total);
        while (pending) {
            const bool mine = (pending >> lane) & 1u;
            int pick = -1;
            if (mine) {
#pragma unroll
                for (int k = 0; k < ZM; ++k)
                    if (k < z && pick < 0 && c.vac()[cand[k]] > 0) pick = cand[k];
            }
            const bool has = mine && pick >= 0;
            const unsigned peers = __match_any_sync(FULL, has ? pick : -1);   // (one dummy value: see the goods market)
            const bool ok = !has || __popc(peers & lt) < c.vac()[pick];
            const unsigned badm = __ballot_sync(FULL, mine && !ok);
            const unsigned commit = badm ? (pending & ((1u << (__ffs(badm) - 1)) - 1u)) : pending;
            __syncwarp();
            if (has && ((commit >> lane) & 1u)) {
                c.Oc()[w] = pick;
                if (listed) atomicOr(&c.emp()[w >> 5], 1u << (w & 31));
                const unsigned grp = peers & commit;
                if ((grp & lt) == 0) { int n = __popc(grp); c.vac()[pick] -= n; c.Leff()[pick] += n; }
            }
            pending &= ~commit;
            __syncwarp();
        }
    }
}

// One consumption-goods seller as the market sees it.  The SoA rows (P, Q, stock) are packed into
// this 32-byte record when production is decided and unpacked when the market closes.
// cls: price class = number of strictly cheaper sellers.  Sorting a household's candidates by
// (cls, draw index) IS the stable sort by price of MODEL.md 5.20, done on small integers.
// The demand expressed at a seller (2^-40 fixed point, order-independent: MODEL.md 4.4) is
// accumulated in the seller's Yd slot of the blob with fire-and-forget 64-bit reductions (RED at
// L2): nobody reads it before the market closes, so no lane ever waits for it.
// rp = 1 / p: a purchase is b * rp (MODEL.md 4.5).  Sales are what is missing from the stock when
// the market closes (Q = Y - stock), so the record carries no running total.
// Laid out as three arrays, not one record per seller: 32 lanes reading the class word of 32 random
// 32-byte records would all land in 4 of the 32 shared-memory banks.
struct Mkt {
    double2 *yr;     // [F] stock, 1 / price
    double *p;       // [F] price
    unsigned *cls;   // [F] price class
    __device__ __forceinline__ Mkt(unsigned char *base, int F)
        : yr(reinterpret_cast<double2 *>(base)), p(reinterpret_cast<double *>(base + 16 * (size_t)F)),
          cls(reinterpret_cast<unsigned *>(base + 24 * (size_t)F)) {}
};

// a11 firms_produce! (cats.py:384-385)
template <class C>
__device__ __forceinline__ void produce_c_loop(C &c, int f_lo, int f_hi)
{
    const auto &L = c.layout();
    const int F = L.F, last = c.last;
    const double alpha = c.par()[P_alpha], k = c.par()[P_k], eta = c.par()[P_eta], theta = c.par()[P_theta];
    const double wb = c.scal()[S_wb], pk = c.scal()[S_price_k];
    Mkt mkt(c.sm + L.s_U, F);
    for (int f = f_lo + c.lane; f < f_hi; f += 32) {
        // loads first (see phase_firm_decisions)
        const double K_f = c.cf(C_K)[f], De = c.cf(C_De)[f], deb = c.cf(C_deb)[f], ir_f = c.cf(C_interest_r)[f];
        double P = c.cf(C_P)[f];
        double Leff = (double)c.Leff()[f];
        double cap_l = Leff * alpha, cap_k = K_f * k;
        double Yp = cap_l < cap_k ? cap_l : cap_k;
        double Y = De < Yp ? De : Yp;
        c.cf(C_Y_prev)[f] = Y;
        double interests = ir_f * deb;
        double wages = wb * Leff;
        c.cf(C_interests)[f] = interests; c.cf(C_wages)[f] = wages;
        double Pl = 0.0;
        if (Y > 0.0) Pl = (1.01 * (((wages + interests) + ((Y / k) * eta) * pk) + theta * deb)) / Y;
        if (P < Pl) { P = Pl; c.cf(C_P)[f] = P; }
        mkt.yr[f] = make_double2(Y, C::kStrict ? 0.0 : 1.0 / P); mkt.p[f] = P; mkt.cls[f] = 0u;   // strict: .y is the Yd accumulator
        c.cf(C_investment)[f] = 0.0; c.at(L.s_cvalinv)[f] = 0.0;
        c.cf(C_Yd)[f] = 0.0;                     // == the all-zero demand accumulator of the goods market
        if (last < 20) c.cf(C_Q)[f] = 0.0;       // else written when the goods market closes
    }
}

template <class C>
__device__ __forceinline__ void produce_k_loop(C &c, int k_lo, int k_hi)
{
    const auto &L = c.layout();
    const int F = L.F;
    const double alpha = c.par()[P_alpha], theta = c.par()[P_theta];
    const double wb = c.scal()[S_wb];
    for (int i = k_lo + c.lane; i < k_hi; i += 32) {
        double Leff = (double)c.Leff()[F + i];
        double want = c.kf(K_De)[i];
        double cap_l = Leff * alpha;
        double Y = want < cap_l ? want : cap_l;
        c.kf(K_Y_prev)[i] = Y;
        double avail = Y + c.kf(K_inv)[i];
        c.at(L.s_kYstock)[i] = avail; c.kf(K_Yavail)[i] = avail;
        double deb = c.kf(K_deb)[i];
        double interests = c.kf(K_interest_r)[i] * deb;
        double wages = wb * Leff;
        c.kf(K_interests)[i] = interests; c.kf(K_wages)[i] = wages;
        double Pl = 0.0;
        if (Y > 0.0) Pl = (1.01 * ((wages + interests) + theta * deb)) / Y;
        double P = c.kf(K_P)[i];
        if (P < Pl) { P = Pl; c.kf(K_P)[i] = P; }
        // the Q slot carries 1 / P through the capital-goods market (MODEL.md 4.5); K-firm
        // sales are Yavail - stock, written by the accounting phase
        c.kf(K_Q)[i] = 1.0 / P; c.kf(K_Yd)[i] = 0.0;
    }
}

template <class C>
__device__ void phase_produce(C &c)
{
    const auto &L = c.layout();
    const int F = L.F, N = L.N, last = c.last;
    Mkt mkt(c.sm + L.s_U, F);
    produce_c_loop(c, 0, F);
    if (last >= 20) {
        // price classes of the sellers (prices are final for the period from here on)
        __syncwarp();
        for (int f0 = c.lane; f0 < F; f0 += 128) {
            const int f1 = f0 + 32, f2 = f0 + 64, f3 = f0 + 96;
            const double pa = mkt.p[f0], pb = f1 < F ? mkt.p[f1] : 0.0, pc = f2 < F ? mkt.p[f2] : 0.0,
                         pd = f3 < F ? mkt.p[f3] : 0.0;
            unsigned na = 0u, nb = 0u, nc = 0u, nd = 0u;
#pragma unroll 2
            for (int g = 0; g < F; ++g) {
                const double pg = mkt.p[g];
                na += pg < pa ? 1u : 0u; nb += pg < pb ? 1u : 0u; nc += pg < pc ? 1u : 0u; nd += pg < pd ? 1u : 0u;
            }
            mkt.cls[f0] = na;
            if (f1 < F) mkt.cls[f1] = nb;
            if (f2 < F) mkt.cls[f2] = nc;
            if (f3 < F) mkt.cls[f3] = nd;
        }
    }
    if (last >= 15) produce_k_loop(c, 0, N);
    __syncwarp();
}

// a12 capital_goods_market! (cats.py:387).  The investing C-firms are compacted in visiting order;
// 32 of them at a time draw and price-sort their z_k suppliers in parallel and then trade one after
// the other (there are ~F * Iprob buyers per period: a short chain).
// a12, one buyer: down its price-sorted candidates while it has funds and wants capital.  kd: capital still wanted,
// cb: funds, lq: liquidity; iv / vi: capital bought and its value.  Sequential per supplier (see capital_go).
template <int ZM, bool STRICT>
__device__ __forceinline__ void capital_trade(int z, const int (&cand)[ZM], const double (&pr)[ZM], const double *RPk, double *Ydk,
                                              double *Ysk, double &kd, double &cb, double &lq, double &iv, double &vi)
{
#pragma unroll
    for (int k = 0; k < ZM; ++k) {
        if (k < z && cb > 0.0 && kd > 0.0) {
            const int best = cand[k];
            const double pk = pr[k], rpk = RPk[best];
            const double afford = STRICT ? cb / pk : cb * rpk;
            const double want = afford < kd ? afford : kd;
            Ydk[best] = Ydk[best] + want;
            const double ys = Ysk[best];
            if (ys > 0.0) {
                const double full = kd * pk;
                const double spend = cb < full ? cb : full;
                const double q = STRICT ? spend / pk : spend * rpk;
                if (ys > q) {
                    Ysk[best] = ys - q;
                    kd = kd - q; cb = cb - spend;
                    lq = lq - spend; iv = iv + q; vi = vi + spend;
                } else {
                    const double cost = ys * pk;
                    kd = kd - ys; cb = cb - cost;
                    lq = lq - cost;
                    iv = iv + ys; vi = vi + cost;
                    Ysk[best] = 0.0;
                }
            }
        }
    }
}
// Buyers trade in visiting order -- but only buyers that share a supplier have to.  Every lane marks its candidates in
// a 32-bit set (supplier mod 32: a false match only delays a lane); a pending lane (bit in m) may trade as soon as no
// pending lane BELOW it has a candidate in common.  The lowest pending lane always may, so rounds end; each round takes
// every buyer whose suppliers are its own for now: ~6 rounds for the ~25 buyers of a default-size period.  Warp collective.
// SetT: unsigned (supplier mod 32; exact up to 32 suppliers: the default configuration has 20) or unsigned long long
// (mod 64) for the instantiations that serve any size.
template <class SetT, int ZM>
__device__ __forceinline__ SetT capital_set(const int (&cand)[ZM], int z)
{
    SetT cset = 0;
#pragma unroll
    for (int k = 0; k < ZM; ++k) if (k < z) cset |= (SetT)1 << (cand[k] & (int)(8 * sizeof(SetT) - 1));
    return cset;
}
template <class SetT>
__device__ __forceinline__ bool capital_go(unsigned m, SetT cset, int lane)
{
    const bool pend = (m >> lane) & 1u;
    SetT below = pend ? cset : (SetT)0;   // becomes the union over the pending lanes up to this one ...
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) { const SetT t = __shfl_up_sync(FULL, below, off); if (lane >= off) below |= t; }
    below = __shfl_up_sync(FULL, below, 1);   // ... and now strictly below it
    if (lane == 0) below = 0;
    return pend && (cset & below) == 0;
}
template <class C> using CapitalSet = std::conditional_t<C::kStatic, unsigned, unsigned long long>;

template <int ZM, class C>
__device__ void phase_capital_market(C &c)
{
    const auto &L = c.layout();
    const int F = L.F, N = L.N, lane = c.lane;
    const int z = c.z_k() < N ? c.z_k() : N;
    OrderT<C::kStatic ? DefaultLayout::kF : 0, !C::kProd> ord; ord.init(c.d, MKT_CAP, c.a->od[MKT_CAP], c.okeys(), lane);
    double *Pk = c.kf(K_P), *Ydk = c.kf(K_Yd), *Ysk = c.at(L.s_kYstock);
    const double *RPk = c.kf(K_Q);   // 1 / P, parked there by phase_produce
    double *Kdem = c.at(L.s_cKdem), *Ftot = c.at(L.s_cFtot), *liq = c.cf(C_liquidity);
    double *inv = c.cf(C_investment), *valinv = c.at(L.s_cvalinv), *wages = c.cf(C_wages);
    int *buyers = c.iat(L.s_ic);
    int total = 0;
    for (int base = 0; base < F; base += 32) {
        const int pos = base + lane;
        int f = -1; bool act = false;
        if (pos < F) { f = ord.at(pos); act = Kdem[f] > 0.0; }
        const unsigned bal = __ballot_sync(FULL, act);
        if (act) buyers[total + __popc(bal & c.lt())] = f;
        total += __popc(bal);
    }
    __syncwarp();
    for (int base = 0; base < total; base += 32) {
        const int idx = base + lane;
        const bool act = idx < total;
        int f = 0; double kd = 0.0, cb = 0.0, lq = 0.0;
        int cand[ZM]; double pr[ZM];
#pragma unroll
        for (int k = 0; k < ZM; ++k) { cand[k] = 0; pr[k] = __longlong_as_double(0x7ff0000000000000LL); }
        if (act) {
            f = buyers[idx]; kd = Kdem[f]; lq = liq[f];
            cb = (Ftot[f] + lq) - wages[f];
            c.d.template candidates<ZM>(PH_CAP, f, N, z, cand);
#pragma unroll
            for (int k = 0; k < ZM; ++k) if (k < z) pr[k] = Pk[cand[k]];
            sort_by_price<ZM>(cand, pr, z);
        }
        unsigned m = __ballot_sync(FULL, act);
        using SetT = CapitalSet<C>;
        const SetT cset = act ? capital_set<SetT, ZM>(cand, z) : (SetT)0;
        while (m) {
            const bool go = capital_go(m, cset, lane);
            m &= ~__ballot_sync(FULL, go);
            if (go) {
                double iv = 0.0, vi = 0.0;   // produce zeroed investment and its value this period
                capital_trade<ZM, C::kStrict>(z, cand, pr, RPk, Ydk, Ysk, kd, cb, lq, iv, vi);
                liq[f] = lq; inv[f] = iv; valinv[f] = vi; Kdem[f] = kd;
            }
            __syncwarp();
        }
    }
    __syncwarp();
}

// a13 gov_adjusts_subsidy! (cats.py:389) and the government side of a14 workers_get_paid!
// (cats.py:391): subsidy level, wage tax and subsidy totals.  The employed head-count is the sum
// of Leff (== #{Oc >= 0}: invariant kept by firing, job search and bankruptcy).
template <class C>
__device__ void phase_gov_income(C &c)
{
    const auto &L = c.layout();
    int n = 0;
    for (int i = c.lane; i < L.F + L.N; i += 32) n += c.Leff()[i];
    n = bfly_i(n);
    if (c.lane == 0) {
        double sub = c.par()[P_subsidy] * c.scal()[S_wb];
        if (c.par()[P_maastricht] > 0.0 && c.scal()[S_deficitGDP] > c.par()[P_target_deficit]) {
            double scale = 1.0 - c.par()[P_maastricht] * (c.scal()[S_deficitGDP] - c.par()[P_target_deficit]);
            if (!(scale > 0.0)) scale = 0.0;
            sub = sub * scale;
        }
        c.scal()[S_gov_subsidy_level] = sub;
        if (c.last >= 18) {
            const double wb = c.scal()[S_wb], tax = c.par()[P_tax_rate];
            c.scal()[S_gov_TA] = c.scal()[S_gov_TA] + (wb * tax) * (double)n;
            c.scal()[S_gov_G] = c.scal()[S_gov_G] + sub * (double)(L.W - n);
        }
    }
    __syncwarp();
}

// debug path only (stop_phase 18 / 19): the household rules without the market
template <class C>
__device__ void phase_households_only(C &c)
{
    const auto &L = c.layout();
    const double wb = c.scal()[S_wb], sub = c.scal()[S_gov_subsidy_level], price = c.scal()[S_price];
    const double xi = c.par()[P_xi], chi = c.par()[P_chi];
    for (int h = c.lane; h < L.H; h += 32) {
        double pa = c.PA(h), pm = c.perm(h);
        double inc;
        if (h < L.W) { inc = c.Oc()[h] >= 0 ? wb * (1.0 - c.par()[P_tax_rate]) : sub; pa = pa + inc; }
        else inc = h < L.W + L.F ? c.cf(C_div_income)[h - L.W] : c.kf(K_div_income)[h - L.W - L.F];
        if (c.last >= 19) {
            const double rpavg = 1.0 / price;
            double ir = C::kStrict ? inc / price : inc * rpavg;
            pm = pm * xi + (1.0 - xi) * ir;
            double target = pm + (C::kStrict ? (chi * pa) / price : (chi * pa) * rpavg);
            double b = target * price;
            if (b > pa) b = pa;
            if (!(b > 0.0)) b = 0.0;
            pa = pa - b;
        }
        c.PA(h) = pa; c.perm(h) = pm;
    }
    __syncwarp();
}

// A household's sorted candidate list, consumed front to back by the goods-market walk: (seller + 1)
// per slot, 0 = end of list.  Generic: 16 bits per slot in two 64-bit words (F <= 16383).  SMALL (the
// seller count is a compile-time constant <= 254): 8 bits per slot in two 32-bit words -- half the
// registers, and a pop is a mask, a funnel shift and a shift.
template <bool SMALL> struct CandList;
template <> struct CandList<false> {
    unsigned long long w0 = 0ull, w1 = 0ull;
    __device__ __forceinline__ void set(int k, unsigned seller)
    {
        if (k < 4) w0 |= (unsigned long long)(seller + 1u) << (16 * k);
        else w1 |= (unsigned long long)(seller + 1u) << (16 * (k - 4));
    }
    __device__ __forceinline__ bool pop(int &seller)
    {
        const unsigned e = (unsigned)(w0 & 0xffffull);
        if (e == 0u) return false;
        seller = (int)e - 1;
        w0 = (w0 >> 16) | (w1 << 48); w1 >>= 16;
        return true;
    }
    // lane `src`'s list (warp collective)
    __device__ __forceinline__ CandList from_lane(int src) const
    {
        CandList r;
        r.w0 = __shfl_sync(FULL, w0, src); r.w1 = __shfl_sync(FULL, w1, src);
        return r;
    }
    // is `seller` still on the list?  (a zero 16-bit slot of list ^ splat(seller + 1); empty slots are 0 and never match)
    __device__ __forceinline__ bool contains(int seller) const
    {
        const unsigned long long x = (unsigned long long)(unsigned)(seller + 1) * 0x0001000100010001ull;
        const unsigned long long a = w0 ^ x, b = w1 ^ x;
        return ((((a - 0x0001000100010001ull) & ~a) | ((b - 0x0001000100010001ull) & ~b)) & 0x8000800080008000ull) != 0ull;
    }
};
template <> struct CandList<true> {
    unsigned w0 = 0u, w1 = 0u;
    __device__ __forceinline__ void set(int k, unsigned seller)
    {
        if (k < 4) w0 |= (seller + 1u) << (8 * k);
        else w1 |= (seller + 1u) << (8 * (k - 4));
    }
    __device__ __forceinline__ bool pop(int &seller)
    {
        const unsigned e = w0 & 0xffu;
        if (e == 0u) return false;
        seller = (int)e - 1;
        w0 = __funnelshift_r(w0, w1, 8); w1 >>= 8;
        return true;
    }
    __device__ __forceinline__ CandList from_lane(int src) const
    {
        CandList r;
        r.w0 = __shfl_sync(FULL, w0, src); r.w1 = __shfl_sync(FULL, w1, src);
        return r;
    }
    __device__ __forceinline__ bool contains(int seller) const
    {
        const unsigned x = (unsigned)(seller + 1) * 0x01010101u;
        const unsigned a = w0 ^ x, b = w1 ^ x;
        return ((((a - 0x01010101u) & ~a) | ((b - 0x01010101u) & ~b)) & 0x80808080u) != 0u;
    }
};

// a14 + a15 + a16: workers_get_paid!, households_find_cons_budget!, consumption_goods_market!
// (cats.py:391-394), fused.  Households are visited in the market's random order, 32 at a time.
//   prepare (lane-parallel): gather wealth / permanent income / employment link (prefetched one
//       chunk ahead), income and budget, z_c Philox candidates, price sort.
//   commit (rounds): every "dirty" lane walks down its sorted candidates: at each seller it
//       registers its demand b / P (fixed-point atomic add: order-free) and stops at the first one
//       with stock, its terminal.  Lanes with the same terminal replay the purchases of the lower
//       lanes in lane order (shuffles), so each knows the stock it will find.  A lane that exhausts its
//       seller (it takes what is left and keeps the rest of its budget) or finds it sold out walks on
//       in the next round -- to a seller still on its list.  Rule R2: a lane commits unless its terminal
//       is on the remaining list of such a lane BELOW it (then the stock it replayed is not final;
//       if it is a full buyer its own list endangers the lanes above it in turn: closed by iteration
//       over the uncertain lanes, two shuffles each).  Round 1 stopped every round at the first
//       sell-out: ~100 rounds per period at the default size; R2 needs ~67 (scripts/round_model.py).
//       Stock, sales and budgets follow the visiting order exactly.
enum { G_net, G_sub, G_ir_emp, G_ir_un, G_xi, G_omxi, G_chi, G_price, G_rpavg };   // goods-market constants (c.gk())

// the market constants: computed once by ONE thread, parked in shared memory and re-read per chunk
template <class C>
__device__ __forceinline__ void goods_constants(C &c)
{
    volatile double *gk = c.gk();
    const double wb = c.scal()[S_wb], sub = c.scal()[S_gov_subsidy_level], price = c.scal()[S_price];
    const double xi = c.par()[P_xi];
    const double net = wb * (1.0 - c.par()[P_tax_rate]);
    const double rp = 1.0 / price;
    gk[G_net] = net; gk[G_sub] = sub;
    gk[G_ir_emp] = C::kStrict ? net / price : net * rp; gk[G_ir_un] = C::kStrict ? sub / price : sub * rp;
    gk[G_xi] = xi; gk[G_omxi] = 1.0 - xi; gk[G_chi] = c.par()[P_chi]; gk[G_price] = price; gk[G_rpavg] = rp;
}

template <class C> using GoodsList = CandList<C::kStatic && DefaultLayout::kF <= 254>;

// prepare (one lane = one household h; h < 0: none): a14 + a15 income, permanent income, consumption budget, then
// a16 the z_c candidates, sorted by price, packed.  pa: wealth in, wealth less the budget out.
template <int ZM, class C>
__device__ __forceinline__ void goods_prepare(C &c, const Mkt &mkt, int z, int h, int oc, double dinc, double &pa, double pm,
                                              double &b, GoodsList<C> &cl)
{
    const auto &L = c.layout();
    const int W = L.W, F = L.F;
    volatile double *gk = c.gk();
    b = 0.0;
    const double rpavg = gk[G_rpavg];
    if (h >= 0) {
        const double price = gk[G_price];
        double ir;
        if (h < W) { const bool emp = oc >= 0; pa = pa + (emp ? gk[G_net] : gk[G_sub]); ir = emp ? gk[G_ir_emp] : gk[G_ir_un]; }
        else ir = C::kStrict ? dinc / price : dinc * rpavg;
        pm = pm * gk[G_xi] + gk[G_omxi] * ir;
        const double target = pm + (C::kStrict ? (gk[G_chi] * pa) / price : (gk[G_chi] * pa) * rpavg);
        b = target * price;
        if (b > pa) b = pa;
        if (!(b > 0.0)) b = 0.0;
        pa = pa - b;
        c.perm(h) = pm;
    }
    if (b > 0.0) {
        int cand[ZM]; unsigned key[ZM];
        c.d.template candidates<ZM>(PH_GOODS, h, F, z, cand);
        // key = price class (14 bits) | draw index (3 bits) | seller (14 bits): F <= 16383
#pragma unroll
        for (int k = 0; k < ZM; ++k)
            key[k] = k < z ? (mkt.cls[cand[k]] << 17) | ((unsigned)k << 14) | (unsigned)cand[k] : 0xffffffffu;
        sort_keys<ZM>(key);
#pragma unroll
        for (int k = 0; k < ZM; ++k)
            if (k < z) cl.set(k, key[k] & 0x3fffu);
    }
}

template <int ZM, class C>
__device__ __forceinline__ void goods_commit_chunk(C &c, Mkt &mkt, unsigned long long *acc, int h, double pa, double b, GoodsList<C> cl);

// The peer table of rule R2 (one word per seller, in the firing / capital-market scratch, idle during the goods
// market): in a round with uncertain lanes, the mask of the lanes whose terminal the seller is; zero between rounds.
template <class C>
__device__ __forceinline__ int *goods_peer_table(const C &c) { return c.iat(c.layout().s_ic); }
template <class C>
__device__ __forceinline__ void goods_peer_table_init(const C &c, int first, int stride)
{
    int *tab = goods_peer_table(c);
    for (int f = first; f < c.layout().F; f += stride) tab[f] = 0;
}

template <int ZM, class C>
__device__ void phase_goods_market(C &c)
{
    const auto &L = c.layout();
    const int W = L.W, F = L.F, H = L.H, lane = c.lane;
    const int z = c.z_c() < F ? c.z_c() : F;
    Mkt mkt(c.sm + L.s_U, F);
    unsigned long long *acc = reinterpret_cast<unsigned long long *>(c.cf(C_Yd));   // GLOBAL, zeroed by produce
    OrderT<C::kStatic ? DefaultLayout::kW + DefaultLayout::kF + DefaultLayout::kN : 0, !C::kProd> ord; ord.init(c.d, MKT_GOODS, c.a->od[MKT_GOODS], c.okeys(), lane);
    volatile double *gk = c.gk();
    if (lane == 0) goods_constants(c);
    goods_peer_table_init(c, lane, 32);
    __syncwarp();
    const int nch = (H + 31) / 32;

    // one chunk of households ahead: id, wealth, permanent income, employment link / dividend
    int h_n = -1, oc_n = -1; double pa_n = 0.0, pm_n = 0.0, dinc_n = 0.0;
    auto fetch = [&](int ch) {
        const int pos = ch * 32 + lane;
        h_n = -1; oc_n = -1; pa_n = 0.0; pm_n = 0.0; dinc_n = 0.0;
        if (pos < H) {
            const int h = ord.at(pos);
            const double2 rec = c.hh()[h];
            h_n = h; pa_n = rec.x; pm_n = rec.y;
            if (h < W) oc_n = ((c.emp()[h >> 5] >> (h & 31)) & 1u) ? 0 : -1;   // employed? (bits set by the job search)
            else if (h < W + F) dinc_n = c.cf(C_div_income)[h - W];
            else dinc_n = c.kf(K_div_income)[h - W - F];
        }
    };
    fetch(0);
    for (int ch = 0; ch < nch; ++ch) {
        const int h = h_n, oc = oc_n;
        double pa = pa_n, pm = pm_n;
        const double dinc = dinc_n;
        if (ch + 1 < nch) fetch(ch + 1);
        double b;
        GoodsList<C> cl;
        goods_prepare<ZM>(c, mkt, z, h, oc, dinc, pa, pm, b, cl);
        goods_commit_chunk<ZM>(c, mkt, acc, h, pa, b, cl);
    }
    // close: unpack the seller records.  The reductions above were issued by other lanes: fence, then
    // read the accumulators where they live (L2)
    __threadfence();
    __syncwarp();
    {
        double *Yd = c.cf(C_Yd), *Q = c.cf(C_Q);
        for (int f = lane; f < F; f += 32) {
            const double2 r = mkt.yr[f];
            const unsigned long long a = __ldcg(&acc[f]);
            Yd[f] = C::kStrict ? r.y : (fx_to(a) * gk[G_price]) * r.y; Q[f] = c.cf(C_Y_prev)[f] - r.x;
        }
    }
    __syncwarp();
}

// commit (rounds) of one chunk of 32 prepared households; ends with what is left of b going back to PA
template <int ZM, class C>
__device__ __forceinline__ void goods_commit_chunk(C &c, Mkt &mkt, unsigned long long *acc, int h, double pa, double b, GoodsList<C> cl)
{
    const int lane = c.lane;
    const unsigned lt = c.lt();
    const double rpavg = c.gk()[G_rpavg];
    {
        unsigned pend = __ballot_sync(FULL, b > 0.0);
        // term: the seller this lane will buy from (>= 0), TERM_NONE (nothing left to visit / done)
        // or TERM_WALK (has budget and must walk down its list to the next seller with stock)
        enum { TERM_NONE = -1, TERM_WALK = -2 };
        int term = b > 0.0 ? TERM_WALK : TERM_NONE;
        double d_t = 0.0;
        unsigned long long xb = C::kStrict ? 0ull : fx_from(b * rpavg);   // this lane's demand in real units, fixed point
        // strict mode: the demand b / P of every visit, in the order of the (sorted) candidate list; added to
        // the sellers' accumulators in visiting order when the chunk is done
        const auto cl0 = cl;
        int nv = 0; double vv[C::kStrict ? ZM : 1];
        while (pend) {
            // ---- walk: dirty lanes register demand seller by seller until one has stock
            // (a round nearly always has a walker -- the chunk's first round, or the lane that just
            // exhausted a seller -- so the test comes after the step)
            auto walk_step = [&]() {
                int cnd;
                if (!cl.pop(cnd)) term = TERM_NONE;   // every candidate is sold out
                else {
                    const double2 yr = mkt.yr[cnd];   // stock, 1 / price
                    if constexpr (C::kStrict) {
                        const double dem = b / mkt.p[cnd];
#pragma unroll
                        for (int k = 0; k < ZM; ++k) if (k == nv) vv[k] = dem;
                        ++nv;
                        if (yr.x > 0.0) { term = cnd; d_t = dem; }
                    } else {
                        fx_red(&acc[cnd], xb);
                        if (yr.x > 0.0) { term = cnd; d_t = b * yr.y; }
                    }
                }
            };
            do {   // up to three steps per vote (measured: 1 -> 2 -> 3 steps +0.9 %, +2.0 %; all five -1 %):
                   // late in the market most walks are longer than one step
                if (term == TERM_WALK) {
                    walk_step();
                    if (term == TERM_WALK) walk_step();
                    if (term == TERM_WALK) walk_step();
                }
            } while (__any_sync(FULL, term == TERM_WALK));
            // ---- lanes at the same seller replay the lower lanes' purchases, in lane order
            const bool mine = (pend >> lane) & 1u;
            const bool has = mine && term >= 0;
            // lanes without a terminal share ONE dummy value (they never read their mask): MATCH.ANY takes as long as
            // there are distinct values in the warp, and from a chunk's second round on most lanes have no terminal
            const unsigned peers = __match_any_sync(FULL, has ? term : -1);
            unsigned rem = has ? (peers & lt) : 0u;
            double ys = 0.0;
            if (has) ys = mkt.yr[term].x;
            int esrc = -1;   // the lane that empties my seller before my turn (>= 0: sold out when I get there)
            while (__any_sync(FULL, rem != 0u)) {
                // branch-free body: a lane without peers left reads lane 31 and changes nothing
                const int src = __ffs(rem) - 1;
                const double di = __shfl_sync(FULL, d_t, src);
                const bool act = rem != 0u, more = act && ys > di;
                if (act && !more) esrc = src;            // sold out before my turn: stop replaying
                rem = more ? (rem & (rem - 1u)) : 0u;
                ys = more ? ys - di : ys;
            }
            const bool sold = esrc >= 0;
            const bool full = has && !sold && ys > d_t;
            // ---- who may commit (rule R2, cats_period_mw.cuh): a lane that exhausts its seller or finds it sold out
            // walks on -- to one of the sellers left on its list; lanes above it whose terminal is on that list
            // cannot be sure of the stock they replayed (and if they buy there in full they are uncertain in turn:
            // their lists count too).  Everybody else commits, also past a seller that sells out.
            const unsigned movers = __ballot_sync(FULL, has && (sold || !full));
            bool endangered = false;
            if (movers) {
                // The lanes at a seller post their peer mask under the seller's row (one plain store each: the peers
                // write the same word); an uncertain lane ORs the rows of the sellers left on its list -- the lanes
                // it endangers, all at once -- and one OR-reduction over the warp gives every endangered lane.  One
                // pass per level of the cascade instead of one per uncertain lane.  The rows are zero between rounds.
                int *tab = goods_peer_table(c);
                if (has) tab[term] = (int)peers;
                __syncwarp();
                const unsigned fullm = __ballot_sync(FULL, full);
                const unsigned openm = __ballot_sync(FULL, has && !sold);
                unsigned unc = movers, fresh = movers, endm = 0u;
                do {
                    unsigned v = 0u;
                    if ((fresh >> lane) & 1u) {
                        auto li = cl;
#pragma unroll
                        for (int k = 0; k < ZM - 1; ++k) { int sl; if (li.pop(sl)) v |= (unsigned)tab[sl]; }
                        v &= ~(lt | (1u << lane));   // the lanes above me
                    }
                    const unsigned hits = __reduce_or_sync(FULL, v) & openm & ~endm;
                    endm |= hits;
                    fresh = hits & fullm & ~unc;
                    unc |= fresh;
                } while (fresh);
                endangered = (endm >> lane) & 1u;
                if (has) tab[term] = 0;   // (every read of the table is behind the last reduction)
            }
            const unsigned commit = __ballot_sync(FULL, mine && (!has || (!sold && !endangered)));
            // branch-free: a FULL lane buys d_t, an EXHAUST lane takes what is left and walks on with the rest of its
            // budget, a lane queued behind a seller that is now empty for good walks on, everybody else keeps its state
            const bool cm = (commit >> lane) & 1u;
            const bool buy = cm && has, exh = buy && !full;
            const bool lastp = (peers & commit & ~lt & ~(1u << lane)) == 0u;   // last committing lane at this seller
            const double left = exh ? b - ys * mkt.p[term] : 0.0;              // EXHAUST: the budget that remains
            if (buy && lastp) mkt.yr[term].x = full ? ys - d_t : 0.0;
            if (buy) b = left;
            if constexpr (!C::kStrict) { if (exh) xb = fx_from(left * rpavg); }
            const bool walk_on = (exh && left > 0.0) || (!cm && sold && ((commit >> esrc) & 1u));
            const bool done = cm && !(exh && left > 0.0);
            term = walk_on ? TERM_WALK : (done ? TERM_NONE : term);
            pend &= ~__ballot_sync(FULL, done);
            __syncwarp();
        }
        if (h >= 0) c.PA(h) = pa + b;   // involuntary saving
        if constexpr (C::kStrict) {
            // Yd += b / P at every seller visited, household by household in visiting order (= lane order)
            unsigned m = __ballot_sync(FULL, nv > 0);
            while (m) {
                const int l = __ffs(m) - 1; m &= m - 1;
                if (lane == l) {
                    auto c2 = cl0;
#pragma unroll
                    for (int k = 0; k < ZM; ++k)
                        if (k < nv) { int sl = 0; c2.pop(sl); mkt.yr[sl].y = mkt.yr[sl].y + vv[k]; }
                }
                __syncwarp();
            }
        }
    }
}

// a17 firms_accounting! (cats.py:397-398)
// The per-firm loops of the accounting phase over C-firms [f_lo, f_hi) and K-firms [k_lo, k_hi).  acc[0..5]: the
// det_sum partials (installments, interests, dividend tax; C then K) of a warp that walks a whole population.
// STAGE (multi-warp kernel): the summands are also left in transient rows that are dead by now (C: instalment in
// the Kdem row, dividend tax in the valinv row; K: instalment in the Ftot row, dividend tax in the Ystock row; the
// interests are a persisted row), for ONE warp to add up in the oracle's order afterwards.
template <bool STAGE, class C>
__device__ __forceinline__ void accounting_loops(C &c, int f_lo, int f_hi, int k_lo, int k_hi, double (&acc)[6])
{
    const auto &L = c.layout();
    const int W = L.W, F = L.F, last = c.last;
    const double delta = c.par()[P_delta], eta = c.par()[P_eta], k = c.par()[P_k], theta = c.par()[P_theta];
    const double divs = c.par()[P_div], taxd = c.par()[P_tax_rate_d];
    const Mkt mkt(c.sm + L.s_U, F);
    double s_in = 0.0, s_int = 0.0, s_dt = 0.0, k_in = 0.0, k_int = 0.0, k_dt = 0.0;   // det_sum partials
    for (int f = f_lo + c.lane; f < f_hi; f += 32) {
        // loads first (see phase_firm_decisions)
        const double Yp = c.cf(C_Y_prev)[f], barYK = c.cf(C_barYK)[f], Kold = c.cf(C_K)[f];
        const double inv_f = c.cf(C_investment)[f], cv = c.cf(C_capital_value)[f], Q_f = c.cf(C_Q)[f];
        const double wages = c.cf(C_wages)[f], interests = c.cf(C_interests)[f], deb = c.cf(C_deb)[f];
        const double liq0 = c.cf(C_liquidity)[f], A0 = c.cf(C_A)[f], pa_owner = c.PA(W + f);
        c.cf(C_barYK)[f] = delta * barYK + (1.0 - delta) * (Yp / k);
        double dep = (eta * Yp) / k;
        c.cf(C_K)[f] = (Kold - dep) + inv_f;
        double dv = 0.0;
        if (Kold != 0.0) dv = (dep * cv) / Kold;
        c.cf(C_capital_value)[f] = (cv - dv) + c.at(L.s_cvalinv)[f];
        double RIC = mkt.p[f] * Q_f;
        double in = theta * deb;
        double liq = ((((liq0 + RIC) + c.at(L.s_cFtot)[f]) - wages) - interests) - in;
        if (STAGE) c.at(L.s_cKdem)[f] = in;
        c.cf(C_deb)[f] = deb - in;
        double pi = ((RIC - wages) - interests) - dv;
        double A = A0 + pi;
        double net = 0.0, dt = 0.0;
        if (pi > 0.0) {
            double dvd = divs * pi;
            liq = liq - dvd; A = A - dvd;
            net = dvd * (1.0 - taxd);
            dt = dvd - net;
            c.PA(W + f) = pa_owner + net;
        }
        c.cf(C_div_income)[f] = net;
        c.cf(C_liquidity)[f] = liq; c.cf(C_A)[f] = A;
        if (STAGE) c.at(L.s_cvalinv)[f] = dt;
        s_in = s_in + in; s_int = s_int + interests; s_dt = s_dt + dt;
    }
    if (last >= 22) {
        for (int i = k_lo + c.lane; i < k_hi; i += 32) {
            const double left = c.at(L.s_kYstock)[i];
            const double Qi = c.kf(K_Yavail)[i] - left;   // sold = available - left over
            c.kf(K_Q)[i] = Qi;
            double RIC = c.kf(K_P)[i] * Qi;
            c.kf(K_inv)[i] = (1.0 - c.par()[P_inventory_depreciation]) * left;
            double wages = c.kf(K_wages)[i], interests = c.kf(K_interests)[i];
            double deb = c.kf(K_deb)[i];
            double in = theta * deb;
            double liq = ((((c.kf(K_liquidity)[i] + RIC) + c.at(L.s_kFtot)[i]) - wages) - interests) - in;
            c.kf(K_deb)[i] = deb - in;
            double pi = (RIC - wages) - interests;
            double A = c.kf(K_A)[i] + pi;
            double net = 0.0, dt = 0.0;
            if (pi > 0.0) {
                double dvd = divs * pi;
                liq = liq - dvd; A = A - dvd;
                net = dvd * (1.0 - taxd);
                dt = dvd - net;
                c.PA(W + F + i) = c.PA(W + F + i) + net;
            }
            c.kf(K_div_income)[i] = net;
            c.kf(K_liquidity)[i] = liq; c.kf(K_A)[i] = A;
            if (STAGE) { c.at(L.s_kFtot)[i] = in; c.at(L.s_kYstock)[i] = dt; }
            k_in = k_in + in; k_int = k_int + interests; k_dt = k_dt + dt;
        }
    }
    acc[0] = s_in; acc[1] = s_int; acc[2] = s_dt; acc[3] = k_in; acc[4] = k_int; acc[5] = k_dt;
}

// the bank's and the government's books take the six sums
template <class C>
__device__ __forceinline__ void accounting_book(C &c, double s_in, double s_int, double s_dt, double k_in, double k_int, double k_dt)
{
    const int last = c.last;
    {
        double loans = c.scal()[S_bank_loans] - s_in;
        double ii = c.scal()[S_bank_interest_income] + s_int;
        double ta = c.scal()[S_gov_TA] + s_dt;
        if (last >= 22) { loans = loans - k_in; ii = ii + k_int; ta = ta + k_dt; }
        c.scal()[S_bank_loans] = loans; c.scal()[S_bank_interest_income] = ii; c.scal()[S_gov_TA] = ta;
    }
}

template <class C>
__device__ void phase_accounting(C &c)
{
    const auto &L = c.layout();
    double acc[6];
    accounting_loops<false>(c, 0, L.F, 0, L.N, acc);
    double s_in = acc[0], s_int = acc[1], s_dt = acc[2], k_in = acc[3], k_int = acc[4], k_dt = acc[5];
    const int last = c.last;
    // six det_sums: installments, interests, dividend tax for C then K
    s_in = bfly(s_in); s_int = bfly(s_int); s_dt = bfly(s_dt);
    k_in = bfly(k_in); k_int = bfly(k_int); k_dt = bfly(k_dt);
    if (c.lane == 0) {
        double loans = c.scal()[S_bank_loans] - s_in;
        double ii = c.scal()[S_bank_interest_income] + s_int;
        double ta = c.scal()[S_gov_TA] + s_dt;
        if (last >= 22) { loans = loans - k_in; ii = ii + k_int; ta = ta + k_dt; }
        c.scal()[S_bank_loans] = loans; c.scal()[S_bank_interest_income] = ii; c.scal()[S_gov_TA] = ta;
    }
    __syncwarp();
}

// a18 Cats._get_reward (cats.py:493-556), evaluated between accounting and tracking
template <class C>
__device__ void phase_reward(C &c)
{
    const KArgs &a = *c.a;
    if (!a.reward || a.n_agents <= 0) return;
    const int F = c.layout().F, nA = a.n_agents;
    const double *P = c.cf(C_P), *Q = c.cf(C_Q);
    if (a.flags & CATS_FLAG_REWARD_RMS) {
        // cats.py:548-549: Python's built-in sum() over the RL firms' revenues, then over the others'.
        // CPython >= 3.12 adds floats with Neumaier's compensated summation (Objects/bltinmodule.c);
        // the two sums are then added plainly.  MODEL.md 5, phase 23.
        double total = 0.0;
        if (c.lane == 0) {
            auto py_sum = [&](int lo, int hi) {
                double acc = 0.0, comp = 0.0;
                for (int f = lo; f < hi; ++f) {
                    const double x = P[f] * Q[f], t = acc + x;
                    if (fabs(acc) >= fabs(x)) comp = comp + ((acc - t) + x); else comp = comp + ((x - t) + acc);
                    acc = t;
                }
                if (comp != 0.0 && isfinite(comp)) acc = acc + comp;
                return acc;
            };
            total = py_sum(0, nA) + py_sum(nA, F);
        }
        total = __shfl_sync(FULL, total, 0);
        for (int i = c.lane; i < nA; i += 32) a.reward[c.b * nA + i] = (P[i] * Q[i]) / total;
        return;
    }
    const double eta = c.par()[P_eta], k = c.par()[P_k];
    for (int i = c.lane; i < nA; i += 32) {
        double dep = (eta * c.cf(C_Y_prev)[i]) / k;
        double den = (c.cf(C_K)[i] + dep) - c.cf(C_investment)[i];
        double dv;
        if (den == 0.0) { dv = 0.0; atomicOr(&c.misc()[M_STATUS], (int)CATS_STATUS_DEPVAL_DIV0); }
        else dv = (dep * c.cf(C_capital_value)[i]) / den;
        double RIC = P[i] * Q[i];
        double profits = ((RIC - c.cf(C_wages)[i]) - c.cf(C_interests)[i]) - dv;
        profits = (profits * c.scal()[S_init_price]) / c.scal()[S_price];
        if (profits != profits) { profits = 0.0; atomicOr(&c.misc()[M_STATUS], (int)CATS_STATUS_PROFIT_NAN); }
        if (c.cf(C_A)[i] <= 0.0) profits = a.bankruptcy_reward;
        a.reward[c.b * nA + i] = profits;
    }
    __syncwarp();
}

// a19 update_tracking_variables! (cats.py:401)
template <class C>
__device__ void phase_tracking(C &c)
{
    const auto &L = c.layout();
    const int F = L.F, N = L.N;
    const double *cP = c.cf(C_P), *cQ = c.cf(C_Q), *cY = c.cf(C_Y_prev), *cK = c.cf(C_K);
    double sQ = 0.0, sPQ = 0.0, sP = 0.0, sY = 0.0, sK = 0.0, sQk = 0.0, sPQk = 0.0, sYk = 0.0;
    int nL = 0;
    for (int f = c.lane; f < F; f += 32) {
        const double p = cP[f], q = cQ[f];
        sQ = sQ + q; sPQ = sPQ + p * q; sP = sP + p; sY = sY + cY[f]; sK = sK + cK[f];
        nL += c.Leff()[f];
    }
    for (int i = c.lane; i < N; i += 32) {
        const double p = c.kf(K_P)[i], q = c.kf(K_Q)[i];
        sQk = sQk + q; sPQk = sPQk + p * q; sYk = sYk + c.kf(K_Y_prev)[i];
        nL += c.Leff()[F + i];
    }
    sQ = bfly(sQ); sPQ = bfly(sPQ); sP = bfly(sP); sY = bfly(sY); sK = bfly(sK);
    sQk = bfly(sQk); sPQk = bfly(sPQk); sYk = bfly(sYk);
    nL = bfly_i(nL);
    if (c.lane == 0) {
        double *s = c.scal();
        double price_prev = s[S_price];
        double price = sQ > 0.0 ? sPQ / sQ : sP / (double)F;
        double price_k = sQk > 0.0 ? sPQk / sQk : s[S_price_k];
        double Yreal = sY * s[S_init_price] + sYk * s[S_init_price_k];
        double Ynom = sY * price + sYk * price_k;
        s[S_price] = price; s[S_price_k] = price_k;
        s[S_Y_real] = Yreal; s[S_Y_nominal_tot] = Ynom;
        s[S_gdp_deflator] = Yreal > 0.0 ? Ynom / Yreal : 0.0;
        s[S_inflationRate] = price / price_prev - 1.0;
        s[S_Un] = 1.0 - (double)nL / (double)L.W;
        s[S_consumption] = sPQ;
        s[S_totK] = sK;
    }
    __syncwarp();
}

// a20 firms_go_bankrupt! (cats.py:404) + a21 _get_terminated (cats.py:483-486)
// flags, the masked sums over survivors / defaulters, the bank's books and the entrants.  One warp.  Returns the
// number of bankrupt firms; their workers are released by bankrupt_release.
template <class C>
__device__ __forceinline__ int bankrupt_replace(C &c)
{
    const auto &L = c.layout();
    const int W = L.W, F = L.F, N = L.N;
    const double thr = c.par()[P_E_threshold_scale] * c.scal()[S_price];
    unsigned char *bad = c.sm + L.s_bad;
    // masked det_sums over the survivors / the defaulters, and survivor counts
    double cK = 0.0, cL = 0.0, cP = 0.0, cY = 0.0, cLoss = 0.0, cDeb = 0.0;
    double kL = 0.0, kP = 0.0, kY = 0.0, kLoss = 0.0, kDeb = 0.0;
    int nsC = 0, nsK = 0;
    for (int f = c.lane; f < F; f += 32) {
        const double A = c.cf(C_A)[f], liq = c.cf(C_liquidity)[f], deb = c.cf(C_deb)[f];
        const bool b = (A <= thr) || (liq < 0.0);
        bad[f] = b ? 1 : 0;
        nsC += b ? 0 : 1;
        double expo = deb + (liq < 0.0 ? -liq : 0.0);
        double loss = A < 0.0 ? -A : 0.0;
        if (loss > expo) loss = expo;
        cK = cK + (b ? 0.0 : c.cf(C_K)[f]);
        cL = cL + (b ? 0.0 : liq);
        cP = cP + (b ? 0.0 : c.cf(C_P)[f]);
        cY = cY + (b ? 0.0 : c.cf(C_Y_prev)[f]);
        cLoss = cLoss + (b ? loss : 0.0);
        cDeb = cDeb + (b ? deb : 0.0);
    }
    for (int i = c.lane; i < N; i += 32) {
        const double A = c.kf(K_A)[i], liq = c.kf(K_liquidity)[i], deb = c.kf(K_deb)[i];
        const bool b = (A <= thr) || (liq < 0.0);
        bad[F + i] = b ? 1 : 0;
        nsK += b ? 0 : 1;
        double expo = deb + (liq < 0.0 ? -liq : 0.0);
        double loss = A < 0.0 ? -A : 0.0;
        if (loss > expo) loss = expo;
        kL = kL + (b ? 0.0 : liq);
        kP = kP + (b ? 0.0 : c.kf(K_P)[i]);
        kY = kY + (b ? 0.0 : c.kf(K_Y_prev)[i]);
        kLoss = kLoss + (b ? loss : 0.0);
        kDeb = kDeb + (b ? deb : 0.0);
    }
    cK = bfly(cK); cL = bfly(cL); cP = bfly(cP); cY = bfly(cY); cLoss = bfly(cLoss); cDeb = bfly(cDeb);
    kL = bfly(kL); kP = bfly(kP); kY = bfly(kY); kLoss = bfly(kLoss); kDeb = bfly(kDeb);
    nsC = bfly_i(nsC); nsK = bfly_i(nsK);
    double mK = 10.0, mL = 0.0, mP = c.scal()[S_price], mY = c.par()[P_alpha];
    if (nsC > 0) { mK = cK / (double)nsC; mL = cL / (double)nsC; mP = cP / (double)nsC; mY = cY / (double)nsC; }
    double mLk = 0.0, mPk = c.scal()[S_price_k], mYk = c.par()[P_alpha];
    if (nsK > 0) { mLk = kL / (double)nsK; mPk = kP / (double)nsK; mYk = kY / (double)nsK; }
    const double price_k = c.scal()[S_price_k];
    __syncwarp();
    if (c.lane == 0) {
        double bd = c.scal()[S_bank_bad_debt] + cLoss;
        double loans = c.scal()[S_bank_loans] - cDeb;
        bd = bd + kLoss;
        loans = loans - kDeb;
        c.scal()[S_bank_bad_debt] = bd; c.scal()[S_bank_loans] = loans;
        c.scal()[S_bankruptcy_rate] = (double)((F - nsC) + (N - nsK)) / (double)(F + N);
    }
    const int nbad = (F - nsC) + (N - nsK);
    if (nbad > 0) {
        for (int f = c.lane; f < F; f += 32) {
            if (!bad[f]) continue;
            const int ow = W + f;
            double inj = mL > 0.0 ? mL : 0.0;
            double pa = c.PA(ow);
            if (inj > pa) inj = pa;
            if (!(inj > 0.0)) inj = 0.0;
            c.PA(ow) = pa - inj;
            double cv = mK * price_k;
            c.cf(C_liquidity)[f] = inj; c.cf(C_K)[f] = mK; c.cf(C_capital_value)[f] = cv;
            c.cf(C_A)[f] = inj + cv; c.cf(C_deb)[f] = 0.0; c.cf(C_P)[f] = mP;
            c.cf(C_Y_prev)[f] = mY; c.cf(C_Yd)[f] = mY; c.cf(C_De)[f] = mY;
            c.cf(C_barYK)[f] = mY / c.par()[P_k]; c.cf(C_interest_r)[f] = c.par()[P_r_f];
            c.cf(C_Q)[f] = 0.0; c.cf(C_investment)[f] = 0.0; c.cf(C_wages)[f] = 0.0;
            c.cf(C_interests)[f] = 0.0; c.cf(C_div_income)[f] = 0.0;
            c.Leff()[f] = 0;
        }
        for (int i = c.lane; i < N; i += 32) {
            if (!bad[F + i]) continue;
            const int ow = W + F + i;
            double inj = mLk > 0.0 ? mLk : 0.0;
            double pa = c.PA(ow);
            if (inj > pa) inj = pa;
            if (!(inj > 0.0)) inj = 0.0;
            c.PA(ow) = pa - inj;
            c.kf(K_liquidity)[i] = inj; c.kf(K_A)[i] = inj; c.kf(K_deb)[i] = 0.0;
            c.kf(K_P)[i] = mPk; c.kf(K_Y_prev)[i] = mYk; c.kf(K_Yd)[i] = mYk; c.kf(K_Yavail)[i] = mYk;
            c.kf(K_De)[i] = mYk; c.kf(K_inv)[i] = 0.0; c.kf(K_interest_r)[i] = c.par()[P_r_f];
            c.kf(K_Q)[i] = 0.0; c.kf(K_wages)[i] = 0.0; c.kf(K_interests)[i] = 0.0;
            c.kf(K_div_income)[i] = 0.0;
            c.Leff()[F + i] = 0;
        }
    }
    return nbad;
}

// the employees of the bankrupt firms lose their jobs: workers start, start + stride, ...
template <class C>
__device__ __forceinline__ void bankrupt_release(C &c, int start, int stride)
{
    const auto &L = c.layout();
    const int W = L.W;
    const unsigned char *bad = c.sm + L.s_bad;
    for (int w0 = start; w0 < W; w0 += stride * 4) {
        int oo[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { const int w = w0 + stride * u; oo[u] = w < W ? c.Oc()[w] : -1; }   // 4 loads in flight
#pragma unroll
        for (int u = 0; u < 4; ++u) if (oo[u] >= 0 && bad[oo[u]]) c.Oc()[w0 + stride * u] = -1;
    }
}

// a21 _get_terminated (cats.py:483-486).  One warp.
template <class C>
__device__ __forceinline__ void bankrupt_terminated(C &c)
{
    const int F = c.layout().F;
    {
        const double *A = c.cf(C_A);
        double s = 0.0;
        for (int f = c.lane; f < F; f += 32) s = s + A[f];
        s = bfly(s);
        if (c.lane == 0) c.scal()[S_terminated] = s == 0.0 ? 1.0 : 0.0;
    }
}

template <class C>
__device__ void phase_bankrupt(C &c)
{
    const int nbad = bankrupt_replace(c);
    if (nbad > 0) bankrupt_release(c, c.lane, 32);
    __syncwarp();
    bankrupt_terminated(c);
    __syncwarp();
}

// a22 bank_accounting!, a23 gov_accounting!, a24 adjust_wages!, a25 timestep (cats.py:407-413)
// the scalar part, by ONE thread; returns each owner household's share of the bank's dividend
template <class C>
__device__ __forceinline__ double closing_scalars(C &c)
{
    const auto &L = c.layout();
    const int last = c.last;
    double share = 0.0;
    {
        double *s = c.scal();
        double gross = (s[S_bank_interest_income] + s[S_gov_r_b] * s[S_gov_bonds]) - s[S_bank_bad_debt];
        double dB = 0.0;
        if (gross > 0.0 && s[S_bank_E] > 0.0) dB = c.par()[P_div] * gross;
        s[S_bank_E] = s[S_bank_E] + (gross - dB);
        s[S_dividendsB] = dB; s[S_profitsB] = gross;
        share = dB > 0.0 ? dB / (double)(L.F + L.N) : 0.0;
        if (last >= 27) {
            s[S_gov_G] = s[S_gov_G] + s[S_gov_r_b] * s[S_gov_bonds];
            double GB = s[S_gov_TA] - s[S_gov_G];
            s[S_GB] = GB;
            s[S_gov_bonds] = s[S_gov_bonds] - GB;
            s[S_deficitGDP] = s[S_Y_nominal_tot] > 0.0 ? (-GB) / s[S_Y_nominal_tot] : 0.0;
        }
        if (last >= 28) {
            double Un = s[S_Un], ut = c.par()[P_u_target];
            if (Un < ut) s[S_wb] = s[S_wb] * (1.0 + c.par()[P_wage_update_up] * (ut - Un));
            else s[S_wb] = s[S_wb] * (1.0 - c.par()[P_wage_update_down] * (Un - ut));
        }
        if (last >= 29) s[S_timestep] = s[S_timestep] + 1.0;
    }
    return share;
}

template <class C>
__device__ void phase_closing(C &c)
{
    const auto &L = c.layout();
    double share = 0.0;
    if (c.lane == 0) share = closing_scalars(c);
    share = __shfl_sync(FULL, share, 0);
    if (share > 0.0)
        for (int h = L.W + c.lane; h < L.H; h += 32) c.PA(h) = c.PA(h) + share;
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// one period; phase ids as docs/MODEL.md section 6 (stop_phase halts after the given id)
template <int ZM, int G, class C>
__device__ void run_period(C &c)
{
    const KArgs &a = *c.a;
    const auto &L = c.layout();
    const int F = L.F, N = L.N, last = c.last;
#ifdef CATS_PROFILE_PHASES
    long long t0 = 0;
    const bool prof = a.cycles && c.lane == 0 && c.b == 0;
    if (prof) t0 = clock64();
#define TICK(slot) do { if (prof) { long long t1 = clock64(); a.cycles[slot] += t1 - t0; t0 = t1; } } while (0)
#else
#define TICK(slot) do { } while (0)
#endif
    // phases 1-2: reset_gov_and_banking_variables!, set_bond_interest_rate! (cats.py:355-356)
    if (c.lane == 0) {
        c.scal()[S_gov_TA] = 0.0; c.scal()[S_gov_G] = 0.0;
        c.scal()[S_bank_interest_income] = 0.0; c.scal()[S_bank_bad_debt] = 0.0;
        if (last >= 2) c.scal()[S_gov_r_b] = c.par()[P_interest_rate];
        c.misc()[M_STATUS] = 0;
    }
    __syncwarp();
    if (last < 3) return;
    phase_firm_decisions(c);  // 3..10
    TICK(0);
    if (last < 11) return;
    phase_fire(c, last >= 12 ? F + N : F);  // 11, 12
    __syncwarp();
    TICK(1);
    if (last < 13) return;
    phase_job_search<ZM>(c);  // 13
    __syncwarp();
    TICK(2);
    if (last < 14) return;
    phase_produce(c);         // 14, 15
    TICK(3);
    if (last < 16) return;
    phase_capital_market<ZM>(c);  // 16
    // household wealth / permanent income are first touched by the goods market: into L2 just ahead
    // of it (any earlier and the lines age out of L2 again before they are used)
    if (c.lane == 0 && last >= 18) bulk_prefetch_l2(c.blob + L.o_PA, (uint32_t)(L.blob - L.o_PA));
    TICK(4);
    if (last < 17) return;
    phase_gov_income(c);      // 17 (+ government side of 18)
    TICK(5);
    if (last < 18) return;
    if (last < 20) { phase_households_only(c); return; }
    phase_goods_market<ZM>(c);  // 18, 19, 20
    TICK(6);
    if (last < 21) return;
    // the goods market is where economies drift apart (its round count differs); the closing phases
    // are a long stretch of once-per-period code again: back in step before them
    // (not in the default-configuration instantiation: its code is a quarter smaller and the wait costs
    // more than the instruction fetches it saves -- measured +1.2 % without, against -4 % for the generic one)
    if (G > 1 && !C::kStatic && (C::kProd || a.bars)) __syncthreads();
    phase_accounting(c);      // 21, 22
    TICK(7);
    if (last < 23) return;
    phase_reward(c);          // 23
    __syncwarp();
    TICK(8);
    if (last < 24) return;
    phase_tracking(c);        // 24
    TICK(9);
    if (last < 25) return;
    phase_bankrupt(c);        // 25
    TICK(10);
    if (last < 26) return;
    phase_closing(c);         // 26..29
    TICK(11);
#undef TICK
}

// G economies per block, one warp each, every warp on its own slice of the block's shared memory.
// The warps of a block never exchange data; they share a block so that they START TOGETHER: economies
// of one size take the same path through the ~170 KB of period code at nearly the same pace, and warps
// that run in step fetch each instruction line once for all of them.  One-warp blocks are rescheduled
// one by one and spread over all phases, which costs a quarter of the throughput in instruction-cache
// misses ("instruction supply").  Multi-period launches re-align at every period.
// MASKED: the launch honours KArgs::active (per-economy resets).  A separate instantiation, so that
// the test at the kernel entry cannot disturb the register allocation of the common, unmasked path
// (measured: the same source with the test compiled in is 3 % slower even when active == nullptr).
// MODE: see CtxT.  The host picks 2 when (W, F, N) == (1000, 100, 20) and the slice is not padded, 0 when
// the call carries a tape, a debug stop or a debug image, 1 otherwise.
template <int ZM, int G, int MINB, bool MASKED, int MODE, bool STRICT = false>
__global__ void __launch_bounds__(32 * G, MINB) cats_period_kernel(const __grid_constant__ KArgs a)
{
#ifdef CATS_EMU
    unsigned char *smem_block = simt::g_blk->smem;
#else
    extern __shared__ __align__(128) unsigned char smem_block[];
#endif
    constexpr bool ST = MODE == 2, PROD = MODE >= 1;
    CtxT<MODE, STRICT> c;
    c.a = &a;
    const auto &L = c.layout();
    const int slice = ST ? ((DefaultLayout::smem + 15) & ~15) : a.slice;
    unsigned char *smem = smem_block + (size_t)(threadIdx.x >> 5) * (size_t)slice;
    c.sm = smem;
    c.lane = threadIdx.x & 31;
    c.last = a.stop_phase > 0 ? a.stop_phase : 99;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + L.s_misc + 8);
    const long long b = (long long)blockIdx.x * G + (threadIdx.x >> 5);
    if (b >= a.B || (MASKED && !a.active[b])) {
        if (G > 1) for (int t = 0; t < a.n_periods; ++t) {   // keep the block's barrier count
            if (t > 0) __syncthreads();
            if (!ST && (PROD || a.bars)) __syncthreads();
        }
        return;
    }
    c.b = b;
    unsigned char *blob = a.blobs + (size_t)b * (size_t)L.blob;
    c.blob = blob;

    // HBM -> smem: one TMA bulk copy of the resident prefix (aggregates, K-firm tables, head-counts)
    if (c.lane == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
        mbar_expect_tx(bar, (uint32_t)L.resident);
        bulk_g2s(smem, blob, (uint32_t)L.resident, bar);
        // the streamed tail is first touched phases apart: pull what the firm phases and the labour
        // market need (employment links, firm rows) into L2 now, household wealth before the goods market
        bulk_prefetch_l2(blob + L.o_Oc, (uint32_t)(L.o_PA - L.o_Oc));
    }
    bool bad_z = false;
    // parameters (shared row or per-economy row) while the copy is in flight
    const double *pr = a.params + (a.params_stride ? (size_t)b * (size_t)a.params_stride : 0);
    for (int i = c.lane; i < N_PARAMS; i += 32) c.par()[i] = pr[i];
    if (c.lane == 0) c.par()[P_log1mtheta] = det_log(1.0 - pr[P_theta]);
    if (ST && c.lane == 0 && ((int)pr[P_z_c] != DefaultLayout::kZc || (int)pr[P_z_k] != DefaultLayout::kZk || (int)pr[P_z_e] != DefaultLayout::kZe))
        bad_z = true;      // the call declared the default z for every economy and this one disagrees
    // every transient is written before it is read within a period; only a debug image of a period
    // stopped half-way shows the ones not reached yet, and those must read as zero
    if (!PROD && a.debug)
        for (int i = L.resident / 4 + c.lane; i < L.s_par / 4; i += 32) reinterpret_cast<int *>(smem)[i] = 0;
    __syncwarp();
    mbar_wait(bar, 0);
    if (c.lane == 0) c.misc()[M_EPISODE] = (int)((uint32_t)c.scal()[S_episode] << 12);
    __syncwarp();

    c.d.rk = a.rk;
    c.d.economy = (uint32_t)(a.economy_base + (unsigned long long)b);
    c.d.tp = &a.T;
    c.d.tape_z = a.tape_z;
    c.d.bad = &c.misc()[M_STATUS];

    unsigned launch_status = 0u;   // lane 0: the status bits raised by THIS launch (what d_status reports)
    for (int t = 0; t < a.n_periods; ++t) {
        if (G > 1 && t > 0) __syncthreads();   // back in step (see above)
        // counter word 2: the period (32 bits); the episode (bumped by a masked reset) has its own 20 bits
        // in word 1, so neither a long episode nor many resets can alias another stream (MODEL.md 3.1)
        c.d.period = (uint32_t)c.scal()[S_timestep];
        c.d.epi = (uint32_t)c.misc()[M_EPISODE];
        if constexpr (!PROD)
            c.d.tape = a.tape ? a.tape + ((size_t)b * (size_t)a.n_periods + (size_t)t) * (size_t)a.T.bytes : nullptr;
        run_period<ZM, G>(c);
        __syncwarp();
        if (a.stats && c.lane < N_STATS) {
            const double *s = c.scal();
            double v;
            switch (c.lane) {
            case 0: v = s[S_Y_real]; break;       case 1: v = s[S_wb]; break;
            case 2: v = s[S_Un]; break;           case 3: v = s[S_bankruptcy_rate]; break;
            case 4: v = s[S_price]; break;        case 5: v = s[S_Y_nominal_tot]; break;
            case 6: v = s[S_gdp_deflator]; break; case 7: v = s[S_inflationRate]; break;
            case 8: v = s[S_consumption]; break;  case 9: v = s[S_totK]; break;
            case 10: v = s[S_dividendsB]; break;  case 11: v = s[S_profitsB]; break;
            case 12: v = s[S_GB]; break;          case 13: v = s[S_deficitGDP]; break;
            case 14: v = s[S_gov_bonds]; break;   default: v = s[S_price_k]; break;
            }
            a.stats[((size_t)b * (size_t)a.n_periods + (size_t)t) * N_STATS + c.lane] = v;
        }
        if (c.lane == 0 && c.misc()[M_STATUS]) {
            launch_status |= (unsigned)c.misc()[M_STATUS];
            c.scal()[S_status] = (double)((unsigned)c.scal()[S_status] | (unsigned)c.misc()[M_STATUS]);   // sticky copy
        }
        __syncwarp();
    }
    // a26 _get_obs (cats.py:466-481 / 655-671), a21 terminated, status
    if (a.obs) {
        for (int i = c.lane; i < a.n_agents; i += 32) {
            double yp = c.cf(C_Y_prev)[i], yd = c.cf(C_Yd)[i], p = c.cf(C_P)[i], ap = c.scal()[S_price];
            double o0, o1;
            if (a.flags & CATS_FLAG_LOG) {
                o0 = det_log(yp + 1e-8) - det_log(yd + 1e-8);
                o1 = det_log(p + 1e-8) - det_log(ap + 1e-8);
            } else { o0 = yp - yd; o1 = p - ap; }
            a.obs[2 * (b * a.n_agents + i)] = o0;
            a.obs[2 * (b * a.n_agents + i) + 1] = o1;
        }
    }
    if (c.lane == 0) {
        if (a.terminated) a.terminated[b] = c.scal()[S_terminated] != 0.0 ? 1 : 0;
        if (a.status) a.status[b] = launch_status | (bad_z ? CATS_STATUS_Z_MISMATCH : 0u);
    }
    if (!PROD && a.debug) {
        // transients image (cats_workset_field_info): Ld vac cFtot cKdem cYstock cvalinv kFtot kYstock
        unsigned char *img = a.debug + (size_t)b * (size_t)L.workset;
        const int nf = L.F + L.N;
        const Mkt mkt(smem + L.s_U, L.F);
        const bool have_mkt = c.last >= 14;
        for (int i = c.lane; i < L.workset / 4; i += 32) reinterpret_cast<int *>(img)[i] = 0;
        __syncwarp();
        for (int i = c.lane; i < nf; i += 32) {
            // Ld == Leff + vac; before the credit phase vac only carries Ld and reads as 0
            const bool isk = i >= L.F;
            reinterpret_cast<int *>(img + L.d_Ld)[i] = c.last >= (isk ? 8 : 7) ? c.Leff()[i] + c.vac()[i] : 0;
            reinterpret_cast<int *>(img + L.d_vac)[i] = c.last >= (isk ? 10 : 9) ? c.vac()[i] : 0;
        }
        for (int f = c.lane; f < L.F; f += 32) {
            reinterpret_cast<double *>(img + L.d_cFtot)[f] = c.at(L.s_cFtot)[f];
            reinterpret_cast<double *>(img + L.d_cKdem)[f] = c.at(L.s_cKdem)[f];
            reinterpret_cast<double *>(img + L.d_cYstock)[f] = have_mkt ? mkt.yr[f].x : 0.0;
            reinterpret_cast<double *>(img + L.d_cvalinv)[f] = c.at(L.s_cvalinv)[f];
        }
        for (int i = c.lane; i < L.N; i += 32) {
            reinterpret_cast<double *>(img + L.d_kFtot)[i] = c.at(L.s_kFtot)[i];
            reinterpret_cast<double *>(img + L.d_kYstock)[i] = c.at(L.s_kYstock)[i];
        }
    }
    if (c.last >= 15 && c.last < 22) {   // debug stop while the K-firm Q slot still carries 1 / P
        for (int i = c.lane; i < L.N; i += 32) c.kf(K_Q)[i] = 0.0;
        __syncwarp();
    }
    // smem -> HBM: bulk store of the resident prefix
    fence_proxy_async();
    __syncwarp();
    if (c.lane == 0) {
        bulk_s2g(blob, smem, (uint32_t)L.resident);
        bulk_commit();
        bulk_wait_all();
    }
}

constexpr int kThreads = 32;    // one warp per economy
// economies resident per SM that the register budget is cut for (__launch_bounds__): 28 x 32 threads
// x 72 registers fill the register file; shared memory (8.2 KB per default-size economy) allows the same
constexpr int kMaxPerSM = 28;
typedef void (*KernelFn)(const KArgs);

template <int ZM, bool MASKED, int MODE, bool STRICT = false>
static KernelFn select_group(int g)
{
    switch (g) {
    case 2: return cats_period_kernel<ZM, 2, kMaxPerSM / 2, MASKED, MODE, STRICT>;
    case 4: return cats_period_kernel<ZM, 4, kMaxPerSM / 4, MASKED, MODE, STRICT>;
    case 8: return cats_period_kernel<ZM, 8, kMaxPerSM / 8, MASKED, MODE, STRICT>;
    case 12: return cats_period_kernel<ZM, 12, kMaxPerSM / 12, MASKED, MODE, STRICT>;
    case 14: return cats_period_kernel<ZM, 14, kMaxPerSM / 14, MASKED, MODE, STRICT>;
    default: return cats_period_kernel<ZM, 1, kMaxPerSM, MASKED, MODE, STRICT>;
    }
}

// The instantiations are spread over several translation units so that they compile in parallel
// (cats_inst_*.cu); each hands out its kernels through one of these.
KernelFn kernel_debug(int max_z, int g);    // MODE 0: tape / debug stop / debug image
KernelFn kernel_prod(int max_z, int g);     // MODE 1: any size
KernelFn kernel_static(int g);              // MODE 2: default size, z <= 5
KernelFn kernel_masked(int g);              // MODE 1 + per-economy active mask (resets), any z
KernelFn kernel_strict(int mode, int g);    // CATS_FLAG_STRICT: MODE 0 / 1 (any z), MODE 2 (default configuration)

}  // namespace cats
#endif
