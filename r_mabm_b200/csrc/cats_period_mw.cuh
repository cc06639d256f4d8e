// cats_period_mw.cuh -- the period kernel for UNDER-FILLED launches: one thread block = one economy,
// NW warps share it (north_star: "one thread block per economy with agents on warps").
//
// The one-warp-per-economy kernel (cats_period.cuh) fills a B200 from ~4,100 economies up.  Below that
// -- BASELINE.json config 2 (1,024 economies), config 3 split over 8 GPUs (2,048 each), config 5 (256
// economies of 10x size) -- the launch time is ONE economy's dependent chain, 60 % of it the
// consumption-goods market.  Here NW warps work on one economy:
//
//   * the goods market takes NW * 32 households per chunk and commits them with a rule that does not
//     stop at the first seller that sells out (R2 below): the chain is ~20 wide rounds per period
//     instead of ~100 warp-wide ones (scripts/round_model.py is the model the rule was worked out on);
//   * the per-firm and per-worker phases split over the warps (contiguous 32-aligned ranges per warp,
//     so rows stay coalesced); the fixed-order fp64 sums (det_sum, MODEL.md 4.3) are taken by ONE
//     warp, in the oracle's lane order, from rows the parallel pass left in shared memory.
//
// Results are bit-identical to the one-warp kernel and the oracle (tests/test_kernel_emulated.py on the
// CPU emulation, tests/test_multiwarp_gpu.py on the device).
//
// R2, the commit rule.  Every pending household ("lane") has a terminal: the first seller of its
// price-sorted list that still has stock.  Lanes with the same terminal replay the lower lanes'
// purchases in lane (= visiting) order and so know the stock they will find: FULL (buys all it wants),
// EXHAUST (buys what is left, walks on with the rest of its budget) or SOLD (a lower lane empties
// the seller first; walks on).  EXHAUST and SOLD lanes are "uncertain": where they buy next is not
// known yet -- but it can only be one of the sellers left on their lists.  A lane is endangered when
// its terminal is on the remaining list of an uncertain LOWER lane; an endangered FULL lane is
// uncertain in turn (it may find less stock than it replayed, sell the seller out and walk on), so
// the set is closed by iteration.  Every lane that is not endangered commits: FULL lanes buy, an
// EXHAUST lane takes the rest of the stock and walks on in the next round.  The committed lanes at
// a seller are always a prefix of its queue, so the stock after the round is the minimum over their
// "stock after my purchase" (an atomic min).  The lowest pending lane always commits: rounds end.
#ifndef CATS_PERIOD_MW_CUH
#define CATS_PERIOD_MW_CUH
#include "cats_period.cuh"

namespace cats {

constexpr int MW_ROWS = 256;   // rows of the peer table and of the danger table (seller id hashed by its low bits)

// shared memory of the multi-warp kernel after the economy's working set (the slice of the one-warp kernel)
template <int NW>
struct MWLayout {
    static constexpr int T = 32 * NW;
    static constexpr int o_mask = 0;                                 // [2][MW_ROWS][NW] u32: lanes whose terminal hashes to the row
    static constexpr int o_dang = o_mask + 2 * MW_ROWS * NW * 4;     // [2][MW_ROWS] i32: lowest uncertain lane that may still come here
    static constexpr int o_d = o_dang + 2 * MW_ROWS * 4;             // [T] f64: what the lane wants at its terminal
    static constexpr int o_term = o_d + T * 8;                       // [T] i32: the lane's terminal (-1: none)
    static constexpr int bytes = (o_term + T * 4 + 15) & ~15;
};

template <int MODE, int NW>
struct CtxMW : CtxT<MODE, false> {
    static constexpr int kWarps = NW;
    static constexpr int NT = 32 * NW;
    int warp, tid;
    unsigned char *mw;   // MWLayout region
};

// a14 + a15 + a16 for NW warps: see the header comment
template <int ZM, class C>
__device__ void mw_goods_market(C &c)
{
    constexpr int NW = C::kWarps, T = 32 * NW;
    using ML = MWLayout<NW>;
    const auto &L = c.layout();
    const int W = L.W, F = L.F, H = L.H, lane = c.lane, warp = c.warp, tid = c.tid;
    const unsigned lt = c.lt();
    const int z = c.z_c() < F ? c.z_c() : F;
    Mkt mkt(c.sm + L.s_U, F);
    unsigned long long *acc = reinterpret_cast<unsigned long long *>(c.cf(C_Yd));   // GLOBAL, zeroed by produce
    OrderT<C::kStatic ? DefaultLayout::kW + DefaultLayout::kF + DefaultLayout::kN : 0, !C::kProd> ord;
    enum { G_net, G_sub, G_ir_emp, G_ir_un, G_xi, G_omxi, G_chi, G_price, G_rpavg };
    volatile double *gk = c.gk();
    unsigned *maskb = reinterpret_cast<unsigned *>(c.mw + ML::o_mask);
    int *dangb = reinterpret_cast<int *>(c.mw + ML::o_dang);
    volatile double *sh_d = reinterpret_cast<volatile double *>(c.mw + ML::o_d);
    volatile int *sh_term = reinterpret_cast<volatile int *>(c.mw + ML::o_term);
    if (warp == 0) {
        ord.init(c.d, MKT_GOODS, c.a->od[MKT_GOODS], c.okeys(), lane);
        if (lane == 0) {
            const double wb = c.scal()[S_wb], sub = c.scal()[S_gov_subsidy_level], price = c.scal()[S_price];
            const double xi = c.par()[P_xi];
            const double net = wb * (1.0 - c.par()[P_tax_rate]);
            const double rp = 1.0 / price;
            gk[G_net] = net; gk[G_sub] = sub; gk[G_ir_emp] = net * rp; gk[G_ir_un] = sub * rp;
            gk[G_xi] = xi; gk[G_omxi] = 1.0 - xi; gk[G_chi] = c.par()[P_chi]; gk[G_price] = price; gk[G_rpavg] = rp;
        }
    } else ord.k = c.okeys();
    for (int i = tid; i < 2 * MW_ROWS * NW; i += T) maskb[i] = 0u;
    for (int i = tid; i < 2 * MW_ROWS; i += T) dangb[i] = 0x7fffffff;
    __syncthreads();
    const int nch = (H + T - 1) / T;
    int parity = 0;

    int h_n = -1, oc_n = -1; double pa_n = 0.0, pm_n = 0.0, dinc_n = 0.0;
    auto fetch = [&](int ch) {
        const int pos = ch * T + tid;
        h_n = -1; oc_n = -1; pa_n = 0.0; pm_n = 0.0; dinc_n = 0.0;
        if (pos < H) {
            const int h = ord.at(pos);
            const double2 rec = c.hh()[h];
            h_n = h; pa_n = rec.x; pm_n = rec.y;
            if (h < W) oc_n = ((c.emp()[h >> 5] >> (h & 31)) & 1u) ? 0 : -1;
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
        // a14 + a15: income, permanent income, consumption budget (as phase_goods_market)
        double b = 0.0;
        const double rpavg = gk[G_rpavg];
        if (h >= 0) {
            const double price = gk[G_price];
            double ir;
            if (h < W) { const bool emp = oc >= 0; pa = pa + (emp ? gk[G_net] : gk[G_sub]); ir = emp ? gk[G_ir_emp] : gk[G_ir_un]; }
            else ir = dinc * rpavg;
            pm = pm * gk[G_xi] + gk[G_omxi] * ir;
            const double target = pm + (gk[G_chi] * pa) * rpavg;
            b = target * price;
            if (b > pa) b = pa;
            if (!(b > 0.0)) b = 0.0;
            pa = pa - b;
            c.perm(h) = pm;
        }
        CandList<C::kStatic && DefaultLayout::kF <= 254> cl;
        if (b > 0.0) {
            int cand[ZM]; unsigned key[ZM];
            c.d.template candidates<ZM>(PH_GOODS, h, F, z, cand);
#pragma unroll
            for (int k = 0; k < ZM; ++k)
                key[k] = k < z ? (mkt.cls[cand[k]] << 17) | ((unsigned)k << 14) | (unsigned)cand[k] : 0xffffffffu;
            sort_keys<ZM>(key);
#pragma unroll
            for (int k = 0; k < ZM; ++k)
                if (k < z) cl.set(k, key[k] & 0x3fffu);
        }
        enum { TERM_NONE = -1, TERM_WALK = -2 };
        enum { ST_IDLE = 0, ST_FULL = 1, ST_EXHAUST = 2, ST_SOLD = 3 };
        bool pend = b > 0.0;
        int term = pend ? TERM_WALK : TERM_NONE;
        double d_t = 0.0;
        unsigned long long xb = fx_from(b * rpavg);
        int any = __syncthreads_or(pend);
        while (any) {
            unsigned *mask = maskb + parity * (MW_ROWS * NW);
            int *dang = dangb + parity * MW_ROWS;
            // ---- walk: to the first seller of the list that has stock, registering demand on the way
            if (pend && term >= 0 && !(mkt.yr[term].x > 0.0)) term = TERM_WALK;   // emptied by a lane that committed
            while (term == TERM_WALK) {
                int cnd;
                if (!cl.pop(cnd)) term = TERM_NONE;
                else {
                    const double2 yr = mkt.yr[cnd];
                    fx_red(&acc[cnd], xb);
                    if (yr.x > 0.0) { term = cnd; d_t = b * yr.y; }
                }
            }
            if (pend && term == TERM_NONE) pend = false;   // nothing left to visit: what is left of b is saved
            const bool has = pend;
            const int row = term & (MW_ROWS - 1);
            if (has) { atomicOr(&mask[row * NW + warp], 1u << lane); sh_d[tid] = d_t; }
            sh_term[tid] = has ? term : -1;
            __syncthreads();
            // ---- replay the lower lanes at my terminal, in lane order
            int status = ST_IDLE;
            double ys = 0.0;
            if (has) {
                ys = mkt.yr[term].x;
                bool sold = false;
                for (int w2 = 0; w2 <= warp && !sold; ++w2) {
                    unsigned m = mask[row * NW + w2];
                    if (w2 == warp) m &= lt;
                    while (m) {
                        const int p = w2 * 32 + __ffs(m) - 1;
                        m &= m - 1;
                        if (sh_term[p] == term) {
                            const double dp = sh_d[p];
                            if (ys > dp) ys = ys - dp; else { sold = true; break; }
                        }
                    }
                }
                status = sold ? ST_SOLD : (ys > d_t ? ST_FULL : ST_EXHAUST);
            }
            // ---- who is uncertain, which sellers are in danger (closed by iteration)
            bool unc = status >= ST_EXHAUST, published = false;
            auto publish = [&]() {
                auto rest = cl;
                int s;
                while (rest.pop(s)) atomicMin(&dang[s & (MW_ROWS - 1)], tid);
                published = true;
            };
            if (unc) publish();
            if (__syncthreads_or(unc)) {
                while (true) {
                    const bool newly = status == ST_FULL && !unc && dang[row] < tid;
                    if (newly) { unc = true; publish(); }
                    if (!__syncthreads_or(newly)) break;
                }
            }
            // ---- commit
            if (status == ST_FULL && !unc) {
                atomicMin(reinterpret_cast<unsigned long long *>(&mkt.yr[term].x), (unsigned long long)__double_as_longlong(ys - d_t));
                b = 0.0; pend = false;
            } else if (status == ST_EXHAUST && !(dang[row] < tid)) {
                b = b - ys * mkt.p[term];
                atomicMin(reinterpret_cast<unsigned long long *>(&mkt.yr[term].x), 0ull);
                if (b > 0.0) { term = TERM_WALK; xb = fx_from(b * rpavg); } else { term = TERM_NONE; pend = false; }
            }
            any = __syncthreads_or(pend);
            // ---- leave this round's tables clean (they are used again two rounds from now)
            if (has) mask[row * NW + warp] = 0u;
            if (published) {
                auto rest = cl;
                int s;
                // the list as it was when it was published: a lane that just exhausted its seller has not walked yet
                while (rest.pop(s)) dang[s & (MW_ROWS - 1)] = 0x7fffffff;
            }
            parity ^= 1;
        }
        if (h >= 0) c.PA(h) = pa + b;   // involuntary saving
    }
    __threadfence();
    __syncthreads();
    {
        double *Yd = c.cf(C_Yd), *Q = c.cf(C_Q);
        for (int f = tid; f < F; f += T) {
            const double2 r = mkt.yr[f];
            const unsigned long long a = __ldcg(&acc[f]);
            Yd[f] = (fx_to(a) * gk[G_price]) * r.y; Q[f] = c.cf(C_Y_prev)[f] - r.x;
        }
    }
    __syncthreads();
}

// One economy per block, NW warps.  MODE as in cats_period_kernel (1 = production, any size; 2 = the
// default configuration with the layout resolved at compile time).  Version 1: warp 0 runs the phases
// around the goods market with the one-warp code; the goods market is the block's.
template <int ZM, int NW, int MINB, int MODE>
__global__ void __launch_bounds__(32 * NW, MINB) cats_period_mw_kernel(const __grid_constant__ KArgs a)
{
#ifdef CATS_EMU
    unsigned char *smem = simt::g_blk->smem;
#else
    extern __shared__ __align__(128) unsigned char smem[];
#endif
    static_assert(MODE >= 1, "the multi-warp kernel is a production kernel (Philox draws, whole periods)");
    constexpr bool ST = MODE == 2;
    CtxMW<MODE, NW> c;
    c.a = &a;
    const auto &L = c.layout();
    const int slice = ST ? ((DefaultLayout::smem + 15) & ~15) : a.slice;
    c.sm = smem;
    c.mw = smem + slice;
    c.tid = threadIdx.x; c.lane = threadIdx.x & 31; c.warp = threadIdx.x >> 5;
    c.last = 99;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + L.s_misc + 8);
    const long long b = blockIdx.x;
    if (b >= a.B) return;
    c.b = b;
    unsigned char *blob = a.blobs + (size_t)b * (size_t)L.blob;
    c.blob = blob;
    if (c.tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
        mbar_expect_tx(bar, (uint32_t)L.resident);
        bulk_g2s(smem, blob, (uint32_t)L.resident, bar);
        bulk_prefetch_l2(blob + L.o_Oc, (uint32_t)(L.o_PA - L.o_Oc));
    }
    bool bad_z = false;
    const double *pr = a.params + (a.params_stride ? (size_t)b * (size_t)a.params_stride : 0);
    for (int i = c.tid; i < N_PARAMS; i += 32 * NW) c.par()[i] = pr[i];
    if (c.tid == 0) c.par()[P_log1mtheta] = det_log(1.0 - pr[P_theta]);
    if (ST && c.tid == 0 && ((int)pr[P_z_c] != DefaultLayout::kZc || (int)pr[P_z_k] != DefaultLayout::kZk || (int)pr[P_z_e] != DefaultLayout::kZe))
        bad_z = true;
    __syncthreads();          // the barrier object is initialised before anybody polls it
    mbar_wait(bar, 0);
    if (c.tid == 0) c.misc()[M_EPISODE] = (int)((uint32_t)c.scal()[S_episode] << 12);
    __syncthreads();
    c.d.rk = a.rk;
    c.d.economy = (uint32_t)(a.economy_base + (unsigned long long)b);
    c.d.tp = &a.T;
    c.d.tape_z = a.tape_z;
    unsigned launch_status = 0u;
    for (int t = 0; t < a.n_periods; ++t) {
        c.d.period = (uint32_t)c.scal()[S_timestep];
        c.d.epi = (uint32_t)c.misc()[M_EPISODE];
        __syncthreads();      // everybody has read the period before thread 0 moves it on
        if (c.warp == 0) {
            if (c.lane == 0) {
                c.scal()[S_gov_TA] = 0.0; c.scal()[S_gov_G] = 0.0;
                c.scal()[S_bank_interest_income] = 0.0; c.scal()[S_bank_bad_debt] = 0.0;
                c.scal()[S_gov_r_b] = c.par()[P_interest_rate];
                c.misc()[M_STATUS] = 0;
            }
            __syncwarp();
            phase_firm_decisions(c);
            phase_fire(c, L.F + L.N);
            __syncwarp();
            phase_job_search<ZM>(c);
            __syncwarp();
            phase_produce(c);
            phase_capital_market<ZM>(c);
            if (c.lane == 0) bulk_prefetch_l2(c.blob + L.o_PA, (uint32_t)(L.blob - L.o_PA));
            phase_gov_income(c);
        }
        __syncthreads();
        mw_goods_market<ZM>(c);
        if (c.warp == 0) {
            phase_accounting(c);
            phase_reward(c);
            __syncwarp();
            phase_tracking(c);
            phase_bankrupt(c);
            phase_closing(c);
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
                c.scal()[S_status] = (double)((unsigned)c.scal()[S_status] | (unsigned)c.misc()[M_STATUS]);
            }
        }
        __syncthreads();
    }
    if (c.warp == 0) {
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
        fence_proxy_async();
        __syncwarp();
        if (c.lane == 0) {
            bulk_s2g(blob, smem, (uint32_t)L.resident);
            bulk_commit();
            bulk_wait_all();
        }
    }
}

KernelFn kernel_mw(int mode, int max_z, int nw);   // cats_inst_mw.cu
int mw_extra_bytes(int nw);

}  // namespace cats
#endif
