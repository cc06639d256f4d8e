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
    static constexpr int o_dang = o_mask + 2 * MW_ROWS * NW * 4;     // [MW_ROWS] i32: round stamp << 10 | lowest uncertain lane that may still come here
    static constexpr int o_d = o_dang + MW_ROWS * 4;                 // [T] f64: what the lane wants at its terminal
    static constexpr int o_term = o_d + T * 8;                       // [T] i32: the lane's terminal (-1: none)
    static constexpr int bytes = (o_term + T * 4 + 15) & ~15;       // then [F] u64: the demand accumulators of the goods market
    static constexpr int total(int F) { return bytes + ((8 * F + 15) & ~15); }
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
    // demand accumulators in SHARED memory here (the one-warp kernel reduces into the Yd row in L2, fire and forget:
    // but a block barrier waits for outstanding reductions to be acknowledged by L2, and this market has several
    // barriers per round)
    unsigned long long *acc = reinterpret_cast<unsigned long long *>(c.mw + ML::bytes);
    OrderT<C::kStatic ? DefaultLayout::kW + DefaultLayout::kF + DefaultLayout::kN : 0, !C::kProd> ord;
    volatile double *gk = c.gk();
    unsigned *maskb = reinterpret_cast<unsigned *>(c.mw + ML::o_mask);
    // danger table: entries carry the stamp of the round that wrote them.  The stamp counts DOWN, so an atomic min
    // lets this round's entries win over stale ones and nobody has to clean the table between rounds.
    int *dangb = reinterpret_cast<int *>(c.mw + ML::o_dang);
    int stamp = 0x1fffff;
    double *sh_d = reinterpret_cast<double *>(c.mw + ML::o_d);   // (block barriers order the accesses)
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
    } else { ord.k = c.okeys(); ord.bad = c.d.bad; }
    for (int i = tid; i < 2 * MW_ROWS * NW; i += T) maskb[i] = 0u;
    for (int i = tid; i < MW_ROWS; i += T) dangb[i] = 0x7fffffff;
    for (int f = tid; f < F; f += T) acc[f] = 0ull;
    __syncthreads();
    const int nch = (H + T - 1) / T;
    int parity = 0;
    const bool exact_rows = F <= MW_ROWS;   // every seller has a table row of its own
#ifdef CATS_PROFILE_PHASES
    long long tq = 0;
    const bool prof = c.a->cycles && tid == 0 && c.b == 0;
    if (prof) tq = clock64();
#define MWTICK(slot) do { if (prof) { long long t1 = clock64(); c.a->cycles[slot] += t1 - tq; tq = t1; } } while (0)
#else
#define MWTICK(slot) do { } while (0)
#endif

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
        MWTICK(12);
        while (any) {
#ifdef CATS_EMU_COUNT
            if (tid == 0) simt::g_count[0]++;
#endif
            unsigned *mask = maskb + parity * (MW_ROWS * NW);
            int *dang = dangb;
            --stamp;
            // lowest uncertain lane that may still come to the seller of table row r this round (T = nobody)
            auto danger_at = [&](int r) { const int v = dang[r]; return (v >> 10) == stamp ? (v & 1023) : 0x7fffffff; };
            // ---- walk: to the first seller of the list that has stock, registering demand on the way
            if (pend && term >= 0 && !(mkt.yr[term].x > 0.0)) term = TERM_WALK;   // emptied by a lane that committed
            if (term == TERM_WALK) {
                // demand is registered at each seller passed and at the terminal.  The head of the list first (it
                // usually has stock); if not, the stock of every seller left, read at once (independent loads)
                int s0;
                if (!cl.pop(s0)) term = TERM_NONE;
                else {
                    const double2 yr0 = mkt.yr[s0];
                    fx_add(&acc[s0], xb);
                    if (yr0.x > 0.0) { term = s0; d_t = b * yr0.y; }
                    else {
                        auto rest = cl;
                        int sl[ZM - 1]; double st[ZM - 1];
#pragma unroll
                        for (int k = 0; k < ZM - 1; ++k) { sl[k] = -1; st[k] = 0.0; if (rest.pop(sl[k])) st[k] = mkt.yr[sl[k]].x; else sl[k] = -1; }
                        term = TERM_NONE;
                        int nvis = 0;
#pragma unroll
                        for (int k = 0; k < ZM - 1; ++k) {
                            if (term == TERM_NONE && sl[k] >= 0) {
                                int dummy;
                                cl.pop(dummy);
                                nvis = k + 1;
                                if (st[k] > 0.0) term = sl[k];
                            }
                        }
                        if (term >= 0) d_t = b * mkt.yr[term].y;
                        // demand at the sellers visited: the low words first (independent atomics in flight
                        // together), then the carries into the high words (see fx_add)
                        const unsigned lo = (unsigned)xb, hi = (unsigned)(xb >> 32);
                        unsigned old[ZM - 1];
#pragma unroll
                        for (int k = 0; k < ZM - 1; ++k)
                            if (k < nvis) old[k] = atomicAdd(reinterpret_cast<unsigned *>(&acc[sl[k]]), lo);
#pragma unroll
                        for (int k = 0; k < ZM - 1; ++k)
                            if (k < nvis) {
                                const unsigned h2 = hi + ((old[k] + lo < old[k]) ? 1u : 0u);
                                if (h2) atomicAdd(reinterpret_cast<unsigned *>(&acc[sl[k]]) + 1, h2);
                            }
                    }
                }
            }
            if (pend && term == TERM_NONE) pend = false;   // nothing left to visit: what is left of b is saved
            const bool has = pend;
            const int row = term & (MW_ROWS - 1);
            if (has) { atomicOr(&mask[row * NW + warp], 1u << lane); sh_d[tid] = d_t; }
            sh_term[tid] = has ? term : -1;
            __syncthreads();
            MWTICK(13);
            // ---- replay the lower lanes at my terminal, in lane order
            int status = ST_IDLE, nxt = 0x7fffffff;
            double ys = 0.0;
            if (has) {
                ys = mkt.yr[term].x;
                bool sold = false;
                unsigned mw_[NW];
#pragma unroll
                for (int w2 = 0; w2 < NW; ++w2) mw_[w2] = mask[row * NW + w2];
                // (exact rows) the next lane above me at this seller: the last lane that commits writes the stock
#pragma unroll
                for (int w2 = NW - 1; w2 >= 0; --w2) {
                    const unsigned up = w2 > warp ? mw_[w2] : (w2 == warp ? mw_[w2] & ~(lt | (1u << lane)) : 0u);
                    if (up) nxt = w2 * 32 + __ffs(up) - 1;
                }
#pragma unroll
                for (int w2 = 0; w2 < NW; ++w2) {
                    unsigned m = w2 < warp ? mw_[w2] : (w2 == warp ? mw_[w2] & lt : 0u);
                    if (exact_rows) {
                        while (m && !sold) {   // two peers' demands in flight per step
                            const int p0 = w2 * 32 + __ffs(m) - 1; m &= m - 1;
                            const double d0 = sh_d[p0];
                            double d1 = 0.0; const bool two = m != 0u;
                            if (two) { const int p1 = w2 * 32 + __ffs(m) - 1; m &= m - 1; d1 = sh_d[p1]; }
                            if (ys > d0) ys = ys - d0; else sold = true;
                            if (two && !sold) { if (ys > d1) ys = ys - d1; else sold = true; }
                        }
                    } else {
                        while (m && !sold) {
                            const int p = w2 * 32 + __ffs(m) - 1;
                            m &= m - 1;
                            if (sh_term[p] == term) {
                                const double dp = sh_d[p];
                                if (ys > dp) ys = ys - dp; else sold = true;
                            }
                        }
                    }
                }
                status = sold ? ST_SOLD : (ys > d_t ? ST_FULL : ST_EXHAUST);
            }
            // ---- who is uncertain, which sellers are in danger (closed by iteration)
            bool unc = status >= ST_EXHAUST;
            auto publish = [&]() {
                auto rest = cl;
                int s;
                while (rest.pop(s)) atomicMin(&dang[s & (MW_ROWS - 1)], (stamp << 10) | tid);
            };
            if (unc) publish();
            if (__syncthreads_or(unc)) {
                while (true) {
                    const bool newly = status == ST_FULL && !unc && danger_at(row) < tid;
                    if (newly) { unc = true; publish(); }
#ifdef CATS_EMU_COUNT
                    if (tid == 0) simt::g_count[1]++;
#endif
                    if (!__syncthreads_or(newly)) break;
                }
            }
            MWTICK(14);
            // ---- commit
            // The committed lanes at a seller are a prefix of its queue.  Exact rows: the last of them -- the lane
            // whose next peer is endangered, or has none -- stores the stock; hashed rows: an atomic min over them.
            const int dmin = status == ST_FULL || status == ST_EXHAUST ? danger_at(row) : 0x7fffffff;
            if (status == ST_FULL && !unc) {
                if (!exact_rows)
                    atomicMin(reinterpret_cast<unsigned long long *>(&mkt.yr[term].x), (unsigned long long)__double_as_longlong(ys - d_t));
                else if (nxt == 0x7fffffff || dmin < nxt) mkt.yr[term].x = ys - d_t;
                b = 0.0; pend = false;
            } else if (status == ST_EXHAUST && !(dmin < tid)) {
                b = b - ys * mkt.p[term];
                if (!exact_rows) atomicMin(reinterpret_cast<unsigned long long *>(&mkt.yr[term].x), 0ull);
                else mkt.yr[term].x = 0.0;
                if (b > 0.0) { term = TERM_WALK; xb = fx_from(b * rpavg); } else { term = TERM_NONE; pend = false; }
            }
            any = __syncthreads_or(pend);
            // ---- leave this round's tables clean (they are used again two rounds from now)
            if (has) mask[row * NW + warp] = 0u;
            parity ^= 1;
            MWTICK(15);
        }
        if (h >= 0) c.PA(h) = pa + b;   // involuntary saving
    }
    __syncthreads();
    {
        double *Yd = c.cf(C_Yd), *Q = c.cf(C_Q);
        for (int f = tid; f < F; f += T) {
            const double2 r = mkt.yr[f];
            const unsigned long long a = acc[f];
            Yd[f] = (fx_to(a) * gk[G_price]) * r.y; Q[f] = c.cf(C_Y_prev)[f] - r.x;
        }
    }
    __syncthreads();
#undef MWTICK
}



// ---------------------------------------------------------------------------------------------
// The goods market as a PIPELINE (default-size economies): warp 0 commits, the other warps prepare.
//
// At F ~ 100 sellers a block-wide round costs 3-4 warp-wide ones (five barrier-separated stages against
// shuffles), so the wide market of mw_goods_market only pays for large F.  Here the market keeps the one-warp
// kernel's commit rounds (goods_commit_chunk: warp collectives, rule R2, ~67 rounds a period) and takes the
// household preparation -- visiting order, record gather, income and budget, Philox candidates, price sort: 40 % of
// the market's instructions, none of them dependent on the sellers' stock -- OFF the chain: producer warps fill a
// ring of prepared chunks in shared memory, the consumer warp only commits.  Hand-over by mbarriers (full / empty per
// slot, one arrival each): the pattern of a TMA pipeline with warps as the producers.
// Every producer owns TWO slots and alternates between them, so it is never more than one barrier phase ahead of
// the consumer (a one-bit phase parity cannot tell two phases apart).
template <int S>
struct PCSlots {   // lives in the MWLayout region (the wide market's tables are not used in this mode)
    uint64_t full[S], empty[S];
    int h[S][32];
    double pa[S][32], b[S][32];
    unsigned long long w0[S][32], w1[S][32];
};
static_assert(sizeof(PCSlots<2>) <= MWLayout<2>::bytes && sizeof(PCSlots<6>) <= MWLayout<4>::bytes &&
              sizeof(PCSlots<14>) <= MWLayout<8>::bytes, "the ring must fit the multi-warp region");

template <int ZM, class C>
__device__ void mw_goods_market_pc(C &c)
{
    constexpr int NW = C::kWarps, P = NW - 1, S = 2 * P;
    const auto &L = c.layout();
    const int W = L.W, F = L.F, H = L.H, lane = c.lane, warp = c.warp;
    const int z = c.z_c() < F ? c.z_c() : F;
    Mkt mkt(c.sm + L.s_U, F);
    unsigned long long *acc = reinterpret_cast<unsigned long long *>(c.cf(C_Yd));   // GLOBAL, zeroed by produce
    PCSlots<S> *ring = reinterpret_cast<PCSlots<S> *>(c.mw);
    // chunk ch is producer (ch % P)'s k-th chunk, k = ch / P: slot (ch % P) + (k & 1) * P, barrier phase k >> 1
    OrderT<C::kStatic ? DefaultLayout::kW + DefaultLayout::kF + DefaultLayout::kN : 0, !C::kProd> ord;
    if (warp == 0) {
        ord.init(c.d, MKT_GOODS, c.a->od[MKT_GOODS], c.okeys(), lane);
        goods_peer_table_init(c, lane, 32);
        if (lane == 0) {
            goods_constants(c);
            for (int s = 0; s < S; ++s) { mbar_init(&ring->full[s], 1); mbar_init(&ring->empty[s], 1); }
            fence_mbar_init();
        }
    } else ord.k = c.okeys();
    __syncthreads();
    const int nch = (H + 31) / 32;
    if (warp == 0) {
        // ---- consumer: chunk after chunk in visiting order
        for (int ch = 0; ch < nch; ++ch) {
            const int k = ch / P, s = ch % P + (k & 1) * P;
            mbar_wait(&ring->full[s], (uint32_t)((k >> 1) & 1));
            const int h = ring->h[s][lane];
            const double pa = ring->pa[s][lane], b = ring->b[s][lane];
            GoodsList<C> cl;
            cl.w0 = (decltype(cl.w0))ring->w0[s][lane]; cl.w1 = (decltype(cl.w1))ring->w1[s][lane];
            __syncwarp();
            if (lane == 0) mbar_arrive(&ring->empty[s]);
            goods_commit_chunk<ZM>(c, mkt, acc, h, pa, b, cl);
        }
    } else {
        // ---- producers: warp p prepares chunks p, p + P, ... (each one chunk of loads ahead)
        const int p = warp - 1;
        int h_n = -1, oc_n = -1; double pa_n = 0.0, pm_n = 0.0, dinc_n = 0.0;
        auto fetch = [&](int ch) {
            const int pos = ch * 32 + lane;
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
        if (p < nch) fetch(p);
        for (int ch = p; ch < nch; ch += P) {
            const int h = h_n, oc = oc_n;
            double pa = pa_n;
            const double pm = pm_n, dinc = dinc_n;
            if (ch + P < nch) fetch(ch + P);
            double b;
            GoodsList<C> cl;
            goods_prepare<ZM>(c, mkt, z, h, oc, dinc, pa, pm, b, cl);
            const int k = ch / P, s = p + (k & 1) * P;
            mbar_wait(&ring->empty[s], (uint32_t)(((k >> 1) & 1) ^ 1));   // a fresh barrier counts as emptied once
            ring->h[s][lane] = h; ring->pa[s][lane] = pa; ring->b[s][lane] = b;
            ring->w0[s][lane] = (unsigned long long)cl.w0; ring->w1[s][lane] = (unsigned long long)cl.w1;
            __syncwarp();
            if (lane == 0) mbar_arrive(&ring->full[s]);
        }
    }
    __threadfence();
    __syncthreads();
    // the ring's memory serves other phases until the next market: its barriers must not stay valid objects
    if (c.tid < S) { mbar_inval(&ring->full[c.tid]); mbar_inval(&ring->empty[c.tid]); }
    {
        volatile double *gk = c.gk();
        double *Yd = c.cf(C_Yd), *Q = c.cf(C_Q);
        for (int f = c.tid; f < F; f += 32 * NW) {
            const double2 r = mkt.yr[f];
            const unsigned long long a = __ldcg(&acc[f]);
            Yd[f] = (fx_to(a) * gk[G_price]) * r.y; Q[f] = c.cf(C_Y_prev)[f] - r.x;
        }
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// The phases around the goods market, split over the warps of the block.


// warp w's 32-aligned share [lo, hi) of a population of n (reversed: the last warp takes the first share, so that a
// small second population -- the K-firms -- lands on a warp that has no C-firms)
template <class C>
__device__ __forceinline__ void mw_range(const C &c, int n, bool reversed, int &lo, int &hi)
{
    constexpr int NW = C::kWarps;
    const int per = ((((n + 31) >> 5) + NW - 1) / NW) << 5;
    const int w = reversed ? NW - 1 - c.warp : c.warp;
    lo = w * per; if (lo > n) lo = n;
    hi = lo + per; if (hi > n) hi = n;
}
// block scratch outside the markets: the visiting-order key block (10 ints) and the goods-market constants (9 doubles)
template <class C> __device__ __forceinline__ volatile int *mw_iscratch(const C &c) { return reinterpret_cast<volatile int *>(c.okeys()); }
template <class C> __device__ __forceinline__ volatile double *mw_dscratch(const C &c) { return c.gk(); }

// a3-a8 (phases 3..10): every warp its share of the firms; the new credit is summed by warp 0 from the Ftot rows
template <class C>
__device__ void mw_firm_decisions(C &c)
{
    const auto &L = c.layout();
    int f_lo, f_hi, k_lo, k_hi;
    mw_range(c, L.F, false, f_lo, f_hi);
    mw_range(c, L.N, true, k_lo, k_hi);
    double accC = 0.0, accK = 0.0;
    firm_decisions_loops(c, f_lo, f_hi, k_lo, k_hi, accC, accK);
    __syncthreads();
    if (c.warp == 0) {
        const double *Fc = c.at(L.s_cFtot), *Fk = c.at(L.s_kFtot);
        double sC = 0.0, sK = 0.0;
        for (int f = c.lane; f < L.F; f += 32) sC = sC + Fc[f];
        for (int i = c.lane; i < L.N; i += 32) sK = sK + Fk[i];
        sC = bfly(sC); sK = bfly(sK);
        if (c.lane == 0) {
            double loans = c.scal()[S_bank_loans] + sC;
            loans = loans + sK;
            c.scal()[S_bank_loans] = loans;
        }
    }
}

// a9 (phases 11, 12): as phase_fire; the passes over the workers and over the pool are the block's, warp 0 cuts the
// batches.  Ends with every thread past the last barrier.
template <class C>
__device__ void mw_fire(C &c)
{
    constexpr int T = C::NT;
    const auto &L = c.layout();
    const int W = L.W, limit = L.F + L.N, tid = c.tid;
    int *ic = c.iat(L.s_ic);
    bool over = false;
    for (int f = tid; f < limit; f += T) { const bool o = c.vac()[f] < 0; over |= o; ic[f] = o ? c.Leff()[f] : 0; }
    if (!__syncthreads_or(over)) return;
    const int cap = L.U_bytes / 8;
    int *pool_w = c.iat(L.s_U);
    unsigned *pool_k = reinterpret_cast<unsigned *>(c.sm + L.s_U) + cap;
    volatile int *scr = mw_iscratch(c);
    int f0 = 0;
    while (f0 < limit) {
        if (c.warp == 0) {
            int f1;
            const int total = fire_batch(c, ic, f0, limit, cap, f1);
            if (c.lane == 0) { scr[0] = f1; scr[1] = total; }
            __syncwarp();
            if (f1 == f0) fire_oversize(c, f0);   // one firm larger than the pool (slow, rare)
        }
        __syncthreads();
        const int f1 = scr[0], total = scr[1];
        if (f1 == f0) { f0 = f0 + 1; __syncthreads(); continue; }
        if (total > 0) {
            for (int w0 = tid; w0 < W; w0 += 4 * T) {
                int ff[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) { const int w = w0 + u * T; ff[u] = w < W ? c.Oc()[w] : -1; }   // 4 loads in flight
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int f = ff[u], w = w0 + u * T;
                    if (f >= f0 && f < f1 && c.vac()[f] < 0) {
                        const int slot = atomicAdd(&ic[f], 0x10000) >> 16;
                        pool_w[slot] = w;
                    }
                }
            }
            __syncthreads();
            for (int i = tid; i < total; i += T) pool_k[i] = c.d.fire_key(pool_w[i]);   // dense: one draw per pooled worker
            __syncthreads();
            for (int i = tid; i < total; i += T) {
                const int w = pool_w[i]; const unsigned key = pool_k[i];
                const int f = c.Oc()[w];
                const int v = ic[f];
                const int end = v >> 16, beg = end - (v & 0xffff);
                int rank = 0;
                for (int j = beg; j < end; ++j) {
                    const unsigned kj = pool_k[j]; const int wj = pool_w[j];
                    rank += (kj < key || (kj == key && wj < w)) ? 1 : 0;
                }
                if (rank < -c.vac()[f]) c.Oc()[w] = -1;
            }
            __syncthreads();
        }
        for (int f = f0 + tid; f < f1; f += T)
            if (c.vac()[f] < 0) { c.Leff()[f] = c.Leff()[f] + c.vac()[f]; c.vac()[f] = 0; }
        __syncthreads();
        f0 = f1;
    }
}

// a10 (phase 13): the block lists the unemployed (any order), asks the visiting order for their positions and ranks
// them; warp 0 runs the matching rounds.  More seekers than the block ranks (start-up periods): the one-warp path.
template <int ZM, class C>
__device__ void mw_job_search(C &c)
{
    constexpr int NW = C::kWarps, T = C::NT;
    const auto &L = c.layout();
    const int W = L.W, tid = c.tid, lane = c.lane;
    OrderT<C::kStatic ? DefaultLayout::kW : 0, !C::kProd> ord;
    if (c.warp == 0) ord.init(c.d, MKT_JOB, c.a->od[MKT_JOB], c.okeys(), lane); else ord.k = c.okeys();
    unsigned short *seekers = reinterpret_cast<unsigned short *>(c.sm + L.s_U);
    const int cap3 = L.U_bytes / 6;
    unsigned short *byidx = seekers + cap3, *posv = seekers + 2 * cap3;
    int *cnt = reinterpret_cast<int *>(const_cast<double *>(mw_dscratch(c)));
    if (tid == 0) cnt[0] = 0;
    __syncthreads();
    const int nwords = (W + 31) >> 5;
    for (int word0 = c.warp; word0 < nwords; word0 += 4 * NW) {
        int oc[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { const int w = (word0 + u * NW) * 32 + lane; oc[u] = w < W ? c.Oc()[w] : 0; }   // 4 loads in flight
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int word = word0 + u * NW, w = word * 32 + lane;
            if (word >= nwords) break;
            const bool un = w < W && oc[u] < 0;
            const unsigned bal = __ballot_sync(FULL, un);
            int base = 0;
            if (lane == 0) {
                c.emp()[word] = ~bal;   // employed bits for the goods market (hires are added by the matching rounds)
                if (bal) base = atomicAdd(&cnt[0], __popc(bal));
            }
            base = __shfl_sync(FULL, base, 0);
            const int at = base + __popc(bal & c.lt());
            if (un && at < cap3) byidx[at] = (unsigned short)w;
        }
    }
    __syncthreads();
    const int U = cnt[0];
    // the seekers' visiting positions go into a bitmap of W bits; a seeker's rank in visiting order is the number of
    // bits below its own (a prefix over the words + a population count): O(U + W / 32) instead of U * U compares.
    // The bitmap and the word prefix live in the goods market's tables, idle until then.
    unsigned *bitmap = reinterpret_cast<unsigned *>(c.mw);
    int *prefix = reinterpret_cast<int *>(c.mw) + nwords;
    if (U > cap3 || 8 * nwords > MWLayout<NW>::o_d) {   // start-up periods (everybody looks for a job) / huge W
        if (c.warp == 0) phase_job_search<ZM>(c);
        return;
    }
    for (int i = tid; i < nwords; i += T) bitmap[i] = 0u;
    __syncthreads();
    for (int i = tid; i < U; i += T) {
        const int pos = ord.position_of((int)byidx[i]);
        posv[i] = (unsigned short)pos;
        atomicOr(&bitmap[pos >> 5], 1u << (pos & 31));
    }
    __syncthreads();
    if (c.warp == 0) {
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
    }
    __syncthreads();
    for (int i = tid; i < U; i += T) {
        const int pos = (int)posv[i];
        const int r = prefix[pos >> 5] + __popc(bitmap[pos >> 5] & ((1u << (pos & 31)) - 1u));
        seekers[r] = byidx[i];
    }
    __syncthreads();
    if (c.warp == 0) job_match<ZM>(c, seekers, U, true);
}

// a11 (phases 14, 15): every warp its share of the firms, then the sellers' price classes by the block
template <class C>
__device__ void mw_produce(C &c)
{
    constexpr int T = C::NT;
    const auto &L = c.layout();
    const int F = L.F;
    int lo, hi;
    mw_range(c, F, false, lo, hi);
    produce_c_loop(c, lo, hi);
    mw_range(c, L.N, true, lo, hi);
    produce_k_loop(c, lo, hi);
    __syncthreads();
    Mkt mkt(c.sm + L.s_U, F);
    // price class = number of strictly cheaper sellers.  Few sellers: count (F * F / T compares per thread).  Many
    // (the 10x economy: 1,000): sort the prices in the goods market's tables (idle until then) with a bitonic
    // network -- a positive double orders like its bit pattern -- and a seller's class is the first position of its
    // price in the sorted row.
    int n2 = 1;
    while (n2 < F) n2 <<= 1;
    if (F <= 256 || n2 * 10 > MWLayout<C::kWarps>::bytes) {
        for (int f = c.tid; f < F; f += T) {
            const double pf = mkt.p[f];
            unsigned n = 0u;
#pragma unroll 4
            for (int g = 0; g < F; ++g) n += mkt.p[g] < pf ? 1u : 0u;
            mkt.cls[f] = n;
        }
        return;
    }
    unsigned long long *key = reinterpret_cast<unsigned long long *>(c.mw);
    unsigned short *idx = reinterpret_cast<unsigned short *>(c.mw + 8 * (size_t)n2);
    for (int i = c.tid; i < n2; i += T) {
        key[i] = i < F ? (unsigned long long)__double_as_longlong(mkt.p[i]) : ~0ull;
        idx[i] = (unsigned short)i;
    }
    __syncthreads();
    for (int k = 2; k <= n2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = c.tid; t < (n2 >> 1); t += T) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));   // the lower index of the t-th pair at distance j
                const int p = i | j;
                const bool up = (i & k) == 0;
                const unsigned long long a = key[i], b = key[p];
                if ((a > b) == up) {
                    key[i] = b; key[p] = a;
                    const unsigned short ia = idx[i]; idx[i] = idx[p]; idx[p] = ia;
                }
            }
            __syncthreads();
        }
    }
    for (int r = c.tid; r < F; r += T) {
        int first = r;
        const unsigned long long kr = key[r];
        while (first > 0 && key[first - 1] == kr) --first;   // equal prices share the class (ties are rare)
        mkt.cls[idx[r]] = (unsigned)first;
    }
}


// a12 (phase 16): the block lists the investing firms in visiting order and prepares them (suppliers drawn and
// price-sorted) into records; ONE thread then trades them one after the other in a tight loop -- the chain is
// sequential by nature (buyer i sees the stock buyers < i left), and a single thread on prepared records walks it
// without a warp hand-over per buyer; the block scatters the results back.  Records live in the goods market's
// tables (idle until then), in batches if need be.
template <int ZM>
struct CapRec {
    int f, n;
    double kd, cb, lq, iv, vi;
    int cand[ZM];
    double pr[ZM];
};

template <int ZM, class C>
__device__ void mw_capital_market(C &c)
{
    constexpr int NW = C::kWarps, T = C::NT;
    const auto &L = c.layout();
    const int F = L.F, N = L.N, lane = c.lane, tid = c.tid;
    const int z = c.z_k() < N ? c.z_k() : N;
    OrderT<C::kStatic ? DefaultLayout::kF : 0, !C::kProd> ord;
    if (c.warp == 0) ord.init(c.d, MKT_CAP, c.a->od[MKT_CAP], c.okeys(), lane); else ord.k = c.okeys();
    double *Pk = c.kf(K_P), *Ydk = c.kf(K_Yd), *Ysk = c.at(L.s_kYstock);
    const double *RPk = c.kf(K_Q);   // 1 / P, parked there by produce
    double *Kdem = c.at(L.s_cKdem), *Ftot = c.at(L.s_cFtot), *liq = c.cf(C_liquidity);
    double *inv = c.cf(C_investment), *valinv = c.at(L.s_cvalinv), *wages = c.cf(C_wages);
    int *buyers = c.iat(L.s_ic);
    const int nchunks = (F + 31) >> 5;
    unsigned *bal = reinterpret_cast<unsigned *>(c.mw);
    int *pre = reinterpret_cast<int *>(c.mw) + nchunks;
    const int rec_off = (8 * nchunks + 4 + 15) & ~15;
    volatile int *tot = pre + nchunks;
    CapRec<ZM> *rec = reinterpret_cast<CapRec<ZM> *>(c.mw + rec_off);
    const int cap = (MWLayout<NW>::bytes - rec_off) / (int)sizeof(CapRec<ZM>);
    if (cap < 1) {   // a population so large that not even the ballots fit: the one-warp code
        __syncthreads();
        if (c.warp == 0) phase_capital_market<ZM>(c);
        return;
    }
    __syncthreads();
    for (int chunk = c.warp; chunk < nchunks; chunk += NW) {
        const int pos = chunk * 32 + lane;
        bool act = false;
        if (pos < F) act = Kdem[ord.at(pos)] > 0.0;
        const unsigned b = __ballot_sync(FULL, act);
        if (lane == 0) bal[chunk] = b;
    }
    __syncthreads();
    if (c.warp == 0) {
        int run = 0;
        for (int base = 0; base < nchunks; base += 32) {
            const int i = base + lane;
            const int n = i < nchunks ? __popc(bal[i]) : 0;
            int incl = n;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) { const int t = __shfl_up_sync(FULL, incl, off); if (lane >= off) incl += t; }
            if (i < nchunks) pre[i] = run + incl - n;
            run += __shfl_sync(FULL, incl, 31);
        }
        if (lane == 0) tot[0] = run;
    }
    __syncthreads();
    for (int chunk = c.warp; chunk < nchunks; chunk += NW) {
        const unsigned b = bal[chunk];
        if ((b >> lane) & 1u) buyers[pre[chunk] + __popc(b & c.lt())] = ord.at(chunk * 32 + lane);
    }
    const int total = tot[0];
    __syncthreads();
    for (int b0 = 0; b0 < total; b0 += cap) {
        const int nb = total - b0 < cap ? total - b0 : cap;
        for (int i = tid; i < nb; i += T) {
            const int f = buyers[b0 + i];
            const double lq = liq[f];
            int cand[ZM]; double pr[ZM];
#pragma unroll
            for (int k = 0; k < ZM; ++k) { cand[k] = 0; pr[k] = __longlong_as_double(0x7ff0000000000000LL); }
            c.d.template candidates<ZM>(PH_CAP, f, N, z, cand);
#pragma unroll
            for (int k = 0; k < ZM; ++k) if (k < z) pr[k] = Pk[cand[k]];
            sort_by_price<ZM>(cand, pr, z);
            CapRec<ZM> &r = rec[i];
            r.f = f; r.n = z; r.kd = Kdem[f]; r.cb = (Ftot[f] + lq) - wages[f]; r.lq = lq; r.iv = 0.0; r.vi = 0.0;
#pragma unroll
            for (int k = 0; k < ZM; ++k) { r.cand[k] = cand[k]; r.pr[k] = pr[k]; }
        }
        __syncthreads();
        if (c.warp == 0) {
            // one warp trades the records, 32 at a time, in rounds of buyers that share no supplier (capital_go)
            for (int base = 0; base < nb; base += 32) {
                const int i = base + lane;
                const bool act = i < nb;
                int cand[ZM]; double pr[ZM];
                double kd = 0.0, cb = 0.0, lq = 0.0;
#pragma unroll
                for (int k = 0; k < ZM; ++k) { cand[k] = 0; pr[k] = 0.0; }
                if (act) {
                    const CapRec<ZM> &r = rec[i];
                    kd = r.kd; cb = r.cb; lq = r.lq;
#pragma unroll
                    for (int k = 0; k < ZM; ++k) { cand[k] = r.cand[k]; pr[k] = r.pr[k]; }
                }
                unsigned m = __ballot_sync(FULL, act);
                using SetT = CapitalSet<C>;
                const SetT cset = act ? capital_set<SetT, ZM>(cand, z) : (SetT)0;
                while (m) {
                    const bool go = capital_go(m, cset, lane);
                    m &= ~__ballot_sync(FULL, go);
                    if (go) {
                        double iv = 0.0, vi = 0.0;
                        capital_trade<ZM, false>(z, cand, pr, RPk, Ydk, Ysk, kd, cb, lq, iv, vi);
                        CapRec<ZM> &r = rec[i];
                        r.kd = kd; r.lq = lq; r.iv = iv; r.vi = vi;
                    }
                    __syncwarp();
                }
            }
        }
        __syncthreads();
        for (int i = tid; i < nb; i += T) {
            const CapRec<ZM> &r = rec[i];
            const int f = r.f;
            liq[f] = r.lq; inv[f] = r.iv; valinv[f] = r.vi; Kdem[f] = r.kd;
        }
        __syncthreads();
    }
}

// a17 (phases 21, 22): every warp its share; warp 0 adds the six sums up from the rows the loops staged
template <class C>
__device__ void mw_accounting(C &c)
{
    const auto &L = c.layout();
    const int F = L.F, N = L.N;
    int f_lo, f_hi, k_lo, k_hi;
    mw_range(c, F, false, f_lo, f_hi);
    mw_range(c, N, true, k_lo, k_hi);
    double acc[6];
    accounting_loops<true>(c, f_lo, f_hi, k_lo, k_hi, acc);
    __syncthreads();
    if (c.warp == 0) {
        const double *c_in = c.at(L.s_cKdem), *c_int = c.cf(C_interests), *c_dt = c.at(L.s_cvalinv);
        const double *k_inr = c.at(L.s_kFtot), *k_intr = c.kf(K_interests), *k_dtr = c.at(L.s_kYstock);
        double s_in = 0.0, s_int = 0.0, s_dt = 0.0, k_in = 0.0, k_int = 0.0, k_dt = 0.0;
        for (int f = c.lane; f < F; f += 32) { s_in = s_in + c_in[f]; s_int = s_int + c_int[f]; s_dt = s_dt + c_dt[f]; }
        for (int i = c.lane; i < N; i += 32) { k_in = k_in + k_inr[i]; k_int = k_int + k_intr[i]; k_dt = k_dt + k_dtr[i]; }
        s_in = bfly(s_in); s_int = bfly(s_int); s_dt = bfly(s_dt);
        k_in = bfly(k_in); k_int = bfly(k_int); k_dt = bfly(k_dt);
        if (c.lane == 0) accounting_book(c, s_in, s_int, s_dt, k_in, k_int, k_dt);
        __syncwarp();
    }
}

// One economy per block, NW warps.  MODE as in cats_period_kernel (1 = production, any size; 2 = the
// default configuration with the layout resolved at compile time).  The matching chains of the labour and
// capital-goods markets, the reward and the aggregate tracking are warp 0's (the one-warp code); everything else
// is the block's.
// PIPE: the goods market as the producer / consumer pipeline (mw_goods_market_pc) instead of the block-wide one
template <int ZM, int NW, int MINB, int MODE, bool PIPE>
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
    c.d.bad = &c.misc()[M_STATUS];
    unsigned launch_status = 0u;
#ifdef CATS_PROFILE_PHASES
    long long tk = 0;
    const bool kprof = a.cycles && c.tid == 0 && b == 0;
    if (kprof) tk = clock64();
#define KTICK(slot) do { if (kprof) { long long t1 = clock64(); a.cycles[slot] += t1 - tk; tk = t1; } } while (0)
#else
#define KTICK(slot) do { } while (0)
#endif
    for (int t = 0; t < a.n_periods; ++t) {
        c.d.period = (uint32_t)c.scal()[S_timestep];
        c.d.epi = (uint32_t)c.misc()[M_EPISODE];
        __syncthreads();      // everybody has read the period before thread 0 moves it on
        if (c.tid == 0) {
            c.scal()[S_gov_TA] = 0.0; c.scal()[S_gov_G] = 0.0;
            c.scal()[S_bank_interest_income] = 0.0; c.scal()[S_bank_bad_debt] = 0.0;
            c.scal()[S_gov_r_b] = c.par()[P_interest_rate];
            c.misc()[M_STATUS] = 0;
        }
        KTICK(11);
        mw_firm_decisions(c);
        __syncthreads();
        KTICK(0);
        mw_fire(c);
        __syncthreads();
        KTICK(1);
        mw_job_search<ZM>(c);
        __syncthreads();
        KTICK(2);
        mw_produce(c);
        __syncthreads();
        KTICK(3);
        mw_capital_market<ZM>(c);
        KTICK(4);
        if (c.warp == 0) {
            if (c.lane == 0) bulk_prefetch_l2(c.blob + L.o_PA, (uint32_t)(L.blob - L.o_PA));
            phase_gov_income(c);
            KTICK(5);
        }
        __syncthreads();
        if constexpr (PIPE) mw_goods_market_pc<ZM>(c); else mw_goods_market<ZM>(c);
        KTICK(6);
        mw_accounting(c);
        KTICK(7);
        if (c.warp == 0) {
            phase_reward(c);
            __syncwarp();
            KTICK(8);
            phase_tracking(c);
            KTICK(9);
            const int nbad = bankrupt_replace(c);
            if (c.lane == 0) mw_iscratch(c)[0] = nbad;
        }
        __syncthreads();
        if (mw_iscratch(c)[0] > 0) bankrupt_release(c, c.tid, 32 * NW);
        if (c.warp == 0) {
            bankrupt_terminated(c);
            __syncwarp();
            KTICK(10);
            if (c.lane == 0) mw_dscratch(c)[0] = closing_scalars(c);
        }
        __syncthreads();
        {
            const double share = mw_dscratch(c)[0];
            if (share > 0.0)
                for (int h = L.W + c.tid; h < L.H; h += 32 * NW) c.PA(h) = c.PA(h) + share;
        }
        if (c.warp == 0) {
            KTICK(11);
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

KernelFn kernel_mw(int mode, int max_z, int nw, bool pipe);   // cats_inst_mw.cu
int mw_extra_bytes(int nw, int F);

}  // namespace cats
#endif
