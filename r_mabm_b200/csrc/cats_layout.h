// cats_layout.h -- economy blob / on-chip working-set layout shared by host and device code.
//
// One economy = one contiguous blob in HBM (docs/MODEL.md section 2, DESIGN.md "data layout"):
//
//   [ scal | hot K-firm rows | Leff |  Oc | C-firm rows | cold K-firm rows | {PA, perm} x H ]
//     `------ resident prefix -----'  `--------------- streamed tail ---------------'
//
// One WARP advances one economy.  The resident prefix (aggregates, K-firm tables, head-counts: what
// the labour and capital-goods markets touch at random) is pulled into shared memory with one TMA
// bulk copy and pushed back with one bulk store.  The tail is streamed through L2: C-firm rows are
// only ever touched by the lane that owns the firm (lane l owns firms l, l+32, ...: coalesced, and
// it makes the fixed-order fp64 reductions register-only), employment links and household wealth /
// permanent income are gathered once per market in visiting order.  The consumption-goods market
// works on a packed per-seller record table in shared memory.  Keeping the tail out of shared
// memory leaves ~10 KB per economy on chip, so residency is bounded by registers, not smem.
// All section offsets are multiples of 16 bytes (bulk-copy granularity).
#pragma once
#include <stdint.h>

namespace cats {

constexpr int NC_FIELDS = 16;
constexpr int NK_FIELDS = 14;
constexpr int N_SCAL = 32;
constexpr int N_PARAMS = 34;
constexpr int N_STATS = 16;
constexpr int MAX_Z = 8;

// parameter slots: Cats._default_parameters key order (/root/reference/rmabm/environments/cats.py:69-104)
enum {
    P_z_c, P_z_k, P_z_e, P_xi, P_chi, P_q_adj, P_p_adj, P_mu, P_eta, P_Iprob, P_phi, P_theta,
    P_delta, P_alpha, P_k, P_div, P_barX, P_inventory_depreciation, P_b1, P_b2, P_b_k1, P_b_k2,
    P_interest_rate, P_subsidy, P_maastricht, P_target_deficit, P_tax_rate, P_wage_update_up,
    P_wage_update_down, P_u_target, P_wb, P_tax_rate_d, P_r_f, P_E_threshold_scale,
    // derived, filled on chip
    P_log1mtheta = 34, P_SLOTS = 36
};

// scalar slots of the blob (model.agg / bank / gov of the reference, cats.py:567-607)
enum {
    S_timestep, S_price, S_price_k, S_init_price, S_init_price_k, S_wb, S_Y_real, S_Y_nominal_tot,
    S_gdp_deflator, S_inflationRate, S_Un, S_consumption, S_totK, S_bankruptcy_rate,
    S_bank_E, S_bank_loans, S_profitsB, S_dividendsB, S_bank_interest_income, S_bank_bad_debt,
    S_gov_bonds, S_GB, S_deficitGDP, S_gov_TA, S_gov_G, S_gov_r_b, S_gov_subsidy_level,
    S_terminated, S_status, S_episode, S_reserved1, S_reserved2
};

// C-firm / K-firm float64 rows
enum { C_De, C_P, C_Y_prev, C_Yd, C_Q, C_A, C_K, C_investment, C_capital_value, C_wages,
       C_interests, C_deb, C_liquidity, C_barYK, C_interest_r, C_div_income };
// K-firm rows: the first NK_HOT are what the capital-goods market touches at random (resident in
// shared memory); the others are only ever touched by the lane that owns the firm (streamed)
enum { K_P, K_Yd, K_Q, K_De, K_Y_prev, K_A, K_wages, K_interests, K_deb, K_liquidity,
       K_interest_r, K_div_income, K_inv, K_Yavail };
constexpr int NK_HOT = 3;

struct Layout {
    int32_t W, F, N, H;
    int32_t rowF, rowN;  // bytes of one firm f64 row
    // persisted sections (byte offsets inside the blob)
    int32_t o_scal, o_kh, o_Leff, resident, o_Oc, o_c, o_kc, o_PA, o_perm, blob;
    // shared-memory working set: [0, resident) mirrors the blob prefix, then
    int32_t s_vac;                  // int32 x (F+N): vacancies (negative = surplus); Ld == Leff + vac
    int32_t s_cFtot, s_cKdem, s_cvalinv, s_kFtot, s_kYstock;  // f64 transients
    int32_t s_ic;                   // int32 x (F+N): firing buckets (count | cursor << 16) / capital-market buyers
    int32_t s_bad;                  // uint8 x (F+N): bankrupt flags
    int32_t s_U, U_bytes;           // union: goods-market seller records | firing pool | job seekers
    int32_t s_emp;                  // uint32 x ceil(W/32): employed bit per worker (job search -> goods market)
    int32_t s_par, s_misc, smem;
    // debug image (cats_workset_*): transients dumped after a stop_phase step
    int32_t d_Ld, d_vac, d_cFtot, d_cKdem, d_cYstock, d_cvalinv, d_kFtot, d_kYstock, workset;
};

__host__ __device__ constexpr int32_t up16(int64_t x) { return (int32_t)((x + 15) & ~(int64_t)15); }

constexpr Layout make_layout(int W, int F, int N)
{
    Layout L{};
    L.W = W; L.F = F; L.N = N; L.H = W + F + N;
    L.rowF = up16(8ll * F); L.rowN = up16(8ll * N);
    int64_t o = 0;
    L.o_scal = (int32_t)o; o += 8 * N_SCAL;
    L.o_kh = (int32_t)o; o += (int64_t)NK_HOT * L.rowN;
    L.o_Leff = (int32_t)o; o += up16(4ll * (F + N));
    L.resident = (int32_t)o;
    L.o_Oc = (int32_t)o; o += up16(4ll * W);
    o = (o + 127) & ~(int64_t)127;
    L.o_c = (int32_t)o; o += (int64_t)NC_FIELDS * L.rowF;
    L.o_kc = (int32_t)o; o += (int64_t)(NK_FIELDS - NK_HOT) * L.rowN;
    // households: one 16-byte record {wealth PA, permanent income perm} each -- the goods market
    // gathers both for a household visited at random, and a pair is one access and one sector
    L.o_PA = (int32_t)o; L.o_perm = L.o_PA + 8; o += 16ll * L.H;
    o = (o + 127) & ~(int64_t)127;
    L.blob = (int32_t)o;
    // shared memory
    o = L.resident;
    L.s_vac = (int32_t)o; o += up16(4ll * (F + N));
    L.s_cFtot = (int32_t)o; o += L.rowF;
    L.s_cKdem = (int32_t)o; o += L.rowF;
    L.s_cvalinv = (int32_t)o; o += L.rowF;
    L.s_kFtot = (int32_t)o; o += L.rowN;
    L.s_kYstock = (int32_t)o; o += L.rowN;
    L.s_ic = (int32_t)o; o += up16(4ll * (F + N));
    L.s_bad = L.s_ic;   // bankrupt flags reuse the firing / capital-market scratch (idle by then)
    // union region: goods-market seller table | firing pool of {worker, key} pairs | job seekers
    // (one uint16 per worker: W <= 65535 in this build)
    int64_t u = 28ll * F;   // seller table: (stock, 1/price) 16 B + price 8 B + class 4 B
    if (2ll * W > u) u = 2ll * W;
    L.s_U = (int32_t)o; L.U_bytes = up16(u); o += L.U_bytes;
    L.s_emp = (int32_t)o; o += up16(4ll * ((W + 31) / 32));
    L.s_par = (int32_t)o; o += 8 * P_SLOTS;
    L.s_misc = (int32_t)o; o += 128;   // status 8 B | mbarrier 8 B | OrderKeys 40 B | goods-market constants 9 x f64
    L.smem = (int32_t)o;
    // debug image
    o = 0;
    L.d_Ld = (int32_t)o; o += up16(4ll * (F + N));
    L.d_vac = (int32_t)o; o += up16(4ll * (F + N));
    L.d_cFtot = (int32_t)o; o += L.rowF;
    L.d_cKdem = (int32_t)o; o += L.rowF;
    L.d_cYstock = (int32_t)o; o += L.rowF;
    L.d_cvalinv = (int32_t)o; o += L.rowF;
    L.d_kFtot = (int32_t)o; o += L.rowN;
    L.d_kYstock = (int32_t)o; o += L.rowN;
    L.workset = (int32_t)o;
    return L;
}


// The reference's default economy (W=1000, F=100, N=20: cats.py:115-117, every BASELINE configuration
// but the 10x one; candidate-list lengths z_c=5, z_k=2, z_e=5) with the layout resolved at compile time: the same member names as `Layout`, as
// static constants, so that `L.F`, `L.o_c`, ... in the phase code fold into immediates (addresses
// become base + immediate, loops get constant bounds).
struct DefaultLayout {
    static constexpr int kW = 1000, kF = 100, kN = 20;
    static constexpr int kZc = 5, kZk = 2, kZe = 5;   // the default z_c, z_k, z_e (cats.py:70-72)
    static constexpr Layout k = make_layout(kW, kF, kN);
    static constexpr int32_t W = k.W;
    static constexpr int32_t F = k.F;
    static constexpr int32_t N = k.N;
    static constexpr int32_t H = k.H;
    static constexpr int32_t rowF = k.rowF;
    static constexpr int32_t rowN = k.rowN;
    static constexpr int32_t o_scal = k.o_scal;
    static constexpr int32_t o_kh = k.o_kh;
    static constexpr int32_t o_Leff = k.o_Leff;
    static constexpr int32_t resident = k.resident;
    static constexpr int32_t o_Oc = k.o_Oc;
    static constexpr int32_t o_c = k.o_c;
    static constexpr int32_t o_kc = k.o_kc;
    static constexpr int32_t o_PA = k.o_PA;
    static constexpr int32_t o_perm = k.o_perm;
    static constexpr int32_t blob = k.blob;
    static constexpr int32_t s_vac = k.s_vac;
    static constexpr int32_t s_cFtot = k.s_cFtot;
    static constexpr int32_t s_cKdem = k.s_cKdem;
    static constexpr int32_t s_cvalinv = k.s_cvalinv;
    static constexpr int32_t s_kFtot = k.s_kFtot;
    static constexpr int32_t s_kYstock = k.s_kYstock;
    static constexpr int32_t s_ic = k.s_ic;
    static constexpr int32_t s_bad = k.s_bad;
    static constexpr int32_t s_U = k.s_U;
    static constexpr int32_t U_bytes = k.U_bytes;
    static constexpr int32_t s_emp = k.s_emp;
    static constexpr int32_t s_par = k.s_par;
    static constexpr int32_t s_misc = k.s_misc;
    static constexpr int32_t smem = k.smem;
    static constexpr int32_t d_Ld = k.d_Ld;
    static constexpr int32_t d_vac = k.d_vac;
    static constexpr int32_t d_cFtot = k.d_cFtot;
    static constexpr int32_t d_cKdem = k.d_cKdem;
    static constexpr int32_t d_cYstock = k.d_cYstock;
    static constexpr int32_t d_cvalinv = k.d_cvalinv;
    static constexpr int32_t d_kFtot = k.d_kFtot;
    static constexpr int32_t d_kYstock = k.d_kYstock;
    static constexpr int32_t workset = k.workset;
};

struct TapeOff {
    uint32_t u_price_c, u_price_k, u_invest, fire_key, job_order, job_cand, cap_order, cap_cand,
        goods_order, goods_cand, bytes;
};

inline TapeOff make_tape(int W, int F, int N, int z_c, int z_k, int z_e)
{
    TapeOff t; uint64_t o = 0; uint64_t H = (uint64_t)W + F + N;
    t.u_price_c = (uint32_t)o; o += 8ull * F;
    t.u_price_k = (uint32_t)o; o += 8ull * N;
    t.u_invest = (uint32_t)o; o += 8ull * F;
    t.fire_key = (uint32_t)o; o += 4ull * W;
    t.job_order = (uint32_t)o; o += 4ull * W;
    t.job_cand = (uint32_t)o; o += 4ull * W * z_e;
    t.cap_order = (uint32_t)o; o += 4ull * F;
    t.cap_cand = (uint32_t)o; o += 4ull * F * z_k;
    t.goods_order = (uint32_t)o; o += 4ull * H;
    t.goods_cand = (uint32_t)o; o += 4ull * H * z_c;
    t.bytes = (uint32_t)((o + 15) & ~15ull);
    return t;
}

}  // namespace cats
