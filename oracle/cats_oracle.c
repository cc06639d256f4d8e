/*
 * cats_oracle.c -- CPU ORACLE for the batched CATS economy step.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this file.  The product path (r_mabm_b200/, include/) never links, imports or calls it.
 *
 * PARITY UNPINNED: the arithmetic of the reference's step lives in the Julia package ABCredit
 * (un-vendored, un-pinned: `Pkg.add("ABCredit")`, /root/reference/rmabm/environments/cats.py:158-162)
 * whose source is not on disk and cannot be installed here (no Julia, no network).  The reference's
 * own tests hold no golden vectors for the step (tests/test_env.py:157-178 assert lengths/types
 * only).  This file therefore restates the *published* CATS model (Assenza, Delli Gatti, Grazzini,
 * JEDC 2015 -- the paper /root/reference/README.md:4-5 names) in the call order and with the
 * field protocol visible at cats.py:355-410 and cats.py:434-626.  Every rule is written down in
 * docs/MODEL.md; this file and the CUDA kernels are two independent implementations of that spec.
 *
 * PINNED against the reference itself: the part of the path the reference implements in Python --
 * the call order of Cats.step, the action override, reward, termination, observations (ph_action,
 * ph_reward, oracle_obs, oracle_step below).  tests/golden/make_reference_driver.py runs the unmodified
 * /root/reference/rmabm code over a juliacall stand-in that forwards each ABCredit call to the ph_*
 * function here; tests/test_reference_driver.py holds oracle_step against what the reference returned.
 *
 * Plain C99, one economy at a time, deliberately simple.  Compile with
 *   gcc -O2 -std=c99 -ffp-contract=off -fPIC -shared (see oracle/Makefile)
 * -ffp-contract=off matters: the spec's arithmetic is "every * and + rounds separately".
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

/* ------------------------------------------------------------------------------------------ */
/* parameters: the 34 keys of Cats._default_parameters, in that order (cats.py:69-104)         */
enum {
    P_z_c, P_z_k, P_z_e, P_xi, P_chi, P_q_adj, P_p_adj, P_mu, P_eta, P_Iprob, P_phi, P_theta,
    P_delta, P_alpha, P_k, P_div, P_barX, P_inventory_depreciation, P_b1, P_b2, P_b_k1, P_b_k2,
    P_interest_rate, P_subsidy, P_maastricht, P_target_deficit, P_tax_rate, P_wage_update_up,
    P_wage_update_down, P_u_target, P_wb, P_tax_rate_d, P_r_f, P_E_threshold_scale, N_PARAMS
};

/* scalar slots (docs/MODEL.md section 2.1) */
enum {
    S_timestep, S_price, S_price_k, S_init_price, S_init_price_k, S_wb, S_Y_real, S_Y_nominal_tot,
    S_gdp_deflator, S_inflationRate, S_Un, S_consumption, S_totK, S_bankruptcy_rate,
    S_bank_E, S_bank_loans, S_profitsB, S_dividendsB, S_bank_interest_income, S_bank_bad_debt,
    S_gov_bonds, S_GB, S_deficitGDP, S_gov_TA, S_gov_G, S_gov_r_b, S_gov_subsidy_level,
    S_terminated, S_status, S_episode, S_reserved1, S_reserved2, N_SCAL
};

/* C-firm fp64 fields, K-firm fp64 fields (persisted) */
enum { C_De, C_P, C_Y_prev, C_Yd, C_Q, C_A, C_K, C_investment, C_capital_value, C_wages,
       C_interests, C_deb, C_liquidity, C_barYK, C_interest_r, C_div_income, NC_FIELDS };
enum { K_De, K_P, K_Y_prev, K_Yd, K_Q, K_A, K_wages, K_interests, K_deb, K_liquidity,
       K_interest_r, K_div_income, K_inv, K_Yavail, NK_FIELDS };

/* draw phases: Philox counter word 1 = (phase << 8) | block  (docs/MODEL.md section 3) */
enum { PH_PRICE_C = 1, PH_PRICE_K = 2, PH_INVEST = 3, PH_FIRE = 4, PH_JOB = 5, PH_CAP = 6,
       PH_GOODS = 7, PH_ORDER = 8 };
enum { MKT_JOB = 0, MKT_CAP = 1, MKT_GOODS = 2 };

/* step flags (same bits as include/cats_b200.h) */
#define FLAG_BURNIN      1u   /* ignore actions (cats.py:369) */
#define FLAG_LOG         2u   /* CatsLog action/observation transform (cats.py:655-704) */
#define FLAG_AVG_PRICE   4u   /* price_change == "avg_price" (cats.py:450-451) */
#define FLAG_REWARD_RMS  8u   /* reward_type == "rms" (cats.py:544-556) */
#define FLAG_STRICT     16u   /* plain arithmetic (docs/MODEL.md 4.6): b / P, sequential Yd, no reciprocals */

#define STATUS_DEPVAL_DIV0  1u  /* cats.py:516-523 ZeroDivisionError branch taken */
#define STATUS_PROFIT_NAN   2u  /* cats.py:533-535 NaN profits set to 0 */

#define MAX_Z 8

typedef struct {
    int W, F, N, H;
    double par[N_PARAMS];
    int z_c, z_k, z_e;
    double log1mtheta;
    uint64_t seed;
    uint32_t economy;
    double scal[N_SCAL];
    double *c[NC_FIELDS];
    double *k[NK_FIELDS];
    int32_t *c_Leff, *k_Leff;
    double *PA, *perm;
    int32_t *Oc;
    /* transients (live inside one period only) */
    int32_t *c_Ld, *c_vac, *k_Ld, *k_vac;
    double *c_Ftot, *c_Kdem, *c_Ystock, *c_valinv;
    double *k_Ftot, *k_Ystock;
    double *income, *budget;
    /* draw tape for the current period (NULL => Philox) */
    const uint8_t *tape;
    int bankruptcy_reward_set;
    double bankruptcy_reward;
    int stop_phase; /* debug: stop the period after this phase id (0 = run all) */
    int strict;     /* FLAG_STRICT of the step in progress (docs/MODEL.md 4.6) */
} Econ;

/* ------------------------------------------------------------------------------------------ */
/* deterministic math (docs/MODEL.md section 4): same operation sequence as the CUDA side       */

static double bits2d(uint64_t b) { double d; memcpy(&d, &b, 8); return d; }
static uint64_t d2bits(double d) { uint64_t b; memcpy(&b, &d, 8); return b; }

static const double LN2_HI = 6.93147180369123816490e-01;
static const double LN2_LO = 1.90821492927058770002e-10;
static const double INV_LN2 = 1.44269504088896338700e+00;

static double det_exp(double x)
{
    if (x != x) return x;
    if (x > 709.0) return INFINITY;
    if (x < -708.0) return 0.0;
    double kf = floor(x * INV_LN2 + 0.5);
    double r = (x - kf * LN2_HI) - kf * LN2_LO;
    /* Horner, degree 13, coefficients 1/n! */
    double p = 1.0 / 6227020800.0;
    p = p * r + 1.0 / 479001600.0;
    p = p * r + 1.0 / 39916800.0;
    p = p * r + 1.0 / 3628800.0;
    p = p * r + 1.0 / 362880.0;
    p = p * r + 1.0 / 40320.0;
    p = p * r + 1.0 / 5040.0;
    p = p * r + 1.0 / 720.0;
    p = p * r + 1.0 / 120.0;
    p = p * r + 1.0 / 24.0;
    p = p * r + 1.0 / 6.0;
    p = p * r + 0.5;
    p = p * r + 1.0;
    p = p * r + 1.0;
    int64_t ki = (int64_t)kf;
    double scale = bits2d((uint64_t)(ki + 1023) << 52);
    return p * scale;
}

static double det_log(double x)
{
    if (x != x) return x;
    if (x < 0.0) return NAN;
    if (x == 0.0) return -INFINITY;
    if (x == INFINITY) return x;
    uint64_t b = d2bits(x);
    int64_t e = (int64_t)((b >> 52) & 0x7ff);
    if (e == 0) { /* subnormal: scale up by 2^54 */
        x = x * 18014398509481984.0;
        b = d2bits(x);
        e = (int64_t)((b >> 52) & 0x7ff) - 54;
    }
    e -= 1023;
    double m = bits2d((b & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL);
    if (m > 1.4142135623730951) { m = m * 0.5; e += 1; }
    double s = (m - 1.0) / (m + 1.0);
    double z = s * s;
    double p = 1.0 / 23.0;
    p = p * z + 1.0 / 21.0;
    p = p * z + 1.0 / 19.0;
    p = p * z + 1.0 / 17.0;
    p = p * z + 1.0 / 15.0;
    p = p * z + 1.0 / 13.0;
    p = p * z + 1.0 / 11.0;
    p = p * z + 1.0 / 9.0;
    p = p * z + 1.0 / 7.0;
    p = p * z + 1.0 / 5.0;
    p = p * z + 1.0 / 3.0;
    p = p * z;
    double lm = 2.0 * (s + s * p);
    double ef = (double)e;
    return ef * LN2_HI + (lm + ef * LN2_LO);
}

/* det_sum: 32 strided partials (lane l adds x[l], x[l+32], ... in order), then xor-butterfly
 * 16,8,4,2,1.  Exactly what one CUDA warp does with __shfl_xor_sync.                           */
static double det_sum(const double *x, int n)
{
    double s[32], t[32];
    for (int l = 0; l < 32; ++l) {
        double a = 0.0;
        for (int i = l; i < n; i += 32) a = a + x[i];
        s[l] = a;
    }
    for (int off = 16; off >= 1; off >>= 1) {
        for (int l = 0; l < 32; ++l) t[l] = s[l] + s[l ^ off];
        memcpy(s, t, sizeof s);
    }
    return s[0];
}

/* ------------------------------------------------------------------------------------------ */
/* Philox-4x32-10 (Salmon et al. 2011), counter = (agent, episode<<12|phase<<8|block, period,   */
/* economy), key = (seed_lo, seed_hi)                                                          */
static void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                          uint32_t k1, uint32_t out[4])
{
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static void econ_philox(const Econ *e, uint32_t agent, uint32_t phase, uint32_t block,
                        uint32_t out[4])
{
    uint32_t period = (uint32_t)e->scal[S_timestep];
    uint32_t epi = (uint32_t)e->scal[S_episode] << 12;
    philox4x32_10(agent, (phase << 8) | block | epi, period, e->economy, (uint32_t)e->seed,
                  (uint32_t)(e->seed >> 32), out);
}

static double u01(uint32_t hi, uint32_t lo)
{
    uint64_t m = ((uint64_t)(hi >> 5) << 26) | (uint64_t)(lo >> 6);
    return (double)m * (1.0 / 9007199254740992.0);
}

/* z <= 8 distinct indices in [0,n): partial Fisher-Yates on a virtual identity array, fed by ONE
 * Philox block (4 words).  Draw k < 4 uses word k: j = k + mulhi32(w_k, n-k); draw k >= 4 re-uses
 * the low half of that product (multiply-shift chain): j = k + mulhi32(lo32(w_{k-4} * (n-k+4)), n-k). */
static void sample_distinct(const uint32_t *words, int n, int z, int32_t *out)
{
    int jpos[MAX_Z], jval[MAX_Z];
    for (int k = 0; k < z; ++k) {
        uint32_t w = words[k & 3];
        if (k >= 4) w = (uint32_t)((uint64_t)w * (uint64_t)(uint32_t)(n - (k - 4)));
        int j = k + (int)(((uint64_t)w * (uint64_t)(uint32_t)(n - k)) >> 32);
        int v = j, vk = k;
        for (int t = 0; t < k; ++t) {
            if (jpos[t] == j) v = jval[t];
            if (jpos[t] == k) vk = jval[t];
        }
        out[k] = v;
        jpos[k] = j; jval[k] = vk;
    }
}

/* keyed bijection on [0,n) (visiting order of a market), docs/MODEL.md 3.3: an alternating
 * 4-round Feistel network on [0,a) x [0,b), a = smallest integer with a*a >= n, b = ceil(n/a),
 * x = L*b + R, cycle-walked into [0,n) (a*b - n < a: almost never).  Round keys: the 4 words of
 * Philox block 0, phase ORDER, agent = market id. */
typedef struct { uint32_t k[4], a, b, m; int n; } Perm;

static void perm_dims(int n, uint32_t *pa, uint32_t *pb, uint32_t *pm)
{
    uint32_t a = 1;
    while ((uint64_t)a * a < (uint64_t)n) ++a;
    uint32_t b = ((uint32_t)n + a - 1) / a;
    *pa = a; *pb = b;
    *pm = b > 1 ? (uint32_t)(((uint64_t)1 << 32) / b) + 1u : 0u;   /* L = mulhi32(x, m) == x / b for x < 2^23 */
}

static uint32_t perm_hash(uint32_t v, uint32_t k)
{
    uint32_t x = v * 0x9E3779B1u + k;
    x ^= x >> 16;
    x *= 0x85EBCA6Bu; /* only the high bits are used (mulhi by the half-domain size) */
    return x;
}

static void perm_init(const Econ *e, int market, int n, Perm *p)
{
    uint32_t w[4];
    econ_philox(e, (uint32_t)market, PH_ORDER, 0, w);
    for (int r = 0; r < 4; ++r) p->k[r] = w[r];
    p->n = n;
    perm_dims(n, &p->a, &p->b, &p->m);
}

static int perm_at(const Perm *p, int pos)
{
    uint32_t x = (uint32_t)pos;
    const uint32_t a = p->a, b = p->b;
    do {
        uint32_t L = b > 1 ? (uint32_t)(((uint64_t)x * p->m) >> 32) : x;
        uint32_t R = x - L * b;
        L += (uint32_t)(((uint64_t)perm_hash(R, p->k[0]) * a) >> 32); if (L >= a) L -= a;
        R += (uint32_t)(((uint64_t)perm_hash(L, p->k[1]) * b) >> 32); if (R >= b) R -= b;
        L += (uint32_t)(((uint64_t)perm_hash(R, p->k[2]) * a) >> 32); if (L >= a) L -= a;
        R += (uint32_t)(((uint64_t)perm_hash(L, p->k[3]) * b) >> 32); if (R >= b) R -= b;
        x = L * b + R;
    } while (x >= (uint32_t)p->n);
    return (int)x;
}

/* ------------------------------------------------------------------------------------------ */
/* draw tape record for one economy-period (docs/MODEL.md section 3.4)                         */
typedef struct {
    size_t u_price_c, u_price_k, u_invest, fire_key, job_order, job_cand, cap_order, cap_cand,
        goods_order, goods_cand, bytes;
} TapeLayout;

static TapeLayout tape_layout(int W, int F, int N, int z_c, int z_k, int z_e)
{
    TapeLayout t; size_t o = 0; int H = W + F + N;
    t.u_price_c = o; o += 8u * (size_t)F;
    t.u_price_k = o; o += 8u * (size_t)N;
    t.u_invest = o; o += 8u * (size_t)F;
    t.fire_key = o; o += 4u * (size_t)W;
    t.job_order = o; o += 4u * (size_t)W;
    t.job_cand = o; o += 4u * (size_t)W * (size_t)z_e;
    t.cap_order = o; o += 4u * (size_t)F;
    t.cap_cand = o; o += 4u * (size_t)F * (size_t)z_k;
    t.goods_order = o; o += 4u * (size_t)H;
    t.goods_cand = o; o += 4u * (size_t)H * (size_t)z_c;
    t.bytes = (o + 15u) & ~(size_t)15u;
    return t;
}

size_t oracle_tape_bytes(int W, int F, int N, int z_c, int z_k, int z_e)
{
    return tape_layout(W, F, N, z_c, z_k, z_e).bytes;
}

static TapeLayout econ_tl(const Econ *e)
{
    return tape_layout(e->W, e->F, e->N, e->z_c, e->z_k, e->z_e);
}

/* --- the draw source: Philox, or the tape when one is installed ----------------------------- */
static double draw_uniform(const Econ *e, uint32_t phase, int agent)
{
    if (e->tape) {
        TapeLayout t = econ_tl(e);
        size_t off = phase == PH_PRICE_C ? t.u_price_c : phase == PH_PRICE_K ? t.u_price_k
                                                                             : t.u_invest;
        double u; memcpy(&u, e->tape + off + 8u * (size_t)agent, 8); return u;
    }
    uint32_t w[4]; econ_philox(e, (uint32_t)agent, phase, 0, w);
    return u01(w[0], w[1]);
}

static uint32_t draw_fire_key(const Econ *e, int worker)
{
    if (e->tape) {
        uint32_t k; memcpy(&k, e->tape + econ_tl(e).fire_key + 4u * (size_t)worker, 4); return k;
    }
    uint32_t w[4]; econ_philox(e, (uint32_t)worker, PH_FIRE, 0, w);
    return w[0];
}

static void draw_candidates(const Econ *e, uint32_t phase, int agent, int n, int z, int32_t *out)
{
    if (z > n) z = n;
    if (e->tape) {
        TapeLayout t = econ_tl(e);
        size_t off; int zz;
        if (phase == PH_JOB) { off = t.job_cand; zz = e->z_e; }
        else if (phase == PH_CAP) { off = t.cap_cand; zz = e->z_k; }
        else { off = t.goods_cand; zz = e->z_c; }
        memcpy(out, e->tape + off + 4u * ((size_t)agent * (size_t)zz), 4u * (size_t)z);
        return;
    }
    uint32_t w[4];
    econ_philox(e, (uint32_t)agent, phase, 0, w);
    sample_distinct(w, n, z, out);
}

typedef struct { const Econ *e; int market; Perm p; const int32_t *tape_order; } Order;

static void order_init(const Econ *e, int market, int n, Order *o)
{
    o->e = e; o->market = market; o->tape_order = NULL;
    if (e->tape) {
        TapeLayout t = econ_tl(e);
        size_t off = market == MKT_JOB ? t.job_order : market == MKT_CAP ? t.cap_order
                                                                         : t.goods_order;
        o->tape_order = (const int32_t *)(e->tape + off);
    } else {
        perm_init(e, market, n, &o->p);
    }
}

static int order_at(const Order *o, int pos)
{
    return o->tape_order ? o->tape_order[pos] : perm_at(&o->p, pos);
}

/* ------------------------------------------------------------------------------------------ */
/* a0: initialise_model (cats.py:224).  Deterministic: no draws.  docs/MODEL.md section 5.      */
static void *xcalloc(size_t n, size_t s) { void *p = calloc(n ? n : 1, s); if (!p) abort(); return p; }

/* a0 initialise_model (cats.py:224): the deterministic initial state, docs/MODEL.md section 5 */
static void init_state(Econ *e)
{
    int W = e->W, F = e->F, N = e->N;
    const double P0 = 3.0, PK0 = 3.0, K0 = 10.0, LIQ0 = 10.0, YC0 = 5.0, YK0 = 3.0, PA0 = 2.0;
    double *s = e->scal;
    s[S_timestep] = 0.0; s[S_price] = P0; s[S_price_k] = PK0; s[S_init_price] = P0;
    s[S_init_price_k] = PK0; s[S_wb] = e->par[P_wb]; s[S_bank_E] = 3000.0;
    s[S_gdp_deflator] = 1.0; s[S_Un] = 1.0; s[S_gov_r_b] = e->par[P_interest_rate];
    for (int f = 0; f < F; ++f) {
        e->c[C_De][f] = 1.0; e->c[C_P][f] = P0; e->c[C_Y_prev][f] = YC0; e->c[C_Yd][f] = YC0;
        e->c[C_K][f] = K0; e->c[C_capital_value][f] = K0 * PK0; e->c[C_liquidity][f] = LIQ0;
        e->c[C_A][f] = LIQ0 + K0 * PK0; e->c[C_barYK][f] = YC0 / e->par[P_k];
        e->c[C_interest_r][f] = e->par[P_r_f];
    }
    for (int i = 0; i < N; ++i) {
        e->k[K_De][i] = 1.0; e->k[K_P][i] = PK0; e->k[K_Y_prev][i] = YK0; e->k[K_Yd][i] = YK0;
        e->k[K_A][i] = LIQ0; e->k[K_liquidity][i] = LIQ0; e->k[K_interest_r][i] = e->par[P_r_f];
        e->k[K_Yavail][i] = YK0;
    }
    for (int h = 0; h < e->H; ++h) { e->PA[h] = PA0; e->perm[h] = 1.0 / P0; }
    for (int w = 0; w < W; ++w) e->Oc[w] = -1;
}

Econ *oracle_create(int W, int F, int N, const double *params34, uint64_t seed, uint32_t economy)
{
    Econ *e = (Econ *)xcalloc(1, sizeof(Econ));
    e->W = W; e->F = F; e->N = N; e->H = W + F + N;
    memcpy(e->par, params34, sizeof e->par);
    e->z_c = (int)e->par[P_z_c]; e->z_k = (int)e->par[P_z_k]; e->z_e = (int)e->par[P_z_e];
    e->log1mtheta = det_log(1.0 - e->par[P_theta]);
    e->seed = seed; e->economy = economy;
    for (int i = 0; i < NC_FIELDS; ++i) e->c[i] = (double *)xcalloc((size_t)F, 8);
    for (int i = 0; i < NK_FIELDS; ++i) e->k[i] = (double *)xcalloc((size_t)N, 8);
    e->c_Leff = (int32_t *)xcalloc((size_t)F, 4); e->k_Leff = (int32_t *)xcalloc((size_t)N, 4);
    e->PA = (double *)xcalloc((size_t)e->H, 8); e->perm = (double *)xcalloc((size_t)e->H, 8);
    e->Oc = (int32_t *)xcalloc((size_t)W, 4);
    e->c_Ld = (int32_t *)xcalloc((size_t)F, 4); e->c_vac = (int32_t *)xcalloc((size_t)F, 4);
    e->k_Ld = (int32_t *)xcalloc((size_t)N, 4); e->k_vac = (int32_t *)xcalloc((size_t)N, 4);
    e->c_Ftot = (double *)xcalloc((size_t)F, 8); e->c_Kdem = (double *)xcalloc((size_t)F, 8);
    e->c_Ystock = (double *)xcalloc((size_t)F, 8); e->c_valinv = (double *)xcalloc((size_t)F, 8);
    e->k_Ftot = (double *)xcalloc((size_t)N, 8); e->k_Ystock = (double *)xcalloc((size_t)N, 8);
    e->income = (double *)xcalloc((size_t)W, 8); e->budget = (double *)xcalloc((size_t)e->H, 8);

    init_state(e);
    e->bankruptcy_reward = -100.0;
    return e;
}

/* Per-economy reset (cats_reset_masked): a new model in place, episode counter + 1 (the salt of
 * the economy's Philox stream, see econ_philox). */
void oracle_reset(Econ *e)
{
    double episode = e->scal[S_episode] + 1.0;
    int F = e->F, N = e->N, W = e->W, H = e->H;
    memset(e->scal, 0, sizeof e->scal);
    for (int i = 0; i < NC_FIELDS; ++i) memset(e->c[i], 0, 8u * (size_t)F);
    for (int i = 0; i < NK_FIELDS; ++i) memset(e->k[i], 0, 8u * (size_t)N);
    memset(e->c_Leff, 0, 4u * (size_t)F); memset(e->k_Leff, 0, 4u * (size_t)N);
    memset(e->PA, 0, 8u * (size_t)H); memset(e->perm, 0, 8u * (size_t)H);
    memset(e->Oc, 0, 4u * (size_t)W);
    init_state(e);
    e->scal[S_episode] = episode;
}

void oracle_destroy(Econ *e)
{
    if (!e) return;
    for (int i = 0; i < NC_FIELDS; ++i) free(e->c[i]);
    for (int i = 0; i < NK_FIELDS; ++i) free(e->k[i]);
    free(e->c_Leff); free(e->k_Leff); free(e->PA); free(e->perm); free(e->Oc);
    free(e->c_Ld); free(e->c_vac); free(e->k_Ld); free(e->k_vac);
    free(e->c_Ftot); free(e->c_Kdem); free(e->c_Ystock); free(e->c_valinv);
    free(e->k_Ftot); free(e->k_Ystock); free(e->income); free(e->budget);
    free(e);
}

/* ------------------------------------------------------------------------------------------ */
/* a1 reset_gov_and_banking_variables! (cats.py:355)                                           */
void ph_reset_gov_bank(Econ *e)
{
    e->scal[S_gov_TA] = 0.0; e->scal[S_gov_G] = 0.0;
    e->scal[S_bank_interest_income] = 0.0; e->scal[S_bank_bad_debt] = 0.0;
}

/* a2 set_bond_interest_rate! (cats.py:356) */
void ph_bond_rate(Econ *e) { e->scal[S_gov_r_b] = e->par[P_interest_rate]; }

/* a3 / a4 firms_decide_price_quantity! (cats.py:363-364) */
static void price_quantity(Econ *e, double *De, double *P, const double *Yprev, const double *avail,
                           const double *Yd, int n, double avg, uint32_t phase)
{
    const double q_adj = e->par[P_q_adj], p_adj = e->par[P_p_adj], alpha = e->par[P_alpha];
    for (int f = 0; f < n; ++f) {
        double stock = avail[f] - Yd[f];
        double u = draw_uniform(e, phase, f);
        double de, p = P[f];
        if (stock <= 0.0) {
            if (p >= avg) de = Yprev[f] + (-stock) * q_adj;
            else { de = Yprev[f]; p = p * (1.0 + u * p_adj); }
        } else {
            if (p > avg) { de = Yprev[f]; p = p * (1.0 - u * p_adj); }
            else de = Yprev[f] - stock * q_adj;
        }
        if (de < alpha) de = alpha;
        De[f] = de; P[f] = p;
    }
}
void ph_price_quantity_c(Econ *e)
{
    price_quantity(e, e->c[C_De], e->c[C_P], e->c[C_Y_prev], e->c[C_Y_prev], e->c[C_Yd], e->F,
                   e->scal[S_price], PH_PRICE_C);
}
void ph_price_quantity_k(Econ *e)
{
    price_quantity(e, e->k[K_De], e->k[K_P], e->k[K_Y_prev], e->k[K_Yavail], e->k[K_Yd], e->N,
                   e->scal[S_price_k], PH_PRICE_K);
}

/* a5 firms_decide_investment! (cats.py:366) */
void ph_investment(Econ *e)
{
    const double Iprob = e->par[P_Iprob], barX = e->par[P_barX], eta = e->par[P_eta];
    for (int f = 0; f < e->F; ++f) {
        double u = draw_uniform(e, PH_INVEST, f);
        double kd = 0.0;
        if (u < Iprob) {
            double K = e->c[C_K][f];
            double Kdes = e->c[C_barYK][f] / barX;
            double depn = ((barX * K) * eta) / Iprob;
            kd = (Kdes - K) + depn;
            if (!(kd > 0.0)) kd = 0.0;
        }
        e->c_Kdem[f] = kd;
    }
}

/* a6 Cats._implement_action / CatsLog._implement_action (cats.py:434-464, 673-704).
 * prev_De/prev_P are the values BEFORE a3 (cats.py:359-360).                                  */
void ph_action(Econ *e, const double *actions, const uint8_t *mask, int n_agents, unsigned flags,
               const double *prev_De, const double *prev_P)
{
    if (flags & FLAG_BURNIN) return;
    for (int a = 0; a < n_agents; ++a) {
        if (!mask[a]) continue; /* "dummy" action */
        double a0 = actions[2 * a], a1 = actions[2 * a + 1];
        double base = (flags & FLAG_AVG_PRICE) ? e->scal[S_price] : prev_P[a];
        double de, p;
        if (flags & FLAG_LOG) {
            de = det_exp(det_log(prev_De[a] + 1e-8) + a0);
            p = det_exp(det_log(base + 1e-8) + a1);
        } else {
            de = prev_De[a] * a0;
            p = base * a1;
        }
        if (de < e->par[P_alpha]) de = e->par[P_alpha];
        e->c[C_De][a] = de; e->c[C_P][a] = p;
    }
}

static int32_t d2count(double x)
{
    if (!(x > 0.0)) return 0;
    if (x > 1.0e9) return 1000000000;
    return (int32_t)x;
}

/* a7 firms_decide_labour! (cats.py:373-374) */
void ph_labour_c(Econ *e)
{
    const double alpha = e->par[P_alpha], k = e->par[P_k];
    for (int f = 0; f < e->F; ++f) {
        double a = ceil(e->c[C_De][f] / alpha);
        double b = ceil((e->c[C_K][f] * k) / alpha);
        e->c_Ld[f] = d2count(a < b ? a : b);
    }
}
void ph_labour_k(Econ *e)
{
    const double alpha = e->par[P_alpha];
    for (int i = 0; i < e->N; ++i) {
        e->k_Ld[i] = d2count(ceil(e->k[K_De][i] / alpha));
    }
}

/* a8 firms_get_credit! (cats.py:376-377): bank prices default risk from leverage, rations by
 * phi * bank equity (equity as of the start of the period => no order dependence).            */
static void credit_one(Econ *e, double need, double b1, double b2, double *deb, double *A,
                       double *interest_r, double *Ftot, double liquidity, int32_t *Ld,
                       int32_t Leff, int32_t *vac)
{
    const double theta = e->par[P_theta], mu = e->par[P_mu], r_f = e->par[P_r_f];
    const double wb = e->scal[S_wb];
    double B = need > 0.0 ? need : 0.0;
    double credit = 0.0;
    if (B > 0.0) {
        double d0 = *deb;
        double lev = (d0 + B) / ((*A + B) + d0);
        double x = b1 + b2 * lev;
        double pr = 1.0 / (1.0 + det_exp(-x));
        double y = 1.0 + 1.0 / pr;
        double pw = det_exp(y * e->log1mtheta);
        double Xi = (1.0 - pw) / theta;
        double rate = mu * ((1.0 + r_f / theta) / Xi - theta);
        double maxL = (e->par[P_phi] * e->scal[S_bank_E] - pr * d0) / pr;
        if (!(maxL > 0.0)) maxL = 0.0;
        credit = B < maxL ? B : maxL;
        double d1 = d0 + credit;
        *deb = d1;
        if (d1 > 0.0) *interest_r = (d0 * (*interest_r) + rate * credit) / d1;
    }
    *Ftot = credit;
    double afford = floor((credit + liquidity) / wb);
    double ld = (double)(*Ld);
    if (afford < ld) ld = afford;
    int32_t L = d2count(ld);
    if (L < 1) L = 1;
    *Ld = L;
    *vac = L - Leff;
}
void ph_credit_c(Econ *e)
{
    const double wb = e->scal[S_wb], pk = e->scal[S_price_k];
    for (int f = 0; f < e->F; ++f) {
        double need = (wb * (double)e->c_Ld[f] + e->c_Kdem[f] * pk) - e->c[C_liquidity][f];
        credit_one(e, need, e->par[P_b1], e->par[P_b2], &e->c[C_deb][f], &e->c[C_A][f],
                   &e->c[C_interest_r][f], &e->c_Ftot[f], e->c[C_liquidity][f], &e->c_Ld[f],
                   e->c_Leff[f], &e->c_vac[f]);
    }
    e->scal[S_bank_loans] = e->scal[S_bank_loans] + det_sum(e->c_Ftot, e->F);
}
void ph_credit_k(Econ *e)
{
    const double wb = e->scal[S_wb];
    for (int i = 0; i < e->N; ++i) {
        double need = wb * (double)e->k_Ld[i] - e->k[K_liquidity][i];
        credit_one(e, need, e->par[P_b_k1], e->par[P_b_k2], &e->k[K_deb][i], &e->k[K_A][i],
                   &e->k[K_interest_r][i], &e->k_Ftot[i], e->k[K_liquidity][i], &e->k_Ld[i],
                   e->k_Leff[i], &e->k_vac[i]);
    }
    e->scal[S_bank_loans] = e->scal[S_bank_loans] + det_sum(e->k_Ftot, e->N);
}

/* a9 firms_fire_workers! (cats.py:379-380): a firm with s = Leff - Ld > 0 dismisses the s
 * employees with the smallest (fire_key, worker index).                                       */
static void fire(Econ *e, int firm_base, int n, int32_t *Ld, int32_t *Leff, int32_t *vac)
{
    for (int f = 0; f < n; ++f) {
        int s = -vac[f];
        if (s <= 0) continue;
        int id = firm_base + f;
        for (int r = 0; r < s; ++r) {
            int best = -1; uint32_t bk = 0;
            for (int w = 0; w < e->W; ++w) {
                if (e->Oc[w] != id) continue;
                uint32_t key = draw_fire_key(e, w);
                if (best < 0 || key < bk) { best = w; bk = key; }
            }
            if (best < 0) break;
            e->Oc[best] = -1;
        }
        Leff[f] = Ld[f];
        vac[f] = 0;
    }
}
void ph_fire_c(Econ *e) { fire(e, 0, e->F, e->c_Ld, e->c_Leff, e->c_vac); }
void ph_fire_k(Econ *e) { fire(e, e->F, e->N, e->k_Ld, e->k_Leff, e->k_vac); }

/* a10 workers_search_job! (cats.py:382) */
void ph_job_search(Econ *e)
{
    Order o; order_init(e, MKT_JOB, e->W, &o);
    int nf = e->F + e->N;
    int z = e->z_e < nf ? e->z_e : nf;
    for (int pos = 0; pos < e->W; ++pos) {
        int w = order_at(&o, pos);
        if (e->Oc[w] >= 0) continue;
        int32_t cand[MAX_Z];
        draw_candidates(e, PH_JOB, w, nf, z, cand);
        for (int k = 0; k < z; ++k) {
            int f = cand[k];
            if (f < e->F) {
                if (e->c_vac[f] > 0) { e->Oc[w] = f; e->c_Leff[f] += 1; e->c_vac[f] -= 1; break; }
            } else {
                int i = f - e->F;
                if (e->k_vac[i] > 0) { e->Oc[w] = f; e->k_Leff[i] += 1; e->k_vac[i] -= 1; break; }
            }
        }
    }
}

/* a11 firms_produce! (cats.py:384-385) */
void ph_produce_c(Econ *e)
{
    const double alpha = e->par[P_alpha], k = e->par[P_k], eta = e->par[P_eta];
    const double theta = e->par[P_theta], wb = e->scal[S_wb], pk = e->scal[S_price_k];
    for (int f = 0; f < e->F; ++f) {
        double cap_l = (double)e->c_Leff[f] * alpha, cap_k = e->c[C_K][f] * k;
        double Yp = cap_l < cap_k ? cap_l : cap_k;
        double De = e->c[C_De][f];
        double Y = De < Yp ? De : Yp;
        e->c[C_Y_prev][f] = Y; e->c_Ystock[f] = Y;
        double interests = e->c[C_interest_r][f] * e->c[C_deb][f];
        double wages = wb * (double)e->c_Leff[f];
        e->c[C_interests][f] = interests; e->c[C_wages][f] = wages;
        double Pl = 0.0;
        if (Y > 0.0)
            Pl = (1.01 * (((wages + interests) + ((Y / k) * eta) * pk) + theta * e->c[C_deb][f])) / Y;
        if (e->c[C_P][f] < Pl) e->c[C_P][f] = Pl;
        e->c[C_Q][f] = 0.0; e->c[C_Yd][f] = 0.0; e->c[C_investment][f] = 0.0; e->c_valinv[f] = 0.0;
    }
}
void ph_produce_k(Econ *e)
{
    const double alpha = e->par[P_alpha], theta = e->par[P_theta], wb = e->scal[S_wb];
    for (int i = 0; i < e->N; ++i) {
        double want = e->k[K_De][i];
        double cap_l = (double)e->k_Leff[i] * alpha;
        double Y = want < cap_l ? want : cap_l;
        e->k[K_Y_prev][i] = Y;
        double avail = Y + e->k[K_inv][i];
        e->k_Ystock[i] = avail; e->k[K_Yavail][i] = avail;
        double interests = e->k[K_interest_r][i] * e->k[K_deb][i];
        double wages = wb * (double)e->k_Leff[i];
        e->k[K_interests][i] = interests; e->k[K_wages][i] = wages;
        double Pl = 0.0;
        if (Y > 0.0) Pl = (1.01 * ((wages + interests) + theta * e->k[K_deb][i])) / Y;
        if (e->k[K_P][i] < Pl) e->k[K_P][i] = Pl;
        e->k[K_Q][i] = 0.0; e->k[K_Yd][i] = 0.0;
    }
}

/* stable insertion sort of candidate list by ascending price */
static void sort_by_price(int32_t *cand, int z, const double *P)
{
    for (int a = 1; a < z; ++a) {
        int32_t c = cand[a]; double pc = P[c]; int b = a - 1;
        while (b >= 0 && pc < P[cand[b]]) { cand[b + 1] = cand[b]; --b; }
        cand[b + 1] = c;
    }
}

/* a12 capital_goods_market! (cats.py:387) */
void ph_capital_market(Econ *e)
{
    Order o; order_init(e, MKT_CAP, e->F, &o);
    int z = e->z_k < e->N ? e->z_k : e->N;
    double *Pk = e->k[K_P];
    double *rp = (double *)xcalloc((size_t)e->N, 8); /* 1 / P, once per seller (MODEL.md 4.5) */
    for (int i = 0; i < e->N; ++i) rp[i] = 1.0 / Pk[i];
    for (int pos = 0; pos < e->F; ++pos) {
        int f = order_at(&o, pos);
        double kd = e->c_Kdem[f];
        if (!(kd > 0.0)) continue;
        double cb = (e->c_Ftot[f] + e->c[C_liquidity][f]) - e->c[C_wages][f];
        int32_t cand[MAX_Z];
        draw_candidates(e, PH_CAP, f, e->N, z, cand);
        sort_by_price(cand, z, Pk);
        for (int k = 0; k < z; ++k) {
            if (!(cb > 0.0 && kd > 0.0)) break;
            int best = cand[k];
            double pk = Pk[best];
            double afford = e->strict ? cb / pk : cb * rp[best];
            double want = afford < kd ? afford : kd;
            e->k[K_Yd][best] = e->k[K_Yd][best] + want;
            double ys = e->k_Ystock[best];
            if (ys > 0.0) {
                double full = kd * pk;
                double spend = cb < full ? cb : full;
                double q = e->strict ? spend / pk : spend * rp[best];
                if (ys > q) {
                    e->k_Ystock[best] = ys - q;
                    kd = kd - q; cb = cb - spend;
                    e->c[C_liquidity][f] = e->c[C_liquidity][f] - spend;
                    e->c[C_investment][f] = e->c[C_investment][f] + q;
                    e->c_valinv[f] = e->c_valinv[f] + spend;
                } else {
                    double cost = ys * pk;
                    kd = kd - ys; cb = cb - cost;
                    e->c[C_liquidity][f] = e->c[C_liquidity][f] - cost;
                    e->c[C_investment][f] = e->c[C_investment][f] + ys;
                    e->c_valinv[f] = e->c_valinv[f] + cost;
                    e->k_Ystock[best] = 0.0;
                }
            }
        }
        e->c_Kdem[f] = kd;
    }
    free(rp);
}

/* a13 gov_adjusts_subsidy! (cats.py:389) */
void ph_gov_subsidy(Econ *e)
{
    double sub = e->par[P_subsidy] * e->scal[S_wb];
    if (e->par[P_maastricht] > 0.0 && e->scal[S_deficitGDP] > e->par[P_target_deficit]) {
        double scale = 1.0 - e->par[P_maastricht] * (e->scal[S_deficitGDP] - e->par[P_target_deficit]);
        if (!(scale > 0.0)) scale = 0.0;
        sub = sub * scale;
    }
    e->scal[S_gov_subsidy_level] = sub;
}

/* a14 workers_get_paid! (cats.py:391) */
void ph_get_paid(Econ *e)
{
    const double wb = e->scal[S_wb], tax = e->par[P_tax_rate], sub = e->scal[S_gov_subsidy_level];
    double net = wb * (1.0 - tax);
    int n_emp = 0;
    for (int w = 0; w < e->W; ++w) {
        double inc;
        if (e->Oc[w] >= 0) { inc = net; ++n_emp; } else inc = sub;
        e->income[w] = inc;
        e->PA[w] = e->PA[w] + inc;
    }
    e->scal[S_gov_TA] = e->scal[S_gov_TA] + (wb * tax) * (double)n_emp;
    e->scal[S_gov_G] = e->scal[S_gov_G] + sub * (double)(e->W - n_emp);
}

/* a15 households_find_cons_budget! (cats.py:393); households = workers, C-owners, K-owners
 * (cats.py:235)                                                                               */
void ph_cons_budget(Econ *e)
{
    const double xi = e->par[P_xi], chi = e->par[P_chi], price = e->scal[S_price];
    const double rpavg = 1.0 / price; /* real values are nominal * (1 / price): MODEL.md 4.5 */
    for (int h = 0; h < e->H; ++h) {
        double inc = h < e->W ? e->income[h]
                   : h < e->W + e->F ? e->c[C_div_income][h - e->W]
                                     : e->k[K_div_income][h - e->W - e->F];
        double ir = e->strict ? inc / price : inc * rpavg;
        double pm = e->perm[h] * xi + (1.0 - xi) * ir;
        e->perm[h] = pm;
        double pa = e->PA[h];
        double target = pm + (e->strict ? (chi * pa) / price : (chi * pa) * rpavg);
        double b = target * price;
        if (b > pa) b = pa;
        if (!(b > 0.0)) b = 0.0;
        e->PA[h] = pa - b;
        e->budget[h] = b;
    }
}

/* Demand expressed at a consumption-goods seller (docs/MODEL.md 4.4).  A visit by a household with
 * budget b registers the demand b / P_seller.  The sum over the visits is order-independent by
 * construction: each visit adds round(b * (1 / Pavg) * 2^40) to a wrap-around uint64 accumulator
 * (Pavg = last period's average price: budgets in real units), and when the market closes
 * Yd = acc * 2^-40 * Pavg / P_seller.  Every other quantity of the market (stock, sales, budgets)
 * follows the visiting order exactly. */
#define FX_ONE 1099511627776.0 /* 2^40 */
static uint64_t fx_from(double x)
{
    if (!(x < 4194304.0)) return (uint64_t)1 << 62; /* saturate (also NaN, +inf) */
    return (uint64_t)(int64_t)llrint(x * FX_ONE);   /* round to nearest even */
}
static double fx_to(uint64_t a) { return (double)(int64_t)a * (1.0 / FX_ONE); }

/* a16 consumption_goods_market! (cats.py:394) */
void ph_goods_market(Econ *e)
{
    Order o; order_init(e, MKT_GOODS, e->H, &o);
    int z = e->z_c < e->F ? e->z_c : e->F;
    const double *P = e->c[C_P];
    const double pavg = e->scal[S_price], rpavg = 1.0 / pavg;
    uint64_t *acc = (uint64_t *)xcalloc((size_t)e->F, 8);
    double *rp = (double *)xcalloc((size_t)e->F, 8); /* 1 / P, once per seller (MODEL.md 4.5) */
    for (int f = 0; f < e->F; ++f) rp[f] = 1.0 / P[f];
    for (int pos = 0; pos < e->H; ++pos) {
        int h = order_at(&o, pos);
        double b = e->budget[h];
        if (!(b > 0.0)) continue;
        int32_t cand[MAX_Z];
        draw_candidates(e, PH_GOODS, h, e->F, z, cand);
        sort_by_price(cand, z, P);
        for (int k = 0; k < z; ++k) {
            if (!(b > 0.0)) break;
            int best = cand[k];
            double p = P[best];
            double dem = 0.0;
            if (e->strict) { dem = b / p; e->c[C_Yd][best] = e->c[C_Yd][best] + dem; }   /* in visiting order */
            else acc[best] += fx_from(b * rpavg);
            double ys = e->c_Ystock[best];
            if (ys > 0.0) {
                double d = e->strict ? dem : b * rp[best];
                if (ys > d) {
                    e->c_Ystock[best] = ys - d;
                    b = 0.0;
                } else {
                    b = b - ys * p;
                    e->c_Ystock[best] = 0.0;
                }
            }
        }
        e->PA[h] = e->PA[h] + b; /* involuntary saving */
        e->budget[h] = b;
    }
    for (int f = 0; f < e->F; ++f) {
        if (!e->strict) e->c[C_Yd][f] = (fx_to(acc[f]) * pavg) * rp[f];
        e->c[C_Q][f] = e->c[C_Y_prev][f] - e->c_Ystock[f]; /* sold = produced - left over */
    }
    free(acc); free(rp);
}

/* a17 firms_accounting! (cats.py:397-398) */
void ph_accounting_c(Econ *e)
{
    const double delta = e->par[P_delta], eta = e->par[P_eta], k = e->par[P_k];
    const double theta = e->par[P_theta], divs = e->par[P_div], taxd = e->par[P_tax_rate_d];
    int F = e->F;
    double *inst = (double *)xcalloc((size_t)F, 8), *dtax = (double *)xcalloc((size_t)F, 8);
    for (int f = 0; f < F; ++f) {
        double Yp = e->c[C_Y_prev][f];
        e->c[C_barYK][f] = delta * e->c[C_barYK][f] + (1.0 - delta) * (Yp / k);
        double dep = (eta * Yp) / k;
        double Kold = e->c[C_K][f];
        e->c[C_K][f] = (Kold - dep) + e->c[C_investment][f];
        double dv = 0.0;
        if (Kold != 0.0) dv = (dep * e->c[C_capital_value][f]) / Kold;
        e->c[C_capital_value][f] = (e->c[C_capital_value][f] - dv) + e->c_valinv[f];
        double RIC = e->c[C_P][f] * e->c[C_Q][f];
        double wages = e->c[C_wages][f], interests = e->c[C_interests][f];
        double in = theta * e->c[C_deb][f];
        inst[f] = in;
        double liq = ((((e->c[C_liquidity][f] + RIC) + e->c_Ftot[f]) - wages) - interests) - in;
        e->c[C_deb][f] = e->c[C_deb][f] - in;
        double pi = ((RIC - wages) - interests) - dv;
        double A = e->c[C_A][f] + pi;
        double net = 0.0; dtax[f] = 0.0;
        if (pi > 0.0) {
            double dvd = divs * pi;
            liq = liq - dvd; A = A - dvd;
            net = dvd * (1.0 - taxd);
            dtax[f] = dvd - net;
            e->PA[e->W + f] = e->PA[e->W + f] + net;
        }
        e->c[C_div_income][f] = net;
        e->c[C_liquidity][f] = liq; e->c[C_A][f] = A;
    }
    e->scal[S_bank_loans] = e->scal[S_bank_loans] - det_sum(inst, F);
    e->scal[S_bank_interest_income] = e->scal[S_bank_interest_income] + det_sum(e->c[C_interests], F);
    e->scal[S_gov_TA] = e->scal[S_gov_TA] + det_sum(dtax, F);
    free(inst); free(dtax);
}
void ph_accounting_k(Econ *e)
{
    const double theta = e->par[P_theta], divs = e->par[P_div], taxd = e->par[P_tax_rate_d];
    int N = e->N;
    double *inst = (double *)xcalloc((size_t)N, 8), *dtax = (double *)xcalloc((size_t)N, 8);
    for (int i = 0; i < N; ++i) {
        e->k[K_Q][i] = e->k[K_Yavail][i] - e->k_Ystock[i]; /* sold = available - left over */
        double RIC = e->k[K_P][i] * e->k[K_Q][i];
        e->k[K_inv][i] = (1.0 - e->par[P_inventory_depreciation]) * e->k_Ystock[i];
        double wages = e->k[K_wages][i], interests = e->k[K_interests][i];
        double in = theta * e->k[K_deb][i];
        inst[i] = in;
        double liq = ((((e->k[K_liquidity][i] + RIC) + e->k_Ftot[i]) - wages) - interests) - in;
        e->k[K_deb][i] = e->k[K_deb][i] - in;
        double pi = (RIC - wages) - interests;
        double A = e->k[K_A][i] + pi;
        double net = 0.0; dtax[i] = 0.0;
        if (pi > 0.0) {
            double dvd = divs * pi;
            liq = liq - dvd; A = A - dvd;
            net = dvd * (1.0 - taxd);
            dtax[i] = dvd - net;
            e->PA[e->W + e->F + i] = e->PA[e->W + e->F + i] + net;
        }
        e->k[K_div_income][i] = net;
        e->k[K_liquidity][i] = liq; e->k[K_A][i] = A;
    }
    e->scal[S_bank_loans] = e->scal[S_bank_loans] - det_sum(inst, N);
    e->scal[S_bank_interest_income] = e->scal[S_bank_interest_income] + det_sum(e->k[K_interests], N);
    e->scal[S_gov_TA] = e->scal[S_gov_TA] + det_sum(dtax, N);
    free(inst); free(dtax);
}

/* a18 Cats._get_reward (cats.py:493-556), evaluated between accounting and tracking */
/* sum(p[f] * q[f] for f in [lo, hi)) as CPython >= 3.12 adds floats: Neumaier's compensated summation
 * (Objects/bltinmodule.c, builtin_sum_impl); CPython <= 3.11 added left to right, a few ulp away. */
static double py_sum(const double *p, const double *q, int lo, int hi)
{
    double acc = 0.0, comp = 0.0;
    for (int f = lo; f < hi; ++f) {
        double x = p[f] * q[f], t = acc + x;
        if (fabs(acc) >= fabs(x)) comp = comp + ((acc - t) + x); else comp = comp + ((x - t) + acc);
        acc = t;
    }
    if (comp != 0.0 && isfinite(comp)) acc = acc + comp;
    return acc;
}

void ph_reward(Econ *e, int n_agents, unsigned flags, double *reward)
{
    if (flags & FLAG_REWARD_RMS) {
        /* cats.py:548-549: Python's built-in sum() over the RL firms' revenues, then over the others';
         * the two sums are added plainly.  py_sum below is CPython >= 3.12's float summation. */
        double total = py_sum(e->c[C_P], e->c[C_Q], 0, n_agents) + py_sum(e->c[C_P], e->c[C_Q], n_agents, e->F);
        for (int a = 0; a < n_agents; ++a) reward[a] = (e->c[C_P][a] * e->c[C_Q][a]) / total;
        return;
    }
    const double eta = e->par[P_eta], k = e->par[P_k];
    for (int a = 0; a < n_agents; ++a) {
        double dep = (eta * e->c[C_Y_prev][a]) / k;
        double den = (e->c[C_K][a] + dep) - e->c[C_investment][a];
        double dv;
        if (den == 0.0) { dv = 0.0; e->scal[S_status] = (double)((unsigned)e->scal[S_status] | STATUS_DEPVAL_DIV0); }
        else dv = (dep * e->c[C_capital_value][a]) / den;
        double RIC = e->c[C_P][a] * e->c[C_Q][a];
        double profits = ((RIC - e->c[C_wages][a]) - e->c[C_interests][a]) - dv;
        profits = (profits * e->scal[S_init_price]) / e->scal[S_price];
        if (profits != profits) { profits = 0.0; e->scal[S_status] = (double)((unsigned)e->scal[S_status] | STATUS_PROFIT_NAN); }
        if (e->c[C_A][a] <= 0.0) profits = e->bankruptcy_reward;
        reward[a] = profits;
    }
}

/* a19 update_tracking_variables! (cats.py:401) */
void ph_tracking(Econ *e)
{
    int F = e->F, N = e->N;
    double *tmp = (double *)xcalloc((size_t)(F > N ? F : N), 8);
    double sQ = det_sum(e->c[C_Q], F);
    for (int f = 0; f < F; ++f) tmp[f] = e->c[C_P][f] * e->c[C_Q][f];
    double sPQ = det_sum(tmp, F);
    double price_prev = e->scal[S_price];
    double price = sQ > 0.0 ? sPQ / sQ : det_sum(e->c[C_P], F) / (double)F;
    double sQk = det_sum(e->k[K_Q], N);
    for (int i = 0; i < N; ++i) tmp[i] = e->k[K_P][i] * e->k[K_Q][i];
    double sPQk = det_sum(tmp, N);
    double price_k = sQk > 0.0 ? sPQk / sQk : e->scal[S_price_k];
    double sY = det_sum(e->c[C_Y_prev], F), sYk = det_sum(e->k[K_Y_prev], N);
    double Yreal = sY * e->scal[S_init_price] + sYk * e->scal[S_init_price_k];
    double Ynom = sY * price + sYk * price_k;
    int L = 0;
    for (int f = 0; f < F; ++f) L += e->c_Leff[f];
    for (int i = 0; i < N; ++i) L += e->k_Leff[i];
    e->scal[S_price] = price; e->scal[S_price_k] = price_k;
    e->scal[S_Y_real] = Yreal; e->scal[S_Y_nominal_tot] = Ynom;
    e->scal[S_gdp_deflator] = Yreal > 0.0 ? Ynom / Yreal : 0.0;
    e->scal[S_inflationRate] = price / price_prev - 1.0;
    e->scal[S_Un] = 1.0 - (double)L / (double)e->W;
    e->scal[S_consumption] = sPQ;
    e->scal[S_totK] = det_sum(e->c[C_K], F);
    free(tmp);
}

/* a20 firms_go_bankrupt! (cats.py:404) + a21 Cats._get_terminated (cats.py:483-486) */
void ph_bankrupt(Econ *e)
{
    int F = e->F, N = e->N, W = e->W;
    const double thr = e->par[P_E_threshold_scale] * e->scal[S_price];
    int M = F > N ? F : N;
    uint8_t *bad = (uint8_t *)xcalloc((size_t)M, 1);
    double *t0 = (double *)xcalloc((size_t)M, 8), *t1 = (double *)xcalloc((size_t)M, 8);
    double *t2 = (double *)xcalloc((size_t)M, 8), *t3 = (double *)xcalloc((size_t)M, 8);
    double *t4 = (double *)xcalloc((size_t)M, 8), *t5 = (double *)xcalloc((size_t)M, 8);
    int nbad_total = 0;

    /* ---- C firms ---- */
    int ns = 0;
    for (int f = 0; f < F; ++f) {
        double A = e->c[C_A][f], liq = e->c[C_liquidity][f], deb = e->c[C_deb][f];
        int b = (A <= thr) || (liq < 0.0);
        bad[f] = (uint8_t)b;
        if (!b) ++ns;
        t0[f] = b ? 0.0 : e->c[C_K][f];
        t1[f] = b ? 0.0 : liq;
        t2[f] = b ? 0.0 : e->c[C_P][f];
        t3[f] = b ? 0.0 : e->c[C_Y_prev][f];
        double expo = deb + (liq < 0.0 ? -liq : 0.0);
        double loss = A < 0.0 ? -A : 0.0;
        if (loss > expo) loss = expo;
        t4[f] = b ? loss : 0.0;
        t5[f] = b ? deb : 0.0;
    }
    double mK = 10.0, mL = 0.0, mP = e->scal[S_price], mY = e->par[P_alpha];
    if (ns > 0) {
        mK = det_sum(t0, F) / (double)ns; mL = det_sum(t1, F) / (double)ns;
        mP = det_sum(t2, F) / (double)ns; mY = det_sum(t3, F) / (double)ns;
    }
    e->scal[S_bank_bad_debt] = e->scal[S_bank_bad_debt] + det_sum(t4, F);
    e->scal[S_bank_loans] = e->scal[S_bank_loans] - det_sum(t5, F);
    for (int f = 0; f < F; ++f) {
        if (!bad[f]) continue;
        ++nbad_total;
        int ow = W + f;
        double inj = mL > 0.0 ? mL : 0.0;
        if (inj > e->PA[ow]) inj = e->PA[ow];
        if (!(inj > 0.0)) inj = 0.0;
        e->PA[ow] = e->PA[ow] - inj;
        e->c[C_liquidity][f] = inj; e->c[C_K][f] = mK;
        e->c[C_capital_value][f] = mK * e->scal[S_price_k];
        e->c[C_A][f] = inj + e->c[C_capital_value][f];
        e->c[C_deb][f] = 0.0; e->c[C_P][f] = mP; e->c[C_Y_prev][f] = mY; e->c[C_Yd][f] = mY;
        e->c[C_De][f] = mY; e->c[C_barYK][f] = mY / e->par[P_k];
        e->c[C_interest_r][f] = e->par[P_r_f];
        e->c[C_Q][f] = 0.0; e->c[C_investment][f] = 0.0; e->c[C_wages][f] = 0.0;
        e->c[C_interests][f] = 0.0; e->c[C_div_income][f] = 0.0;
        e->c_Leff[f] = 0;
    }
    for (int w = 0; w < W; ++w) { int o = e->Oc[w]; if (o >= 0 && o < F && bad[o]) e->Oc[w] = -1; }

    /* ---- K firms ---- */
    ns = 0;
    for (int i = 0; i < N; ++i) {
        double A = e->k[K_A][i], liq = e->k[K_liquidity][i], deb = e->k[K_deb][i];
        int b = (A <= thr) || (liq < 0.0);
        bad[i] = (uint8_t)b;
        if (!b) ++ns;
        t1[i] = b ? 0.0 : liq;
        t2[i] = b ? 0.0 : e->k[K_P][i];
        t3[i] = b ? 0.0 : e->k[K_Y_prev][i];
        double expo = deb + (liq < 0.0 ? -liq : 0.0);
        double loss = A < 0.0 ? -A : 0.0;
        if (loss > expo) loss = expo;
        t4[i] = b ? loss : 0.0;
        t5[i] = b ? deb : 0.0;
    }
    mL = 0.0; mP = e->scal[S_price_k]; mY = e->par[P_alpha];
    if (ns > 0) {
        mL = det_sum(t1, N) / (double)ns; mP = det_sum(t2, N) / (double)ns;
        mY = det_sum(t3, N) / (double)ns;
    }
    e->scal[S_bank_bad_debt] = e->scal[S_bank_bad_debt] + det_sum(t4, N);
    e->scal[S_bank_loans] = e->scal[S_bank_loans] - det_sum(t5, N);
    for (int i = 0; i < N; ++i) {
        if (!bad[i]) continue;
        ++nbad_total;
        int ow = W + F + i;
        double inj = mL > 0.0 ? mL : 0.0;
        if (inj > e->PA[ow]) inj = e->PA[ow];
        if (!(inj > 0.0)) inj = 0.0;
        e->PA[ow] = e->PA[ow] - inj;
        e->k[K_liquidity][i] = inj; e->k[K_A][i] = inj; e->k[K_deb][i] = 0.0;
        e->k[K_P][i] = mP; e->k[K_Y_prev][i] = mY; e->k[K_Yd][i] = mY; e->k[K_Yavail][i] = mY;
        e->k[K_De][i] = mY; e->k[K_inv][i] = 0.0; e->k[K_interest_r][i] = e->par[P_r_f];
        e->k[K_Q][i] = 0.0; e->k[K_wages][i] = 0.0; e->k[K_interests][i] = 0.0;
        e->k[K_div_income][i] = 0.0;
        e->k_Leff[i] = 0;
    }
    for (int w = 0; w < W; ++w) { int o = e->Oc[w]; if (o >= F && bad[o - F]) e->Oc[w] = -1; }

    e->scal[S_bankruptcy_rate] = (double)nbad_total / (double)(F + N);
    e->scal[S_terminated] = det_sum(e->c[C_A], F) == 0.0 ? 1.0 : 0.0;
    free(bad); free(t0); free(t1); free(t2); free(t3); free(t4); free(t5);
}

/* a22 bank_accounting! (cats.py:407) */
void ph_bank_accounting(Econ *e)
{
    double *s = e->scal;
    double gross = (s[S_bank_interest_income] + s[S_gov_r_b] * s[S_gov_bonds]) - s[S_bank_bad_debt];
    double dB = 0.0;
    if (gross > 0.0 && s[S_bank_E] > 0.0) dB = e->par[P_div] * gross;
    s[S_bank_E] = s[S_bank_E] + (gross - dB);
    s[S_dividendsB] = dB; s[S_profitsB] = gross;
    if (dB > 0.0) {
        double share = dB / (double)(e->F + e->N);
        for (int h = e->W; h < e->H; ++h) e->PA[h] = e->PA[h] + share;
    }
}

/* a23 gov_accounting! (cats.py:408) */
void ph_gov_accounting(Econ *e)
{
    double *s = e->scal;
    s[S_gov_G] = s[S_gov_G] + s[S_gov_r_b] * s[S_gov_bonds];
    double GB = s[S_gov_TA] - s[S_gov_G];
    s[S_GB] = GB;
    s[S_gov_bonds] = s[S_gov_bonds] - GB;
    s[S_deficitGDP] = s[S_Y_nominal_tot] > 0.0 ? (-GB) / s[S_Y_nominal_tot] : 0.0;
}

/* a24 adjust_wages! (cats.py:410) */
void ph_adjust_wages(Econ *e)
{
    double *s = e->scal;
    double Un = s[S_Un], ut = e->par[P_u_target];
    if (Un < ut) s[S_wb] = s[S_wb] * (1.0 + e->par[P_wage_update_up] * (ut - Un));
    else s[S_wb] = s[S_wb] * (1.0 - e->par[P_wage_update_down] * (Un - ut));
}

/* a26 Cats._get_obs / CatsLog._get_obs (cats.py:466-481, 655-671) */
void oracle_obs(const Econ *e, int n_agents, unsigned flags, double *obs)
{
    for (int a = 0; a < n_agents; ++a) {
        double yp = e->c[C_Y_prev][a], yd = e->c[C_Yd][a], p = e->c[C_P][a], ap = e->scal[S_price];
        if (flags & FLAG_LOG) {
            obs[2 * a] = det_log(yp + 1e-8) - det_log(yd + 1e-8);
            obs[2 * a + 1] = det_log(p + 1e-8) - det_log(ap + 1e-8);
        } else {
            obs[2 * a] = yp - yd;
            obs[2 * a + 1] = p - ap;
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* One period in the order of Cats.step (cats.py:355-419).  Phase ids are the ones docs/MODEL.md
 * section 6 lists; stop_phase lets a test halt after any of them.                             */
#define PHASE(id, call) do { call; if (e->stop_phase == (id)) return; } while (0)

void oracle_step(Econ *e, const double *actions, const uint8_t *mask, int n_agents, unsigned flags,
                 const uint8_t *tape, double *obs, double *reward)
{
    double prevDe[64], prevP[64];
    double *pde = prevDe, *pp = prevP;
    if (n_agents > 64) { pde = (double *)xcalloc((size_t)n_agents, 8); pp = (double *)xcalloc((size_t)n_agents, 8); }
    e->tape = tape;
    e->strict = (flags & FLAG_STRICT) != 0;
    /* transients carry nothing across periods: clear them so that debug images are well defined */
    memset(e->c_Ld, 0, 4u * (size_t)e->F); memset(e->c_vac, 0, 4u * (size_t)e->F);
    memset(e->k_Ld, 0, 4u * (size_t)e->N); memset(e->k_vac, 0, 4u * (size_t)e->N);
    memset(e->c_Ftot, 0, 8u * (size_t)e->F); memset(e->c_Kdem, 0, 8u * (size_t)e->F);
    memset(e->c_Ystock, 0, 8u * (size_t)e->F); memset(e->c_valinv, 0, 8u * (size_t)e->F);
    memset(e->k_Ftot, 0, 8u * (size_t)e->N); memset(e->k_Ystock, 0, 8u * (size_t)e->N);
    memset(e->income, 0, 8u * (size_t)e->W); memset(e->budget, 0, 8u * (size_t)e->H);
    PHASE(1, ph_reset_gov_bank(e));
    PHASE(2, ph_bond_rate(e));
    for (int a = 0; a < n_agents; ++a) { pde[a] = e->c[C_De][a]; pp[a] = e->c[C_P][a]; }
    PHASE(3, ph_price_quantity_c(e));
    PHASE(4, ph_price_quantity_k(e));
    PHASE(5, ph_investment(e));
    PHASE(6, ph_action(e, actions, mask, n_agents, flags, pde, pp));
    if (pde != prevDe) { free(pde); free(pp); }
    PHASE(7, ph_labour_c(e));
    PHASE(8, ph_labour_k(e));
    PHASE(9, ph_credit_c(e));
    PHASE(10, ph_credit_k(e));
    PHASE(11, ph_fire_c(e));
    PHASE(12, ph_fire_k(e));
    PHASE(13, ph_job_search(e));
    PHASE(14, ph_produce_c(e));
    PHASE(15, ph_produce_k(e));
    PHASE(16, ph_capital_market(e));
    PHASE(17, ph_gov_subsidy(e));
    PHASE(18, ph_get_paid(e));
    PHASE(19, ph_cons_budget(e));
    PHASE(20, ph_goods_market(e));
    PHASE(21, ph_accounting_c(e));
    PHASE(22, ph_accounting_k(e));
    if (reward) ph_reward(e, n_agents, flags, reward);
    if (e->stop_phase == 23) return;
    PHASE(24, ph_tracking(e));
    PHASE(25, ph_bankrupt(e));
    PHASE(26, ph_bank_accounting(e));
    PHASE(27, ph_gov_accounting(e));
    PHASE(28, ph_adjust_wages(e));
    e->scal[S_timestep] = e->scal[S_timestep] + 1.0;
    if (obs) oracle_obs(e, n_agents, flags, obs);
    e->tape = NULL;
}

void oracle_set_stop_phase(Econ *e, int p) { e->stop_phase = p; }
void oracle_set_bankruptcy_reward(Econ *e, double r) { e->bankruptcy_reward = r; }

/* ------------------------------------------------------------------------------------------ */
/* Write the draws the Philox source WOULD produce for the coming period into a tape record.
 * Used by the tests to show that tape replay == Philox, and as the format example for a
 * foreign (Julia-side) recorder.                                                              */
void oracle_fill_tape(Econ *e, uint8_t *rec)
{
    const uint8_t *saved = e->tape; e->tape = NULL;
    TapeLayout t = econ_tl(e);
    memset(rec, 0, t.bytes);
    for (int f = 0; f < e->F; ++f) { double u = draw_uniform(e, PH_PRICE_C, f); memcpy(rec + t.u_price_c + 8u * (size_t)f, &u, 8); }
    for (int i = 0; i < e->N; ++i) { double u = draw_uniform(e, PH_PRICE_K, i); memcpy(rec + t.u_price_k + 8u * (size_t)i, &u, 8); }
    for (int f = 0; f < e->F; ++f) { double u = draw_uniform(e, PH_INVEST, f); memcpy(rec + t.u_invest + 8u * (size_t)f, &u, 8); }
    for (int w = 0; w < e->W; ++w) { uint32_t k = draw_fire_key(e, w); memcpy(rec + t.fire_key + 4u * (size_t)w, &k, 4); }
    Order o; int32_t cand[MAX_Z];
    int nf = e->F + e->N;
    int ze = e->z_e < nf ? e->z_e : nf, zk = e->z_k < e->N ? e->z_k : e->N, zc = e->z_c < e->F ? e->z_c : e->F;
    order_init(e, MKT_JOB, e->W, &o);
    for (int p = 0; p < e->W; ++p) { int32_t v = order_at(&o, p); memcpy(rec + t.job_order + 4u * (size_t)p, &v, 4); }
    for (int w = 0; w < e->W; ++w) { draw_candidates(e, PH_JOB, w, nf, ze, cand); memcpy(rec + t.job_cand + 4u * (size_t)w * (size_t)e->z_e, cand, 4u * (size_t)ze); }
    order_init(e, MKT_CAP, e->F, &o);
    for (int p = 0; p < e->F; ++p) { int32_t v = order_at(&o, p); memcpy(rec + t.cap_order + 4u * (size_t)p, &v, 4); }
    for (int f = 0; f < e->F; ++f) { draw_candidates(e, PH_CAP, f, e->N, zk, cand); memcpy(rec + t.cap_cand + 4u * (size_t)f * (size_t)e->z_k, cand, 4u * (size_t)zk); }
    order_init(e, MKT_GOODS, e->H, &o);
    for (int p = 0; p < e->H; ++p) { int32_t v = order_at(&o, p); memcpy(rec + t.goods_order + 4u * (size_t)p, &v, 4); }
    for (int h = 0; h < e->H; ++h) { draw_candidates(e, PH_GOODS, h, e->F, zc, cand); memcpy(rec + t.goods_cand + 4u * (size_t)h * (size_t)e->z_c, cand, 4u * (size_t)zc); }
    e->tape = saved;
}

void oracle_tape_offsets(int W, int F, int N, int z_c, int z_k, int z_e, uint64_t *out11)
{
    TapeLayout t = tape_layout(W, F, N, z_c, z_k, z_e);
    out11[0] = t.u_price_c; out11[1] = t.u_price_k; out11[2] = t.u_invest; out11[3] = t.fire_key;
    out11[4] = t.job_order; out11[5] = t.job_cand; out11[6] = t.cap_order; out11[7] = t.cap_cand;
    out11[8] = t.goods_order; out11[9] = t.goods_cand; out11[10] = t.bytes;
}

/* ------------------------------------------------------------------------------------------ */
/* field access by name (tests compare these with the GPU blob's views)                        */
static const char *C_NAMES[NC_FIELDS] = { "De", "P", "Y_prev", "Yd", "Q", "A", "K", "investment",
    "capital_value", "wages", "interests", "deb", "liquidity", "barYK", "interest_r", "div_income" };
static const char *K_NAMES[NK_FIELDS] = { "De", "P", "Y_prev", "Yd", "Q", "A", "wages", "interests",
    "deb", "liquidity", "interest_r", "div_income", "inv", "Yavail" };

double *oracle_f64(Econ *e, const char *name, int *n)
{
    if (!strncmp(name, "c_", 2)) {
        for (int i = 0; i < NC_FIELDS; ++i) if (!strcmp(name + 2, C_NAMES[i])) { *n = e->F; return e->c[i]; }
        if (!strcmp(name, "c_Ftot")) { *n = e->F; return e->c_Ftot; }
        if (!strcmp(name, "c_Kdem")) { *n = e->F; return e->c_Kdem; }
        if (!strcmp(name, "c_Ystock")) { *n = e->F; return e->c_Ystock; }
        if (!strcmp(name, "c_valinv")) { *n = e->F; return e->c_valinv; }
    }
    if (!strncmp(name, "k_", 2)) {
        for (int i = 0; i < NK_FIELDS; ++i) if (!strcmp(name + 2, K_NAMES[i])) { *n = e->N; return e->k[i]; }
        if (!strcmp(name, "k_Ftot")) { *n = e->N; return e->k_Ftot; }
        if (!strcmp(name, "k_Ystock")) { *n = e->N; return e->k_Ystock; }
    }
    if (!strcmp(name, "PA")) { *n = e->H; return e->PA; }
    if (!strcmp(name, "perm")) { *n = e->H; return e->perm; }
    if (!strcmp(name, "budget")) { *n = e->H; return e->budget; }
    if (!strcmp(name, "income")) { *n = e->W; return e->income; }
    if (!strcmp(name, "scal")) { *n = N_SCAL; return e->scal; }
    *n = 0; return NULL;
}

int32_t *oracle_i32(Econ *e, const char *name, int *n)
{
    if (!strcmp(name, "c_Leff")) { *n = e->F; return e->c_Leff; }
    if (!strcmp(name, "k_Leff")) { *n = e->N; return e->k_Leff; }
    if (!strcmp(name, "Oc")) { *n = e->W; return e->Oc; }
    if (!strcmp(name, "c_Ld")) { *n = e->F; return e->c_Ld; }
    if (!strcmp(name, "c_vac")) { *n = e->F; return e->c_vac; }
    if (!strcmp(name, "k_Ld")) { *n = e->N; return e->k_Ld; }
    if (!strcmp(name, "k_vac")) { *n = e->N; return e->k_vac; }
    *n = 0; return NULL;
}

/* primitives exposed for known-answer tests */
double oracle_det_exp(double x) { return det_exp(x); }
double oracle_det_log(double x) { return det_log(x); }
double oracle_det_sum(const double *x, int n) { return det_sum(x, n); }
void oracle_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t *out) { philox4x32_10(c0, c1, c2, c3, k0, k1, out); }
void oracle_sample_distinct(const uint32_t *words, int n, int z, int32_t *out) { sample_distinct(words, n, z, out); }
void oracle_perm(Econ *e, int market, int n, int32_t *out)
{
    Perm p; perm_init(e, market, n, &p);
    for (int i = 0; i < n; ++i) out[i] = perm_at(&p, i);
}

/* ------------------------------------------------------------------------------------------ */
/* batch driver: B independent economies, heuristic (burn-in style) periods, optional OpenMP.
 * Used for the CPU baseline timing and for large parity runs.  stats: [B][n_periods][16].      */
#define N_STATS 16
static void fill_stats(const Econ *e, double *st)
{
    const double *s = e->scal;
    st[0] = s[S_Y_real]; st[1] = s[S_wb]; st[2] = s[S_Un]; st[3] = s[S_bankruptcy_rate];
    st[4] = s[S_price]; st[5] = s[S_Y_nominal_tot]; st[6] = s[S_gdp_deflator];
    st[7] = s[S_inflationRate]; st[8] = s[S_consumption]; st[9] = s[S_totK];
    st[10] = s[S_dividendsB]; st[11] = s[S_profitsB]; st[12] = s[S_GB]; st[13] = s[S_deficitGDP];
    st[14] = s[S_gov_bonds]; st[15] = s[S_price_k];
}

/* The same driver with a draw tape [B][n_periods] records and step flags (FLAG_STRICT ...).
 * fill != 0: each record is first written with what the Philox source would draw (oracle_fill_tape) and
 * then replayed; fill == 0: the records are the caller's (a foreign generator).  tape == NULL: Philox. */
void oracle_run_batch_tape(Econ **econs, int B, int n_periods, uint8_t *tape, int fill, unsigned flags,
                           double *stats, int n_threads)
{
    (void)n_threads;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads > 0 ? n_threads : 1)
#endif
    for (int b = 0; b < B; ++b) {
        const size_t tb = tape ? econ_tl(econs[b]).bytes : 0;
        for (int t = 0; t < n_periods; ++t) {
            uint8_t *rec = tape ? tape + ((size_t)b * (size_t)n_periods + (size_t)t) * tb : NULL;
            if (rec && fill) oracle_fill_tape(econs[b], rec);
            oracle_step(econs[b], NULL, NULL, 0, FLAG_BURNIN | flags, rec, NULL, NULL);
            if (stats) fill_stats(econs[b], stats + ((size_t)b * (size_t)n_periods + (size_t)t) * N_STATS);
        }
    }
}

void oracle_run_batch(Econ **econs, int B, int n_periods, double *stats, int n_threads)
{
    (void)n_threads;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads > 0 ? n_threads : 1)
#endif
    for (int b = 0; b < B; ++b) {
        for (int t = 0; t < n_periods; ++t) {
            oracle_step(econs[b], NULL, NULL, 0, FLAG_BURNIN, NULL, NULL, NULL);
            if (stats) fill_stats(econs[b], stats + ((size_t)b * (size_t)n_periods + (size_t)t) * N_STATS);
        }
    }
}
