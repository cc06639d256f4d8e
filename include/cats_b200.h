/*
 * cats_b200.h -- C-ABI of the B200-native batched CATS economy engine (libcats_b200.so).
 *
 * This is the drop-in boundary for the ONE hot path of Brusa99/R-MABM: the per-period model
 * step that `rmabm.environments.Cats.step` delegates to ABCredit.jl through juliacall
 * (/root/reference/rmabm/environments/cats.py:355-410).  Every entry point names the reference
 * interface it replaces.  Plain pointers and sizes only: no torch, no C++ types.
 *
 * Conventions
 *   - All `d_*` pointers are DEVICE pointers borrowed from the caller (e.g. torch tensors'
 *     data_ptr()); the library never allocates, frees or retains them.
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous on it and never
 *     synchronises the host (except the *_host convenience calls, which say so).
 *   - Return value: 0 on success, a negative CATS_E* code otherwise; cats_last_error() gives text.
 *   - State lives in one contiguous "economy blob" per economy (docs/MODEL.md section 2): the
 *     caller allocates B * cats_blob_bytes(W,F,N) bytes, 128-byte aligned.  Field views are
 *     described by cats_field_info().
 *   - Economies are independent.  `economy_base` is the GLOBAL id of blob 0 of this call, so a
 *     shard on GPU g passes its slab offset and gets bit-identical economies at any sharding.
 */
#ifndef CATS_B200_H
#define CATS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CATS_ABI_VERSION 2
#define CATS_N_PARAMS 34   /* Cats._default_parameters, cats.py:69-104, in that key order */
#define CATS_N_SCAL 32
#define CATS_N_STATS 16
#define CATS_MAX_Z 8

/* error codes */
#define CATS_OK 0
#define CATS_EINVAL (-1)   /* bad argument (the Python side raises ValueError) */
#define CATS_ECUDA (-2)    /* CUDA runtime error; see cats_last_error() */
#define CATS_ESMEM (-3)    /* economy too large for the on-chip working set of this build */

/* step flags */
#define CATS_FLAG_BURNIN 1u      /* cats.py:369 -- `burnin=True`: actions ignored */
#define CATS_FLAG_LOG 2u         /* CatsLog transforms, cats.py:655-704 */
#define CATS_FLAG_AVG_PRICE 4u   /* price_change == "avg_price", cats.py:450-451 */
#define CATS_FLAG_REWARD_RMS 8u  /* reward_type == "rms", cats.py:544-556 */
#define CATS_FLAG_STRICT 16u     /* plain arithmetic (docs/MODEL.md 4.6): purchases and real values divide by the price, demand at a
                                    seller is an fp64 sum in visiting order.  Slower; the mode a tape recorded from ABCredit.jl
                                    would be replayed in if the reference divides.  Not with d_active. */

/* per-economy status bits (replace the reference's warnings.warn, cats.py:518,535) */
#define CATS_STATUS_DEPVAL_DIV0 1u
#define CATS_STATUS_PROFIT_NAN 2u
#define CATS_STATUS_TAPE_RANGE 8u    /* replay: the tape held a candidate or a visiting-order entry outside its range (taken as 0): results invalid */
#define CATS_STATUS_Z_MISMATCH 4u   /* cats_step_args.z_all was set but this economy's parameters disagree: its results are invalid */

/* columns of the per-period statistics row (the aggregates Cats._get_info reads, cats.py:567-607) */
enum cats_stat {
    CATS_STAT_Y_real = 0, CATS_STAT_wb, CATS_STAT_Un, CATS_STAT_bankruptcy_rate,
    CATS_STAT_avg_price, CATS_STAT_Y_nominal_tot, CATS_STAT_gdp_deflator, CATS_STAT_inflationRate,
    CATS_STAT_consumption, CATS_STAT_totK, CATS_STAT_dividendsB, CATS_STAT_profitsB, CATS_STAT_GB,
    CATS_STAT_deficitGDP, CATS_STAT_bonds, CATS_STAT_price_k
};

int cats_abi_version(void);
const char *cats_last_error(void);

/* Size in bytes of one economy blob (multiple of 128).  This is the state `initialise_model`
 * builds (cats.py:224); 2 * cats_blob_bytes() is the algorithmic HBM traffic of one
 * economy-period (docs/DESIGN.md, CATS_ALGO_BYTES). */
size_t cats_blob_bytes(int W, int F, int N);

/* Describe a named state field inside the blob.  Names: "scal" (32 doubles, see docs/MODEL.md),
 * C-firm fields "c_De c_P c_Y_prev c_Yd c_Q c_A c_K c_investment c_capital_value c_wages
 * c_interests c_deb c_liquidity c_barYK c_interest_r c_div_income" (f64 x F), "c_Leff" (i32 x F),
 * K-firm fields "k_De k_P k_Y_prev k_Yd k_Q k_A k_wages k_interests k_deb k_liquidity
 * k_interest_r k_div_income k_inv k_Yavail" (f64 x N), "k_Leff" (i32 x N), "PA", "perm" (f64 x H),
 * "Oc" (i32 x W).  This replaces the juliacall struct-attribute protocol (cats.py:227-243,
 * 359-360, 472-473, 486, 512-538, 564-594).  dtype: 0 = float64, 1 = int32. */
int cats_field_info(int W, int F, int N, const char *name, uint64_t *byte_offset, int32_t *count,
                    int32_t *dtype);
/* enumerate field names: returns NULL past the end */
const char *cats_field_name(int index);

/* Bytes of one draw-tape record (one economy-period) and the byte offsets of its 10 sections
 * (u_price_c u_price_k u_invest fire_key job_order job_cand cap_order cap_cand goods_order
 * goods_cand), then the record size.  docs/MODEL.md section 3.4. */
size_t cats_tape_bytes(int W, int F, int N, int z_c, int z_k, int z_e);
int cats_tape_offsets(int W, int F, int N, int z_c, int z_k, int z_e, uint64_t out11[11]);

/* Replaces ABCredit.initialise_model(W, F, N, params)  (cats.py:224).
 * d_params: device [B][34] doubles when params_stride == 34, or one shared [34] row when
 * params_stride == 0. */
int cats_init(void *d_blobs, int64_t B, int W, int F, int N, const double *d_params,
              int64_t params_stride, void *stream);

/* Per-economy reset, for episodes that end at different times (a new model per `Cats.reset`,
 * cats.py:283-315, one economy at a time): re-initialises the economies whose d_active[b] != 0 and
 * leaves the others untouched.  A reset economy's EPISODE counter (scal slot "episode") goes up by
 * one; it is part of the Philox counter (docs/MODEL.md 3.1), so the new episode draws a fresh stream
 * while staying a pure function of (seed, global economy id, episode, period).  cats_init() puts
 * the counter back to zero.  Follow with cats_step(n_periods = t_burnin, BURNIN flag, the same
 * d_active) for the burn-in of the reset economies only. */
int cats_reset_masked(void *d_blobs, int64_t B, int W, int F, int N, const double *d_params,
                      int64_t params_stride, const uint8_t *d_active, void *stream);

typedef struct cats_step_args {
    void *d_blobs;              /* [B] economy blobs */
    int64_t B;
    int32_t W, F, N;
    const double *d_params;     /* as for cats_init */
    int64_t params_stride;
    uint64_t seed;              /* Philox key (cats.py:306 seeds Julia's global RNG instead) */
    uint64_t economy_base;      /* global id of blob 0 */
    int32_t n_periods;          /* periods advanced by this launch (>1 only with BURNIN flag or n_agents==0) */
    uint32_t flags;             /* CATS_FLAG_* */
    int32_t n_agents;           /* RL-controlled C-firms: the first n_agents slots (cats.py:242) */
    const double *d_actions;    /* [B][n_agents][2] or NULL */
    const uint8_t *d_action_mask; /* [B][n_agents], 0 = "dummy" (cats.py:439-440); NULL = all real */
    double bankruptcy_reward;   /* cats.py:124,538-539 */
    double *d_obs;              /* out [B][n_agents][2]  (cats.py:466-481) or NULL */
    double *d_reward;           /* out [B][n_agents]     (cats.py:493-556) or NULL */
    uint8_t *d_terminated;      /* out [B]               (cats.py:483-486) or NULL */
    uint32_t *d_status;         /* out [B] or NULL: the CATS_STATUS_* bits raised by THIS call (the reference warns on the step
                                   where the event occurs, cats.py:516-536); scal slot "status" keeps the sticky OR since cats_init */
    double *d_stats;            /* out [B][n_periods][CATS_N_STATS] or NULL */
    const uint8_t *d_tape;      /* replay: [B][n_periods] records of cats_tape_bytes(); NULL = Philox */
    int32_t tape_z[3];          /* z_c, z_k, z_e the tape records were laid out with (only read when d_tape != NULL) */
    int32_t max_z;              /* upper bound on z_c, z_k, z_e over all economies of the call (1..8);
                                   0 = unknown (8).  <= 5 selects the leaner kernel instantiation. */
    int32_t stop_phase;         /* debug: halt the (single) period after this phase id; 0 = off */
    void *d_debug;              /* debug: [B][cats_workset_bytes] image of the on-chip working set */
    void *stream;
    const uint8_t *d_active;    /* DEVICE [B] or NULL: economies with a zero entry sit this call out (state
                                   and outputs untouched); NULL = all.  See cats_reset_masked().  Cannot be combined with
                                   d_tape, stop_phase or d_debug (CATS_EINVAL). */
    int32_t z_all[3];           /* z_c, z_k, z_e when EVERY economy of the call has exactly these values; zeros =
                                   unknown or mixed.  Only a hint for picking a kernel instantiation (the reference's
                                   defaults 5, 2, 5 at the default size have one with the lists unrolled); an economy
                                   that disagrees gets CATS_STATUS_Z_MISMATCH. */
} cats_step_args;

/* Replaces the 26 ABCredit calls of one Cats.step (cats.py:355-413) plus _implement_action,
 * _get_reward, _get_terminated and _get_obs, for B economies, n_periods times. */
int cats_step(const cats_step_args *args);

/* Same call with HOST action/obs/reward/terminated/stats buffers (pageable or pinned) and a stream
 * synchronisation at the end: the outputs are in the host buffers when it returns.  The actions are copied
 * host->device ahead of the step.  Page-locked OUTPUT buffers (cudaHostAlloc / cudaHostRegister / torch
 * pin_memory: mapped under unified addressing) are written IN PLACE: every economy posts its outputs over
 * PCIe while the others are still computing, so no copy is left when the kernel ends.  Pageable output
 * buffers are staged through d_scratch and copied out.  d_scratch must hold
 * cats_host_scratch_bytes(B, n_agents, n_periods) device bytes either way. */
size_t cats_host_scratch_bytes(int64_t B, int n_agents, int n_periods);
/* Byte offsets of the seven sections of that staging block (actions, mask, obs, reward, terminated,
 * status, stats).  A caller whose HOST buffers sit at the same offsets inside one (pinned) block gets a
 * single copy in each direction from cats_step_host instead of one per buffer. */
int cats_host_block_offsets(int64_t B, int n_agents, int n_periods, uint64_t out7[7]);
int cats_step_host(const cats_step_args *args_with_host_io, void *d_scratch);

/* ---- learned firms: tabular Q-learning policy, one table per (economy, agent) --------------------
 * The batched twin of rmabm.agents.QLearner (/root/reference/rmabm/agents/qlearners.py:63-166) as ONE
 * launch per rollout step: the TD(0) update for the transition just observed (`train`,
 * qlearners.py:130-166), then the epsilon-greedy choice for the new observation (`get_action`,
 * qlearners.py:97-128).  Observations, rewards and actions are the device buffers cats_step reads and
 * writes: nothing leaves the GPU.
 *   binning: index of the nearest of n poles linearly spaced over [lo, hi]; the lower pole wins a tie,
 *            values outside clamp, NaN -> 0 (qlearners.py:63-95)
 *   train:   Q[s,a] += alpha * ((r + (terminated ? 0 : gamma * max Q[s', .])) - Q[s,a])
 *   act:     one Philox-4x32-10 block per (economy, agent, step): counter (agent, 0x51, step, global
 *            economy id), key = seed;  explore <=> w0 * 2^-32 < epsilon, then a0 = mulhi(w1, n2),
 *            a1 = mulhi(w2, n3); otherwise the first argmax of Q[s', ., .].  The action values are
 *            lo + index * (hi - lo) / (n - 1), as qlearners.py:121-126. */
typedef struct cats_q_args {
    double *d_Q;                 /* [B][A][n0][n1][n2][n3] in/out */
    int64_t B;
    int32_t A;
    int32_t n[4];                /* bins: firm_stock, price_delta, production factor, price factor (each >= 2) */
    double obs_lo[2], obs_hi[2]; /* gym_spaces_bounds obs_firm_stock / obs_price_delta (cats.py:106-111) */
    double act_lo[2], act_hi[2]; /* act_production_factor / act_price_factor */
    double alpha, gamma, epsilon;
    uint64_t seed;
    uint64_t economy_base;       /* global id of economy 0 of this call */
    uint32_t step;               /* rollout step counter (exploration stream position) */
    int32_t do_train, do_act;
    const double *d_obs;         /* [B][A][2]: the observation AFTER the step (s'); the state to act on */
    const double *d_reward;      /* [B][A]  (train only) */
    const uint8_t *d_terminated; /* [B]     (train only) */
    int32_t *d_last;             /* [B][A][4] in/out: (s0, s1, a0, a1) of the action in flight */
    double *d_actions;           /* out [B][A][2] (act only): feeds cats_step_args.d_actions */
    const uint8_t *d_active;     /* [B] or NULL: economies with a zero entry are skipped */
    void *stream;
} cats_q_args;
int cats_q_step(const cats_q_args *args);

/* Size of the on-chip working set image written to d_debug (blob + transients). */
size_t cats_workset_bytes(int W, int F, int N);
/* Element stride of a named state field, in elements of its dtype: 1 for contiguous rows, 2 for
 * "PA" and "perm" (households are stored as {PA, perm} records: element i of the field lives at
 * byte_offset + i * stride * sizeof(dtype)).  0 for an unknown name. */
int cats_field_stride(const char *name);

/* Describe a transient field inside the working-set image (same contract as cats_field_info).
 * Names: c_Ld c_vac k_Ld k_vac (i32), c_Ftot c_Kdem c_Ystock c_valinv k_Ftot k_Ystock (f64). */
int cats_workset_field_info(int W, int F, int N, const char *name, uint64_t *byte_offset,
                            int32_t *count, int32_t *dtype);

/* Launch geometry actually used for (W,F,N): threads per block (32 per economy, several economies
 * per block), dynamic shared bytes per block, economies resident per SM, number of SMs. */
int cats_launch_info(int W, int F, int N, int32_t *threads, int32_t *smem_bytes,
                     int32_t *economies_per_sm, int32_t *num_sms);

/* The kernel shape cats_step picks for a production launch of B economies of this size on the current device: warps
 * per economy (1 = the one-warp-per-economy kernel, 4 / 8 = one thread block per economy), economies per block (one-warp
 * kernel), and whether the multi-warp kernel runs its goods market as the producer / consumer pipeline. */
int cats_launch_shape(int W, int F, int N, int64_t B, int32_t *warps_per_economy, int32_t *economies_per_block,
                      int32_t *pipelined_goods_market);

/* Number of kernels this library has launched since load (bench.py's gpu_launches). */
uint64_t cats_launch_count(void);

/* Test and profiling hooks (process-wide settings that never change results) live in cats_b200_debug.h: they are
 * not part of the drop-in boundary. */

#ifdef __cplusplus
}
#endif
#endif /* CATS_B200_H */
