// cats_aux.cuh -- the two small kernels beside the period kernel: the initial state (a0, initialise_model)
// and the fused tabular-Q policy step (qlearners.py:63-166).  Launched by cats_kernels.cu.
#ifndef CATS_AUX_CUH
#define CATS_AUX_CUH
#include "cats_period.cuh"

namespace cats {

// a0: initialise_model (cats.py:224) -- deterministic initial state, docs/MODEL.md section 5
// `active` (nullable): only economies with a non-zero entry are re-initialised, and their episode
// counter -- the salt of their Philox stream -- goes up by one instead of back to zero.
__global__ void cats_init_kernel(unsigned char *blobs, long long B, Layout L, const double *params,
                                 long long params_stride, const unsigned char *active)
{
    const long long b = blockIdx.x;
    if (b >= B || (active && !active[b])) return;
    unsigned char *blob = blobs + (size_t)b * (size_t)L.blob;
    const double episode = active ? reinterpret_cast<const double *>(blob + L.o_scal)[S_episode] + 1.0 : 0.0;
    __syncthreads();
    const double *par = params + (params_stride ? (size_t)b * (size_t)params_stride : 0);
    const int tid = threadIdx.x, nt = blockDim.x;
    // blobs are multiples of 128 bytes and 128-byte aligned: 16-byte stores
    for (int i = tid; i < L.blob / 16; i += nt) reinterpret_cast<uint4 *>(blob)[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    const double P0 = 3.0, PK0 = 3.0, K0 = 10.0, LIQ0 = 10.0, YC0 = 5.0, YK0 = 3.0, PA0 = 2.0;
    double *scal = reinterpret_cast<double *>(blob + L.o_scal);
    if (tid == 0) {
        scal[S_price] = P0; scal[S_price_k] = PK0; scal[S_init_price] = P0; scal[S_init_price_k] = PK0;
        scal[S_wb] = par[P_wb]; scal[S_bank_E] = 3000.0; scal[S_gdp_deflator] = 1.0; scal[S_Un] = 1.0;
        scal[S_gov_r_b] = par[P_interest_rate];
        scal[S_episode] = episode;
    }
    auto cf = [&](int f) { return reinterpret_cast<double *>(blob + L.o_c + f * L.rowF); };
    auto kf = [&](int f) { return reinterpret_cast<double *>(blob + (f < NK_HOT ? L.o_kh + f * L.rowN : L.o_kc + (f - NK_HOT) * L.rowN)); };
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
    double2 *hh = reinterpret_cast<double2 *>(blob + L.o_PA);
    for (int h = tid; h < L.H; h += nt) hh[h] = make_double2(PA0, 1.0 / P0);
    int *Oc = reinterpret_cast<int *>(blob + L.o_Oc);
    for (int w = tid; w < L.W; w += nt) Oc[w] = -1;
}

// ---------------------------------------------------------------------------------------------
// Tabular Q policy (cats_q_step): one warp per (economy, agent).  The row Q[s', ., .] is read once for
// the bootstrap maximum; the TD update touches one cell; the greedy choice re-reads the row only when
// the updated cell lies in it (s == s').
struct QArgs {
    double *Q; long long B; int A; int n[4];
    double obs_lo[2], obs_step[2], act_lo[2], act_hi[2];
    double alpha, gamma, epsilon;
    uint32_t rk[20];
    unsigned long long economy_base; uint32_t step; int do_train, do_act;
    const double *obs, *reward; const unsigned char *terminated; int *last; double *actions;
    const unsigned char *active;
};

__device__ __forceinline__ int q_bin(double x, double lo, double step, int n)
{
    double idx = ceil((x - lo) / step - 0.5);
    if (idx != idx) idx = 0.0;                       // NaN -> 0
    if (idx < 0.0) idx = 0.0;
    if (idx > (double)(n - 1)) idx = (double)(n - 1);
    return (int)idx;
}

// (max value, first index of it) over row[0, m): lanes stride, then a butterfly on (value, index)
__device__ __forceinline__ void q_row_max(const double *row, int m, int lane, double &best, int &arg)
{
    best = __longlong_as_double(0xfff0000000000000LL);   // -inf
    arg = 0x7fffffff;
    for (int i = lane; i < m; i += 32) {
        const double v = row[i];
        if (v > best || (v == best && i < arg)) { best = v; arg = i; }
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const double ov = __shfl_xor_sync(FULL, best, off);
        const int oa = __shfl_xor_sync(FULL, arg, off);
        if (ov > best || (ov == best && oa < arg)) { best = ov; arg = oa; }
    }
    if (arg == 0x7fffffff) arg = 0;   // a row of NaNs compares false everywhere: act on cell 0, as numpy's argmax of such a row does
}

__global__ void cats_q_kernel(const __grid_constant__ QArgs q)
{
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;   // (economy, agent) pair
    const int lane = threadIdx.x & 31;
    if (w >= q.B * q.A) return;
    const long long b = w / q.A; const int a = (int)(w % q.A);
    if (q.active && !q.active[b]) return;
    const int n0 = q.n[0], n1 = q.n[1], n2 = q.n[2], n3 = q.n[3], m = n2 * n3;
    double *Qa = q.Q + (size_t)w * (size_t)n0 * (size_t)n1 * (size_t)m;
    const double o0 = q.obs[2 * w], o1 = q.obs[2 * w + 1];
    const int s0n = q_bin(o0, q.obs_lo[0], q.obs_step[0], n0), s1n = q_bin(o1, q.obs_lo[1], q.obs_step[1], n1);
    double *row = Qa + ((size_t)s0n * n1 + s1n) * m;
    double best; int arg;
    q_row_max(row, m, lane, best, arg);
    int *last = q.last + 4 * w;
    if (q.do_train) {
        const int s0 = last[0], s1 = last[1], a0 = last[2], a1 = last[3];
        __syncwarp();
        if (lane == 0) {
            const double boot = q.terminated[b] ? 0.0 : q.gamma * best;
            const double target = q.reward[w] + boot;
            double *cell = Qa + (((size_t)s0 * n1 + s1) * n2 + a0) * n3 + a1;
            const double cur = *cell;
            *cell = cur + q.alpha * (target - cur);
        }
        __syncwarp();
        if (q.do_act && s0 == s0n && s1 == s1n) q_row_max(row, m, lane, best, arg);   // the row changed
    }
    if (q.do_act && lane == 0) {
        const uint4 r = philox4x32_10((uint32_t)a, 0x51u, q.step, (uint32_t)(q.economy_base + (unsigned long long)b), q.rk);
        int a0 = arg / n3, a1 = arg - a0 * n3;
        if ((double)r.x * (1.0 / 4294967296.0) < q.epsilon) { a0 = (int)__umulhi(r.y, (uint32_t)n2); a1 = (int)__umulhi(r.z, (uint32_t)n3); }
        last[0] = s0n; last[1] = s1n; last[2] = a0; last[3] = a1;
        q.actions[2 * w] = q.act_lo[0] + (double)a0 * (q.act_hi[0] - q.act_lo[0]) / (double)(n2 - 1);
        q.actions[2 * w + 1] = q.act_lo[1] + (double)a1 * (q.act_hi[1] - q.act_lo[1]) / (double)(n3 - 1);
    }
}


// the kernel parameter block of one cats_step call (everything but the launch shape: slice, bars)
inline void fill_kargs(const cats_step_args *s, KArgs &a)
{
    a.L = make_layout(s->W, s->F, s->N);
    a.T = TapeOff{};
    a.od[MKT_JOB] = make_order_dims(s->W); a.od[MKT_CAP] = make_order_dims(s->F);
    a.od[MKT_GOODS] = make_order_dims(s->W + s->F + s->N);
    a.tape_z[0] = a.tape_z[1] = a.tape_z[2] = 0;
    a.blobs = (unsigned char *)s->d_blobs; a.B = s->B;
    a.params = s->d_params; a.params_stride = s->params_stride;
    for (uint32_t r = 0; r < 10; ++r) {
        a.rk[2 * r] = (uint32_t)s->seed + r * 0x9E3779B9u;
        a.rk[2 * r + 1] = (uint32_t)(s->seed >> 32) + r * 0xBB67AE85u;
    }
    a.economy_base = s->economy_base;
    a.n_periods = s->n_periods; a.flags = s->flags; a.n_agents = s->n_agents;
    a.actions = s->d_actions; a.mask = s->d_action_mask;
    a.bankruptcy_reward = s->bankruptcy_reward;
    a.obs = s->d_obs; a.reward = s->d_reward; a.terminated = s->d_terminated; a.status = s->d_status;
    a.stats = s->d_stats; a.tape = s->d_tape; a.stop_phase = s->stop_phase;
    a.debug = (unsigned char *)s->d_debug;
    a.cycles = nullptr;
    a.active = s->d_active;
    if (s->d_tape) {
        a.T = make_tape(s->W, s->F, s->N, s->tape_z[0], s->tape_z[1], s->tape_z[2]);
        a.tape_z[0] = s->tape_z[0]; a.tape_z[1] = s->tape_z[1]; a.tape_z[2] = s->tape_z[2];
    }
    a.slice = ((a.L.smem + 15) & ~15);
    a.bars = s->stop_phase ? 0 : 1;
}

}  // namespace cats
#endif
