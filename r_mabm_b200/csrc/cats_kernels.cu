// cats_kernels.cu -- C-ABI of libcats_b200.so: argument checks, launch geometry, the initial-state kernel.
// The period kernel is in cats_period.cuh; its instantiations are compiled in cats_inst_*.cu.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

#include <atomic>
#include <mutex>

#include "cats_period.cuh"
#include "../../include/cats_b200_debug.h"

#include "cats_aux.cuh"
#include "cats_period_mw.cuh"

// =============================================================================================
// C-ABI
using namespace cats;

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};
static std::atomic<long long *> g_cycles{nullptr};  // test / profiling hook: cats_debug_phase_cycles()
static std::atomic<int> g_group{0};    // test hook: cats_debug_set_group() (0 = automatic)
static std::atomic<int> g_static{1};   // test hook: cats_debug_set_static(0) keeps the generic kernel for the default size
static std::atomic<int> g_warps{0};    // test hook: cats_debug_set_warps(): warps per economy (0 = chosen per launch)
static std::atomic<int> g_zero_copy{2};   // test hook: cats_debug_set_host_zero_copy(): cats_step_host on mapped pinned buffers (bit 0: inputs read in place, bit 1: outputs written in place)
static std::atomic<int> g_pipe{-1};    // test hook: cats_debug_set_pipe(): goods market of the multi-warp kernel (-1 = chosen per launch)
// development aid: CATS_SMEM_PAD=<bytes> pads the dynamic shared memory of every launch (caps residency)
static int smem_pad()
{
    static const int pad = [] { const char *e = getenv("CATS_SMEM_PAD"); const int v = e ? atoi(e) : 0; return v < 0 ? 0 : v; }();
    return pad;
}
static bool is_default_size(const Layout &L)
{
    // development aid: CATS_NO_STATIC=1 keeps the generic kernel for every size
    static const int off = [] { const char *e = getenv("CATS_NO_STATIC"); return e ? atoi(e) : 0; }();
    return !off && g_static.load() && L.W == DefaultLayout::kW && L.F == DefaultLayout::kF && L.N == DefaultLayout::kN && smem_pad() == 0;
}
// production: Philox draws, no debug stop, no debug image
static KernelFn select_kernel(const Layout &L, int max_z, int g, bool masked = false, bool production = true,
                              bool default_z = false, bool strict = false)
{
    if (strict) return kernel_strict(!production ? 0 : (default_z && is_default_size(L)) ? 2 : 1, g);
    if (masked) return kernel_masked(g);   // resets are rare: one instantiation serves every size and z
    if (!production) return kernel_debug(max_z, g);
    if (default_z && is_default_size(L)) return kernel_static(g);
    return kernel_prod(max_z, g);
}
static int slice_bytes(const Layout &L) { return ((L.smem + smem_pad()) + 15) & ~15; }
// Economies per block (see cats_period_kernel).  Shared memory and the register file bound how many
// economies fit on an SM (at most kMaxPerSM).  The launch takes as long as its busiest SM; a period
// on an SM with n resident economies costs about (12 + n) units (measured: latency + contention), and
// economies that run in step save instruction fetches (measured factors below).  The cheapest shape
// wins.  CATS_GROUP=<1|2|4|8|12|14> overrides (development aid).
static int choose_group(const Layout &L, long long B, int sms)
{
    static const int env_forced = [] { const char *e = getenv("CATS_GROUP"); return e ? atoi(e) : 0; }();
    const int gg = g_group.load();
    const int forced = gg ? gg : env_forced;
    const int slice = slice_bytes(L);
    const int sm_bytes = 228 * 1024, block_max = 227 * 1024, reserve = 1024;
    const int cand[6] = {14, 12, 8, 4, 2, 1};
    if (sms < 1) sms = 1;
    int best = 1; double best_score = 1e300;
    for (int i = 0; i < 6; ++i) {
        const int g = cand[i];
        if (g > 1 && g * slice > block_max) continue;
        if (forced == g) return g;
        long long fit = sm_bytes / (g * (long long)slice + reserve);     // blocks resident per SM
        if (fit > kMaxPerSM / g) fit = kMaxPerSM / g;
        if (fit < 1) fit = 1;
        const long long blocks = (B + g - 1) / g;
        long long per_sm = (blocks + sms - 1) / sms;                     // blocks on the busiest SM ...
        const double waves = (double)((per_sm + fit - 1) / fit);         // ... in this many rounds of `fit`
        if (per_sm > fit) per_sm = fit;
        const double step[6] = {0.60, 0.62, 0.68, 0.75, 0.85, 1.0};       // g = 14, 12, 8, 4, 2, 1
        const double score = (waves > 1.0 ? waves : 1.0) * (12.0 + (double)(per_sm * g)) * step[i];
        if (score < best_score) { best_score = score; best = g; }
    }
    return best;
}

// Warps per economy.  The one-warp kernel fills the GPU from about 28 x SMs economies up; a smaller launch takes as
// long as ONE economy's dependent chain, so several warps share an economy instead (cats_period_mw.cuh): as many as
// keep ~28 warps on every SM, if the block's shared memory still lets every economy of the launch be resident.
// the multi-warp kernel's goods market: a pipeline (producer warps prepare, one warp commits with the one-warp
// kernel's rounds) or the block-wide market (cats_period_mw.cuh).  Since the one-warp rounds find the endangered lanes
// through the peer table the pipeline is the faster one at every size measured (1,024 default-size
// economies at 4 warps: 8.26e6 against 7.38e6; 256 economies of 10x size at 8 warps: 3.35e5 against 3.21e5; one
// default-size economy per SM at 8 warps: equal); the block-wide market stays selectable (cats_debug_set_pipe).
static bool choose_pipe(const Layout &, int)
{
    const int forced = g_pipe.load();
    return forced >= 0 ? forced != 0 : true;
}

static int choose_warps(const Layout &L, long long B, int sms)
{
    static const int env_forced = [] { const char *e = getenv("CATS_WARPS"); return e ? atoi(e) : 0; }();
    const int gw = g_warps.load();
    const int forced = gw ? gw : env_forced;
    if (forced == 1 || forced == 2 || forced == 4 || forced == 8) return forced;
    if (sms < 1) sms = 1;
    const long long want = (long long)kMaxPerSM * sms / (B > 0 ? B : 1);
    // two warps per economy never paid off, with either goods market: four or eight, or the one-warp kernel
    const int cand[2] = {8, 4};
    for (int i = 0; i < 2; ++i) {
        const int nw = cand[i];
        if (want < nw) continue;
        const long long bytes = slice_bytes(L) + mw_extra_bytes(nw, L.F) + 1024;
        long long fit = (228 * 1024) / bytes;
        const long long by_regs = nw == 8 ? 3 : 7;   // blocks per SM the kernels' __launch_bounds__ are cut for (cats_inst_mw.cu)
        if (fit > by_regs) fit = by_regs;
        if (bytes <= 227 * 1024 && fit * sms >= B) return nw;
    }
    return 1;
}

static int fail(int code, const char *fmt, const char *detail)
{
    snprintf(g_err, sizeof g_err, fmt, detail);
    return code;
}

static int check_sizes(int W, int F, int N)
{
    if (W < 1 || F < 1 || N < 1) return fail(CATS_EINVAL, "W, F, N must be >= 1%s", "");
    if (W > 65535 || F + N > 32767 || F > 16383)
        return fail(CATS_EINVAL, "economy too large for this build (W <= 65535, F <= 16383, F + N <= 32767)%s", "");
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
void cats_debug_set_group(int economies_per_block) { g_group = economies_per_block; }

void cats_debug_set_static(int on) { g_static = on ? 1 : 0; }
void cats_debug_set_warps(int warps_per_economy) { g_warps = warps_per_economy; }
void cats_debug_set_pipe(int mode) { g_pipe = mode; }
void cats_debug_set_host_zero_copy(int on) { g_zero_copy = on; }

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
        case 2: *byte_offset = d.idx < NK_HOT ? (uint64_t)L.o_kh + (uint64_t)d.idx * L.rowN
                                              : (uint64_t)L.o_kc + (uint64_t)(d.idx - NK_HOT) * L.rowN;
                *count = N; *dtype = 0; break;
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

int cats_field_stride(const char *name)
{
    if (!name) return 0;
    if (!strcmp(name, "PA") || !strcmp(name, "perm")) return 2;   // {PA, perm} records
    for (int i = 0; i < kNumFields; ++i) if (!strcmp(name, kFields[i].name)) return 1;
    return 0;
}

int cats_workset_field_info(int W, int F, int N, const char *name, uint64_t *byte_offset, int32_t *count,
                            int32_t *dtype)
{
    if (int rc = check_sizes(W, F, N)) return rc;
    if (!name || !byte_offset || !count || !dtype) return fail(CATS_EINVAL, "null argument%s", "");
    Layout L = make_layout(W, F, N);
    struct T { const char *n; uint64_t off; int cnt; int dt; } t[] = {
        {"c_Ld", (uint64_t)L.d_Ld, F, 1}, {"k_Ld", (uint64_t)L.d_Ld + 4ull * F, N, 1},
        {"c_vac", (uint64_t)L.d_vac, F, 1}, {"k_vac", (uint64_t)L.d_vac + 4ull * F, N, 1},
        {"c_Ftot", (uint64_t)L.d_cFtot, F, 0}, {"c_Kdem", (uint64_t)L.d_cKdem, F, 0},
        {"c_Ystock", (uint64_t)L.d_cYstock, F, 0}, {"c_valinv", (uint64_t)L.d_cvalinv, F, 0},
        {"k_Ftot", (uint64_t)L.d_kFtot, N, 0}, {"k_Ystock", (uint64_t)L.d_kYstock, N, 0},
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


// ---- launch configuration, cached: the attributes of a kernel have to be set once per (device, kernel, shared-memory
// size), the device's limits read once per device.  A period of a small batch is ~150 us: setting two function
// attributes and asking for the occupancy on every call (as round 1 did) is a visible part of it.
struct DevInfo { int known, sms, max_optin; };
static DevInfo g_dev[64];
struct CfgEntry { int dev; KernelFn fn; int smem; int occ; };
static CfgEntry g_cfg[128];
static int g_ncfg = 0;
static std::mutex g_cfg_mutex;

static int device_info(int *dev_out, DevInfo *out)
{
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return fail(CATS_ECUDA, "cudaGetDevice: %s", cudaGetErrorString(e));
    std::lock_guard<std::mutex> lock(g_cfg_mutex);
    if (dev < 0 || dev >= 64) return fail(CATS_ECUDA, "device ordinal out of range%s", "");
    if (!g_dev[dev].known) {
        cudaDeviceGetAttribute(&g_dev[dev].max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        cudaDeviceGetAttribute(&g_dev[dev].sms, cudaDevAttrMultiProcessorCount, dev);
        g_dev[dev].known = 1;
    }
    *dev_out = dev; *out = g_dev[dev];
    return CATS_OK;
}

// sets the kernel up for `smem` bytes of dynamic shared memory (once); *occ (nullable): resident blocks per SM
static int configure(KernelFn fn, int threads, int smem, int *occ_out, int *sms_out)
{
    int dev = 0; DevInfo di;
    if (int rc = device_info(&dev, &di)) return rc;
    if (sms_out) *sms_out = di.sms;
    {
        std::lock_guard<std::mutex> lock(g_cfg_mutex);
        for (int i = 0; i < g_ncfg; ++i)
            if (g_cfg[i].dev == dev && g_cfg[i].fn == fn && g_cfg[i].smem == smem) { if (occ_out) *occ_out = g_cfg[i].occ; return CATS_OK; }
    }
    if (smem > di.max_optin) return fail(CATS_ESMEM, "economy working set exceeds shared memory per block%s", "");
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return fail(CATS_ECUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    // the whole unified L1 as shared memory: residency (economies per SM) is what hides latency here,
    // and the streamed tail goes through L2 anyway
    e = cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return fail(CATS_ECUDA, "cudaFuncSetAttribute(carveout): %s", cudaGetErrorString(e));
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, threads, smem);
    if (e != cudaSuccess) return fail(CATS_ECUDA, "occupancy query: %s", cudaGetErrorString(e));
    if (occ < 1) return fail(CATS_ESMEM, "kernel cannot be resident with this working set%s", "");
    {
        // a kernel keeps ONE dynamic-shared-memory limit: entries of the same kernel with another size are stale now
        std::lock_guard<std::mutex> lock(g_cfg_mutex);
        int n = 0;
        for (int i = 0; i < g_ncfg; ++i) if (!(g_cfg[i].dev == dev && g_cfg[i].fn == fn)) g_cfg[n++] = g_cfg[i];
        g_ncfg = n;
        if (g_ncfg < 128) g_cfg[g_ncfg++] = CfgEntry{dev, fn, smem, occ};
    }
    if (occ_out) *occ_out = occ;
    return CATS_OK;
}

static int configure_smem(KernelFn fn, int threads, int smem) { return configure(fn, threads, smem, nullptr, nullptr); }

static int configure_kernel(const Layout &L, KernelFn fn, int g, int *blocks_per_sm, int *num_sms)
{
    int occ = 0, sms = 0;
    if (int rc = configure(fn, kThreads * g, slice_bytes(L) * g, &occ, &sms)) return rc;
    *blocks_per_sm = occ * g; *num_sms = sms;   // economies resident per SM
    return CATS_OK;
}

int cats_launch_info(int W, int F, int N, int32_t *threads, int32_t *smem_bytes, int32_t *blocks_per_sm,
                     int32_t *num_sms)
{
    if (int rc = check_sizes(W, F, N)) return rc;
    Layout L = make_layout(W, F, N);
    int occ = 0, sms = 0;
    int dev = 0, nsm = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    const int g = choose_group(L, 1ll << 40, nsm);   // the shape of a launch that fills the GPU
    if (int rc = configure_kernel(L, select_kernel(L, 5, g, false, true, true), g, &occ, &sms)) return rc;
    if (threads) *threads = kThreads * g;            // per block: g economies, one warp each
    if (smem_bytes) *smem_bytes = slice_bytes(L) * g;
    if (blocks_per_sm) *blocks_per_sm = occ;
    if (num_sms) *num_sms = sms;
    return CATS_OK;
}

int cats_launch_shape(int W, int F, int N, int64_t B, int32_t *warps_per_economy, int32_t *economies_per_block,
                      int32_t *pipelined_goods_market)
{
    if (int rc = check_sizes(W, F, N)) return rc;
    Layout L = make_layout(W, F, N);
    int dev = 0; DevInfo di;
    if (int rc = device_info(&dev, &di)) return rc;
    const int nw = choose_warps(L, B, di.sms);
    if (warps_per_economy) *warps_per_economy = nw;
    if (economies_per_block) *economies_per_block = nw > 1 ? 1 : choose_group(L, B, di.sms);
    if (pipelined_goods_market) *pipelined_goods_market = nw > 1 && choose_pipe(L, nw) ? 1 : 0;
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
                                                                   params_stride, nullptr);
    g_launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(CATS_ECUDA, "cats_init launch: %s", cudaGetErrorString(e));
    return CATS_OK;
}

int cats_reset_masked(void *d_blobs, int64_t B, int W, int F, int N, const double *d_params,
                      int64_t params_stride, const uint8_t *d_active, void *stream)
{
    if (int rc = check_sizes(W, F, N)) return rc;
    if (!d_blobs || !d_params || !d_active || B < 0) return fail(CATS_EINVAL, "cats_reset_masked: null pointer or negative B%s", "");
    if (params_stride != 0 && params_stride < CATS_N_PARAMS) return fail(CATS_EINVAL, "params_stride must be 0 or >= 34%s", "");
    if (B == 0) return CATS_OK;
    Layout L = make_layout(W, F, N);
    cats_init_kernel<<<(unsigned)B, 256, 0, (cudaStream_t)stream>>>((unsigned char *)d_blobs, B, L, d_params,
                                                                   params_stride, d_active);
    g_launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(CATS_ECUDA, "cats_reset_masked launch: %s", cudaGetErrorString(e));
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
    // the masked kernel is a production instantiation (Philox draws, whole periods, no debug image): a call that
    // combines d_active with a replay tape or a debug stop would silently get Philox-drawn whole periods
    if (s->d_active && (s->d_tape || s->stop_phase || s->d_debug))
        return fail(CATS_EINVAL, "d_active cannot be combined with d_tape, stop_phase or d_debug%s", "");
    if (s->d_active && (s->flags & CATS_FLAG_STRICT))
        return fail(CATS_EINVAL, "d_active cannot be combined with CATS_FLAG_STRICT%s", "");
    if (s->B == 0) return CATS_OK;

    KArgs a;
    fill_kargs(s, a);
    a.cycles = g_cycles.load();
    int occ = 0, sms = 0;
    int dev = 0, nsm = 148;
    { DevInfo di; if (int rc = device_info(&dev, &di)) return rc; nsm = di.sms; }
    a.slice = slice_bytes(a.L);
    a.bars = s->stop_phase ? 0 : 1;
    const bool default_z = s->z_all[0] == DefaultLayout::kZc && s->z_all[1] == DefaultLayout::kZk && s->z_all[2] == DefaultLayout::kZe;
    const bool production = !s->d_tape && !s->stop_phase && !s->d_debug;
    const bool strict = (s->flags & CATS_FLAG_STRICT) != 0;
    const int nw = (production && !strict && !s->d_active) ? choose_warps(a.L, s->B, nsm) : 1;
    if (nw > 1) {
        // one block = one economy, nw warps (cats_period_mw.cuh)
        KernelFn fn = kernel_mw((default_z && is_default_size(a.L)) ? 2 : 1, s->max_z, nw, choose_pipe(a.L, nw));
        const int smem = a.slice + mw_extra_bytes(nw, a.L.F);
        if (int rc = configure_smem(fn, kThreads * nw, smem)) return rc;
        fn<<<(unsigned)s->B, kThreads * nw, smem, (cudaStream_t)s->stream>>>(a);
    } else {
        const int g = choose_group(a.L, s->B, nsm);
        KernelFn fn = select_kernel(a.L, s->max_z, g, s->d_active != nullptr, production, default_z, strict);
        if (int rc = configure_kernel(a.L, fn, g, &occ, &sms)) return rc;
        fn<<<(unsigned)((s->B + g - 1) / g), kThreads * g, a.slice * g, (cudaStream_t)s->stream>>>(a);
    }
    g_launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(CATS_ECUDA, "cats_step launch: %s", cudaGetErrorString(e));
    return CATS_OK;
}

int cats_q_step(const cats_q_args *s)
{
    if (!s) return fail(CATS_EINVAL, "null args%s", "");
    if (!s->d_Q || !s->d_obs || !s->d_last || s->B < 0 || s->A < 1) return fail(CATS_EINVAL, "cats_q_step: null pointer, negative B or A < 1%s", "");
    for (int i = 0; i < 4; ++i) if (s->n[i] < 2 || s->n[i] > 4096) return fail(CATS_EINVAL, "cats_q_step: bins must be in [2, 4096]%s", "");
    if (s->do_train && (!s->d_reward || !s->d_terminated)) return fail(CATS_EINVAL, "cats_q_step: train needs reward and terminated%s", "");
    if (s->do_act && !s->d_actions) return fail(CATS_EINVAL, "cats_q_step: act needs an action buffer%s", "");
    if (s->B == 0 || (!s->do_train && !s->do_act)) return CATS_OK;
    QArgs q;
    q.Q = s->d_Q; q.B = s->B; q.A = s->A;
    for (int i = 0; i < 4; ++i) q.n[i] = s->n[i];
    for (int i = 0; i < 2; ++i) {
        q.obs_lo[i] = s->obs_lo[i]; q.obs_step[i] = (s->obs_hi[i] - s->obs_lo[i]) / (double)(s->n[i] - 1);
        q.act_lo[i] = s->act_lo[i]; q.act_hi[i] = s->act_hi[i];
    }
    q.alpha = s->alpha; q.gamma = s->gamma; q.epsilon = s->epsilon;
    for (uint32_t r = 0; r < 10; ++r) {
        q.rk[2 * r] = (uint32_t)s->seed + r * 0x9E3779B9u;
        q.rk[2 * r + 1] = (uint32_t)(s->seed >> 32) + r * 0xBB67AE85u;
    }
    q.economy_base = s->economy_base; q.step = s->step; q.do_train = s->do_train; q.do_act = s->do_act;
    q.obs = s->d_obs; q.reward = s->d_reward; q.terminated = s->d_terminated; q.last = s->d_last;
    q.actions = s->d_actions; q.active = s->d_active;
    const long long warps = s->B * s->A;
    cats_q_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, (cudaStream_t)s->stream>>>(q);
    g_launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(CATS_ECUDA, "cats_q_step launch: %s", cudaGetErrorString(e));
    return CATS_OK;
}

// sections of the staging block: actions | mask | obs | reward | terminated | status | stats, each 256-byte aligned
static void host_block_layout(int64_t B, int n_agents, int n_periods, size_t off[8])
{
    const size_t nA = (size_t)(n_agents > 0 ? n_agents : 0), b = (size_t)(B > 0 ? B : 0);
    const size_t np = (size_t)(n_periods > 0 ? n_periods : 1);
    const size_t sz[7] = {b * nA * 16, b * nA, b * nA * 16, b * nA * 8, b, b * 4, b * np * N_STATS * 8};
    size_t o = 0;
    for (int i = 0; i < 7; ++i) { off[i] = o; o += (sz[i] + 255) & ~(size_t)255; }
    off[7] = o;
}

size_t cats_host_scratch_bytes(int64_t B, int n_agents, int n_periods)
{
    size_t off[8];
    host_block_layout(B, n_agents, n_periods, off);
    return off[7];
}

int cats_host_block_offsets(int64_t B, int n_agents, int n_periods, uint64_t out7[7])
{
    if (!out7) return fail(CATS_EINVAL, "null argument%s", "");
    size_t off[8];
    host_block_layout(B, n_agents, n_periods, off);
    for (int i = 0; i < 7; ++i) out7[i] = off[i];
    return CATS_OK;
}

// The device alias of a HOST pointer the kernel can dereference: page-locked memory (cudaHostAlloc / cudaHostRegister,
// torch's pin_memory) is mapped into the device's address space under unified addressing.  nullptr: pageable memory.
static void *mapped_alias(const void *p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return at.type == cudaMemoryTypeHost ? at.devicePointer : nullptr;
}

int cats_step_host(const cats_step_args *h, void *d_scratch)
{
    if (!h || !d_scratch) return fail(CATS_EINVAL, "null args%s", "");
    cudaStream_t st = (cudaStream_t)h->stream;
    const size_t nA = (size_t)(h->n_agents > 0 ? h->n_agents : 0), b = (size_t)(h->B > 0 ? h->B : 0);
    const size_t np = (size_t)(h->n_periods > 0 ? h->n_periods : 1);
    size_t off[8];
    host_block_layout(h->B, h->n_agents, h->n_periods, off);
    unsigned char *base = (unsigned char *)d_scratch;
    cats_step_args s = *h;
    cudaError_t e = cudaSuccess;
    const bool up_act = h->d_actions && nA, up_msk = h->d_action_mask && nA;
    // inputs and outputs: {host pointer, section of the staging block, bytes}
    struct Io { const void *host; int sec; size_t bytes; };
    const Io ins[2] = {{up_act ? h->d_actions : nullptr, 0, b * nA * 16}, {up_msk ? h->d_action_mask : nullptr, 1, b * nA}};
    const Io outs[5] = {
        {(h->d_obs && nA) ? h->d_obs : nullptr, 2, b * nA * 16}, {(h->d_reward && nA) ? h->d_reward : nullptr, 3, b * nA * 8},
        {h->d_terminated, 4, b}, {h->d_status, 5, b * 4}, {h->d_stats, 6, b * np * N_STATS * 8}};
    // Zero copy: when EVERY output buffer is mapped page-locked memory the kernel writes it in place -- each economy
    // posts its outputs over PCIe as it finishes, while the others are still computing, and nothing is left to copy
    // when the kernel ends (measured, 16,384 economies: 1.445e7 -> 1.492e7 economy-periods/s end to end).  The
    // inputs are still copied ahead of the kernel: reading them in place puts a PCIe round trip on every economy's
    // chain (1.38e7).  Pageable buffers are staged through d_scratch with one copy per buffer or block;
    // cats_debug_set_host_zero_copy() forces either treatment for either direction.
    void *in_dev[2] = {nullptr, nullptr}, *out_dev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    const int zc_mode = g_zero_copy.load();
    bool zc_in = (zc_mode & 1) != 0, zc_out = (zc_mode & 2) != 0;
    for (int i = 0; i < 2 && zc_in; ++i)
        if (ins[i].host && !(in_dev[i] = mapped_alias(ins[i].host))) zc_in = false;
    for (int i = 0; i < 5 && zc_out; ++i)
        if (outs[i].host && !(out_dev[i] = mapped_alias(outs[i].host))) zc_out = false;
    if (zc_in) {
        if (up_act) s.d_actions = (const double *)in_dev[0];
        if (up_msk) s.d_action_mask = (const uint8_t *)in_dev[1];
    } else {
        // A caller that lays its host buffers out like the staging block (cats_host_block_offsets) gets one copy
        // in each direction instead of one per buffer.
        if (up_act && up_msk &&
            (const unsigned char *)h->d_action_mask - (const unsigned char *)h->d_actions == (ptrdiff_t)(off[1] - off[0])) {
            e = cudaMemcpyAsync(base + off[0], h->d_actions, (off[1] - off[0]) + b * nA, cudaMemcpyHostToDevice, st);
        } else {
            for (int i = 0; i < 2 && e == cudaSuccess; ++i)
                if (ins[i].host) e = cudaMemcpyAsync(base + off[ins[i].sec], ins[i].host, ins[i].bytes, cudaMemcpyHostToDevice, st);
        }
        if (e != cudaSuccess) return fail(CATS_ECUDA, "H2D copy: %s", cudaGetErrorString(e));
        if (up_act) s.d_actions = (const double *)(base + off[0]);
        if (up_msk) s.d_action_mask = base + off[1];
    }
    auto out_ptr = [&](int i) -> void * { return !outs[i].host ? nullptr : zc_out ? out_dev[i] : (void *)(base + off[outs[i].sec]); };
    s.d_obs = (double *)out_ptr(0);
    s.d_reward = (double *)out_ptr(1);
    s.d_terminated = (uint8_t *)out_ptr(2);
    s.d_status = (uint32_t *)out_ptr(3);
    s.d_stats = (double *)out_ptr(4);
    if (int rc = cats_step(&s)) return rc;
    if (!zc_out) {
        int first = -1, last = -1; bool mirrored = true;
        for (int i = 0; i < 5; ++i) {
            if (!outs[i].host) continue;
            if (first < 0) first = i;
            last = i;
            if ((const unsigned char *)outs[i].host - (const unsigned char *)outs[first].host != (ptrdiff_t)(off[outs[i].sec] - off[outs[first].sec]))
                mirrored = false;
        }
        if (first >= 0 && mirrored && first != last) {
            e = cudaMemcpyAsync((void *)outs[first].host, base + off[outs[first].sec],
                                (off[outs[last].sec] - off[outs[first].sec]) + outs[last].bytes, cudaMemcpyDeviceToHost, st);
        } else {
            for (int i = 0; i < 5 && e == cudaSuccess; ++i)
                if (outs[i].host) e = cudaMemcpyAsync((void *)outs[i].host, base + off[outs[i].sec], outs[i].bytes, cudaMemcpyDeviceToHost, st);
        }
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return fail(CATS_ECUDA, "D2H copy / sync: %s", cudaGetErrorString(e));
    return CATS_OK;
}

}  // extern "C"
