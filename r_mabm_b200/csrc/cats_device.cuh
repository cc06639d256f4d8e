// cats_device.cuh -- device-side primitives of the CATS period kernel:
// deterministic math, Philox-4x32-10, draw sources (Philox or replay tape), warp reductions,
// TMA bulk-copy / mbarrier wrappers.  Everything here implements docs/MODEL.md sections 3-4.
//
// Build note: this translation unit is compiled with -fmad=false.  The spec's arithmetic is
// "every * and + rounds separately", which is what makes the kernels bit-exact against the CPU
// oracle (gcc -ffp-contract=off).
//
// CATS_EMU: defined only by tests/simt (a CPU emulation of warps and blocks that compiles this source with
// g++ to check the kernel's logic without a GPU).  The inline PTX below then has plain-C++ stand-ins.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "cats_layout.h"

namespace cats {

constexpr unsigned FULL = 0xffffffffu;

// draw phases (Philox counter word 1 = phase << 8 | block) and market ids
enum { PH_PRICE_C = 1, PH_PRICE_K = 2, PH_INVEST = 3, PH_FIRE = 4, PH_JOB = 5, PH_CAP = 6,
       PH_GOODS = 7, PH_ORDER = 8 };
enum { MKT_JOB = 0, MKT_CAP = 1, MKT_GOODS = 2 };

// ---------------------------------------------------------------------------------------------
// deterministic exp / log  (docs/MODEL.md section 4)
__device__ __forceinline__ double det_exp(double x)
{
    if (x != x) return x;
    if (x > 709.0) return __longlong_as_double(0x7ff0000000000000LL);
    if (x < -708.0) return 0.0;
    const double LN2_HI = 6.93147180369123816490e-01;
    const double LN2_LO = 1.90821492927058770002e-10;
    const double INV_LN2 = 1.44269504088896338700e+00;
    double kf = floor(x * INV_LN2 + 0.5);
    double r = (x - kf * LN2_HI) - kf * LN2_LO;
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
    long long ki = (long long)kf;
    double scale = __longlong_as_double((ki + 1023) << 52);
    return p * scale;
}

__device__ __forceinline__ double det_log(double x)
{
    if (x != x) return x;
    if (x < 0.0) return __longlong_as_double(0x7ff8000000000000LL);
    if (x == 0.0) return __longlong_as_double(0xfff0000000000000LL);
    if (x == __longlong_as_double(0x7ff0000000000000LL)) return x;
    const double LN2_HI = 6.93147180369123816490e-01;
    const double LN2_LO = 1.90821492927058770002e-10;
    unsigned long long b = (unsigned long long)__double_as_longlong(x);
    long long e = (long long)((b >> 52) & 0x7ff);
    if (e == 0) {
        x = x * 18014398509481984.0;
        b = (unsigned long long)__double_as_longlong(x);
        e = (long long)((b >> 52) & 0x7ff) - 54;
    }
    e -= 1023;
    double m = __longlong_as_double((long long)((b & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL));
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

// ---------------------------------------------------------------------------------------------
// det_sum: lane l adds x[l], x[l+32], ... in order, then an xor butterfly 16,8,4,2,1.
// Must be called by all 32 lanes of a warp; every lane gets the result.
template <class Fn>
__device__ __forceinline__ double warp_det_sum(int n, int lane, Fn val)
{
    double a = 0.0;
    for (int i = lane; i < n; i += 32) a = a + val(i);
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) a = a + __shfl_xor_sync(FULL, a, off);
    return a;
}

template <class Fn>
__device__ __forceinline__ int warp_int_sum(int n, int lane, Fn val)
{
    int a = 0;
    for (int i = lane; i < n; i += 32) a += val(i);
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) a += __shfl_xor_sync(FULL, a, off);
    return a;
}

// ---------------------------------------------------------------------------------------------
// Philox-4x32-10
// rk: the 10 round keys {k0 + r * 0x9E3779B9, k1 + r * 0xBB67AE85}, expanded once on the host into the
// kernel's parameter block: a round is two wide multiplies and two three-input xors whose key operand
// comes straight from the constant bank (no key-schedule adds, no key registers).
__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               const uint32_t *rk)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ rk[2 * r], n2 = hi0 ^ c3 ^ rk[2 * r + 1];
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    }
    return make_uint4(c0, c1, c2, c3);
}

__device__ __forceinline__ double u01(uint32_t hi, uint32_t lo)
{
    unsigned long long m = ((unsigned long long)(hi >> 5) << 26) | (unsigned long long)(lo >> 6);
    return (double)m * (1.0 / 9007199254740992.0);
}

// the draw source of one economy-period
// TAPE = false: the instantiation never reads a tape (the `tape` member is dead and costs nothing)
template <bool TAPE>
struct DrawsT {
    static constexpr bool kTape = TAPE;
    const unsigned char *tape;  // nullptr => Philox
    const uint32_t *rk;         // Philox round keys (kernel parameter space)
    uint32_t period, economy;
    uint32_t epi;               // episode << 12: rides in counter word 1 above the phase and block bits
    const TapeOff *tp;          // tape record layout (kernel parameter space: costs no registers)
    const int *tape_z;          // z_c, z_k, z_e the tape was laid out with
    int *bad;                   // tape mode: status word that takes CATS_STATUS_TAPE_RANGE when a taped index is out of range

    __device__ __forceinline__ uint4 philox(uint32_t agent, uint32_t phase, uint32_t block) const
    {
        return philox4x32_10(agent, (phase << 8) | block | epi, period, economy, rk);
    }
    __device__ __forceinline__ double uniform(uint32_t phase, int agent) const
    {
        if (TAPE && tape) {
            uint32_t off = phase == PH_PRICE_C ? tp->u_price_c : phase == PH_PRICE_K ? tp->u_price_k : tp->u_invest;
            return *reinterpret_cast<const double *>(tape + off + 8u * (uint32_t)agent);
        }
        uint4 w = philox((uint32_t)agent, phase, 0);
        return u01(w.x, w.y);
    }
    __device__ __forceinline__ uint32_t fire_key(int worker) const
    {
        if (TAPE && tape) return *reinterpret_cast<const uint32_t *>(tape + tp->fire_key + 4u * (uint32_t)worker);
        return philox((uint32_t)worker, PH_FIRE, 0).x;
    }
    // z <= 8 distinct indices in [0,n): partial Fisher-Yates on a virtual identity array (register only)
    template <int ZM>
    __device__ __forceinline__ void candidates(uint32_t phase, int agent, int n, int z, int (&out)[ZM]) const
    {
        if (TAPE && tape) {
            uint32_t off; int zz;
            if (phase == PH_JOB) { off = tp->job_cand; zz = tape_z[2]; }
            else if (phase == PH_CAP) { off = tp->cap_cand; zz = tape_z[1]; }
            else { off = tp->goods_cand; zz = tape_z[0]; }
            const int *src = reinterpret_cast<const int *>(tape + off) + (size_t)agent * (size_t)zz;
            // a foreign tape is untrusted input: an index outside [0, n) would be a wild shared-memory access
            bool oob = false;
#pragma unroll
            for (int k = 0; k < ZM; ++k) {
                int v = k < z ? src[k] : 0;
                if ((unsigned)v >= (unsigned)n) { v = 0; oob = true; }
                out[k] = v;
            }
            if (oob) atomicOr(bad, 8);   // CATS_STATUS_TAPE_RANGE
            return;
        }
        // ONE Philox block: draw k < 4 uses word k, draw k >= 4 re-uses the low half of that product
        const uint4 a = philox((uint32_t)agent, phase, 0);
        uint32_t w[8];
        w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
#pragma unroll
        for (int k = 4; k < 8; ++k) w[k] = w[k - 4] * (uint32_t)(n - (k - 4));
        int jpos[ZM], jval[ZM];
#pragma unroll
        for (int k = 0; k < ZM; ++k) {
            out[k] = 0; jpos[k] = -1; jval[k] = 0;
            if (k < z) {
                int j = k + (int)__umulhi(w[k], (uint32_t)(n - k));
                int v = j, vk = k;
#pragma unroll
                for (int q = 0; q < k; ++q) {
                    if (jpos[q] == j) v = jval[q];
                    if (jpos[q] == k) vk = jval[q];
                }
                out[k] = v; jpos[k] = j; jval[k] = vk;
            }
        }
    }
};

// visiting order of a market: tape permutation, or the keyed bijection of docs/MODEL.md 3.3 (an
// alternating 4-round Feistel network on [0,a) x [0,b), a*b >= n, cycle-walked into [0,n)).
// The round keys live in shared memory (OrderKeys) and are re-read at every use: they are needed
// once per 32 agents, and registers are what bounds how many economies are resident per SM.
struct OrderDims { uint32_t a, b, m, n; };   // computed on the host: they depend on the market size only
struct OrderKeys {
    uint32_t k[4], a, b, m, n;
    const int *tape_order;
};
__host__ __device__ constexpr OrderDims make_order_dims(int n)
{
    OrderDims d{}; uint32_t a = 1;
    while ((uint64_t)a * a < (uint64_t)n) ++a;
    d.a = a; d.b = ((uint32_t)n + a - 1) / a; d.n = (uint32_t)n;
    d.m = d.b > 1 ? (uint32_t)(((uint64_t)1 << 32) / d.b) + 1u : 0u;   // mulhi32(x, m) == x / b for x < 2^23
    return d;
}
__device__ __forceinline__ uint32_t perm_hash(uint32_t v, uint32_t k)
{
    uint32_t x = v * 0x9E3779B1u + k;
    x ^= x >> 16;
    x *= 0x85EBCA6Bu;   // only the high bits of x are used (mulhi by the half-domain size)
    return x;
}
// NS > 0: the market size is a compile-time constant (default-size economies) and so are the network
// dimensions; NS == 0: they are read from the keys block.
template <uint32_t NS, bool TAPE>
struct OrderT {
    volatile OrderKeys *k;
    int *bad;   // tape mode: the status word (see DrawsT::bad)
    struct Dims { uint32_t a, b, m, n; };
    __device__ __forceinline__ Dims dims() const
    {
        if constexpr (NS > 0) {
            constexpr OrderDims d = make_order_dims((int)NS);
            return Dims{d.a, d.b, d.m, d.n};
        } else {
            return Dims{k->a, k->b, k->m, k->n};
        }
    }
    template <class D>
    __device__ __forceinline__ void init(const D &d, int market, const OrderDims &dims, OrderKeys *keys, int lane)
    {
        k = keys;
        bad = d.bad;
        if (lane == 0) {
            keys->a = dims.a; keys->b = dims.b; keys->m = dims.m; keys->n = dims.n;
            keys->tape_order = nullptr;
            if (TAPE && d.tape) {
                uint32_t off = market == MKT_JOB ? d.tp->job_order : market == MKT_CAP ? d.tp->cap_order : d.tp->goods_order;
                keys->tape_order = reinterpret_cast<const int *>(d.tape + off);
            } else {
                const uint4 w = d.philox((uint32_t)market, PH_ORDER, 0);
                keys->k[0] = w.x; keys->k[1] = w.y; keys->k[2] = w.z; keys->k[3] = w.w;
            }
        }
        __syncwarp();
    }
    __device__ __forceinline__ int at(int pos) const
    {
        if (TAPE) {
            const int *to = k->tape_order;
            if (to) {
                int v = to[pos];
                if ((unsigned)v >= k->n) { v = 0; atomicOr(bad, 8); }   // untrusted input: CATS_STATUS_TAPE_RANGE
                return v;
            }
        }
        const Dims dm = dims();
        const uint32_t a = dm.a, b = dm.b, m = dm.m, n = dm.n;
        const uint32_t k0 = k->k[0], k1 = k->k[1], k2 = k->k[2], k3 = k->k[3];
        uint32_t x = (uint32_t)pos;
        do {
            uint32_t L = b > 1 ? __umulhi(x, m) : x;
            uint32_t R = x - L * b;
            L += __umulhi(perm_hash(R, k0), a); if (L >= a) L -= a;
            R += __umulhi(perm_hash(L, k1), b); if (R >= b) R -= b;
            L += __umulhi(perm_hash(R, k2), a); if (L >= a) L -= a;
            R += __umulhi(perm_hash(L, k3), b); if (R >= b) R -= b;
            x = L * b + R;
        } while (x >= n);
        return (int)x;
    }
    // inverse of at(): the position at which agent `who` is visited (Feistel rounds undone in reverse
    // order, cycle-walked the same way).  Only for the keyed bijection, not for a tape order.
    __device__ __forceinline__ int position_of(int who) const
    {
        const Dims dm = dims();
        const uint32_t a = dm.a, b = dm.b, m = dm.m, n = dm.n;
        const uint32_t k0 = k->k[0], k1 = k->k[1], k2 = k->k[2], k3 = k->k[3];
        uint32_t x = (uint32_t)who;
        do {
            uint32_t L = b > 1 ? __umulhi(x, m) : x;
            uint32_t R = x - L * b;
            uint32_t t;
            t = __umulhi(perm_hash(L, k3), b); R = R >= t ? R - t : R + b - t;
            t = __umulhi(perm_hash(R, k2), a); L = L >= t ? L - t : L + a - t;
            t = __umulhi(perm_hash(L, k1), b); R = R >= t ? R - t : R + b - t;
            t = __umulhi(perm_hash(R, k0), a); L = L >= t ? L - t : L + a - t;
            x = L * b + R;
        } while (x >= n);
        return (int)x;
    }
    __device__ __forceinline__ bool from_tape() const { return TAPE && k->tape_order != nullptr; }
};

// stable ascending sort of ZM (candidate, price) pairs by price, register only: a sorting network
// on the total order (price, draw index) -- the same order a stable insertion sort produces.
// Slots k >= z must hold price = +inf (they stay behind the real candidates).
template <int ZM>
__device__ __forceinline__ void sort_by_price(int (&cand)[ZM], double (&pr)[ZM], int z)
{
    int idx[ZM];
#pragma unroll
    for (int k = 0; k < ZM; ++k) idx[k] = k;
    auto cx = [&](int i, int j) {
        const bool sw = pr[i] > pr[j] || (pr[i] == pr[j] && idx[i] > idx[j]);
        const double pi = pr[i], pj = pr[j];
        const int ci = cand[i], cj = cand[j], ii = idx[i], ij = idx[j];
        pr[i] = sw ? pj : pi; pr[j] = sw ? pi : pj;
        cand[i] = sw ? cj : ci; cand[j] = sw ? ci : cj;
        idx[i] = sw ? ij : ii; idx[j] = sw ? ii : ij;
    };
    (void)z;
    if (ZM == 2) { cx(0, 1); }
    else if (ZM == 5) {
        cx(0, 1); cx(3, 4); cx(2, 4); cx(2, 3); cx(1, 4); cx(0, 3); cx(0, 2); cx(1, 3); cx(1, 2);
    } else {
        static_assert(ZM == 2 || ZM == 5 || ZM == 8, "sorting networks exist for 2, 5 and 8 slots");
        cx(0, 1); cx(2, 3); cx(4, 5); cx(6, 7); cx(0, 2); cx(1, 3); cx(4, 6); cx(5, 7); cx(1, 2); cx(5, 6);
        cx(0, 4); cx(1, 5); cx(2, 6); cx(3, 7); cx(2, 4); cx(3, 5); cx(1, 2); cx(3, 4); cx(5, 6);
    }
}

// ascending sort of ZM packed 32-bit keys, register only (min / max sorting network)
template <int ZM>
__device__ __forceinline__ void sort_keys(unsigned (&key)[ZM])
{
    auto cx = [&](int i, int j) { const unsigned a = key[i], b = key[j]; key[i] = min(a, b); key[j] = max(a, b); };
    if (ZM == 2) { cx(0, 1); }
    else if (ZM == 5) {
        cx(0, 1); cx(3, 4); cx(2, 4); cx(2, 3); cx(1, 4); cx(0, 3); cx(0, 2); cx(1, 3); cx(1, 2);
    } else {
        static_assert(ZM == 2 || ZM == 5 || ZM == 8, "sorting networks exist for 2, 5 and 8 slots");
        cx(0, 1); cx(2, 3); cx(4, 5); cx(6, 7); cx(0, 2); cx(1, 3); cx(4, 6); cx(5, 7); cx(1, 2); cx(5, 6);
        cx(0, 4); cx(1, 5); cx(2, 6); cx(3, 7); cx(2, 4); cx(3, 5); cx(1, 2); cx(3, 4); cx(5, 6);
    }
}

// demand registered at a consumption-goods seller: 2^-40 fixed point, wrap-around add (MODEL.md 4.4)
__device__ __forceinline__ unsigned long long fx_from(double x)
{
    if (!(x < 4194304.0)) return 1ull << 62;
    return (unsigned long long)__double2ll_rn(x * 1099511627776.0);
}
__device__ __forceinline__ double fx_to(unsigned long long a)
{
    return (double)(long long)a * (1.0 / 1099511627776.0);
}
// 64-bit integer add on a shared-memory accumulator from two native 32-bit atomics (a 64-bit
// shared atomicAdd compiles to a compare-and-swap spin loop).  Integer adds commute, and every
// carry out of the low word is added to the high word exactly once, so the final value is exact.
__device__ __forceinline__ void fx_add(unsigned long long *acc, unsigned long long x)
{
    unsigned *w = reinterpret_cast<unsigned *>(acc);
    const unsigned lo = (unsigned)x;
    unsigned hi = (unsigned)(x >> 32);
    const unsigned old = atomicAdd(&w[0], lo);
    hi += (old + lo < old) ? 1u : 0u;
    if (hi) atomicAdd(&w[1], hi);
}

// the same add on an accumulator in global memory: one fire-and-forget 64-bit reduction at L2
__device__ __forceinline__ void fx_red(unsigned long long *acc, unsigned long long x)
{
#ifdef CATS_EMU
    *acc += x;
#else
    asm volatile("red.global.add.u64 [%0], %1;" ::"l"(acc), "l"(x) : "memory");
#endif
}

// a / b, bit-identical to the IEEE quotient.  The compiler's fp64 division drops into a ~60
// instruction subroutine for the WHOLE warp as soon as one lane has a zero (or denormal) dividend --
// and zero dividends are common here (idle lanes, households without wealth, firms without a
// dividend).  +-0 / (positive finite) is +-0, so those lanes divide 1.0 instead and keep their zero.
__device__ __forceinline__ double div_nz(double a, double b)
{
    const bool z = (a == 0.0) && (b > 0.0) && (b < __longlong_as_double(0x7ff0000000000000LL));
    double num = z ? 1.0 : a;
#ifndef CATS_EMU
    asm volatile("" : "+d"(num));   // keep the compiler from folding this back into a / b
#endif
    const double q = num / b;
    return z ? a : q;
}

__device__ __forceinline__ int d2count(double x)
{
    if (!(x > 0.0)) return 0;
    if (x > 1.0e9) return 1000000000;
    return (int)x;
}

// ---------------------------------------------------------------------------------------------
// TMA bulk copy (cp.async.bulk, SASS UBLKCP) + mbarrier
#ifdef CATS_EMU
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) { simt::mbar_init(bar, count); }
__device__ __forceinline__ void mbar_inval(uint64_t *) {}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) { simt::mbar_arrive(bar); }
__device__ __forceinline__ void fence_mbar_init() {}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t) { simt::mbar_arrive(bar); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) { simt::mbar_wait(bar, parity); }
__device__ __forceinline__ void mbar_wait_backoff(uint64_t *bar, uint32_t parity) { simt::mbar_wait(bar, parity); }
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *) { memcpy(smem_dst, gsrc, bytes); }
__device__ __forceinline__ void bulk_s2g(void *gdst, const void *smem_src, uint32_t bytes) { memcpy(gdst, smem_src, bytes); }
__device__ __forceinline__ void bulk_prefetch_l2(const void *, uint32_t) {}
__device__ __forceinline__ void bulk_commit() {}
__device__ __forceinline__ void bulk_wait_all() {}
__device__ __forceinline__ void fence_proxy_async() {}
#else
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// an mbarrier object has to be invalidated before its memory is used for anything else (or initialised again)
__device__ __forceinline__ void mbar_inval(uint64_t *bar)
{
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// has the phase with this parity completed?  (no labels in the PTX: may be inlined any number of times)
__device__ __forceinline__ bool mbar_test(uint64_t *bar, uint32_t parity)
{
    uint32_t done;
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return done != 0u;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    while (!mbar_test(bar, parity)) { }
}
// producer-side wait: poll with a sleep so that the spinning warp does not steal issue slots
__device__ __forceinline__ void mbar_wait_backoff(uint64_t *bar, uint32_t parity)
{
    uint32_t done = 0;
    while (true) {
        asm volatile(
            "{\n"
            ".reg .pred P1;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
            "selp.u32 %0, 1, 0, P1;\n"
            "}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (done) break;
#ifndef CATS_PRODUCER_SLEEP_NS
#define CATS_PRODUCER_SLEEP_NS 256
#endif
        __nanosleep(CATS_PRODUCER_SLEEP_NS);
    }
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *gdst, const void *smem_src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
// L2 prefetch of a contiguous region (16-byte granularity): one instruction, no destination, no wait
__device__ __forceinline__ void bulk_prefetch_l2(const void *gsrc, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until the committed bulk stores have READ their shared-memory source (the block may then
// exit and give the memory back; the writes themselves complete with the grid)
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

#endif  // CATS_EMU

}  // namespace cats
