// proto_tpe.cu -- feasibility probe (development aid, not product code): the consumption-goods
// market with ONE THREAD PER ECONOMY (a warp = 32 economies, seller tables interleaved by lane in
// shared memory), to be timed against the warp-per-economy market of cats_kernels.cu.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../r_mabm_b200/csrc/cats_device.cuh"
#include "../r_mabm_b200/csrc/cats_layout.h"

using namespace cats;

struct PArgs {
    Layout L;
    OrderDims od;
    unsigned char *blobs;
    long long B;
    const double *params;
    uint32_t k0, k1;
};

__device__ __forceinline__ int order_at_reg(uint32_t pos, const OrderDims &d, uint32_t k0, uint32_t k1, uint32_t k2, uint32_t k3)
{
    uint32_t x = pos;
    do {
        uint32_t Lh = d.b > 1 ? __umulhi(x, d.m) : x;
        uint32_t R = x - Lh * d.b;
        Lh += __umulhi(perm_hash(R, k0), d.a); if (Lh >= d.a) Lh -= d.a;
        R += __umulhi(perm_hash(Lh, k1), d.b); if (R >= d.b) R -= d.b;
        Lh += __umulhi(perm_hash(R, k2), d.a); if (Lh >= d.a) Lh -= d.a;
        R += __umulhi(perm_hash(Lh, k3), d.b); if (R >= d.b) R -= d.b;
        x = Lh * d.b + R;
    } while (x >= d.n);
    return (int)x;
}

template <int ZM, int ILV>
__global__ void __launch_bounds__(32) proto_goods(const __grid_constant__ PArgs a)
{
    extern __shared__ __align__(16) unsigned char sm[];
    const Layout &L = a.L;
    const int lane = threadIdx.x;
    const long long b = (long long)blockIdx.x * 32 + lane;
    if (b >= a.B) return;
    const int W = L.W, F = L.F, H = L.H;
    unsigned char *blob = a.blobs + (size_t)b * (size_t)L.blob;
    double *ys = reinterpret_cast<double *>(sm);                       // [F][32]
    unsigned short *cls = reinterpret_cast<unsigned short *>(ys + (size_t)F * 32);   // [F][32]
    const double *scal = reinterpret_cast<const double *>(blob + L.o_scal);
    double *cP = reinterpret_cast<double *>(blob + L.o_c + C_P * L.rowF);
    double *cY = reinterpret_cast<double *>(blob + L.o_c + C_Y_prev * L.rowF);
    double *cYd = reinterpret_cast<double *>(blob + L.o_c + C_Yd * L.rowF);
    double *cQ = reinterpret_cast<double *>(blob + L.o_c + C_Q * L.rowF);
    double *cRP = reinterpret_cast<double *>(blob + L.o_c + C_investment * L.rowF);   // probe: 1/P parked in a dead row
    const double *cdiv = reinterpret_cast<const double *>(blob + L.o_c + C_div_income * L.rowF);
    const double *kdiv = reinterpret_cast<const double *>(blob + L.o_kc + (K_div_income - NK_HOT) * L.rowN);
    double *PA = reinterpret_cast<double *>(blob + L.o_PA), *perm = reinterpret_cast<double *>(blob + L.o_perm);
    const int *Oc = reinterpret_cast<const int *>(blob + L.o_Oc);
    unsigned long long *acc = reinterpret_cast<unsigned long long *>(cYd);

    // open the market: stock, reciprocal price, price class
    for (int f = 0; f < F; ++f) {
        const double p = cP[f];
        ys[f * 32 + lane] = cY[f];
        cRP[f] = 1.0 / p;
        cYd[f] = 0.0;
        int n = 0;
        for (int g = 0; g < F; ++g) n += cP[g] < p ? 1 : 0;
        cls[f * 32 + lane] = (unsigned short)n;
    }
    const double price = scal[S_price], wb = scal[S_wb];
    const double rpavg = 1.0 / price;
    const double net = wb * (1.0 - a.params[P_tax_rate]), sub = a.params[P_subsidy] * wb;
    const double ir_emp = net * rpavg, ir_un = sub * rpavg;
    const double xi = a.params[P_xi], omxi = 1.0 - xi, chi = a.params[P_chi];
    const int z = ZM < F ? ZM : F;
    Draws d; d.tape = nullptr; d.k0 = a.k0; d.k1 = a.k1; d.period = (uint32_t)scal[S_timestep]; d.economy = (uint32_t)b;
    d.tp = nullptr; d.tape_z = nullptr;
    const uint4 ok = d.philox(MKT_GOODS, PH_ORDER, 0);

    for (int pos0 = 0; pos0 < H; pos0 += ILV) {
        int hh[ILV]; double bb[ILV], pav[ILV]; unsigned key[ILV][ZM]; double rpc[ILV][ZM];
        // ---- prepare ILV households (independent work: instruction-level parallelism)
#pragma unroll
        for (int u = 0; u < ILV; ++u) {
            const int pos = pos0 + u;
            hh[u] = -1; bb[u] = 0.0; pav[u] = 0.0;
            if (pos < H) {
                const int h = order_at_reg((uint32_t)pos, a.od, ok.x, ok.y, ok.z, ok.w);
                hh[u] = h;
                double pa = PA[h], pm = perm[h], ir;
                if (h < W) { const bool emp = Oc[h] >= 0; pa = pa + (emp ? net : sub); ir = emp ? ir_emp : ir_un; }
                else ir = (h < W + F ? cdiv[h - W] : kdiv[h - W - F]) * rpavg;
                pm = pm * xi + omxi * ir;
                const double target = pm + (chi * pa) * rpavg;
                double bud = target * price;
                if (bud > pa) bud = pa;
                if (!(bud > 0.0)) bud = 0.0;
                pa = pa - bud;
                perm[h] = pm;
                bb[u] = bud; pav[u] = pa;
                int cand[ZM];
                d.candidates<ZM>(PH_GOODS, h, F, z, cand);
#pragma unroll
                for (int k = 0; k < ZM; ++k)
                    key[u][k] = k < z ? ((unsigned)cls[cand[k] * 32 + lane] << 17) | ((unsigned)k << 14) | (unsigned)cand[k] : 0xffffffffu;
                sort_keys<ZM>(key[u]);
#pragma unroll
                for (int k = 0; k < ZM; ++k) rpc[u][k] = k < z ? cRP[key[u][k] & 0x3fffu] : 0.0;
            }
        }
        // ---- commit, in visiting order
#pragma unroll
        for (int u = 0; u < ILV; ++u) {
            if (hh[u] < 0) continue;
            double bud = bb[u];
            if (bud > 0.0) {
                unsigned long long xb = fx_from(bud * rpavg);
#pragma unroll
                for (int k = 0; k < ZM; ++k) {
                    if (k < z && bud > 0.0) {
                        const int cnd = (int)(key[u][k] & 0x3fffu);
                        fx_red(&acc[cnd], xb);
                        const double y = ys[cnd * 32 + lane];
                        if (y > 0.0) {
                            const double dq = bud * rpc[u][k];
                            if (y > dq) { ys[cnd * 32 + lane] = y - dq; bud = 0.0; }
                            else { bud = bud - y * cP[cnd]; ys[cnd * 32 + lane] = 0.0; xb = fx_from(bud * rpavg); }
                        }
                    }
                }
            }
            PA[hh[u]] = pav[u] + bud;
        }
    }
    __threadfence();
    for (int f = 0; f < F; ++f) {
        const unsigned long long v = __ldcg(&acc[f]);
        cQ[f] = cY[f] - ys[f * 32 + lane];
        cYd[f] = (fx_to(v) * price) * cRP[f];
    }
}

extern "C" int proto_goods_launch(void *blobs, long long B, int W, int F, int N, const double *params,
                                  unsigned long long seed, int ilv, void *stream)
{
    PArgs a;
    a.L = make_layout(W, F, N);
    a.od = make_order_dims(W + F + N);
    a.blobs = (unsigned char *)blobs; a.B = B; a.params = params;
    a.k0 = (uint32_t)seed; a.k1 = (uint32_t)(seed >> 32);
    const int smem = F * 32 * 8 + F * 32 * 2;
    void (*fn)(const PArgs) = ilv == 1 ? proto_goods<5, 1> : ilv == 2 ? proto_goods<5, 2> : proto_goods<5, 4>;
    cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, 32, smem);
    fn<<<(unsigned)((B + 31) / 32), 32, smem, (cudaStream_t)stream>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { fprintf(stderr, "proto launch: %s\n", cudaGetErrorString(e)); return -1; }
    return occ;
}
