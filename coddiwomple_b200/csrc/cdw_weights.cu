// cdw_weights.cu — K2 (incremental work / log-weight update, log-sum-exp + ESS reduction), K3a (integer cdf scan +
// vectorised search for multinomial / systematic resampling), K3b (coalesced particle gather) for sm_100a.
//
// Reference semantics: TargetFactory.compute_incremental_work (coddiwomple/distribution_factories.py:124-129),
// nESS / vanilla_floor_threshold (coddiwomple/utils.py:41-72), Resampler.resample (coddiwomple/resamplers.py:61-133),
// MultinomialResampler._resample (coddiwomple/resamplers.py:254-272).
//
// Exactness: ancestors come from an INTEGER cdf (weights quantised with det_exp + rint, uint64 prefix sum, integer
// search), so they are independent of summation order, block schedule and GPU count, and equal oracle/resampling.py
// bit for bit.
#include <stdlib.h>

#include "cdw_common.h"
#include "cdw_resample.cuh"

namespace cdw {

constexpr int kBlock = 256;
constexpr int kItems = 8;
constexpr int kTile = kBlock * kItems;
constexpr int kMaxStatBlocks = 2048;

// workspace header (uint64 words)
enum { WS_TICKET = 0, WS_T1 = 1, WS_SHIFT = 2, WS_TOTAL = 3, WS_HEADER_WORDS = 8 };

struct WsView {
    unsigned long long* hdr;
    double* part_e;
    double* part_e2;
    unsigned long long* tile_sum;
    unsigned long long* tile_off;
    unsigned long long* coarse;  // cdf sampled at the end of every group of kGroup entries: coarse[g] = cdf[min(n, kGroup (g + 1)) - 1]
};

constexpr int kGroup = 32, kGroupLog2 = 5;  // coarse table = n/4 bytes: L2-resident up to ~4e8 particles
__host__ __device__ inline long long n_groups_for(long long n) { return (n + kGroup - 1) / kGroup; }

__host__ __device__ inline long long n_tiles_for(long long n) { return (n + kTile - 1) / kTile; }

__host__ __device__ inline WsView ws_view(void* ws, long long n) {
    WsView v;
    unsigned char* b = (unsigned char*)ws;
    v.hdr = (unsigned long long*)b; b += sizeof(unsigned long long) * WS_HEADER_WORDS;
    v.part_e = (double*)b; b += sizeof(double) * kMaxStatBlocks;
    v.part_e2 = (double*)b; b += sizeof(double) * kMaxStatBlocks;
    v.tile_sum = (unsigned long long*)b; b += sizeof(unsigned long long) * n_tiles_for(n);
    v.tile_off = (unsigned long long*)b; b += sizeof(unsigned long long) * (n_tiles_for(n) + 1);
    b = (unsigned char*)(((size_t)b + 15) & ~(size_t)15);  // the coarse table is written with 16-byte stores
    v.coarse = (unsigned long long*)b;
    return v;
}


template <typename T, typename Op>
__device__ __forceinline__ T block_reduce(T v, Op op, T* sh /*[kBlock/32]*/) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    T r = sh[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) r = op(r, sh[w]);
    return r;  // same value in every thread, combined in a fixed order
}

// ------------------------------------------------------------------------------------------------ K2 pass 1
__global__ void __launch_bounds__(kBlock) weight_update_kernel(const double* __restrict__ u_new, const double* __restrict__ aux,
                                                               const double* __restrict__ proposal, const double* __restrict__ cum,
                                                               long long n, double* __restrict__ inc_out, double* __restrict__ w_out,
                                                               unsigned long long* wmin_key) {
    __shared__ unsigned long long sh[kBlock / 32];
    unsigned long long m = ~0ull;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double inc = u_new[i] + (aux ? aux[i] : 0.0);
        if (proposal) inc += proposal[i];
        double w = (cum ? cum[i] : 0.0) + inc;
        if (inc_out) inc_out[i] = inc;
        w_out[i] = w;
        unsigned long long k = ordered_key(w);
        m = k < m ? k : m;
    }
    m = block_reduce(m, [](unsigned long long a, unsigned long long b) { return a < b ? a : b; }, sh);
    if (threadIdx.x == 0 && wmin_key && m != ~0ull) atomicMin(wmin_key, m);
}

// ------------------------------------------------------------------------------------------------ K2 pass 2
__global__ void __launch_bounds__(kBlock) weight_stats_kernel(const double* __restrict__ w, long long n, double threshold,
                                                              const unsigned long long* __restrict__ wmin_key, double* stats, WsView ws,
                                                              int s1) {
    __shared__ double shd[kBlock / 32];
    __shared__ unsigned long long shu[kBlock / 32];
    __shared__ bool is_last;
    const double wmin = ordered_key_inverse(*wmin_key);
    const double scale1 = pow2(s1);
    double se = 0.0, se2 = 0.0;
    unsigned long long t1 = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double e = det_exp(CDW_ADD(wmin, -w[i]));
        se += e;
        se2 += e * e;
        t1 += __double2ull_rn(CDW_MUL(e, scale1));
    }
    se = block_reduce(se, [](double a, double b) { return a + b; }, shd);
    se2 = block_reduce(se2, [](double a, double b) { return a + b; }, shd);
    t1 = block_reduce(t1, [](unsigned long long a, unsigned long long b) { return a + b; }, shu);
    if (threadIdx.x == 0) {
        ws.part_e[blockIdx.x] = se;
        ws.part_e2[blockIdx.x] = se2;
        atomicAdd(&ws.hdr[WS_T1], t1);
        __threadfence();
        unsigned long long ticket = atomicAdd(&ws.hdr[WS_TICKET], 1ull);
        is_last = ticket == (unsigned long long)gridDim.x - 1;
    }
    __syncthreads();
    if (is_last && threadIdx.x == 0) {
        __threadfence();
        double sum_e = 0.0, sum_e2 = 0.0;
        for (unsigned b = 0; b < gridDim.x; ++b) {  // fixed order: run-to-run deterministic
            sum_e += ((volatile double*)ws.part_e)[b];
            sum_e2 += ((volatile double*)ws.part_e2)[b];
        }
        unsigned long long T1 = ((volatile unsigned long long*)ws.hdr)[WS_T1];
        ws.hdr[WS_T1] = 0; ws.hdr[WS_TICKET] = 0;  // leave the accumulators at rest: the next call needs no memset
        int bitlen = 64 - __clzll((long long)T1);
        int shift = 60 - bitlen;
        shift = shift < 0 ? 0 : shift;
        ws.hdr[WS_SHIFT] = (unsigned long long)(s1 + shift);
        double ness = (sum_e * sum_e) / ((double)n * sum_e2);
        stats[CDW_STAT_WMIN] = wmin;
        stats[CDW_STAT_SUM_E] = sum_e;
        stats[CDW_STAT_SUM_E2] = sum_e2;
        stats[CDW_STAT_NESS] = ness;
        stats[CDW_STAT_MEAN] = wmin - log(sum_e) + log((double)n);
        stats[CDW_STAT_RESAMPLE] = (threshold < 0.0 || ness <= threshold) ? 1.0 : 0.0;
        stats[CDW_STAT_TOTAL] = 0.0;
        stats[CDW_STAT_NNAN] = 0.0;
    }
}

// ------------------------------------------------------------------------------------------------ K3a: integer cdf
__device__ __forceinline__ void load_tile_q(const double* __restrict__ w, long long n, long long base, double wmin, double scale,
                                            unsigned long long (&q)[kItems]) {
    const long long i0 = base + (long long)threadIdx.x * kItems;
    if (i0 + kItems <= n) {
        const double2* p = (const double2*)(w + i0);  // i0 is a multiple of 8 -> 64-byte aligned
#pragma unroll
        for (int k = 0; k < kItems / 2; ++k) {
            double2 v = __ldg(p + k);
            q[2 * k] = quantise(v.x, wmin, scale);
            q[2 * k + 1] = quantise(v.y, wmin, scale);
        }
    } else {
#pragma unroll
        for (int k = 0; k < kItems; ++k) q[k] = (i0 + k < n) ? quantise(w[i0 + k], wmin, scale) : 0ull;
    }
}

__global__ void __launch_bounds__(kBlock) cdf_tile_sums_kernel(const double* __restrict__ w, long long n, const double* __restrict__ stats,
                                                               WsView ws) {
    if (stats[CDW_STAT_RESAMPLE] == 0.0) return;
    __shared__ unsigned long long sh[kBlock / 32];
    const double wmin = stats[CDW_STAT_WMIN];
    const double scale = pow2((int)ws.hdr[WS_SHIFT]);
    for (long long tile = blockIdx.x; tile < n_tiles_for(n); tile += gridDim.x) {
        unsigned long long q[kItems];
        load_tile_q(w, n, tile * kTile, wmin, scale, q);
        unsigned long long s = 0;
#pragma unroll
        for (int k = 0; k < kItems; ++k) s += q[k];
        s = block_reduce(s, [](unsigned long long a, unsigned long long b) { return a + b; }, sh);
        if (threadIdx.x == 0) ws.tile_sum[tile] = s;
    }
}

__global__ void __launch_bounds__(1024) cdf_scan_tiles_kernel(long long n, double* stats, WsView ws) {
    if (stats[CDW_STAT_RESAMPLE] == 0.0) return;
    __shared__ unsigned long long warp_tot[32];
    __shared__ unsigned long long carry_sh;
    const long long nt = n_tiles_for(n);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_sh = 0;
    __syncthreads();
    for (long long base = 0; base < nt; base += 1024) {
        long long i = base + threadIdx.x;
        unsigned long long v = i < nt ? ws.tile_sum[i] : 0ull;
        unsigned long long incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            unsigned long long t = warp_tot[lane], s = t;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                unsigned long long u = __shfl_up_sync(0xffffffffu, s, o);
                if (lane >= o) s += u;
            }
            warp_tot[lane] = s - t;  // exclusive prefix of warp totals
        }
        __syncthreads();
        const unsigned long long carry = carry_sh;
        const unsigned long long excl = carry + warp_tot[warp] + incl - v;
        if (i < nt) ws.tile_off[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry_sh = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        ws.hdr[WS_TOTAL] = carry_sh;
        stats[CDW_STAT_TOTAL] = (double)carry_sh;
    }
}

// cdw_smc_resample draws its own uniforms (Philox, keyed by slot) inside the search kernels and lets them re-arm the
// running-min key: no separate uniforms / fill launches.
struct SearchGen {
    int gen;                              // 0: read `uniforms`; 1: generate (and store to uniforms_out if not NULL)
    unsigned long long seed, stream_id;
    double* uniforms_out;
    unsigned long long* rearm_key;        // set to ~0 by the kernel (may be NULL)
};

__device__ __forceinline__ void search_rearm(const SearchGen& g) {
    if (g.rearm_key && blockIdx.x == 0 && threadIdx.x == 0) *g.rearm_key = ~0ull;
}

// Systematic resampling needs no search at all: its targets are sorted, so the tile scan that has cdf[j - 1] and cdf[j] in
// registers can emit item j's slots directly, [first(cdf[j - 1]), first(cdf[j])) with first(c) = min{i : target_i >= c},
// target_i = floor(((i + u0) / n) T).  first() is estimated in floating point and corrected with the exact forward map,
// so the ancestors are bit-identical to the search form.  No second pass over the cdf.
struct EmitArgs {
    const double* uniforms;   // [1] u0 when gen.gen == 0
    const double* cum_in; double* cum_out; int* anc;
    SearchGen gen;
};

__device__ __forceinline__ unsigned long long systematic_target(long long i, double u0, long long n, unsigned long long total) {
    return fixed_target(__ddiv_rn(__dadd_rn((double)i, u0), (double)n), total);
}

__device__ __forceinline__ long long first_slot(unsigned long long c, unsigned long long total, double u0, long long n, double n_over_total) {
    if (c == 0) return 0;
    long long i = (long long)ceil((double)c * n_over_total - u0);
    i = i < 0 ? 0 : (i > n ? n : i);
    while (i > 0 && systematic_target(i - 1, u0, n, total) >= c) --i;
    while (i < n && systematic_target(i, u0, n, total) < c) ++i;
    return i;
}

__device__ __forceinline__ void emit_slot(const EmitArgs& e, long long i, long long j, double mean) {
    e.anc[i] = (int)j;
    if (e.cum_out) {
        const double c = e.cum_in ? e.cum_in[i] : 0.0;
        e.cum_out[i] = CDW_ADD(c, CDW_ADD(mean, -c));  // _update_particle: cum += (mean - cum) (resamplers.py:196-197)
    }
}

constexpr int kHeavyRange = 16;   // longer slot ranges (one ancestor copied many times) are written by the whole block
constexpr int kHeavyList = 64;

template <bool EMIT>
__global__ void __launch_bounds__(kBlock) cdf_tile_scan_kernel(const double* __restrict__ w, long long n, const double* __restrict__ stats,
                                                               WsView ws, unsigned long long* __restrict__ cdf, EmitArgs em) {
    if (EMIT) search_rearm(em.gen);
    if (stats[CDW_STAT_RESAMPLE] == 0.0) {
        if (EMIT) {
            for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
                em.anc[i] = (int)i;
                if (em.cum_out) em.cum_out[i] = w[i];  // null update: cum += inc (resamplers.py:136-156)
            }
        }
        return;
    }
    __shared__ unsigned long long warp_tot[kBlock / 32];
    __shared__ long long heavy[3 * kHeavyList];
    __shared__ int n_heavy;
    const unsigned long long total = EMIT ? ws.hdr[WS_TOTAL] : 0ull;
    const double mean = stats[CDW_STAT_MEAN];
    const double u0 = EMIT ? (em.gen.gen ? philox_uniform53(em.gen.seed, em.gen.stream_id, 0) : em.uniforms[0]) : 0.0;
    const double n_over_total = EMIT ? (double)n / (double)total : 0.0;
    if (EMIT && em.gen.gen && em.gen.uniforms_out && blockIdx.x == 0 && threadIdx.x == 0) em.gen.uniforms_out[0] = u0;
    if (EMIT && threadIdx.x == 0) n_heavy = 0;
    const double wmin = stats[CDW_STAT_WMIN];
    const double scale = pow2((int)ws.hdr[WS_SHIFT]);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long long tile = blockIdx.x; tile < n_tiles_for(n); tile += gridDim.x) {
        unsigned long long q[kItems];
        load_tile_q(w, n, tile * kTile, wmin, scale, q);
#pragma unroll
        for (int k = 1; k < kItems; ++k) q[k] += q[k - 1];
        const unsigned long long tot = q[kItems - 1];
        unsigned long long incl = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        __syncthreads();
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        unsigned long long off = ws.tile_off[tile] + incl - tot;
        for (int wi = 0; wi < warp; ++wi) off += warp_tot[wi];
        const long long i0 = tile * kTile + (long long)threadIdx.x * kItems;
        if (i0 + kItems <= n) {
            ulonglong2* o2 = (ulonglong2*)(cdf + i0);
#pragma unroll
            for (int k = 0; k < kItems / 2; ++k) o2[k] = make_ulonglong2(q[2 * k] + off, q[2 * k + 1] + off);
        } else {
#pragma unroll
            for (int k = 0; k < kItems; ++k)
                if (i0 + k < n) cdf[i0 + k] = q[k] + off;
        }
        // coarse table for the hierarchical search: the last cdf entry of every group of kGroup.  A group ends on the last
        // item of a thread (kGroup is a multiple of kItems) or at n - 1.  The index uses an unsigned shift on purpose:
        // nvcc 12.9 dropped the rounding of a signed `i / kGroup` in the n - 1 case (uniform datapath), giving a
        // misaligned store for n = 2..7.
        static_assert((1 << kGroupLog2) == kGroup && kGroup % kItems == 0, "groups are whole threads");
#pragma unroll
        for (int k = 0; k < kItems; ++k) {
            const unsigned long long i = (unsigned long long)(i0 + k);
            const bool ends = (k == kItems - 1 && ((i + 1) & (kGroup - 1)) == 0) || i + 1 == (unsigned long long)n;
            if (i < (unsigned long long)n && ends) ws.coarse[i >> kGroupLog2] = q[k] + off;
        }
        if (EMIT) {
            long long lo = first_slot(off, total, u0, n, n_over_total);
#pragma unroll
            for (int k = 0; k < kItems; ++k) {
                const long long j = i0 + k;
                const long long hi = j < n ? first_slot(q[k] + off, total, u0, n, n_over_total) : lo;
                if (hi - lo > kHeavyRange) {
                    const int slot = atomicAdd(&n_heavy, 1);
                    if (slot < kHeavyList) { heavy[3 * slot] = j; heavy[3 * slot + 1] = lo; heavy[3 * slot + 2] = hi; }
                    else for (long long i = lo; i < hi; ++i) emit_slot(em, i, j, mean);
                } else {
                    for (long long i = lo; i < hi; ++i) emit_slot(em, i, j, mean);
                }
                lo = hi;
            }
            __syncthreads();
            const int nh = n_heavy < kHeavyList ? n_heavy : kHeavyList;
            for (int h = 0; h < nh; ++h)
                for (long long i = heavy[3 * h + 1] + threadIdx.x; i < heavy[3 * h + 2]; i += blockDim.x) emit_slot(em, i, heavy[3 * h], mean);
            __syncthreads();
            if (threadIdx.x == 0) n_heavy = 0;
        }
    }
}

// ancestor = first j with cdf[j] > target  (== searchsorted(cdf, target, side='right')), found hierarchically:
//   level 0: a <= 2048-entry sample of the coarse table in shared memory            (11 steps, no global traffic)
//   level 1: the coarse table (cdf at the end of every 32 entries; n/4 bytes: L2-resident)
//   level 2: the 32 cdf entries of the group = one 256-byte run of the cdf            (5 steps)
// so a lookup costs one DRAM sector of the cdf instead of log2(n) dependent probes scattered over 8n bytes.
constexpr int kTop = 2048;

struct TopTable {
    long long m, stride;
    int n_top;
};

__device__ __forceinline__ TopTable load_top(const unsigned long long* __restrict__ coarse, long long n, unsigned long long* top) {
    TopTable t;
    t.m = n_groups_for(n);
    t.n_top = t.m < kTop ? (int)t.m : kTop;
    t.stride = (t.m + t.n_top - 1) / t.n_top;
    for (int k = threadIdx.x; k < t.n_top; k += blockDim.x) {
        long long g = (long long)(k + 1) * t.stride - 1;
        top[k] = __ldg(coarse + (g < t.m ? g : t.m - 1));
    }
    __syncthreads();
    return t;
}

__device__ __forceinline__ long long hier_search(unsigned long long target, const unsigned long long* __restrict__ cdf,
                                                 const unsigned long long* __restrict__ coarse, const unsigned long long* top, const TopTable& t,
                                                 long long n) {
    int lo = 0, hi = t.n_top - 1;  // first k with top[k] > target (top[n_top - 1] == total > target)
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (top[mid] > target) hi = mid; else lo = mid + 1;
    }
    long long glo = (long long)lo * t.stride, ghi = glo + t.stride - 1;
    if (ghi > t.m - 1) ghi = t.m - 1;
    while (glo < ghi) {  // first group whose last entry exceeds the target
        const long long mid = (glo + ghi) >> 1;
        if (__ldg(coarse + mid) > target) ghi = mid; else glo = mid + 1;
    }
    long long jlo = glo * kGroup, jhi = jlo + kGroup - 1;
    if (jhi > n - 1) jhi = n - 1;
    while (jlo < jhi) {
        const long long mid = (jlo + jhi) >> 1;
        if (__ldg(cdf + mid) > target) jhi = mid; else jlo = mid + 1;
    }
    return jlo;
}

// Multinomial (unsorted uniforms): one slot per thread, hierarchical search.
__global__ void __launch_bounds__(kBlock) search_kernel(int kind, const double* __restrict__ w, const double* __restrict__ uniforms,
                                                        long long n, const double* __restrict__ stats, WsView ws,
                                                        const unsigned long long* __restrict__ cdf, const double* __restrict__ cum_in,
                                                        double* __restrict__ cum_out, int* __restrict__ anc, SearchGen g) {
    __shared__ unsigned long long top[kTop];
    const bool resample = stats[CDW_STAT_RESAMPLE] != 0.0;
    const double mean = stats[CDW_STAT_MEAN];
    search_rearm(g);
    if (!resample) {
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
            if (anc) anc[i] = (int)i;
            if (cum_out) cum_out[i] = w[i];  // null update: cum += inc (resamplers.py:136-156)
        }
        return;
    }
    const unsigned long long total = ws.hdr[WS_TOTAL];
    const double u0 = kind == CDW_RESAMPLE_SYSTEMATIC ? (g.gen ? philox_uniform53(g.seed, g.stream_id, 0) : uniforms[0]) : 0.0;
    if (g.gen && g.uniforms_out && kind == CDW_RESAMPLE_SYSTEMATIC && blockIdx.x == 0 && threadIdx.x == 0) g.uniforms_out[0] = u0;
    const TopTable t = load_top(ws.coarse, n, top);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double u;
        if (kind == CDW_RESAMPLE_SYSTEMATIC) u = __ddiv_rn(__dadd_rn((double)i, u0), (double)n);
        else if (g.gen) {
            u = philox_uniform53(g.seed, g.stream_id, i);
            if (g.uniforms_out) g.uniforms_out[i] = u;
        } else u = uniforms[i];
        anc[i] = (int)hier_search(fixed_target(u, total), cdf, ws.coarse, top, t, n);
        if (cum_out) {
            const double c = cum_in ? cum_in[i] : 0.0;
            cum_out[i] = CDW_ADD(c, CDW_ADD(mean, -c));  // _update_particle: cum += (mean - cum) (resamplers.py:196-197)
        }
    }
}

// Systematic (sorted targets): a block owns kSysSlots consecutive slots; their ancestors lie in one contiguous window
// of the cdf, which is staged once in shared memory (8 B of cdf read per slot instead of log2(n) probes).  Windows that do
// not fit (very uneven weights) fall back to the hierarchical search.
constexpr int kSysPerThread = 4;
constexpr int kSysSlots = kBlock * kSysPerThread;
constexpr int kSysWindow = 3072;

__global__ void __launch_bounds__(kBlock) search_systematic_kernel(const double* __restrict__ w, const double* __restrict__ uniforms, long long n,
                                                                   const double* __restrict__ stats, WsView ws,
                                                                   const unsigned long long* __restrict__ cdf, const double* __restrict__ cum_in,
                                                                   double* __restrict__ cum_out, int* __restrict__ anc, SearchGen g) {
    __shared__ unsigned long long top[kTop];
    __shared__ unsigned long long win[kSysWindow];
    __shared__ long long bounds[2];
    search_rearm(g);
    if (stats[CDW_STAT_RESAMPLE] == 0.0) {
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
            if (anc) anc[i] = (int)i;
            if (cum_out) cum_out[i] = w[i];
        }
        return;
    }
    const double mean = stats[CDW_STAT_MEAN];
    const unsigned long long total = ws.hdr[WS_TOTAL];
    const double u0 = g.gen ? philox_uniform53(g.seed, g.stream_id, 0) : uniforms[0];
    if (g.gen && g.uniforms_out && blockIdx.x == 0 && threadIdx.x == 0) g.uniforms_out[0] = u0;
    const TopTable t = load_top(ws.coarse, n, top);
    const long long n_chunks = (n + kSysSlots - 1) / kSysSlots;
    for (long long chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
        const long long i0 = chunk * kSysSlots;
        const long long i1 = i0 + kSysSlots < n ? i0 + kSysSlots : n;  // slots [i0, i1)
        __syncthreads();
        if (threadIdx.x < 2) {
            const long long i = threadIdx.x == 0 ? i0 : i1 - 1;
            const double u = __ddiv_rn(__dadd_rn((double)i, u0), (double)n);
            bounds[threadIdx.x] = hier_search(fixed_target(u, total), cdf, ws.coarse, top, t, n);
        }
        __syncthreads();
        const long long jlo = bounds[0], jhi = bounds[1];
        const bool fits = jhi - jlo + 1 <= kSysWindow;
        if (fits)
            for (long long j = jlo + threadIdx.x; j <= jhi; j += blockDim.x) win[j - jlo] = __ldg(cdf + j);
        __syncthreads();
        for (long long i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
            const double u = __ddiv_rn(__dadd_rn((double)i, u0), (double)n);
            const unsigned long long target = fixed_target(u, total);
            long long a;
            if (fits) {
                int lo = 0, hi = (int)(jhi - jlo);
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (win[mid] > target) hi = mid; else lo = mid + 1;
                }
                a = jlo + lo;
            } else {
                a = hier_search(target, cdf, ws.coarse, top, t, n);
            }
            anc[i] = (int)a;
            if (cum_out) {
                const double c = cum_in ? cum_in[i] : 0.0;
                cum_out[i] = CDW_ADD(c, CDW_ADD(mean, -c));
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ K2 + K3a fused (small n)
// One launch of 8 cooperating CTAs (cdw_resample.cuh) does everything after the propagate kernel for n <= kFusedMax
// particles: log-sum-exp / nESS / decision, Philox uniforms, integer cdf, search, cum update, re-arming of the running-min key.
// Same integer algorithm as the multi-kernel path => same ancestors bit for bit; for batches this small the
// multi-kernel path is pure launch latency (7 launches for ~30 us of work at n = 1e4).
constexpr int kFusedThreads = 128;   // x 8 cooperating CTAs = the 1024 virtual threads of the canonical statistics order
constexpr int kFusedMax = 24576;  // cdf in shared memory: 192 KB (covers 2 x 10^4 particles sharded over two GPUs)

__global__ void __launch_bounds__(kFusedThreads) fused_resample_kernel(ResampleArgs r) {
    extern __shared__ __align__(16) unsigned long long scratch[];  // 3 x 1024 words: the statistics tree of CTA 0
    resample_coop<kFusedThreads>(r, (int)blockIdx.x, (int)gridDim.x, scratch);
}

// ------------------------------------------------------------------------------------------------ K3b: gather
__global__ void __launch_bounds__(kBlock) gather_kernel(const float4* __restrict__ pos_in, const float4* __restrict__ vel_in,
                                                        const int* __restrict__ anc, float4* __restrict__ pos_out, float4* __restrict__ vel_out,
                                                        const double* __restrict__ aux_in, double* __restrict__ aux_out,
                                                        const int* __restrict__ label_in, int* __restrict__ label_out, long long n, int A) {
    const long long total = n * A;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long p = idx / A;
        const int a = (int)(idx - p * A);
        const long long src = anc[p];
        pos_out[idx] = __ldg(pos_in + src * A + a);
        if (vel_in) vel_out[idx] = __ldg(vel_in + src * A + a);
        if (a == 0) {
            if (aux_in) aux_out[p] = aux_in[src];
            if (label_in) label_out[p] = label_in[src];
        }
    }
}

__global__ void __launch_bounds__(kBlock) gather_peer_kernel(const float4* const* __restrict__ peer_pos, const float4* const* __restrict__ peer_vel,
                                                             const double* const* __restrict__ peer_aux, const int* const* __restrict__ peer_label,
                                                             long long n_per_rank, const int* __restrict__ anc, float4* __restrict__ pos_out,
                                                             float4* __restrict__ vel_out, double* __restrict__ aux_out, int* __restrict__ label_out,
                                                             long long n, int A) {
    const long long total = n * A;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long p = idx / A;
        const int a = (int)(idx - p * A);
        const long long g = anc[p];
        const int r = (int)(g / n_per_rank);
        const long long src = g - (long long)r * n_per_rank;
        pos_out[idx] = peer_pos[r][src * A + a];  // NVLink load when r is a peer (P2P-mapped pointer)
        if (peer_vel) vel_out[idx] = peer_vel[r][src * A + a];
        if (a == 0) {
            if (peer_aux) aux_out[p] = peer_aux[r][src];
            if (peer_label) label_out[p] = peer_label[r][src];
        }
    }
}

__global__ void fill_u64_kernel(unsigned long long* p, unsigned long long v, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}

__global__ void philox_uniforms_kernel(unsigned long long seed, unsigned long long stream_id, long long n, double* out) {
    const long long pairs = (n + 1) / 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < pairs; i += (long long)gridDim.x * blockDim.x) {
        u4 c;
        c.x = (uint32_t)i; c.y = (uint32_t)((unsigned long long)i >> 32); c.z = (uint32_t)stream_id; c.w = 0x5EEDu;
        u4 r = philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
        unsigned long long a = (((unsigned long long)r.x << 32) | r.y) >> 11, b = (((unsigned long long)r.z << 32) | r.w) >> 11;
        out[2 * i] = (double)a * 1.1102230246251565e-16;
        if (2 * i + 1 < n) out[2 * i + 1] = (double)b * 1.1102230246251565e-16;
    }
}

template <typename T>
__global__ void fma_peak_kernel(T* out, int iters) {
    T a0 = (T)threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const T m = (T)0.999, c = (T)0.001;
    for (int i = 0; i < iters; ++i) {
        a0 = a0 * m + c; a1 = a1 * m + c; a2 = a2 * m + c; a3 = a3 * m + c;
        a4 = a4 * m + c; a5 = a5 * m + c; a6 = a6 * m + c; a7 = a7 * m + c;
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

static int blocks_for(long long n, int per_block, int cap) {
    long long b = (n + per_block - 1) / per_block;
    if (b < 1) b = 1;
    return (int)(b < cap ? b : cap);
}

static int sm_count() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    return sms;
}

}  // namespace cdw

using namespace cdw;

extern "C" size_t cdw_weight_workspace_bytes(int64_t n) {
    if (n < 0) n = 0;
    return sizeof(unsigned long long) * WS_HEADER_WORDS + 2 * sizeof(double) * kMaxStatBlocks +
           2 * sizeof(unsigned long long) * (size_t)(n_tiles_for(n) + 1) + sizeof(unsigned long long) * (size_t)(n_groups_for(n) + 4) + 32;
}

extern "C" int cdw_weight_update(const double* u_new, const double* aux, const double* proposal, const double* cum, int64_t n,
                                 double* inc_out, double* w_out, uint64_t* wmin_key, cdw_stream stream) {
    CDW_REQUIRE(u_new && w_out && n >= 0, "null argument");
    if (n == 0) return CDW_OK;
    int grid = blocks_for(n, kBlock * 4, sm_count() * 8);
    weight_update_kernel<<<grid, kBlock, 0, (cudaStream_t)stream>>>(u_new, aux, proposal, cum, n, inc_out, w_out, (unsigned long long*)wmin_key);
    CDW_CUDA(cudaGetLastError());
    return CDW_OK;
}

static int weight_stats_launch(const double* w, int64_t n, double floor_threshold, const uint64_t* wmin_key, double* stats, void* workspace,
                               size_t workspace_bytes, cdw_stream stream, bool clear_header) {
    CDW_REQUIRE(w && wmin_key && stats && workspace && n > 0, "null argument or n == 0");
    CDW_REQUIRE(n < (1ll << 31), "n must fit int32 ancestors");
    CDW_REQUIRE(workspace_bytes >= cdw_weight_workspace_bytes(n), "workspace too small");
    WsView ws = ws_view(workspace, n);
    // every kernel that uses the header leaves it at rest (zero); a caller-facing call clears it anyway
    if (clear_header) CDW_CUDA(cudaMemsetAsync(ws.hdr, 0, sizeof(unsigned long long) * WS_HEADER_WORDS, (cudaStream_t)stream));
    int b = 0;
    while ((1ll << b) < n) ++b;  // ceil(log2 n)
    int grid = blocks_for(n, kBlock * 4, sm_count() * 8);
    if (grid > kMaxStatBlocks) grid = kMaxStatBlocks;
    weight_stats_kernel<<<grid, kBlock, 0, (cudaStream_t)stream>>>(w, n, floor_threshold, (const unsigned long long*)wmin_key, stats, ws, 61 - b);
    CDW_CUDA(cudaGetLastError());
    return CDW_OK;
}

extern "C" int cdw_weight_stats(const double* w, int64_t n, double floor_threshold, const uint64_t* wmin_key, double* stats,
                                void* workspace, size_t workspace_bytes, cdw_stream stream) {
    return weight_stats_launch(w, n, floor_threshold, wmin_key, stats, workspace, workspace_bytes, stream, true);
}

static int resample_indices_launch(int kind, const double* w, const double* uniforms, int64_t n, const double* stats, const double* cum_in,
                                   double* cum_out, int32_t* anc_out, uint64_t* cdf_out, void* workspace, size_t workspace_bytes,
                                   cdw_stream stream, SearchGen gen) {
    CDW_REQUIRE(w && stats && anc_out && cdf_out && workspace && n > 0, "null argument or n == 0");
    CDW_REQUIRE(kind == CDW_RESAMPLE_MULTINOMIAL || kind == CDW_RESAMPLE_SYSTEMATIC, "unknown resampler kind");
    CDW_REQUIRE(uniforms || gen.gen, "uniforms are required");
    CDW_REQUIRE(n < (1ll << 31), "n must fit int32 ancestors");
    CDW_REQUIRE(workspace_bytes >= cdw_weight_workspace_bytes(n), "workspace too small");
    WsView ws = ws_view(workspace, n);
    cudaStream_t st = (cudaStream_t)stream;
    const long long nt = n_tiles_for(n);
    int grid_t = (int)(nt < (long long)sm_count() * 16 ? nt : (long long)sm_count() * 16);
    static const bool dbg = getenv("CDW_DEBUG_SYNC") != nullptr;
#define CDW_DBG(name)                                                                       \
    if (dbg) {                                                                               \
        cudaError_t e_ = cudaStreamSynchronize(st);                                          \
        if (e_ == cudaSuccess) e_ = cudaGetLastError();                                      \
        if (e_ != cudaSuccess) { set_error("%s: %s", name, cudaGetErrorString(e_)); return CDW_ECUDA; } \
    }
    cdf_tile_sums_kernel<<<grid_t, kBlock, 0, st>>>(w, n, stats, ws);
    CDW_DBG("cdf_tile_sums_kernel")
    cdf_scan_tiles_kernel<<<1, 1024, 0, st>>>(n, (double*)stats, ws);
    CDW_DBG("cdf_scan_tiles_kernel")
    static const bool sys_search = getenv("CDW_SYSTEMATIC_SEARCH") != nullptr;   // A/B: the windowed search form
    if (kind == CDW_RESAMPLE_SYSTEMATIC && !sys_search) {
        EmitArgs em = {};
        em.uniforms = uniforms; em.cum_in = cum_in; em.cum_out = cum_out; em.anc = anc_out; em.gen = gen;
        cdf_tile_scan_kernel<true><<<grid_t, kBlock, 0, st>>>(w, n, stats, ws, (unsigned long long*)cdf_out, em);
        CDW_CUDA(cudaGetLastError());
        return CDW_OK;
    }
    cdf_tile_scan_kernel<false><<<grid_t, kBlock, 0, st>>>(w, n, stats, ws, (unsigned long long*)cdf_out, EmitArgs{});
    CDW_DBG("cdf_tile_scan_kernel")
    if (kind == CDW_RESAMPLE_SYSTEMATIC) {
        int grid_s = blocks_for(n, kSysSlots, sm_count() * 8);
        search_systematic_kernel<<<grid_s, kBlock, 0, st>>>(w, uniforms, n, stats, ws, (const unsigned long long*)cdf_out, cum_in, cum_out, anc_out, gen);
    } else {
        int grid_s = blocks_for(n, kBlock * 4, sm_count() * 8);
        search_kernel<<<grid_s, kBlock, 0, st>>>(kind, w, uniforms, n, stats, ws, (const unsigned long long*)cdf_out, cum_in, cum_out, anc_out, gen);
    }
    CDW_CUDA(cudaGetLastError());
    return CDW_OK;
}

extern "C" int cdw_resample_indices(int kind, const double* w, const double* uniforms, int64_t n, const double* stats, const double* cum_in,
                                    double* cum_out, int32_t* anc_out, uint64_t* cdf_out, void* workspace, size_t workspace_bytes,
                                    cdw_stream stream) {
    CDW_REQUIRE(uniforms, "uniforms are required");
    SearchGen none = {};
    return resample_indices_launch(kind, w, uniforms, n, stats, cum_in, cum_out, anc_out, cdf_out, workspace, workspace_bytes, stream, none);
}

extern "C" int cdw_smc_resample(int kind, const double* w, int64_t n, double floor_threshold, uint64_t* wmin_key, uint64_t seed, uint64_t stream_id,
                                double* stats, const double* cum_in, double* cum_out, int32_t* anc_out, uint64_t* cdf_out, double* uniforms,
                                void* workspace, size_t workspace_bytes, cdw_stream stream) {
    CDW_REQUIRE(w && wmin_key && stats && anc_out && cdf_out && uniforms && workspace && n > 0, "null argument or n == 0");
    CDW_REQUIRE(kind == CDW_RESAMPLE_MULTINOMIAL || kind == CDW_RESAMPLE_SYSTEMATIC, "unknown resampler kind");
    CDW_REQUIRE(n < (1ll << 31), "n must fit int32 ancestors");
    cudaStream_t st = (cudaStream_t)stream;
    static const char* no_fuse = getenv("CDW_NO_FUSED_RESAMPLE");
    if (n <= kFusedMax && !no_fuse) {
        int b = 0;
        while ((1ll << b) < n) ++b;
        ResampleArgs r = {};
        r.kind = kind; r.n = (int)n; r.s1 = 61 - b; r.threshold = floor_threshold; r.w = w; r.wmin_key = (unsigned long long*)wmin_key;
        r.seed = seed; r.stream_id = stream_id; r.stats = stats; r.cum_in = cum_in; r.cum_out = cum_out; r.anc = anc_out; r.uniforms_out = uniforms;
        r.cdf_spill = (unsigned long long*)cdf_out; r.ws = workspace;
        CDW_REQUIRE(workspace_bytes >= cdw_weight_workspace_bytes(n), "workspace too small");
        fused_resample_kernel<<<kVirtualThreads / kFusedThreads, kFusedThreads, sizeof(unsigned long long) * 3 * kVirtualThreads, st>>>(r);
        CDW_CUDA(cudaGetLastError());
        return CDW_OK;
    }
    // multi-kernel form: statistics, three scan kernels, search (which also draws the uniforms and re-arms the key): 5 launches
    int rc = weight_stats_launch(w, n, floor_threshold, wmin_key, stats, workspace, workspace_bytes, stream, false);
    if (rc) return rc;
    SearchGen gen = {};
    gen.gen = 1; gen.seed = seed; gen.stream_id = stream_id; gen.uniforms_out = uniforms; gen.rearm_key = (unsigned long long*)wmin_key;
    return resample_indices_launch(kind, w, uniforms, n, stats, cum_in, cum_out, anc_out, cdf_out, workspace, workspace_bytes, stream, gen);
}

extern "C" int cdw_gather_particles(const float* pos_in, const float* vel_in, const int32_t* anc, float* pos_out, float* vel_out,
                                    const double* aux_in, double* aux_out, const int32_t* label_in, int32_t* label_out, int64_t n,
                                    int32_t n_atoms, cdw_stream stream) {
    CDW_REQUIRE(pos_in && pos_out && anc && n >= 0 && n_atoms > 0, "null argument");
    CDW_REQUIRE(pos_in != pos_out && (!vel_in || (vel_out && vel_in != vel_out)), "gather needs distinct in/out buffers");
    CDW_REQUIRE((!aux_in || aux_out) && (!label_in || label_out), "missing output");
    if (n == 0) return CDW_OK;
    int grid = blocks_for(n * n_atoms, kBlock, sm_count() * 16);
    gather_kernel<<<grid, kBlock, 0, (cudaStream_t)stream>>>((const float4*)pos_in, (const float4*)vel_in, anc, (float4*)pos_out, (float4*)vel_out,
                                                            aux_in, aux_out, label_in, label_out, n, n_atoms);
    CDW_CUDA(cudaGetLastError());
    return CDW_OK;
}

extern "C" int cdw_gather_particles_peer(const float* const* peer_pos, const float* const* peer_vel, const double* const* peer_aux,
                                         const int32_t* const* peer_label, int32_t n_ranks, int64_t n_per_rank, const int32_t* anc,
                                         float* pos_out, float* vel_out, double* aux_out, int32_t* label_out, int64_t n_local, int32_t n_atoms,
                                         cdw_stream stream) {
    CDW_REQUIRE(peer_pos && pos_out && anc && n_ranks > 0 && n_per_rank > 0 && n_local >= 0 && n_atoms > 0, "null argument");
    CDW_REQUIRE((!peer_vel || vel_out) && (!peer_aux || aux_out) && (!peer_label || label_out), "missing output");
    if (n_local == 0) return CDW_OK;
    int grid = blocks_for(n_local * n_atoms, kBlock, sm_count() * 16);
    gather_peer_kernel<<<grid, kBlock, 0, (cudaStream_t)stream>>>((const float4* const*)peer_pos, (const float4* const*)peer_vel, peer_aux, peer_label,
                                                                 n_per_rank, anc, (float4*)pos_out, (float4*)vel_out, aux_out, label_out, n_local, n_atoms);
    CDW_CUDA(cudaGetLastError());
    return CDW_OK;
}

extern "C" int cdw_fill_u64(uint64_t* ptr, uint64_t value, int64_t n, cdw_stream stream) {
    CDW_REQUIRE(ptr && n >= 0, "null argument");
    if (n == 0) return CDW_OK;
    fill_u64_kernel<<<blocks_for(n, 256, 1024), 256, 0, (cudaStream_t)stream>>>((unsigned long long*)ptr, value, n);
    CDW_CUDA(cudaGetLastError());
    return CDW_OK;
}

extern "C" int cdw_philox_uniforms(uint64_t seed, uint64_t stream_id, int64_t n, double* out, cdw_stream stream) {
    CDW_REQUIRE(out && n >= 0, "null argument");
    if (n == 0) return CDW_OK;
    philox_uniforms_kernel<<<blocks_for((n + 1) / 2, 256, sm_count() * 16), 256, 0, (cudaStream_t)stream>>>(seed, stream_id, n, out);
    CDW_CUDA(cudaGetLastError());
    return CDW_OK;
}

extern "C" int cdw_ipc_get_handle(const void* dev_ptr, void* handle64_host) {
    CDW_REQUIRE(dev_ptr && handle64_host, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    CDW_CUDA(cudaIpcGetMemHandle((cudaIpcMemHandle_t*)handle64_host, (void*)dev_ptr));
    return CDW_OK;
}
extern "C" int cdw_ipc_open_handle(const void* handle64_host, void** dev_ptr_out) {
    CDW_REQUIRE(handle64_host && dev_ptr_out, "null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64_host, sizeof(h));
    CDW_CUDA(cudaIpcOpenMemHandle(dev_ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
    return CDW_OK;
}
extern "C" int cdw_ipc_close_handle(void* dev_ptr) {
    CDW_REQUIRE(dev_ptr, "null argument");
    CDW_CUDA(cudaIpcCloseMemHandle(dev_ptr));
    return CDW_OK;
}
extern "C" int cdw_device_malloc(size_t bytes, void** out) {
    CDW_REQUIRE(out && bytes > 0, "null argument");
    CDW_CUDA(cudaMalloc(out, bytes));
    return CDW_OK;
}
extern "C" int cdw_device_free(void* ptr) {
    if (ptr) CDW_CUDA(cudaFree(ptr));
    return CDW_OK;
}

extern "C" int cdw_fma_peak(int fp64, int iters, double* tflops_out, double* ms_out) {
    CDW_REQUIRE(tflops_out && iters > 0, "null argument");
    const int sms = sm_count(), blocks = sms * 8, threads = 512;
    void* buf = nullptr;
    CDW_CUDA(cudaMalloc(&buf, (size_t)blocks * threads * sizeof(double)));
    cudaEvent_t e0, e1;
    CDW_CUDA(cudaEventCreate(&e0));
    CDW_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        CDW_CUDA(cudaEventRecord(e0));
        if (fp64) fma_peak_kernel<double><<<blocks, threads>>>((double*)buf, iters);
        else fma_peak_kernel<float><<<blocks, threads>>>((float*)buf, iters);
        CDW_CUDA(cudaEventRecord(e1));
        CDW_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        CDW_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(buf);
    double flops = 2.0 * 8.0 * (double)iters * (double)blocks * (double)threads;
    *tflops_out = flops / (best * 1e-3) / 1e12;
    if (ms_out) *ms_out = best;
    return CDW_OK;
}
