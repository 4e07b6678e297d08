// cdw_weights.cu — K2 (incremental work / log-weight update, log-sum-exp + ESS reduction), K3a (integer cdf scan, direct
// emission of systematic ancestors, guide-table lookup for multinomial ancestors), K3b (coalesced particle gather) for sm_100a.
//
// Reference semantics: TargetFactory.compute_incremental_work (coddiwomple/distribution_factories.py:124-129),
// nESS / vanilla_floor_threshold (coddiwomple/utils.py:41-72), Resampler.resample (coddiwomple/resamplers.py:61-133),
// MultinomialResampler._resample (coddiwomple/resamplers.py:254-272).
//
// Exactness: EVERYTHING after the exponential is integer arithmetic, hence exactly associative — results do
// not depend on grid size, block schedule, summation order or GPU count, and equal oracle/resampling.py bit for bit:
//   * e_i = det_exp(min W - W_i) is evaluated ONCE per particle (correctly rounded mul/add/rint only: same bits on CPU and GPU) and
//     kept as q_i = rint(e_i 2^62) (uint64);
//   * sum e and sum e^2 (nESS, the normaliser) are 128-bit integer sums of q_i and rint(e_i^2 2^62);
//   * the cdf is the uint64 prefix sum of q_i >> k (rounded), k chosen from the exact sum so that 2^59 <= total < 2^61;
//   * ancestors are searched in integers (floor(u total) in 128-bit arithmetic).
//
// Passes over n particles (HBM bytes per particle):
//   weight_update / weight_compose   W = cum + inc, running min                          R 16-20  W 8-16
//   weight_stats                     q, the two integer sums; last block: statistics      R 8      W 8
//   cdf_tile_sums                    per-tile sums of q >> k; last block scans them       R 8
//   cdf_tile_scan<EMIT>              cdf; systematic ancestors or the multinomial guide   R 8      W 8 + 4
//   lookup (multinomial only)        bucket -> guide -> <= a few cdf probes               R 8 + ~64 (random sectors)  W 4
#include <stdlib.h>
#include <string.h>

#include "cdw_common.h"

namespace cdw {

constexpr int kBlock = 256;
constexpr int kItems = 8;
constexpr int kTile = kBlock * kItems;

// workspace header (uint64 words); every kernel that uses an accumulator leaves it at zero
enum { WS_TICKET = 0, WS_S1_LO = 1, WS_S1_HI = 2, WS_S2_LO = 3, WS_S2_HI = 4, WS_SHIFT = 5, WS_TOTAL = 6, WS_TICKET2 = 7, WS_HEADER_WORDS = 8 };

struct WsView {
    unsigned long long* hdr;
    unsigned long long* tile_sum;
    unsigned long long* tile_off;
    unsigned long long* q;   // [n] rint(e_i 2^62)
    int* guide;              // [n] multinomial: guide[b] = first j with cdf[j] > floor((b / n) total)
};

__host__ __device__ inline long long n_tiles_for(long long n) { return (n + kTile - 1) / kTile; }

__host__ __device__ inline size_t align16u(size_t v) { return (v + 15) & ~(size_t)15; }

__host__ __device__ inline WsView ws_view(void* ws, long long n) {
    WsView v;
    unsigned char* b = (unsigned char*)ws;
    v.hdr = (unsigned long long*)b; b += sizeof(unsigned long long) * WS_HEADER_WORDS;
    v.tile_sum = (unsigned long long*)b; b += align16u(sizeof(unsigned long long) * n_tiles_for(n));
    v.tile_off = (unsigned long long*)b; b += align16u(sizeof(unsigned long long) * (n_tiles_for(n) + 1));
    v.q = (unsigned long long*)b; b += align16u(sizeof(unsigned long long) * n);
    v.guide = (int*)b;
    return v;
}

__device__ __forceinline__ double philox_uniform53(unsigned long long seed, unsigned long long stream_id, long long i) {
    u4 c;
    const long long pair = i >> 1;
    c.x = (uint32_t)pair; c.y = (uint32_t)((unsigned long long)pair >> 32); c.z = (uint32_t)stream_id; c.w = 0x5EEDu;
    const u4 r = philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    const unsigned long long bits = (i & 1) ? ((((unsigned long long)r.z << 32) | r.w) >> 11) : ((((unsigned long long)r.x << 32) | r.y) >> 11);
    return (double)bits * 1.1102230246251565e-16;
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

// ------------------------------------------------------------------------------------------------ 128-bit unsigned sums
struct u128 {
    unsigned long long lo, hi;
};
__device__ __forceinline__ void add128(u128& a, unsigned long long v) {
    a.lo += v;
    a.hi += a.lo < v ? 1ull : 0ull;
}
__device__ __forceinline__ void add128(u128& a, const u128& b) {
    a.lo += b.lo;
    a.hi += b.hi + (a.lo < b.lo ? 1ull : 0ull);
}
__device__ __forceinline__ u128 block_sum128(u128 v, unsigned long long* sh /*[2 * kBlock/32]*/) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        u128 t;
        t.lo = __shfl_xor_sync(0xffffffffu, v.lo, o);
        t.hi = __shfl_xor_sync(0xffffffffu, v.hi, o);
        add128(v, t);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) { sh[2 * warp] = v.lo; sh[2 * warp + 1] = v.hi; }
    __syncthreads();
    u128 r;
    r.lo = sh[0]; r.hi = sh[1];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
        u128 t;
        t.lo = sh[2 * w]; t.hi = sh[2 * w + 1];
        add128(r, t);
    }
    return r;
}
// integers commute: any number of blocks may add their partial sums in any order
__device__ __forceinline__ void atomic_add128(unsigned long long* lo, unsigned long long* hi, const u128& v) {
    const unsigned long long old = atomicAdd(lo, v.lo);
    const unsigned long long carry = (old + v.lo < old) ? 1ull : 0ull;
    if (v.hi + carry) atomicAdd(hi, v.hi + carry);
}
// (hi 2^64 + lo) 2^-62 with two correctly rounded conversions, one add and an exact scaling (oracle/resampling.py:u128_to_double)
__device__ __forceinline__ double u128_to_double_2m62(const u128& v) {
    return __dmul_rn(__dadd_rn(__dmul_rn(__ull2double_rn(v.hi), 18446744073709551616.0), __ull2double_rn(v.lo)), 2.168404344971008868e-19);
}
// bit length of a 128-bit integer
__device__ __forceinline__ int bitlen128(const u128& v) { return v.hi ? 128 - __clzll((long long)v.hi) : 64 - __clzll((long long)v.lo); }

// q >> k rounded half up (k = 0: q itself)
__device__ __forceinline__ unsigned long long shift_round(unsigned long long q, int k) { return k ? ((q >> (k - 1)) + 1ull) >> 1 : q; }

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

// The look-ahead loop (cdw_smc_lookahead): the incremental work of slot i was computed one launch earlier, at the slot its state
// then lived in: W[i] = cum[i] + inc_src[anc_prev[i]]  (Resampler.resample's `cumulative_work + incremental_work`, resamplers.py:99-100).
__global__ void __launch_bounds__(kBlock) weight_compose_kernel(const double* __restrict__ inc_src, const int* __restrict__ anc_prev,
                                                                const double* __restrict__ cum, long long n, double* __restrict__ inc_out,
                                                                double* __restrict__ w_out, unsigned long long* wmin_key) {
    __shared__ unsigned long long sh[kBlock / 32];
    unsigned long long m = ~0ull;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double inc = inc_src[anc_prev ? (long long)anc_prev[i] : i];
        const double w = (cum ? cum[i] : 0.0) + inc;
        if (inc_out) inc_out[i] = inc;
        w_out[i] = w;
        const unsigned long long k = ordered_key(w);
        m = k < m ? k : m;
    }
    m = block_reduce(m, [](unsigned long long a, unsigned long long b) { return a < b ? a : b; }, sh);
    if (threadIdx.x == 0 && m != ~0ull) atomicMin(wmin_key, m);
}

// ------------------------------------------------------------------------------------------------ K2 pass 2
// One exponential per particle; everything it feeds is kept as integers.  `rearm` (may be NULL): the running-min key is reset to
// ~0 by the last block (every block has read it by then), so that the next iteration's weight pass can fold into it again.
__global__ void __launch_bounds__(kBlock) weight_stats_kernel(const double* __restrict__ w, long long n, double threshold,
                                                              const unsigned long long* __restrict__ wmin_key, double* stats, WsView ws,
                                                              unsigned long long* rearm) {
    __shared__ unsigned long long sh[2 * (kBlock / 32)];
    __shared__ bool is_last;
    const double wmin = ordered_key_inverse(*(const volatile unsigned long long*)wmin_key);
    const double two62 = 4611686018427387904.0;
    u128 s1 = {0ull, 0ull}, s2 = {0ull, 0ull};
    const long long n2 = n >> 1;
    // two 16-byte loads in flight and four independent exponentials per trip: the polynomial is a chain of 26 dependent (unfused, for
    // bit-reproducibility) fp64 operations, which one element pair alone cannot keep the pipe busy with
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
#define CDW_STATS_PAIR(V, I)                                                                                                    \
    {                                                                                                                           \
        const double ea = det_exp(CDW_ADD(wmin, -(V).x)), eb = det_exp(CDW_ADD(wmin, -(V).y));                                  \
        const unsigned long long qa = __double2ull_rn(CDW_MUL(ea, two62)), qb = __double2ull_rn(CDW_MUL(eb, two62));            \
        ((ulonglong2*)ws.q)[(I)] = make_ulonglong2(qa, qb);                                                                     \
        add128(s1, qa); add128(s1, qb);                                                                                         \
        add128(s2, __double2ull_rn(CDW_MUL(CDW_MUL(ea, ea), two62)));                                                           \
        add128(s2, __double2ull_rn(CDW_MUL(CDW_MUL(eb, eb), two62)));                                                           \
    }
    for (; i + stride < n2; i += 2 * stride) {
        const double2 v0 = __ldg((const double2*)w + i), v1 = __ldg((const double2*)w + i + stride);
        CDW_STATS_PAIR(v0, i)
        CDW_STATS_PAIR(v1, i + stride)
    }
    for (; i < n2; i += stride) {
        const double2 v = __ldg((const double2*)w + i);
        CDW_STATS_PAIR(v, i)
    }
#undef CDW_STATS_PAIR
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const double e = det_exp(CDW_ADD(wmin, -w[n - 1]));
        const unsigned long long q = __double2ull_rn(CDW_MUL(e, two62));
        ws.q[n - 1] = q;
        add128(s1, q);
        add128(s2, __double2ull_rn(CDW_MUL(CDW_MUL(e, e), two62)));
    }
    s1 = block_sum128(s1, sh);
    s2 = block_sum128(s2, sh);
    if (threadIdx.x == 0) {
        atomic_add128(&ws.hdr[WS_S1_LO], &ws.hdr[WS_S1_HI], s1);
        atomic_add128(&ws.hdr[WS_S2_LO], &ws.hdr[WS_S2_HI], s2);
        __threadfence();
        const unsigned long long ticket = atomicAdd(&ws.hdr[WS_TICKET], 1ull);
        is_last = ticket == (unsigned long long)gridDim.x - 1;
    }
    __syncthreads();
    if (is_last && threadIdx.x == 0) {
        __threadfence();
        volatile unsigned long long* h = ws.hdr;
        u128 S1, S2;
        S1.lo = h[WS_S1_LO]; S1.hi = h[WS_S1_HI]; S2.lo = h[WS_S2_LO]; S2.hi = h[WS_S2_HI];
        h[WS_S1_LO] = 0; h[WS_S1_HI] = 0; h[WS_S2_LO] = 0; h[WS_S2_HI] = 0; h[WS_TICKET] = 0;  // at rest for the next call: no memset needed
        const int b = bitlen128(S1);
        h[WS_SHIFT] = (unsigned long long)(b > 60 ? b - 60 : 0);   // total = sum (q >> k, rounded) then has 59..61 bits
        const double sum_e = u128_to_double_2m62(S1), sum_e2 = u128_to_double_2m62(S2);
        const double ness = (sum_e * sum_e) / ((double)n * sum_e2);
        stats[CDW_STAT_WMIN] = wmin;
        stats[CDW_STAT_SUM_E] = sum_e;
        stats[CDW_STAT_SUM_E2] = sum_e2;
        stats[CDW_STAT_NESS] = ness;
        stats[CDW_STAT_MEAN] = wmin - log(sum_e) + log((double)n);
        stats[CDW_STAT_RESAMPLE] = (threshold < 0.0 || ness <= threshold) ? 1.0 : 0.0;
        stats[CDW_STAT_TOTAL] = 0.0;
        stats[CDW_STAT_NNAN] = 0.0;
        if (rearm) *rearm = ~0ull;
    }
}

// ------------------------------------------------------------------------------------------------ K3a: integer cdf
__device__ __forceinline__ void load_tile_q(const unsigned long long* __restrict__ q_in, long long n, long long base, int k, unsigned long long (&q)[kItems]) {
    const long long i0 = base + (long long)threadIdx.x * kItems;
    if (i0 + kItems <= n) {
        const ulonglong2* p = (const ulonglong2*)(q_in + i0);  // i0 is a multiple of 8 -> 64-byte aligned
#pragma unroll
        for (int j = 0; j < kItems / 2; ++j) {
            const ulonglong2 v = __ldcg(p + j);
            q[2 * j] = shift_round(v.x, k);
            q[2 * j + 1] = shift_round(v.y, k);
        }
    } else {
#pragma unroll
        for (int j = 0; j < kItems; ++j) q[j] = (i0 + j < n) ? shift_round(__ldcg(q_in + i0 + j), k) : 0ull;
    }
}

// per-tile sums; the last block to finish scans them (exclusive offsets + grand total)
__global__ void __launch_bounds__(kBlock) cdf_tile_sums_kernel(long long n, double* stats, WsView ws) {
    if (stats[CDW_STAT_RESAMPLE] == 0.0) return;
    __shared__ unsigned long long sh[kBlock / 32];
    __shared__ unsigned long long carry_sh;
    __shared__ bool is_last;
    const int k = (int)ws.hdr[WS_SHIFT];
    const long long nt = n_tiles_for(n);
    for (long long tile = blockIdx.x; tile < nt; tile += gridDim.x) {
        unsigned long long q[kItems];
        load_tile_q(ws.q, n, tile * kTile, k, q);
        unsigned long long s = 0;
#pragma unroll
        for (int j = 0; j < kItems; ++j) s += q[j];
        s = block_reduce(s, [](unsigned long long a, unsigned long long b) { return a + b; }, sh);
        if (threadIdx.x == 0) ws.tile_sum[tile] = s;
    }
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned long long ticket = atomicAdd(&ws.hdr[WS_TICKET2], 1ull);
        is_last = ticket == (unsigned long long)gridDim.x - 1;
        carry_sh = 0;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long long base = 0; base < nt; base += kTile) {   // kItems consecutive tile sums per thread
        unsigned long long v[kItems], run = 0;
#pragma unroll
        for (int j = 0; j < kItems; ++j) {
            const long long i = base + (long long)threadIdx.x * kItems + j;
            v[j] = i < nt ? __ldcg(ws.tile_sum + i) : 0ull;
            run += v[j];
        }
        unsigned long long incl = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        __syncthreads();
        if (lane == 31) sh[warp] = incl;
        __syncthreads();
        unsigned long long off = carry_sh + incl - run;
        for (int wi = 0; wi < warp; ++wi) off += sh[wi];
#pragma unroll
        for (int j = 0; j < kItems; ++j) {
            const long long i = base + (long long)threadIdx.x * kItems + j;
            if (i < nt) ws.tile_off[i] = off;
            off += v[j];
        }
        __syncthreads();
        if (threadIdx.x == kBlock - 1) carry_sh = off;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        ws.hdr[WS_TOTAL] = carry_sh;
        ws.hdr[WS_TICKET2] = 0;
        stats[CDW_STAT_TOTAL] = (double)carry_sh;
    }
}

// The final kernel of a resampling call may draw its own uniforms (Philox, keyed by slot).
struct SearchGen {
    int gen;                              // 0: read `uniforms`; 1: generate (and store to uniforms_out if not NULL)
    unsigned long long seed, stream_id;
    double* uniforms_out;
};

// Sorted targets need no search: the tile scan holds cdf[j - 1] and cdf[j] in registers and emits item j's slots directly,
// [first(cdf[j - 1]), first(cdf[j])) with first(c) = min{i : target_i >= c}, target_i = floor(((i + u0) (1/n)) total).
// first() is estimated in floating point and corrected with the exact forward map.  Two uses:
//   EMIT_ANCESTORS  systematic resampling: slot i's ancestor is j (u0 = the one uniform of the draw);
//   EMIT_GUIDE      multinomial resampling: with u0 = 0 the same emission is a guide table, guide[b] = first j with
//                   cdf[j] > floor((b / n) total), which bounds the search of any target that falls into bucket b.
enum { EMIT_NONE = 0, EMIT_ANCESTORS = 1, EMIT_GUIDE = 2 };
struct EmitArgs {
    const double* uniforms;   // [1] u0 when gen.gen == 0
    const double* cum_in; double* cum_out; int* anc;
    SearchGen gen;
};

__device__ __forceinline__ unsigned long long systematic_target(long long i, double u0, double inv_n, unsigned long long total) {
    return fixed_target(__dmul_rn(__dadd_rn((double)i, u0), inv_n), total);
}

// first(c) = min{i : target(i) >= c}.  In real arithmetic that is ceil(c n / total - u0); the exact forward map differs from the real one by
// less than n 2^-50 slots (two roundings and the 53-bit truncation of (i + u0) / n, three roundings of the estimate), so the estimate IS the
// answer unless it lies within `margin` = n 2^-48 of an integer — only then (a few items in 10^7) the exact map is walked.
__device__ __forceinline__ long long first_slot(unsigned long long c, unsigned long long total, double u0, long long n, double inv_n, double n_over_total,
                                                double margin) {
    if (c == 0) return 0;
    const double x = (double)c * n_over_total - u0;
    const double up = ceil(x);
    long long i = (long long)up;
    i = i < 0 ? 0 : (i > n ? n : i);
    const double d = up - x;
    if (d > margin && d < 1.0 - margin) return i;
    while (i > 0 && systematic_target(i - 1, u0, inv_n, total) >= c) --i;
    while (i < n && systematic_target(i, u0, inv_n, total) < c) ++i;
    return i;
}

template <int EMIT>
__device__ __forceinline__ void emit_slot(const EmitArgs& e, long long i, long long j, double mean) {
    e.anc[i] = (int)j;
    if (EMIT == EMIT_ANCESTORS && e.cum_out) {
        const double c = e.cum_in ? e.cum_in[i] : 0.0;
        e.cum_out[i] = CDW_ADD(c, CDW_ADD(mean, -c));  // _update_particle: cum += (mean - cum) (resamplers.py:196-197)
    }
}

constexpr int kHeavyRange = 16;   // longer slot ranges (one ancestor copied many times) are written by the whole block
constexpr int kHeavyList = 64;

template <int EMIT>
__global__ void __launch_bounds__(kBlock) cdf_tile_scan_kernel(const double* __restrict__ w, long long n, const double* __restrict__ stats,
                                                               WsView ws, unsigned long long* __restrict__ cdf, EmitArgs em) {
    if (stats[CDW_STAT_RESAMPLE] == 0.0) {
        if (EMIT == EMIT_ANCESTORS) {
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
    const double u0 = EMIT == EMIT_ANCESTORS ? (em.gen.gen ? philox_uniform53(em.gen.seed, em.gen.stream_id, 0) : em.uniforms[0]) : 0.0;
    const double inv_n = __ddiv_rn(1.0, (double)n);
    const double n_over_total = EMIT ? (double)n / (double)total : 0.0;
    const double margin = (double)n * 3.5527136788005009e-15 + 1e-12;   // n 2^-48 (first_slot)
    if (EMIT == EMIT_ANCESTORS && em.gen.gen && em.gen.uniforms_out && blockIdx.x == 0 && threadIdx.x == 0) em.gen.uniforms_out[0] = u0;
    if (EMIT && threadIdx.x == 0) n_heavy = 0;
    const int k = (int)ws.hdr[WS_SHIFT];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long long tile = blockIdx.x; tile < n_tiles_for(n); tile += gridDim.x) {
        unsigned long long q[kItems];
        load_tile_q(ws.q, n, tile * kTile, k, q);
#pragma unroll
        for (int j = 1; j < kItems; ++j) q[j] += q[j - 1];
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
            for (int j = 0; j < kItems / 2; ++j) o2[j] = make_ulonglong2(q[2 * j] + off, q[2 * j + 1] + off);
        } else {
#pragma unroll
            for (int j = 0; j < kItems; ++j)
                if (i0 + j < n) cdf[i0 + j] = q[j] + off;
        }
        if (EMIT) {
            long long lo = first_slot(off, total, u0, n, inv_n, n_over_total, margin);
#pragma unroll
            for (int j = 0; j < kItems; ++j) {
                const long long item = i0 + j;
                const long long hi = item < n ? first_slot(q[j] + off, total, u0, n, inv_n, n_over_total, margin) : lo;
                if (hi - lo > kHeavyRange) {
                    const int slot = atomicAdd(&n_heavy, 1);
                    if (slot < kHeavyList) { heavy[3 * slot] = item; heavy[3 * slot + 1] = lo; heavy[3 * slot + 2] = hi; }
                    else for (long long i = lo; i < hi; ++i) emit_slot<EMIT>(em, i, item, mean);
                } else {
                    for (long long i = lo; i < hi; ++i) emit_slot<EMIT>(em, i, item, mean);
                }
                lo = hi;
            }
            __syncthreads();
            const int nh = n_heavy < kHeavyList ? n_heavy : kHeavyList;
            for (int h = 0; h < nh; ++h)
                for (long long i = heavy[3 * h + 1] + threadIdx.x; i < heavy[3 * h + 2]; i += blockDim.x) emit_slot<EMIT>(em, i, heavy[3 * h], mean);
            __syncthreads();
            if (threadIdx.x == 0) n_heavy = 0;
        }
    }
}

// Multinomial (unsorted uniforms): ancestor = first j with cdf[j] > target (== searchsorted(cdf, target, side='right')).
// target = floor(u total) falls into bucket b (floor((b/n) total) <= target < floor(((b+1)/n) total)), and the ancestor into
// [guide[b], guide[b + 1]] — on average one or two cdf entries: two dependent memory accesses instead of log2(n).
__global__ void __launch_bounds__(kBlock) lookup_kernel(const double* __restrict__ w, const double* __restrict__ uniforms, long long n,
                                                        const double* __restrict__ stats, WsView ws, const unsigned long long* __restrict__ cdf,
                                                        const double* __restrict__ cum_in, double* __restrict__ cum_out, int* __restrict__ anc, SearchGen g) {
    if (stats[CDW_STAT_RESAMPLE] == 0.0) {
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
            anc[i] = (int)i;
            if (cum_out) cum_out[i] = w[i];  // null update: cum += inc (resamplers.py:136-156)
        }
        return;
    }
    const double mean = stats[CDW_STAT_MEAN];
    const unsigned long long total = ws.hdr[WS_TOTAL];
    const double inv_n = __ddiv_rn(1.0, (double)n), nd = (double)n;
    const int* __restrict__ guide = ws.guide;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double u;
        if (g.gen) {
            u = philox_uniform53(g.seed, g.stream_id, i);
            if (g.uniforms_out) g.uniforms_out[i] = u;
        } else u = uniforms[i];
        const unsigned long long t = fixed_target(u, total);
        long long b = (long long)(u * nd);
        b = b < 0 ? 0 : (b > n - 1 ? n - 1 : b);
        while (b > 0 && systematic_target(b, 0.0, inv_n, total) > t) --b;
        while (b + 1 < n && systematic_target(b + 1, 0.0, inv_n, total) <= t) ++b;
        long long lo = __ldcg(guide + b), hi = b + 1 < n ? (long long)__ldcg(guide + b + 1) : n - 1;
        while (lo < hi) {
            const long long mid = (lo + hi) >> 1;
            if (__ldcg(cdf + mid) > t) hi = mid; else lo = mid + 1;
        }
        anc[i] = (int)lo;
        if (cum_out) {
            const double c = cum_in ? cum_in[i] : 0.0;
            cum_out[i] = CDW_ADD(c, CDW_ADD(mean, -c));  // _update_particle: cum += (mean - cum) (resamplers.py:196-197)
        }
    }
}

// ------------------------------------------------------------------------------------------------ K3b: gather
// One thread per float4 of the destination rows, kGatherUnroll elements in flight per thread: all ancestor indices first, then
// all row loads (the gather is latency-bound: bytes in flight per SM are what sets the bandwidth).
constexpr int kGatherUnroll = 4;

template <bool PEER>
__global__ void __launch_bounds__(kBlock) gather_kernel(const float4* __restrict__ pos_in, const float4* __restrict__ vel_in,
                                                        const float4* const* __restrict__ peer_pos, const float4* const* __restrict__ peer_vel,
                                                        const double* __restrict__ aux_in, const int* __restrict__ label_in,
                                                        const double* const* __restrict__ peer_aux, const int* const* __restrict__ peer_label,
                                                        long long n_per_rank, const int* __restrict__ anc, float4* __restrict__ pos_out,
                                                        float4* __restrict__ vel_out, double* __restrict__ aux_out, int* __restrict__ label_out,
                                                        long long n, int A) {
    const long long total = n * A;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const bool with_vel = PEER ? peer_vel != nullptr : vel_in != nullptr;
    for (long long base = (long long)blockIdx.x * blockDim.x + threadIdx.x; base < total; base += stride * kGatherUnroll) {
        long long p[kGatherUnroll], src[kGatherUnroll];
        int a[kGatherUnroll];
#pragma unroll
        for (int k = 0; k < kGatherUnroll; ++k) {
            const long long idx = base + k * stride;
            p[k] = idx < total ? idx / A : -1;
            a[k] = idx < total ? (int)(idx - p[k] * A) : 0;
            src[k] = idx < total ? (long long)__ldg(anc + p[k]) : 0;
        }
        float4 x[kGatherUnroll], v[kGatherUnroll];
#pragma unroll
        for (int k = 0; k < kGatherUnroll; ++k) {
            x[k] = make_float4(0.f, 0.f, 0.f, 0.f); v[k] = x[k];
            if (p[k] < 0) continue;
            if (PEER) {
                const int r = (int)(src[k] / n_per_rank);
                const long long l = src[k] - (long long)r * n_per_rank;
                x[k] = peer_pos[r][l * A + a[k]];  // NVLink load when r is a peer (P2P-mapped pointer)
                if (with_vel) v[k] = peer_vel[r][l * A + a[k]];
            } else {
                x[k] = __ldg(pos_in + src[k] * A + a[k]);
                if (with_vel) v[k] = __ldg(vel_in + src[k] * A + a[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < kGatherUnroll; ++k) {
            if (p[k] < 0) continue;
            const long long idx = base + k * stride;
            pos_out[idx] = x[k];
            if (with_vel) vel_out[idx] = v[k];
            if (a[k] == 0) {
                if (PEER) {
                    const int r = (int)(src[k] / n_per_rank);
                    const long long l = src[k] - (long long)r * n_per_rank;
                    if (peer_aux) aux_out[p[k]] = peer_aux[r][l];
                    if (peer_label) label_out[p[k]] = peer_label[r][l];
                } else {
                    if (aux_in) aux_out[p[k]] = aux_in[src[k]];
                    if (label_in) label_out[p[k]] = label_in[src[k]];
                }
            }
        }
    }
}

// Per-iteration history spill (Particle._incremental_works / _ancestry, particles.py:38-47): after iteration t slot i holds
// cumulative work cum[i] and the root label of its (possibly resampled) state, label[anc[i]].
__global__ void __launch_bounds__(kBlock) record_lineage_kernel(const double* __restrict__ cum, const int* __restrict__ label, const int* __restrict__ anc,
                                                                double* __restrict__ cum_row, int* __restrict__ label_row, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        cum_row[i] = cum[i];
        label_row[i] = label[anc ? anc[i] : i];
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

// One block: tell every peer "this rank has finished its epoch e" and wait until every peer has said so.  flags_local[r] is written
// by rank r (through its peer-mapped pointer to this rank's array).  The epoch is a device-side counter (*epoch_counter, one
// increment per call), so the launch carries no host-side constant and can be replayed from a CUDA graph; epochs only grow, the
// flag arrays never need clearing.  The wait is bounded: after ~2 s it gives up and records the failure in *status (the caller
// checks it; a hang would cost the whole node).
// skip_if_zero_a / _b (may be NULL): device flags that are identical on every rank (the resampling decisions of the two previous
// iterations); when both are given and zero the ancestors involved are the identity, no rank touches a peer's rows, and the wait
// is skipped (the signal is still sent: epochs stay in step).
__global__ void peer_barrier_kernel(unsigned long long* const* peer_flags, volatile unsigned long long* flags_local, int rank, int world,
                                    unsigned long long* epoch_counter, int* status, const double* skip_if_zero_a, const double* skip_if_zero_b) {
    __shared__ unsigned long long epoch_sh;
    if (threadIdx.x == 0) {
        epoch_sh = *epoch_counter + 1ull;
        *epoch_counter = epoch_sh;
    }
    __syncthreads();
    const unsigned long long epoch = epoch_sh;
    const int r = threadIdx.x;
    if (r >= world) return;
    __threadfence_system();
    if (r != rank) {
        *(volatile unsigned long long*)(peer_flags[r] + rank) = epoch;
    } else {
        flags_local[rank] = epoch;
    }
    __threadfence_system();
    if (skip_if_zero_a && skip_if_zero_b && *skip_if_zero_a == 0.0 && *skip_if_zero_b == 0.0) return;
    long long spins = 0;
    while (flags_local[r] < epoch) {
        __nanosleep(100);
        if (++spins > 20000000ll) {
            atomicExch(status, 1);
            break;
        }
    }
    __threadfence_system();
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

// ---- K3b over peer memory with TMA.  A particle row is 16 A contiguous bytes on the owning GPU; a chunk of kTmaRows destination rows is 16 A
// kTmaRows contiguous bytes here.  Each of the first kTmaRows threads asks for its ancestor's position and velocity rows with one bulk copy each
// (cp.async.bulk global -> shared, completion on an mbarrier: NVLink sees 16 A-byte requests instead of one 16-byte load per thread), and one
// elected thread writes the whole chunk back with two bulk stores (shared -> global).  Several small CTAs per SM overlap each other's phases.
constexpr int kTmaRows = 32;
constexpr int kTmaThreads = 64;

template <bool PEER>
__global__ void __launch_bounds__(kTmaThreads) gather_rows_tma_kernel(const float4* __restrict__ pos_in, const float4* __restrict__ vel_in,
                                                                      const double* __restrict__ aux_in, const int* __restrict__ label_in,
                                                                      const float4* const* __restrict__ peer_pos, const float4* const* __restrict__ peer_vel,
                                                                      const double* const* __restrict__ peer_aux, const int* const* __restrict__ peer_label,
                                                                      long long n_per_rank, const int* __restrict__ anc, float4* __restrict__ pos_out,
                                                                      float4* __restrict__ vel_out, double* __restrict__ aux_out, int* __restrict__ label_out,
                                                                      long long n, int A) {
    extern __shared__ __align__(128) unsigned char stage[];
    __shared__ __align__(8) unsigned long long bar;
    const unsigned row_bytes = 16u * (unsigned)A;
    unsigned char* spos = stage;
    unsigned char* svel = stage + (size_t)kTmaRows * row_bytes;
    const unsigned b = (unsigned)__cvta_generic_to_shared(&bar);
    const int tid = (int)threadIdx.x;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    unsigned parity = 0;
    const long long n_chunks = (n + kTmaRows - 1) / kTmaRows;
    for (long long c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        const long long base = c * kTmaRows;
        const int cnt = (int)(n - base < kTmaRows ? n - base : kTmaRows);
        if (tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"((unsigned)cnt * row_bytes * 2u) : "memory");
        if (tid < cnt) {
            const long long src = (long long)__ldg(anc + base + tid);
            const int r = PEER ? (int)(src / n_per_rank) : 0;
            const long long l = PEER ? src - (long long)r * n_per_rank : src;
            const float4* prow = (PEER ? peer_pos[r] : pos_in) + l * A;
            const float4* vrow = (PEER ? peer_vel[r] : vel_in) + l * A;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             (unsigned)__cvta_generic_to_shared(spos + (size_t)tid * row_bytes)), "l"(prow), "r"(row_bytes), "r"(b) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             (unsigned)__cvta_generic_to_shared(svel + (size_t)tid * row_bytes)), "l"(vrow), "r"(row_bytes), "r"(b) : "memory");
            if (aux_out) aux_out[base + tid] = PEER ? peer_aux[r][l] : __ldg(aux_in + l);
            if (label_out) label_out[base + tid] = PEER ? peer_label[r][l] : __ldg(label_in + l);
        }
        unsigned ok = 0;
        while (!ok) {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(b), "r"(parity) : "memory");
        }
        parity ^= 1u;
        if (tid == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(pos_out + base * A), "r"((unsigned)__cvta_generic_to_shared(spos)),
                         "r"((unsigned)cnt * row_bytes) : "memory");
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(vel_out + base * A), "r"((unsigned)__cvta_generic_to_shared(svel)),
                         "r"((unsigned)cnt * row_bytes) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the staging area may be overwritten
        }
        __syncthreads();
    }
}

static size_t gather_stage_bytes(int n_atoms) { return (size_t)2 * kTmaRows * 16 * n_atoms; }
static bool gather_tma_ok(int n_atoms) {
    static const bool no_tma = getenv("CDW_GATHER_NO_TMA") != nullptr;   // A/B switch
    return !no_tma && gather_stage_bytes(n_atoms) <= 96 * 1024;
}

static int gather_ctas_per_sm() {
    static const char* env = getenv("CDW_GATHER_CTAS");
    return env && atoi(env) > 0 ? atoi(env) : 8;   // measured (profiles/r2_notes.md): more CTAs in flight thrash L2 on randomly chosen rows
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
    const size_t nt = (size_t)n_tiles_for(n);
    return sizeof(unsigned long long) * WS_HEADER_WORDS + align16u(sizeof(unsigned long long) * nt) + align16u(sizeof(unsigned long long) * (nt + 1)) +
           align16u(sizeof(unsigned long long) * (size_t)n) + align16u(sizeof(int) * (size_t)n) + 32;
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
                               size_t workspace_bytes, cdw_stream stream, bool clear_header, uint64_t* rearm) {
    CDW_REQUIRE(w && wmin_key && stats && workspace && n > 0, "null argument or n == 0");
    CDW_REQUIRE(n < (1ll << 31), "n must fit int32 ancestors");
    CDW_REQUIRE(workspace_bytes >= cdw_weight_workspace_bytes(n), "workspace too small");
    CDW_REQUIRE(((size_t)w & 15) == 0 && ((size_t)workspace & 15) == 0, "w and workspace must be 16-byte aligned");
    WsView ws = ws_view(workspace, n);
    // every kernel that uses the header leaves it at rest (zero); a caller-facing call clears it anyway
    if (clear_header) CDW_CUDA(cudaMemsetAsync(ws.hdr, 0, sizeof(unsigned long long) * WS_HEADER_WORDS, (cudaStream_t)stream));
    int grid = blocks_for(n, kBlock * 8, sm_count() * 8);
    weight_stats_kernel<<<grid, kBlock, 0, (cudaStream_t)stream>>>(w, n, floor_threshold, (const unsigned long long*)wmin_key, stats, ws,
                                                                   (unsigned long long*)rearm);
    CDW_CUDA(cudaGetLastError());
    return CDW_OK;
}

extern "C" int cdw_weight_stats(const double* w, int64_t n, double floor_threshold, const uint64_t* wmin_key, double* stats,
                                void* workspace, size_t workspace_bytes, cdw_stream stream) {
    return weight_stats_launch(w, n, floor_threshold, wmin_key, stats, workspace, workspace_bytes, stream, true, nullptr);
}

static int resample_indices_launch(int kind, const double* w, const double* uniforms, int64_t n, double* stats, const double* cum_in,
                                   double* cum_out, int32_t* anc_out, uint64_t* cdf_out, void* workspace, size_t workspace_bytes,
                                   cdw_stream stream, SearchGen gen) {
    CDW_REQUIRE(w && stats && anc_out && cdf_out && workspace && n > 0, "null argument or n == 0");
    CDW_REQUIRE(kind == CDW_RESAMPLE_MULTINOMIAL || kind == CDW_RESAMPLE_SYSTEMATIC, "unknown resampler kind");
    CDW_REQUIRE(uniforms || gen.gen, "uniforms are required");
    CDW_REQUIRE(n < (1ll << 31), "n must fit int32 ancestors");
    CDW_REQUIRE(workspace_bytes >= cdw_weight_workspace_bytes(n), "workspace too small");
    CDW_REQUIRE(((size_t)cdf_out & 15) == 0 && ((size_t)workspace & 15) == 0, "cdf_out and workspace must be 16-byte aligned");
    WsView ws = ws_view(workspace, n);
    cudaStream_t st = (cudaStream_t)stream;
    const long long nt = n_tiles_for(n);
    int grid_t = (int)(nt < (long long)sm_count() * 8 ? nt : (long long)sm_count() * 8);
    static const bool dbg = getenv("CDW_DEBUG_SYNC") != nullptr;
#define CDW_DBG(name)                                                                       \
    if (dbg) {                                                                               \
        cudaError_t e_ = cudaStreamSynchronize(st);                                          \
        if (e_ == cudaSuccess) e_ = cudaGetLastError();                                      \
        if (e_ != cudaSuccess) { set_error("%s: %s", name, cudaGetErrorString(e_)); return CDW_ECUDA; } \
    }
    cdf_tile_sums_kernel<<<grid_t, kBlock, 0, st>>>(n, stats, ws);
    CDW_DBG("cdf_tile_sums_kernel")
    EmitArgs em = {};
    em.uniforms = uniforms; em.cum_in = cum_in; em.cum_out = cum_out; em.gen = gen;
    if (kind == CDW_RESAMPLE_SYSTEMATIC) {
        em.anc = anc_out;
        cdf_tile_scan_kernel<EMIT_ANCESTORS><<<grid_t, kBlock, 0, st>>>(w, n, stats, ws, (unsigned long long*)cdf_out, em);
        CDW_CUDA(cudaGetLastError());
        return CDW_OK;
    }
    em.anc = ws.guide;
    cdf_tile_scan_kernel<EMIT_GUIDE><<<grid_t, kBlock, 0, st>>>(w, n, stats, ws, (unsigned long long*)cdf_out, em);
    CDW_DBG("cdf_tile_scan_kernel")
    int grid_s = blocks_for(n, kBlock * 2, sm_count() * 8);
    lookup_kernel<<<grid_s, kBlock, 0, st>>>(w, uniforms, n, stats, ws, (const unsigned long long*)cdf_out, cum_in, cum_out, anc_out, gen);
    CDW_CUDA(cudaGetLastError());
    return CDW_OK;
}

extern "C" int cdw_resample_indices(int kind, const double* w, const double* uniforms, int64_t n, const double* stats, const double* cum_in,
                                    double* cum_out, int32_t* anc_out, uint64_t* cdf_out, void* workspace, size_t workspace_bytes,
                                    cdw_stream stream) {
    CDW_REQUIRE(uniforms, "uniforms are required");
    SearchGen none = {};
    return resample_indices_launch(kind, w, uniforms, n, (double*)stats, cum_in, cum_out, anc_out, cdf_out, workspace, workspace_bytes, stream, none);
}

extern "C" int cdw_smc_resample(int kind, const double* w, int64_t n, double floor_threshold, uint64_t* wmin_key, uint64_t seed, uint64_t stream_id,
                                double* stats, const double* cum_in, double* cum_out, int32_t* anc_out, uint64_t* cdf_out, double* uniforms,
                                void* workspace, size_t workspace_bytes, cdw_stream stream) {
    CDW_REQUIRE(w && wmin_key && stats && anc_out && cdf_out && uniforms && workspace && n > 0, "null argument or n == 0");
    CDW_REQUIRE(kind == CDW_RESAMPLE_MULTINOMIAL || kind == CDW_RESAMPLE_SYSTEMATIC, "unknown resampler kind");
    CDW_REQUIRE(n < (1ll << 31), "n must fit int32 ancestors");
    // statistics (the last block re-arms the running-min key), tile sums + their scan, tile scan with direct emission, and for
    // multinomial resampling the lookup, which draws the Philox uniforms itself: 3 or 4 launches, no host decision in between
    int rc = weight_stats_launch(w, n, floor_threshold, wmin_key, stats, workspace, workspace_bytes, stream, false, wmin_key);
    if (rc) return rc;
    SearchGen gen = {};
    gen.gen = 1; gen.seed = seed; gen.stream_id = stream_id; gen.uniforms_out = uniforms;
    return resample_indices_launch(kind, w, uniforms, n, stats, cum_in, cum_out, anc_out, cdf_out, workspace, workspace_bytes, stream, gen);
}

extern "C" int cdw_smc_weigh_resample(int kind, const double* inc_src, const int32_t* anc_prev, int64_t n, double floor_threshold, uint64_t* wmin_key,
                                      uint64_t seed, uint64_t stream_id, double* stats, double* cum, double* inc_out, double* w_out, int32_t* anc_out,
                                      uint64_t* cdf_out, double* uniforms, void* workspace, size_t workspace_bytes, cdw_stream stream) {
    CDW_REQUIRE(inc_src && cum && w_out && wmin_key && n > 0, "null argument or n == 0");
    CDW_REQUIRE(anc_prev != anc_out, "the previous and the new ancestors need distinct buffers");
    int grid = blocks_for(n, kBlock * 4, sm_count() * 8);
    weight_compose_kernel<<<grid, kBlock, 0, (cudaStream_t)stream>>>(inc_src, anc_prev, cum, n, inc_out, w_out, (unsigned long long*)wmin_key);
    CDW_CUDA(cudaGetLastError());
    return cdw_smc_resample(kind, w_out, n, floor_threshold, wmin_key, seed, stream_id, stats, cum, cum, anc_out, cdf_out, uniforms, workspace,
                            workspace_bytes, stream);
}

extern "C" int cdw_gather_particles(const float* pos_in, const float* vel_in, const int32_t* anc, float* pos_out, float* vel_out,
                                    const double* aux_in, double* aux_out, const int32_t* label_in, int32_t* label_out, int64_t n,
                                    int32_t n_atoms, cdw_stream stream) {
    CDW_REQUIRE(pos_in && pos_out && anc && n >= 0 && n_atoms > 0, "null argument");
    CDW_REQUIRE(pos_in != pos_out && (!vel_in || (vel_out && vel_in != vel_out)), "gather needs distinct in/out buffers");
    CDW_REQUIRE((!aux_in || aux_out) && (!label_in || label_out), "missing output");
    if (n == 0) return CDW_OK;
    if (gather_tma_ok(n_atoms) && vel_in && vel_out && (!aux_in == !aux_out) && (!label_in == !label_out)) {
        const long long chunks = (n + kTmaRows - 1) / kTmaRows, cap = (long long)sm_count() * gather_ctas_per_sm();
        CDW_CUDA(cudaFuncSetAttribute(gather_rows_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        gather_rows_tma_kernel<false><<<(int)(chunks < cap ? chunks : cap), kTmaThreads, gather_stage_bytes(n_atoms), (cudaStream_t)stream>>>(
            (const float4*)pos_in, (const float4*)vel_in, aux_in, label_in, nullptr, nullptr, nullptr, nullptr, 1, anc, (float4*)pos_out, (float4*)vel_out, aux_out,
            label_out, n, n_atoms);
        CDW_CUDA(cudaGetLastError());
        return CDW_OK;
    }
    int grid = blocks_for(n * n_atoms, kBlock * kGatherUnroll, sm_count() * 8);
    gather_kernel<false><<<grid, kBlock, 0, (cudaStream_t)stream>>>((const float4*)pos_in, (const float4*)vel_in, nullptr, nullptr, aux_in, label_in, nullptr,
                                                                   nullptr, 1, anc, (float4*)pos_out, (float4*)vel_out, aux_out, label_out, n, n_atoms);
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
    if (gather_tma_ok(n_atoms) && peer_vel && (!peer_aux == !aux_out) && (!peer_label == !label_out)) {
        const long long chunks = (n_local + kTmaRows - 1) / kTmaRows, cap = (long long)sm_count() * gather_ctas_per_sm();
        CDW_CUDA(cudaFuncSetAttribute(gather_rows_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        gather_rows_tma_kernel<true><<<(int)(chunks < cap ? chunks : cap), kTmaThreads, gather_stage_bytes(n_atoms), (cudaStream_t)stream>>>(
            nullptr, nullptr, nullptr, nullptr, (const float4* const*)peer_pos, (const float4* const*)peer_vel, peer_aux, peer_label, n_per_rank, anc,
            (float4*)pos_out, (float4*)vel_out, aux_out, label_out, n_local, n_atoms);
        CDW_CUDA(cudaGetLastError());
        return CDW_OK;
    }
    int grid = blocks_for(n_local * n_atoms, kBlock * kGatherUnroll, sm_count() * 8);
    gather_kernel<true><<<grid, kBlock, 0, (cudaStream_t)stream>>>(nullptr, nullptr, (const float4* const*)peer_pos, (const float4* const*)peer_vel, nullptr,
                                                                  nullptr, peer_aux, peer_label, n_per_rank, anc, (float4*)pos_out, (float4*)vel_out, aux_out,
                                                                  label_out, n_local, n_atoms);
    CDW_CUDA(cudaGetLastError());
    return CDW_OK;
}

extern "C" int cdw_record_lineage(const double* cum, const int32_t* label, const int32_t* anc, double* cum_row, int32_t* label_row, int64_t n,
                                  cdw_stream stream) {
    CDW_REQUIRE(cum && label && cum_row && label_row && n >= 0, "null argument");
    if (n == 0) return CDW_OK;
    record_lineage_kernel<<<blocks_for(n, kBlock * 4, sm_count() * 8), kBlock, 0, (cudaStream_t)stream>>>(cum, label, anc, cum_row, label_row, n);
    CDW_CUDA(cudaGetLastError());
    return CDW_OK;
}

extern "C" int cdw_peer_barrier(uint64_t* const* peer_flags, uint64_t* flags_local, int32_t rank, int32_t world, uint64_t* epoch_counter,
                                int32_t* status, const double* skip_if_zero_a, const double* skip_if_zero_b, cdw_stream stream) {
    CDW_REQUIRE(peer_flags && flags_local && epoch_counter && status && world > 0 && world <= 32 && rank >= 0 && rank < world, "bad argument");
    peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((unsigned long long* const*)peer_flags, (volatile unsigned long long*)flags_local, rank, world,
                                                           (unsigned long long*)epoch_counter, status, skip_if_zero_a, skip_if_zero_b);
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
