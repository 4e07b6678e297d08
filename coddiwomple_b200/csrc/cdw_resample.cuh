// cdw_resample.cuh — resampling (second half of K2 + K3a) by a team of cooperating CTAs, as a device function shared by
//   * fused_resample_kernel (cdw_weights.cu): 8 CTAs x 128 threads launched between two propagate kernels, and
//   * the resample roles of smc_propagate_kernel (cdw_propagate.cu): 8 extra 128-thread CTAs of the SAME launch that wait
//     until every replica's opening evaluation has produced its weight and then resample while the other CTAs are
//     still integrating — an SMC iteration is then ONE kernel launch.
// Both produce the same bits: log-sum-exp statistics are accumulated by 1024 VIRTUAL threads (element i belongs to
// virtual thread i mod 1024, summed in index order; the team's 1024 real threads are exactly those) and combined by a
// fixed binary tree; everything after that is integer arithmetic (exact in any order).
//
// Reference semantics: nESS / vanilla_floor_threshold (coddiwomple/utils.py:41-72), Resampler.resample
// (coddiwomple/resamplers.py:61-133), MultinomialResampler._resample (coddiwomple/resamplers.py:254-272).
#pragma once
#include "cdw_common.h"

namespace cdw {

constexpr int kVirtualThreads = 1024;

__device__ __forceinline__ double pow2(int e) { return bits_to_double((uint64_t)(e + 1023) << 52); }

__device__ __forceinline__ unsigned long long quantise(double w, double wmin, double scale) {
    double e = det_exp(CDW_ADD(wmin, -w));
    return __double2ull_rn(CDW_MUL(e, scale));
}

__device__ __forceinline__ double philox_uniform53(unsigned long long seed, unsigned long long stream_id, long long i) {
    u4 c;
    const long long pair = i >> 1;
    c.x = (uint32_t)pair; c.y = (uint32_t)((unsigned long long)pair >> 32); c.z = (uint32_t)stream_id; c.w = 0x5EEDu;
    const u4 r = philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    const unsigned long long bits = (i & 1) ? ((((unsigned long long)r.z << 32) | r.w) >> 11) : ((((unsigned long long)r.x << 32) | r.y) >> 11);
    return (double)bits * 1.1102230246251565e-16;
}

// Scratch shared by the cooperating CTAs (in the weight workspace, see ws_view in cdw_weights.cu: 8 header words, then
// two arrays of 2048 doubles).
struct CoopScratch {
    unsigned int* barrier;        // header word 5
    double* red_e;                // [1024]
    unsigned long long* red_t;    // [1024]
    double* red_e2;               // [1024]
    unsigned long long* tot;      // [kMaxCoop] slice totals
    double* bcast;                // [4] resample flag, mean, shift
};
constexpr int kMaxCoop = 8;

__device__ __forceinline__ CoopScratch coop_scratch(void* ws) {
    CoopScratch c;
    unsigned char* b = (unsigned char*)ws;
    c.barrier = (unsigned int*)(b + 5 * 8);
    double* part_e = (double*)(b + 8 * 8);
    double* part_e2 = part_e + 2048;
    c.red_e = part_e; c.red_t = (unsigned long long*)(part_e + 1024);
    c.red_e2 = part_e2; c.tot = (unsigned long long*)(part_e2 + 1024); c.bcast = part_e2 + 1024 + kMaxCoop;
    return c;
}

// barrier among the R co-resident CTAs of the resampling team; `phase` counts barriers since the counter was last zero
__device__ __forceinline__ void coop_barrier(unsigned int* ctr, unsigned int R, unsigned int phase) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(ctr, 1u);
        while (atomicAdd(ctr, 0u) < R * phase) __nanosleep(64);
        __threadfence();
    }
    __syncthreads();
}

// One-iteration resampling by a team of R CTAs of NT threads, R * NT == 1024: real thread g = c * NT + tid IS virtual
// thread g of the canonical statistics order.  `c` = this CTA's index in the team.  The cdf lives in global memory
// (r.cdf_spill, L2-resident) as per-slice inclusive prefixes; slice offsets are added on the fly by the search.
// Weights are read with ld.global.cg: in the in-launch role they were written by other CTAs of the same kernel.
template <int NT>
__device__ void resample_coop(const ResampleArgs& r, int c, int R, unsigned long long* sm) {
    static_assert(kVirtualThreads % NT == 0 && NT % 32 == 0, "thread count");
    constexpr int kPer = 4;                  // consecutive elements per thread and scan chunk
    __shared__ unsigned long long warp_tot[32];
    __shared__ unsigned long long carry_sh;
    __shared__ unsigned long long off_sh[kMaxCoop + 1];
    const CoopScratch S = coop_scratch(r.ws);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n = r.n, g = c * NT + tid;
    const double wmin = ordered_key_inverse(*(volatile unsigned long long*)r.wmin_key);
    const double scale1 = pow2(r.s1);
    unsigned int phase = 0;
    // ---- log-sum-exp statistics: virtual thread g sums its elements in index order
    {
        double se = 0.0, se2 = 0.0;
        unsigned long long t1 = 0;
        for (int i = g; i < n; i += kVirtualThreads) {
            const double e = det_exp(CDW_ADD(wmin, -__ldcg(r.w + i)));
            se += e;
            se2 += e * e;
            t1 += __double2ull_rn(CDW_MUL(e, scale1));
        }
        S.red_e[g] = se; S.red_e2[g] = se2; S.red_t[g] = t1;
    }
    coop_barrier(S.barrier, R, ++phase);
    if (c == 0) {
        // the fixed binary tree over the 1024 partials: red[k] += red[k + s], s = 512 .. 1 (three arrays side by side in smem)
        double* te = (double*)sm; double* te2 = te + kVirtualThreads; unsigned long long* tt = sm + 2 * kVirtualThreads;
        for (int k = tid; k < kVirtualThreads; k += NT) { te[k] = __ldcg(S.red_e + k); te2[k] = __ldcg(S.red_e2 + k); tt[k] = __ldcg(S.red_t + k); }
        __syncthreads();
        for (int s_ = kVirtualThreads / 2; s_ > 0; s_ >>= 1) {
            for (int k = tid; k < s_; k += NT) { te[k] += te[k + s_]; te2[k] += te2[k + s_]; tt[k] += tt[k + s_]; }
            __syncthreads();
        }
        if (tid == 0) {
            const double sum_e = te[0], sum_e2 = te2[0];
            const unsigned long long T1 = tt[0];
            const int bitlen = 64 - __clzll((long long)T1);
            int shift = 60 - bitlen;
            shift = shift < 0 ? 0 : shift;
            const double ness = (sum_e * sum_e) / ((double)n * sum_e2);
            const double mean = wmin - log(sum_e) + log((double)n);
            const double res = (r.threshold < 0.0 || ness <= r.threshold) ? 1.0 : 0.0;
            double* stats = r.stats;
            stats[CDW_STAT_WMIN] = wmin; stats[CDW_STAT_SUM_E] = sum_e; stats[CDW_STAT_SUM_E2] = sum_e2; stats[CDW_STAT_NESS] = ness;
            stats[CDW_STAT_MEAN] = mean; stats[CDW_STAT_RESAMPLE] = res; stats[CDW_STAT_TOTAL] = 0.0; stats[CDW_STAT_NNAN] = 0.0;
            S.bcast[0] = res; S.bcast[1] = mean; S.bcast[2] = (double)(r.s1 + shift);
        }
        __syncthreads();
    }
    coop_barrier(S.barrier, R, ++phase);
    const bool resample = __ldcg(S.bcast) != 0.0;
    const double mean = __ldcg(S.bcast + 1);
    const double scale = pow2((int)__ldcg(S.bcast + 2));
    if (!resample) {
        for (int i = g; i < n; i += kVirtualThreads) {
            r.anc[i] = i;
            if (r.cum_out) r.cum_out[i] = __ldcg(r.w + i);  // null update (resamplers.py:136-156)
        }
    } else {
        // ---- integer cdf: CTA c scans the contiguous slice [c m, (c + 1) m) in chunks of NT * kPer elements
        const int m = ((n + R - 1) / R + kPer - 1) / kPer * kPer;
        const int j0 = c * m, j1 = j0 + m < n ? j0 + m : n;
        if (tid == 0) carry_sh = 0;
        __syncthreads();
        for (int base = j0; base < j1; base += NT * kPer) {
            const int i0 = base + tid * kPer;
            unsigned long long q[kPer];
            unsigned long long run = 0;
#pragma unroll
            for (int k = 0; k < kPer; ++k) {
                run += (i0 + k < j1) ? quantise(__ldcg(r.w + i0 + k), wmin, scale) : 0ull;
                q[k] = run;
            }
            unsigned long long incl = run;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            if (lane == 31) warp_tot[warp] = incl;
            __syncthreads();
            if (warp == 0) {
                const unsigned long long t = lane < NT / 32 ? warp_tot[lane] : 0ull;
                unsigned long long sc = t;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const unsigned long long u = __shfl_up_sync(0xffffffffu, sc, o);
                    if (lane >= o) sc += u;
                }
                if (lane < NT / 32) warp_tot[lane] = sc - t;  // exclusive prefix of the warp totals
            }
            __syncthreads();
            const unsigned long long off = carry_sh + warp_tot[warp] + incl - run;
#pragma unroll
            for (int k = 0; k < kPer; ++k)
                if (i0 + k < j1) r.cdf_spill[i0 + k] = q[k] + off;   // slice-local inclusive prefix
            __syncthreads();
            if (tid == NT - 1) carry_sh = off + run;
            __syncthreads();
        }
        if (tid == 0) S.tot[c] = carry_sh;
        coop_barrier(S.barrier, R, ++phase);
        if (tid == 0) {
            unsigned long long acc = 0;
            for (int q = 0; q < R; ++q) { off_sh[q] = acc; acc += __ldcg(S.tot + q); }
            off_sh[R] = acc;
        }
        __syncthreads();
        const unsigned long long total = off_sh[R];
        if (g == 0) r.stats[CDW_STAT_TOTAL] = (double)total;
        const bool systematic = r.kind == CDW_RESAMPLE_SYSTEMATIC;
        const double u0 = systematic ? philox_uniform53(r.seed, r.stream_id, 0) : 0.0;
        if (g == 0 && r.uniforms_out && systematic) r.uniforms_out[0] = u0;
        // ---- ancestors: first j with cdf[j] > floor(u * total): the slice by its end value, then a binary search of the slice's
        // local prefixes (L2 reads; kSlots independent searches per thread and trip hide the latency)
        constexpr int kSlots = 4;
        for (int i0 = g; i0 < n; i0 += kVirtualThreads * kSlots) {
            unsigned long long target[kSlots];
            int lo[kSlots], hi[kSlots];
#pragma unroll
            for (int k = 0; k < kSlots; ++k) {
                const int i = i0 + k * kVirtualThreads;
                double u = 0.0;
                if (i < n) {
                    if (systematic) u = __ddiv_rn(__dadd_rn((double)i, u0), (double)n);
                    else {
                        u = philox_uniform53(r.seed, r.stream_id, i);
                        if (r.uniforms_out) r.uniforms_out[i] = u;
                    }
                }
                const unsigned long long t = fixed_target(u, total);
                int sl = 0;
                while (sl < R - 1 && off_sh[sl + 1] <= t) ++sl;   // first slice whose end exceeds the target
                target[k] = t - off_sh[sl];
                lo[k] = sl * m; hi[k] = (sl * m + m < n ? sl * m + m : n) - 1;
            }
            bool more = true;
            while (more) {
                more = false;
#pragma unroll
                for (int k = 0; k < kSlots; ++k) {
                    if (lo[k] < hi[k]) {
                        const int mid = (lo[k] + hi[k]) >> 1;
                        if (__ldcg(r.cdf_spill + mid) > target[k]) hi[k] = mid; else lo[k] = mid + 1;
                        more |= lo[k] < hi[k];
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < kSlots; ++k) {
                const int i = i0 + k * kVirtualThreads;
                if (i < n) {
                    r.anc[i] = lo[k];
                    if (r.cum_out) {
                        const double cc = r.cum_in ? r.cum_in[i] : 0.0;
                        r.cum_out[i] = CDW_ADD(cc, CDW_ADD(mean, -cc));  // _update_particle: cum += (mean - cum) (resamplers.py:196-197)
                    }
                }
            }
        }
    }
    // ---- leave the shared state at rest for the next call: the last CTA out re-arms the running min and zeroes the counters
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        const unsigned int arrived = atomicAdd(S.barrier, 1u) + 1u;
        if (arrived == R * (phase + 1)) {
            *r.wmin_key = ~0ull;
            if (r.done) *r.done = 0u;
            __threadfence();
            *S.barrier = 0u;
        }
    }
}

}  // namespace cdw
