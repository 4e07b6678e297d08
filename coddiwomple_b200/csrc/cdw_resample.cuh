// cdw_resample.cuh — one-CTA resampling (second half of K2 + K3a) as a device function, shared by
//   * fused_resample_kernel (cdw_weights.cu): a 1024-thread CTA launched between two propagate kernels, and
//   * the resample role of smc_propagate_kernel (cdw_propagate.cu): one extra 128-thread CTA of the SAME launch that waits
//     until every replica's opening evaluation has produced its weight and then resamples while the other CTAs are
//     still integrating — an SMC iteration is then ONE kernel launch.
// Both produce the same bits: log-sum-exp statistics are accumulated by 1024 VIRTUAL threads (element i belongs to
// virtual thread i mod 1024, summed in index order) and combined by a fixed binary tree, whatever the real thread count;
// everything after that is integer arithmetic (exact in any order).
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

// fixed binary tree over the 1024 virtual-thread partials held in red[0..1023]: red[k] += red[k + s], s = 512 .. 1
template <int NT, typename T>
__device__ __forceinline__ T tree_reduce(T* red) {
    __syncthreads();
    for (int s = kVirtualThreads / 2; s > 0; s >>= 1) {
        for (int k = threadIdx.x; k < s; k += NT) red[k] += red[k + s];
        __syncthreads();
    }
    const T r = red[0];
    __syncthreads();
    return r;
}

// `sm`: the CTA's dynamic shared memory, `cap` 8-byte entries (cap >= 1024); cdf entries >= cap go to r.cdf_spill.
// Weights are read with ld.global.cg: in the in-launch role they were written by other CTAs of the same kernel.
template <int NT>
__device__ void resample_cta(const ResampleArgs& r, unsigned long long* sm, int cap) {
    static_assert(kVirtualThreads % NT == 0 && NT % 32 == 0, "thread count");
    constexpr int M = kVirtualThreads / NT;  // virtual threads per real thread
    constexpr int kPer = 4;                  // consecutive elements per thread and scan chunk
    __shared__ unsigned long long warp_tot[32];
    __shared__ unsigned long long carry_sh;
    __shared__ double bc[2];
    __shared__ int shift_sh;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n = r.n;
    const double wmin = ordered_key_inverse(*(volatile unsigned long long*)r.wmin_key);
    const double scale1 = pow2(r.s1);
    // ---- log-sum-exp statistics by 1024 virtual threads
    double se[M], se2[M];
    unsigned long long t1[M];
#pragma unroll
    for (int m = 0; m < M; ++m) {
        se[m] = 0.0; se2[m] = 0.0; t1[m] = 0;
        for (int i = tid + NT * m; i < n; i += kVirtualThreads) {
            const double e = det_exp(CDW_ADD(wmin, -__ldcg(r.w + i)));
            se[m] += e;
            se2[m] += e * e;
            t1[m] += __double2ull_rn(CDW_MUL(e, scale1));
        }
    }
    double* red = (double*)sm;
#pragma unroll
    for (int m = 0; m < M; ++m) red[tid + NT * m] = se[m];
    const double sum_e = tree_reduce<NT>(red);
#pragma unroll
    for (int m = 0; m < M; ++m) red[tid + NT * m] = se2[m];
    const double sum_e2 = tree_reduce<NT>(red);
#pragma unroll
    for (int m = 0; m < M; ++m) sm[tid + NT * m] = t1[m];
    const unsigned long long T1 = tree_reduce<NT>(sm);
    if (tid == 0) {
        const int bitlen = 64 - __clzll((long long)T1);
        int shift = 60 - bitlen;
        shift = shift < 0 ? 0 : shift;
        shift_sh = r.s1 + shift;
        const double ness = (sum_e * sum_e) / ((double)n * sum_e2);
        const double mean = wmin - log(sum_e) + log((double)n);
        const double res = (r.threshold < 0.0 || ness <= r.threshold) ? 1.0 : 0.0;
        double* stats = r.stats;
        stats[CDW_STAT_WMIN] = wmin; stats[CDW_STAT_SUM_E] = sum_e; stats[CDW_STAT_SUM_E2] = sum_e2; stats[CDW_STAT_NESS] = ness;
        stats[CDW_STAT_MEAN] = mean; stats[CDW_STAT_RESAMPLE] = res; stats[CDW_STAT_TOTAL] = 0.0; stats[CDW_STAT_NNAN] = 0.0;
        bc[0] = res; bc[1] = mean;
        carry_sh = 0;
        *r.wmin_key = ~0ull;  // re-arm the running min for the next propagate launch
    }
    __syncthreads();
    const bool resample = bc[0] != 0.0;
    const double mean = bc[1];
    if (!resample) {
        for (int i = tid; i < n; i += NT) {
            r.anc[i] = i;
            if (r.cum_out) r.cum_out[i] = __ldcg(r.w + i);  // null update (resamplers.py:136-156)
        }
        return;
    }
    const double scale = pow2(shift_sh);
    // ---- integer cdf: chunks of NT * kPer consecutive elements, block-wide inclusive scan + running carry
    for (int base = 0; base < n; base += NT * kPer) {
        const int i0 = base + tid * kPer;
        unsigned long long q[kPer];
        unsigned long long run = 0;
#pragma unroll
        for (int k = 0; k < kPer; ++k) {
            run += (i0 + k < n) ? quantise(__ldcg(r.w + i0 + k), wmin, scale) : 0ull;
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
        for (int k = 0; k < kPer; ++k) {
            const int j = i0 + k;
            if (j < n) {
                if (j < cap) sm[j] = q[k] + off; else r.cdf_spill[j] = q[k] + off;
            }
        }
        __syncthreads();
        if (tid == NT - 1) carry_sh = off + run;
        __syncthreads();
    }
    const unsigned long long total = carry_sh;
    if (tid == 0) r.stats[CDW_STAT_TOTAL] = (double)total;
    const bool systematic = r.kind == CDW_RESAMPLE_SYSTEMATIC;
    const double u0 = systematic ? philox_uniform53(r.seed, r.stream_id, 0) : 0.0;
    if (tid == 0 && r.uniforms_out && systematic) r.uniforms_out[0] = u0;
    // ---- ancestors: first j with cdf[j] > floor(u * total); kSlots independent searches per thread and trip
    constexpr int kSlots = 4;
    for (int i0 = tid; i0 < n; i0 += NT * kSlots) {
        unsigned long long target[kSlots];
        int lo[kSlots], hi[kSlots];
#pragma unroll
        for (int k = 0; k < kSlots; ++k) {
            const int i = i0 + k * NT;
            double u = 0.0;
            if (i < n) {
                if (systematic) u = __ddiv_rn(__dadd_rn((double)i, u0), (double)n);
                else {
                    u = philox_uniform53(r.seed, r.stream_id, i);
                    if (r.uniforms_out) r.uniforms_out[i] = u;
                }
            }
            target[k] = fixed_target(u, total);
            lo[k] = 0; hi[k] = n - 1;
        }
        bool more = true;
        while (more) {
            more = false;
#pragma unroll
            for (int k = 0; k < kSlots; ++k) {
                if (lo[k] < hi[k]) {
                    const int mid = (lo[k] + hi[k]) >> 1;
                    const unsigned long long c = mid < cap ? sm[mid] : r.cdf_spill[mid];
                    if (c > target[k]) hi[k] = mid; else lo[k] = mid + 1;
                    more |= lo[k] < hi[k];
                }
            }
        }
#pragma unroll
        for (int k = 0; k < kSlots; ++k) {
            const int i = i0 + k * NT;
            if (i < n) {
                r.anc[i] = lo[k];
                if (r.cum_out) {
                    const double c = r.cum_in ? r.cum_in[i] : 0.0;
                    r.cum_out[i] = CDW_ADD(c, CDW_ADD(mean, -c));  // _update_particle: cum += (mean - cum) (resamplers.py:196-197)
                }
            }
        }
    }
}

}  // namespace cdw
