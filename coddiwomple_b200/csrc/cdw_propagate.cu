// cdw_propagate.cu — K1 (batched reduced potential) and K4 (fused SMC propagate: gather-on-load, energy+force,
// incremental work, BAOAB + SHAKE/RATTLE + Philox, closing energy) for sm_100a.
//
// Mapping: one warp per replica, kWarps replicas in flight per CTA, persistent CTAs looping over
// replicas.  The per-replica algorithm lives in cdw_replica.cuh (shared with the host test harness); this file is
// the kernel shells and the C ABI.
//   * prologue (once per CTA): the lambda vector is resolved into "effective" term records in shared memory;
//   * a replica's positions are loaded as coalesced float4 (one lane per atom) into fp64 shared arrays and stay on
//     chip for all substeps; forces are evaluated in fp64, written per term to an fp32 scratch slot (no atomics:
//     shared-memory float/double atomicAdd is a CAS loop on sm_100) and summed per atom through a CSR list;
//   * constraints: one lane per constraint cluster, Gauss-Seidel SHAKE / velocity RATTLE.
#include "cdw_common.h"
#include "cdw_replica.cuh"

namespace cdw {

__global__ void __launch_bounds__(kThreads) reduced_potential_kernel(SysDev s, PropArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const Layout L = make_layout(s, false, kWarps);
    const Tables T = table_pointers(L, smem);
    build_tables(s, T, a.lambda, 1.0, false);
    const int warp = threadIdx.x >> 5;
    const WarpMem W = warp_pointers(L, smem, warp);
    for (long long p = (long long)blockIdx.x * kWarps + warp; p < a.n; p += (long long)gridDim.x * kWarps) replica_potential(s, T, W, a, p);
}

__global__ void __launch_bounds__(kThreads, CDW_MINBLOCKS) smc_propagate_kernel(SysDev s, PropArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ unsigned long long cta_min[kWarps];
    const Layout L = make_layout(s, true, kWarps);
    const Tables T = table_pointers(L, smem);
    build_tables(s, T, a.lambda, a.kT, true);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const WarpMem W = warp_pointers(L, smem, warp);
    unsigned long long my_min = ~0ull;
    for (long long p = (long long)blockIdx.x * kWarps + warp; p < a.n; p += (long long)gridDim.x * kWarps) {
        unsigned long long key = replica_step(s, T, W, L.ns_pad, a, p);
        my_min = key < my_min ? key : my_min;
    }
    // one atomic per CTA for the running min of W (the log-sum-exp shift of K2)
    if (lane == 0) cta_min[warp] = my_min;
    __syncthreads();
    if (a.wmin_key && threadIdx.x == 0) {
        unsigned long long m = cta_min[0];
        for (int w = 1; w < kWarps; ++w) m = cta_min[w] < m ? cta_min[w] : m;
        if (m != ~0ull) atomicMin(a.wmin_key, m);
    }
}

static int grid_for(const void* kernel, size_t smem, long long n) {
    int dev = 0, sms = 148, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    long long need = (n + kWarps - 1) / kWarps;
    long long cap = (long long)sms * per_sm;
    return (int)(need < cap ? (need > 0 ? need : 1) : cap);
}

static int launch(const cdw_system* sys, PropArgs& a, bool forces, cudaStream_t st) {
    const Layout L = make_layout(sys->dev, forces, kWarps);
    const void* k = forces ? (const void*)smc_propagate_kernel : (const void*)reduced_potential_kernel;
    if (L.total > 227 * 1024) {
        set_error("system needs %zu bytes of shared memory per CTA (> 227 KB)", L.total);
        return CDW_EUNSUPPORTED;
    }
    CDW_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
    if (a.n == 0) return CDW_OK;
    int grid = grid_for(k, L.total, a.n);
    if (forces)
        smc_propagate_kernel<<<grid, kThreads, L.total, st>>>(sys->dev, a);
    else
        reduced_potential_kernel<<<grid, kThreads, L.total, st>>>(sys->dev, a);
    CDW_CUDA(cudaGetLastError());
    return CDW_OK;
}

}  // namespace cdw

using namespace cdw;

extern "C" int cdw_reduced_potential(const cdw_system* sys, const float* pos, int64_t n, const double* lambda, double beta,
                                     double* u_out, cdw_stream stream) {
    CDW_REQUIRE(sys && n >= 0, "null argument");
    if (n == 0) return CDW_OK;
    CDW_REQUIRE(pos && u_out, "null argument");
    CDW_REQUIRE(lambda || sys->dev.n_lambda == 0, "lambda is required");
    PropArgs a = {};
    a.pos_in = (const float4*)pos; a.n = n; a.lambda = lambda; a.beta = beta; a.energy_out = u_out;
    return launch(sys, a, false, (cudaStream_t)stream);
}

extern "C" int cdw_energy_forces(const cdw_system* sys, const float* pos, int64_t n, const double* lambda, double* energy_out,
                                 float* force_out, cdw_stream stream) {
    CDW_REQUIRE(sys && n >= 0, "null argument");
    if (n == 0) return CDW_OK;
    CDW_REQUIRE(pos, "null argument");
    CDW_REQUIRE(lambda || sys->dev.n_lambda == 0, "lambda is required");
    PropArgs a = {};
    a.pos_in = (const float4*)pos; a.n = n; a.lambda = lambda; a.kT = 1.0; a.dt = 0.0; a.n_steps = 0; a.tol = 1e-6;
    a.energy_out = energy_out; a.force_out = (float4*)force_out;
    return launch(sys, a, true, (cudaStream_t)stream);
}

extern "C" int cdw_smc_propagate(const cdw_system* sys, const float* pos_in, const float* vel_in, float* pos_out, float* vel_out,
                                 int64_t n, const int32_t* anc, const double* aux_in, const double* cum_in, const int32_t* label_in,
                                 const double* lambda, const cdw_langevin_params* prm, uint64_t seed, uint64_t step0, int64_t gid0,
                                 double* inc_out, double* w_out, double* aux_out, int32_t* label_out, double* u_first_out,
                                 double* u_last_out, double* shadow_out, double* proposal_out, int32_t* nan_flags, uint64_t* wmin_key,
                                 cdw_stream stream) {
    CDW_REQUIRE(sys && prm && n >= 0, "null argument");
    if (n == 0) return CDW_OK;
    CDW_REQUIRE(pos_in, "null argument");
    CDW_REQUIRE(lambda || sys->dev.n_lambda == 0, "lambda is required");
    CDW_REQUIRE(prm->n_steps >= 0 && prm->dt >= 0.0 && prm->kT > 0.0 && prm->gamma >= 0.0 && prm->constraint_tol > 0.0, "bad Langevin parameters");
    CDW_REQUIRE(!(anc && (pos_in == pos_out || (vel_in && vel_in == vel_out))), "gather-on-load needs distinct in/out buffers");
    CDW_REQUIRE(vel_in || (prm->flags & CDW_FLAG_RANDOMIZE_VELOCITIES) || prm->n_steps == 0, "velocities required unless they are randomized");
    CDW_REQUIRE(!pos_out || prm->n_steps == 0 || vel_out, "vel_out is required when positions are written");
    PropArgs a = {};
    a.pos_in = (const float4*)pos_in; a.vel_in = (const float4*)vel_in; a.pos_out = (float4*)pos_out; a.vel_out = (float4*)vel_out;
    a.n = n; a.anc = anc; a.aux_in = aux_in; a.cum_in = cum_in; a.label_in = label_in; a.lambda = lambda;
    a.dt = prm->dt; a.gamma = prm->gamma; a.kT = prm->kT; a.tol = prm->constraint_tol; a.n_steps = prm->n_steps; a.flags = prm->flags;
    a.seed = seed; a.step0 = step0; a.gid0 = gid0;
    a.inc_out = inc_out; a.w_out = w_out; a.aux_out = aux_out; a.label_out = label_out; a.u_first_out = u_first_out; a.u_last_out = u_last_out;
    a.shadow_out = shadow_out; a.proposal_out = proposal_out; a.nan_flags = nan_flags; a.wmin_key = (unsigned long long*)wmin_key;
    return launch(sys, a, true, (cudaStream_t)stream);
}

extern "C" int cdw_baoab(const cdw_system* sys, float* pos, float* vel, int64_t n, const double* lambda, const cdw_langevin_params* prm,
                         uint64_t seed, uint64_t step0, int64_t gid0, double* u_first_out, double* u_last_out, double* shadow_out,
                         double* proposal_out, int32_t* nan_flags, cdw_stream stream) {
    // in place is safe without anc: every replica is read and written by the same warp
    return cdw_smc_propagate(sys, pos, vel, pos, vel, n, nullptr, nullptr, nullptr, nullptr, lambda, prm, seed, step0, gid0, nullptr,
                             nullptr, nullptr, nullptr, u_first_out, u_last_out, shadow_out, proposal_out, nan_flags, nullptr, stream);
}
