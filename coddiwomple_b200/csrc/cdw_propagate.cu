// cdw_propagate.cu — K1 (batched reduced potential) and K4 (fused SMC propagate: gather-on-load, energy+force,
// incremental work, BAOAB + SHAKE/RATTLE + Philox, closing energy) and the look-ahead iteration (the same, with the evaluation
// that opens iteration t + 1 moved to the end of launch t: the next weights are a by-product) for sm_100a.
//
// Mapping: a TEAM of T lanes per replica (T = 4 for 13-22 atom systems), 32 replica columns per CTA.  The
// per-replica algorithm lives in cdw_lane.cuh (shared with the host test harness); this file is the kernel shells, the
// launch logic and the C ABI.
//   * prologue (once per CTA): the lambda-resolved "effective" term records arrive in shared memory, either by one TMA bulk
//     copy of a prebuilt image (registered protocols) or built in place;
//   * a batch of rows is gathered by the whole CTA (4-byte cp.async straight into [component][replica] shared arrays; row
//     index = ancestor), where the state stays on chip for all substeps; every lane of a warp evaluates the same term of
//     different replicas or, inside a team, the next term of the same replica (no atomics, per-lane fp32 force copies);
//   * constraints: exact k x k velocity projection and Gauss-Seidel SHAKE per cluster;
//   * smc_lookahead_kernel: a second table image (lambda_{t+1}) replaces the first one on chip before the trailing evaluation.
#include <stdlib.h>

#include "cdw_common.h"
#include "cdw_lane.cuh"

namespace cdw {

constexpr int kMaxThreads = 256;

struct Shape {
    int R, T;   // replicas staged per CTA, lanes per replica; blockDim = R * T
    int metro;  // the program has a Metropolised block: the layout holds the state saved at "{"
    int bulk;   // rows arrive by TMA bulk copies (cta_load_rows_bulk)
};

__global__ void __launch_bounds__(kMaxThreads) reduced_potential_kernel(SysDev s, PropArgs a, Shape sh) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int T_ = sh.T, sw = 32 / T_;
    const Layout L = make_layout(s, false, kR, T_);
    const Tables T = table_pointers(L, smem);
    build_tables(s, T, a.lambda, 1.0, false);
    LaneMem M = lane_pointers(L, smem, T_, threadIdx.x);
    for (long long p0 = (long long)blockIdx.x * kR; p0 < a.n; p0 += (long long)gridDim.x * kR) {
        const long long left = a.n - p0;
        const int count = left < kR ? (int)left : kR;
        cta_load_rows(s, a, smem, L, sw, p0, count, false);
        const bool valid = M.rid < count;
        M.mask = __ballot_sync(0xffffffffu, valid);
        if (valid) {
            const double e = lane_energy<false>(s, T, M);
            if (M.tl == 0) a.energy_out[p0 + M.rid] = a.beta * e;
        }
        __syncthreads();
    }
}

// ---- prebuilt tables: one block per protocol row resolves lambda into the table image once; the propagate kernel then
// fetches the image with a single TMA bulk copy (cp.async.bulk, completion on an mbarrier) that overlaps the row loads.
__global__ void __launch_bounds__(128) build_tables_kernel(SysDev s, const double* lambda_rows, double kT, unsigned char* out, size_t stride) {
    extern __shared__ __align__(16) unsigned char smem[];
    const Layout L = make_layout(s, true, kR, 1);
    const Tables T = table_pointers(L, smem);
    build_tables(s, T, lambda_rows + (size_t)blockIdx.x * s.n_lambda, kT, true);
    __syncthreads();
    const uint4* src = (const uint4*)smem;
    uint4* dst = (uint4*)(out + (size_t)blockIdx.x * stride);
    for (size_t i = threadIdx.x; i < L.x / 16; i += blockDim.x) dst[i] = src[i];
}

__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// One elected thread announces `bytes` and starts the bulk copy.  The barrier is initialised once per CTA (init_table_barrier +
// __syncthreads); every copy is the next phase of the same barrier (wait_tables takes the phase parity).
__device__ __forceinline__ void init_table_barrier(unsigned long long* bar) {
    const unsigned b = smem_addr(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void bulk_load_tables(unsigned char* smem, const unsigned char* src, unsigned bytes, unsigned long long* bar) {
    const unsigned b = smem_addr(bar);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic-proxy reads/writes of the image area are ordered before the async-proxy writes
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(smem)), "l"(src), "r"(bytes), "r"(b)
                 : "memory");
}

__device__ __forceinline__ void wait_tables(unsigned long long* bar, unsigned parity) {
    const unsigned b = smem_addr(bar);
    unsigned ok = 0;
    while (!ok) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(b), "r"(parity) : "memory");
    }
}

// ---- TMA row gather.  Every particle row is 16 A contiguous bytes (float4 [A]) in HBM — local, or on a peer GPU when the run is sharded
// and the ancestor lives there.  One bulk copy per row and array (cp.async.bulk, completion on an mbarrier) brings the rows into a
// row-major staging area that borrows the (still unused) force copies, and the CTA transposes them into the [component][replica] columns:
// every byte crosses NVLink once, in 16 A-byte requests.  (The 4-byte cp.async path requests every 32-byte sector of a row three times —
// once per component — and peer reads are not cached on this side: on 8 GPUs that redundancy saturated NVLink, profiles/r2_notes.md.)
// Needs T >= 4 (three staging arrays of 512 A bytes inside the T force copies of 384 A bytes, the force rows clear of copy 0).
__device__ __forceinline__ void cta_load_rows_bulk(const SysDev& s, const PropArgs& a, unsigned char* smem, const Layout& L, int sw, long long p0, int count,
                                                   unsigned long long* bar, unsigned parity) {
    const int A = s.n_atoms, tid = (int)threadIdx.x, nt = (int)blockDim.x;
    const unsigned row_bytes = 16u * (unsigned)A;
    const size_t arr = (size_t)kR * row_bytes;
    float* X = (float*)(smem + L.x);
    float* V = (float*)(smem + L.v);
    float* Fc = (float*)(smem + L.f);
    long long* SRC = (long long*)(smem + L.src);
    const float4* SP = (const float4*)(smem + L.f);
    const float4* SV = (const float4*)(smem + L.f + arr);
    const float4* SF = (const float4*)(smem + L.f + 2 * arr);
    const bool with_vel = a.vel_in != nullptr, with_f = a.fs_in != nullptr;
    const unsigned b = smem_addr(bar);
    if (tid == 0) {
        const unsigned total = (unsigned)count * row_bytes * (1u + (with_vel ? 1u : 0u) + (with_f ? 1u : 0u));
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(total) : "memory");
    }
    if (tid < count) {
        const long long src = a.anc ? (long long)a.anc[p0 + tid] : p0 + tid;
        SRC[tid] = src;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the staging area was last touched through the generic proxy (force copies)
        const unsigned dst = smem_addr(smem + L.f) + (unsigned)tid * row_bytes;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                     "l"(row_ptr(a, a.pos_in, a.peer_pos, src, A)), "r"(row_bytes), "r"(b) : "memory");
        if (with_vel)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + (unsigned)arr),
                         "l"(row_ptr(a, a.vel_in, a.peer_vel, src, A)), "r"(row_bytes), "r"(b) : "memory");
        if (with_f)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + 2u * (unsigned)arr),
                         "l"(row_ptr(a, a.fs_in, a.peer_fs, src, A)), "r"(row_bytes), "r"(b) : "memory");
    }
    wait_tables(bar, parity);
    // transpose, replica index fastest: conflict-free 16-byte reads (row stride 16 A bytes, A odd or not) and conflict-free column writes
    for (int idx = tid; idx < count * A; idx += nt) {
        const int at = idx / count, r = idx - at * count;
        const int c = CDW_BASE_R(at, r);
        const float4 q = SP[r * A + at];
        X[c] = q.x; X[c + kR] = q.y; X[c + 2 * kR] = q.z;
        if (with_vel) {
            const float4 w = SV[r * A + at];
            V[c] = w.x; V[c + kR] = w.y; V[c + 2 * kR] = w.z;
        } else {
            V[c] = 0.f; V[c + kR] = 0.f; V[c + 2 * kR] = 0.f;
        }
    }
    __syncthreads();  // the position / velocity staging (which overlaps force copy 0) is dead
    if (with_f) {
        for (int idx = tid; idx < count * A; idx += nt) {
            const int at = idx / count, r = idx - at * count;
            const int c = CDW_BASE_R(at, r);
            const float4 g = SF[r * A + at];
            Fc[c] = g.x; Fc[c + kR] = g.y; Fc[c + 2 * kR] = g.z;
        }
    }
    __syncthreads();
}

// Two register budgets: teams of <= 4 lanes run 128-thread CTAs, three of which must be co-resident for a batch of
// 10^4 replicas to run as one wave (<= 170 registers); teams of 8 run 256-thread CTAs (<= 128 registers, two per SM).
#ifndef CDW_MINBLOCKS_SMALL
#define CDW_MINBLOCKS_SMALL 3
#endif
#ifndef CDW_MINBLOCKS_DENSE
#define CDW_MINBLOCKS_DENSE 4
#endif
template <int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) smc_propagate_kernel(SysDev s, PropArgs a, Shape sh) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ unsigned long long cta_min[MAXT / 32];
    __shared__ __align__(8) unsigned long long table_bar;
    const int T_ = sh.T, sw = 32 / T_;
    const Layout L = make_layout(s, true, kR, T_, sh.metro != 0);
    const Tables T = table_pointers(L, smem);
    bool tables_in_flight = false;
    if (a.tables) {
        if (threadIdx.x == 0) {
            init_table_barrier(&table_bar);
            bulk_load_tables(smem, a.tables, (unsigned)L.x, &table_bar);
        }
        tables_in_flight = true;
        __syncthreads();  // the barrier is initialised before anyone polls it
    } else {
        build_tables(s, T, a.lambda, a.kT, true);
    }
    LaneMem M = lane_pointers(L, smem, T_, threadIdx.x);
    unsigned long long my_min = ~0ull;
    __shared__ __align__(8) unsigned long long row_bar;
    const bool bulk_rows = sh.bulk != 0;
    unsigned row_parity = 0;
    if (bulk_rows) {
        if (threadIdx.x == 0) init_table_barrier(&row_bar);
        __syncthreads();
    }
    for (long long p0 = (long long)blockIdx.x * kR; p0 < a.n; p0 += (long long)gridDim.x * kR) {
        const long long left = a.n - p0;
        const int count = left < kR ? (int)left : kR;
        if (bulk_rows) { cta_load_rows_bulk(s, a, smem, L, sw, p0, count, &row_bar, row_parity & 1u); ++row_parity; }
        else cta_load_rows(s, a, smem, L, sw, p0, count, true);
        if (tables_in_flight) {
            wait_tables(&table_bar, 0);
            tables_in_flight = false;
        }
        const bool valid = M.rid < count;
        M.mask = __ballot_sync(0xffffffffu, valid);
        if (valid) {
            const unsigned long long key = a.program ? lane_step_program(s, T, M, a, p0 + M.rid) : lane_step(s, T, M, a, p0 + M.rid);
            my_min = key < my_min ? key : my_min;
        }
        __syncthreads();
        cta_store_rows(s, a, smem, L, sw, p0, count);
    }
    // one atomic per CTA for the running min of W (the log-sum-exp shift of K2)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, my_min, o);
        my_min = other < my_min ? other : my_min;
    }
    if ((threadIdx.x & 31) == 0) cta_min[threadIdx.x >> 5] = my_min;
    __syncthreads();
    if (a.wmin_key && threadIdx.x == 0) {
        unsigned long long m = cta_min[0];
        for (int w = 1; w < ((int)blockDim.x + 31) / 32; ++w) m = cta_min[w] < m ? cta_min[w] : m;
        if (m != ~0ull) atomicMin(a.wmin_key, m);
    }
}

// The look-ahead iteration (cdw_lane.cuh: lane_ahead_main / lane_ahead_next).  One CTA per batch of 32 replicas (no persistent loop:
// CTAs retire continuously, so a small kernel on another stream — the resampling of this very iteration — finds room beside it).
// The table image of lambda_t arrives by TMA while the rows are gathered; after part A the image of lambda_{t+1} replaces it.
template <int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) smc_lookahead_kernel(SysDev s, PropArgs a, Shape sh) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ __align__(8) unsigned long long table_bar;
    const int T_ = sh.T, sw = 32 / T_;
    const Layout L = make_layout(s, true, kR, T_);
    const Tables T = table_pointers(L, smem);
    LaneMem M = lane_pointers(L, smem, T_, threadIdx.x);
    const bool have_next = a.have_next != 0;
    unsigned parity = 0;
    __shared__ __align__(8) unsigned long long row_bar;
    const bool bulk_rows = sh.bulk != 0;
    unsigned row_parity = 0;
    if (threadIdx.x == 0) {
        init_table_barrier(&table_bar);
        if (bulk_rows) init_table_barrier(&row_bar);
    }
    __syncthreads();  // the barriers are initialised before anyone polls them
    for (long long p0 = (long long)blockIdx.x * kR; p0 < a.n; p0 += (long long)gridDim.x * kR) {
        const long long left = a.n - p0;
        const int count = left < kR ? (int)left : kR;
        if (a.tables) {
            if (threadIdx.x == 0) bulk_load_tables(smem, a.tables, (unsigned)L.x, &table_bar);
        } else {
            build_tables(s, T, a.lambda, a.kT, true);
        }
        if (bulk_rows) { cta_load_rows_bulk(s, a, smem, L, sw, p0, count, &row_bar, row_parity & 1u); ++row_parity; }
        else cta_load_rows(s, a, smem, L, sw, p0, count, true);
        if (a.tables) { wait_tables(&table_bar, parity & 1u); ++parity; }
        const bool valid = M.rid < count;
        M.mask = __ballot_sync(0xffffffffu, valid);
        AheadState S;
        S.U_keep = 0.0; S.Us = 0.0; S.bad = 0;
        if (valid) S = lane_ahead_main(s, T, M, a, p0 + M.rid);
        __syncthreads();  // nobody reads the tables of lambda_t any more
        if (have_next) {
            if (a.tables_next) {
                if (threadIdx.x == 0) bulk_load_tables(smem, a.tables_next, (unsigned)L.x, &table_bar);
                wait_tables(&table_bar, parity & 1u);
                ++parity;
            } else {
                build_tables(s, T, a.lambda_next, a.kT, true);
            }
        }
        if (valid) lane_ahead_next(s, T, M, a, p0 + M.rid, S, have_next);
        __syncthreads();
        cta_store_rows_ahead(s, a, smem, L, sw, p0, count);
    }
}

// Launch shape.  T lanes per replica (power of two), kR = 32 replica columns per CTA (blockDim = 32 T <= 256, so T <= 8).
// CDW_TEAM overrides the choice (tuning).
static int default_team(const SysDev& dev) { return dev.n_atoms >= 8 ? 4 : (dev.n_atoms >= 4 ? 2 : 1); }

static Shape pick_shape(const SysDev& dev, bool forces, long long n, int sms, int pinned = 0) {
    static const char* env = getenv("CDW_TEAM");
    Shape sh;
    sh.R = 32;
    sh.metro = 0;
    sh.bulk = 0;
    if (env && atoi(env) > 0) {
        sh.T = atoi(env);
    } else if (pinned > 0) {
        sh.T = pinned;
    } else {
        // measured on B200 (tools/tune_step.py): T = 4 is the sweet spot for 13-22 atom systems (fewer lanes
        // starve the SM of resident warps, more lanes pay for force copies and idle tails); tiny systems have nothing to
        // split.  The choice depends on the SYSTEM only, never on the batch size: the summation order of energies and
        // forces follows the team size, and results must not change with how particles are sharded over GPUs.
        sh.T = default_team(dev);
    }
    if (sh.T > 8) sh.T = 8;
    while (sh.T > 1 && make_layout(dev, forces, sh.R, sh.T).total > 227 * 1024) sh.T /= 2;
    return sh;
}

// `lambda` points at a row of the registered protocol (and kT matches): use that row's prebuilt tables
static const unsigned char* registered_tables(const cdw_system* sys, const double* lambda, double kT) {
    if (!sys->proto_tables || !lambda || kT != sys->proto_kT || sys->dev.n_lambda <= 0) return nullptr;
    const ptrdiff_t off = lambda - sys->proto_lambda;
    if (lambda < sys->proto_lambda || off % sys->dev.n_lambda != 0) return nullptr;
    const ptrdiff_t row = off / sys->dev.n_lambda;
    if (row >= sys->proto_rows) return nullptr;
    return sys->proto_tables + (size_t)row * sys->proto_stride;
}

enum LaunchKind { LAUNCH_ENERGY = 0, LAUNCH_PROPAGATE = 1, LAUNCH_LOOKAHEAD = 2 };

static int launch(const cdw_system* sys, PropArgs& a, LaunchKind kind, cudaStream_t st) {
    static const bool no_tables = getenv("CDW_NO_PREBUILT_TABLES") != nullptr;
    const bool forces = kind != LAUNCH_ENERGY, ahead = kind == LAUNCH_LOOKAHEAD;
    if (forces && !no_tables) {
        a.tables = registered_tables(sys, a.lambda, a.kT);
        if (ahead && a.lambda_next) a.tables_next = registered_tables(sys, a.lambda_next, a.kT);
    }
    int devid = 0, sms = 148;
    cudaGetDevice(&devid);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, devid);
    Shape sh = pick_shape(sys->dev, forces, a.n, sms, forces ? sys->team : 0);
    for (unsigned long long w = a.program; w & 15ull; w >>= 4) sh.metro |= (w & 15ull) == CDW_STEP_OPEN;
    static const bool diag_local = getenv("CDW_DIAG_LOCAL_ROWS") != nullptr;
    if (diag_local && a.peer_pos && a.n_per_rank > 0) a.diag_local_rank = (int)(a.gid0 / a.n_per_rank) + 1;
    static const char* bulk_env = getenv("CDW_BULK_ROWS");   // 0: never, 1: always (where possible), unset: rows that may live on a peer GPU
    const bool bulk_ok = forces && sh.T >= 4;
    sh.bulk = bulk_ok && (bulk_env ? atoi(bulk_env) != 0 : a.peer_pos != nullptr) ? 1 : 0;
    const int nthr = sh.R * sh.T;
    const Layout L = make_layout(sys->dev, forces, sh.R, sh.T, sh.metro != 0);
    const bool small = nthr <= 128;
    // 128-thread CTAs come in two register budgets: <= 170 registers (three CTAs per SM: enough for one wave of a
    // 10^4 batch, fewer spills) or <= 128 registers (four CTAs per SM) when shared memory leaves room for a fourth CTA
    // and the batch is large enough to use it (measured: 13-atom system, 10^5 replicas: 333 us vs 375 us)
    // (64-thread CTAs — teams of 2 — are limited by shared memory long before the 142-register variant is: they keep it.  Measured,
    // fluorobenzene hybrid, 1e5 replicas, T = 2: 271 us per iteration with 142 registers, 295 us with 96.)
    const bool dense = nthr == 128 && L.total * CDW_MINBLOCKS_DENSE <= 227 * 1024 && (a.n + kR - 1) / kR >= (long long)sms * 4;
    const void* k = !forces ? (const void*)reduced_potential_kernel
                  : ahead ? (!small ? (const void*)smc_lookahead_kernel<256, 2>
                                    : (dense ? (const void*)smc_lookahead_kernel<128, CDW_MINBLOCKS_DENSE> : (const void*)smc_lookahead_kernel<128, CDW_MINBLOCKS_SMALL>))
                          : (!small ? (const void*)smc_propagate_kernel<256, 2>
                                    : (dense ? (const void*)smc_propagate_kernel<128, 4> : (const void*)smc_propagate_kernel<128, CDW_MINBLOCKS_SMALL>));
    if (L.total > 227 * 1024) {
        set_error("system needs %zu bytes of shared memory per CTA (> 227 KB)", L.total);
        return CDW_EUNSUPPORTED;
    }
    CDW_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
    if (a.n == 0) return CDW_OK;
    int per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, nthr, L.total) != cudaSuccess || per_sm < 1) per_sm = 1;
    const long long need = (a.n + sh.R - 1) / sh.R, cap = (long long)sms * per_sm;
    // the look-ahead kernel runs one CTA per batch (CTAs retire continuously: the resampling kernel of the same iteration, on another
    // stream, finds room beside it); the others are persistent
    static const bool ahead_persistent = getenv("CDW_AHEAD_PERSISTENT") != nullptr;   // A/B switch
    int grid = (int)(ahead && !ahead_persistent ? (need < (1ll << 30) ? need : (1ll << 30)) : (need < cap ? need : cap));
    void* args[] = {(void*)&sys->dev, (void*)&a, (void*)&sh};
    CDW_CUDA(cudaLaunchKernel(k, dim3(grid), dim3(nthr), args, L.total, st));
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
    return launch(sys, a, LAUNCH_ENERGY, (cudaStream_t)stream);
}

extern "C" int cdw_system_register_protocol(cdw_system* sys, const double* lambda_rows, int32_t n_rows, double kT, cdw_stream stream) {
    CDW_REQUIRE(sys && lambda_rows && n_rows > 0 && kT > 0.0, "null argument");
    CDW_REQUIRE(sys->dev.n_lambda > 0, "the system has no global parameters");
    const Layout L = make_layout(sys->dev, true, kR, 1);
    CDW_REQUIRE(L.x <= 200 * 1024, "table image too large");
    if (sys->proto_tables && sys->proto_lambda == lambda_rows && sys->proto_rows == n_rows && sys->proto_kT == kT && sys->proto_stride == L.x) {
        // same protocol array again: rebuild the images in place (the pointers captured graphs hold stay valid)
        build_tables_kernel<<<n_rows, 128, L.x, (cudaStream_t)stream>>>(sys->dev, lambda_rows, kT, sys->proto_tables, L.x);
        CDW_CUDA(cudaGetLastError());
        return CDW_OK;
    }
    cudaFree(sys->proto_tables);
    sys->proto_tables = nullptr; sys->proto_lambda = nullptr; sys->proto_rows = 0;
    unsigned char* buf = nullptr;
    CDW_CUDA(cudaMalloc(&buf, L.x * (size_t)n_rows));
    cudaFuncSetAttribute(build_tables_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.x);
    build_tables_kernel<<<n_rows, 128, L.x, (cudaStream_t)stream>>>(sys->dev, lambda_rows, kT, buf, L.x);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        cudaFree(buf);
        set_error("build_tables_kernel: %s", cudaGetErrorString(e));
        return CDW_ECUDA;
    }
    sys->proto_tables = buf; sys->proto_stride = L.x; sys->proto_lambda = lambda_rows; sys->proto_rows = n_rows; sys->proto_kT = kT;
    return CDW_OK;
}

// Lanes per replica.  The summation order of energies and forces follows the team size, so results are bitwise reproducible only
// between runs that use the same one: the library's own choice depends on the system alone (never on the batch size or on how a
// batch is sharded); a caller that knows the size of the WHOLE ensemble can pin a better one for large ensembles.
extern "C" int cdw_system_set_team(cdw_system* sys, int32_t lanes) {
    CDW_REQUIRE(sys && (lanes == 0 || lanes == 1 || lanes == 2 || lanes == 4 || lanes == 8), "lanes must be 0 (automatic), 1, 2, 4 or 8");
    if (lanes > 0) CDW_REQUIRE(make_layout(sys->dev, true, kR, lanes).total <= 227 * 1024, "this team size does not fit shared memory");
    sys->team = lanes;
    return CDW_OK;
}

extern "C" int cdw_system_team(const cdw_system* sys) { return sys ? (sys->team > 0 ? sys->team : default_team(sys->dev)) : CDW_EINVAL; }

// shared memory per CTA (32 replicas) of the propagate kernels for a team size: what residency, hence the best team size, depends on
extern "C" size_t cdw_system_smem_bytes(const cdw_system* sys, int32_t lanes) { return sys && lanes > 0 ? make_layout(sys->dev, true, kR, lanes).total : 0; }

extern "C" int cdw_system_clear_protocol(cdw_system* sys) {
    CDW_REQUIRE(sys, "null argument");
    cudaFree(sys->proto_tables);
    sys->proto_tables = nullptr; sys->proto_lambda = nullptr; sys->proto_rows = 0; sys->proto_stride = 0;
    return CDW_OK;
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
    return launch(sys, a, LAUNCH_PROPAGATE, (cudaStream_t)stream);
}

static int apply_cache(PropArgs& a, const cdw_static_cache* c, bool sharded, bool gather) {
    if (!c) return CDW_OK;
    CDW_REQUIRE(c->force_out && c->energy_out, "cdw_static_cache: force_out and energy_out are required");
    CDW_REQUIRE(!(a.flags & CDW_FLAG_TRACK_WORKS), "cdw_static_cache cannot be combined with CDW_FLAG_TRACK_WORKS");
    const bool have_in = sharded ? (c->peer_force_in && c->peer_energy_in) : (c->force_in && c->energy_in);
    CDW_REQUIRE(have_in || a.n_steps == 0, "cdw_static_cache: seed the cache with an n_steps = 0 call before stepping");
    CDW_REQUIRE(!(gather && have_in && !sharded && c->force_in == c->force_out), "cdw_static_cache: gather-on-load needs distinct in/out rows");
    a.fs_out = (float4*)c->force_out; a.us_out = c->energy_out;
    if (have_in && a.n_steps > 0) {
        if (sharded) { a.peer_fs = (const float4* const*)c->peer_force_in; a.peer_us = c->peer_energy_in; a.fs_in = (const float4*)1; a.us_in = (const double*)1; }
        else { a.fs_in = (const float4*)c->force_in; a.us_in = c->energy_in; }
    }
    return CDW_OK;
}

// copies the dense-twist descriptor of a call into the kernel arguments (validation first)
static int apply_dense_twist(PropArgs& a, const cdw_langevin_params* prm, bool ula) {
    const cdw_dense_twist* d = prm->dense_twist;
    if (!d) return CDW_OK;
    CDW_REQUIRE(ula, "a twist applies to the Euler-Maruyama proposal only (CDW_FLAG_ULA)");
    CDW_REQUIRE(!prm->twist && !(prm->flags & CDW_FLAG_TWIST_REFERENCE_NORMALIZER), "dense_twist excludes twist and CDW_FLAG_TWIST_REFERENCE_NORMALIZER");
    CDW_REQUIRE(d->theta && d->chol && d->A && d->b && d->beta && d->scratch, "cdw_dense_twist: null array");
    a.tw_theta = d->theta; a.tw_chol = d->chol; a.tw_A = d->A; a.tw_b = d->b; a.tw_beta = d->beta;
    a.tw_logk0 = d->logk0; a.tw_cd = d->cd; a.tw_scratch = d->scratch;
    return CDW_OK;
}

extern "C" int cdw_smc_propagate(const cdw_system* sys, const float* pos_in, const float* vel_in, float* pos_out, float* vel_out, int64_t n,
                                 const int32_t* anc, const double* aux_in, const double* cum_in, const int32_t* label_in, const double* lambda,
                                 const cdw_langevin_params* prm, uint64_t seed, uint64_t step0, int64_t gid0, double* inc_out, double* w_out,
                                 double* aux_out, int32_t* label_out, double* u_first_out, double* u_last_out, double* shadow_out,
                                 double* proposal_out, int32_t* nan_flags, uint64_t* wmin_key, cdw_stream stream) {
    return cdw_smc_propagate_cached(sys, pos_in, vel_in, pos_out, vel_out, n, anc, aux_in, cum_in, label_in, lambda, prm, seed, step0, gid0, inc_out, w_out,
                                    aux_out, label_out, u_first_out, u_last_out, shadow_out, proposal_out, nan_flags, wmin_key, nullptr, stream);
}

extern "C" int cdw_smc_propagate_cached(const cdw_system* sys, const float* pos_in, const float* vel_in, float* pos_out, float* vel_out,
                                 int64_t n, const int32_t* anc, const double* aux_in, const double* cum_in, const int32_t* label_in,
                                 const double* lambda, const cdw_langevin_params* prm, uint64_t seed, uint64_t step0, int64_t gid0,
                                 double* inc_out, double* w_out, double* aux_out, int32_t* label_out, double* u_first_out,
                                 double* u_last_out, double* shadow_out, double* proposal_out, int32_t* nan_flags, uint64_t* wmin_key,
                                 const cdw_static_cache* cache, cdw_stream stream) {
    CDW_REQUIRE(sys && prm && n >= 0, "null argument");
    if (n == 0) return CDW_OK;
    CDW_REQUIRE(pos_in, "null argument");
    CDW_REQUIRE(lambda || sys->dev.n_lambda == 0, "lambda is required");
    CDW_REQUIRE(prm->n_steps >= 0 && prm->dt >= 0.0 && prm->kT > 0.0 && prm->gamma >= 0.0 && prm->constraint_tol > 0.0, "bad Langevin parameters");
    CDW_REQUIRE(!(anc && (pos_in == pos_out || (vel_in && vel_in == vel_out))), "gather-on-load needs distinct in/out buffers");
    const bool ula = (prm->flags & CDW_FLAG_ULA) != 0;  // the Euler-Maruyama proposal has no velocities: the buffers are left untouched
    CDW_REQUIRE(ula || vel_in || (prm->flags & CDW_FLAG_RANDOMIZE_VELOCITIES) || prm->n_steps == 0, "velocities required unless they are randomized");
    CDW_REQUIRE(ula || !pos_out || prm->n_steps == 0 || vel_out, "vel_out is required when positions are written");
    CDW_REQUIRE(!(ula && (prm->flags & CDW_FLAG_TRACK_WORKS)), "CDW_FLAG_TRACK_WORKS does not apply to the Euler-Maruyama proposal");
    CDW_REQUIRE(ula || !prm->twist, "a twist applies to the Euler-Maruyama proposal only (CDW_FLAG_ULA)");
    PropArgs a = {};
    a.twist = prm->twist;
    if (int rc = apply_dense_twist(a, prm, ula)) return rc;
    a.pos_in = (const float4*)pos_in; a.vel_in = ula ? nullptr : (const float4*)vel_in; a.pos_out = (float4*)pos_out;
    a.vel_out = ula ? nullptr : (float4*)vel_out;
    a.n = n; a.anc = anc; a.aux_in = aux_in; a.cum_in = cum_in; a.label_in = label_in; a.lambda = lambda;
    a.dt = prm->dt; a.gamma = prm->gamma; a.kT = prm->kT; a.tol = prm->constraint_tol; a.n_steps = prm->n_steps; a.flags = prm->flags;
    a.seed = seed; a.step0 = step0; a.gid0 = gid0;
    a.inc_out = inc_out; a.w_out = w_out; a.aux_out = aux_out; a.label_out = label_out; a.u_first_out = u_first_out; a.u_last_out = u_last_out;
    a.shadow_out = shadow_out; a.proposal_out = proposal_out; a.nan_flags = nan_flags; a.wmin_key = (unsigned long long*)wmin_key;
    if (prm->splitting) {
        CDW_REQUIRE(!ula && !cache, "a splitting program cannot be combined with the Euler-Maruyama proposal or a static-term cache");
        int depth = 0, n = 0;
        for (unsigned long long w = prm->splitting; w & 15ull; w >>= 4, ++n) {
            const unsigned tok = (unsigned)(w & 15ull);
            CDW_REQUIRE(tok >= CDW_STEP_V && tok <= CDW_STEP_CLOSE, "unknown sub-step code in cdw_langevin_params.splitting");
            depth += tok == CDW_STEP_OPEN ? 1 : (tok == CDW_STEP_CLOSE ? -1 : 0);
            CDW_REQUIRE(depth == 0 || depth == 1, "Metropolised blocks cannot be nested and must be opened before they are closed");
            // shadow-work generating steps outside the block are rejected like the reference does (openmm/integrators.py:248-276)
        }
        CDW_REQUIRE(depth == 0, "the Metropolised block is not closed");
        a.program = prm->splitting; a.trial_counts = prm->trial_counts;
    }
    if (int rc = apply_cache(a, cache, false, anc != nullptr)) return rc;
    return launch(sys, a, LAUNCH_PROPAGATE, (cudaStream_t)stream);
}

// ---- the look-ahead iteration
static int lookahead_common(const cdw_system* sys, PropArgs& a, const cdw_langevin_params* prm, const double* lambda, const double* lambda_next,
                            uint64_t seed, uint64_t step0, int64_t gid0, float* pos_out, float* vel_out, float* force_out, double* inc_next_out,
                            double* aux_out, int32_t* label_out, int32_t* nan_flags, cudaStream_t st) {
    CDW_REQUIRE(lambda || sys->dev.n_lambda == 0, "lambda is required");
    CDW_REQUIRE(prm->n_steps >= 0 && prm->dt >= 0.0 && prm->kT > 0.0 && prm->gamma >= 0.0 && prm->constraint_tol > 0.0, "bad Langevin parameters");
    CDW_REQUIRE(!(prm->flags & ~CDW_FLAG_RANDOMIZE_VELOCITIES), "cdw_smc_lookahead: BAOAB moves only (flags other than CDW_FLAG_RANDOMIZE_VELOCITIES are not supported)");
    CDW_REQUIRE(!prm->twist && !prm->dense_twist, "a twist applies to the Euler-Maruyama proposal only (cdw_smc_propagate with CDW_FLAG_ULA)");
    CDW_REQUIRE(force_out, "force_out is required");
    CDW_REQUIRE(prm->n_steps == 0 || (pos_out && vel_out), "pos_out and vel_out are required when n_steps > 0");
    CDW_REQUIRE(!inc_next_out || lambda_next || sys->dev.n_lambda == 0, "lambda_next is required with inc_next_out");
    a.lambda = lambda; a.lambda_next = lambda_next; a.have_next = inc_next_out != nullptr;
    a.dt = prm->dt; a.gamma = prm->gamma; a.kT = prm->kT; a.tol = prm->constraint_tol; a.n_steps = prm->n_steps; a.flags = prm->flags;
    a.seed = seed; a.step0 = step0; a.gid0 = gid0;
    a.pos_out = (float4*)pos_out; a.vel_out = (float4*)vel_out; a.fs_out = (float4*)force_out;
    a.inc_next_out = inc_next_out; a.aux_out = aux_out; a.label_out = label_out; a.nan_flags = nan_flags;
    return launch(sys, a, LAUNCH_LOOKAHEAD, st);
}

extern "C" int cdw_smc_lookahead(const cdw_system* sys, const float* pos_in, const float* vel_in, const float* force_in, float* pos_out, float* vel_out,
                                 float* force_out, int64_t n, const int32_t* anc, const int32_t* label_in, const double* lambda, const double* lambda_next,
                                 const cdw_langevin_params* prm, uint64_t seed, uint64_t step0, int64_t gid0, double* inc_next_out, double* aux_out,
                                 int32_t* label_out, int32_t* nan_flags, cdw_stream stream) {
    CDW_REQUIRE(sys && prm && n >= 0, "null argument");
    if (n == 0) return CDW_OK;
    CDW_REQUIRE(pos_in, "null argument");
    CDW_REQUIRE(prm->n_steps == 0 || (force_in && (vel_in || (prm->flags & CDW_FLAG_RANDOMIZE_VELOCITIES))), "force_in (and velocities, unless they are randomized) are required when n_steps > 0");
    CDW_REQUIRE(!(anc && (pos_in == pos_out || (vel_in && vel_in == vel_out) || (force_in && force_in == force_out))), "gather-on-load needs distinct in/out buffers");
    PropArgs a = {};
    a.pos_in = (const float4*)pos_in; a.vel_in = (const float4*)vel_in; a.fs_in = prm->n_steps > 0 ? (const float4*)force_in : nullptr;
    a.n = n; a.anc = anc; a.label_in = label_in;
    return lookahead_common(sys, a, prm, lambda, lambda_next, seed, step0, gid0, pos_out, vel_out, force_out, inc_next_out, aux_out, label_out, nan_flags,
                            (cudaStream_t)stream);
}

extern "C" int cdw_smc_lookahead_sharded(const cdw_system* sys, const float* const* peer_pos, const float* const* peer_vel, const float* const* peer_force,
                                         const int32_t* const* peer_label, int32_t n_ranks, int64_t n_per_rank, float* pos_out, float* vel_out,
                                         float* force_out, int64_t n, const int32_t* anc, const double* lambda, const double* lambda_next,
                                         const cdw_langevin_params* prm, uint64_t seed, uint64_t step0, int64_t gid0, double* inc_next_out,
                                         double* aux_out, int32_t* label_out, int32_t* nan_flags, cdw_stream stream) {
    CDW_REQUIRE(sys && prm && n >= 0, "null argument");
    if (n == 0) return CDW_OK;
    CDW_REQUIRE(peer_pos && peer_vel && peer_force && anc && n_ranks > 0 && n_per_rank > 0, "null argument");
    CDW_REQUIRE(prm->n_steps > 0, "the sharded look-ahead iteration needs n_steps > 0 (equip the initial sample with cdw_smc_lookahead)");
    PropArgs a = {};
    a.peer_pos = (const float4* const*)peer_pos; a.peer_vel = (const float4* const*)peer_vel; a.peer_fs = (const float4* const*)peer_force;
    a.peer_label = peer_label; a.n_per_rank = n_per_rank;
    a.vel_in = (const float4*)1; a.fs_in = (const float4*)1;  // they exist (read through the peer tables); the local pointers are never dereferenced
    a.n = n; a.anc = anc;
    return lookahead_common(sys, a, prm, lambda, lambda_next, seed, step0, gid0, pos_out, vel_out, force_out, inc_next_out, aux_out, label_out, nan_flags,
                            (cudaStream_t)stream);
}

extern "C" int cdw_smc_propagate_sharded(const cdw_system* sys, const float* const* peer_pos, const float* const* peer_vel, const double* const* peer_aux,
                                         const int32_t* const* peer_label, int32_t n_ranks, int64_t n_per_rank, float* pos_out, float* vel_out, int64_t n,
                                         const int32_t* anc, const double* cum_in, const double* lambda, const cdw_langevin_params* prm, uint64_t seed,
                                         uint64_t step0, int64_t gid0, double* inc_out, double* w_out, double* aux_out, int32_t* label_out,
                                         int32_t* nan_flags, uint64_t* wmin_key, cdw_stream stream) {
    return cdw_smc_propagate_sharded_cached(sys, peer_pos, peer_vel, peer_aux, peer_label, n_ranks, n_per_rank, pos_out, vel_out, n, anc, cum_in, lambda, prm,
                                            seed, step0, gid0, inc_out, w_out, aux_out, label_out, nan_flags, wmin_key, nullptr, stream);
}

extern "C" int cdw_smc_propagate_sharded_cached(const cdw_system* sys, const float* const* peer_pos, const float* const* peer_vel, const double* const* peer_aux,
                                         const int32_t* const* peer_label, int32_t n_ranks, int64_t n_per_rank, float* pos_out, float* vel_out, int64_t n,
                                         const int32_t* anc, const double* cum_in, const double* lambda, const cdw_langevin_params* prm, uint64_t seed,
                                         uint64_t step0, int64_t gid0, double* inc_out, double* w_out, double* aux_out, int32_t* label_out,
                                         int32_t* nan_flags, uint64_t* wmin_key, const cdw_static_cache* cache, cdw_stream stream) {
    CDW_REQUIRE(sys && prm && n >= 0, "null argument");
    if (n == 0) return CDW_OK;
    const bool ula = (prm->flags & CDW_FLAG_ULA) != 0;
    CDW_REQUIRE(peer_pos && anc && n_ranks > 0 && n_per_rank > 0 && pos_out && (ula || (peer_vel && vel_out)), "null argument");
    CDW_REQUIRE(!(ula && (prm->flags & CDW_FLAG_TRACK_WORKS)), "CDW_FLAG_TRACK_WORKS does not apply to the Euler-Maruyama proposal");
    CDW_REQUIRE(lambda || sys->dev.n_lambda == 0, "lambda is required");
    CDW_REQUIRE(prm->n_steps >= 0 && prm->dt >= 0.0 && prm->kT > 0.0 && prm->gamma >= 0.0 && prm->constraint_tol > 0.0, "bad Langevin parameters");
    PropArgs a = {};
    a.peer_pos = (const float4* const*)peer_pos; a.peer_vel = (const float4* const*)peer_vel; a.peer_aux = peer_aux; a.peer_label = peer_label;
    a.n_per_rank = n_per_rank;
    CDW_REQUIRE(ula || !prm->twist, "a twist applies to the Euler-Maruyama proposal only (CDW_FLAG_ULA)");
    a.twist = prm->twist;
    if (int rc = apply_dense_twist(a, prm, ula)) return rc;
    a.vel_in = ula ? nullptr : (const float4*)1;  // velocities exist (read through peer_vel); the local pointer itself is never dereferenced
    a.pos_out = (float4*)pos_out; a.vel_out = ula ? nullptr : (float4*)vel_out; a.n = n; a.anc = anc; a.cum_in = cum_in; a.lambda = lambda;
    a.dt = prm->dt; a.gamma = prm->gamma; a.kT = prm->kT; a.tol = prm->constraint_tol; a.n_steps = prm->n_steps; a.flags = prm->flags;
    a.seed = seed; a.step0 = step0; a.gid0 = gid0;
    a.inc_out = inc_out; a.w_out = w_out; a.aux_out = aux_out; a.label_out = label_out; a.nan_flags = nan_flags; a.wmin_key = (unsigned long long*)wmin_key;
    if (int rc = apply_cache(a, cache, true, true)) return rc;
    return launch(sys, a, LAUNCH_PROPAGATE, (cudaStream_t)stream);
}

extern "C" int cdw_baoab(const cdw_system* sys, float* pos, float* vel, int64_t n, const double* lambda, const cdw_langevin_params* prm,
                         uint64_t seed, uint64_t step0, int64_t gid0, double* u_first_out, double* u_last_out, double* shadow_out,
                         double* proposal_out, int32_t* nan_flags, cdw_stream stream) {
    // in place is safe without anc: every replica is read and written by the same warp
    return cdw_smc_propagate(sys, pos, vel, pos, vel, n, nullptr, nullptr, nullptr, nullptr, lambda, prm, seed, step0, gid0, nullptr,
                             nullptr, nullptr, nullptr, u_first_out, u_last_out, shadow_out, proposal_out, nan_flags, nullptr, stream);
}
