// harness.cpp — TEST-ONLY host build of the per-replica device algorithm (cdw_lane.cuh compiled by g++; one host
// "thread" per replica).  It lets the CPU test-suite check lambda lowering, term energies/forces, the
// scratch-slot force gather, SHAKE/RATTLE, BAOAB and the Philox/det_exp/fixed-point helpers against the fp64 oracle
// on the build host, where there is no GPU.  It is NOT part of the product: coddiwomple_b200 never loads it and has
// no CPU fallback.
#include <stdlib.h>

#include <barrier>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#define CDW_HOST_TEAMS 1

#include "../../coddiwomple_b200/csrc/cdw_pack.h"
#include "../../coddiwomple_b200/csrc/cdw_lane.cuh"

using namespace cdw;

struct hh_system {
    std::vector<unsigned char> blob;
    SysDev dev;
};

static std::string g_err;

extern "C" const char* hh_last_error() { return g_err.c_str(); }

extern "C" int hh_system_create(const cdw_system_desc* d, hh_system** out) {
    hh_system* s = new hh_system();
    int rc = pack_system(d, s->blob, s->dev, g_err);
    if (rc != CDW_OK) {
        delete s;
        return rc;
    }
    rebase_system(s->dev, s->blob.data());
    *out = s;
    return CDW_OK;
}

extern "C" int hh_system_destroy(hh_system* s) {
    delete s;
    return 0;
}

extern "C" int hh_system_info(const hh_system* s, int* n_slots, int* n_clusters, size_t* smem_energy, size_t* smem_step) {
    *n_slots = s->dev.n_bonds_s + s->dev.n_angles_s + s->dev.n_torsions_s + s->dev.n_pairs_s + s->dev.n_ext_s;  // static terms
    *n_clusters = s->dev.n_clusters;
    *smem_energy = make_layout(s->dev, false, 32, 8).total;
    *smem_step = make_layout(s->dev, true, 32, 8).total;
    return 0;
}

// ---- team emulation: the T lanes of one replica run as T host threads; CDW_TEAM_SYNC is a real barrier and the xor
// shuffle goes through a small exchange buffer (write, barrier, read partner, barrier), so the host executes the same
// team algorithm (term partition, per-lane force copies, reductions) as a warp does.
static thread_local std::barrier<>* g_team_barrier = nullptr;
static thread_local double* g_team_exchange = nullptr;
static int g_team = 1;

void cdw_host_team_sync() {
    if (g_team_barrier) g_team_barrier->arrive_and_wait();
}
double cdw_host_shfl_xor(double v, int tl, int o) {
    if (!g_team_barrier) return v;
    g_team_exchange[tl] = v;
    g_team_barrier->arrive_and_wait();
    const double r = g_team_exchange[tl ^ o];
    g_team_barrier->arrive_and_wait();
    return r;
}

extern "C" void hh_set_team(int T) { g_team = T < 1 ? 1 : T; }

static bool has_block(const PropArgs& a) {
    for (unsigned long long w = a.program; w & 15ull; w >>= 4)
        if ((w & 15ull) == CDW_STEP_OPEN) return true;
    return false;
}

static int run_team(const hh_system* s, PropArgs& a) {
    const int T_ = g_team;
    const Layout L = make_layout(s->dev, true, 1, T_, has_block(a));
    std::vector<unsigned char> smem(L.total + 64, 0);
    unsigned char* base = (unsigned char*)(((size_t)smem.data() + 15) & ~size_t(15));
    const Tables T = table_pointers(L, base);
    build_tables(s->dev, T, a.lambda, a.kT, true);
    unsigned long long mn = ~0ull;
    for (long long p = 0; p < a.n; ++p) {
        cta_load_rows(s->dev, a, base, L, 0, p, 1, true);
        std::barrier<> bar(T_);
        std::vector<double> exchange(T_, 0.0);
        std::vector<unsigned long long> keys(T_, ~0ull);
        std::vector<std::thread> lanes;
        for (int tl = 0; tl < T_; ++tl)
            lanes.emplace_back([&, tl]() {
                g_team_barrier = &bar; g_team_exchange = exchange.data();
                const LaneMem M = lane_pointers(L, base, T_, tl);
                keys[tl] = a.program ? lane_step_program(s->dev, T, M, a, p) : lane_step(s->dev, T, M, a, p);
                g_team_barrier = nullptr; g_team_exchange = nullptr;
            });
        for (auto& t : lanes) t.join();
        mn = keys[0] < mn ? keys[0] : mn;
        cta_store_rows(s->dev, a, base, L, 0, p, 1);
    }
    if (a.wmin_key && mn < *a.wmin_key) *a.wmin_key = mn;
    return 0;
}

static int run(const hh_system* s, PropArgs& a, bool forces) {
    if (forces && g_team > 1) return run_team(s, a);
    const Layout L = make_layout(s->dev, forces, 1, 1, forces && has_block(a));
    std::vector<unsigned char> smem(L.total + 64, 0);
    unsigned char* base = (unsigned char*)(((size_t)smem.data() + 15) & ~size_t(15));
    const Tables T = table_pointers(L, base);
    build_tables(s->dev, T, a.lambda, forces ? a.kT : 1.0, forces);
    const LaneMem M = lane_pointers(L, base, 1, 0);
    unsigned long long mn = ~0ull;
    for (long long p = 0; p < a.n; ++p) {
        cta_load_rows(s->dev, a, base, L, 0, p, 1, forces);
        if (forces) {
            unsigned long long k = a.program ? lane_step_program(s->dev, T, M, a, p) : lane_step(s->dev, T, M, a, p);
            mn = k < mn ? k : mn;
            cta_store_rows(s->dev, a, base, L, 0, p, 1);
        } else {
            a.energy_out[p] = a.beta * lane_energy<false>(s->dev, T, M);
        }
    }
    if (forces && a.wmin_key && mn < *a.wmin_key) *a.wmin_key = mn;
    return 0;
}

extern "C" int hh_reduced_potential(const hh_system* s, const float* pos, int64_t n, const double* lambda, double beta, double* u_out) {
    PropArgs a = {};
    a.pos_in = (const float4*)pos; a.n = n; a.lambda = lambda; a.beta = beta; a.energy_out = u_out;
    return run(s, a, false);
}

extern "C" int hh_energy_forces(const hh_system* s, const float* pos, int64_t n, const double* lambda, double* energy_out, float* force_out) {
    PropArgs a = {};
    a.pos_in = (const float4*)pos; a.n = n; a.lambda = lambda; a.kT = 1.0; a.tol = 1e-6;
    a.energy_out = energy_out; a.force_out = (float4*)force_out;
    return run(s, a, true);
}

extern "C" int hh_smc_propagate(const hh_system* s, const float* pos_in, const float* vel_in, float* pos_out, float* vel_out, int64_t n,
                                const int32_t* anc, const double* aux_in, const double* cum_in, const int32_t* label_in, const double* lambda,
                                const cdw_langevin_params* prm, uint64_t seed, uint64_t step0, int64_t gid0, double* inc_out, double* w_out,
                                double* aux_out, int32_t* label_out, double* u_first_out, double* u_last_out, double* shadow_out,
                                double* proposal_out, int32_t* nan_flags, uint64_t* wmin_key) {
    PropArgs a = {};
    a.pos_in = (const float4*)pos_in; a.vel_in = (const float4*)vel_in; a.pos_out = (float4*)pos_out; a.vel_out = (float4*)vel_out;
    a.n = n; a.anc = anc; a.aux_in = aux_in; a.cum_in = cum_in; a.label_in = label_in; a.lambda = lambda;
    a.dt = prm->dt; a.gamma = prm->gamma; a.kT = prm->kT; a.tol = prm->constraint_tol; a.n_steps = prm->n_steps; a.flags = prm->flags;
    a.seed = seed; a.step0 = step0; a.gid0 = gid0;
    a.inc_out = inc_out; a.w_out = w_out; a.aux_out = aux_out; a.label_out = label_out; a.u_first_out = u_first_out; a.u_last_out = u_last_out;
    a.shadow_out = shadow_out; a.proposal_out = proposal_out; a.nan_flags = nan_flags; a.wmin_key = (unsigned long long*)wmin_key;
    if (a.flags & CDW_FLAG_ULA) {  // as cdw_smc_propagate does
        a.vel_in = nullptr; a.vel_out = nullptr; a.twist = prm->twist;
        if (const cdw_dense_twist* d = prm->dense_twist) {
            a.tw_theta = d->theta; a.tw_chol = d->chol; a.tw_A = d->A; a.tw_b = d->b; a.tw_beta = d->beta;
            a.tw_logk0 = d->logk0; a.tw_cd = d->cd; a.tw_scratch = d->scratch;
        }
    }
    a.program = prm->splitting; a.trial_counts = prm->trial_counts;
    return run(s, a, true);
}

// the same with a static-term cache (cdw_smc_propagate_cached): fs/us rows in and out
extern "C" int hh_smc_propagate_cached(const hh_system* s, const float* pos_in, const float* vel_in, float* pos_out, float* vel_out, int64_t n,
                                       const int32_t* anc, const double* aux_in, const double* cum_in, const double* lambda,
                                       const cdw_langevin_params* prm, uint64_t seed, uint64_t step0, int64_t gid0, double* inc_out, double* w_out,
                                       double* aux_out, double* u_first_out, double* u_last_out, double* proposal_out, int32_t* nan_flags,
                                       const float* fs_in, const double* us_in, float* fs_out, double* us_out) {
    PropArgs a = {};
    a.pos_in = (const float4*)pos_in; a.vel_in = (const float4*)vel_in; a.pos_out = (float4*)pos_out; a.vel_out = (float4*)vel_out;
    a.n = n; a.anc = anc; a.aux_in = aux_in; a.cum_in = cum_in; a.lambda = lambda;
    a.dt = prm->dt; a.gamma = prm->gamma; a.kT = prm->kT; a.tol = prm->constraint_tol; a.n_steps = prm->n_steps; a.flags = prm->flags;
    a.seed = seed; a.step0 = step0; a.gid0 = gid0;
    a.inc_out = inc_out; a.w_out = w_out; a.aux_out = aux_out; a.u_first_out = u_first_out; a.u_last_out = u_last_out;
    a.proposal_out = proposal_out; a.nan_flags = nan_flags;
    if (a.flags & CDW_FLAG_ULA) { a.vel_in = nullptr; a.vel_out = nullptr; }
    a.fs_out = (float4*)fs_out; a.us_out = us_out;
    if (fs_in && a.n_steps > 0) { a.fs_in = (const float4*)fs_in; a.us_in = us_in; }
    return run(s, a, true);
}

// the look-ahead iteration (cdw_smc_lookahead): part A with the tables of lambda, part B with those of lambda_next
extern "C" int hh_smc_lookahead(const hh_system* s, const float* pos_in, const float* vel_in, const float* force_in, float* pos_out, float* vel_out,
                                float* force_out, int64_t n, const int32_t* anc, const int32_t* label_in, const double* lambda, const double* lambda_next,
                                const cdw_langevin_params* prm, uint64_t seed, uint64_t step0, int64_t gid0, double* inc_next_out, double* aux_out,
                                int32_t* label_out, int32_t* nan_flags) {
    PropArgs a = {};
    a.pos_in = (const float4*)pos_in; a.vel_in = (const float4*)vel_in; a.fs_in = prm->n_steps > 0 ? (const float4*)force_in : nullptr;
    a.pos_out = (float4*)pos_out; a.vel_out = (float4*)vel_out; a.fs_out = (float4*)force_out;
    a.n = n; a.anc = anc; a.label_in = label_in; a.lambda = lambda; a.lambda_next = lambda_next; a.have_next = inc_next_out != nullptr;
    a.dt = prm->dt; a.gamma = prm->gamma; a.kT = prm->kT; a.tol = prm->constraint_tol; a.n_steps = prm->n_steps; a.flags = prm->flags;
    a.seed = seed; a.step0 = step0; a.gid0 = gid0;
    a.inc_next_out = inc_next_out; a.aux_out = aux_out; a.label_out = label_out; a.nan_flags = nan_flags;
    const int T_ = g_team;
    const Layout L = make_layout(s->dev, true, 1, T_);
    std::vector<unsigned char> smem(L.total + 64, 0);
    unsigned char* base = (unsigned char*)(((size_t)smem.data() + 15) & ~size_t(15));
    const Tables T = table_pointers(L, base);
    for (long long p = 0; p < a.n; ++p) {
        build_tables(s->dev, T, a.lambda, a.kT, true);
        cta_load_rows(s->dev, a, base, L, 0, p, 1, true);
        std::vector<AheadState> st(T_);
        auto team = [&](auto&& body) {
            if (T_ == 1) { body(0); return; }
            std::barrier<> bar(T_);
            std::vector<double> exchange(T_, 0.0);
            std::vector<std::thread> lanes;
            for (int tl = 0; tl < T_; ++tl)
                lanes.emplace_back([&, tl]() {
                    g_team_barrier = &bar; g_team_exchange = exchange.data();
                    body(tl);
                    g_team_barrier = nullptr; g_team_exchange = nullptr;
                });
            for (auto& t : lanes) t.join();
        };
        team([&](int tl) { st[tl] = lane_ahead_main(s->dev, T, lane_pointers(L, base, T_, tl), a, p); });
        if (a.have_next) build_tables(s->dev, T, a.lambda_next, a.kT, true);
        team([&](int tl) { lane_ahead_next(s->dev, T, lane_pointers(L, base, T_, tl), a, p, st[tl], a.have_next != 0); });
        cta_store_rows_ahead(s->dev, a, base, L, 0, p, 1);
    }
    return 0;
}

extern "C" void hh_det_exp(const double* x, int64_t n, double* out) {
    for (int64_t i = 0; i < n; ++i) out[i] = det_exp(x[i]);
}

extern "C" void hh_fixed_target(const double* u, int64_t n, uint64_t total, uint64_t* out) {
    for (int64_t i = 0; i < n; ++i) out[i] = fixed_target(u[i], total);
}

extern "C" void hh_ordered_key(const double* x, int64_t n, uint64_t* key, double* back) {
    for (int64_t i = 0; i < n; ++i) {
        key[i] = ordered_key(x[i]);
        back[i] = ordered_key_inverse(key[i]);
    }
}

extern "C" void hh_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out4) {
    u4 c; c.x = c0; c.y = c1; c.z = c2; c.w = c3;
    u4 r = philox4x32_10(c, k0, k1);
    out4[0] = r.x; out4[1] = r.y; out4[2] = r.z; out4[3] = r.w;
}

extern "C" void hh_normals(uint64_t seed, uint64_t gid, uint32_t step, uint32_t atom, uint32_t kind, double* out3) {
    normals3(seed, gid, step, atom, kind, out3[0], out3[1], out3[2]);
}
