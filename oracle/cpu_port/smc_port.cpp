// smc_port.cpp — CPU BASELINE ("B2"): the whole SMC anneal of the reference's hot loop
// (coddiwomple/openmm/coddiwomple.py:312-331) compiled for the host, batched over particles, on all cores (OpenMP).
// TEST / BENCHMARK INFRASTRUCTURE ONLY — used by bench.py's cpu_baseline / --impl reference leg and by tests; the product never
// loads it (the product has no CPU path).
//
// The per-particle arithmetic is the library's own per-replica code (coddiwomple_b200/csrc/cdw_lane.cuh compiled by g++, one
// "lane" per particle: the same energy terms, BAOAB "V R O R V" (openmm/integrators.py:297-350), SHAKE / RATTLE, Philox noise),
// i.e. the GPU and the CPU arm do the same work per particle-step; the loop around it follows the reference:
//     inc = u_t(x_{t-1}) + aux            distribution_factories.py:124-129
//     propagate (n_steps BAOAB steps)     openmm/propagators.py:136-154
//     nESS, resample iff nESS <= thr      utils.py:41-72, resamplers.py:61-133 (multinomial: cdf search, resamplers.py:254-272)
//     aux = -u_t(x_t)                     openmm/coddiwomple.py:326-327
// kind: "port" (OpenMM, where the reference's arithmetic lives, cannot be installed offline).
#include <math.h>
#include <omp.h>
#include <stdlib.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../coddiwomple_b200/csrc/cdw_pack.h"
#include "../../coddiwomple_b200/csrc/cdw_lane.cuh"

using namespace cdw;

struct port_system {
    std::vector<unsigned char> blob;
    SysDev dev;
};

static std::string g_err;

extern "C" const char* port_last_error() { return g_err.c_str(); }

extern "C" int port_system_create(const cdw_system_desc* d, port_system** out) {
    port_system* s = new port_system();
    int rc = pack_system(d, s->blob, s->dev, g_err);
    if (rc != CDW_OK) {
        delete s;
        return rc;
    }
    rebase_system(s->dev, s->blob.data());
    *out = s;
    return CDW_OK;
}

extern "C" int port_system_destroy(port_system* s) {
    delete s;
    return 0;
}

extern "C" int port_threads() { return omp_get_max_threads(); }
// torchrun exports OMP_NUM_THREADS=1 to every rank; the baseline is meant to use every host core it is allowed to
extern "C" void port_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }

static double uniform53(unsigned long long seed, unsigned long long stream_id, long long i) {
    u4 c;
    const long long pair = i >> 1;
    c.x = (uint32_t)pair; c.y = (uint32_t)((unsigned long long)pair >> 32); c.z = (uint32_t)stream_id; c.w = 0x5EEDu;
    const u4 r = philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    const unsigned long long bits = (i & 1) ? ((((unsigned long long)r.z << 32) | r.w) >> 11) : ((((unsigned long long)r.x << 32) | r.y) >> 11);
    return (double)bits * 1.1102230246251565e-16;
}

// One anneal.  pos: [P][A][4] float rows (in: samples of the first target; out: final particles), lambda_rows: [T][n_lambda].
// Returns 0; cum_out[P] = cumulative works, stats_out[T][2] = (nESS, resampled) per iteration.
extern "C" int port_anneal(const port_system* s, float* pos, int64_t P, const double* lambda_rows, int32_t T, const cdw_langevin_params* prm,
                           double floor_threshold, uint64_t seed, uint64_t resample_seed, double* cum_out, double* stats_out) {
    const int A = s->dev.n_atoms, nl = s->dev.n_lambda;
    std::vector<float> vel((size_t)P * A * 4, 0.f), pos2((size_t)P * A * 4), vel2((size_t)P * A * 4), frc((size_t)P * A * 4), frc2((size_t)P * A * 4);
    std::vector<double> aux(P), aux2(P), inc(P), inc_next(P), cum(P, 0.0), W(P), cdf(P);
    std::vector<int> anc(P), label(P), label2(P);
    for (int64_t p = 0; p < P; ++p) { anc[p] = (int)p; label[p] = (int)p; }
    float* px = pos; float* pv = vel.data(); float* pf = frc.data();
    float* qx = pos2.data(); float* qv = vel2.data(); float* qf = frc2.data();
    const Layout L = make_layout(s->dev, true, 1, 1);
    auto launch = [&](int t, int n_steps, unsigned flags, bool gather, const double* lam, const double* lam_next) {
#pragma omp parallel
        {
            std::vector<unsigned char> smem(L.total + 64, 0);
            unsigned char* base = (unsigned char*)(((size_t)smem.data() + 15) & ~size_t(15));
            const Tables Tb = table_pointers(L, base);
            std::vector<unsigned char> img_t(L.x), img_n(L.x);
            PropArgs a = {};
            a.pos_in = (const float4*)px; a.vel_in = (const float4*)pv; a.fs_in = n_steps > 0 ? (const float4*)pf : nullptr;
            a.pos_out = n_steps > 0 ? (float4*)qx : nullptr; a.vel_out = n_steps > 0 ? (float4*)qv : nullptr; a.fs_out = n_steps > 0 ? (float4*)qf : (float4*)pf;
            a.n = P; a.anc = gather ? anc.data() : nullptr; a.label_in = label.data(); a.lambda = lam; a.lambda_next = lam_next; a.have_next = lam_next != nullptr;
            a.dt = prm->dt; a.gamma = prm->gamma; a.kT = prm->kT; a.tol = prm->constraint_tol; a.n_steps = n_steps; a.flags = flags;
            a.seed = seed; a.step0 = (unsigned long long)t * (unsigned long long)std::max(prm->n_steps, 1); a.gid0 = 0;
            a.inc_next_out = lam_next ? inc_next.data() : nullptr; a.aux_out = n_steps > 0 ? aux2.data() : aux.data();
            a.label_out = n_steps > 0 ? label2.data() : nullptr;
            // the two table images are built once per thread and launch, then copied in per particle (what the TMA copy does on the GPU)
            build_tables(s->dev, Tb, lam, a.kT, true);
            memcpy(img_t.data(), base, L.x);
            if (lam_next) { build_tables(s->dev, Tb, lam_next, a.kT, true); memcpy(img_n.data(), base, L.x); }
            const LaneMem M = lane_pointers(L, base, 1, 0);
#pragma omp for schedule(dynamic, 64)
            for (int64_t p = 0; p < P; ++p) {
                memcpy(base, img_t.data(), L.x);
                cta_load_rows(s->dev, a, base, L, 0, p, 1, true);
                const AheadState st = lane_ahead_main(s->dev, Tb, M, a, p);
                if (lam_next) memcpy(base, img_n.data(), L.x);
                lane_ahead_next(s->dev, Tb, M, a, p, st, lam_next != nullptr);
                cta_store_rows_ahead(s->dev, a, base, L, 0, p, 1);
            }
        }
    };
    // equip the initial sample: aux = -u_0(x_0), forces and incremental work for iteration 1
    launch(0, 0, 0, false, lambda_rows, T > 1 ? lambda_rows + nl : nullptr);
    for (int t = 1; t < T; ++t) {
        // weights of iteration t (the works the previous launch left behind, through the previous ancestors)
        double wmin = INFINITY;
        for (int64_t p = 0; p < P; ++p) { inc[p] = inc_next[anc[p]]; W[p] = cum[p] + inc[p]; wmin = std::min(wmin, W[p]); }
        double se = 0.0, se2 = 0.0;
#pragma omp parallel for reduction(+ : se, se2)
        for (int64_t p = 0; p < P; ++p) { const double e = exp(wmin - W[p]); se += e; se2 += e * e; cdf[p] = e; }
        const double ness = se * se / ((double)P * se2);
        const bool resample = floor_threshold < 0.0 || ness <= floor_threshold;
        stats_out[2 * t] = ness; stats_out[2 * t + 1] = resample ? 1.0 : 0.0;
        // propagate with the ancestors of iteration t - 1 (gather-on-load), forces and works for iteration t + 1
        launch(t, prm->n_steps, t == 1 ? CDW_FLAG_RANDOMIZE_VELOCITIES : 0u, t > 1, lambda_rows + (size_t)t * nl, t + 1 < T ? lambda_rows + (size_t)(t + 1) * nl : nullptr);
        std::swap(px, qx); std::swap(pv, qv); std::swap(pf, qf); aux.swap(aux2); label.swap(label2);
        if (resample) {
            const double mean = wmin - log(se) + log((double)P);
            double run = 0.0;
            for (int64_t p = 0; p < P; ++p) { run += cdf[p]; cdf[p] = run; }
#pragma omp parallel for
            for (int64_t p = 0; p < P; ++p) {
                const double target = uniform53(resample_seed, (unsigned long long)t, p) * run;
                anc[p] = (int)std::min<int64_t>(std::upper_bound(cdf.begin(), cdf.end(), target) - cdf.begin(), P - 1);
                cum[p] = cum[p] + (mean - cum[p]);
            }
        } else {
            for (int64_t p = 0; p < P; ++p) { anc[p] = (int)p; cum[p] = W[p]; }
        }
    }
    // apply the last ancestors
    if (T > 1) {
        for (int64_t p = 0; p < P; ++p) memcpy(qx + (size_t)p * A * 4, px + (size_t)anc[p] * A * 4, sizeof(float) * A * 4);
        std::swap(px, qx);
    }
    if (px != pos) memcpy(pos, px, sizeof(float) * (size_t)P * A * 4);
    for (int64_t p = 0; p < P; ++p) cum_out[p] = cum[p];
    return 0;
}
