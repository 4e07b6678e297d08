// cdw_lane.cuh — what a TEAM of T lanes (T = 1, 2, 4, ..., 32; same warp) does for ONE replica.
//
// Why this mapping (profiles/prof_propagate_r01_warp_per_replica_details.csv): with one warp per
// replica the profiler showed 3-29 active lanes per instruction (constraints 3.2, force gather 10.5, angles 18,
// torsions 23.5, pairs 29) and ~15 000 warp-instructions per replica-iteration, two thirds of them bookkeeping
// (scratch-slot stores, CSR gather, warp syncs, lane loops).  With one thread per replica every lane of a warp
// executes the SAME term of 32 different replicas: 32/32 lanes active in every phase, term parameters are
// warp-uniform broadcast loads from shared memory, forces are accumulated by plain read-modify-write on the
// thread's own column (no atomics, no scratch, no gather, no __syncwarp), and constraint clusters are solved
// sequentially by the thread that owns the replica.
//
// One thread per replica needs ~12 B of shared memory per degree of freedom per thread, which caps residency at a
// few warps per SM for 20-atom systems, and a batch of 10^4 replicas is only 313 warps.  So the code is written for
// a TEAM of T lanes per replica: the T lanes split the term lists, atoms and constraint clusters of their replica
// round-robin, each accumulates forces into its own fp32 copy (summed after the evaluation; still no atomics), and
// energies are combined with log2(T) shuffles.  T = 1 is the pure lane-per-replica case (and the host harness);
// the launcher picks T from the system (atom count, shared-memory footprint), never from the batch size, so that
// results do not depend on how a batch is sharded.
//
// Layout: the CTA stages the lambda-resolved term tables once (shared by all its threads); every per-replica
// array (x, v, f, r0) is stored [element][thread] in shared memory, so a warp touching element k of its 32 replicas
// hits 32 consecutive banks.  Global rows (float4 per atom) are moved by the whole CTA with coalesced loads and
// transposed into that layout; the resampling gather is that load (row index = ancestor).
//
// Numerical contract: particle state (x, v) is fp32 wherever it is stored, on chip and off; all
// arithmetic on it (energies, forces, constraints, integrator) is fp64; forces are accumulated in fp32.
//
// Reference semantics: OMMLI "V R O R V" (coddiwomple/openmm/integrators.py:297-350) driven by OMMBIP.apply
// (coddiwomple/openmm/propagators.py:127-219); reduced potential (coddiwomple/openmm/states.py:116-136);
// incremental work (coddiwomple/distribution_factories.py:124-129).
#pragma once
#include "cdw_math.cuh"

#if defined(__CUDACC__)
#define CDW_NOINLINE_HD __host__ __device__ __noinline__
#else
#define CDW_NOINLINE_HD
#endif
#if defined(__CUDA_ARCH__)
#define CDW_CTA_THREADS(tid, nt) for (int tid = (int)threadIdx.x, nt = (int)blockDim.x, _cdw_once = 1; _cdw_once; _cdw_once = 0, __syncthreads())
#define CDW_LDG(p) __ldg(p)
#define CDW_TEAM_SYNC(M) __syncwarp((M).mask)
#define CDW_SHFL_XOR(M, v, o) __shfl_xor_sync((M).mask, (v), (o))
#else
#define CDW_CTA_THREADS(tid, nt) for (int tid = 0, nt = 1; tid < 1; ++tid)
#define CDW_LDG(p) (*(p))
#if defined(CDW_HOST_TEAMS)
// host harness with teams: the T lanes of a replica are T host threads; the harness supplies a barrier and an exchange
void cdw_host_team_sync();
double cdw_host_shfl_xor(double v, int tl, int o);
#define CDW_TEAM_SYNC(M) cdw_host_team_sync()
#define CDW_SHFL_XOR(M, v, o) cdw_host_shfl_xor((v), (M).tl, (o))
#else
#define CDW_TEAM_SYNC(M) ((void)0)
#define CDW_SHFL_XOR(M, v, o) (v)
#endif
#endif
#if !defined(__CUDACC__)
struct alignas(16) float4 {
    float x, y, z, w;
};
inline float4 make_float4(float x, float y, float z, float w) { float4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
#endif

namespace cdw {

constexpr int kMaxConstraintIters = 500;

// ------------------------------------------------------------------------------------------------ shared-memory layout
// kR = replicas staged per CTA (blockDim = kR * T).  Component c of atom a of replica column r lives at
// arr[(3 a + c) * kR + ((r + a * sw) & (kR - 1))], sw = 32 / T: the per-atom rotation makes the T lanes of a team,
// which touch T different atoms of the SAME column, hit T different banks, without padding; the three components
// of an atom are kR words apart (immediate offsets).  r0 and the force copies use the same rule with the
// constraint / atom index.
#if defined(__CUDA_ARCH__)
constexpr int kR = 32;
#else
constexpr int kR = 1;  // host harness: one replica at a time
#endif
struct Layout {
    size_t lam, nbeff, inv_m, sig_v, bonds, angles, torsions, pairs, exts, cl_start, cons_atoms, cons_d2, cl_coef_start, cl_coef;  // CTA-shared tables
    size_t x, v, f, r0, src, bad;                                                                                                  // [element][replica]
    size_t xo, vo;                                                                                                                 // Metropolised programs: state at "{"
    size_t total;
};

CDW_HD size_t align16(size_t v) { return (v + 15) & ~size_t(15); }

CDW_HD Layout make_layout(const SysDev& s, bool forces, int R, int T, bool metro = false) {  // R: kR on the device, 1 on the host
    Layout L;
    size_t o = 0;
    const int A = s.n_atoms;
    L.lam = o; o = align16(o + sizeof(double) * (s.n_lambda > 0 ? s.n_lambda : 1));
    L.nbeff = o; o = align16(o + sizeof(double) * 3 * A);
    L.inv_m = o; o = align16(o + sizeof(double) * A);
    L.sig_v = o; o = align16(o + sizeof(double) * A);
    L.bonds = o; o = align16(o + sizeof(EffBond) * s.n_bonds);
    L.angles = o; o = align16(o + sizeof(EffAngle) * s.n_angles);
    L.torsions = o; o = align16(o + sizeof(EffTorsion) * s.n_torsions);
    L.pairs = o; o = align16(o + sizeof(EffPair) * s.n_pairs);
    L.exts = o; o = align16(o + sizeof(EffExt) * s.n_ext);
    L.cl_start = o; o = align16(o + sizeof(int) * (forces ? s.n_clusters + 1 : 0));
    L.cons_atoms = o; o = align16(o + sizeof(int) * (forces ? 2 * s.n_cons : 0));
    L.cons_d2 = o; o = align16(o + sizeof(double) * (forces ? s.n_cons : 0));
    L.cl_coef_start = o; o = align16(o + sizeof(int) * (forces ? s.n_clusters + 1 : 0));
    L.cl_coef = o; o = align16(o + sizeof(double) * (forces ? s.n_coef : 0));
    const size_t col = sizeof(float) * R;
    L.x = o; o = align16(o + col * 3 * A);
    L.v = L.f = L.r0 = L.src = L.bad = o;
    if (forces) {
        L.v = o; o = align16(o + col * 3 * A);
        L.f = o; o = align16(o + col * 3 * A * T);  // one force copy per team lane
        L.r0 = o; o = align16(o + col * 3 * (s.n_cons > 0 ? s.n_cons : 0));
        L.src = o; o = align16(o + sizeof(long long) * R);
        L.bad = o; o = align16(o + sizeof(int) * R);
    }
    L.xo = L.vo = o;
    if (forces && metro) {
        L.xo = o; o = align16(o + col * 3 * A);
        L.vo = o; o = align16(o + col * 3 * A);
    }
    L.total = o;
    return L;
}

struct Tables {  // pointers into the CTA-shared tables
    double* lam; double* nbeff; double* inv_m; double* sig_v;
    EffBond* bonds; EffAngle* angles; EffTorsion* torsions; EffPair* pairs; EffExt* exts;
    int* cl_start; int* cons_atoms; double* cons_d2; int* cl_coef_start; double* cl_coef;
};

struct LaneMem {  // this thread's view of its replica's columns
    float* x; float* v; float* f; float* r0;
    float* xo; float* vo;  // Metropolised programs only
    long long* src; int* bad;
    int rid;           // this replica's column (0 .. kR-1)
    int T, tl, sw;     // team size, lane within the team, bank rotation (32 / T, 0 on the host)
    unsigned mask;     // lanes of this warp whose replica is valid
};

CDW_HD Tables table_pointers(const Layout& L, unsigned char* smem) {
    Tables T;
    T.lam = (double*)(smem + L.lam); T.nbeff = (double*)(smem + L.nbeff); T.inv_m = (double*)(smem + L.inv_m); T.sig_v = (double*)(smem + L.sig_v);
    T.bonds = (EffBond*)(smem + L.bonds); T.angles = (EffAngle*)(smem + L.angles); T.torsions = (EffTorsion*)(smem + L.torsions);
    T.pairs = (EffPair*)(smem + L.pairs); T.exts = (EffExt*)(smem + L.exts);
    T.cl_start = (int*)(smem + L.cl_start); T.cons_atoms = (int*)(smem + L.cons_atoms); T.cons_d2 = (double*)(smem + L.cons_d2);
    T.cl_coef_start = (int*)(smem + L.cl_coef_start); T.cl_coef = (double*)(smem + L.cl_coef);
    return T;
}

CDW_HD LaneMem lane_pointers(const Layout& L, unsigned char* smem, int T, int tid) {
    LaneMem M;
    M.x = (float*)(smem + L.x); M.v = (float*)(smem + L.v); M.f = (float*)(smem + L.f); M.r0 = (float*)(smem + L.r0);
    M.src = (long long*)(smem + L.src); M.bad = (int*)(smem + L.bad);
    M.xo = (float*)(smem + L.xo); M.vo = (float*)(smem + L.vo);
    M.T = T; M.rid = tid / T; M.tl = tid - M.rid * T; M.sw = kR > 1 ? 32 / T : 0;
    M.mask = 0xffffffffu;
    return M;
}

// word offset of group g (an atom or a constraint: three consecutive components) in column r (see Layout)
#define CDW_BASE_R(g, r) ((g) * (3 * kR) + (((r) + (g) * sw) & (kR - 1)))
#define CDW_BASE(g) CDW_BASE_R(g, rid)
#define CDW_UNPACK(M) const int rid = (M).rid, T_ = (M).T, tl = (M).tl, sw = (M).sw; (void)T_; (void)tl; (void)rid; (void)sw

// true if the predicate holds in ANY valid lane of the warp (host: in this lane).  Control flow that leads to a team barrier
// (CDW_TEAM_SYNC synchronises all valid lanes of the warp) has to be decided with this, never per team.
CDW_HD bool lane_any(const LaneMem& M, bool pred) {
#if defined(__CUDA_ARCH__)
    return __any_sync(M.mask, pred) != 0;
#else
    (void)M;
    return pred;
#endif
}

CDW_HD double team_sum(const LaneMem& M, double v) {
    for (int o = M.T >> 1; o > 0; o >>= 1) v += CDW_SHFL_XOR(M, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------------ prologue (once per CTA)
// Resolve the lambda vector into effective term records: the per-replica code never touches lambda again.
CDW_NOINLINE_HD inline void build_tables(const SysDev& s, const Tables& T, const double* lam_g, double kT, bool forces) {
    CDW_CTA_THREADS(tid, nt) {
        for (int q = tid; q < s.n_lambda; q += nt) T.lam[q] = lam_g[q];
    }
    const double* lam = T.lam;
    CDW_CTA_THREADS(tid, nt) {
        // per-atom effective NonbondedForce parameters (ParticleOffsets)
        for (int a = tid; a < s.n_atoms; a += nt) {
            double q = s.nb_atom_par[3 * a], sg = s.nb_atom_par[3 * a + 1], ep = s.nb_atom_par[3 * a + 2];
            for (int o = s.atom_off_start[a]; o < s.atom_off_start[a + 1]; ++o) {
                double l = lam[s.atom_off_slot[o]];
                q += l * s.atom_off_par[3 * o]; sg += l * s.atom_off_par[3 * o + 1]; ep += l * s.atom_off_par[3 * o + 2];
            }
            T.nbeff[3 * a] = q; T.nbeff[3 * a + 1] = sg; T.nbeff[3 * a + 2] = ep;
            T.inv_m[a] = s.inv_mass[a];
            T.sig_v[a] = sqrt(kT * s.inv_mass[a]);
        }
        for (int q = tid; q < s.n_bonds; q += nt) {
            EffBond b;
            b.i = s.bond_atoms[2 * q]; b.j = s.bond_atoms[2 * q + 1];
            b.K = affine(s.bond_par[4 * q], s.bond_par[4 * q + 1], s.bond_slot[q], lam);
            b.L = affine(s.bond_par[4 * q + 2], s.bond_par[4 * q + 3], s.bond_slot[q], lam);
            T.bonds[q] = b;
        }
        for (int q = tid; q < s.n_angles; q += nt) {
            EffAngle a;
            a.i = s.angle_atoms[3 * q]; a.j = s.angle_atoms[3 * q + 1]; a.k = s.angle_atoms[3 * q + 2]; a.pad = 0;
            a.K = affine(s.angle_par[4 * q], s.angle_par[4 * q + 1], s.angle_slot[q], lam);
            a.T0 = affine(s.angle_par[4 * q + 2], s.angle_par[4 * q + 3], s.angle_slot[q], lam);
            T.angles[q] = a;
        }
        for (int q = tid; q < s.n_torsions; q += nt) {
            EffTorsion t;
            t.i = s.torsion_atoms[4 * q]; t.j = s.torsion_atoms[4 * q + 1]; t.k = s.torsion_atoms[4 * q + 2]; t.l = s.torsion_atoms[4 * q + 3];
            t.kk = affine(s.torsion_par[3 * q], s.torsion_par[3 * q + 1], s.torsion_slot[q], lam);
            double ph = s.torsion_par[3 * q + 2];
            t.c0 = cos(ph); t.s0 = sin(ph);
            t.n = s.torsion_n[q]; t.pad = 0;
            T.torsions[q] = t;
        }
        for (int q = tid; q < s.n_ext; q += nt) {
            EffExt e;
            e.atom = s.ext_atom[q]; e.pad = 0;
            e.K = affine(s.ext_par[6 * q], s.ext_par[6 * q + 1], s.ext_slot[3 * q], lam);
            e.x0 = affine(s.ext_par[6 * q + 2], s.ext_par[6 * q + 3], s.ext_slot[3 * q + 1], lam);
            e.U0 = affine(s.ext_par[6 * q + 4], s.ext_par[6 * q + 5], s.ext_slot[3 * q + 2], lam);
            T.exts[q] = e;
        }
        if (forces) {
            for (int q = tid; q < s.n_clusters + 1; q += nt) T.cl_start[q] = s.cluster_start[q];
            for (int q = tid; q < 2 * s.n_cons; q += nt) T.cons_atoms[q] = s.cons_atoms[q];
            for (int q = tid; q < s.n_cons; q += nt) T.cons_d2[q] = s.cons_d2[q];
            for (int q = tid; q < s.n_clusters + 1; q += nt) T.cl_coef_start[q] = s.cluster_coef_start[q];
            for (int q = tid; q < s.n_coef; q += nt) T.cl_coef[q] = s.cluster_coef[q];
        }
    }
    CDW_CTA_THREADS(tid, nt) {  // nbeff is complete: pair mixing rule / exceptions / soft-core
        for (int q = tid; q < s.n_pairs; q += nt) {
            EffPair p;
            p.i = s.pair_atoms[2 * q]; p.j = s.pair_atoms[2 * q + 1]; p.flags = 0; p.pad = 0;
            double sig = 0.0, eps = 0.0;
            p.qqC = 0.0;
            p.sc_sig2 = p.sc_eps4 = p.sc_rlj = p.sc_F = p.sc_U = 0.0;
            int nbk = s.pair_nb[q];
            if (nbk == -2) {
                p.qqC = CDW_ONE_4PI_EPS0 * (T.nbeff[3 * p.i] * T.nbeff[3 * p.j]);
                sig = 0.5 * (T.nbeff[3 * p.i + 1] + T.nbeff[3 * p.j + 1]);
                eps = sqrt(T.nbeff[3 * p.i + 2] * T.nbeff[3 * p.j + 2]);
                p.flags |= CDW_PAIR_NB;
            } else if (nbk >= 0) {
                double qq = s.exc_par[3 * nbk], sg = s.exc_par[3 * nbk + 1], ep = s.exc_par[3 * nbk + 2];
                for (int o = s.exc_off_start[nbk]; o < s.exc_off_start[nbk + 1]; ++o) {
                    double l = lam[s.exc_off_slot[o]];
                    qq += l * s.exc_off_par[3 * o]; sg += l * s.exc_off_par[3 * o + 1]; ep += l * s.exc_off_par[3 * o + 2];
                }
                p.qqC = CDW_ONE_4PI_EPS0 * qq; sig = sg; eps = ep;
                p.flags |= CDW_PAIR_NB;
            }
            if (eps != 0.0) p.flags |= CDW_PAIR_LJ;
            p.sig2 = sig * sig; p.eps4 = 4.0 * eps;
            int cls = s.pair_sc_class[q];
            if (cls >= 0)
                lower_softcore(p, cls, s.pair_sc_par + 4 * q, lam[s.slot_core], lam[s.slot_insert], lam[s.slot_delete], lam[s.slot_alpha]);
            T.pairs[q] = p;
        }
    }
}

// ------------------------------------------------------------------------------------------------ energy / forces
CDW_HD d3 ld3(const float* arr, int base) { return mk3((double)arr[base], (double)arr[base + kR], (double)arr[base + 2 * kR]); }
CDW_HD void st3(float* arr, int base, d3 v) { arr[base] = (float)v.x; arr[base + kR] = (float)v.y; arr[base + 2 * kR] = (float)v.z; }
CDW_HD void add3f(float* arr, int base, d3 v) { arr[base] += (float)v.x; arr[base + kR] += (float)v.y; arr[base + 2 * kR] += (float)v.z; }
struct f3 {
    float x, y, z;
};
CDW_HD f3 ldf3(const float* arr, int base) { f3 r; r.x = arr[base]; r.y = arr[base + kR]; r.z = arr[base + 2 * kR]; return r; }
CDW_HD d3 widen(f3 a) { return mk3((double)a.x, (double)a.y, (double)a.z); }
CDW_HD void add3ff(float* arr, int base, f3 v) { arr[base] += v.x; arr[base + kR] += v.y; arr[base + 2 * kR] += v.z; }

// Potential energy (kJ/mol) of the team's replica (same value in every lane of the team).  With F the forces end up
// in force copy 0: each lane accumulates the terms it evaluates into its own copy, then the copies are summed.
// Entered and left with the team synchronised.
//
// `part` selects the terms: PART_ALL, PART_STATIC (the leading lambda-independent terms of every table) or PART_DYNAMIC
// (the rest).  PART_DYNAMIC ADDS to what force copy 0 already holds (the static forces: preloaded from the particle's
// cache row or left there by a PART_STATIC pass), so STATIC followed by DYNAMIC equals ALL up to summation order.
enum { PART_ALL = 0, PART_STATIC = 1, PART_DYNAMIC = 2 };

template <bool F>
CDW_HD double lane_energy(const SysDev& s, const Tables& T, const LaneMem& M, int part = PART_ALL) {
    CDW_UNPACK(M);
    const int A = s.n_atoms;
    const float* __restrict__ x = M.x;
    float* f = M.f + tl * (3 * kR) * A;  // this lane's force copy (NOT __restrict__: lane 0's copy is also read and written through M.f by the reduction below)
    const bool dyn = part == PART_DYNAMIC, sta = part == PART_STATIC;
    const int b0 = dyn ? s.n_bonds_s : 0, b1 = sta ? s.n_bonds_s : s.n_bonds;
    const int a0 = dyn ? s.n_angles_s : 0, a1 = sta ? s.n_angles_s : s.n_angles;
    const int t0 = dyn ? s.n_torsions_s : 0, t1 = sta ? s.n_torsions_s : s.n_torsions;
    const int p0 = dyn ? s.n_pairs_s : 0, p1 = sta ? s.n_pairs_s : s.n_pairs;
    const int e0 = dyn ? s.n_ext_s : 0, e1 = sta ? s.n_ext_s : s.n_ext;
    CDW_TEAM_SYNC(M);
    if (F && !(dyn && tl == 0)) {
        for (int a = 0; a < A; ++a) {
            const int b = CDW_BASE(a);
            f[b] = 0.f; f[b + kR] = 0.f; f[b + 2 * kR] = 0.f;
        }
    }
    double e = 0.0;
    for (int q = b0 + tl; q < b1; q += T_) {
        const EffBond b = T.bonds[q];
        const int bi = CDW_BASE(b.i), bj = CDW_BASE(b.j);
        d3 fi;
        e += bond_term<F>(ld3(x, bi), ld3(x, bj), b.K, b.L, fi);
        if (F) { add3f(f, bi, fi); add3f(f, bj, -1.0 * fi); }
    }
    for (int q = a0 + tl; q < a1; q += T_) {
        const EffAngle a = T.angles[q];
        const int bi = CDW_BASE(a.i), bj = CDW_BASE(a.j), bk = CDW_BASE(a.k);
        d3 fi, fk;
        e += angle_term<F>(ld3(x, bi), ld3(x, bj), ld3(x, bk), a.K, a.T0, fi, fk);
        if (F) { add3f(f, bi, fi); add3f(f, bj, -1.0 * (fi + fk)); add3f(f, bk, fk); }
    }
    for (int q = t0 + tl; q < t1; q += T_) {
        const EffTorsion t = T.torsions[q];
        const int bi = CDW_BASE(t.i), bj = CDW_BASE(t.j), bk = CDW_BASE(t.k), bl = CDW_BASE(t.l);
        d3 fi, fj, fk, fl;
        e += torsion_term<F>(ld3(x, bi), ld3(x, bj), ld3(x, bk), ld3(x, bl), t.kk, t.c0, t.s0, t.n, fi, fj, fk, fl);
        if (F) { add3f(f, bi, fi); add3f(f, bj, fj); add3f(f, bk, fk); add3f(f, bl, fl); }
    }
    {
        // pairs are sorted by (i, j); each lane takes a contiguous block, so runs of pairs share atom i: its position is
        // loaded once and its force is accumulated in registers (fp64) and flushed when the run ends
        const int per = (p1 - p0 + T_ - 1) / T_;
        const int q0 = p0 + tl * per, q1 = q0 + per < p1 ? q0 + per : p1;
        // Geometry and energy in fp64 from the fp32 coordinates; the FORCE vector is formed in fp32 (g rounded once, fp32
        // coordinate differences) because it is accumulated in fp32 anyway: one fp64->fp32 conversion per pair instead of
        // three (conversions run on the quarter-rate XU pipe, the busiest pipe of this kernel).
        int cur = -1, bi = 0;
        f3 xi32 = {0.f, 0.f, 0.f}, fi = {0.f, 0.f, 0.f};
        d3 xi = mk3(0.0, 0.0, 0.0);
        int q = q0;
#define CDW_PAIR_RUN(P)                                                              \
        if ((P).i != cur) {                                                          \
            if (F && cur >= 0) add3ff(f, bi, fi);                                    \
            cur = (P).i; bi = CDW_BASE(cur); xi32 = ldf3(x, bi); xi = widen(xi32);   \
            fi.x = fi.y = fi.z = 0.f;                                                \
        }
#define CDW_PAIR_APPLY(G, XJ32, BJ)                                                  \
        {                                                                            \
            const float g32 = (float)(G);                                            \
            f3 fj;                                                                   \
            fj.x = g32 * (xi32.x - (XJ32).x); fj.y = g32 * (xi32.y - (XJ32).y); fj.z = g32 * (xi32.z - (XJ32).z); \
            fi.x -= fj.x; fi.y -= fj.y; fi.z -= fj.z;                                \
            add3ff(f, (BJ), fj);                                                     \
        }
        for (; q + 1 < q1; q += 2) {  // two pairs per trip: their arithmetic is independent and interleaves
            const EffPair& pa = T.pairs[q];
            const EffPair& pb = T.pairs[q + 1];
            CDW_PAIR_RUN(pa)
            const int bja = CDW_BASE(pa.j), bjb = CDW_BASE(pb.j);
            const f3 xja = ldf3(x, bja), xjb = ldf3(x, bjb);
            if (pb.i == cur) {
                double ga, gb;
                e += pair_term<F>(pa, xi - widen(xja), ga);
                e += pair_term<F>(pb, xi - widen(xjb), gb);
                if (F) { CDW_PAIR_APPLY(ga, xja, bja) CDW_PAIR_APPLY(gb, xjb, bjb) }
            } else {
                double ga, gb;
                e += pair_term<F>(pa, xi - widen(xja), ga);
                if (F) CDW_PAIR_APPLY(ga, xja, bja)
                CDW_PAIR_RUN(pb)
                e += pair_term<F>(pb, xi - widen(xjb), gb);
                if (F) CDW_PAIR_APPLY(gb, xjb, bjb)
            }
        }
        for (; q < q1; ++q) {
            const EffPair& p = T.pairs[q];
            CDW_PAIR_RUN(p)
            const int bj = CDW_BASE(p.j);
            const f3 xj = ldf3(x, bj);
            double g;
            e += pair_term<F>(p, xi - widen(xj), g);
            if (F) CDW_PAIR_APPLY(g, xj, bj)
        }
        if (F && cur >= 0) add3ff(f, bi, fi);
#undef CDW_PAIR_RUN
#undef CDW_PAIR_APPLY
    }
    for (int q = e0 + tl; q < e1; q += T_) {
        const EffExt w = T.exts[q];
        const int bw = CDW_BASE(w.atom);
        d3 fw;
        e += ext_term<F>(ld3(x, bw), w.K, w.x0, w.U0, fw);
        if (F) add3f(f, bw, fw);
    }
    e = team_sum(M, e);
    if (F && T_ > 1) {
        CDW_TEAM_SYNC(M);
        const int copy = (3 * kR) * A;
        for (int a = tl; a < A; a += T_) {
            const int b = CDW_BASE(a);
            float ax = M.f[b], ay = M.f[b + kR], az = M.f[b + 2 * kR];
            for (int c = 1; c < T_; ++c) { ax += M.f[c * copy + b]; ay += M.f[c * copy + b + kR]; az += M.f[c * copy + b + 2 * kR]; }
            M.f[b] = ax; M.f[b + kR] = ay; M.f[b + 2 * kR] = az;
        }
    }
    CDW_TEAM_SYNC(M);
    return e;
}

// The look-ahead kernel evaluates in part A and in part B.  Both call sites are inlined: an out-of-line copy (-DCDW_ENERGY_NOINLINE,
// 1 k fewer SASS instructions) measured 17 % SLOWER on B200 (fluorobenzene, 1e5 replicas: 339 vs 289 us per iteration; the
// 22-atom system, 1e4 replicas: 128 vs 110 us) — the call pins the table pointers and counts in local memory.
#if defined(CDW_ENERGY_NOINLINE)
CDW_NOINLINE_HD inline double lane_energy_forces(const SysDev& s, const Tables& T, const LaneMem& M, int part) { return lane_energy<true>(s, T, M, part); }
#else
CDW_HD double lane_energy_forces(const SysDev& s, const Tables& T, const LaneMem& M, int part) { return lane_energy<true>(s, T, M, part); }
#endif

// ------------------------------------------------------------------------------------------------ constraints
// SHAKE: Gauss-Seidel sweeps per cluster until every distance is inside OpenMM's (1 +- tol)^2 window.  The
// displacement is applied to x and, divided by the drift time, to v ("v += (x - x1)/(dt/2)", integrators.py:310).
// r0 = constraint vectors before the drift (saved by the caller).  Clusters are dealt round-robin to the team lanes.
CDW_HD void lane_shake(const SysDev& s, const Tables& T, const LaneMem& M, double tol, double inv_h) {
    CDW_UNPACK(M);
    const double lower = 1.0 - 2.0 * tol + tol * tol, upper = 1.0 + 2.0 * tol + tol * tol;
    for (int c = tl; c < s.n_clusters; c += T_) {
        const int q0 = T.cl_start[c], q1 = T.cl_start[c + 1];
        for (int it = 0; it < kMaxConstraintIters; ++it) {
            bool updated = false;
            for (int q = q0; q < q1; ++q) {
                const int i = T.cons_atoms[2 * q], j = T.cons_atoms[2 * q + 1];
                const int bi = CDW_BASE(i), bj = CDW_BASE(j);
                const double d2 = T.cons_d2[q];
                const d3 xi = ld3(M.x, bi), xj = ld3(M.x, bj);
                const d3 sv = xi - xj;
                const double r2 = dot(sv, sv);
                if (r2 < lower * d2 || r2 > upper * d2) {
                    const d3 r0 = ld3(M.r0, CDW_BASE(q));
                    const double imi = T.inv_m[i], imj = T.inv_m[j];
                    const double g = (d2 - r2) / (2.0 * (imi + imj) * dot(sv, r0));
                    const double gi = g * imi, gj = g * imj;
                    st3(M.x, bi, xi + gi * r0);
                    st3(M.x, bj, xj - gj * r0);
                    st3(M.v, bi, ld3(M.v, bi) + (gi * inv_h) * r0);
                    st3(M.v, bj, ld3(M.v, bj) - (gj * inv_h) * r0);
                    updated = true;
                }
            }
            if (!updated) break;
        }
    }
    CDW_TEAM_SYNC(M);
}

// Velocity constraints are linear: a cluster of k <= CDW_MAX_DIRECT coupled constraints is projected EXACTLY by one
// k x k symmetric positive-definite solve, A_qr = (s_q . s_r) c_qr, A g = -(s_q . v_rel,q); no sweeps, no tolerance.
template <int K>
CDW_HD void lane_rattle_cluster(const Tables& T, const LaneMem& M, int q0, const double* coef) {
    CDW_UNPACK(M);
    d3 sv[K];
    double b[K], A[K][K];
    int bi[K], bj[K];
    double mi[K], mj[K];
#pragma unroll
    for (int q = 0; q < K; ++q) {
        const int i = T.cons_atoms[2 * (q0 + q)], j = T.cons_atoms[2 * (q0 + q) + 1];
        bi[q] = CDW_BASE(i); bj[q] = CDW_BASE(j);
        mi[q] = T.inv_m[i]; mj[q] = T.inv_m[j];
        sv[q] = ld3(M.x, bi[q]) - ld3(M.x, bj[q]);
        b[q] = -dot(sv[q], ld3(M.v, bi[q]) - ld3(M.v, bj[q]));
    }
#pragma unroll
    for (int q = 0; q < K; ++q)
#pragma unroll
        for (int r = 0; r < K; ++r) A[q][r] = dot(sv[q], sv[r]) * coef[q * K + r];
#pragma unroll
    for (int p = 0; p < K; ++p) {
        const double ip = 1.0 / A[p][p];
#pragma unroll
        for (int r = p + 1; r < K; ++r) {
            const double fct = A[r][p] * ip;
#pragma unroll
            for (int c = p + 1; c < K; ++c) A[r][c] -= fct * A[p][c];
            b[r] -= fct * b[p];
        }
        A[p][p] = ip;
    }
#pragma unroll
    for (int p = K - 1; p >= 0; --p) {
        double t = b[p];
#pragma unroll
        for (int c = p + 1; c < K; ++c) t -= A[p][c] * b[c];
        b[p] = t * A[p][p];
    }
#pragma unroll
    for (int q = 0; q < K; ++q) {  // sequential: constraints of a cluster share atoms
        st3(M.v, bi[q], ld3(M.v, bi[q]) + (b[q] * mi[q]) * sv[q]);
        st3(M.v, bj[q], ld3(M.v, bj[q]) - (b[q] * mj[q]) * sv[q]);
    }
}

CDW_HD void lane_rattle(const SysDev& s, const Tables& T, const LaneMem& M, double tol) {
    CDW_UNPACK(M);
    for (int c = tl; c < s.n_clusters; c += T_) {
        const int q0 = T.cl_start[c], k = T.cl_start[c + 1] - q0;
        const double* coef = T.cl_coef + T.cl_coef_start[c];
        if (k == 1) lane_rattle_cluster<1>(T, M, q0, coef);
        else if (k == 2) lane_rattle_cluster<2>(T, M, q0, coef);
        else if (k == 3) lane_rattle_cluster<3>(T, M, q0, coef);
        else {
            // large clusters: Gauss-Seidel sweeps to |s . v_rel| <= tol d^2
            for (int it = 0; it < kMaxConstraintIters; ++it) {
                bool updated = false;
                for (int q = q0; q < q0 + k; ++q) {
                    const int i = T.cons_atoms[2 * q], j = T.cons_atoms[2 * q + 1];
                    const int bi = CDW_BASE(i), bj = CDW_BASE(j);
                    const double d2 = T.cons_d2[q];
                    const d3 sv = ld3(M.x, bi) - ld3(M.x, bj);
                    const d3 vi = ld3(M.v, bi), vj = ld3(M.v, bj);
                    const double sd = dot(sv, vi - vj);
                    if (fabs(sd) > tol * d2) {
                        const double imi = T.inv_m[i], imj = T.inv_m[j];
                        const double g = -sd / ((imi + imj) * dot(sv, sv));
                        st3(M.v, bi, vi + (g * imi) * sv);
                        st3(M.v, bj, vj - (g * imj) * sv);
                        updated = true;
                    }
                }
                if (!updated) break;
            }
        }
    }
    CDW_TEAM_SYNC(M);
}

CDW_HD double lane_kinetic(const SysDev& s, const Tables& T, const LaneMem& M) {
    CDW_UNPACK(M);
    double k = 0.0;
    for (int a = tl; a < s.n_atoms; a += T_) {
        const d3 v = ld3(M.v, CDW_BASE(a));
        k += dot(v, v) / T.inv_m[a];
    }
    return 0.5 * team_sum(M, k);
}

// ------------------------------------------------------------------------------------------------ kernel arguments
struct PropArgs {
    const float4* pos_in; const float4* vel_in; float4* pos_out; float4* vel_out;
    long long n;
    const int* anc; const double* aux_in; const double* cum_in; const int* label_in;
    const double* lambda;
    const unsigned char* tables;  // prebuilt image of the CTA-shared table region for this lambda (or NULL: build in the prologue)
    double dt, gamma, kT, tol;
    int n_steps; unsigned flags;
    unsigned long long seed, step0; long long gid0;
    double* inc_out; double* w_out; double* aux_out; int* label_out;
    double* u_first_out; double* u_last_out; double* shadow_out; double* proposal_out;
    int* nan_flags; unsigned long long* wmin_key;
    double* energy_out; float4* force_out;  // cdw_energy_forces
    double beta;                             // cdw_reduced_potential
    // sharded runs: ancestors are GLOBAL ids; rank g owns [g * n_per_rank, (g + 1) * n_per_rank) and its rows are
    // reachable through peer-mapped pointers (CUDA IPC over NVLink).  NULL peer_pos = single-GPU addressing.
    const float4* const* peer_pos; const float4* const* peer_vel; const double* const* peer_aux; const int* const* peer_label;
    long long n_per_rank;
    // static-term cache (cdw_static_cache): forces / energy of the lambda-independent terms at the INPUT rows (may be NULL:
    // then they are evaluated) and at the OUTPUT rows.  Indexed like pos_in / pos_out (input through the ancestor).
    const float4* fs_in; const double* us_in; float4* fs_out; double* us_out;
    const float4* const* peer_fs; const double* const* peer_us;
    // look-ahead iteration (cdw_smc_lookahead, smc_lookahead_kernel): fs_in / peer_fs rows hold the TOTAL forces F(x_{t-1}; lambda_t) of the
    // input rows, fs_out receives F(x_t; lambda_{t+1}); lambda_next / tables_next describe lambda_{t+1} (NULL on the last iteration) and
    // inc_next_out[p] = u_{t+1}(x_t) - u_t(x_t), the incremental work of the NEXT iteration (distribution_factories.py:124-129)
    const double* lambda_next; const unsigned char* tables_next; double* inc_next_out;
    int have_next;  // part B runs (lambda_next may still be NULL for a system without global parameters)
    // programmable splitting (cdw_langevin_params.splitting != 0, lane_step_program): 4 bits per sub-step, CDW_STEP_*; 0 ends the program
    unsigned long long program; int* trial_counts;  // [P][2] (trials, accepted) of the Metropolised block, may be NULL
    // quadratic twist of the Euler-Maruyama proposal (cdw_langevin_params.twist; troubleshooting/csmc.py:460-623): a[3A] b[3A] c d, or NULL
    const double* twist;
    // dense, position-independent twist (cdw_dense_twist, include/cdw.h): Theta = Sigma~ S^-1, the lower Cholesky factor of Sigma~, A, b,
    // beta = Sigma~ b (all [D][D] / [D], D = 3 A, row-major), logk0 = 0.5 log det Theta + 0.5 b^T Sigma~ b - c - d, cd = c + d, and a
    // scratch area of 2 D doubles per particle of the call
    const double* tw_theta; const double* tw_chol; const double* tw_A; const double* tw_b; const double* tw_beta;
    double tw_logk0, tw_cd; double* tw_scratch;
    int diag_local_rank;  // timing diagnostic (CDW_DIAG_LOCAL_ROWS, results are wrong): rank + 1 = read every ancestor's row from this rank
};

// where row `src` (a global particle id when the run is sharded) lives
CDW_HD const float4* row_ptr(const PropArgs& a, const float4* local, const float4* const* peers, long long src, int A) {
    if (!a.peer_pos) return local + src * A;
    const long long r = src / a.n_per_rank;
    return peers[a.diag_local_rank ? a.diag_local_rank - 1 : r] + (src - r * a.n_per_rank) * A;
}

// (x, y, z) of one float4 row element -> three shared-memory columns
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ void cdw_copy3(float* arr, int base, const float* g) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(arr + base);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(g) : "memory");
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d + 4u * kR), "l"(g + 1) : "memory");
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d + 8u * kR), "l"(g + 2) : "memory");
}
__device__ __forceinline__ void cdw_copy_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }
#else
inline void cdw_copy3(float* arr, int base, const float* g) { arr[base] = g[0]; arr[base + kR] = g[1]; arr[base + 2 * kR] = g[2]; }
inline void cdw_copy_wait() {}
#endif

// CTA-cooperative, coalesced load of `count` rows starting at particle p0 into the [element][replica] layout.
// Row r comes from particle anc[p0 + r] when anc != NULL: this load IS the resampling gather (resamplers.py:123-129).
CDW_HD void cta_load_rows(const SysDev& s, const PropArgs& a, unsigned char* smem, const Layout& L, int sw, long long p0, int count, bool with_v) {
    const int A = s.n_atoms;
    float* X = (float*)(smem + L.x);
    float* V = (float*)(smem + L.v);
    long long* SRC = (long long*)(smem + L.src);
    CDW_CTA_THREADS(tid, nt) {
        if (with_v) {
            for (int r = tid; r < count; r += nt) SRC[r] = a.anc ? (long long)a.anc[p0 + r] : p0 + r;
        }
    }
    // The rows are gathered (latency, not bandwidth, is the cost): on the device every component goes global -> shared
    // with a 4-byte cp.async (LDGSTS), no register staging, so ALL of a thread's loads are in flight at once (measured
    // alternatives, profiles/r2_notes.md: a register-staged 16-byte load for rows that come over NVLink, +20-30 % per anneal on 2
    // GPUs; lanes copying consecutive words of a row so that every sector is requested once, +4 % on one GPU, +9 % on two); one
    // wait_all before the CTA barrier.
    CDW_CTA_THREADS(tid, nt) {
        float* Fc = (float*)(smem + L.f);
        const bool with_f = with_v && a.fs_in != nullptr;
        for (int idx = tid; idx < count * A; idx += nt) {
            const int r = idx / A, at = idx - r * A;
            const long long src = with_v ? SRC[r] : p0 + r;
            const int b = CDW_BASE_R(at, r);
            cdw_copy3(X, b, (const float*)(row_ptr(a, a.pos_in, a.peer_pos, src, A) + at));  // NVLink load when the ancestor lives on a peer
            if (with_v) {
                if (a.vel_in) cdw_copy3(V, b, (const float*)(row_ptr(a, a.vel_in, a.peer_vel, src, A) + at));
                else { V[b] = 0.f; V[b + kR] = 0.f; V[b + 2 * kR] = 0.f; }
            }
            if (with_f) cdw_copy3(Fc, b, (const float*)(row_ptr(a, a.fs_in, a.peer_fs, src, A) + at));  // cached forces of the ancestor's coordinates -> copy 0
        }
        cdw_copy_wait();
    }
}

// CTA-cooperative, coalesced store; a replica whose move produced a non-finite energy keeps the caller's state
// (propagators.py:166-186).
CDW_HD void cta_store_rows(const SysDev& s, const PropArgs& a, unsigned char* smem, const Layout& L, int sw, long long p0, int count) {
    const int A = s.n_atoms;
    const float* X = (const float*)(smem + L.x);
    const float* V = (const float*)(smem + L.v);
    const float* Fv = (const float*)(smem + L.f);
    const long long* SRC = (const long long*)(smem + L.src);
    const int* BAD = (const int*)(smem + L.bad);
    CDW_CTA_THREADS(tid, nt) {
        for (int idx = tid; idx < count * A; idx += nt) {
            const int r = idx / A, at = idx - r * A;
            const int b = CDW_BASE_R(at, r);
            if (a.force_out) a.force_out[(p0 + r) * A + at] = make_float4(Fv[b], Fv[b + kR], Fv[b + 2 * kR], 0.f);
            if (a.pos_out) {
                float4 q, w;
                if (!BAD[r]) {
                    q = make_float4(X[b], X[b + kR], X[b + 2 * kR], 0.f);
                    w = make_float4(V[b], V[b + kR], V[b + 2 * kR], 0.f);
                } else {
                    q = row_ptr(a, a.pos_in, a.peer_pos, SRC[r], A)[at];
                    w = a.vel_in ? row_ptr(a, a.vel_in, a.peer_vel, SRC[r], A)[at] : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                a.pos_out[(p0 + r) * A + at] = q;
                if (a.vel_out) a.vel_out[(p0 + r) * A + at] = w;
                if (a.fs_out && a.fs_in && BAD[r]) a.fs_out[(p0 + r) * A + at] = row_ptr(a, a.fs_in, a.peer_fs, SRC[r], A)[at];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ integrator sub-steps
// V : v += h f / m                                                  (integrators.py:317-336)
CDW_HD void lane_kick(const SysDev& s, const Tables& T, const LaneMem& M, double h) {
    CDW_UNPACK(M);
    for (int at = tl; at < s.n_atoms; at += T_) {
        const int b = CDW_BASE(at);
        st3(M.v, b, ld3(M.v, b) + (h * T.inv_m[at]) * ld3(M.f, b));
    }
}
// R : x += h v ; SHAKE ; v += (x - x1) / h                          (integrators.py:297-315)
CDW_HD void lane_drift(const SysDev& s, const Tables& T, const LaneMem& M, double h, double tol) {
    CDW_UNPACK(M);
    for (int q = tl; q < s.n_cons; q += T_) {  // constraint vectors before the drift (fp32 differences)
        const int bi = CDW_BASE(T.cons_atoms[2 * q]), bj = CDW_BASE(T.cons_atoms[2 * q + 1]), bq = CDW_BASE(q);
        M.r0[bq] = M.x[bi] - M.x[bj]; M.r0[bq + kR] = M.x[bi + kR] - M.x[bj + kR]; M.r0[bq + 2 * kR] = M.x[bi + 2 * kR] - M.x[bj + 2 * kR];
    }
    CDW_TEAM_SYNC(M);
    for (int at = tl; at < s.n_atoms; at += T_) {
        const int b = CDW_BASE(at);
        st3(M.x, b, ld3(M.x, b) + h * ld3(M.v, b));
    }
    CDW_TEAM_SYNC(M);
    if (s.n_cons) lane_shake(s, T, M, tol, 1.0 / h);
}
// O : v = cv v + cn sqrt(kT/m) xi, xi = Philox(seed; gid, step, atom, kind)   (integrators.py:338-350; cv = 0, cn = 1: Maxwell-Boltzmann)
CDW_HD void lane_noise(const SysDev& s, const Tables& T, const LaneMem& M, double cv, double cn, unsigned long long seed, unsigned long long gid,
                       uint32_t step, uint32_t kind) {
    CDW_UNPACK(M);
    for (int at = tl; at < s.n_atoms; at += T_) {
        double z0, z1, z2;
        normals3(seed, gid, step, (uint32_t)at, kind, z0, z1, z2);
        const int b = CDW_BASE(at);
        const double sg = cn * T.sig_v[at];
        st3(M.v, b, cv * ld3(M.v, b) + sg * mk3(z0, z1, z2));
    }
}

// ---- dense twist of the Euler-Maruyama move (troubleshooting/csmc.py:460-623 with a full matrix A_t; the uncontrolled kernel is N(f, S),
// S = diag(s), s = 2 tau per coordinate, f = x + tau beta F).  With Sigma~ = (S^-1 + 2 A)^-1 and Theta = Sigma~ S^-1 (Theta_t, :460-478):
//     x' ~ N(Theta f - Sigma~ b, Sigma~)                                          twisted_forward_proposal (:503-528)
//     log K_t(psi_t)(x) = 0.5 log det Theta - (Theta f).(A f) - b.(Theta f) + 0.5 b^T Sigma~ b - c - d
// which is twisted_forward_log_normalizer (:531-561) with the two O(|f|^2 / s) quadratic forms, whose difference it is, expanded
// analytically (Sigma~ = S - 2 Sigma~ A S): nothing cancels in floating point.  The team splits the rows of the three matrix-vector products;
// vectors travel through the per-particle scratch (2 D doubles).  Returns this lane's share of the x-dependent part of log K; x_old is left in
// the v column, the unrounded-then-fp32 x' in the x column (position constraints are the caller's).
CDW_HD double lane_twisted_move_dense(const SysDev& s, const Tables& T, const LaneMem& M, const PropArgs& a, long long p, unsigned long long gid,
                                      uint32_t step, double inv_kT) {
    CDW_UNPACK(M);
    const int A = s.n_atoms, D = 3 * A;
    double* S0 = a.tw_scratch + (size_t)p * 2 * D;   // f, later xi
    double* S1 = S0 + D;                              // the proposal mean
    for (int at = tl; at < A; at += T_) {
        const int b = CDW_BASE(at);
        const double tau = 0.5 * a.kT * a.dt * a.dt * T.inv_m[at];
        const d3 x = ld3(M.x, b);
        st3(M.v, b, x);
        const d3 f = x + (tau * inv_kT) * ld3(M.f, b);
        S0[3 * at] = f.x; S0[3 * at + 1] = f.y; S0[3 * at + 2] = f.z;
    }
    CDW_TEAM_SYNC(M);
    double lk = 0.0;
    for (int at = tl; at < A; at += T_) {   // the three rows of an atom share every load of f: six independent accumulation chains
        const int i = 3 * at;
        const double* th = a.tw_theta + (size_t)i * D;
        const double* am = a.tw_A + (size_t)i * D;
        double u0 = 0.0, u1 = 0.0, u2 = 0.0, w0 = 0.0, w1 = 0.0, w2 = 0.0;
        for (int j = 0; j < D; ++j) {
            const double fj = S0[j];
            u0 = fma(th[j], fj, u0); u1 = fma(th[D + j], fj, u1); u2 = fma(th[2 * D + j], fj, u2);
            w0 = fma(am[j], fj, w0); w1 = fma(am[D + j], fj, w1); w2 = fma(am[2 * D + j], fj, w2);
        }
        lk -= u0 * w0 + a.tw_b[i] * u0;
        lk -= u1 * w1 + a.tw_b[i + 1] * u1;
        lk -= u2 * w2 + a.tw_b[i + 2] * u2;
        S1[i] = u0 - a.tw_beta[i]; S1[i + 1] = u1 - a.tw_beta[i + 1]; S1[i + 2] = u2 - a.tw_beta[i + 2];
    }
    CDW_TEAM_SYNC(M);   // every lane is done with f
    for (int at = tl; at < A; at += T_) {
        double z0, z1, z2;
        normals3(a.seed, gid, step, (uint32_t)at, CDW_KIND_ULA, z0, z1, z2);
        S0[3 * at] = z0; S0[3 * at + 1] = z1; S0[3 * at + 2] = z2;
    }
    CDW_TEAM_SYNC(M);
    for (int at = tl; at < A; at += T_) {
        const int i = 3 * at;
        const double* L = a.tw_chol + (size_t)i * D;
        double x0 = S1[i], x1 = S1[i + 1], x2 = S1[i + 2];
        for (int j = 0; j <= i; ++j) {   // (lower triangle: rows i + 1, i + 2 have one / two more entries)
            const double zj = S0[j];
            x0 = fma(L[j], zj, x0); x1 = fma(L[D + j], zj, x1); x2 = fma(L[2 * D + j], zj, x2);
        }
        x1 = fma(L[D + i + 1], S0[i + 1], x1);
        x2 = fma(L[2 * D + i + 1], S0[i + 1], x2);
        x2 = fma(L[2 * D + i + 2], S0[i + 2], x2);
        st3(M.x, CDW_BASE(at), mk3(x0, x1, x2));
    }
    return lk;
}

// phi_t(x') - c - d = x'^T A x' + b^T x' of the dense twist: this lane's rows (twisted_t_log_weight, csmc.py:593-623)
CDW_HD double lane_twist_phi_dense(const SysDev& s, const LaneMem& M, const PropArgs& a) {
    CDW_UNPACK(M);
    const int A = s.n_atoms, D = 3 * A;
    double phi = 0.0;
    for (int at = tl; at < A; at += T_) {
        const d3 xi = ld3(M.x, CDW_BASE(at));
        const int i = 3 * at;
        const double* am = a.tw_A + (size_t)i * D;
        double w0 = 0.0, w1 = 0.0, w2 = 0.0;
        for (int aj = 0; aj < A; ++aj) {
            const d3 xj = ld3(M.x, CDW_BASE(aj));
            const int j = 3 * aj;
            w0 = fma(am[j], xj.x, w0); w0 = fma(am[j + 1], xj.y, w0); w0 = fma(am[j + 2], xj.z, w0);
            w1 = fma(am[D + j], xj.x, w1); w1 = fma(am[D + j + 1], xj.y, w1); w1 = fma(am[D + j + 2], xj.z, w1);
            w2 = fma(am[2 * D + j], xj.x, w2); w2 = fma(am[2 * D + j + 1], xj.y, w2); w2 = fma(am[2 * D + j + 2], xj.z, w2);
        }
        phi += xi.x * (w0 + a.tw_b[i]);
        phi += xi.y * (w1 + a.tw_b[i + 1]);
        phi += xi.z * (w2 + a.tw_b[i + 2]);
    }
    return phi;
}

// K4 body (+ first half of K2) for the replica already staged in this team's column.
// Returns the order-preserving key of W[p] (for the running min).
//
// The integrator is a tiny op-code loop  E (V R O R E V)^n  with ONE call site per routine (force evaluation, SHAKE,
// velocity projection, kinetic energy): the fully inlined body stays inside the instruction cache.
//   E : energy + forces at the current x.   First E: E1, F1 at (x_{t-1}, lambda_t) -> the incremental work
//       (distribution_factories.py:120-129); later E's close a step.
//   V : v += (dt/2) f / m                                            (integrators.py:317-336)
//   R : x += (dt/2) v ; SHAKE ; v += (x - x1)/(dt/2)                 (integrators.py:297-315)
//   O : v = a v + b sqrt(kT/m) xi                                    (integrators.py:338-350)
//   every V, R, O is followed by the velocity-constraint projection, as every step of the reference program is.
CDW_HD unsigned long long lane_step(const SysDev& s, const Tables& T, const LaneMem& M, const PropArgs& a, long long p) {
    CDW_UNPACK(M);
    const int A = s.n_atoms;
    const long long src = M.src[rid];
    const unsigned long long gid = (unsigned long long)(a.gid0 + p);
    const double h = 0.5 * a.dt;
    const double ca = exp(-a.gamma * a.dt), cb = sqrt(1.0 - exp(-2.0 * a.gamma * a.dt));
    const bool track = (a.flags & CDW_FLAG_TRACK_WORKS) != 0;
    const double inv_kT = 1.0 / a.kT;
    enum { OP_E = 0, OP_V = 1, OP_R = 2, OP_O = 3, OP_MB = 4, OP_U = 5, OP_B = 6 };
    const bool ula = (a.flags & CDW_FLAG_ULA) != 0;
    // Static-term cache: with `split` every evaluation is two passes of the ONE lane_energy call site, static terms then
    // dynamic terms (ops E E), and the opening evaluation skips the static pass when the particle's cache row is given.
    const bool split = a.fs_out != nullptr && !track;
    const bool cached_open = split && a.fs_in != nullptr && a.n_steps > 0;
    const int n_open = split && !cached_open ? 2 : 1;                 // [ES] ED | E
    const int n_body = (ula ? 3 : 6) + (split ? 1 : 0);               // V R O R [ES] E(D) V | U [ES] E(D) B
    const int e_slot = ula ? 1 : 4;                                   // slot of the (first) evaluation inside a step
    double U = 0.0, Us = 0.0, u_first = 0.0, shadow = 0.0, proposal = 0.0, pending = 0.0, twist_lk = 0.0;
    double aux = 0.0, inc = 0.0, w = 0.0;
    int label = (int)src;
    if (a.peer_pos) {
        const long long r = src / a.n_per_rank, l = src - r * a.n_per_rank;
        if (a.peer_aux) aux = a.peer_aux[r][l];
        if (a.peer_label) label = a.peer_label[r][l];
    } else {
        if (a.aux_in) aux = a.aux_in[src];
        if (a.label_in) label = a.label_in[src];
    }
    if (cached_open) Us = a.peer_pos ? a.peer_us[src / a.n_per_rank][src % a.n_per_rank] : a.us_in[src];
    const int n_ops = n_open + n_body * a.n_steps;
    // op -1 (only with RANDOMIZE): Context.setVelocitiesToTemperature (propagators.py:137-138)
    for (int op = (!ula && (a.flags & CDW_FLAG_RANDOMIZE_VELOCITIES)) ? -1 : 0; op < n_ops; ++op) {
        int kind, slot = 0, st = 0, part = PART_ALL;
        if (op < 0) kind = OP_MB;
        else if (op < n_open) {
            kind = OP_E;
            part = !split ? PART_ALL : (cached_open || op == 1 ? PART_DYNAMIC : PART_STATIC);
        } else {
            slot = (op - n_open) % n_body; st = (op - n_open) / n_body;
            if (split && slot > e_slot) slot -= 1, part = PART_DYNAMIC;   // the second pass of the evaluation
            else if (split && slot == e_slot) part = PART_STATIC;
            if (ula) kind = slot == 0 ? OP_U : (slot == 1 ? OP_E : OP_B);                                               // U E B
            else kind = slot == 0 || slot == 5 ? OP_V : (slot == 1 || slot == 3 ? OP_R : (slot == 2 ? OP_O : OP_E));   // V R O R E V
            if (kind != OP_E) part = PART_ALL;
        }
        if (kind == OP_E) {
            const double e = lane_energy<true>(s, T, M, part);
            if (part == PART_STATIC) {
                Us = e;
                if (st == a.n_steps - 1 || a.n_steps == 0) {  // static forces (force copy 0 right now) at the output coordinates -> cache row
                    for (int at = tl; at < A; at += T_) {
                        const int b = CDW_BASE(at);
                        a.fs_out[p * A + at] = make_float4(M.f[b], M.f[b + kR], M.f[b + 2 * kR], 0.f);
                    }
                }
                continue;
            }
            U = part == PART_DYNAMIC ? Us + e : e;
            if (op < n_open) {
                u_first = U;
                if (!ula) {
                    // MCMC move: work_inc = u_t(x_{t-1}) - u_{t-1}(x_{t-1}) (distribution_factories.py:98-131)
                    inc = u_first * inv_kT + aux;
                    w = (a.cum_in ? a.cum_in[p] : 0.0) + inc;
                    if (tl == 0) {
                        if (a.inc_out) a.inc_out[p] = inc;
                        if (a.w_out) a.w_out[p] = w;
                    }
                }
            } else if (track) shadow += pending + U;  // closes the second R of the step: (KE + U_new) - (KE_old + U_mid)
            continue;
        }
        if (kind == OP_U || kind == OP_B) {
            // Euler-Maruyama proposal (integrators.py:676-680, 741-745, 829-888); the v column holds x_old.
            //   U : x' = x + tau beta F(x) + sqrt(2 tau) xi, tau_a = kT dt^2 / (2 m_a); constrain positions;
            //       pending = log K(x'|x)  = -sum |x' - x - tau beta F(x)|^2 / (4 tau)
            //   B : proposal += pending - log L(x|x'),  log L = -sum |x - x' - tau beta F(x')|^2 / (4 tau)
            const bool fwd = kind == OP_U;
            if (fwd) {
                for (int q = tl; q < s.n_cons; q += T_) {
                    const int bi = CDW_BASE(T.cons_atoms[2 * q]), bj = CDW_BASE(T.cons_atoms[2 * q + 1]), bq = CDW_BASE(q);
                    M.r0[bq] = M.x[bi] - M.x[bj]; M.r0[bq + kR] = M.x[bi + kR] - M.x[bj + kR]; M.r0[bq + 2 * kR] = M.x[bi + 2 * kR] - M.x[bj + 2 * kR];
                }
                CDW_TEAM_SYNC(M);  // every constraint vector is taken before any lane moves an atom
                double lk = 0.0;
                if (a.tw_theta) lk = lane_twisted_move_dense(s, T, M, a, p, gid, (uint32_t)(a.step0 + st), inv_kT);
                else for (int at = tl; at < A; at += T_) {
                    const int b = CDW_BASE(at);
                    const double tau = 0.5 * a.kT * a.dt * a.dt * T.inv_m[at];
                    double z0, z1, z2;
                    normals3(a.seed, gid, (uint32_t)(a.step0 + st), (uint32_t)at, CDW_KIND_ULA, z0, z1, z2);
                    const d3 x = ld3(M.x, b);
                    st3(M.v, b, x);
                    if (!a.twist) {
                        st3(M.x, b, x + ((tau * inv_kT) * ld3(M.f, b) + sqrt(2.0 * tau) * mk3(z0, z1, z2)));
                    } else {
                        // twisted move (csmc.py:460-528 with the diagonal A_t, b_t of this iteration; the uncontrolled kernel is N(f, s I),
                        // f = x + tau beta F (f_t, :480-500), s = 2 tau):  Theta = 1 / (1 + 2 s a)  (Theta_t, :460-478),
                        // x' ~ N(Theta (f - s b), s Theta)  (twisted_forward_proposal, :503-528), and its log normaliser
                        // log K_t(psi_t)(x) = sum 0.5 log Theta + (f - s b)^2 Theta / (2 s) - f^2 / (2 s)  - c - d  (:531-561)
                        const d3 f = x + (tau * inv_kT) * ld3(M.f, b);
                        const double sc = 2.0 * tau;
                        const double fk[3] = {f.x, f.y, f.z}, zk[3] = {z0, z1, z2};
                        double xn[3];
                        for (int k = 0; k < 3; ++k) {
                            const double th = 1.0 / (1.0 + 2.0 * sc * a.twist[3 * at + k]);
                            const double g = fk[k] - sc * a.twist[3 * A + 3 * at + k];
                            xn[k] = th * g + sqrt(sc * th) * zk[k];
                            const double quad = (a.flags & CDW_FLAG_TWIST_REFERENCE_NORMALIZER) ? g * g / th : g * g * th;  // (:556 passes Theta^-1)
                            lk += 0.5 * log(th) + (quad - fk[k] * fk[k]) / (2.0 * sc);
                        }
                        st3(M.x, b, mk3(xn[0], xn[1], xn[2]));
                    }
                }
                if (a.twist) twist_lk = team_sum(M, lk) - a.twist[6 * A] - a.twist[6 * A + 1];
                if (a.tw_theta) twist_lk = team_sum(M, lk) + a.tw_logk0;
                CDW_TEAM_SYNC(M);
                if (s.n_cons) lane_shake(s, T, M, a.tol, 0.0);
            }
            double acc = 0.0;
            for (int at = tl; at < A; at += T_) {
                const int b = CDW_BASE(at);
                const double tau = 0.5 * a.kT * a.dt * a.dt * T.inv_m[at];
                const d3 xn = ld3(M.x, b), xo = ld3(M.v, b);
                const d3 d = (fwd ? xn - xo : xo - xn) - (tau * inv_kT) * ld3(M.f, b);
                acc += dot(d, d) / (4.0 * tau);
            }
            acc = team_sum(M, acc);
            if (fwd) pending = -acc;
            else proposal += pending + acc;
            if (!fwd && a.twist) {
                // twisted incremental weight (twisted_t_log_weight, csmc.py:593-623): log w = log w_uncontrolled + log K_t(psi_t)(x) + phi_t(x'),
                // phi_t(x') = x'^T A x' + b^T x' + c + d; the incremental WORK is -log w
                double phi = 0.0;
                for (int at = tl; at < A; at += T_) {
                    const d3 xn = ld3(M.x, CDW_BASE(at));
                    const double* ta = a.twist + 3 * at;
                    const double* tb = a.twist + 3 * A + 3 * at;
                    phi += (ta[0] * xn.x + tb[0]) * xn.x + (ta[1] * xn.y + tb[1]) * xn.y + (ta[2] * xn.z + tb[2]) * xn.z;
                }
                phi = team_sum(M, phi) + a.twist[6 * A] + a.twist[6 * A + 1];
                proposal -= twist_lk + phi;
            }
            if (!fwd && a.tw_theta) proposal -= twist_lk + (team_sum(M, lane_twist_phi_dense(s, M, a)) + a.tw_cd);
            continue;
        }
        const double ke_before = track ? lane_kinetic(s, T, M) : 0.0;
        if (kind == OP_V) lane_kick(s, T, M, h);
        else if (kind == OP_R) lane_drift(s, T, M, h, a.tol);
        else {  // OP_O / OP_MB: fresh Gaussian noise, counter = (seed; gid, step, atom, kind)
            const bool mb = kind == OP_MB;
            lane_noise(s, T, M, mb ? 0.0 : ca, mb ? 1.0 : cb, a.seed, gid, (uint32_t)(a.step0 + st), mb ? CDW_KIND_MAXWELL_BOLTZMANN : CDW_KIND_O_STEP);
        }
        CDW_TEAM_SYNC(M);
        if (s.n_cons) lane_rattle(s, T, M, a.tol);
        if (track && kind != OP_MB) {
            const double ke_after = lane_kinetic(s, T, M);
            if (kind == OP_V) shadow += ke_after - ke_before;                    // integrators.py:336
            else if (kind == OP_O) proposal -= ke_after - ke_before;             // integrators.py:350
            else if (slot == 1) {                                                // first R: needs the energy at the new x (integrators.py:313-315)
                const double U_mid = lane_energy<true>(s, T, M);
                shadow += (ke_after + U_mid) - (ke_before + U);
                U = U_mid;
            } else pending = ke_after - (ke_before + U);                         // second R: U here is U_mid; closed by the next E
        }
    }
    const double u_last = U;
    const double fin = ula ? u_last + proposal : u_last;
    const bool bad = a.n_steps > 0 && !(fin - fin == 0.0);  // NaN or inf
    if (ula) {
        // Euler-Maruyama proposal: u_t(x_t) - u_{t-1}(x_{t-1}) + proposal work (distribution_factories.py:98-131); a reverted move is a
        // rejected one (zero proposal work, state kept)
        inc = (!bad ? u_last * inv_kT + proposal : u_first * inv_kT) + aux;
        w = (a.cum_in ? a.cum_in[p] : 0.0) + inc;
    }
    if (tl == 0) {
        M.bad[rid] = bad ? 1 : 0;
        if (ula && a.inc_out) a.inc_out[p] = inc;
        if (ula && a.w_out) a.w_out[p] = w;
        if (a.aux_out) a.aux_out[p] = -(bad ? u_first : u_last) * inv_kT;
        if (a.label_out) a.label_out[p] = label;
        if (a.u_first_out) a.u_first_out[p] = u_first * inv_kT;
        if (a.u_last_out) a.u_last_out[p] = u_last * inv_kT;
        if (a.shadow_out) a.shadow_out[p] = shadow * inv_kT;
        if (a.proposal_out) a.proposal_out[p] = ula ? proposal : proposal * inv_kT;
        if (a.nan_flags) a.nan_flags[p] = bad ? 1 : 0;
        if (a.energy_out) a.energy_out[p] = u_first;
        if (split && a.us_out) {  // a reverted move keeps the ancestor's cache (its static-force row is restored by cta_store_rows)
            double us_keep = Us;
            if (bad && cached_open) us_keep = a.peer_pos ? a.peer_us[src / a.n_per_rank][src % a.n_per_rank] : a.us_in[src];
            a.us_out[p] = us_keep;
        }
    }
    return ordered_key(w);
}

// ------------------------------------------------------------------------------------------------ programmable splitting
// OMMLI builds its integrator from a splitting string (openmm/integrators.py:207-217, 363-425): any sequence of R, V, O and at most
// one Metropolised block "{ ... }" that must hold every R and V (openmm/integrators.py:241-276, 427-444).  This body interprets such
// a program (CDW_STEP_* codes, 4 bits each) with the reference's bookkeeping:
//   V, R, O   as in lane_step; forces / energy are re-evaluated lazily, when a V or a work increment needs them at moved positions
//             (OpenMM's `f` and `energy`); shadow_work += dKE (V), d(KE + PE) (R); proposal_work -= dKE (O)
//   {         x_old = x, v_old = v                                                                   (:427-430)
//   }         accept iff exp(-shadow_work / kT) >= uniform; else x = x_old, v = -v_old; shadow_work = 0   (:432-444)
// "V R O R V" gives the same trajectory as lane_step (one extra evaluation per step for the R bookkeeping, as with TRACK_WORKS).
#define CDW_KIND_METROPOLIS 3u   // (CDW_STEP_*: include/cdw.h)

CDW_HD unsigned long long lane_step_program(const SysDev& s, const Tables& T, const LaneMem& M, const PropArgs& a, long long p) {
    CDW_UNPACK(M);
    const int A = s.n_atoms;
    const long long src = M.src[rid];
    const unsigned long long gid = (unsigned long long)(a.gid0 + p);
    const double h_r = a.dt, inv_kT = 1.0 / a.kT;
    int n_v = 0, n_r = 0, n_o = 0;
    for (unsigned long long w = a.program; w & 15ull; w >>= 4) {
        const unsigned tok = (unsigned)(w & 15ull);
        n_v += tok == CDW_STEP_V; n_r += tok == CDW_STEP_R; n_o += tok == CDW_STEP_O;
    }
    // every kind of sub-step shares the time step equally (openmm/integrators.py:300, 320, 158-159: dt / n_R, dt / n_V, dt / n_O)
    const double hv = n_v ? a.dt / n_v : 0.0, hr = n_r ? h_r / n_r : 0.0, ho = n_o ? a.dt / n_o : 0.0;
    const double ca = exp(-a.gamma * ho), cb = sqrt(1.0 - exp(-2.0 * a.gamma * ho));
    double aux = 0.0;
    int label = (int)src;
    if (a.aux_in) aux = a.aux_in[src];
    if (a.label_in) label = a.label_in[src];
    // opening evaluation: the incremental work of an MCMC move (distribution_factories.py:98-131)
    double U = lane_energy<true>(s, T, M);
    bool fvalid = true;
    const double u_first = U;
    const double inc = u_first * inv_kT + aux, w = (a.cum_in ? a.cum_in[p] : 0.0) + inc;
    if (tl == 0) {
        if (a.inc_out) a.inc_out[p] = inc;
        if (a.w_out) a.w_out[p] = w;
    }
    if (a.n_steps > 0 && (a.flags & CDW_FLAG_RANDOMIZE_VELOCITIES)) {
        lane_noise(s, T, M, 0.0, 1.0, a.seed, gid, (uint32_t)a.step0, CDW_KIND_MAXWELL_BOLTZMANN);
        CDW_TEAM_SYNC(M);
        if (s.n_cons) lane_rattle(s, T, M, a.tol);
    }
    double shadow = 0.0, proposal = 0.0;
    int trials = 0, accepted = 0;
    for (int st = 0; st < a.n_steps; ++st) {
        int o_count = 0;
        for (unsigned long long prog = a.program; prog & 15ull; prog >>= 4) {
            const unsigned tok = (unsigned)(prog & 15ull);
            if (tok == CDW_STEP_OPEN) {
                for (int at = tl; at < A; at += T_) {
                    const int b = CDW_BASE(at);
                    M.xo[b] = M.x[b]; M.xo[b + kR] = M.x[b + kR]; M.xo[b + 2 * kR] = M.x[b + 2 * kR];
                    M.vo[b] = M.v[b]; M.vo[b + kR] = M.v[b + kR]; M.vo[b + 2 * kR] = M.v[b + 2 * kR];
                }
                CDW_TEAM_SYNC(M);
                continue;
            }
            if (tok == CDW_STEP_CLOSE) {
                u4 c;
                c.x = (uint32_t)gid; c.y = (uint32_t)(gid >> 32); c.z = (uint32_t)(a.step0 + st); c.w = CDW_KIND_METROPOLIS << 24;
                const u4 r = philox4x32_10(c, (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
                const double u = (double)uniform24(r.x);
                const bool accept = exp(-shadow * inv_kT) - u >= 0.0;   // step(exp(-shadow_work / kT) - uniform); a NaN shadow work rejects
                ++trials;
                if (accept) ++accepted;
                else {
                    for (int at = tl; at < A; at += T_) {
                        const int b = CDW_BASE(at);
                        M.x[b] = M.xo[b]; M.x[b + kR] = M.xo[b + kR]; M.x[b + 2 * kR] = M.xo[b + 2 * kR];
                        M.v[b] = -M.vo[b]; M.v[b + kR] = -M.vo[b + kR]; M.v[b + 2 * kR] = -M.vo[b + 2 * kR];
                    }
                    fvalid = false;
                }
                shadow = 0.0;
                CDW_TEAM_SYNC(M);
                continue;
            }
            if (tok != CDW_STEP_O && lane_any(M, !fvalid)) {   // V needs the forces, R the energy, at the current positions
                U = lane_energy<true>(s, T, M);             // (decided per WARP: the team barriers inside must stay uniform; re-evaluating
                fvalid = true;                              //  is idempotent for the teams whose forces were still valid)
            }
            const double ke_before = lane_kinetic(s, T, M);
            if (tok == CDW_STEP_V) lane_kick(s, T, M, hv);
            else if (tok == CDW_STEP_R) lane_drift(s, T, M, hr, a.tol);
            else lane_noise(s, T, M, ca, cb, a.seed, gid, (uint32_t)(a.step0 + st), CDW_KIND_O_STEP | ((uint32_t)o_count++ << 4));
            CDW_TEAM_SYNC(M);
            if (s.n_cons) lane_rattle(s, T, M, a.tol);
            const double ke_after = lane_kinetic(s, T, M);
            if (tok == CDW_STEP_V) shadow += ke_after - ke_before;
            else if (tok == CDW_STEP_O) proposal -= ke_after - ke_before;
            else {
                const double U_new = lane_energy<true>(s, T, M);
                shadow += (ke_after + U_new) - (ke_before + U);
                U = U_new;
            }
        }
    }
    if (lane_any(M, !fvalid)) U = lane_energy<true>(s, T, M);
    const double u_last = U;
    const bool bad = a.n_steps > 0 && !(u_last - u_last == 0.0);
    if (tl == 0) {
        M.bad[rid] = bad ? 1 : 0;
        if (a.aux_out) a.aux_out[p] = -(bad ? u_first : u_last) * inv_kT;
        if (a.label_out) a.label_out[p] = label;
        if (a.u_first_out) a.u_first_out[p] = u_first * inv_kT;
        if (a.u_last_out) a.u_last_out[p] = u_last * inv_kT;
        if (a.shadow_out) a.shadow_out[p] = shadow * inv_kT;
        if (a.proposal_out) a.proposal_out[p] = proposal * inv_kT;
        if (a.nan_flags) a.nan_flags[p] = bad ? 1 : 0;
        if (a.trial_counts) { a.trial_counts[2 * p] = trials; a.trial_counts[2 * p + 1] = accepted; }
    }
    return ordered_key(w);
}

// ------------------------------------------------------------------------------------------------ look-ahead iteration
// One SMC iteration with the weights of the NEXT iteration as a by-product (cdw_smc_lookahead).  For an MCMC move the incremental
// work of iteration t + 1 is u_{t+1}(x_t) - u_t(x_t) (distribution_factories.py:124-129 with the aux mechanism of
// openmm/coddiwomple.py:326-327): both energies are taken at the coordinates this launch ends with, so the launch that makes x_t
// can also deliver it, and the resampling of iteration t + 1 then runs BESIDE the launch of iteration t + 1 instead of between two
// launches.  The evaluation that used to open launch t + 1 (forces at lambda_{t+1}) closes launch t instead; the forces travel
// with the particle (one float4 row, gathered through the same ancestor index as x and v).
//
//   part A (tables of lambda_t):      [MB] (V R O R Es Ed V)^n   Es / Ed = static / dynamic terms; Es leaves its forces in fs_out[p]
//   part B (tables of lambda_{t+1}):  Ed' on top of the static forces  ->  fs_out[p] = F(x_t; lambda_{t+1}),  inc_next
//
// A move that ends in a non-finite energy is reverted (propagators.py:166-186): the team reloads its input row and evaluates the
// energies there, so that aux and inc_next are those of the state that is kept.  The redo is decided per WARP (every team of the
// warp repeats the closing evaluation, which is idempotent for the teams that were fine) so that the team barriers stay uniform.
struct AheadState {
    double U_keep, Us;  // u_t at the coordinates that are kept (kJ/mol), and its static part
    int bad;
};

CDW_HD void lane_restore_row(const SysDev& s, const LaneMem& M, const PropArgs& a, long long src) {
    CDW_UNPACK(M);
    for (int at = tl; at < s.n_atoms; at += T_) {
        const int b = CDW_BASE(at);
        const float4 q = row_ptr(a, a.pos_in, a.peer_pos, src, s.n_atoms)[at];
        M.x[b] = q.x; M.x[b + kR] = q.y; M.x[b + 2 * kR] = q.z;
        if (a.vel_in) {
            const float4 w = row_ptr(a, a.vel_in, a.peer_vel, src, s.n_atoms)[at];
            M.v[b] = w.x; M.v[b + kR] = w.y; M.v[b + 2 * kR] = w.z;
        }
    }
    // (no team barrier here: this runs for the reverted teams only; the caller synchronises the whole warp afterwards)
}

CDW_HD AheadState lane_ahead_main(const SysDev& s, const Tables& T, const LaneMem& M, const PropArgs& a, long long p) {
    CDW_UNPACK(M);
    const int A = s.n_atoms;
    const unsigned long long gid = (unsigned long long)(a.gid0 + p);
    const double h = 0.5 * a.dt;
    const double ca = exp(-a.gamma * a.dt), cb = sqrt(1.0 - exp(-2.0 * a.gamma * a.dt));
    enum { OP_V = 0, OP_R = 1, OP_O = 2, OP_ES = 3, OP_ED = 4, OP_MB = 5 };
    const int n_ops = a.n_steps > 0 ? 7 * a.n_steps : 2;   // (V R O R Es Ed V)^n, or Es Ed alone (the call that equips the initial sample)
    double U = 0.0, Us = 0.0;
    bool reverted = false;
    int was_bad = 0;
    int op = (a.n_steps > 0 && (a.flags & CDW_FLAG_RANDOMIZE_VELOCITIES)) ? -1 : 0;
    for (;;) {
        for (; op < n_ops; ++op) {
            int kind, st = 0;
            if (op < 0) kind = OP_MB;
            else if (a.n_steps == 0) kind = op == 0 ? OP_ES : OP_ED;
            else {
                const int slot = op % 7;
                st = op / 7;
                kind = slot == 0 || slot == 6 ? OP_V : (slot == 1 || slot == 3 ? OP_R : (slot == 2 ? OP_O : (slot == 4 ? OP_ES : OP_ED)));
            }
            if (kind == OP_ES || kind == OP_ED) {
                const double e = lane_energy_forces(s, T, M, kind == OP_ES ? PART_STATIC : PART_DYNAMIC);
                if (kind == OP_ED) { U = Us + e; continue; }
                Us = e;
                if (st == a.n_steps - 1 || a.n_steps == 0) {  // static forces (force copy 0 right now) at the final coordinates: part B starts from them
                    for (int at = tl; at < A; at += T_) {
                        const int b = CDW_BASE(at);
                        a.fs_out[p * A + at] = make_float4(M.f[b], M.f[b + kR], M.f[b + 2 * kR], 0.f);
                    }
                }
                continue;
            }
            if (reverted) continue;  // the redo only repeats the closing evaluation
            if (kind == OP_V) lane_kick(s, T, M, h);
            else if (kind == OP_R) lane_drift(s, T, M, h, a.tol);
            else {
                const bool mb = kind == OP_MB;
                lane_noise(s, T, M, mb ? 0.0 : ca, mb ? 1.0 : cb, a.seed, gid, (uint32_t)(a.step0 + st), mb ? CDW_KIND_MAXWELL_BOLTZMANN : CDW_KIND_O_STEP);
            }
            CDW_TEAM_SYNC(M);
            if (s.n_cons) lane_rattle(s, T, M, a.tol);
        }
        const bool bad = a.n_steps > 0 && !reverted && !(U - U == 0.0);  // NaN or inf
        const bool redo = lane_any(M, bad);
        if (bad) {  // (uniform within the team: the team sum leaves the same bits in every lane)
            was_bad = 1;
            lane_restore_row(s, M, a, M.src[rid]);
        }
        if (!redo) break;
        CDW_TEAM_SYNC(M);
        reverted = true;
        op = n_ops - 3;  // Es Ed (V skipped)
    }
    AheadState S;
    S.U_keep = U; S.Us = Us; S.bad = was_bad;
    return S;
}

// part B: the dynamic terms at lambda_{t+1} on top of the static forces saved by part A.
CDW_HD void lane_ahead_next(const SysDev& s, const Tables& T, const LaneMem& M, const PropArgs& a, long long p, const AheadState& S, bool have_next) {
    CDW_UNPACK(M);
    const int A = s.n_atoms;
    const double inv_kT = 1.0 / a.kT;
    const double aux_next = -S.U_keep * inv_kT;  // -u_t(x_t) (openmm/coddiwomple.py:326-327)
    if (have_next) {
        for (int at = tl; at < A; at += T_) {  // (each lane reads back the rows it wrote itself)
            const int b = CDW_BASE(at);
            const float4 q = a.fs_out[p * A + at];
            M.f[b] = q.x; M.f[b + kR] = q.y; M.f[b + 2 * kR] = q.z;
        }
        const double e = lane_energy_forces(s, T, M, PART_DYNAMIC);
        for (int at = tl; at < A; at += T_) {
            const int b = CDW_BASE(at);
            a.fs_out[p * A + at] = make_float4(M.f[b], M.f[b + kR], M.f[b + 2 * kR], 0.f);
        }
        if (tl == 0 && a.inc_next_out) a.inc_next_out[p] = (S.Us + e) * inv_kT + aux_next;  // distribution_factories.py:124-129
    }
    if (tl == 0) {
        if (a.aux_out) a.aux_out[p] = aux_next;
        if (a.u_last_out) a.u_last_out[p] = S.U_keep * inv_kT;
        if (a.nan_flags) a.nan_flags[p] = S.bad;
        if (a.label_out) {
            const long long src = M.src[rid];
            int label = (int)src;
            if (a.peer_pos) {
                const long long r = src / a.n_per_rank, l = src - r * a.n_per_rank;
                if (a.peer_label) label = a.peer_label[r][l];
            } else if (a.label_in) label = a.label_in[src];
            a.label_out[p] = label;
        }
    }
}

// coalesced store of the rows of a look-ahead launch (reverted replicas were restored on chip: no special case)
CDW_HD void cta_store_rows_ahead(const SysDev& s, const PropArgs& a, unsigned char* smem, const Layout& L, int sw, long long p0, int count) {
    const int A = s.n_atoms;
    const float* X = (const float*)(smem + L.x);
    const float* V = (const float*)(smem + L.v);
    CDW_CTA_THREADS(tid, nt) {
        for (int idx = tid; idx < count * A; idx += nt) {
            const int r = idx / A, at = idx - r * A;
            const int b = CDW_BASE_R(at, r);
            if (a.pos_out) a.pos_out[(p0 + r) * A + at] = make_float4(X[b], X[b + kR], X[b + 2 * kR], 0.f);
            if (a.vel_out) a.vel_out[(p0 + r) * A + at] = make_float4(V[b], V[b + kR], V[b + 2 * kR], 0.f);
        }
    }
}

}  // namespace cdw
