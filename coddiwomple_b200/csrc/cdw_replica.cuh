// cdw_replica.cuh — what ONE warp does for ONE replica, written as lane-parallel phases.
//
// On the device a phase `CDW_LANES(lane) { ... }` runs once per thread with lane = threadIdx.x & 31 and ends in
// __syncwarp(); on the build host (tests/host_harness, no GPU) the same phase is a sequential loop over the 32
// lanes, so the whole per-replica algorithm — lambda lowering, term evaluation, scratch-slot force gather, SHAKE /
// RATTLE, BAOAB, Philox noise, NaN guard — is unit-tested against the fp64 oracle before it ever runs on a B200.
// Values that cross a phase boundary live in (shared) memory arrays, never in lane-private registers, except
// per-lane partial sums, which are combined with CDW_WARP_SUM at the end of the phase that produced them.
//
// Reference semantics: OMMLI "V R O R V" (coddiwomple/openmm/integrators.py:297-350) driven by OMMBIP.apply
// (coddiwomple/openmm/propagators.py:127-219); reduced potential (coddiwomple/openmm/states.py:116-136);
// incremental work (coddiwomple/distribution_factories.py:124-129).
#pragma once
#include "cdw_math.cuh"

#if defined(__CUDA_ARCH__)
#define CDW_LANES(lane) for (int lane = (int)(threadIdx.x & 31u), _cdw_once = 1; _cdw_once; _cdw_once = 0, __syncwarp())
#define CDW_CTA_THREADS(tid, nt) for (int tid = (int)threadIdx.x, nt = (int)blockDim.x, _cdw_once = 1; _cdw_once; _cdw_once = 0, __syncthreads())
#define CDW_WARP_SUM(v) cdw::warp_sum_dev(v)
#define CDW_LDG(p) __ldg(p)
#define CDW_NOINLINE __device__ __noinline__
#else
#define CDW_LANES(lane) for (int lane = 0; lane < 32; ++lane)
#define CDW_CTA_THREADS(tid, nt) for (int tid = 0, nt = 1; tid < 1; ++tid)
#define CDW_WARP_SUM(v) (v)
#define CDW_LDG(p) (*(p))
#define CDW_NOINLINE inline
#endif
#if !defined(__CUDACC__)
struct alignas(16) float4 {
    float x, y, z, w;
};
inline float4 make_float4(float x, float y, float z, float w) { float4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
#endif

namespace cdw {

#if defined(__CUDA_ARCH__)
__device__ __forceinline__ double warp_sum_dev(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
#endif

#ifndef CDW_KWARPS
#define CDW_KWARPS 8
#endif
#ifndef CDW_MINBLOCKS
#define CDW_MINBLOCKS 2
#endif
constexpr int kWarps = CDW_KWARPS;  // replicas in flight per CTA
constexpr int kThreads = kWarps * 32;
constexpr int kMaxConstraintIters = 500;

// ------------------------------------------------------------------------------------------------ shared-memory layout
struct Layout {
    size_t lam, nbeff, inv_m, sig_v, bonds, angles, torsions, pairs, exts, slot_start, slots, cl_start, cons_atoms, cons_d2, cl_coef_start, cl_coef;  // CTA-shared
    size_t warp0, warp_stride, x, v, f, xref, fs;                                                                                // per warp
    int ns_pad;
    size_t total;
};

CDW_HD size_t align16(size_t v) { return (v + 15) & ~size_t(15); }

CDW_HD Layout make_layout(const SysDev& s, bool forces, int warps) {
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
    L.slot_start = o; o = align16(o + sizeof(int) * (forces ? A + 1 : 0));
    L.slots = o; o = align16(o + sizeof(int) * (forces ? s.n_slots : 0));
    L.cl_start = o; o = align16(o + sizeof(int) * (forces ? s.n_clusters + 1 : 0));
    L.cons_atoms = o; o = align16(o + sizeof(int) * (forces ? 2 * s.n_cons : 0));
    L.cons_d2 = o; o = align16(o + sizeof(double) * (forces ? s.n_cons : 0));
    L.cl_coef_start = o; o = align16(o + sizeof(int) * (forces ? s.n_clusters + 1 : 0));
    L.cl_coef = o; o = align16(o + sizeof(double) * (forces ? s.n_coef : 0));
    L.warp0 = o;
    size_t w = 0;
    const size_t row = sizeof(double) * 3 * s.a_pad;
    L.x = w; w += row;
    L.ns_pad = (s.n_slots + 3) & ~3;
    L.v = L.f = L.xref = L.fs = 0;
    if (forces) {
        L.v = w; w += row;
        L.f = w; w += row;
        L.xref = w; w += row;
        L.fs = w; w += sizeof(float) * 3 * L.ns_pad;
    }
    L.warp_stride = align16(w);
    L.total = L.warp0 + L.warp_stride * warps;
    return L;
}

struct Tables {  // pointers into the CTA-shared part
    double* lam; double* nbeff; double* inv_m; double* sig_v;
    EffBond* bonds; EffAngle* angles; EffTorsion* torsions; EffPair* pairs; EffExt* exts;
    int* slot_start; int* slots; int* cl_start; int* cons_atoms; double* cons_d2; int* cl_coef_start; double* cl_coef;
};

struct WarpMem {  // pointers into one warp's private part
    double* x; double* v; double* f; double* xref; float* fs;
};

CDW_HD Tables table_pointers(const Layout& L, unsigned char* smem) {
    Tables T;
    T.lam = (double*)(smem + L.lam); T.nbeff = (double*)(smem + L.nbeff); T.inv_m = (double*)(smem + L.inv_m); T.sig_v = (double*)(smem + L.sig_v);
    T.bonds = (EffBond*)(smem + L.bonds); T.angles = (EffAngle*)(smem + L.angles); T.torsions = (EffTorsion*)(smem + L.torsions);
    T.pairs = (EffPair*)(smem + L.pairs); T.exts = (EffExt*)(smem + L.exts);
    T.slot_start = (int*)(smem + L.slot_start); T.slots = (int*)(smem + L.slots);
    T.cl_start = (int*)(smem + L.cl_start); T.cons_atoms = (int*)(smem + L.cons_atoms); T.cons_d2 = (double*)(smem + L.cons_d2);
    T.cl_coef_start = (int*)(smem + L.cl_coef_start); T.cl_coef = (double*)(smem + L.cl_coef);
    return T;
}

CDW_HD WarpMem warp_pointers(const Layout& L, unsigned char* smem, int warp) {
    unsigned char* wb = smem + L.warp0 + L.warp_stride * warp;
    WarpMem W;
    W.x = (double*)(wb + L.x); W.v = (double*)(wb + L.v); W.f = (double*)(wb + L.f); W.xref = (double*)(wb + L.xref); W.fs = (float*)(wb + L.fs);
    return W;
}

// ------------------------------------------------------------------------------------------------ prologue (once per CTA)
// Resolve the lambda vector into effective term records: the per-replica loops never touch lambda again.
CDW_HD void build_tables(const SysDev& s, const Tables& T, const double* lam_g, double kT, bool forces) {
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
            for (int q = tid; q < s.n_atoms + 1; q += nt) T.slot_start[q] = s.atom_slot_start[q];
            for (int q = tid; q < s.n_slots; q += nt) T.slots[q] = s.atom_slots[q];
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
            p.qqC = p.sig = p.eps = 0.0;
            p.sc_sig = p.sc_eps = p.sc_rlj = p.sc_F = p.sc_U = 0.0;
            int nbk = s.pair_nb[q];
            if (nbk == -2) {
                p.qqC = CDW_ONE_4PI_EPS0 * (T.nbeff[3 * p.i] * T.nbeff[3 * p.j]);
                p.sig = 0.5 * (T.nbeff[3 * p.i + 1] + T.nbeff[3 * p.j + 1]);
                p.eps = sqrt(T.nbeff[3 * p.i + 2] * T.nbeff[3 * p.j + 2]);
                p.flags |= CDW_PAIR_NB;
            } else if (nbk >= 0) {
                double qq = s.exc_par[3 * nbk], sg = s.exc_par[3 * nbk + 1], ep = s.exc_par[3 * nbk + 2];
                for (int o = s.exc_off_start[nbk]; o < s.exc_off_start[nbk + 1]; ++o) {
                    double l = lam[s.exc_off_slot[o]];
                    qq += l * s.exc_off_par[3 * o]; sg += l * s.exc_off_par[3 * o + 1]; ep += l * s.exc_off_par[3 * o + 2];
                }
                p.qqC = CDW_ONE_4PI_EPS0 * qq; p.sig = sg; p.eps = ep;
                p.flags |= CDW_PAIR_NB;
            }
            if (p.eps != 0.0) p.flags |= CDW_PAIR_LJ;
            int cls = s.pair_sc_class[q];
            if (cls >= 0)
                lower_softcore(p, cls, s.pair_sc_par + 4 * q, lam[s.slot_core], lam[s.slot_insert], lam[s.slot_delete], lam[s.slot_alpha]);
            T.pairs[q] = p;
        }
    }
}

// ------------------------------------------------------------------------------------------------ energy / forces
// Lane `lane` of the warp evaluates its share of the terms of one replica.  x: [3][a_pad] doubles.  Returns the
// lane's partial energy.  With F, per-term forces go to the fp32 scratch fs: [3][ns_pad] (one slot per
// (term, atom-in-term), laid out so that consecutive lanes write consecutive words: no bank conflicts, no atomics).
template <bool F>
CDW_HD double eval_terms(const SysDev& s, const Tables& T, const double* x, float* fs, int ns_pad, int lane) {
    const int Ap = s.a_pad;
    double e = 0.0;
    const int nb = s.n_bonds, na = s.n_angles, nt = s.n_torsions, np = s.n_pairs, ne = s.n_ext;
    const int base_a = 2 * nb, base_t = base_a + 3 * na, base_p = base_t + 4 * nt, base_e = base_p + 2 * np;
#define CDW_X(a) mk3(x[(a)], x[Ap + (a)], x[2 * Ap + (a)])
#define CDW_PUT(slot, fv)                         \
    do {                                          \
        fs[(slot)] = (float)(fv).x;               \
        fs[ns_pad + (slot)] = (float)(fv).y;      \
        fs[2 * ns_pad + (slot)] = (float)(fv).z;  \
    } while (0)
    for (int q = lane; q < nb; q += 32) {
        EffBond b = T.bonds[q];
        d3 fi;
        e += bond_term<F>(CDW_X(b.i), CDW_X(b.j), b.K, b.L, fi);
        if (F) { d3 fj = -1.0 * fi; CDW_PUT(q, fi); CDW_PUT(nb + q, fj); }
    }
    for (int q = lane; q < na; q += 32) {
        EffAngle a = T.angles[q];
        d3 fi, fk;
        e += angle_term<F>(CDW_X(a.i), CDW_X(a.j), CDW_X(a.k), a.K, a.T0, fi, fk);
        if (F) { d3 fj = -1.0 * (fi + fk); CDW_PUT(base_a + q, fi); CDW_PUT(base_a + na + q, fj); CDW_PUT(base_a + 2 * na + q, fk); }
    }
    for (int q = lane; q < nt; q += 32) {
        EffTorsion t = T.torsions[q];
        d3 fi, fj, fk, fl;
        e += torsion_term<F>(CDW_X(t.i), CDW_X(t.j), CDW_X(t.k), CDW_X(t.l), t.kk, t.c0, t.s0, t.n, fi, fj, fk, fl);
        if (F) { CDW_PUT(base_t + q, fi); CDW_PUT(base_t + nt + q, fj); CDW_PUT(base_t + 2 * nt + q, fk); CDW_PUT(base_t + 3 * nt + q, fl); }
    }
    for (int q = lane; q < np; q += 32) {
        const EffPair& p = T.pairs[q];
        d3 d = CDW_X(p.i) - CDW_X(p.j);
        double g;
        e += pair_term<F>(p, d, g);
        if (F) { d3 fi = (-g) * d; d3 fj = g * d; CDW_PUT(base_p + q, fi); CDW_PUT(base_p + np + q, fj); }
    }
    for (int q = lane; q < ne; q += 32) {
        EffExt w = T.exts[q];
        d3 f;
        e += ext_term<F>(CDW_X(w.atom), w.K, w.x0, w.U0, f);
        if (F) CDW_PUT(base_e + q, f);
    }
#undef CDW_X
#undef CDW_PUT
    return e;
}

CDW_HD void gather_forces(const SysDev& s, const Tables& T, const float* fs, int ns_pad, double* f, int lane) {
    const int Ap = s.a_pad;
    for (int a = lane; a < s.n_atoms; a += 32) {
        double fx = 0.0, fy = 0.0, fz = 0.0;
        for (int q = T.slot_start[a]; q < T.slot_start[a + 1]; ++q) {
            int sl = T.slots[q];
            fx += (double)fs[sl]; fy += (double)fs[ns_pad + sl]; fz += (double)fs[2 * ns_pad + sl];
        }
        f[a] = fx; f[Ap + a] = fy; f[2 * Ap + a] = fz;
    }
}

// potential energy of the replica in W.x (kJ/mol, warp-uniform); with F also the forces in W.f
template <bool F>
CDW_HD double energy_forces(const SysDev& s, const Tables& T, const WarpMem& W, int ns_pad) {
    double e = 0.0;
    CDW_LANES(lane) { e += eval_terms<F>(s, T, W.x, W.fs, ns_pad, lane); }
    if (F) {
        CDW_LANES(lane) { gather_forces(s, T, W.fs, ns_pad, W.f, lane); }
    }
    return CDW_WARP_SUM(e);
}

// ------------------------------------------------------------------------------------------------ constraints
struct ConsCtx {  // what the constraint routines need, passed BY VALUE so the noinline calls keep everything in registers
    const int* cl_start; const int* cons_atoms; const int* cl_coef_start; const double* cl_coef; const double* cons_d2; const double* inv_m;
    int n_clusters, Ap;
};
CDW_HD ConsCtx cons_ctx(const SysDev& s, const Tables& T) {
    ConsCtx c;
    c.cl_start = T.cl_start; c.cons_atoms = T.cons_atoms; c.cl_coef_start = T.cl_coef_start; c.cl_coef = T.cl_coef; c.cons_d2 = T.cons_d2; c.inv_m = T.inv_m;
    c.n_clusters = s.n_cons ? s.n_clusters : 0; c.Ap = s.a_pad;
    return c;
}

// One lane per constraint cluster (connected component of the constraint graph): Gauss-Seidel sweeps in constraint
// order until every distance is inside OpenMM's (1 +- tol)^2 window.
CDW_NOINLINE void shake(const ConsCtx T, double* x, const double* xref, double tol) {
    if (T.n_clusters == 0) return;
    const int Ap = T.Ap;
    const double lower = 1.0 - 2.0 * tol + tol * tol, upper = 1.0 + 2.0 * tol + tol * tol;
    CDW_LANES(lane) {
        for (int c = lane; c < T.n_clusters; c += 32) {
            for (int it = 0; it < kMaxConstraintIters; ++it) {
                bool updated = false;
                for (int q = T.cl_start[c]; q < T.cl_start[c + 1]; ++q) {
                    int i = T.cons_atoms[2 * q], j = T.cons_atoms[2 * q + 1];
                    double d2 = T.cons_d2[q];
                    d3 sv = mk3(x[i] - x[j], x[Ap + i] - x[Ap + j], x[2 * Ap + i] - x[2 * Ap + j]);
                    double r2 = dot(sv, sv);
                    if (r2 < lower * d2 || r2 > upper * d2) {
                        d3 r0 = mk3(xref[i] - xref[j], xref[Ap + i] - xref[Ap + j], xref[2 * Ap + i] - xref[2 * Ap + j]);
                        double imi = T.inv_m[i], imj = T.inv_m[j];
                        double g = (d2 - r2) / (2.0 * (imi + imj) * dot(sv, r0));
                        x[i] += g * imi * r0.x; x[Ap + i] += g * imi * r0.y; x[2 * Ap + i] += g * imi * r0.z;
                        x[j] -= g * imj * r0.x; x[Ap + j] -= g * imj * r0.y; x[2 * Ap + j] -= g * imj * r0.z;
                        updated = true;
                    }
                }
                if (!updated) break;
            }
        }
    }
}

// Velocity constraints are linear: a cluster of k <= CDW_MAX_DIRECT coupled constraints is projected EXACTLY by one
// k x k symmetric positive-definite solve, A_qr = (s_q . s_r) c_qr, A g = -(s_q . v_rel,q); no sweeps, no tolerance.
template <int K>
CDW_HD void rattle_cluster(const ConsCtx& T, const double* x, double* v, int Ap, int q0, const double* coef) {
    d3 sv[K];
    double b[K], A[K][K];
    int ia[K], ja[K];
#pragma unroll
    for (int q = 0; q < K; ++q) {
        const int i = T.cons_atoms[2 * (q0 + q)], j = T.cons_atoms[2 * (q0 + q) + 1];
        ia[q] = i; ja[q] = j;
        sv[q] = mk3(x[i] - x[j], x[Ap + i] - x[Ap + j], x[2 * Ap + i] - x[2 * Ap + j]);
        b[q] = -dot(sv[q], mk3(v[i] - v[j], v[Ap + i] - v[Ap + j], v[2 * Ap + i] - v[2 * Ap + j]));
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
    for (int q = 0; q < K; ++q) {
        const int i = ia[q], j = ja[q];
        const double gi = b[q] * T.inv_m[i], gj = b[q] * T.inv_m[j];
        v[i] += gi * sv[q].x; v[Ap + i] += gi * sv[q].y; v[2 * Ap + i] += gi * sv[q].z;
        v[j] -= gj * sv[q].x; v[Ap + j] -= gj * sv[q].y; v[2 * Ap + j] -= gj * sv[q].z;
    }
}

CDW_NOINLINE void rattle_v(const ConsCtx T, const double* x, double* v, double tol) {
    if (T.n_clusters == 0) return;
    const int Ap = T.Ap;
    CDW_LANES(lane) {
        for (int c = lane; c < T.n_clusters; c += 32) {
            const int q0 = T.cl_start[c], k = T.cl_start[c + 1] - q0;
            const double* coef = T.cl_coef + T.cl_coef_start[c];
            if (k == 1) rattle_cluster<1>(T, x, v, Ap, q0, coef);
            else if (k == 2) rattle_cluster<2>(T, x, v, Ap, q0, coef);
            else if (k == 3) rattle_cluster<3>(T, x, v, Ap, q0, coef);
            else {
                // large clusters: Gauss-Seidel sweeps to |s . v_rel| <= tol d^2
                for (int it = 0; it < kMaxConstraintIters; ++it) {
                    bool updated = false;
                    for (int q = q0; q < q0 + k; ++q) {
                        int i = T.cons_atoms[2 * q], j = T.cons_atoms[2 * q + 1];
                        double d2 = T.cons_d2[q];
                        d3 sv = mk3(x[i] - x[j], x[Ap + i] - x[Ap + j], x[2 * Ap + i] - x[2 * Ap + j]);
                        d3 vr = mk3(v[i] - v[j], v[Ap + i] - v[Ap + j], v[2 * Ap + i] - v[2 * Ap + j]);
                        double sd = dot(sv, vr);
                        if (fabs(sd) > tol * d2) {
                            double imi = T.inv_m[i], imj = T.inv_m[j];
                            double g = -sd / ((imi + imj) * dot(sv, sv));
                            v[i] += g * imi * sv.x; v[Ap + i] += g * imi * sv.y; v[2 * Ap + i] += g * imi * sv.z;
                            v[j] -= g * imj * sv.x; v[Ap + j] -= g * imj * sv.y; v[2 * Ap + j] -= g * imj * sv.z;
                            updated = true;
                        }
                    }
                    if (!updated) break;
                }
            }
        }
    }
}

CDW_HD double kinetic_energy(const SysDev& s, const Tables& T, const double* v) {
    const int Ap = s.a_pad;
    double k = 0.0;
    CDW_LANES(lane) {
        for (int a = lane; a < s.n_atoms; a += 32) k += (v[a] * v[a] + v[Ap + a] * v[Ap + a] + v[2 * Ap + a] * v[2 * Ap + a]) / T.inv_m[a];
    }
    return 0.5 * CDW_WARP_SUM(k);
}

CDW_HD void kick(const SysDev& s, const Tables& T, const WarpMem& W, double h) {
    const int Ap = s.a_pad;
    CDW_LANES(lane) {
        for (int at = lane; at < s.n_atoms; at += 32) {
            double c = h * T.inv_m[at];
            W.v[at] += c * W.f[at]; W.v[Ap + at] += c * W.f[Ap + at]; W.v[2 * Ap + at] += c * W.f[2 * Ap + at];
        }
    }
}

// ------------------------------------------------------------------------------------------------ one replica, one call
struct PropArgs {
    const float4* pos_in; const float4* vel_in; float4* pos_out; float4* vel_out;
    long long n;
    const int* anc; const double* aux_in; const double* cum_in; const int* label_in;
    const double* lambda;
    double dt, gamma, kT, tol;
    int n_steps; unsigned flags;
    unsigned long long seed, step0; long long gid0;
    double* inc_out; double* w_out; double* aux_out; int* label_out;
    double* u_first_out; double* u_last_out; double* shadow_out; double* proposal_out;
    int* nan_flags; unsigned long long* wmin_key;
    double* energy_out; float4* force_out;  // cdw_energy_forces
    double beta;                             // cdw_reduced_potential
};

// K1 body: u[p] = beta * U(x_p; lambda)
CDW_HD void replica_potential(const SysDev& s, const Tables& T, const WarpMem& W, const PropArgs& a, long long p) {
    const int A = s.n_atoms, Ap = s.a_pad;
    CDW_LANES(lane) {
        for (int at = lane; at < A; at += 32) {
            float4 q = CDW_LDG(a.pos_in + p * A + at);
            W.x[at] = (double)q.x; W.x[Ap + at] = (double)q.y; W.x[2 * Ap + at] = (double)q.z;
        }
    }
    double e = energy_forces<false>(s, T, W, 0);
    CDW_LANES(lane) {
        if (lane == 0) a.energy_out[p] = a.beta * e;
    }
}

// K4 body (+ gather-on-load + first half of K2).  Returns the order-preserving key of W[p] (for the running min).
CDW_HD unsigned long long replica_step(const SysDev& s, const Tables& T, const WarpMem& W, int ns, const PropArgs& a, long long p) {
    const int A = s.n_atoms, Ap = s.a_pad;
    const long long src = a.anc ? (long long)a.anc[p] : p;
    const unsigned long long gid = (unsigned long long)(a.gid0 + p);
    const double h = 0.5 * a.dt;
    const double ca = exp(-a.gamma * a.dt), cb = sqrt(1.0 - exp(-2.0 * a.gamma * a.dt));
    const bool track = (a.flags & CDW_FLAG_TRACK_WORKS) != 0;
    const bool randomize = (a.flags & CDW_FLAG_RANDOMIZE_VELOCITIES) != 0;
    const double inv_kT = 1.0 / a.kT;
    double* x = W.x; double* v = W.v; double* f = W.f; double* xref = W.xref;

    // ---- coalesced float4 load of the ancestor's row (this IS the resampling gather, resamplers.py:123-129)
    CDW_LANES(lane) {
        for (int at = lane; at < A; at += 32) {
            float4 q = CDW_LDG(a.pos_in + src * A + at);
            x[at] = (double)q.x; x[Ap + at] = (double)q.y; x[2 * Ap + at] = (double)q.z;
            if (randomize) {  // Context.setVelocitiesToTemperature (propagators.py:137-138)
                double z0, z1, z2;
                normals3(a.seed, gid, (uint32_t)a.step0, (uint32_t)at, CDW_KIND_MAXWELL_BOLTZMANN, z0, z1, z2);
                double sg = T.sig_v[at];
                v[at] = sg * z0; v[Ap + at] = sg * z1; v[2 * Ap + at] = sg * z2;
            } else {
                float4 w = a.vel_in ? CDW_LDG(a.vel_in + src * A + at) : make_float4(0.f, 0.f, 0.f, 0.f);
                v[at] = (double)w.x; v[Ap + at] = (double)w.y; v[2 * Ap + at] = (double)w.z;
            }
        }
    }
    const ConsCtx cc = cons_ctx(s, T);
    if (randomize) rattle_v(cc, x, v, a.tol);
    double U = 0.0, u_first = 0.0, shadow = 0.0, proposal = 0.0, U_mid = 0.0;
    // One force evaluation per pass (a single call site keeps the kernel inside the instruction cache):
    //   pass 0     : E1, F1 at (x_{t-1}, lambda_t) -> the incremental work (distribution_factories.py:120-129)
    //   pass st + 1: closes step st (second half-kick), i.e. "V R O R" of step st happen between pass st and st + 1.
    for (int pass = 0;; ++pass) {
        U = energy_forces<true>(s, T, W, ns);
        if (pass == 0) {
            u_first = U;
            if (a.force_out) {
                CDW_LANES(lane) {
                    for (int at = lane; at < A; at += 32) a.force_out[p * A + at] = make_float4((float)f[at], (float)f[Ap + at], (float)f[2 * Ap + at], 0.f);
                }
            }
        } else {
            if (track) shadow += kinetic_energy(s, T, v) + U;  // closes the bookkeeping of the second R (its -(KE + U_mid) was added below)
            // closing V (integrators.py:317-336)
            double ke0 = track ? kinetic_energy(s, T, v) : 0.0;
            kick(s, T, W, h);
            rattle_v(cc, x, v, a.tol);
            if (track) shadow += kinetic_energy(s, T, v) - ke0;
        }
        if (pass == a.n_steps) break;
        const int st = pass;
        const bool last = st == a.n_steps - 1;
        // V
        double ke0 = track ? kinetic_energy(s, T, v) : 0.0;
        kick(s, T, W, h);
        rattle_v(cc, x, v, a.tol);
        if (track) shadow += kinetic_energy(s, T, v) - ke0;
        for (int half = 0; half < 2; ++half) {
            // R (integrators.py:297-315)
            if (track) shadow -= kinetic_energy(s, T, v) + (half == 0 ? U : U_mid);
            CDW_LANES(lane) {
                for (int at = lane; at < A; at += 32) {
                    for (int c = 0; c < 3; ++c) {
                        double xr = x[c * Ap + at];
                        xref[c * Ap + at] = xr;
                        x[c * Ap + at] = xr + h * v[c * Ap + at];
                    }
                }
            }
            if (s.n_cons) {
                shake(cc, x, xref, a.tol);
                // v += (x - x1)/(dt/2) with x1 = xref + (dt/2) v   <=>   v = (x - xref)/(dt/2)
                const double ih = 1.0 / h;
                CDW_LANES(lane) {
                    for (int at = lane; at < A; at += 32)
                        for (int c = 0; c < 3; ++c) v[c * Ap + at] = (x[c * Ap + at] - xref[c * Ap + at]) * ih;
                }
                rattle_v(cc, x, v, a.tol);
            }
            if (half == 1 && last) {
                // particles are stored as float4: evaluate the closing energy at exactly the stored coordinates
                CDW_LANES(lane) {
                    for (int at = lane; at < A; at += 32)
                        for (int c = 0; c < 3; ++c) x[c * Ap + at] = (double)(float)x[c * Ap + at];
                }
            }
            if (half == 0) {
                if (track) {
                    U_mid = energy_forces<false>(s, T, W, ns);
                    shadow += kinetic_energy(s, T, v) + U_mid;
                }
                // O (integrators.py:338-350)
                double k0 = track ? kinetic_energy(s, T, v) : 0.0;
                CDW_LANES(lane) {
                    for (int at = lane; at < A; at += 32) {
                        double z0, z1, z2;
                        normals3(a.seed, gid, (uint32_t)(a.step0 + st), (uint32_t)at, CDW_KIND_O_STEP, z0, z1, z2);
                        double sg = cb * T.sig_v[at];
                        v[at] = ca * v[at] + sg * z0; v[Ap + at] = ca * v[Ap + at] + sg * z1; v[2 * Ap + at] = ca * v[2 * Ap + at] + sg * z2;
                    }
                }
                rattle_v(cc, x, v, a.tol);
                if (track) proposal -= kinetic_energy(s, T, v) - k0;
            }
        }
        // the second R's  +(KE + U_new)  is added at the top of the next pass, once U_new is known
    }
    const double u_last = U;
    const bool bad = a.n_steps > 0 && !(u_last - u_last == 0.0);  // NaN or inf
    // ---- store; NaN guard reverts to the pre-move state (propagators.py:166-186)
    if (a.pos_out) {
        CDW_LANES(lane) {
            for (int at = lane; at < A; at += 32) {
                float4 q, w;
                if (!bad) {
                    q = make_float4((float)x[at], (float)x[Ap + at], (float)x[2 * Ap + at], 0.f);
                    w = make_float4((float)v[at], (float)v[Ap + at], (float)v[2 * Ap + at], 0.f);
                } else {
                    q = CDW_LDG(a.pos_in + src * A + at);
                    w = a.vel_in ? CDW_LDG(a.vel_in + src * A + at) : make_float4(0.f, 0.f, 0.f, 0.f);  // the caller's state
                }
                a.pos_out[p * A + at] = q;
                if (a.vel_out) a.vel_out[p * A + at] = w;
            }
        }
    }
    const double aux = a.aux_in ? a.aux_in[src] : 0.0;
    const double inc = u_first * inv_kT + aux;
    const double w = (a.cum_in ? a.cum_in[p] : 0.0) + inc;
    CDW_LANES(lane) {
        if (lane == 0) {
            if (a.inc_out) a.inc_out[p] = inc;
            if (a.w_out) a.w_out[p] = w;
            if (a.aux_out) a.aux_out[p] = -(bad ? u_first : u_last) * inv_kT;
            if (a.label_out) a.label_out[p] = a.label_in ? a.label_in[src] : (int)src;
            if (a.u_first_out) a.u_first_out[p] = u_first * inv_kT;
            if (a.u_last_out) a.u_last_out[p] = u_last * inv_kT;
            if (a.shadow_out) a.shadow_out[p] = shadow * inv_kT;
            if (a.proposal_out) a.proposal_out[p] = proposal * inv_kT;
            if (a.nan_flags) a.nan_flags[p] = bad ? 1 : 0;
            if (a.energy_out) a.energy_out[p] = u_first;
        }
    }
    return ordered_key(w);
}

}  // namespace cdw
