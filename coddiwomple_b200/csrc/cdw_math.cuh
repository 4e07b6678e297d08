// cdw_math.cuh — per-term physics, lambda lowering, deterministic exp, Philox.  __host__ __device__ so that the
// same arithmetic can be unit-tested on the build host (tests/host_harness) before it runs on a B200.
//
// Functional forms follow the OpenMM 7.4 System XML the reference feeds to OpenMM
// (coddiwomple/data/perses_data/*.factory.pkl, `_hybrid_system`):
//   bonds (K/2)(r-L)^2, angles (K/2)(theta-T)^2, torsions k(1+cos(n phi - phase)), Coulomb 138.935456 qq / r,
//   LJ 4 eps ((s/r)^12-(s/r)^6), perses soft-core LJ v2 (quadratic cap inside r_LJ), external harmonic well.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define CDW_HD __host__ __device__ __forceinline__
#else
#define CDW_HD inline
#endif

#define CDW_ONE_4PI_EPS0 138.935456

#if defined(__CUDA_ARCH__)
// 1/sqrt(x) for normal positive x (squared distances): hardware seed (MUFU.RSQ64H, ~2^-20) + one third-order
// correction, with NO special-case branch (the library rsqrt() ends in a conditional call that splits basic blocks
// and blocks instruction-level parallelism).  ~1 ulp.
__device__ __forceinline__ double cdw_rsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x * y, y, 1.0);            // 1 - x y^2
    return fma(y * e, fma(e, 0.375, 0.5), y);        // y (1 + e/2 + 3 e^2 / 8)
}
#define CDW_RSQRT(x) cdw_rsqrt(x)
#else
#define CDW_RSQRT(x) (1.0 / sqrt(x))
#endif

namespace cdw {

struct d3 {
    double x, y, z;
};
CDW_HD d3 mk3(double x, double y, double z) { d3 r; r.x = x; r.y = y; r.z = z; return r; }
CDW_HD d3 operator+(d3 a, d3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
CDW_HD d3 operator-(d3 a, d3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
CDW_HD d3 operator*(double s, d3 a) { return mk3(s * a.x, s * a.y, s * a.z); }
CDW_HD double dot(d3 a, d3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
CDW_HD d3 cross(d3 a, d3 b) { return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }

// ---------------------------------------------------------------- effective (lambda-resolved) term records
struct EffBond {
    int i, j;
    double K, L;
};
struct EffAngle {
    int i, j, k, pad;
    double K, T0;
};
struct EffTorsion {
    int i, j, k, l;
    double kk, c0, s0;  // force constant, cos(phase), sin(phase)
    int n, pad;
};
#define CDW_PAIR_NB 1   // Coulomb (+LJ) from NonbondedForce
#define CDW_PAIR_LJ 2   // NonbondedForce epsilon != 0
#define CDW_PAIR_SC 4   // perses soft-core LJ v2 term
#define CDW_PAIR_CAP 8  // r_LJ > 0: quadratic cap is reachable
struct EffPair {
    int i, j, flags, pad;
    // as the pair loop consumes them (prepare_pair): sigma^2 and 4 epsilon instead of sigma and epsilon
    double qqC, sig2, eps4;                        // 138.935456*qq, sigma^2, 4 epsilon
    double sc_sig2, sc_eps4, sc_rlj, sc_F, sc_U;   // soft-core sigma^2, 4 epsilon, r_LJ, Force, U_sterics_cut
};
struct EffExt {
    int atom, pad;
    double K, x0, U0;
};

CDW_HD double affine(double c0, double c1, int slot, const double* lam) { return slot < 0 ? c0 : c0 + lam[slot] * c1; }

// soft-core parameters of one pair (CustomNonbondedForce expression in the fixture XML)
CDW_HD void lower_softcore(EffPair& e, int cls, const double* par /*sigA,epsA,sigB,epsB*/, double l_core, double l_insert,
                           double l_delete, double alpha) {
    double ls = cls == 0 ? l_core : (cls == 1 ? l_insert : l_delete);
    double ldep = cls == 0 ? 0.0 : (cls == 1 ? (1.0 - l_insert) : l_delete);
    double eps = (1.0 - ls) * par[1] + ls * par[3];
    double sig = (1.0 - ls) * par[0] + ls * par[2];
    e.sc_sig2 = sig * sig;
    e.sc_eps4 = 4.0 * eps;
    e.flags |= CDW_PAIR_SC;
    double s2 = sig * sig, s6 = s2 * s2 * s2;
    double arg = (26.0 / 7.0) * s6 * ldep;
    double rlj = arg > 0.0 ? alpha * pow(arg, 1.0 / 6.0) : 0.0;
    e.sc_rlj = rlj;
    e.sc_F = 0.0;
    e.sc_U = 0.0;
    if (rlj > 0.0) {
        double ir = 1.0 / rlj, ir2 = ir * ir, ir6 = ir2 * ir2 * ir2;
        double x = s6 * ir6;
        e.sc_F = -4.0 * eps * (-12.0 * x * x * ir + 6.0 * x * ir);
        e.sc_U = 4.0 * eps * x * (x - 1.0);
        e.flags |= CDW_PAIR_CAP;
    }
}

// ---------------------------------------------------------------- term energies and forces
// Each returns the energy and (if F) the force on every participating atom.

template <bool F>
CDW_HD double bond_term(d3 xi, d3 xj, double K, double L, d3& fi) {
    d3 d = xi - xj;
    double r2 = dot(d, d);
    double ir = CDW_RSQRT(r2);
    double dr = r2 * ir - L;
    if (F) fi = (-K * dr * ir) * d;
    return 0.5 * K * dr * dr;
}

template <bool F>
CDW_HD double angle_term(d3 xi, d3 xj, d3 xk, double K, double T0, d3& fi, d3& fk) {
    d3 a = xi - xj, b = xk - xj;
    double a2 = dot(a, a), b2 = dot(b, b);
    const double iab = CDW_RSQRT(a2 * b2);  // 1 / (|a| |b|)
    double c = dot(a, b) * iab;
    c = c > 1.0 ? 1.0 : (c < -1.0 ? -1.0 : c);
    double th = acos(c);
    double dth = th - T0;
    if (F) {
        d3 cp = cross(a, b);
        double cp2 = dot(cp, cp);
        cp2 = cp2 < 1e-24 ? 1e-24 : cp2;
        const double de = K * dth * CDW_RSQRT(cp2) * (iab * iab);  // 1/a2 = iab^2 b2, 1/b2 = iab^2 a2: no divisions
        fi = (de * b2) * cross(cp, a);
        fk = (de * a2) * cross(b, cp);
    }
    return 0.5 * K * dth * dth;
}

// k (1 + cos(n phi - phase)) with integer n: cos/sin(n phi) by angle-addition recurrence from the cross products,
// no atan2 / sincos.  phi sign convention = OpenMM's (sign of b1 . (b2 x b3) with b's pointing p1->p2->p3->p4).
template <bool F>
CDW_HD double torsion_term(d3 xi, d3 xj, d3 xk, d3 xl, double kk, double c0, double s0, int n, d3& fi, d3& fj, d3& fk, d3& fl) {
    d3 b1 = xj - xi, b2 = xk - xj, b3 = xl - xk;
    d3 n1 = cross(b1, b2), n2 = cross(b2, b3);
    double n1sq = dot(n1, n1), n2sq = dot(n2, n2), b2sq = dot(b2, b2);
    double irb2 = CDW_RSQRT(b2sq);
    double rb2 = b2sq * irb2;
    double inv = CDW_RSQRT(n1sq * n2sq);
    double c = dot(n1, n2) * inv;
    double s = dot(b1, n2) * rb2 * inv;
    double cn = 1.0, sn = 0.0;
    for (int q = 0; q < n; ++q) {
        double t = cn * c - sn * s;
        sn = sn * c + cn * s;
        cn = t;
    }
    if (F) {
        const double de = -kk * (double)n * (sn * c0 - cn * s0) * rb2 * (inv * inv);  // dE/dphi |b2| / (n1sq n2sq)
        fi = (de * n2sq) * n1;    // 1/n1sq = inv^2 n2sq: no divisions
        fl = (-de * n1sq) * n2;
        double ib2 = irb2 * irb2;
        double sa = dot(b1, b2) * ib2, ta = dot(b3, b2) * ib2;
        fj = (-1.0 - sa) * fi + ta * fl;
        fk = (-1.0 - ta) * fl + sa * fi;
    }
    return kk * (1.0 + cn * c0 + sn * s0);
}

// nonbonded pair: returns energy; g = (dE/dr)/r so that force on i is -g * (xi - xj).
// Branch-free over the three parts (Coulomb, plain LJ, soft-core LJ): an absent part has zero coefficients, so all
// lanes of a warp run the same instruction stream whatever pair they hold; only the quadratic cap (r < r_LJ, which
// exists for lambda-softened pairs only) is a select.
template <bool F>
CDW_HD double pair_term(const EffPair& p, d3 d, double& g) {
    const double r2 = dot(d, d);
    const double ir = CDW_RSQRT(r2);
    const double ir2 = ir * ir;
    const double ec = p.qqC * ir;                                   // 138.935456 qq / r
    const double s2 = p.sig2 * ir2, s6 = s2 * s2 * s2;              // plain LJ (NonbondedForce): 4 eps (s6^2 - s6)
    const double t2 = p.sc_sig2 * ir2, x6 = t2 * t2 * t2;           // soft-core LJ outside r_LJ
    const double e = fma(p.eps4, fma(s6, s6, -s6), ec);
    double esc = p.sc_eps4 * fma(x6, x6, -x6);
    double gsc = 0.0;
    if (F) {                                                        // g = -(dU/dr)/r: 4 eps (6 s6 - 12 s6^2) / r^2 - ec / r^2
        g = fma(p.eps4, s6 * fma(-12.0, s6, 6.0), -ec);
        gsc = p.sc_eps4 * (x6 * fma(-12.0, x6, 6.0));
    }
    // quadratic cap inside r_LJ (r_LJ = 0 when the pair is not softened: never taken); selects, not branches, so that
    // two pairs can be evaluated in one basic block (instruction-level parallelism)
    const double dr = fma(r2, ir, -p.sc_rlj);
    const bool cap = dr < 0.0;
    esc = cap ? fma(p.sc_F, dr * fma(0.5, dr, -1.0), p.sc_U) : esc;  // F (dr^2 / 2 - dr) + U_cut
    if (F) {
        const double gcap = p.sc_F * (dr - 1.0) * (r2 * ir);        // expressed per r^2 like the other two: (F (dr - 1) / r) r^2
        g = (g + (cap ? gcap : gsc)) * ir2;
    }
    return e + esc;
}

template <bool F>
CDW_HD double ext_term(d3 x, double K, double x0, double U0, d3& f) {
    d3 d = mk3(x.x - x0, x.y, x.z);
    if (F) f = (-K) * d;
    return 0.5 * K * dot(d, d) + U0;
}

// ---------------------------------------------------------------- deterministic exp (see oracle/resampling.py:det_exp)
#if defined(__CUDA_ARCH__)
#define CDW_MUL(a, b) __dmul_rn((a), (b))
#define CDW_ADD(a, b) __dadd_rn((a), (b))
#else
// host: compile with -ffp-contract=off so these stay separate roundings
#define CDW_MUL(a, b) ((a) * (b))
#define CDW_ADD(a, b) ((a) + (b))
#endif

CDW_HD double bits_to_double(uint64_t b) {
    union { uint64_t u; double d; } c;
    c.u = b;
    return c.d;
}
CDW_HD uint64_t double_to_bits(double d) {
    union { uint64_t u; double d; } c;
    c.d = d;
    return c.u;
}

// exp(x) for x <= 0 from correctly rounded mul/add/rint only: identical bits on CPU and GPU.
CDW_HD double det_exp(double x) {
    if (x < -700.0) return 0.0;
    const double LOG2E = 1.4426950408889634, LN2_HI = 6.93147180369123816490e-01, LN2_LO = 1.90821492927058770002e-10;
    double k = rint(CDW_MUL(x, LOG2E));
    double r = CDW_ADD(CDW_ADD(x, -CDW_MUL(k, LN2_HI)), -CDW_MUL(k, LN2_LO));
    // 1/n!, n = 13..0 (same literals as the oracle computes with np.prod)
    const double c[14] = {1.0, 1.0, 0.5, 1.0 / 6.0, 1.0 / 24.0, 1.0 / 120.0, 1.0 / 720.0, 1.0 / 5040.0, 1.0 / 40320.0,
                          1.0 / 362880.0, 1.0 / 3628800.0, 1.0 / 39916800.0, 1.0 / 479001600.0, 1.0 / 6227020800.0};
    double p = c[13];
#pragma unroll
    for (int n = 12; n >= 0; --n) p = CDW_ADD(CDW_MUL(p, r), c[n]);
    long long ki = (long long)k + 1023;
    ki = ki < 1 ? 1 : (ki > 2046 ? 2046 : ki);
    return CDW_MUL(p, bits_to_double((uint64_t)ki << 52));
}

// order-preserving map double -> uint64 (for atomicMin/Max on doubles)
CDW_HD uint64_t ordered_key(double v) {
    uint64_t b = double_to_bits(v);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
CDW_HD double ordered_key_inverse(uint64_t k) {
    return bits_to_double((k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k);
}

// floor(u * total) for u in [0,1) (53-bit) and total < 2^62, exact in integers; clamped to total-1
CDW_HD uint64_t mulhi64(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * (unsigned __int128)b) >> 64);
#endif
}
CDW_HD uint64_t fixed_target(double u, uint64_t total) {
    uint64_t u53 = (uint64_t)(u * 9007199254740992.0);
    if (u53 > 9007199254740991ull) u53 = 9007199254740991ull;
    uint64_t hi = mulhi64(u53, total), lo = u53 * total;
    uint64_t t = (hi << 11) | (lo >> 53);
    return t < total ? t : total - 1;
}

// ---------------------------------------------------------------- Philox4x32-10 (Salmon et al. SC'11)
struct u4 {
    uint32_t x, y, z, w;
};
CDW_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}
CDW_HD u4 philox4x32_10(u4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = mulhi32(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        uint32_t hi1 = mulhi32(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        u4 n;
        n.x = hi1 ^ c.y ^ k0;
        n.y = lo1;
        n.z = hi0 ^ c.w ^ k1;
        n.w = lo0;
        c = n;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return c;
}
#define CDW_KIND_O_STEP 0u
#define CDW_KIND_MAXWELL_BOLTZMANN 1u
#define CDW_KIND_ULA 2u

CDW_HD float uniform24(uint32_t x) { return ((float)(x >> 8) + 0.5f) * 5.9604644775390625e-08f; }

// three standard normals for (seed; gid, step, atom, kind) — same stream as oracle/rng.py:normals
CDW_HD void normals3(uint64_t seed, uint64_t gid, uint32_t step, uint32_t atom, uint32_t kind, double& z0, double& z1, double& z2) {
    u4 c;
    c.x = (uint32_t)gid;
    c.y = (uint32_t)(gid >> 32);
    c.z = step;
    c.w = atom | (kind << 24);
    u4 r = philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    float r0 = sqrtf(-2.0f * logf(uniform24(r.x)));
    float a0 = 6.283185307179586f * uniform24(r.y);
    float r1 = sqrtf(-2.0f * logf(uniform24(r.z)));
    float a1 = 6.283185307179586f * uniform24(r.w);
    float s0, c0;
#if defined(__CUDA_ARCH__)
    sincosf(a0, &s0, &c0);
#else
    s0 = sinf(a0);
    c0 = cosf(a0);
#endif
    z0 = (double)(r0 * c0);
    z1 = (double)(r0 * s0);
    z2 = (double)(r1 * cosf(a1));
}

#define CDW_MAX_DIRECT 3  // constraint clusters up to this size are projected by a direct k x k solve

// Device-resident raw tables of one replica's force field (global memory) + layout of the per-launch shared memory.
struct SysDev {
    int n_atoms, n_lambda;
    int n_bonds, n_angles, n_torsions, n_pairs, n_ext, n_exceptions, n_cons, n_clusters;
    int n_atom_off, n_exc_off, n_coef;
    int slot_core, slot_insert, slot_delete, slot_alpha;
    int n_bonds_s, n_angles_s, n_torsions_s, n_pairs_s, n_ext_s;  // leading STATIC (lambda-independent) terms of each table
    int a_pad;         // n_atoms rounded up to even (16 B alignment of double rows)
    const double* masses;      // [A]
    const double* inv_mass;    // [A]
    const int* bond_atoms; const int* bond_slot; const double* bond_par;
    const int* angle_atoms; const int* angle_slot; const double* angle_par;
    const int* torsion_atoms; const int* torsion_slot; const int* torsion_n; const double* torsion_par;
    const double* nb_atom_par;
    const int* atom_off_start;  // CSR by atom [A+1]
    const int* atom_off_slot;   // [n_atom_off]
    const double* atom_off_par; // [n_atom_off][3]
    const double* exc_par;
    const int* exc_off_start;   // CSR by exception [n_exc+1]
    const int* exc_off_slot;
    const double* exc_off_par;
    const int* pair_atoms; const int* pair_nb; const int* pair_sc_class; const double* pair_sc_par;
    const int* ext_atom; const int* ext_slot; const double* ext_par;
    // constraints grouped into clusters (connected components); one lane owns one cluster
    const int* cluster_start;   // [n_clusters+1]
    const int* cluster_coef_start;  // [n_clusters+1]: offset of the k x k coupling coefficients (k <= CDW_MAX_DIRECT), else -1
    const double* cluster_coef;     // c_qr = d(iq,ir)/m_ir - d(iq,jr)/m_jr - d(jq,ir)/m_ir + d(jq,jr)/m_jr
    const int* cons_atoms;      // [n_cons][2], sorted by cluster
    const double* cons_d2;      // [n_cons] d^2
};


}  // namespace cdw
