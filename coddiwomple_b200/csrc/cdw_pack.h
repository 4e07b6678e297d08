// cdw_pack.h — host-side construction of one replica's device tables (pure C++, no CUDA): validation, CSR of the
// lambda offsets, constraint clusters, static/dynamic term split; everything packed into ONE blob so that the same
// bytes serve the GPU (cudaMemcpy) and the host test harness.
#pragma once
#include <string.h>

#include <algorithm>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/cdw.h"
#include "cdw_math.cuh"

namespace cdw {

struct Packer {
    std::vector<unsigned char> buf;
    template <typename T>
    size_t add(const std::vector<T>& v) {
        size_t off = (buf.size() + 15) & ~size_t(15);
        buf.resize(off + std::max<size_t>(v.size() * sizeof(T), 16));
        if (!v.empty()) memcpy(buf.data() + off, v.data(), v.size() * sizeof(T));
        return off;
    }
};

template <typename T>
static std::vector<T> vec(const T* p, size_t n) {
    return p && n ? std::vector<T>(p, p + n) : std::vector<T>();
}

#define CDW_PACK_REQUIRE(cond, msg) \
    do {                            \
        if (!(cond)) {              \
            err = msg;              \
            return CDW_EINVAL;      \
        }                           \
    } while (0)

// Fills `blob` and a SysDev whose pointers are OFFSETS into the blob (rebase with rebase_system()).
inline int pack_system(const cdw_system_desc* d, std::vector<unsigned char>& blob, SysDev& v, std::string& err) {
    CDW_PACK_REQUIRE(d != nullptr, "null argument");

    CDW_PACK_REQUIRE(d->n_atoms > 0 && d->n_lambda >= 0 && d->n_lambda <= 64, "n_atoms must be > 0 and n_lambda in [0, 64]");
    const int A = d->n_atoms;
    auto in_range = [&](const int32_t* idx, int n, int width) {
        for (int q = 0; q < n * width; ++q)
            if (idx[q] < 0 || idx[q] >= A) return false;
        return true;
    };
    CDW_PACK_REQUIRE(in_range(d->bond_atoms, d->n_bonds, 2) && in_range(d->angle_atoms, d->n_angles, 3) &&
                    in_range(d->torsion_atoms, d->n_torsions, 4) && in_range(d->pair_atoms, d->n_pairs, 2) &&
                    in_range(d->ext_atom, d->n_ext, 1) && in_range(d->cons_atoms, d->n_constraints, 2),
                "atom index out of range");
    auto slots_ok = [&](const int32_t* s, int n) {
        for (int q = 0; q < n; ++q)
            if (s[q] >= d->n_lambda) return false;
        return true;
    };
    CDW_PACK_REQUIRE(slots_ok(d->bond_slot, d->n_bonds) && slots_ok(d->angle_slot, d->n_angles) && slots_ok(d->torsion_slot, d->n_torsions) &&
                    slots_ok(d->ext_slot, 3 * d->n_ext),
                "lambda slot out of range");
    for (int q = 0; q < d->n_torsions; ++q) CDW_PACK_REQUIRE(d->torsion_n[q] >= 0 && d->torsion_n[q] <= 6, "torsion periodicity must be 0..6");
    for (int q = 0; q < A; ++q) CDW_PACK_REQUIRE(d->masses[q] > 0.0, "massless particles are not supported");
    bool any_sc = false;
    for (int q = 0; q < d->n_pairs; ++q) {
        CDW_PACK_REQUIRE(d->pair_nb[q] >= -2 && d->pair_nb[q] < d->n_exceptions, "pair_nb out of range");
        CDW_PACK_REQUIRE(d->pair_sc_class[q] >= -1 && d->pair_sc_class[q] <= 2, "pair_sc_class out of range");
        any_sc |= d->pair_sc_class[q] >= 0;
    }
    if (any_sc)
        CDW_PACK_REQUIRE(d->slot_sterics_core >= 0 && d->slot_sterics_core < d->n_lambda && d->slot_sterics_insert >= 0 &&
                        d->slot_sterics_insert < d->n_lambda && d->slot_sterics_delete >= 0 && d->slot_sterics_delete < d->n_lambda &&
                        d->slot_softcore_alpha >= 0 && d->slot_softcore_alpha < d->n_lambda,
                    "soft-core pairs need the four sterics slots");

    // ---- CSR of offsets
    std::vector<int> atom_off_start(A + 1, 0), atom_off_slot;
    std::vector<double> atom_off_par;
    for (int a = 0; a < A; ++a) {
        for (int q = 0; q < d->n_atom_offsets; ++q)
            if (d->atom_off_idx[2 * q] == a) {
                CDW_PACK_REQUIRE(d->atom_off_idx[2 * q + 1] >= 0 && d->atom_off_idx[2 * q + 1] < d->n_lambda, "offset slot out of range");
                atom_off_slot.push_back(d->atom_off_idx[2 * q + 1]);
                for (int c = 0; c < 3; ++c) atom_off_par.push_back(d->atom_off_par[3 * q + c]);
            }
        atom_off_start[a + 1] = (int)atom_off_slot.size();
    }
    std::vector<int> exc_off_start(d->n_exceptions + 1, 0), exc_off_slot;
    std::vector<double> exc_off_par;
    {
        std::vector<std::vector<int>> by(d->n_exceptions);
        for (int q = 0; q < d->n_exc_offsets; ++q) {
            int e = d->exc_off_idx[2 * q];
            CDW_PACK_REQUIRE(e >= 0 && e < d->n_exceptions, "exception offset index out of range");
            CDW_PACK_REQUIRE(d->exc_off_idx[2 * q + 1] >= 0 && d->exc_off_idx[2 * q + 1] < d->n_lambda, "offset slot out of range");
            by[e].push_back(q);
        }
        for (int e = 0; e < d->n_exceptions; ++e) {
            for (int q : by[e]) {
                exc_off_slot.push_back(d->exc_off_idx[2 * q + 1]);
                for (int c = 0; c < 3; ++c) exc_off_par.push_back(d->exc_off_par[3 * q + c]);
            }
            exc_off_start[e + 1] = (int)exc_off_slot.size();
        }
    }

    // ---- constraint clusters (connected components of the constraint graph)
    std::vector<int> parent(A);
    std::iota(parent.begin(), parent.end(), 0);
    auto find = [&](int x) {
        while (parent[x] != x) x = parent[x] = parent[parent[x]];
        return x;
    };
    for (int q = 0; q < d->n_constraints; ++q) {
        int a = find(d->cons_atoms[2 * q]), b = find(d->cons_atoms[2 * q + 1]);
        CDW_PACK_REQUIRE(d->cons_atoms[2 * q] != d->cons_atoms[2 * q + 1] && d->cons_dist[q] > 0, "bad constraint");
        if (a != b) parent[std::max(a, b)] = std::min(a, b);
    }
    std::vector<int> order(d->n_constraints);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int p, int q) { return find(d->cons_atoms[2 * p]) < find(d->cons_atoms[2 * q]); });
    std::vector<int> cluster_start, cons_atoms;
    std::vector<double> cons_d2;
    int last_root = -1;
    for (int q : order) {
        int root = find(d->cons_atoms[2 * q]);
        if (root != last_root) {
            cluster_start.push_back((int)cons_d2.size());
            last_root = root;
        }
        cons_atoms.push_back(d->cons_atoms[2 * q]);
        cons_atoms.push_back(d->cons_atoms[2 * q + 1]);
        cons_d2.push_back(d->cons_dist[q] * d->cons_dist[q]);
    }
    const int n_clusters = (int)cluster_start.size();
    cluster_start.push_back((int)cons_d2.size());
    // coupling coefficients of each small cluster: how a multiplier on constraint r changes the relative velocity /
    // displacement along constraint q (direct k x k projection instead of Gauss-Seidel sweeps)
    std::vector<int> cluster_coef_start(n_clusters + 1, 0);
    std::vector<double> cluster_coef;
    for (int c = 0; c < n_clusters; ++c) {
        const int q0 = cluster_start[c], k = cluster_start[c + 1] - q0;
        cluster_coef_start[c] = (int)cluster_coef.size();
        if (k > CDW_MAX_DIRECT) continue;
        for (int q = 0; q < k; ++q)
            for (int r = 0; r < k; ++r) {
                const int iq = cons_atoms[2 * (q0 + q)], jq = cons_atoms[2 * (q0 + q) + 1];
                const int ir = cons_atoms[2 * (q0 + r)], jr = cons_atoms[2 * (q0 + r) + 1];
                double cf = 0.0;
                if (iq == ir) cf += 1.0 / d->masses[ir];
                if (iq == jr) cf -= 1.0 / d->masses[jr];
                if (jq == ir) cf -= 1.0 / d->masses[ir];
                if (jq == jr) cf += 1.0 / d->masses[jr];
                cluster_coef.push_back(cf);
            }
    }
    cluster_coef_start[n_clusters] = (int)cluster_coef.size();

    // ---- static / dynamic split.  A term is STATIC when no global parameter enters it: its energy and forces at given
    // coordinates are the same for every lambda, so the closing evaluation of one SMC iteration can hand them to the
    // opening evaluation of the next (same coordinates, next lambda).  Every term table is reordered static-first
    // (stable: pairs stay sorted by (i, j) inside each part).
    const int nb = d->n_bonds, na = d->n_angles, nt = d->n_torsions, np = d->n_pairs, ne = d->n_ext;
    auto static_first = [](int n, const std::vector<char>& dyn) {
        std::vector<int> perm(n);
        std::iota(perm.begin(), perm.end(), 0);
        std::stable_sort(perm.begin(), perm.end(), [&](int p, int q) { return dyn[p] < dyn[q]; });
        return perm;
    };
    std::vector<char> dyn_b(nb), dyn_a(na), dyn_t(nt), dyn_p(np), dyn_e(ne);
    for (int q = 0; q < nb; ++q) dyn_b[q] = d->bond_slot[q] >= 0 && (d->bond_par[4 * q + 1] != 0.0 || d->bond_par[4 * q + 3] != 0.0);
    for (int q = 0; q < na; ++q) dyn_a[q] = d->angle_slot[q] >= 0 && (d->angle_par[4 * q + 1] != 0.0 || d->angle_par[4 * q + 3] != 0.0);
    for (int q = 0; q < nt; ++q) dyn_t[q] = d->torsion_slot[q] >= 0 && d->torsion_par[3 * q + 1] != 0.0;
    for (int q = 0; q < ne; ++q) {
        dyn_e[q] = 0;
        for (int c = 0; c < 3; ++c) dyn_e[q] |= d->ext_slot[3 * q + c] >= 0 && d->ext_par[6 * q + 2 * c + 1] != 0.0;
    }
    for (int q = 0; q < np; ++q) {
        const int i = d->pair_atoms[2 * q], j = d->pair_atoms[2 * q + 1], k = d->pair_nb[q];
        bool dy = d->pair_sc_class[q] >= 0;
        if (k == -2) dy |= atom_off_start[i + 1] > atom_off_start[i] || atom_off_start[j + 1] > atom_off_start[j];
        if (k >= 0) dy |= exc_off_start[k + 1] > exc_off_start[k];
        dyn_p[q] = dy;
    }
    const std::vector<int> perm_b = static_first(nb, dyn_b), perm_a = static_first(na, dyn_a), perm_t = static_first(nt, dyn_t),
                           perm_p = static_first(np, dyn_p), perm_e = static_first(ne, dyn_e);
    auto n_static = [](const std::vector<char>& dyn) { return (int)std::count(dyn.begin(), dyn.end(), (char)0); };
    auto permuted_i = [](const int32_t* src, const std::vector<int>& perm, int width) {
        std::vector<int> out;
        for (int q : perm)
            for (int c = 0; c < width; ++c) out.push_back(src[width * q + c]);
        return out;
    };
    auto permuted_d = [](const double* src, const std::vector<int>& perm, int width) {
        std::vector<double> out;
        for (int q : perm)
            for (int c = 0; c < width; ++c) out.push_back(src[width * q + c]);
        return out;
    };

    std::vector<double> masses = vec(d->masses, A), inv_mass(A);
    for (int a = 0; a < A; ++a) inv_mass[a] = 1.0 / masses[a];

    // ---- pack everything into one device blob
    Packer pk;
    size_t o_m = pk.add(masses), o_im = pk.add(inv_mass);
    size_t o_ba = pk.add(permuted_i(d->bond_atoms, perm_b, 2)), o_bs = pk.add(permuted_i(d->bond_slot, perm_b, 1)), o_bp = pk.add(permuted_d(d->bond_par, perm_b, 4));
    size_t o_aa = pk.add(permuted_i(d->angle_atoms, perm_a, 3)), o_as = pk.add(permuted_i(d->angle_slot, perm_a, 1)), o_ap = pk.add(permuted_d(d->angle_par, perm_a, 4));
    size_t o_ta = pk.add(permuted_i(d->torsion_atoms, perm_t, 4)), o_ts = pk.add(permuted_i(d->torsion_slot, perm_t, 1)),
           o_tn = pk.add(permuted_i(d->torsion_n, perm_t, 1)), o_tp = pk.add(permuted_d(d->torsion_par, perm_t, 3));
    std::vector<double> nbpar = d->nb_atom_par ? vec(d->nb_atom_par, 3 * A) : std::vector<double>(3 * A, 0.0);
    size_t o_nb = pk.add(nbpar), o_aos = pk.add(atom_off_start), o_aosl = pk.add(atom_off_slot), o_aop = pk.add(atom_off_par);
    size_t o_ep = pk.add(vec(d->exc_par, 3 * d->n_exceptions)), o_eos = pk.add(exc_off_start), o_eosl = pk.add(exc_off_slot),
           o_eop = pk.add(exc_off_par);
    size_t o_pa = pk.add(permuted_i(d->pair_atoms, perm_p, 2)), o_pn = pk.add(permuted_i(d->pair_nb, perm_p, 1)),
           o_pc = pk.add(permuted_i(d->pair_sc_class, perm_p, 1)),
           o_pp = pk.add(d->pair_sc_par ? permuted_d(d->pair_sc_par, perm_p, 4) : std::vector<double>(4 * (size_t)np, 0.0));
    size_t o_ea = pk.add(permuted_i(d->ext_atom, perm_e, 1)), o_es = pk.add(permuted_i(d->ext_slot, perm_e, 3)), o_epar = pk.add(permuted_d(d->ext_par, perm_e, 6));
    size_t o_cs = pk.add(cluster_start), o_ca = pk.add(cons_atoms), o_cd = pk.add(cons_d2);
    size_t o_ccs = pk.add(cluster_coef_start), o_cc = pk.add(cluster_coef);


    const unsigned char* B = nullptr;
    v.n_atoms = A; v.n_lambda = d->n_lambda;
    v.n_bonds = nb; v.n_angles = na; v.n_torsions = nt; v.n_pairs = np; v.n_ext = ne; v.n_exceptions = d->n_exceptions;
    v.n_cons = d->n_constraints; v.n_clusters = n_clusters;
    v.n_atom_off = (int)atom_off_slot.size(); v.n_exc_off = (int)exc_off_slot.size(); v.n_coef = (int)cluster_coef.size();
    v.slot_core = d->slot_sterics_core; v.slot_insert = d->slot_sterics_insert; v.slot_delete = d->slot_sterics_delete; v.slot_alpha = d->slot_softcore_alpha;
    v.n_bonds_s = n_static(dyn_b); v.n_angles_s = n_static(dyn_a); v.n_torsions_s = n_static(dyn_t); v.n_pairs_s = n_static(dyn_p); v.n_ext_s = n_static(dyn_e);
    v.a_pad = (A + 1) & ~1;
    v.masses = (const double*)(B + o_m); v.inv_mass = (const double*)(B + o_im);
    v.bond_atoms = (const int*)(B + o_ba); v.bond_slot = (const int*)(B + o_bs); v.bond_par = (const double*)(B + o_bp);
    v.angle_atoms = (const int*)(B + o_aa); v.angle_slot = (const int*)(B + o_as); v.angle_par = (const double*)(B + o_ap);
    v.torsion_atoms = (const int*)(B + o_ta); v.torsion_slot = (const int*)(B + o_ts); v.torsion_n = (const int*)(B + o_tn); v.torsion_par = (const double*)(B + o_tp);
    v.nb_atom_par = (const double*)(B + o_nb); v.atom_off_start = (const int*)(B + o_aos); v.atom_off_slot = (const int*)(B + o_aosl); v.atom_off_par = (const double*)(B + o_aop);
    v.exc_par = (const double*)(B + o_ep); v.exc_off_start = (const int*)(B + o_eos); v.exc_off_slot = (const int*)(B + o_eosl); v.exc_off_par = (const double*)(B + o_eop);
    v.pair_atoms = (const int*)(B + o_pa); v.pair_nb = (const int*)(B + o_pn); v.pair_sc_class = (const int*)(B + o_pc); v.pair_sc_par = (const double*)(B + o_pp);
    v.ext_atom = (const int*)(B + o_ea); v.ext_slot = (const int*)(B + o_es); v.ext_par = (const double*)(B + o_epar);
    v.cluster_coef_start = (const int*)(B + o_ccs); v.cluster_coef = (const double*)(B + o_cc);
    v.cluster_start = (const int*)(B + o_cs); v.cons_atoms = (const int*)(B + o_ca); v.cons_d2 = (const double*)(B + o_cd);

    blob.swap(pk.buf);
    return CDW_OK;
}

inline void rebase_system(SysDev& v, const unsigned char* base) {
    const size_t b = (size_t)base;
#define CDW_RB(field) v.field = (decltype(v.field))((size_t)v.field + b)
    CDW_RB(masses); CDW_RB(inv_mass);
    CDW_RB(bond_atoms); CDW_RB(bond_slot); CDW_RB(bond_par);
    CDW_RB(angle_atoms); CDW_RB(angle_slot); CDW_RB(angle_par);
    CDW_RB(torsion_atoms); CDW_RB(torsion_slot); CDW_RB(torsion_n); CDW_RB(torsion_par);
    CDW_RB(nb_atom_par); CDW_RB(atom_off_start); CDW_RB(atom_off_slot); CDW_RB(atom_off_par);
    CDW_RB(exc_par); CDW_RB(exc_off_start); CDW_RB(exc_off_slot); CDW_RB(exc_off_par);
    CDW_RB(pair_atoms); CDW_RB(pair_nb); CDW_RB(pair_sc_class); CDW_RB(pair_sc_par);
    CDW_RB(ext_atom); CDW_RB(ext_slot); CDW_RB(ext_par);
    CDW_RB(cluster_coef_start); CDW_RB(cluster_coef);
    CDW_RB(cluster_start); CDW_RB(cons_atoms); CDW_RB(cons_d2);
#undef CDW_RB
}

}  // namespace cdw
