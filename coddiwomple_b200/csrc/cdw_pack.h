// cdw_pack.h — host-side construction of one replica's device tables (pure C++, no CUDA): validation, CSR of the
// lambda offsets, constraint clusters, force-scratch slot map; everything packed into ONE blob so that the same
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

    // ---- force scratch slots: slot(type, pos, idx) = base(type) + pos * n_type + idx; CSR atom -> slots
    const int nb = d->n_bonds, na = d->n_angles, nt = d->n_torsions, np = d->n_pairs, ne = d->n_ext;
    const int base_b = 0, base_a = 2 * nb, base_t = base_a + 3 * na, base_p = base_t + 4 * nt, base_e = base_p + 2 * np;
    const int n_slots = base_e + ne;
    std::vector<std::vector<int>> per_atom(A);
    for (int q = 0; q < nb; ++q)
        for (int c = 0; c < 2; ++c) per_atom[d->bond_atoms[2 * q + c]].push_back(base_b + c * nb + q);
    for (int q = 0; q < na; ++q)
        for (int c = 0; c < 3; ++c) per_atom[d->angle_atoms[3 * q + c]].push_back(base_a + c * na + q);
    for (int q = 0; q < nt; ++q)
        for (int c = 0; c < 4; ++c) per_atom[d->torsion_atoms[4 * q + c]].push_back(base_t + c * nt + q);
    for (int q = 0; q < np; ++q)
        for (int c = 0; c < 2; ++c) per_atom[d->pair_atoms[2 * q + c]].push_back(base_p + c * np + q);
    for (int q = 0; q < ne; ++q) per_atom[d->ext_atom[q]].push_back(base_e + q);
    std::vector<int> atom_slot_start(A + 1, 0), atom_slots;
    for (int a = 0; a < A; ++a) {
        atom_slots.insert(atom_slots.end(), per_atom[a].begin(), per_atom[a].end());
        atom_slot_start[a + 1] = (int)atom_slots.size();
    }

    std::vector<double> masses = vec(d->masses, A), inv_mass(A);
    for (int a = 0; a < A; ++a) inv_mass[a] = 1.0 / masses[a];

    // ---- pack everything into one device blob
    Packer pk;
    size_t o_m = pk.add(masses), o_im = pk.add(inv_mass);
    size_t o_ba = pk.add(vec(d->bond_atoms, 2 * nb)), o_bs = pk.add(vec(d->bond_slot, nb)), o_bp = pk.add(vec(d->bond_par, 4 * nb));
    size_t o_aa = pk.add(vec(d->angle_atoms, 3 * na)), o_as = pk.add(vec(d->angle_slot, na)), o_ap = pk.add(vec(d->angle_par, 4 * na));
    size_t o_ta = pk.add(vec(d->torsion_atoms, 4 * nt)), o_ts = pk.add(vec(d->torsion_slot, nt)), o_tn = pk.add(vec(d->torsion_n, nt)),
           o_tp = pk.add(vec(d->torsion_par, 3 * nt));
    std::vector<double> nbpar = d->nb_atom_par ? vec(d->nb_atom_par, 3 * A) : std::vector<double>(3 * A, 0.0);
    size_t o_nb = pk.add(nbpar), o_aos = pk.add(atom_off_start), o_aosl = pk.add(atom_off_slot), o_aop = pk.add(atom_off_par);
    size_t o_ep = pk.add(vec(d->exc_par, 3 * d->n_exceptions)), o_eos = pk.add(exc_off_start), o_eosl = pk.add(exc_off_slot),
           o_eop = pk.add(exc_off_par);
    size_t o_pa = pk.add(vec(d->pair_atoms, 2 * np)), o_pn = pk.add(vec(d->pair_nb, np)), o_pc = pk.add(vec(d->pair_sc_class, np)),
           o_pp = pk.add(vec(d->pair_sc_par, 4 * np));
    size_t o_ea = pk.add(vec(d->ext_atom, ne)), o_es = pk.add(vec(d->ext_slot, 3 * ne)), o_epar = pk.add(vec(d->ext_par, 6 * ne));
    size_t o_cs = pk.add(cluster_start), o_ca = pk.add(cons_atoms), o_cd = pk.add(cons_d2);
    size_t o_ccs = pk.add(cluster_coef_start), o_cc = pk.add(cluster_coef);
    size_t o_ss = pk.add(atom_slot_start), o_sl = pk.add(atom_slots);


    const unsigned char* B = nullptr;
    v.n_atoms = A; v.n_lambda = d->n_lambda;
    v.n_bonds = nb; v.n_angles = na; v.n_torsions = nt; v.n_pairs = np; v.n_ext = ne; v.n_exceptions = d->n_exceptions;
    v.n_cons = d->n_constraints; v.n_clusters = n_clusters;
    v.n_atom_off = (int)atom_off_slot.size(); v.n_exc_off = (int)exc_off_slot.size(); v.n_coef = (int)cluster_coef.size();
    v.slot_core = d->slot_sterics_core; v.slot_insert = d->slot_sterics_insert; v.slot_delete = d->slot_sterics_delete; v.slot_alpha = d->slot_softcore_alpha;
    v.n_slots = n_slots;
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
    v.atom_slot_start = (const int*)(B + o_ss); v.atom_slots = (const int*)(B + o_sl);

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
    CDW_RB(atom_slot_start); CDW_RB(atom_slots);
#undef CDW_RB
}

}  // namespace cdw
