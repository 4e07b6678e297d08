"""BAOAB Langevin propagation with distance constraints, vectorised over particles.  TEST INFRASTRUCTURE ONLY.

Restates the integrator program `OMMLI` emits for splitting "V R O R V" (coddiwomple/openmm/integrators.py:207-217,
297-350; pretty-printed listing in troubleshooting/openmm_smc.ipynb JSON lines 153-192) as driven by `OMMBIP.apply`
(coddiwomple/openmm/propagators.py:127-219):

    V: v += (dt/2) f/m ; constrain velocities                                  integrators.py:317-336
    R: x += (dt/2) v ; x1 = x ; constrain positions ; v += (x-x1)/(dt/2) ; constrain velocities   :297-315
    O: v = a v + b sqrt(kT/m) xi ; constrain velocities                         :338-350
       a = exp(-gamma dt), b = sqrt(1 - exp(-2 gamma dt))                       :158-159
    bookkeeping: shadow_work += dKE (V) / d(KE+PE) (R) ; proposal_work -= dKE (O)   :315,336,350

`addConstrainPositions/Velocities` are OpenMM built-ins (un-vendored): restated as SHAKE (Gauss-Seidel sweeps in
constraint order, relative distance tolerance as OpenMM's (1 +- tol)^2 window) and its velocity analogue.  Noise is
Philox (oracle/rng.py) because OpenMM's generator is not reproducible outside OpenMM: parity with the reference's
trajectories is statistical only; parity with the CUDA kernel is numerical (same counters -> same normals).

Arithmetic mirrors the kernel contract: particle state (x, v) lives on the fp32 grid wherever it is
stored (float4 rows in HBM, fp32 columns in shared memory); every update is computed in fp64 and rounded back.
"""
import numpy as np

from . import rng

MAX_SHAKE_ITERS = 500


def r32(a):
    """Round to the fp32 grid (the stored precision of particle state) and continue in fp64."""
    return np.asarray(a, dtype=np.float64).astype(np.float32).astype(np.float64)


def _shake(x, v, r0, inv_m, constraints, tol, inv_h):
    """Gauss-Seidel SHAKE in constraint order; every update moves x and, divided by the drift time, v
    ("v += (x - x1)/(dt/2)", integrators.py:310).  State is re-rounded to fp32 after each update."""
    lower, upper = 1.0 - 2.0 * tol + tol * tol, 1.0 + 2.0 * tol + tol * tol
    for _ in range(MAX_SHAKE_ITERS):
        any_update = False
        for q, (i, j, d) in enumerate(constraints):
            s = x[:, i] - x[:, j]
            r2 = (s * s).sum(-1)
            d2 = d * d
            bad = (r2 < lower * d2) | (r2 > upper * d2)
            if not bad.any():
                continue
            any_update = True
            g = (d2 - r2) / (2.0 * (inv_m[i] + inv_m[j]) * (s * r0[q]).sum(-1))
            g = np.where(bad, g, 0.0)[:, None]
            x[:, i] = r32(x[:, i] + g * inv_m[i] * r0[q])
            x[:, j] = r32(x[:, j] - g * inv_m[j] * r0[q])
            v[:, i] = r32(v[:, i] + g * inv_m[i] * inv_h * r0[q])
            v[:, j] = r32(v[:, j] - g * inv_m[j] * inv_h * r0[q])
        if not any_update:
            return
    raise RuntimeError("oracle SHAKE did not converge")


def constraint_clusters(constraints):
    """Connected components of the constraint graph, each as a list of constraint indices (in input order)."""
    parent = {}

    def find(a):
        parent.setdefault(a, a)
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a

    for (i, j, _) in constraints:
        a, b = find(i), find(j)
        if a != b:
            parent[max(a, b)] = min(a, b)
    groups = {}
    for q, (i, _, _) in enumerate(constraints):
        groups.setdefault(find(i), []).append(q)
    return [groups[k] for k in sorted(groups)]


def _rattle_v(x, v, inv_m, constraints, tol):
    """Remove the velocity components along the constraints.  The condition is linear, so each cluster of coupled
    constraints is projected exactly by one k x k solve (what `addConstrainVelocities` converges to):
        A_qr = (s_q . s_r) c_qr,   c_qr = d(iq,ir)/m_ir - d(iq,jr)/m_jr - d(jq,ir)/m_ir + d(jq,jr)/m_jr,   A g = -(s_q . v_rel,q)
    """
    for cl in constraint_clusters(constraints):
        k = len(cl)
        idx = [(constraints[q][0], constraints[q][1]) for q in cl]
        s = np.stack([x[:, i] - x[:, j] for (i, j) in idx], 1)            # (P, k, 3)
        b = -np.stack([((x[:, i] - x[:, j]) * (v[:, i] - v[:, j])).sum(-1) for (i, j) in idx], 1)
        coef = np.zeros((k, k))
        for q, (iq, jq) in enumerate(idx):
            for r, (ir, jr) in enumerate(idx):
                coef[q, r] = (iq == ir) * inv_m[ir] - (iq == jr) * inv_m[jr] - (jq == ir) * inv_m[ir] + (jq == jr) * inv_m[jr]
        A = np.einsum("pqc,prc->pqr", s, s) * coef[None]
        g = np.linalg.solve(A, b[..., None])[..., 0]
        for q, (i, j) in enumerate(idx):
            v[:, i] = r32(v[:, i] + (g[:, q] * inv_m[i])[:, None] * s[:, q])
            v[:, j] = r32(v[:, j] - (g[:, q] * inv_m[j])[:, None] * s[:, q])


def kinetic_energy(v, masses):
    return 0.5 * (masses[None, :, None] * v * v).sum((1, 2))


def maxwell_boltzmann(x, masses, constraints, kT, seed, gids, step, tol=1e-6):
    """Context.setVelocitiesToTemperature as used by propagators.py:137-138: v ~ N(0, kT/m), then velocity constraints."""
    xi = rng.normals(seed, gids, step, len(masses), rng.KIND_MAXWELL_BOLTZMANN)
    v = r32(np.sqrt(kT / masses)[None, :, None] * xi)
    if constraints:
        _rattle_v(np.asarray(x, dtype=np.float64), v, 1.0 / masses, constraints, tol)
    return v


def baoab(energy_forces, x, v, masses, constraints, dt, gamma, kT, n_steps, seed, gids, step0, tol=1e-6,
          randomize_velocities=False, track_works=False):
    """n_steps of "V R O R V".  energy_forces(x) -> (U[P] kJ/mol, F[P,A,3]).

    Returns dict(x, v, u_first, u_last, shadow_work, proposal_work, nan) with energies in kJ/mol; `nan[p]` marks
    particles whose closing energy was not finite — their (x, v) are reverted to the input, which is what
    propagators.py:166-186 does with the default n_restart_attempts = 0.
    Precision contract shared with the CUDA kernel: x and v live on the fp32 grid (they are float4 rows in HBM and
    fp32 columns in shared memory); every update is computed in fp64 and rounded back.
    """
    x0 = r32(x)
    x = x0.copy()
    masses = np.asarray(masses, dtype=np.float64)
    inv_m = 1.0 / masses
    P, A = x.shape[0], x.shape[1]
    v0 = r32(v)  # a reverted move returns the caller's state (propagators.py:184)
    if randomize_velocities:
        v = maxwell_boltzmann(x, masses, constraints, kT, seed, gids, step0, tol)
    v = r32(v)
    a = np.exp(-gamma * dt)
    b = np.sqrt(1.0 - np.exp(-2.0 * gamma * dt))
    sig = np.sqrt(kT / masses)[None, :, None]
    h = 0.5 * dt
    shadow = np.zeros(P)
    proposal = np.zeros(P)
    U, F = energy_forces(x)
    u_first = U.copy()
    for s in range(n_steps):
        # V
        ke0 = kinetic_energy(v, masses) if track_works else None
        v = r32(v + h * F * inv_m[None, :, None])
        if constraints:
            _rattle_v(x, v, inv_m, constraints, tol)
        if track_works:
            shadow += kinetic_energy(v, masses) - ke0
        for half in (0, 1):
            # R
            if track_works:
                e0 = kinetic_energy(v, masses) + (U if half == 0 else U_mid)
            r0 = [(x[:, i].astype(np.float32) - x[:, j].astype(np.float32)).astype(np.float64) for (i, j, _) in constraints]
            x = r32(x + h * v)
            if constraints:
                _shake(x, v, r0, inv_m, constraints, tol, 1.0 / h)
                _rattle_v(x, v, inv_m, constraints, tol)
            if track_works or half == 1:
                if half == 0:
                    U_mid = energy_forces(x)[0]
                else:
                    U, F = energy_forces(x)
            if track_works:
                shadow += kinetic_energy(v, masses) + (U_mid if half == 0 else U) - e0
            if half == 0:
                # O
                ke0 = kinetic_energy(v, masses) if track_works else None
                xi = rng.normals(seed, gids, step0 + s, A, rng.KIND_O_STEP)
                v = r32(a * v + b * sig * xi)
                if constraints:
                    _rattle_v(x, v, inv_m, constraints, tol)
                if track_works:
                    proposal -= kinetic_energy(v, masses) - ke0
        # closing V
        ke0 = kinetic_energy(v, masses) if track_works else None
        v = r32(v + h * F * inv_m[None, :, None])
        if constraints:
            _rattle_v(x, v, inv_m, constraints, tol)
        if track_works:
            shadow += kinetic_energy(v, masses) - ke0
    u_last = U.copy()
    nan = ~np.isfinite(u_last) if n_steps > 0 else np.zeros(P, dtype=bool)
    if nan.any():
        x[nan] = x0[nan]
        v[nan] = v0[nan]
    return dict(x=x, v=v, u_first=u_first, u_last=u_last, shadow_work=shadow, proposal_work=proposal, nan=nan)


STEP_CODES = {"V": 1, "R": 2, "O": 3, "{": 4, "}": 5}
KIND_METROPOLIS = 3


def encode_splitting(splitting):
    """'V R O R V' -> the 4-bit-per-sub-step program word of cdw_langevin_params.splitting (first sub-step in the lowest bits)."""
    word = 0
    for k, tok in enumerate(splitting.upper().split()):
        word |= STEP_CODES[tok] << (4 * k)
    return word


def langevin_program(energy_forces, x, v, masses, constraints, dt, gamma, kT, n_steps, seed, gids, step0, splitting="V R O R V", tol=1e-6,
                     randomize_velocities=False):
    """n_steps of an arbitrary OMMLI splitting (openmm/integrators.py:207-217, 363-425), including ONE Metropolised block:

        V: v += (dt / n_V) f / m        R: x += (dt / n_R) v (+ constraints)        O: v = a v + b sqrt(kT/m) xi over dt / n_O
        shadow_work += dKE (V) / d(KE + PE) (R);  proposal_work -= dKE (O)                               (:315, 336, 350)
        "{": x_old = x, v_old = v                                                                         (:427-430)
        "}": accept iff exp(-shadow_work / kT) >= uniform, else x = x_old, v = -v_old; shadow_work = 0    (:432-444)

    Forces and energies are re-evaluated lazily, when a V or a work increment needs them at moved positions (OpenMM's `f` / `energy`).
    The uniform of the acceptance test is Philox(seed; gid, step, kind 3) like the Gaussian noise.
    Returns dict(x, v, u_first, u_last, shadow_work, proposal_work, trials, accepted, nan)."""
    toks = splitting.upper().split()
    n_v, n_r, n_o = toks.count("V"), toks.count("R"), toks.count("O")
    x0 = r32(x)
    x = x0.copy()
    masses = np.asarray(masses, dtype=np.float64)
    inv_m = 1.0 / masses
    P, A = x.shape[0], x.shape[1]
    v0 = r32(v)
    if randomize_velocities and n_steps > 0:
        v = maxwell_boltzmann(x, masses, constraints, kT, seed, gids, step0, tol)
    v = r32(v)
    hv, hr, ho = (dt / n_v if n_v else 0.0), (dt / n_r if n_r else 0.0), (dt / n_o if n_o else 0.0)
    a, b = np.exp(-gamma * ho), np.sqrt(1.0 - np.exp(-2.0 * gamma * ho))
    sig = np.sqrt(kT / masses)[None, :, None]
    shadow, proposal = np.zeros(P), np.zeros(P)
    trials, accepted = np.zeros(P, dtype=np.int64), np.zeros(P, dtype=np.int64)
    U, F = energy_forces(x)
    fvalid = np.ones(P, dtype=bool)
    u_first = U.copy()
    x_old, v_old = x.copy(), v.copy()

    def refresh():
        nonlocal U, F
        if not fvalid.all():
            Un, Fn = energy_forces(x)
            U = np.where(fvalid, U, Un)
            F = np.where(fvalid[:, None, None], F, Fn)
            fvalid[:] = True

    for s in range(n_steps):
        o_count = 0
        for tok in toks:
            if tok == "{":
                x_old, v_old = x.copy(), v.copy()
                continue
            if tok == "}":
                c0 = (gids & 0xFFFFFFFF).astype(np.uint64)
                c1 = (np.asarray(gids, dtype=np.uint64) >> np.uint64(32))
                r0_, _, _, _ = rng.philox4x32_10(c0, c1, np.uint64((step0 + s) & 0xFFFFFFFF), np.uint64(KIND_METROPOLIS << 24), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
                u = rng.uniform24(r0_).astype(np.float64)
                with np.errstate(over="ignore", invalid="ignore"):
                    accept = np.exp(-shadow / kT) - u >= 0.0
                trials += 1
                accepted += accept
                rej = ~accept
                x[rej] = x_old[rej]
                v[rej] = -v_old[rej]
                fvalid[rej] = False
                shadow[:] = 0.0
                continue
            if tok != "O":
                refresh()
            ke0 = kinetic_energy(v, masses)
            if tok == "V":
                v = r32(v + hv * F * inv_m[None, :, None])
            elif tok == "R":
                r0 = [(x[:, i].astype(np.float32) - x[:, j].astype(np.float32)).astype(np.float64) for (i, j, _) in constraints]
                x = r32(x + hr * v)
                if constraints:
                    _shake(x, v, r0, inv_m, constraints, tol, 1.0 / hr)
            else:
                xi = rng.normals(seed, gids, step0 + s, A, rng.KIND_O_STEP | (o_count << 4))
                o_count += 1
                v = r32(a * v + b * sig * xi)
            if constraints:
                _rattle_v(x, v, inv_m, constraints, tol)
            ke1 = kinetic_energy(v, masses)
            if tok == "V":
                shadow += ke1 - ke0
            elif tok == "O":
                proposal -= ke1 - ke0
            else:
                U_old = U
                U, F = energy_forces(x)
                fvalid[:] = True
                shadow += (ke1 + U) - (ke0 + U_old)
    refresh()
    u_last = U.copy()
    nan = ~np.isfinite(u_last) if n_steps > 0 else np.zeros(P, dtype=bool)
    return dict(x=x, v=v, u_first=u_first, u_last=u_last, shadow_work=shadow, proposal_work=proposal, trials=trials, accepted=accepted, nan=nan)
