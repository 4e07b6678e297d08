/* cdw.h — C ABI of libcdw_b200.so: the B200-native SMC annealing hot path of choderalab/coddiwomple.
 *
 * The reference (pure Python) has no FFI for this path; it reaches OpenMM through openmmtools objects.  Each entry
 * point below replaces the reference interface cited beside it (paths relative to the reference checkout) and is what
 * a ctypes binding in the reference would call (see INTEGRATION.md for the stub).
 *
 * Conventions
 *   - every function returns 0 on success or a negative CDW_E* code; cdw_last_error() gives a thread-local message;
 *   - all array arguments are CALLER-OWNED DEVICE pointers (e.g. torch tensors' data_ptr()) unless marked "host";
 *     nothing is retained past the call except by cdw_system_create (which copies its host tables to the device);
 *   - every call only ENQUEUES work on `stream` (a cudaStream_t passed as void*) and never synchronises, so the
 *     calls can be captured into a CUDA graph; results are ordered on that stream;
 *   - particle coordinates / velocities are rows of float4: pos[p][a] = (x, y, z, unused) in nm, nm/ps;
 *   - energies are kJ/mol unless a `beta`/`kT` argument says they are reduced (units of kT);
 *   - a cdw_system handle is immutable after creation and may be shared by threads; other state is the caller's.
 */
#ifndef CDW_H
#define CDW_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CDW_VERSION 1

#if defined(__GNUC__)
#define CDW_API __attribute__((visibility("default")))
#else
#define CDW_API
#endif

#define CDW_OK 0
#define CDW_EINVAL (-1)   /* bad argument */
#define CDW_ECUDA (-2)    /* CUDA runtime error (message in cdw_last_error) */
#define CDW_ENOMEM (-3)
#define CDW_EUNSUPPORTED (-4)

typedef struct cdw_system cdw_system; /* opaque, device-resident force field of ONE replica */
typedef void* cdw_stream;             /* cudaStream_t */

/* Host-side description of one replica's potential: the OpenMM System of the reference lowered to term tables.
 * Every lambda-dependent parameter is AFFINE in one global parameter:  value = c0 + lambda[slot] * c1  (slot < 0 ->
 * constant c0), which covers every alchemical form in the reference's hybrid systems
 * (coddiwomple/data/perses_data/*.factory.pkl: `_hybrid_system` XML) and the harmonic test system
 * (coddiwomple/tests/utils.py:27-72).  All pointers are HOST pointers, copied by cdw_system_create. */
typedef struct cdw_system_desc {
    int32_t n_atoms;
    int32_t n_lambda;            /* number of global parameters (lambda vector length) */
    const double* masses;        /* [n_atoms] amu */

    /* harmonic bonds  E = K/2 (r-L)^2       (HarmonicBondForce, CustomBondForce "(K/2)*(r-length)^2") */
    int32_t n_bonds;
    const int32_t* bond_atoms;   /* [n_bonds][2] */
    const int32_t* bond_slot;    /* [n_bonds]    */
    const double* bond_par;      /* [n_bonds][4] = K0, K1, L0, L1 */

    /* harmonic angles E = K/2 (theta-T)^2   (HarmonicAngleForce, CustomAngleForce) */
    int32_t n_angles;
    const int32_t* angle_atoms;  /* [n_angles][3], vertex in the middle */
    const int32_t* angle_slot;
    const double* angle_par;     /* [n_angles][4] = K0, K1, T0, T1 */

    /* periodic torsions E = k (1 + cos(n phi - phase))   (PeriodicTorsionForce; each half of CustomTorsionForce) */
    int32_t n_torsions;
    const int32_t* torsion_atoms; /* [n_torsions][4] */
    const int32_t* torsion_slot;
    const int32_t* torsion_n;     /* integer periodicity 0..6 */
    const double* torsion_par;    /* [n_torsions][3] = k0, k1, phase */

    /* NonbondedForce (NoCutoff): per-atom (q, sigma, epsilon) + ParticleOffsets; exceptions + ExceptionOffsets */
    const double* nb_atom_par;     /* [n_atoms][3] */
    int32_t n_atom_offsets;
    const int32_t* atom_off_idx;   /* [n][2] = atom, slot */
    const double* atom_off_par;    /* [n][3] = dq, dsigma, depsilon */
    int32_t n_exceptions;
    const double* exc_par;         /* [n_exceptions][3] = chargeProd, sigma, epsilon */
    int32_t n_exc_offsets;
    const int32_t* exc_off_idx;    /* [n][2] = exception, slot */
    const double* exc_off_par;     /* [n][3] */

    /* unified pair list: every atom pair with a possibly non-zero nonbonded or soft-core contribution */
    int32_t n_pairs;
    const int32_t* pair_atoms;     /* [n_pairs][2] */
    const int32_t* pair_nb;        /* [n_pairs] -1: none, -2: mixing rule from nb_atom_par, >=0: exception index */
    const int32_t* pair_sc_class;  /* [n_pairs] -1: none, 0: core, 1: touches a unique-new atom, 2: unique-old */
    const double* pair_sc_par;     /* [n_pairs][4] = sigmaA, epsilonA, sigmaB, epsilonB (already mixed per pair) */
    /* perses soft-core LJ v2 (CustomNonbondedForce "U_sterics = select(step(r - r_LJ), ...)") global slots */
    int32_t slot_sterics_core, slot_sterics_insert, slot_sterics_delete, slot_softcore_alpha;

    /* external harmonic wells E = K/2 ((x-x0)^2 + y^2 + z^2) + U0   (openmmtools HarmonicOscillator) */
    int32_t n_ext;
    const int32_t* ext_atom;       /* [n_ext] */
    const int32_t* ext_slot;       /* [n_ext][3] = slot of K, x0, U0 */
    const double* ext_par;         /* [n_ext][6] = K0, K1, x00, x01, U00, U01 */

    /* holonomic distance constraints */
    int32_t n_constraints;
    const int32_t* cons_atoms;     /* [n_constraints][2] */
    const double* cons_dist;       /* [n_constraints] nm */
} cdw_system_desc;

/* Dense, position-independent quadratic twist of ONE iteration of the Euler-Maruyama proposal (controlled SMC, troubleshooting/csmc.py:460-623
 * with a full matrix A_t):  psi_t = exp(-(x^T A_t x + b_t^T x + c_t + d_t)).  The uncontrolled kernel is N(f, S), f = x + tau beta F(x),
 * S = diag(2 tau) per coordinate (tau_a = kT dt^2 / (2 m_a), openmm/integrators.py:676-680).  The caller precomputes, with
 * Sigma~ = (S^-1 + 2 A_t)^-1 (must be positive definite) and D = 3 A, all row-major DEVICE arrays:
 *   theta [D][D] = Sigma~ S^-1 (Theta_t, :460-478)    chol [D][D] = lower Cholesky factor of Sigma~    A [D][D] = A_t (symmetric)
 *   b [D] = b_t                                        beta [D] = Sigma~ b_t
 *   logk0 = 0.5 log det theta + 0.5 b^T Sigma~ b - c - d          cd = c + d
 * and provides `scratch`, 2 D doubles per particle of the call.  The move becomes x' ~ N(theta f - beta, Sigma~)
 * (twisted_forward_proposal, :503-528) and the incremental work loses log K_t(psi_t)(x) + phi_t(x') (:531-561, :593-623), with
 * log K_t(psi_t)(x) = logk0 - (theta f).(A f) - b.(theta f)  (the reference's two large quadratic forms, expanded so that nothing cancels). */
typedef struct cdw_dense_twist {
    const double* theta;
    const double* chol;
    const double* A;
    const double* b;
    const double* beta;
    double logk0;
    double cd;
    double* scratch;
} cdw_dense_twist;

/* Langevin "V R O R V" parameters — OMMLI.__init__ (coddiwomple/openmm/integrators.py:75-131) */
typedef struct cdw_langevin_params {
    double dt;              /* ps */
    double gamma;           /* collision rate, 1/ps */
    double kT;              /* kJ/mol */
    double constraint_tol;  /* relative distance tolerance (setConstraintTolerance) */
    int32_t n_steps;        /* integrator steps per call (OMMBIP.apply n_steps, default 1) */
    uint32_t flags;         /* CDW_FLAG_* */
    /* Splitting program (OMMLI's `splitting` string, openmm/integrators.py:207-217, 363-425): 0 = "V R O R V"; otherwise one
     * CDW_STEP_* code per sub-step, 4 bits each, first sub-step in the lowest bits, at most 15 sub-steps.  cdw_smc_propagate /
     * cdw_baoab only; shadow and proposal works are always tracked.  A Metropolised block "{ ... }" accepts the sub-steps it holds
     * iff exp(-shadow_work / kT) >= uniform, else restores x and flips v (openmm/integrators.py:427-444). */
    uint64_t splitting;
    int32_t* trial_counts;  /* device [P][2] (trials, accepted) of the Metropolised block, or NULL */
    /* Quadratic twist of the Euler-Maruyama proposal (controlled SMC, troubleshooting/csmc.py:460-623), CDW_FLAG_ULA only; NULL = none.
     * DEVICE array of 6 A + 2 doubles for THIS iteration: a[3A] (the diagonal of A_t, 1/nm^2), b[3A] (1/nm), c, d, the coefficients of
     *   phi_t(x) = x^T A_t x + b_t^T x + c_t + d_t,  psi_t = exp(-phi_t).
     * The move becomes x' ~ N(Theta (f - s b), s Theta) with f = x + tau beta F(x), s = 2 tau, Theta = (1 + 2 s a)^-1 per coordinate
     * (Theta_t :460-478, f_t :480-500, twisted_forward_proposal :503-528) and the incremental work loses
     *   log K_t(psi_t)(x) + phi_t(x')   (twisted_forward_log_normalizer :531-561, twisted_t_log_weight :593-623).
     * 1 + 2 s a must be positive for every coordinate. */
    const double* twist;
    /* The same with a full matrix A_t (HOST pointer to the descriptor above, or NULL); excludes `twist` and
     * CDW_FLAG_TWIST_REFERENCE_NORMALIZER. */
    const cdw_dense_twist* dense_twist;
} cdw_langevin_params;

#define CDW_STEP_V 1u      /* v += (dt / n_V) f / m */
#define CDW_STEP_R 2u      /* x += (dt / n_R) v, constraints */
#define CDW_STEP_O 3u      /* Ornstein-Uhlenbeck over dt / n_O */
#define CDW_STEP_OPEN 4u   /* "{" */
#define CDW_STEP_CLOSE 5u  /* "}" */

#define CDW_FLAG_RANDOMIZE_VELOCITIES 1u /* Context.setVelocitiesToTemperature before stepping (propagators.py:137-138) */
#define CDW_FLAG_TRACK_WORKS 2u          /* accumulate shadow_work / proposal_work (integrators.py:315,336,350) */
/* Euler-Maruyama (unadjusted Langevin) proposal instead of BAOAB (integrators.py:676-680, 741-745, 829-888):
 *   x' = x + tau beta F(x) + sqrt(2 tau) xi, tau_a = kT dt^2 / (2 m_a), then position constraints; n_steps moves per call.
 * proposal_out = sum over moves of log K(x'|x) - log L(x|x') (dimensionless), and the incremental work becomes
 *   inc = u_t(x_t) + aux + proposal_work   (distribution_factories.py:98-131; troubleshooting/csmc.py:174-191).
 * Velocities are not used (vel_in / vel_out may be NULL); gamma is ignored. */
#define CDW_FLAG_ULA 4u
/* With a twist: evaluate the quadratic form of the forward log normaliser with Theta^-1, as troubleshooting/csmc.py:556 literally does
 * (mahalanobis(f, dt b, inv(theta))).  The integral of N(x'; f, s) exp(-phi(x')) has Theta there (the default); the two agree when A = 0,
 * the only regime the reference's own test exercises (troubleshooting/test_csmc.py:165-226). */
#define CDW_FLAG_TWIST_REFERENCE_NORMALIZER 8u
#define CDW_RESAMPLE_MULTINOMIAL 0
#define CDW_RESAMPLE_SYSTEMATIC 1

/* layout of the `stats` block written by cdw_weight_stats / read by cdw_resample_decide (doubles) */
#define CDW_STAT_WMIN 0      /* min_i W_i                                  */
#define CDW_STAT_SUM_E 1     /* sum_i exp(-(W_i - Wmin))                   */
#define CDW_STAT_SUM_E2 2    /* sum_i exp(-2 (W_i - Wmin))                 */
#define CDW_STAT_NESS 3      /* (sum e)^2 / (P sum e^2)   (utils.py:41-64) */
#define CDW_STAT_MEAN 4      /* -logsumexp(-W) + log P   (resamplers.py:107) */
#define CDW_STAT_RESAMPLE 5  /* 1.0 if this iteration resamples else 0.0   */
#define CDW_STAT_TOTAL 6     /* integer cdf total T as a double (diagnostic) */
#define CDW_STAT_NNAN 7      /* number of particles whose move was reverted (NaN guard) */
#define CDW_STATS_LEN 8

CDW_API int cdw_version(void);
CDW_API const char* cdw_last_error(void);
/* sm_count / compute capability of the current device; fails (CDW_ECUDA) when no CUDA device is usable */
CDW_API int cdw_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* replaces: OpenMMPDFState.__init__ building the ThermodynamicState + Context (openmm/states.py:50-86) */
CDW_API int cdw_system_create(const cdw_system_desc* desc, cdw_system** out);
CDW_API int cdw_system_destroy(cdw_system* sys);
/* Optional.  Precompute the effective term tables of every row of a lambda protocol: `lambda_rows` is a DEVICE array
 * [n_rows][n_lambda] (the parameter_sequence of openmm/coddiwomple.py:283-297 lowered to numbers), kT in kJ/mol.
 * Afterwards any cdw_smc_propagate / cdw_smc_propagate_sharded / cdw_baoab call whose `lambda` points AT A ROW of that
 * array and whose kT matches fetches the row's prebuilt tables with one bulk copy instead of re-deriving them in every
 * thread block (what OpenMMPDFState.set_parameters + apply_to_context cost the reference P times per iteration,
 * openmm/states.py:89-103).  Results are bit-identical either way.  The array must stay unchanged until
 * cdw_system_clear_protocol or the next registration; registering is not thread-safe against concurrent launches on
 * the same handle; a registration with a different array frees the previous images (CUDA graphs captured with them must
 * be re-captured).  Enqueues on `stream`. */
CDW_API int cdw_system_register_protocol(cdw_system* sys, const double* lambda_rows, int32_t n_rows, double kT, cdw_stream stream);
CDW_API int cdw_system_clear_protocol(cdw_system* sys);
/* Lanes of a warp that share one replica in the propagate kernels (1, 2, 4, 8; 0 = the library's choice, which depends on the
 * system alone).  Energies and forces are summed in team order: runs are bitwise comparable only at equal team size, which is why
 * the choice never follows the batch size or the sharding.  A caller that knows the size of the WHOLE ensemble may pin the team:
 * small teams waste fewer lanes and suit ensembles that fill the GPU many times over, large teams shorten the per-replica
 * dependency chain and suit ensembles of about one wave.  cdw_system_smem_bytes: shared memory per CTA of 32 replicas. */
CDW_API int cdw_system_set_team(cdw_system* sys, int32_t lanes);
CDW_API int cdw_system_team(const cdw_system* sys);
CDW_API size_t cdw_system_smem_bytes(const cdw_system* sys, int32_t lanes);
CDW_API int cdw_system_n_atoms(const cdw_system* sys);
CDW_API int cdw_system_n_lambda(const cdw_system* sys);

/* K1. replaces: OpenMMPDFState.reduced_potential (openmm/states.py:116-136), batched over P replicas.
 * u_out[p] = beta * U(pos[p]; lambda).  lambda: device pointer to n_lambda doubles. */
CDW_API int cdw_reduced_potential(const cdw_system* sys, const float* pos, int64_t n_particles, const double* lambda,
                          double beta, double* u_out, cdw_stream stream);

/* energy (kJ/mol) and forces (float4 rows, kJ/mol/nm) — what `f` / `energy` are inside the CustomIntegrator
 * (openmm/integrators.py:301,327); exposed for the Euler-Maruyama proposal and for tests. */
CDW_API int cdw_energy_forces(const cdw_system* sys, const float* pos, int64_t n_particles, const double* lambda,
                      double* energy_out, float* force_out, cdw_stream stream);

/* K4 (+K3b on load, +K2 first half).  One SMC iteration's per-particle work, fused:
 *     slot p loads (x, v, aux, label) of its ancestor anc[p] (NULL = identity)           resamplers.py:123-129,189-200
 *     E1,F1 at (x, lambda);  inc[p] = E1/kT + aux ;  W[p] = cum[p] + inc[p]             distribution_factories.py:98-131
 *     n_steps of BAOAB with SHAKE/RATTLE and Philox noise                               integrators.py:297-350
 *     E2 at the new x;  aux_out[p] = -E2/kT                                             openmm/coddiwomple.py:326-327
 *     NaN guard: a non-finite E2 reverts (x, v) and sets nan_flags[p]                   openmm/propagators.py:163-190
 * pos_in/vel_in and pos_out/vel_out must be different buffers when anc != NULL.
 * Optional outputs (NULL to skip): inc_out, w_out, aux_out, label_out, u_first_out, u_last_out (reduced, /kT),
 * shadow_out, proposal_out (reduced), nan_flags, wmin_key (one uint64, atomically min-ed with the order-preserving
 * key of W; initialise to ~0ull).  Optional inputs: anc, aux_in, cum_in, label_in.
 * Noise counters: particle gid = gid0 + p, step = step0 + s. */
CDW_API int cdw_smc_propagate(const cdw_system* sys, const float* pos_in, const float* vel_in, float* pos_out, float* vel_out,
                      int64_t n_particles, const int32_t* anc, const double* aux_in, const double* cum_in,
                      const int32_t* label_in, const double* lambda, const cdw_langevin_params* params, uint64_t seed,
                      uint64_t step0, int64_t gid0, double* inc_out, double* w_out, double* aux_out, int32_t* label_out,
                      double* u_first_out, double* u_last_out, double* shadow_out, double* proposal_out,
                      int32_t* nan_flags, uint64_t* wmin_key, cdw_stream stream);

/* Static-term cache.  Terms into which no global parameter enters (in a perses hybrid: every bond, angle, torsion and pair
 * of the unchanged environment) have the same energy and forces at given coordinates for every lambda, and the closing
 * evaluation of SMC iteration t (lambda_t, x_t) sees the same coordinates as the opening evaluation of iteration t + 1
 * (lambda_{t+1}, x_t).  With a cache the kernel evaluates every force as a static pass plus a dynamic pass, writes the
 * static part at the OUTPUT rows to (force_out, energy_out), and, when (force_in, energy_in) hold the static part at the
 * INPUT rows, skips the static pass of the opening evaluation.  Rows are float4 [P][A] (kJ/mol/nm) and doubles [P]
 * (kJ/mol), indexed like pos_in / pos_out (force_in through `anc`), and must be distinct buffers.
 * Seed a cache with an n_steps = 0 call (force_in = NULL); calls with n_steps > 0 need force_in.  Results equal the
 * uncached call up to summation order.  For the sharded form the inputs are per-rank base pointers. */
typedef struct cdw_static_cache {
    const float* force_in;
    const double* energy_in;
    float* force_out;
    double* energy_out;
    const float* const* peer_force_in;   /* cdw_smc_propagate_sharded_cached only */
    const double* const* peer_energy_in;
} cdw_static_cache;

/* cdw_smc_propagate with a static-term cache (cache == NULL: identical to cdw_smc_propagate). */
CDW_API int cdw_smc_propagate_cached(const cdw_system* sys, const float* pos_in, const float* vel_in, float* pos_out, float* vel_out,
                             int64_t n_particles, const int32_t* anc, const double* aux_in, const double* cum_in,
                             const int32_t* label_in, const double* lambda, const cdw_langevin_params* params, uint64_t seed,
                             uint64_t step0, int64_t gid0, double* inc_out, double* w_out, double* aux_out, int32_t* label_out,
                             double* u_first_out, double* u_last_out, double* shadow_out, double* proposal_out,
                             int32_t* nan_flags, uint64_t* wmin_key, const cdw_static_cache* cache, cdw_stream stream);

/* The look-ahead iteration: the hot loop of the reference (openmm/coddiwomple.py:312-331) for an MCMC move (OMMLI "V R O R V",
 * openmm/integrators.py:297-350, driven by OMMBIP.apply, openmm/propagators.py:89-219), cut at a different place.
 * The incremental work of iteration t + 1, u_{t+1}(x_t) + aux with aux = -u_t(x_t) (distribution_factories.py:124-129;
 * openmm/coddiwomple.py:315, 326-327), only needs the coordinates iteration t ENDS with.  This call therefore does
 *     [gather rows by anc]  (V R O R E_t V)^n_steps  E_{t+1}
 * on forces that arrive with the particle: force_in holds F(x_{t-1}; lambda_t) of the input rows, force_out receives
 * F(x_t; lambda_{t+1}), inc_next_out[p] = u_{t+1}(x_t) - u_t(x_t), aux_out[p] = -u_t(x_t).  The resampling of iteration t + 1
 * (cdw_smc_weigh_resample) then has all its inputs before launch t + 1 starts and runs BESIDE it on another stream: nothing
 * but propagate launches is left on the critical path, for any batch size and any number of GPUs.
 *   n_steps = 0 equips the initial sample (ProposalFactory.equip_initial_sample, distribution_factories.py:180-220):
 *     aux_out = -u_0(x_0), force_out = F(x_0; lambda_1), inc_next_out = u_1(x_0) - u_0(x_0); force_in is not read.
 *   lambda_next = NULL / inc_next_out = NULL on the last iteration (force_out is then scratch).
 * Energies and forces of the terms that no global parameter enters are evaluated once per call and shared by E_t and E_{t+1}.
 * A move that ends in a non-finite energy is reverted (openmm/propagators.py:166-186): outputs are those of the input state.
 * flags: CDW_FLAG_RANDOMIZE_VELOCITIES only.  Rows are float4 [P][A]; in/out buffers must be distinct when anc != NULL. */
CDW_API int cdw_smc_lookahead(const cdw_system* sys, const float* pos_in, const float* vel_in, const float* force_in, float* pos_out,
                              float* vel_out, float* force_out, int64_t n_particles, const int32_t* anc, const int32_t* label_in,
                              const double* lambda, const double* lambda_next, const cdw_langevin_params* params, uint64_t seed,
                              uint64_t step0, int64_t gid0, double* inc_next_out, double* aux_out, int32_t* label_out,
                              int32_t* nan_flags, cdw_stream stream);

/* Sharded form for one rank of a multi-GPU run (see cdw_smc_propagate_sharded): anc holds GLOBAL ancestors, the rows, forces
 * and labels of an ancestor are read from the owning rank through peer-mapped pointers (NVLink). */
CDW_API int cdw_smc_lookahead_sharded(const cdw_system* sys, const float* const* peer_pos, const float* const* peer_vel,
                                      const float* const* peer_force, const int32_t* const* peer_label, int32_t n_ranks,
                                      int64_t n_per_rank, float* pos_out, float* vel_out, float* force_out, int64_t n_particles,
                                      const int32_t* anc, const double* lambda, const double* lambda_next,
                                      const cdw_langevin_params* params, uint64_t seed, uint64_t step0, int64_t gid0,
                                      double* inc_next_out, double* aux_out, int32_t* label_out, int32_t* nan_flags, cdw_stream stream);

/* Sharded form of cdw_smc_propagate for one rank of a multi-GPU run (particles are partitioned in contiguous blocks
 * of n_per_rank by global id; this rank propagates its n local slots, whose global ids start at gid0).
 * `anc[p]` is the GLOBAL ancestor of local slot p; its row, aux and label are read straight from the owning rank's
 * buffers through peer-mapped pointers (peer_*[rank], CUDA IPC over NVLink): the resampling gather, the particle
 * migration and the propagation are ONE kernel.  cum_in points at this rank's slice of the cumulative works. */
CDW_API int cdw_smc_propagate_sharded(const cdw_system* sys, const float* const* peer_pos, const float* const* peer_vel,
                                      const double* const* peer_aux, const int32_t* const* peer_label, int32_t n_ranks,
                                      int64_t n_per_rank, float* pos_out, float* vel_out, int64_t n_particles, const int32_t* anc,
                                      const double* cum_in, const double* lambda, const cdw_langevin_params* params, uint64_t seed,
                                      uint64_t step0, int64_t gid0, double* inc_out, double* w_out, double* aux_out,
                                      int32_t* label_out, int32_t* nan_flags, uint64_t* wmin_key, cdw_stream stream);

CDW_API int cdw_smc_propagate_sharded_cached(const cdw_system* sys, const float* const* peer_pos, const float* const* peer_vel,
                                             const double* const* peer_aux, const int32_t* const* peer_label, int32_t n_ranks,
                                             int64_t n_per_rank, float* pos_out, float* vel_out, int64_t n_particles, const int32_t* anc,
                                             const double* cum_in, const double* lambda, const cdw_langevin_params* params, uint64_t seed,
                                             uint64_t step0, int64_t gid0, double* inc_out, double* w_out, double* aux_out,
                                             int32_t* label_out, int32_t* nan_flags, uint64_t* wmin_key, const cdw_static_cache* cache,
                                             cdw_stream stream);

/* replaces: OMMBIP.apply (openmm/propagators.py:89-219) for a batch, in place.  Thin wrapper over the kernel above. */
CDW_API int cdw_baoab(const cdw_system* sys, float* pos, float* vel, int64_t n_particles, const double* lambda,
              const cdw_langevin_params* params, uint64_t seed, uint64_t step0, int64_t gid0, double* u_first_out,
              double* u_last_out, double* shadow_out, double* proposal_out, int32_t* nan_flags, cdw_stream stream);

/* K2. replaces: TargetFactory.compute_incremental_work's sum (distribution_factories.py:124-129) + nESS
 * (utils.py:41-64) + the normaliser of Resampler.resample (resamplers.py:102-107).
 * cdw_weight_update: inc[p] = u_new[p] + aux[p] (+ proposal[p] if not NULL); w_out[p] = cum[p] + inc[p]; min W -> wmin_key.
 * cdw_weight_stats : second pass over W: ONE deterministic exponential per particle, kept in the workspace as the integer
 *                    rint(e 2^62); sum e and sum e^2 are exact 128-bit integer sums (any grid, any order, any number of GPUs
 *                    give the same bits); the last block finalises stats[] including the resampling decision
 *                    (threshold < 0: always resample, as observable=None does).  w and workspace 16-byte aligned. */
CDW_API int cdw_weight_update(const double* u_new, const double* aux, const double* proposal, const double* cum, int64_t n,
                      double* inc_out, double* w_out, uint64_t* wmin_key, cdw_stream stream);
CDW_API size_t cdw_weight_workspace_bytes(int64_t n);
CDW_API int cdw_weight_stats(const double* w, int64_t n, double floor_threshold, const uint64_t* wmin_key, double* stats,
                     void* workspace, size_t workspace_bytes, cdw_stream stream);

/* K3a. replaces: MultinomialResampler._resample (resamplers.py:254-272); systematic is new.
 * Must follow cdw_weight_stats on the same workspace (it reads the quantised weights from there).  The cdf is the uint64
 * prefix sum of those integers (cdf_out[n], caller-provided, 16-byte aligned).  Systematic ancestors (targets
 * floor(((i + u0) (1/n)) total), sorted) are emitted by the scan itself; multinomial ancestors are looked up through a guide
 * table the scan emits the same way.  uniforms: n doubles in [0,1) (multinomial) or 1 double (systematic).
 * If stats[CDW_STAT_RESAMPLE] == 0 the kernels write the identity ancestors and cum_out = W (the null update,
 * resamplers.py:136-156); otherwise cum_out[p] = cum_in[p] + (mean - cum_in[p]) (resamplers.py:196-197). */
CDW_API int cdw_resample_indices(int kind, const double* w, const double* uniforms, int64_t n, const double* stats,
                         const double* cum_in, double* cum_out, int32_t* anc_out, uint64_t* cdf_out, void* workspace,
                         size_t workspace_bytes, cdw_stream stream);

/* K2 (second half) + K3a in one call: everything Resampler.resample does between two propagate calls of the SMC loop
 * (openmm/coddiwomple.py:318).  Uniforms come from Philox (seed, stream_id): n draws for multinomial, one for systematic
 * (written to `uniforms` for inspection).  *wmin_key (the running min left by the weight pass) is consumed and re-armed
 * to ~0.  Three launches (systematic) or four (multinomial), no host decision in between.  `workspace` must be zero when
 * first used; the kernels leave it at rest. */
CDW_API int cdw_smc_resample(int kind, const double* w, int64_t n, double floor_threshold, uint64_t* wmin_key, uint64_t seed,
                             uint64_t stream_id, double* stats, const double* cum_in, double* cum_out, int32_t* anc_out,
                             uint64_t* cdf_out, double* uniforms, void* workspace, size_t workspace_bytes, cdw_stream stream);

/* The resampling of the look-ahead loop: W[i] = cum[i] + inc_src[anc_prev[i]] (the incremental work cdw_smc_lookahead left at
 * the slot the state lived in before the previous resampling moved it to slot i; anc_prev NULL: identity), then
 * cdw_smc_resample on W with `cum` updated in place.  inc_out (may be NULL) receives the per-slot incremental works. */
CDW_API int cdw_smc_weigh_resample(int kind, const double* inc_src, const int32_t* anc_prev, int64_t n, double floor_threshold,
                                   uint64_t* wmin_key, uint64_t seed, uint64_t stream_id, double* stats, double* cum, double* inc_out,
                                   double* w_out, int32_t* anc_out, uint64_t* cdf_out, double* uniforms, void* workspace,
                                   size_t workspace_bytes, cdw_stream stream);

/* K3b. replaces: the per-particle deepcopy loop of Resampler.resample (resamplers.py:123-129,193). */
CDW_API int cdw_gather_particles(const float* pos_in, const float* vel_in, const int32_t* anc, float* pos_out, float* vel_out,
                         const double* aux_in, double* aux_out, const int32_t* label_in, int32_t* label_out,
                         int64_t n_particles, int32_t n_atoms, cdw_stream stream);

/* multi-GPU variant of K3b: anc holds GLOBAL ancestor ids; rank r owns [r*n_per_rank, (r+1)*n_per_rank);
 * peer_pos/peer_vel/peer_aux/peer_label: device arrays of n_ranks base pointers (peer-mapped via CUDA IPC). */
CDW_API int cdw_gather_particles_peer(const float* const* peer_pos, const float* const* peer_vel, const double* const* peer_aux,
                              const int32_t* const* peer_label, int32_t n_ranks, int64_t n_per_rank, const int32_t* anc,
                              float* pos_out, float* vel_out, double* aux_out, int32_t* label_out, int64_t n_local,
                              int32_t n_atoms, cdw_stream stream);

/* History spill. replaces: the per-iteration appends to Particle._incremental_works / Particle._ancestry
 * (particles.py:38-47; resamplers.py:115,129,196-200): cum_row[i] = cum[i], label_row[i] = label[anc[i]] (anc NULL: identity).
 * `label` holds the root labels of the states BEFORE this iteration's ancestors are applied. */
CDW_API int cdw_record_lineage(const double* cum, const int32_t* label, const int32_t* anc, double* cum_row, int32_t* label_row,
                               int64_t n, cdw_stream stream);

/* Cross-GPU ordering point of the sharded loop (one tiny launch): every rank tells every peer that its launches up to here
 * have completed (a store into the peer's flag array through a peer-mapped pointer) and waits until all peers have said so.
 * peer_flags: device array of `world` base pointers to the ranks' flag arrays (uint64[world] each, zero-initialised);
 * *epoch_counter (device, zero-initialised, private to the rank) is incremented by the launch itself, so the call can be
 * captured in a CUDA graph and replayed; every rank must call it the same number of times.  *status is set to 1 if a peer did
 * not show up within ~2 s (the wait is bounded).  skip_if_zero_a / _b (device, may be NULL): when both are given and 0.0 the wait
 * is skipped (the signal is not) — pass the CDW_STAT_RESAMPLE entries of the two previous iterations, which are identical on every
 * rank: if neither resampled, every ancestor is the identity and no rank reads or overwrites rows a peer still needs. */
CDW_API int cdw_peer_barrier(uint64_t* const* peer_flags, uint64_t* flags_local, int32_t rank, int32_t world, uint64_t* epoch_counter,
                             int32_t* status, const double* skip_if_zero_a, const double* skip_if_zero_b, cdw_stream stream);

/* utilities */
CDW_API int cdw_fill_u64(uint64_t* ptr, uint64_t value, int64_t n, cdw_stream stream);
/* Philox uniforms with 53 random bits (same stream as oracle/rng.py:uniforms53) */
CDW_API int cdw_philox_uniforms(uint64_t seed, uint64_t stream_id, int64_t n, double* out, cdw_stream stream);
/* CUDA IPC plumbing for the peer gather (handles are 64 opaque bytes) */
CDW_API int cdw_ipc_get_handle(const void* dev_ptr, void* handle64_host);
CDW_API int cdw_ipc_open_handle(const void* handle64_host, void** dev_ptr_out);
CDW_API int cdw_ipc_close_handle(void* dev_ptr);
CDW_API int cdw_device_malloc(size_t bytes, void** out);
CDW_API int cdw_device_free(void* ptr);
/* FP64 / FP32 FMA-pipe micro-benchmark used for the roofline denominator: returns achieved TFLOP/s */
CDW_API int cdw_fma_peak(int fp64, int iters, double* tflops_out, double* ms_out);

#ifdef __cplusplus
}
#endif
#endif /* CDW_H */
