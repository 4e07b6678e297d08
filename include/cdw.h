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

/* Langevin "V R O R V" parameters — OMMLI.__init__ (coddiwomple/openmm/integrators.py:75-131) */
typedef struct cdw_langevin_params {
    double dt;              /* ps */
    double gamma;           /* collision rate, 1/ps */
    double kT;              /* kJ/mol */
    double constraint_tol;  /* relative distance tolerance (setConstraintTolerance) */
    int32_t n_steps;        /* integrator steps per call (OMMBIP.apply n_steps, default 1) */
    uint32_t flags;         /* CDW_FLAG_* */
} cdw_langevin_params;

#define CDW_FLAG_RANDOMIZE_VELOCITIES 1u /* Context.setVelocitiesToTemperature before stepping (propagators.py:137-138) */
#define CDW_FLAG_TRACK_WORKS 2u          /* accumulate shadow_work / proposal_work (integrators.py:315,336,350) */
/* Euler-Maruyama (unadjusted Langevin) proposal instead of BAOAB (integrators.py:676-680, 741-745, 829-888):
 *   x' = x + tau beta F(x) + sqrt(2 tau) xi, tau_a = kT dt^2 / (2 m_a), then position constraints; n_steps moves per call.
 * proposal_out = sum over moves of log K(x'|x) - log L(x|x') (dimensionless), and the incremental work becomes
 *   inc = u_t(x_t) + aux + proposal_work   (distribution_factories.py:98-131; troubleshooting/csmc.py:174-191).
 * Velocities are not used (vel_in / vel_out may be NULL); gamma is ignored. */
#define CDW_FLAG_ULA 4u
/* Split launches of one MCMC propagate call (used by the sharded loop to overlap the weight exchange and the resampling
 * with the integration): PHASE_OPEN gathers the rows, runs the opening evaluation only and writes inc_out / w_out (and the
 * running min) plus the hand-over buffers cdw_static_cache.force_open / u_open; PHASE_REST gathers the rows again, takes
 * the opening forces and energy from those buffers and does everything else (integration, closing evaluation, outputs).
 * OPEN followed by REST equals one unsplit call bit for bit.  Both need a cdw_static_cache with force_open / u_open. */
#define CDW_FLAG_PHASE_OPEN 8u
#define CDW_FLAG_PHASE_REST 16u

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
    float* force_open;                   /* CDW_FLAG_PHASE_OPEN (out) / CDW_FLAG_PHASE_REST (in): [P][A][4] and [P], local slots */
    double* u_open;
} cdw_static_cache;

/* cdw_smc_propagate with a static-term cache (cache == NULL: identical to cdw_smc_propagate). */
CDW_API int cdw_smc_propagate_cached(const cdw_system* sys, const float* pos_in, const float* vel_in, float* pos_out, float* vel_out,
                             int64_t n_particles, const int32_t* anc, const double* aux_in, const double* cum_in,
                             const int32_t* label_in, const double* lambda, const cdw_langevin_params* params, uint64_t seed,
                             uint64_t step0, int64_t gid0, double* inc_out, double* w_out, double* aux_out, int32_t* label_out,
                             double* u_first_out, double* u_last_out, double* shadow_out, double* proposal_out,
                             int32_t* nan_flags, uint64_t* wmin_key, const cdw_static_cache* cache, cdw_stream stream);

/* One whole SMC iteration of the reference loop (openmm/coddiwomple.py:314-328): cdw_smc_propagate_cached followed by
 * cdw_smc_resample on its weights, with `cum` updated in place and the new ancestors in anc_out (anc_in, the ancestors
 * still to be applied to the input rows, may be the same buffer or NULL).  For MCMC moves the incremental work is final
 * right after the opening evaluation; when the whole batch fits the GPU at once the library therefore adds ONE thread
 * block to the propagate launch that waits for the weights and resamples while the other blocks are still integrating:
 * the iteration is a single kernel launch and the resampling costs no time.  Otherwise (Euler-Maruyama proposal, batches
 * of several waves) the two kernels are launched back to back.  Results are identical either way.  `workspace` as for
 * cdw_weight_stats (zero-initialised once); w_out, wmin_key (armed with ~0 before the first call), stats, cdf, uniforms
 * are required. */
CDW_API int cdw_smc_iteration(const cdw_system* sys, const float* pos_in, const float* vel_in, float* pos_out, float* vel_out,
                              int64_t n_particles, const int32_t* anc_in, const double* aux_in, double* cum, const int32_t* label_in,
                              const double* lambda, const cdw_langevin_params* params, uint64_t seed, uint64_t step0, int64_t gid0,
                              double* inc_out, double* w_out, double* aux_out, int32_t* label_out, int32_t* nan_flags, uint64_t* wmin_key,
                              const cdw_static_cache* cache, int kind, double floor_threshold, uint64_t resample_seed, uint64_t stream_id,
                              double* stats, int32_t* anc_out, uint64_t* cdf, double* uniforms, void* workspace, size_t workspace_bytes,
                              cdw_stream stream);

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
 * cdw_weight_stats : second pass over W: sum e, sum e^2, integer total; the last block finalises stats[] including
 *                    the resampling decision (threshold < 0: always resample, as observable=None does). */
CDW_API int cdw_weight_update(const double* u_new, const double* aux, const double* proposal, const double* cum, int64_t n,
                      double* inc_out, double* w_out, uint64_t* wmin_key, cdw_stream stream);
CDW_API size_t cdw_weight_workspace_bytes(int64_t n);
CDW_API int cdw_weight_stats(const double* w, int64_t n, double floor_threshold, const uint64_t* wmin_key, double* stats,
                     void* workspace, size_t workspace_bytes, cdw_stream stream);

/* K3a. replaces: MultinomialResampler._resample (resamplers.py:254-272); systematic is new.
 * Uses the exactly-associative integer cdf; cdf_out[n] uint64 is caller-provided scratch/output.
 * uniforms: n doubles in [0,1) (multinomial) or 1 double (systematic).  If stats[CDW_STAT_RESAMPLE] == 0 the kernels
 * write the identity ancestors and cum_out = W (the null update, resamplers.py:136-156); otherwise
 * cum_out[p] = cum_in[p] + (mean - cum_in[p]) (resamplers.py:196-197). */
CDW_API int cdw_resample_indices(int kind, const double* w, const double* uniforms, int64_t n, const double* stats,
                         const double* cum_in, double* cum_out, int32_t* anc_out, uint64_t* cdf_out, void* workspace,
                         size_t workspace_bytes, cdw_stream stream);

/* K2 (second half) + K3a in one call: everything between two propagate launches of the SMC loop
 * (openmm/coddiwomple.py:318).  Uniforms come from Philox (seed, stream_id): n draws for multinomial, one for systematic
 * (written to `uniforms` for inspection).  *wmin_key (the running min left by cdw_smc_propagate) is consumed and re-armed
 * to ~0.  For n <= 24576 this is ONE launch of 8 cooperating thread blocks (the routine cdw_smc_iteration runs inside the
 * propagate launch); larger batches run statistics + three scan kernels + a search kernel that draws the uniforms itself
 * (5 launches).  Same ancestors either way.  `workspace` must be zero when first used; the kernels leave it at rest. */
CDW_API int cdw_smc_resample(int kind, const double* w, int64_t n, double floor_threshold, uint64_t* wmin_key, uint64_t seed,
                             uint64_t stream_id, double* stats, const double* cum_in, double* cum_out, int32_t* anc_out,
                             uint64_t* cdf_out, double* uniforms, void* workspace, size_t workspace_bytes, cdw_stream stream);

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
