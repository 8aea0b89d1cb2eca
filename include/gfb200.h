/* gfb200.h -- C ABI of libgfb200.so, the B200-native backend for Gaugefields.jl's quenched
 * SU(3) Wilson update path (staples -> TA force -> exp(eps P) U -> plaquette/action reduction).
 *
 * This is the drop-in boundary: Julia reaches these entry points through `ccall` from a
 * `B200Backend` field type (see INTEGRATION.md and gaugefields.jl_b200/julia/B200Backend.jl);
 * in this repository the same symbols are bound by ctypes in gaugefields.jl_b200/gfb200/.
 * Each entry point cites the reference interface it stands behind (file:line relative to the
 * Gaugefields.jl v1.0.5 tree).
 *
 * Conventions
 *   - every function returns 0 on success and a non-zero gfb_status otherwise; no exception or
 *     abort crosses the boundary; gfb_last_error() returns the message of the last failure.
 *   - the library owns all device memory; callers hold opaque handles and free them explicitly.
 *   - host buffers are caller-owned, touched only inside upload/download calls, in the
 *     reference's gathered-array layout (src/API.jl:516-529, 625-629):
 *        links   : ComplexF64[3,3,NX,NY,NZ,NT]  column-major, interleaved (re,im)
 *        momenta : Float64[8,1,NX,NY,NZ,NT]      (TA_gaugefields_4D_MPILattice.jl:34-35)
 *     `mu` is 0-based here (Julia's direction mu+1).
 *   - scalar-returning calls synchronise; all others are asynchronous and stream-ordered.
 *   - one thread at a time per context.
 *   - the 4D lattice is split into contiguous t-slabs over the GPUs of the context
 *     (SURVEY.md section 8e).  gfb_init drives several GPUs from one process;
 *     gfb_init_rank is the one-process-per-GPU form (rank r owns global t in
 *     [r*NT/nranks, (r+1)*NT/nranks)).  Upload/download always take the GLOBAL host array and
 *     touch only the t-range(s) owned by this context.
 */
#ifndef GFB200_H
#define GFB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gfb_ctx gfb_ctx;
typedef struct gfb_gauge gfb_gauge; /* a gauge configuration: 4 link fields (Vector of Gaugefields_4D, src/API.jl:178-251) */
typedef struct gfb_mom gfb_mom;     /* 4 traceless anti-Hermitian fields, 8 real coefficients (TA_gaugefields_4D_MPILattice.jl:4-52) */

enum gfb_status {
    GFB_OK = 0,
    GFB_ERR_ARG = 1,     /* invalid argument (the Julia wrapper rethrows ArgumentError, molecular_dynamics.jl:447-465) */
    GFB_ERR_CUDA = 2,    /* CUDA runtime error */
    GFB_ERR_NCCL = 3,    /* NCCL error */
    GFB_ERR_NODEVICE = 4, /* no usable GPU: there is no CPU fallback */
    GFB_ERR_NUMERIC = 5   /* a site update failed (heatbath acceptance / normalisation), as the reference's error() at heatbathmodule.jl:1852-1857 */
};

enum gfb_integrator { GFB_QPQ = 0, GFB_PQP = 1 }; /* md_step! src/molecular_dynamics.jl:604-616 */
enum gfb_rng { GFB_PHILOX4X32 = 0 };              /* SiteRNGAlgorithm default, src/API.jl:189 */
enum gfb_energy_kind { GFB_E_CLOVER = 0, GFB_E_PLAQUETTE = 1 };

/* ---- context ------------------------------------------------------------------------------ */
int gfb_version(void);
/* single process driving ngpu local devices (devices==NULL -> 0..ngpu-1) */
int gfb_init(int ngpu, const int* devices, gfb_ctx** out);
/* one process per GPU: rank 0 calls gfb_nccl_unique_id and shares the 128 bytes (the host code uses
 * torch.distributed / MPI.bcast for that, like the reference's seed share AbstractGaugefields.jl:135-150) */
int gfb_nccl_unique_id(char* out128);
int gfb_init_rank(int rank, int nranks, const char* id128, int device, gfb_ctx** out);
int gfb_finalize(gfb_ctx* ctx);
const char* gfb_last_error(const gfb_ctx* ctx); /* ctx may be NULL: last error of a failed init */
int gfb_sync(gfb_ctx* ctx);
int gfb_num_slabs(const gfb_ctx* ctx, int* local, int* total);
/* stream timing of the asynchronous calls issued between tic and toc (CUDA events on the compute stream) */
int gfb_timer_tic(gfb_ctx* ctx);
int gfb_timer_toc(gfb_ctx* ctx, double* milliseconds);
/* number of kernels this library has launched so far on this context */
int gfb_kernel_launches(const gfb_ctx* ctx, long long* count);
/* pinned host memory for upload/download buffers */
int gfb_host_alloc(void** ptr, size_t bytes);
int gfb_host_free(void* ptr);

/* ---- fields ------------------------------------------------------------------------------- */
/* gauge_configuration / similar(U)  (src/API.jl:178-251) */
int gfb_gauge_alloc(gfb_ctx* ctx, int nx, int ny, int nz, int nt, gfb_gauge** out);
int gfb_gauge_free(gfb_gauge* g);
/* gauge_momenta / initialize_TA_Gaugefields (src/API.jl:331, src/TA_Gaugefields.jl:140-195); zero-filled */
int gfb_mom_alloc(gfb_ctx* ctx, int nx, int ny, int nz, int nt, gfb_mom** out);
int gfb_mom_free(gfb_mom* p);
/* host <-> device in the gathered layout (gather_matrix, src/API.jl:516-533); host = global array of one direction */
int gfb_gauge_upload(gfb_gauge* g, int mu, const double* host);
int gfb_gauge_download(const gfb_gauge* g, int mu, double* host);
/* ILDG binary payload of the whole lattice (big-endian, [t][z][y][x][mu][row][col] complex, 32- or 64-bit floats: what
 * _save_binarydata writes and load_gaugefield! reads site by site, src/output/ildg_format.jl:697-746, 67-83) <-> device.
 * Byte swap, precision conversion and the layout transpose run on the GPU; in one-process-per-GPU mode every rank passes the
 * global payload pointer and reads/writes only its own time-slices.  The LIME container is the host's business. */
int gfb_gauge_upload_ildg(gfb_gauge* g, const void* payload, int precision);
int gfb_gauge_download_ildg(const gfb_gauge* g, void* payload, int precision);
int gfb_mom_upload(gfb_mom* p, int mu, const double* host);
int gfb_mom_download(const gfb_mom* p, int mu, double* host);
/* copy_configuration! / substitute_U! (src/API.jl:307-322) */
int gfb_gauge_copy(gfb_gauge* dst, const gfb_gauge* src);
int gfb_mom_copy(gfb_mom* dst, const gfb_mom* src);
int gfb_mom_zero(gfb_mom* p); /* clear_U! on momenta, TA_gaugefields_4D_MPILattice.jl:285-291 */
/* add_U!(P, t, F) on momenta (TA_gaugefields_4D_serial.jl:150-173) */
int gfb_mom_axpy(gfb_mom* p, double t, const gfb_mom* f);
/* the same on one direction: add_U!(P[mu], t, F[mu]) (molecular_dynamics.jl:580-582) */
int gfb_mom_axpy_dir(gfb_mom* p, int mu, double t, const gfb_mom* f);

/* ---- initial fields and random numbers ------------------------------------------------------ */
int gfb_set_cold(gfb_gauge* g);                              /* IdentityGauges, AbstractGaugefields.jl:431-548 */
int gfb_set_hot(gfb_gauge* g, uint64_t seed, int rng_alg);   /* hot start keyed per global site, gaugefields_4D_MPILattice.jl:430-472 */
/* gaussian_momenta!(p; sigma, seed, sweep, rng) (src/API.jl:341-368, TA_gaugefields_4D_MPILattice.jl:157-193) */
int gfb_gaussian_momenta(gfb_mom* p, uint64_t seed, uint64_t sweep, double sigma, int rng_alg);
int gfb_reunitarize(gfb_gauge* g);                           /* normalize_U!, gaugefields_4D_nowing.jl:2387-2458 */

/* ---- observables (write one value to host, synchronous) ------------------------------------- */
/* calculate_Plaquette: un-normalised sum_{x,mu<nu} Re tr P (AbstractGaugefields.jl:2684-2699) */
int gfb_plaquette_sum(gfb_gauge* g, double* out);
/* real(evaluate_GaugeAction) for the action beta/2*(plaq+plaq') = beta * plaquette_sum (GaugeActions.jl:132-142) */
int gfb_wilson_action(gfb_gauge* g, double beta, double* out);
/* p*p = sum c_a^2 (TA_Gaugefields.jl:127-137) */
int gfb_kinetic(gfb_mom* p, double* out);
/* md_hamiltonian = -(beta/3)*plaquette_sum + p*p/2 (molecular_dynamics.jl:494-505, :247-249) */
int gfb_hamiltonian(gfb_gauge* g, gfb_mom* p, double beta, double* out);
/* clover energy density (samples/measurements/energydensity.jl:4-78) or its plaquette form */
int gfb_energy_density(gfb_gauge* g, int kind, double* out);
/* calculate_Polyakov_loop (AbstractGaugefields.jl:2929-2956): out2 = (re, im), averaged over spatial sites */
int gfb_polyakov(gfb_gauge* g, double* out2);

/* ---- updates (asynchronous, stream-ordered) -------------------------------------------------- */
/* md_force!(F, ::GaugeAction, U, ws) for the Wilson action (molecular_dynamics.jl:251-267) */
int gfb_force(gfb_mom* f, gfb_gauge* g, double beta);
/* update_momenta!(P, U, eps, driver) (molecular_dynamics.jl:539-551): fused staple->TA->kick */
int gfb_update_momenta(gfb_mom* p, gfb_gauge* g, double eps, double beta);
/* update_gaugefields!(U, P, eps, driver) (molecular_dynamics.jl:513-531) */
int gfb_update_links(gfb_gauge* g, const gfb_mom* p, double eps);
/* md_trajectory!(U, p, driver) (molecular_dynamics.jl:712-730): `steps` md_step!s of size tau/steps.
 * fused = 0: each step issues the reference's op sequence (QPQ: link, kick, link).
 * fused = 1: the kick and the following link update run in one kernel and adjacent half link
 *            updates are merged (same trajectory up to rounding).  H = {initial, final}; NULL skips
 *            the two Hamiltonian evaluations (diagnostics=false). */
int gfb_md_trajectory(gfb_gauge* g, gfb_mom* p, double beta, int steps, double tau, int integrator, int fused, double* H);
/* flow!(U, ::Gradientflow) (src/smearing/gradientflow.jl:171-238): nsteps Luescher RK3 steps of size eps */
int gfb_flow(gfb_gauge* g, double eps, int nsteps);
/* add_force!(F, U; plaqonly=true) after clear (AbstractGaugefields.jl:2717-2762): F = TA(U_mu V_mu^dag) */
int gfb_flow_force(gfb_mom* f, gfb_gauge* g);
/* exp_aF_U!(W, a, F, U) (AbstractGaugefields.jl:2810-2841); w may alias u */
int gfb_exp_aF_U(gfb_gauge* w, double a, const gfb_mom* f, const gfb_gauge* u);
/* STOUT_Layer forward! (src/smearing/stout_fast.jl:250-274): out = exp(TA(rho*C_mu U_mu^dag)) U_mu;
 * q (may be NULL) receives the 8 coefficients of Q_mu, the tape the backward pass needs */
int gfb_stout_forward(gfb_gauge* out, gfb_gauge* in, double rho, gfb_mom* q);
/* back_prop! through one stout layer (src/smearing/Abstractsmearing.jl:352-411, stout_fast.jl:317-407):
 * given d_out = dS/dU' (matrix field in the reference's dSdU convention) produce d_in = dS/dU */
int gfb_stout_backward(gfb_gauge* d_in, gfb_gauge* d_out, gfb_gauge* in, double rho);
/* momentum kick from an explicit derivative field: P_mu += factor * TAcoeffs(U_mu * dSdU_mu)
 * (molecular_dynamics.jl:255-265 with an external dSdU, test/HMCstout_test_nowing.jl:99-118) */
int gfb_kick_from_dSdU(gfb_mom* p, gfb_gauge* u, gfb_gauge* dsdu, double factor);
/* calc_dSdUmu! for the Wilson action at coefficient beta/2 (GaugeActions.jl:95-123): d = (beta/2) * sum of 6 staples */
int gfb_wilson_dSdU(gfb_gauge* d, gfb_gauge* g, double beta);

/* ---- general-action path: plaquette + rectangle terms (Symanzik / Iwasaki / DBW2 type actions) ----------------------------
 * A GaugeAction with the terms (c_plaq, plaquette + plaquette') and (c_rect, rectangular + rectangular')
 * (GaugeAction/push!, src/action/GaugeActions.jl:23-62; "rectangular" = the 1x2 and 2x1 loops of every plane,
 * src/autostaples/wilsonloops.jl:233-245).  dSdU_mu = c_plaq * (6 plaquette staples) + c_rect * (18 rectangle staples)
 * (calc_dSdUmu!, GaugeActions.jl:95-123).  c_rect = 0 runs the Wilson kernels.  On a t-slab decomposition the rectangle kernels
 * work on a copy of each slab with two halo slices either side (the reference's NDW = 2 wing), so a slab needs >= 2 slices. */
/* out2 = { sum_{x, mu<nu} Re tr P_munu, sum_x Re tr of the 12 rectangle loops }: evaluate_GaugeAction's building blocks
 * (GaugeActions.jl:132-142); Re evaluate_GaugeAction = 2 * (c_plaq * out2[0] + c_rect * out2[1]) */
int gfb_loop_sums(gfb_gauge* g, double* out2);
/* md_force!(F, action, U, ws) (molecular_dynamics.jl:251-267): F_mu = -(1/NC) TAcoeffs(U_mu dSdU_mu) */
int gfb_force_general(gfb_mom* f, gfb_gauge* g, double c_plaq, double c_rect);
/* update_momenta!(P, U, eps, driver) for that action (molecular_dynamics.jl:539-551) */
int gfb_update_momenta_general(gfb_mom* p, gfb_gauge* g, double eps, double c_plaq, double c_rect);
/* md_hamiltonian = -(1/NC) Re evaluate_GaugeAction + p*p/2 (molecular_dynamics.jl:494-505, 247-249) */
int gfb_hamiltonian_general(gfb_gauge* g, gfb_mom* p, double c_plaq, double c_rect, double* out);
/* md_trajectory! for that action; arguments as gfb_md_trajectory */
int gfb_md_trajectory_general(gfb_gauge* g, gfb_mom* p, double c_plaq, double c_rect, int steps, double tau, int integrator, int fused, double* H);
/* flow!(U, ::Gradientflow_general) (src/smearing/gradientflow.jl:240-316) with link values (c_plaq, c_rect) for the loop sets
 * ("plaquette", "rectangular"); (1, 0) is gfb_flow */
int gfb_flow_general(gfb_gauge* g, double eps, int nsteps, double c_plaq, double c_rect);
/* topological_charge(U; method) / topological_charge_density(U; method) (src/AbstractGaugefields.jl:1447-1490):
 * method 0 :plaquette, 1 :clover, 2 :improved.  host_density = Float64[NX,NY,NZ,NT]; the reference supports serial fields only
 * (AbstractGaugefields.jl:1403-1434), this backend also t-slab decompositions. */
enum gfb_topo_method { GFB_Q_PLAQUETTE = 0, GFB_Q_CLOVER = 1, GFB_Q_IMPROVED = 2 };
int gfb_topological_charge(gfb_gauge* g, int method, double* out);
int gfb_topological_charge_density(gfb_gauge* g, int method, double* host_density);

/* ---- heatbath and overrelaxation (the other quenched updater) ---------------------------------------------------------------
 * heatbath!(U, h::Heatbath) / overrelaxation!(U, h) for the Wilson action (src/heatbath/heatbathmodule.jl:481-650): one sweep =
 * 4 directions x 2 checkerboard colours; SU(3) by the subgroup sequence (1,2),(2,3),(1,3) with Kennedy-Pendleton sampling
 * (src/heatbath/portable/kernels.jl:63-270); overrelaxation by three random-subgroup reflections (heatbathmodule.jl:1243-1322).
 * Streams are keyed by (seed, sweep, direction, colour, subgroup) and the GLOBAL site, like the reference's
 * (heatbathmodule.jl:1815-1818): a sweep does not depend on the t-slab decomposition.  Even extents required. */
int gfb_heatbath(gfb_gauge* g, double beta, uint64_t seed, uint64_t sweep, int rng_alg);
int gfb_overrelaxation(gfb_gauge* g, double beta, uint64_t seed, uint64_t sweep, int rng_alg);

/* ---- primitive table ---------------------------------------------------------------------------
 * The element-wise operations that Gaugefields.jl's generic (un-fused) algorithms are written in; each forwards to one
 * LatticeMatrices kernel in the reference (src/4D/mpi_jacc/gaugefields_4D_MPILattice.jl:474-842,
 * TA_gaugefields_4D_MPILattice.jl:147-291; semantics: SURVEY.md Appendix A).  A gfb_field is ONE 3x3 complex matrix
 * field (one Gaugefields_4D: a link direction or a temporary from similar(U[1])).  Asynchronous unless they return a scalar. */
typedef struct gfb_field gfb_field;
int gfb_field_alloc(gfb_ctx* ctx, int nx, int ny, int nz, int nt, gfb_field** out);   /* similar(U[1]) */
/* U[mu] of a configuration as a field that aliases the configuration's memory (no copy); free it before the configuration */
int gfb_field_view(gfb_gauge* g, int mu, gfb_field** out);
int gfb_field_free(gfb_field* f);
int gfb_field_upload(gfb_field* f, const double* host);       /* ComplexF64[3,3,NX,NY,NZ,NT] */
int gfb_field_download(gfb_field* f, double* host);
int gfb_field_clear(gfb_field* f);                            /* clear_U!  (gaugefields_4D_MPILattice.jl:731-733) */
int gfb_field_unit(gfb_field* f);                             /* unit_U!   (:811-815) */
int gfb_field_copy(gfb_field* dst, gfb_field* src, const int* shift4, int dagger); /* substitute_U!(a, b | shifted | adjoint) (:509-575) */
/* mul!(C, A, B, alpha, beta): C = alpha * op(A(x+shiftA)) * op(B(x+shiftB)) + beta * C, op = identity or dagger
 * (src/AbstractGaugefields.jl:2082-2105, 2906-2914; shift_U :647-675).  shift = NULL means no shift; C must not alias A or B. */
int gfb_mul(gfb_field* c, gfb_field* a, const int* shift_a4, int dag_a, gfb_field* b, const int* shift_b4, int dag_b, double alpha_re,
            double alpha_im, double beta_re, double beta_im);
/* add_U!(C, alpha, A | A') : C += alpha * op(A)  (:739-772) */
int gfb_axpy(gfb_field* c, double alpha_re, double alpha_im, gfb_field* a, int dag_a);
/* tr(A) and tr(A, B) = sum_x tr(A(x) B(x))  (:721-728); out2 = (re, im) */
int gfb_tr(gfb_field* a, double* out2);
int gfb_tr2(gfb_field* a, gfb_field* b, double* out2);
/* Traceless_antihermitian!(Q, M) matrix -> matrix (:774-781) */
int gfb_ta_project(gfb_field* q, gfb_field* m);
/* Traceless_antihermitian_add!(P[mu], factor, M) matrix -> 8 coefficients (TA_gaugefields_4D_MPILattice.jl:263-283) */
int gfb_ta_coeffs_add(gfb_mom* p, int mu, double factor, gfb_field* m);
/* exptU!(E, t, Q) with Q a matrix field: E = exp(t * TA(Q)) (:798-808);  exptU!(E, t, P[mu]) for momenta (TA_...:196-210) */
int gfb_exp(gfb_field* e, double t, gfb_field* q);
int gfb_exp_mom(gfb_field* e, double t, gfb_mom* p, int mu);

#ifdef __cplusplus
}
#endif
#endif /* GFB200_H */
