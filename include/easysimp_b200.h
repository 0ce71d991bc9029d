/*
 * easysimp_b200.h -- C ABI of libeasysimp_b200.so
 *
 * B200-native (sm_100a) replacement for the per-iteration SIMP loop of jezekon/EasySIMP.jl.
 * The reference has no FFI of its own (pure Julia); every entry point below names the Julia
 * function (file:line under the reference tree) whose work it takes over.  The Julia host
 * keeps simp_optimize / OptimizationParameters / FilterCache / create_filter_cache /
 * assemble_stiffness_matrix_simp! and reaches this library through `ccall`
 * (binding shown in INTEGRATION.md and easysimp.jl_b200/julia/).
 *
 * Conventions
 *   - every function returns int: 0 = OK, < 0 = error (esimp_last_error gives the text),
 *     > 0 = numerical warning (ESIMP_WARN_*), results are still valid.
 *   - all pointers are HOST memory; the library copies on entry and never keeps them.
 *   - indices are 0-based.  Cell arrays are in the caller's cell order, dof arrays in the
 *     caller's dof numbering (Ferrite DofHandler numbering when `dof_of_node` is given,
 *     3*node+component otherwise).
 *   - a context is bound to one CUDA device and is not thread-safe.
 *   - no CPU fallback exists: without a usable sm_100 device esimp_create fails.
 */
#ifndef EASYSIMP_B200_H
#define EASYSIMP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ESIMP_VERSION 100 /* 0.1.0 */

/* status codes */
#define ESIMP_OK 0
#define ESIMP_ERR_ARG (-1)
#define ESIMP_ERR_CUDA (-2)
#define ESIMP_ERR_STATE (-3)
#define ESIMP_ERR_NCCL (-4)
#define ESIMP_ERR_NODEVICE (-5)
#define ESIMP_WARN_PCG_MAXIT 1   /* PCG stopped at pcg_maxit before reaching pcg_rtol            */
#define ESIMP_WARN_OC_MAXIT 2    /* OC bisection used all 200 steps (OptimalityCriteria.jl:140) */

/* filter_type (OptimizationParameters.filter_type::Symbol, Optimization.jl:67) */
#define ESIMP_FILTER_SENSITIVITY 0
#define ESIMP_FILTER_DENSITY 1

/* cell types for esimp_set_mesh_unstructured */
#define ESIMP_CELL_TET4 4
#define ESIMP_CELL_HEX8 8

/* matvec kernel variants (structured grids) */
#define ESIMP_MATVEC_AUTO 0
#define ESIMP_MATVEC_DENSE_ATOMIC 1 /* one thread per element, dense 24x24, atomics (debug reference) */
#define ESIMP_MATVEC_FUSED 2        /* x-marching mode-space kernel, owner-computes tiles               */
#define ESIMP_MATVEC_FUSED_ATOMIC 3 /* same, tile perimeters combined with red.global.add.f64           */
#define ESIMP_MATVEC_FUSED_8x32 4   /* owner-computes with 8x32 (y,z) tiles instead of 16x16            */
#define ESIMP_MATVEC_FUSED_8x16 5   /* owner-computes with 8x16 tiles, 128-thread CTAs, 4 CTAs per SM  */
#define ESIMP_MATVEC_SWEEP 6        /* z-sweeping TMA / mbarrier kernel (AUTO picks it where eligible)  */

/* CG preconditioners */
#define ESIMP_PRECOND_JACOBI 0 /* diagonal scaling                                                        */
#define ESIMP_PRECOND_GMG 1    /* geometric multigrid V-cycle on structured grids (csrc/multigrid.cu)     */
#define ESIMP_PRECOND_GMG_FP32 2 /* the same V-cycle evaluated in single precision inside the FP64 CG (csrc/multigrid32.cu);
                                    opt-in, written for round 2 and not yet validated on a GPU                */

typedef struct esimp_ctx esimp_ctx;

/* POD mirror of OptimizationParameters (Optimization.jl:53-123) + solver knobs the
 * reference does not have (it solves directly with CHOLMOD, Optimization.jl:317). */
typedef struct esimp_params {
    double E0;              /* :55  */
    double Emin;            /* :56  */
    double nu;              /* :57  */
    double p;               /* :58  */
    double volume_fraction; /* :61  */
    double move_limit;      /* :70  */
    double damping;         /* :71  */
    double filter_radius;   /* ABSOLUTE radius R = ratio * estimate_element_size (FilterCommon.jl:72-73) */
    int32_t filter_type;    /* ESIMP_FILTER_*  (:67) */
    int32_t pcg_maxit;      /* max PCG iterations per solve                */
    double pcg_rtol;        /* stop when ||r||_2 <= pcg_rtol * ||f||_2      */
    int32_t pcg_check_every; /* residual is read back every this many iterations (>=1) */
    int32_t matvec_variant;  /* ESIMP_MATVEC_*                                          */
    int32_t warm_start;      /* 1: PCG starts from the previous displacement            */
    int32_t precond;         /* ESIMP_PRECOND_*                                         */
} esimp_params;

/* per-iteration scalars the reference logs (OptimizationLogger.jl:26-29,
 * Optimization.jl:320-325,374-379) + solver diagnostics + phase timers (CUDA events). */
typedef struct esimp_iter_stats {
    double energy;   /* 0.5 u'Ku of the BC-modified K (Optimization.jl:320)      */
    double volume;   /* dot(rho_phys, V)   (:321)                                */
    double change;   /* max |rho_new - rho_old| (:374)                           */
    double lambda;   /* last bisection midpoint (OptimalityCriteria.jl:145)     */
    double grayness; /* #{0.1 < rho_phys < 0.9} / n_cells (:377-378)             */
    double max_u;    /* max |u| (:379)                                           */
    double pcg_relres;
    int32_t pcg_iters;
    int32_t oc_steps;
    int32_t status; /* OR of ESIMP_WARN_* */
    int32_t reserved;
    double ms_filter, ms_solve, ms_sens, ms_oc, ms_total;
} esimp_iter_stats;

/* ---- lifecycle ------------------------------------------------------------------------ */
int esimp_version(void);
/* device < 0: use the current CUDA device */
int esimp_create(esimp_ctx** out, int device);
int esimp_destroy(esimp_ctx* ctx);
const char* esimp_last_error(const esimp_ctx* ctx); /* ctx may be NULL: last create error */

/* ---- multi-GPU: one process per GPU, x-slab partition ----------------------------------
 * Call esimp_comm_unique_id on rank 0, ship the 128 bytes to every rank (torch.distributed /
 * MPI / files), then esimp_comm_init on every rank BEFORE esimp_set_grid_structured.
 * NCCL (libnccl.so.2) is dlopen'ed on first use; single-GPU runs never need it. */
int esimp_comm_unique_id(void* id128);
int esimp_comm_init(esimp_ctx* ctx, int rank, int nranks, const void* id128);

/* ---- mesh -----------------------------------------------------------------------------
 * Structured: the grid Ferrite's generate_grid(Hexahedron,(nx,ny,nz),left,right) builds
 * (nodes and cells x-fastest).  dof_of_node (3*n_nodes, [node][comp]) is the DofHandler
 * numbering (FiniteElementAnalysis.jl:147-149, :257-279); NULL means dof = 3*node+comp.
 * Replaces setup_problem's allocate_matrix (FiniteElementAnalysis.jl:153): no K is built. */
int esimp_set_grid_structured(esimp_ctx* ctx, int nx, int ny, int nz, double hx, double hy, double hz,
                              const int64_t* dof_of_node);
/* Unstructured hex8 / tet4 (import_mesh grids).  conn: npe*n_cells node ids, celldofs:
 * (3*npe)*n_cells dof ids, both cell-major.  Builds colouring + per-element unit matrices
 * (initialize_element_cache, FiniteElementAnalysis.jl:608-630). Replicas only (no slabs). */
int esimp_set_mesh_unstructured(esimp_ctx* ctx, int celltype, int64_t n_cells, int64_t n_nodes,
                                const double* xyz, const int64_t* conn, const int64_t* celldofs);

/* ---- physics / BCs / load ------------------------------------------------------------- */
int esimp_default_params(esimp_params* p); /* OptimizationParameters defaults, Optimization.jl:87-102 */
int esimp_set_params(esimp_ctx* ctx, const esimp_params* p);
/* union of ch.prescribed_dofs of all ConstraintHandlers (homogeneous Dirichlet,
 * FiniteElementAnalysis.jl:293-340 + Ferrite apply!, Optimization.jl:581-583) */
int esimp_set_dirichlet(esimp_ctx* ctx, const int64_t* dofs, int64_t n);
/* assembled load vector (apply_load_condition!, LoadConditions.jl:112-154), length n_dofs */
int esimp_set_load(esimp_ctx* ctx, const double* f);
/* variable-density body force = simp_optimize's `acceleration_data = (acceleration_vector, base_density)`:
 * on EVERY assembly f = load + sum_e rho_phys_e * base_density * b * int N dOmega, cells whose density is below 1e-6
 * skipped (apply_variable_density_volume_force!, FiniteElementAnalysis.jl:486-526, called at Optimization.jl:300-311
 * and :517-528).  accel3 = NULL switches it off.  A body force alone is a valid right-hand side. */
int esimp_set_body_force(esimp_ctx* ctx, const double* accel3, double base_density);
/* stage call: f_body (length n_dofs, caller numbering) for density_data[n_cells] = rho_phys .* base_density */
int esimp_body_force(esimp_ctx* ctx, const double* accel3, const double* density_data, double* f_body);

/* ---- filter (FilterCache, FilterCommon.jl:40-98) ---------------------------------------- */
/* structured: stencil from R = params.filter_radius; uniform element volume */
int esimp_filter_build_structured(esimp_ctx* ctx);
/* general: CSR of FilterCache.neighbor_lists + centres (3*n_cells, [cell][xyz]) + volumes */
int esimp_filter_build_csr(esimp_ctx* ctx, const int64_t* rowptr, const int64_t* col, const double* centers,
                           const double* volumes, double R);

/* ---- stage calls (parity tests; assemble_stiffness_matrix_simp! drop-in) ---------------- */
/* physical densities -> E_e = Emin + (E0-Emin) rho^p (FiniteElementAnalysis.jl:107); this is
 * what assemble_stiffness_matrix_simp! (:637-712) reduces to in matrix-free form. */
int esimp_set_density(esimp_ctx* ctx, const double* rho_phys);
/* y = P K(rho) P u + m (I-P) u, m = mean|diag K|  (assembly + apply!) */
int esimp_matvec(esimp_ctx* ctx, const double* u, double* y);
/* diag of the BC-modified operator (Jacobi preconditioner) */
int esimp_diagonal(esimp_ctx* ctx, double* d);
/* u = K \ f by FP64 Jacobi-PCG (replaces cholesky(Symmetric(K,:L)) \ f, Optimization.jl:317).
 * u_inout: NULL or initial guess in / solution out. */
int esimp_solve(esimp_ctx* ctx, double* u_inout, int* iters, double* relres);
/* dc_e = -p rho^(p-1) (E0-Emin) u_e'K0u_e (SensitivityAnalysis.jl:31-80) with the context's
 * current density and displacement; energy = 0.5 sum E_e u_e'K0u_e (Optimization.jl:320) */
int esimp_sensitivities(esimp_ctx* ctx, double* dc, double* energy);
int esimp_filter_sens(esimp_ctx* ctx, const double* rho, const double* dc, double* out); /* SensitivityFilter.jl:33-67 */
int esimp_filter_dens(esimp_ctx* ctx, const double* rho, double* out);                   /* DensityFilter.jl:30-60   */
int esimp_filter_dens_T(esimp_ctx* ctx, const double* s, double* out);                   /* DensityFilter.jl:77-117  */
/* optimality_criteria_update (OptimalityCriteria.jl:69-146); dv = volume sensitivities */
int esimp_oc(esimp_ctx* ctx, const double* rho, const double* dc, const double* dv, double* rho_new,
             double* lambda, int* steps);

/* ---- the loop body (Optimization.jl:271-379) ------------------------------------------- */
/* rho = NULL: fill(volume_fraction) (Optimization.jl:222).  Also computes dV (:242-248). */
int esimp_init_design(esimp_ctx* ctx, const double* rho);
/* one SIMP iteration, state stays on the device */
int esimp_iterate(esimp_ctx* ctx, esimp_iter_stats* out);
/* final analysis (Optimization.jl:494-535): re-filter, re-solve, energy, volume */
int esimp_final_analysis(esimp_ctx* ctx, double* energy, double* volume, int* pcg_iters);

/* ---- post-solve fields for the export branches (Optimization.jl:407-477, :537-539) ----- */
/* For the displacement and physical densities on the device (after esimp_iterate / esimp_final_analysis / esimp_solve):
 *   von_mises_q1[n_cells]   von Mises stress of the FIRST Gauss point of each cell = the exported "von_mises_stress"
 *                           field (calculate_von_mises_stresses, PostProcessing.jl:204-222)
 *   von_mises_avg[n_cells]  von Mises stress of the Gauss-point average, whose maximum and first cell attaining it are
 *                           what calculate_stresses_simp returns (FiniteElementAnalysis.jl:566-597); max_cell is
 *                           0-based, -1 when every value is 0
 *   element_energy[n_cells] 1/2 sum_q sigma:eps dOmega (calculate_element_energy, PostProcessing.jl:172-198)
 * Lame parameters per cell from E(rho_phys) like material_model(density) (:577).  Any output pointer may be NULL. */
int esimp_stress_fields(esimp_ctx* ctx, double* von_mises_q1, double* von_mises_avg, double* element_energy,
                        double* max_von_mises, int64_t* max_cell);
/* the `stress_field` Dict itself, for the caller's cells [cell0, cell0+n): sigma[n][nq][6], nq = esimp_stress_nq
 * (8 Gauss points on hexahedra, first coordinate fastest; 1 on linear tetrahedra whose 4 points carry the same tensor),
 * 6 = SymmetricTensor{2,3} storage order (11, 21, 31, 22, 32, 33).  Chunk the range to bound host memory. */
int esimp_stress_tensors(esimp_ctx* ctx, int64_t cell0, int64_t n, double* sigma);
int esimp_stress_nq(const esimp_ctx* ctx);

/* ---- state I/O ------------------------------------------------------------------------- */
/* design variables.  On a multi-rank structured context only the x-slab this rank owns moves: esimp_get_design writes the
 * entries of its own cells and leaves the rest of `rho` untouched, esimp_set_design reads only those entries (ghost layers come
 * from their owners) -- 8 B per owned cell each way instead of the whole array on every rank.  The other getters keep
 * writing zeros outside the owned slab, so that per-rank results can be summed. */
int esimp_get_design(esimp_ctx* ctx, double* rho);
int esimp_get_physical_density(esimp_ctx* ctx, double* rho); /* filtered (== design for :sensitivity) */
int esimp_get_displacement(esimp_ctx* ctx, double* u);       /* caller's dof numbering     */
int esimp_set_displacement(esimp_ctx* ctx, const double* u);
/* overwrite the design variables only (the loop state a host-driven loop carries between iterations) */
int esimp_set_design(esimp_ctx* ctx, const double* rho);
int64_t esimp_n_dofs(const esimp_ctx* ctx);
int64_t esimp_n_cells(const esimp_ctx* ctx);

/* ---- multi-GPU from ONE process (the reference is a single Julia session, Optimization.jl:178) ----------------
 * An esimp_multi owns one context per device, an NCCL communicator initialised from N threads of this process and the x-slab
 * partition; every call below takes the same GLOBAL host arrays as its esimp_* namesake and fans out to one host thread per
 * GPU.  devices = NULL: devices 0 .. n_devices-1.  The scalars of esimp_iter_stats are allreduced inside the library (identical
 * on every GPU), its phase timers are those of the slowest GPU.  esimp_multi_rank_ctx exposes a rank's own context for the stage
 * calls that have no multi wrapper; calls that contain collectives must then be made on ALL ranks concurrently (one thread each). */
typedef struct esimp_multi esimp_multi;
int esimp_multi_create(esimp_multi** out, const int* devices, int n_devices);
int esimp_multi_destroy(esimp_multi* m);
const char* esimp_multi_last_error(const esimp_multi* m); /* m may be NULL: last create error */
int esimp_multi_n_gpus(const esimp_multi* m);
esimp_ctx* esimp_multi_rank_ctx(esimp_multi* m, int rank);
int64_t esimp_multi_n_dofs(const esimp_multi* m);
int64_t esimp_multi_n_cells(const esimp_multi* m);
int esimp_multi_set_params(esimp_multi* m, const esimp_params* p);
int esimp_multi_set_grid_structured(esimp_multi* m, int nx, int ny, int nz, double hx, double hy, double hz,
                                    const int64_t* dof_of_node);
int esimp_multi_set_dirichlet(esimp_multi* m, const int64_t* dofs, int64_t n);
int esimp_multi_set_load(esimp_multi* m, const double* f);
int esimp_multi_set_body_force(esimp_multi* m, const double* accel3, double base_density);
int esimp_multi_filter_build_structured(esimp_multi* m);
int esimp_multi_init_design(esimp_multi* m, const double* rho);
int esimp_multi_set_design(esimp_multi* m, const double* rho);
int esimp_multi_iterate(esimp_multi* m, esimp_iter_stats* out);
int esimp_multi_final_analysis(esimp_multi* m, double* energy, double* volume, int* pcg_iters);
int esimp_multi_get_design(esimp_multi* m, double* rho);
int esimp_multi_get_physical_density(esimp_multi* m, double* rho);
int esimp_multi_get_displacement(esimp_multi* m, double* u);

/* ---- measurement helpers (device-resident timing with CUDA events, bench.py) ----------- */
/* `reps` operator applies on device-resident vectors; ms_avg = mean launch-to-launch time.
 * bytes_per_apply = algorithmic bytes (16*n_dof + 8*n_cell). */
int esimp_bench_matvec(esimp_ctx* ctx, int reps, double* ms_avg, double* bytes_per_apply);
/* the same for one store mode of the structured operator kernel on the fine grid: 0 plain, 1 plain + <u,y>, 2 multigrid
 * smoother step fused into the store (reads r, Dinv, d; writes z', d), 3 residual r - A z, 4 smoother step + <r,z'>.
 * bytes_per_apply = 16 B/dof + 8 B/cell for modes 0-1, 48 B/dof + 8 B/cell for 2 and 4, 24 B/dof + 8 B/cell for 3. */
int esimp_bench_operator_mode(esimp_ctx* ctx, int mode, int reps, double* ms_avg, double* bytes_per_apply);
/* in-region timing: while enabled, every launch of the structured operator kernel is bracketed by
 * CUDA events on the context's stream; esimp_profile_matvec returns the summed kernel time, the number
 * of timed launches and the algorithmic bytes one launch moves on this rank. */
/* on = 1: only the plain / plain + dot operator launches are clocked (about 50 event pairs per SIMP step: cheap enough to stay
 * inside a timed region); on = 2: every category of esimp_profile_kernels as well (thousands of event pairs per step -- costs a few
 * per cent where the step is latency bound, N >= 4 GPUs); 0 = off */
int esimp_profile_enable(esimp_ctx* ctx, int on);
int esimp_profile_matvec(esimp_ctx* ctx, double* ms_sum, int64_t* count, double* bytes_per_apply);
/* same clock over EVERY fine-grid launch of the operator kernel, including the modes that fuse the multigrid smoother
 * step / residual into the store phase (those move more bytes, so they are kept out of the roofline figure) */
int esimp_profile_operator_all(esimp_ctx* ctx, double* ms_sum_all, int64_t* count_all);
/* the same clock per kernel category, fine grid only (sum of kernel milliseconds and number of launches since
 * esimp_profile_enable): 0 operator plain / plain + dot, 1 operator with the smoother step fused in, 2 operator residual
 * mode, 3 filter stencil, 4 sensitivities + energy, 5 OC candidate volumes, 6 prolongation onto the fine grid,
 * 7 restriction of the fine residual, 8 first smoother step fused with the CG update, 9 CG direction update,
 * 10 one whole coloured gather-scatter operator apply on an imported mesh (memset + one launch per colour).
 * n <= 16 entries are written. */
int esimp_profile_kernels(esimp_ctx* ctx, int n, double* ms_sum, int64_t* count);
/* number of kernels this context has launched so far */
int64_t esimp_launch_count(const esimp_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* EASYSIMP_B200_H */
