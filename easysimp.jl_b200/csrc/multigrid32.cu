// multigrid32.cu -- FP64 CG preconditioned by the geometric-multigrid V-cycle evaluated in FP32
// (esimp_params.precond = ESIMP_PRECOND_GMG_FP32).
//
// STATUS: opt-in and NOT YET RUN ON A GPU (written after the round-1 GPU budget was spent; the double
// instantiations of the shared kernel were checked to compile to bit-identical SASS).  Nothing selects it by default.
//
// Why: the solve (Optimization.jl:317) dominates the SIMP iteration and, inside it, the smoother / residual
// kernels of the V-cycle, which are bound by the FP64 pipe.  The preconditioner only has to be a good, fixed,
// (nearly) symmetric approximation of A^-1: evaluated in single precision it needs the same number of CG
// iterations at the tolerances the loop uses (scripts/gmg_fp32_study.py: 15/15, 43/43, 36/36 at rtol 1e-8 with a
// modulus contrast of 1e9), while every vector of the cycle is half as wide and the arithmetic leaves the FP64
// pipe.  The outer CG -- operator, dot products, x and r updates -- stays FP64, so the displacement converges to
// the same residual test as with the other preconditioners.
//
// How: the hierarchy, the coarse moduli, the inverse diagonals and lmax come from mg_setup (multigrid.cu) in
// FP64 and are mirrored to float per level; r is scaled to unit 2-norm on entry (the cycle is linear), z is scaled
// back inside the direction / dot kernels, so no intermediate leaves the FP32 range.  The smoother and residual are
// the float instantiation of the fused operator kernel (hex8_structured.cu).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>

#include "ctx.cuh"

#define LAUNCH_STREAM(kern, ...)                                        \
    do {                                                                \
        kern<<<RED_BLOCKS, RED_THREADS, 0, ctx->stream>>>(__VA_ARGS__); \
        ctx->launches++;                                                \
        ESIMP_CUDA(cudaGetLastError());                                 \
    } while (0)

// ---------------------------------------------------------------------------- kernels
#include "mg32_kernels.cuh"

// ---------------------------------------------------------------------------- float mirror of the hierarchy
// While a replicated level is processed the context behaves like a single-rank one (see multigrid.cu)
struct LocalScope32 {
    esimp_ctx* c;
    int saved;
    LocalScope32(esimp_ctx* ctx, bool on) : c(ctx), saved(ctx->nranks) {
        if (on) c->nranks = 1;
    }
    ~LocalScope32() { c->nranks = saved; }
};

// allocate (once) and refresh (every solve: E and Dinv change with the densities) the float arrays of every level
static int mg32_refresh(esimp_ctx* ctx) {
    for (auto& L : ctx->mg) {
        if (!L.f_E) {
            ESIMP_CUDA(cudaMalloc(&L.f_E, sizeof(float) * L.ncell_loc));
            float** vecs[] = {&L.f_dinv, &L.f_r, &L.f_z, &L.f_z2, &L.f_d, &L.f_q};
            for (float** v : vecs) {
                ESIMP_CUDA(cudaMalloc(v, sizeof(float) * L.ndof_loc));
                ESIMP_CUDA(cudaMemsetAsync(*v, 0, sizeof(float) * L.ndof_loc, ctx->stream));
            }
        }
        LAUNCH_STREAM(k32_convert, L.d_E, L.f_E, L.ncell_loc);
        LAUNCH_STREAM(k32_convert, L.d_dinv, L.f_dinv, L.ndof_loc);
    }
    return ESIMP_OK;
}

static int halo32(esimp_ctx* ctx, const MGLevel& L, float* v) {
    if (ctx->nranks <= 1) return ESIMP_OK;
    return comm_halo_nodes_lv_f32(ctx, L.vw.g, L.vw.own_p0, L.vw.own_p1, L.has_left, L.has_right, v);
}

// k Chebyshev-Jacobi iterations on level L for A z = r, all in float; first_done: the zero-guess first step was
// already applied by the caller (k32_enter)
static int mg32_smooth(esimp_ctx* ctx, MGLevel& L, int k, bool zero_guess, bool first_done = false) {
    const double lmax = L.lmax, lmin = lmax / ctx->mg_ratio;
    const double theta = 0.5 * (lmax + lmin), delta = 0.5 * (lmax - lmin), sigma = theta / delta;
    double rho = 1.0 / sigma;
    int done = 0;
    if (zero_guess) {
        if (!first_done) LAUNCH_STREAM(k32_cheb_first, L.f_r, L.f_dinv, (float)(1.0 / theta), L.f_d, L.f_z, L.own_dof0, L.own_dof1);
        done = 1;
    }
    for (int i = done; i < k; ++i) {
        double c1, c2;
        if (i == 0) {  // restart of the recurrence from the current iterate
            c1 = 0.0;
            c2 = 1.0 / theta;
        } else {
            double rho_new = 1.0 / (2.0 * sigma - rho);
            c1 = rho_new * rho;
            c2 = 2.0 * rho_new / delta;
            rho = rho_new;
        }
        // z' = z + c1 d + c2 Dinv (r - A z), fused into the operator's store
        ESIMP_TRY(halo32(ctx, L, L.f_z));
        ESIMP_TRY(hex8_smooth_view_f32(ctx, L.vw, L.f_E, L.f_z, L.f_z2, L.f_r, L.f_dinv, L.f_d, (float)c1, (float)c2));
        std::swap(L.f_z, L.f_z2);
    }
    return ESIMP_OK;
}

// z_l = V(r_l) in float
static int mg32_vcycle(esimp_ctx* ctx, int l, bool first_done = false) {
    MGLevel& L = ctx->mg[l];
    if (l == (int)ctx->mg.size() - 1) return mg32_smooth(ctx, L, l == 0 ? ctx->mg_nu : ctx->mg_coarse_its, true, first_done);
    MGLevel& C = ctx->mg[l + 1];
    ESIMP_TRY(mg32_smooth(ctx, L, ctx->mg_nu, true, first_done));
    // residual fused into the operator's store: q = r - A z; prescribed rows are zeroed by the fix-up (z = 0 there)
    ESIMP_TRY(halo32(ctx, L, L.f_z));
    ESIMP_TRY(hex8_resid_view_f32(ctx, L.vw, L.f_E, L.f_z, L.f_q, L.f_r));
    if (L.n_fixed > 0) {
        k32_fixup<<<(unsigned)((L.n_fixed + 255) / 256), 256, 0, ctx->stream>>>(L.d_fixed, L.n_fixed, (float)L.diag_value, L.f_z, L.f_q);
        ctx->launches++;
        ESIMP_CUDA(cudaGetLastError());
    }
    ESIMP_TRY(halo32(ctx, L, L.f_q));
    if (C.replicated && !L.replicated) {
        // distributed -> replicated (see mg_vcycle in multigrid.cu): partial restriction, one allreduce, then no communication
        const int fp0g = L.vw.g.gp0 + L.vw.own_p0, fp1g = L.vw.g.gp0 + L.vw.own_p1;
        ESIMP_CUDA(cudaMemsetAsync(C.f_r, 0, sizeof(float) * C.ndof_loc, ctx->stream));
        LAUNCH_STREAM(k32_restrict, L.vw.g, C.vw.g, (fp0g + 1) / 2, (fp1g - 1) / 2 + 1, L.f_q, C.d_mask, C.f_r);
        ESIMP_TRY(comm_allreduce_f32(ctx, C.f_r, C.ndof_loc));
        LocalScope32 ls(ctx, true);
        ESIMP_TRY(mg32_vcycle(ctx, l + 1));
    } else {
        LAUNCH_STREAM(k32_restrict, L.vw.g, C.vw.g, C.vw.own_p0, C.vw.own_p1, L.f_q, C.d_mask, C.f_r);
        ESIMP_TRY(mg32_vcycle(ctx, l + 1));
        ESIMP_TRY(halo32(ctx, C, C.f_z));
    }
    {
        const int nyz = (C.vw.g.ny + 1) * (C.vw.g.nz + 1);
        dim3 grid((unsigned)std::min(64, (nyz + 255) / 256), (unsigned)std::min(65535, (L.vw.own_p1 - L.vw.own_p0) * 3));
        k32_prolong_add<<<grid, 256, 0, ctx->stream>>>(L.vw.g, C.vw.g, L.vw.own_p0, L.vw.own_p1, C.f_z, L.d_mask, L.f_z);
        ctx->launches++;
        ESIMP_CUDA(cudaGetLastError());
    }
    return mg32_smooth(ctx, L, ctx->mg_nu, false);
}

// z32 = V(r / |r|) on level 0; the caller scales by s = |r| (rr = |r|^2, global)
static int mg32_precondition(esimp_ctx* ctx, double rr, double* s_out) {
    MGLevel& L0 = ctx->mg[0];
    const double s = sqrt(rr);
    *s_out = s;
    const double lmax0 = L0.lmax, theta0 = 0.5 * (lmax0 + lmax0 / ctx->mg_ratio);
    LAUNCH_STREAM(k32_enter, ctx->d_r, 1.0 / s, L0.f_dinv, (float)(1.0 / theta0), L0.f_r, L0.f_d, L0.f_z, ctx->own_dof0, ctx->own_dof1);
    return mg32_vcycle(ctx, 0, true);
}

static int apply_fine(esimp_ctx* ctx, MGLevel& L0, double* in, double* out, int dot_slot) {
    if (ctx->nranks > 1) ESIMP_TRY(comm_halo_nodes_lv(ctx, L0.vw.g, L0.vw.own_p0, L0.vw.own_p1, L0.has_left, L0.has_right, in));
    ESIMP_TRY(hex8_matvec_view(ctx, L0.vw, in, out, ctx->prm.matvec_variant, dot_slot));
    return fixup_lv(ctx, L0.d_fixed, L0.n_fixed, L0.diag_value, in, out);
}

// Same contract as pcg_solve_mg (multigrid.cu); the hierarchy must have been built by mg_setup.
int pcg_solve_mg32(esimp_ctx* ctx, int* iters, double* relres) {
    if (ctx->mg.empty()) return esimp_fail(ctx, ESIMP_ERR_STATE, "multigrid hierarchy not built");
    ESIMP_TRY(mg32_refresh(ctx));
    MGLevel& L0 = ctx->mg[0];
    const long long i0 = ctx->own_dof0, i1 = ctx->own_dof1;
    double* S = ctx->d_scal;
    static const int RZ[2] = {SC_RZ0, SC_RZ1};
    ESIMP_TRY(apply_fine(ctx, L0, ctx->d_u, ctx->d_q, -1));
    {
        RedOut<2> out{{S + SC_RR0, S + SC_FF}};
        LAUNCH_STREAM(k32_init, ctx->d_f, ctx->d_q, ctx->d_r, i0, i1, ctx->d_partial, ctx->d_ticket, out);
        ESIMP_TRY(allreduce_scalars(ctx, SC_RR0, 1, false));
        ESIMP_TRY(allreduce_scalars(ctx, SC_FF, 1, false));
    }
    ESIMP_TRY(fetch_scalars(ctx));
    const double ff = ctx->h_scal[SC_FF];
    ctx->f_norm2 = ff;
    const double tol2 = ctx->prm.pcg_rtol * ctx->prm.pcg_rtol * ff;
    double rr = ctx->h_scal[SC_RR0];
    int it = 0, status = ESIMP_OK;
    if (ff == 0.0) {
        ESIMP_CUDA(cudaMemsetAsync(ctx->d_u, 0, sizeof(double) * ctx->ndof_loc, ctx->stream));
        rr = 0.0;
    }
    double s = 1.0;
    if (rr > tol2) {
        ESIMP_TRY(mg32_precondition(ctx, rr, &s));  // z = M^-1 r = s z32
        RedOut<1> out{{S + RZ[0]}};
        LAUNCH_STREAM(k32_dot_rz, ctx->d_r, L0.f_z, s, i0, i1, ctx->d_partial, ctx->d_ticket, out);
        ESIMP_TRY(allreduce_scalars(ctx, RZ[0], 1, false));
        LAUNCH_STREAM(k32_direction, L0.f_z, s, ctx->d_p, ctx->d_u, i0, i1, S + RZ[0], S + RZ[0], S + SC_PQ, 1);
    }
    const bool fused_dot = ctx->prm.matvec_variant != ESIMP_MATVEC_DENSE_ATOMIC;
    while (rr > tol2 && it < ctx->prm.pcg_maxit) {
        const int cur = it & 1, nxt = cur ^ 1;
        if (fused_dot) {  // q = A p with <p, q> fused into the operator's store
            ESIMP_TRY(apply_fine(ctx, L0, ctx->d_p, ctx->d_q, SC_PQ));
            ESIMP_TRY(allreduce_scalars(ctx, SC_PQ, 1, false));
        } else {
            ESIMP_TRY(apply_fine(ctx, L0, ctx->d_p, ctx->d_q, -1));
            ESIMP_TRY(vec_dot(ctx, ctx->d_p, ctx->d_q, SC_PQ));
        }
        {
            RedOut<1> out{{S + SC_RR0}};
            LAUNCH_STREAM(k32_update, ctx->d_q, ctx->d_r, i0, i1, S + RZ[cur], S + SC_PQ, ctx->d_partial, ctx->d_ticket, out);
            ESIMP_TRY(allreduce_scalars(ctx, SC_RR0, 1, false));
        }
        ++it;
        ESIMP_TRY(fetch_scalars(ctx));
        rr = ctx->h_scal[SC_RR0];
        if (getenv("ESIMP_MG_DEBUG") && ctx->rank == 0)
            fprintf(stderr, "[mg32] it %d rr %.6e pq %.6e rz %.6e\n", it, rr, ctx->h_scal[SC_PQ], ctx->h_scal[RZ[cur]]);
        if (!(rr == rr)) return esimp_fail(ctx, ESIMP_ERR_STATE, "PCG breakdown: residual is NaN");
        if (rr <= tol2 || it >= ctx->prm.pcg_maxit) {  // only the deferred x += alpha p is left to do
            LAUNCH_STREAM(k32_direction, L0.f_z, s, ctx->d_p, ctx->d_u, i0, i1, S + RZ[cur], S + RZ[cur], S + SC_PQ, 2);
            break;
        }
        ESIMP_TRY(mg32_precondition(ctx, rr, &s));
        {
            RedOut<1> out{{S + RZ[nxt]}};
            LAUNCH_STREAM(k32_dot_rz, ctx->d_r, L0.f_z, s, i0, i1, ctx->d_partial, ctx->d_ticket, out);
            ESIMP_TRY(allreduce_scalars(ctx, RZ[nxt], 1, false));
        }
        LAUNCH_STREAM(k32_direction, L0.f_z, s, ctx->d_p, ctx->d_u, i0, i1, S + RZ[nxt], S + RZ[cur], S + SC_PQ, 0);
    }
    if (rr > tol2) status = ESIMP_WARN_PCG_MAXIT;
    if (iters) *iters = it;
    if (relres) *relres = ff > 0 ? sqrt(rr / ff) : 0.0;
    return status;
}
