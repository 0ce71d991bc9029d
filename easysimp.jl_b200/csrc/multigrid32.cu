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
__global__ void k32_convert(const double* __restrict__ src, float* __restrict__ dst, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dst[i] = (float)src[i];
}
// level-0 entry of the cycle: r32 = r / |r|, and the first (zero-guess) Chebyshev step d = z = Dinv r32 / theta
__global__ void k32_enter(const double* __restrict__ r, double inv_s, const float* __restrict__ dinv, float inv_theta,
                          float* __restrict__ r32, float* __restrict__ d, float* __restrict__ z, long long i0, long long i1) {
    for (long long i = i0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < i1; i += (long long)gridDim.x * blockDim.x) {
        const float ri = (float)(r[i] * inv_s);
        r32[i] = ri;
        const float v = dinv[i] * ri * inv_theta;
        d[i] = v;
        z[i] = v;
    }
}
__global__ void k32_cheb_first(const float* __restrict__ r, const float* __restrict__ dinv, float inv_theta, float* __restrict__ d,
                               float* __restrict__ z, long long i0, long long i1) {
    for (long long i = i0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < i1; i += (long long)gridDim.x * blockDim.x) {
        const float v = dinv[i] * r[i] * inv_theta;
        d[i] = v;
        z[i] = v;
    }
}
// rows of prescribed dofs: out[d] = m in[d] (in is 0 there, so this zeroes the residual on them)
__global__ void k32_fixup(const long long* __restrict__ fixed, long long n, float m, const float* __restrict__ in,
                          float* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long d = fixed[i];
        out[d] = m * in[d];
    }
}
// r_c = P' res_f on the coarse planes [cp0, cp1) (27-point full weighting), 0 on prescribed coarse dofs
__global__ void k32_restrict(GridDev gf, GridDev gc, int cp0, int cp1, const float* __restrict__ resf,
                             const uint8_t* __restrict__ maskc, float* __restrict__ rc) {
    long long n = (long long)(cp1 - cp0) * 3 * gc.plane;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
        int K = (int)(idx % (gc.nz + 1));
        long long t = idx / (gc.nz + 1);
        int J = (int)(t % (gc.ny + 1));
        t /= (gc.ny + 1);
        int c = (int)(t % 3);
        int Lp = cp0 + (int)(t / 3);
        float acc = 0.0f;
        if (!((maskc[gc.node(Lp, J, K)] >> c) & 1)) {
            const int Gf = 2 * (gc.gp0 + Lp);
            const int lpf = Gf - gf.gp0;
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
                if (Gf + dx < 0 || Gf + dx > gf.nx || lpf + dx < 0 || lpf + dx >= gf.NPX) continue;
#pragma unroll
                for (int dy = -1; dy <= 1; ++dy) {
                    int jf = 2 * J + dy;
                    if (jf < 0 || jf > gf.ny) continue;
#pragma unroll
                    for (int dz = -1; dz <= 1; ++dz) {
                        int kf = 2 * K + dz;
                        if (kf < 0 || kf > gf.nz) continue;
                        float w = (dx ? 0.5f : 1.0f) * (dy ? 0.5f : 1.0f) * (dz ? 0.5f : 1.0f);
                        acc = fmaf(w, resf[gf.dof(lpf + dx, c, jf, kf)], acc);
                    }
                }
            }
        }
        rc[gc.dof(Lp, c, J, K)] = acc;
    }
}
// z_f += P z_c on the owned fine nodes (trilinear interpolation), prescribed fine dofs untouched; one thread per coarse
// (J,K) position of a (plane, component), updating the 2x2 patch of fine nodes (2J+a, 2K+b)
__global__ void __launch_bounds__(256) k32_prolong_add(GridDev gf, GridDev gc, int fp0, int fp1, const float* __restrict__ zc,
                                                       const uint8_t* __restrict__ maskf, float* __restrict__ zf) {
    const int nJ = gc.ny + 1, nK = gc.nz + 1, nJK = nJ * nK, nzf = gf.nz + 1;
    for (int pc = blockIdx.y; pc < (fp1 - fp0) * 3; pc += gridDim.y) {
        const int lp = fp0 + pc / 3, c = pc % 3;
        const int Gf = gf.gp0 + lp;
        const bool two = Gf & 1;  // odd fine plane: average of two coarse planes
        const float* __restrict__ c0 = zc + (long long)(((Gf >> 1) - gc.gp0) * 3 + c) * gc.plane;
        const float* __restrict__ c1 = zc + (long long)((((Gf + 1) >> 1) - gc.gp0) * 3 + c) * gc.plane;
        float* __restrict__ zrow = zf + (long long)(lp * 3 + c) * gf.plane;
        const uint8_t* __restrict__ mrow = maskf + (long long)lp * gf.plane;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nJK; i += gridDim.x * blockDim.x) {
            const int J = i / nK, K = i - J * nK;
            const bool hasJ = J < gc.ny, hasK = K < gc.nz;
            const int o00 = J * nK + K, o01 = o00 + (hasK ? 1 : 0), o10 = o00 + (hasJ ? nK : 0), o11 = o10 + (hasK ? 1 : 0);
            float v00 = c0[o00], v01 = c0[o01], v10 = c0[o10], v11 = c0[o11];
            if (two) {
                v00 = 0.5f * (v00 + c1[o00]);
                v01 = 0.5f * (v01 + c1[o01]);
                v10 = 0.5f * (v10 + c1[o10]);
                v11 = 0.5f * (v11 + c1[o11]);
            }
            const int f00 = (2 * J) * nzf + 2 * K;
            if (!((mrow[f00] >> c) & 1)) zrow[f00] += v00;
            if (hasK && !((mrow[f00 + 1] >> c) & 1)) zrow[f00 + 1] += 0.5f * (v00 + v01);
            if (hasJ) {
                const int f10 = f00 + nzf;
                if (!((mrow[f10] >> c) & 1)) zrow[f10] += 0.5f * (v00 + v10);
                if (hasK && !((mrow[f10 + 1] >> c) & 1)) zrow[f10 + 1] += 0.25f * ((v00 + v01) + (v10 + v11));
            }
        }
    }
}
// ---- FP64 pieces of the outer CG (the preconditioned residual arrives as s * z32)
__global__ void k32_init(const double* __restrict__ f, const double* __restrict__ q, double* __restrict__ r, long long i0, long long i1,
                         double* partial, unsigned int* ticket, RedOut<2> out) {
    double acc[2] = {0.0, 0.0};
    for (long long i = i0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < i1; i += (long long)gridDim.x * blockDim.x) {
        double fi = f[i], ri = fi - q[i];
        r[i] = ri;
        acc[0] = fma(ri, ri, acc[0]);
        acc[1] = fma(fi, fi, acc[1]);
    }
    grid_reduce<2, false>(acc, partial, ticket, out);
}
// r -= alpha q, <r, r>
__global__ void k32_update(const double* __restrict__ q, double* __restrict__ r, long long i0, long long i1,
                           const double* __restrict__ s_rz, const double* __restrict__ s_pq, double* partial, unsigned int* ticket,
                           RedOut<1> out) {
    const double alpha = *s_rz / *s_pq;
    double acc[1] = {0.0};
    for (long long i = i0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < i1; i += (long long)gridDim.x * blockDim.x) {
        double ri = fma(-alpha, q[i], r[i]);
        r[i] = ri;
        acc[0] = fma(ri, ri, acc[0]);
    }
    grid_reduce<1, false>(acc, partial, ticket, out);
}
// <r, s z32>
__global__ void k32_dot_rz(const double* __restrict__ r, const float* __restrict__ z, double s, long long i0, long long i1,
                           double* partial, unsigned int* ticket, RedOut<1> out) {
    double acc[1] = {0.0};
    for (long long i = i0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < i1; i += (long long)gridDim.x * blockDim.x)
        acc[0] = fma(r[i], s * (double)z[i], acc[0]);
    grid_reduce<1, false>(acc, partial, ticket, out);
}
// x += alpha p;  p = s z32 + beta p   (mode 0: both; 1: first call, p = s z32; 2: converged, only the deferred x update)
__global__ void k32_direction(const float* __restrict__ z, double s, double* __restrict__ p, double* __restrict__ x, long long i0,
                              long long i1, const double* __restrict__ s_rz_new, const double* __restrict__ s_rz_old,
                              const double* __restrict__ s_pq, int mode) {
    const double beta = mode == 0 ? *s_rz_new / *s_rz_old : 0.0, alpha = mode == 1 ? 0.0 : *s_rz_old / *s_pq;
    for (long long i = i0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < i1; i += (long long)gridDim.x * blockDim.x) {
        const double pi = p[i];
        if (mode != 1) x[i] = fma(alpha, pi, x[i]);
        if (mode != 2) {
            const double zi = s * (double)z[i];
            p[i] = mode == 1 ? zi : fma(beta, pi, zi);
        }
    }
}

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
