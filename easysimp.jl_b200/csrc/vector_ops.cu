// vector_ops.cu -- FP64 Jacobi-PCG and the streaming vector kernels around the operator
//
// Replaces `u .= cholesky(Symmetric(K, :L)) \ f` (Optimization.jl:317, :532) when the Jacobi preconditioner is
// selected (imported meshes always; structured grids when esimp_params.precond == ESIMP_PRECOND_JACOBI -- the
// multigrid-preconditioned variant lives in multigrid.cu).  All PCG scalars (alpha, beta, <p,q>, <r,z>, <r,r>) stay
// on the device; the host only reads ||r||^2 back every `pcg_check_every` iterations.  One iteration is
//   operator(p -> q) with <p,q> fused into its store | fix-up of prescribed rows |
//   update(r -= a q, <r,Dinv r>, <r,r>)              | direction(x += a p, p = Dinv r + b p)
// = 96 B/dof instead of the 120 B/dof of the textbook sequence; the two reductions of the update share one
// 2-double allreduce.  Every kernel is a 128-bit grid-stride stream over the rank's OWNED dof range (contiguous in
// the [plane][comp][y][z] layout); reductions are two-stage and deterministic (ctx.cuh:grid_reduce).
#include <math.h>

#include "ctx.cuh"

int esimp_fail(esimp_ctx* ctx, int code, const std::string& msg);

// ---------------------------------------------------------------------------- numbering maps
// device dof index -> caller dof index (structured grids)
// returns -1 for the pad entries of a node row (k > nz): they carry no dof
__device__ __forceinline__ long long caller_dof(const GridDev& g, const long long* __restrict__ dof_of_node, long long d) {
    int k = (int)(d % g.nzp);
    if (k > g.nz) return -1;
    long long t = d / g.nzp;
    int j = (int)(t % (g.ny + 1));
    t /= (g.ny + 1);
    int c = (int)(t % 3);
    int lp = (int)(t / 3);
    long long nref = (long long)(g.gp0 + lp) + (long long)(g.nx + 1) * (j + (long long)(g.ny + 1) * k);
    return dof_of_node ? dof_of_node[3 * nref + c] : 3 * nref + c;
}
__device__ __forceinline__ long long caller_cell(const GridDev& g, long long e) {
    int k = (int)(e % g.nz);
    long long t = e / g.nz;
    int j = (int)(t % g.ny);
    int le = (int)(t / g.ny);
    return (long long)(g.ge0 + le) + (long long)g.nx * (j + (long long)g.ny * k);
}

__global__ void k_gather_dofs(GridDev g, const long long* __restrict__ map, const double* __restrict__ src,
                              double* __restrict__ dst, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    {
        const long long s = caller_dof(g, map, i);
        dst[i] = s >= 0 ? src[s] : 0.0;
    }
}
__global__ void k_scatter_dofs(GridDev g, const long long* __restrict__ map, const double* __restrict__ src,
                               double* __restrict__ dst, long long i0, long long i1) {
    for (long long i = i0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < i1; i += (long long)gridDim.x * blockDim.x)
    {
        const long long s = caller_dof(g, map, i);
        if (s >= 0) dst[s] = src[i];
    }
}
__global__ void k_gather_cells(GridDev g, const double* __restrict__ src, double* __restrict__ dst, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dst[i] = src[caller_cell(g, i)];
}
__global__ void k_scatter_cells(GridDev g, const double* __restrict__ src, double* __restrict__ dst, long long i0, long long i1) {
    for (long long i = i0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < i1; i += (long long)gridDim.x * blockDim.x)
        dst[caller_cell(g, i)] = src[i];
}

#define LAUNCH_STREAM(kern, ...)                                            \
    do {                                                                    \
        kern<<<RED_BLOCKS, RED_THREADS, 0, ctx->stream>>>(__VA_ARGS__);     \
        ctx->launches++;                                                    \
        ESIMP_CUDA(cudaGetLastError());                                     \
    } while (0)

int vec_gather_to_device(esimp_ctx* ctx, const double* d_src_global, double* d_dst_local) {
    if (ctx->unstructured) {
        ESIMP_CUDA(cudaMemcpyAsync(d_dst_local, d_src_global, sizeof(double) * ctx->ndof_loc, cudaMemcpyDeviceToDevice, ctx->stream));
        return ESIMP_OK;
    }
    LAUNCH_STREAM(k_gather_dofs, ctx->g, (const long long*)ctx->d_dof_of_node, d_src_global, d_dst_local, ctx->ndof_loc);
    return ESIMP_OK;
}
int vec_scatter_from_device(esimp_ctx* ctx, const double* d_src_local, double* d_dst_global) {
    if (ctx->unstructured) {
        ESIMP_CUDA(cudaMemcpyAsync(d_dst_global, d_src_local, sizeof(double) * ctx->ndof_loc, cudaMemcpyDeviceToDevice, ctx->stream));
        return ESIMP_OK;
    }
    LAUNCH_STREAM(k_scatter_dofs, ctx->g, (const long long*)ctx->d_dof_of_node, d_src_local, d_dst_global, ctx->own_dof0, ctx->own_dof1);
    return ESIMP_OK;
}
int cell_gather_to_device(esimp_ctx* ctx, const double* d_src_global, double* d_dst_local) {
    if (ctx->unstructured) {
        ESIMP_CUDA(cudaMemcpyAsync(d_dst_local, d_src_global, sizeof(double) * ctx->ncell_loc, cudaMemcpyDeviceToDevice, ctx->stream));
        return ESIMP_OK;
    }
    LAUNCH_STREAM(k_gather_cells, ctx->g, d_src_global, d_dst_local, ctx->ncell_loc);
    return ESIMP_OK;
}
int cell_scatter_from_device(esimp_ctx* ctx, const double* d_src_local, double* d_dst_global) {
    if (ctx->unstructured) {
        ESIMP_CUDA(cudaMemcpyAsync(d_dst_global, d_src_local, sizeof(double) * ctx->ncell_loc, cudaMemcpyDeviceToDevice, ctx->stream));
        return ESIMP_OK;
    }
    LAUNCH_STREAM(k_scatter_cells, ctx->g, d_src_local, d_dst_global, ctx->own_cell0, ctx->own_cell1);
    return ESIMP_OK;
}

// ---------------------------------------------------------------------------- scalars
int fetch_scalars(esimp_ctx* ctx) {
    ESIMP_CUDA(cudaMemcpyAsync(ctx->h_scal, ctx->d_scal, sizeof(double) * N_SCAL, cudaMemcpyDeviceToHost, ctx->stream));
    ESIMP_CUDA(cudaStreamSynchronize(ctx->stream));
    return prof_collect(ctx);
}
// the stream is idle here: every recorded event pair has completed
int prof_collect(esimp_ctx* ctx) {
    for (int i = 0; i < ctx->prof_n; ++i) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ctx->prof_ev[2 * i], ctx->prof_ev[2 * i + 1]) == cudaSuccess) {
            const int cat = ctx->prof_tag[i] < PROF_NCAT ? ctx->prof_tag[i] : PROF_NCAT - 1;
            ctx->prof_ms_cat[cat] += ms;
            ctx->prof_cnt_cat[cat]++;
        }
    }
    ctx->prof_n = 0;
    if (ctx->prof_coarse_pending) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ctx->prof_coarse_ev[0], ctx->prof_coarse_ev[1]) == cudaSuccess) {
            ctx->prof_ms_cat[PROF_MG_COARSE] += ms;
            ctx->prof_cnt_cat[PROF_MG_COARSE]++;
        }
        ctx->prof_coarse_pending = false;
    }
    return ESIMP_OK;
}
int allreduce_scalars(esimp_ctx* ctx, int slot0, int n, bool is_max) {
    if (ctx->nranks <= 1) return ESIMP_OK;
    return comm_allreduce(ctx, ctx->d_scal + slot0, n, is_max);
}

// ---------------------------------------------------------------------------- reductions
__global__ void k_dot(const double* __restrict__ a, const double* __restrict__ b, long long i0, long long i1, double* partial,
                      unsigned int* ticket, RedOut<1> out) {
    double acc[1] = {0.0};
    for (long long i = i0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < i1; i += (long long)gridDim.x * blockDim.x)
        acc[0] = fma(a[i], b[i], acc[0]);
    grid_reduce<1, false>(acc, partial, ticket, out);
}
int vec_dot(esimp_ctx* ctx, const double* a, const double* b, int slot) {
    RedOut<1> out{{ctx->d_scal + slot}};
    LAUNCH_STREAM(k_dot, a, b, ctx->own_dof0, ctx->own_dof1, ctx->d_partial, ctx->d_ticket, out);
    return allreduce_scalars(ctx, slot, 1, false);
}

// sum_i scale * a[i] * (w ? w[i] : 1) over [i0, i1)
__global__ void k_sum_scaled(const double* __restrict__ a, const double* __restrict__ w, double scale, long long i0, long long i1,
                             double* partial, unsigned int* ticket, RedOut<1> out) {
    double acc[1] = {0.0};
    for (long long i = i0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < i1; i += (long long)gridDim.x * blockDim.x)
        acc[0] += w ? a[i] * w[i] : a[i];
    acc[0] *= scale;
    grid_reduce<1, false>(acc, partial, ticket, out);
}
int reduce_sum_scaled(esimp_ctx* ctx, const double* a, const double* w, double scale, int64_t i0, int64_t i1, int slot) {
    RedOut<1> out{{ctx->d_scal + slot}};
    LAUNCH_STREAM(k_sum_scaled, a, w, scale, (long long)i0, (long long)i1, ctx->d_partial, ctx->d_ticket, out);
    return allreduce_scalars(ctx, slot, 1, false);
}

// change = max|rho_new - rho_old|, gray = #{0.1 < rho_phys < 0.9}  (Optimization.jl:374-378)
__global__ void k_metrics_cells(const double* __restrict__ rn, const double* __restrict__ ro, const double* __restrict__ ph,
                                long long i0, long long i1, double* partial, unsigned int* ticket, RedOut<1> outmax,
                                RedOut<1> outsum, unsigned int* ticket2, double* partial2) {
    double mx[1] = {0.0}, cnt[1] = {0.0};
    for (long long i = i0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < i1; i += (long long)gridDim.x * blockDim.x) {
        mx[0] = fmax(mx[0], fabs(rn[i] - ro[i]));
        double p = ph[i];
        cnt[0] += (p > 0.1 && p < 0.9) ? 1.0 : 0.0;
    }
    grid_reduce<1, true>(mx, partial, ticket, outmax);
    grid_reduce<1, false>(cnt, partial2, ticket2, outsum);
}
__global__ void k_maxabs(const double* __restrict__ a, long long i0, long long i1, double* partial, unsigned int* ticket,
                         RedOut<1> out) {
    double mx[1] = {0.0};
    for (long long i = i0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < i1; i += (long long)gridDim.x * blockDim.x)
        mx[0] = fmax(mx[0], fabs(a[i]));
    grid_reduce<1, true>(mx, partial, ticket, out);
}
int reduce_metrics(esimp_ctx* ctx) {
    RedOut<1> omax{{ctx->d_scal + SC_CHANGE}}, osum{{ctx->d_scal + SC_GRAY}}, omu{{ctx->d_scal + SC_MAXU}};
    LAUNCH_STREAM(k_metrics_cells, ctx->d_rho, ctx->d_rho_old, ctx->d_phys, ctx->own_cell0, ctx->own_cell1, ctx->d_partial,
                  ctx->d_ticket, omax, osum, ctx->d_ticket + 1, ctx->d_partial + RED_BLOCKS);
    LAUNCH_STREAM(k_maxabs, ctx->d_u, ctx->own_dof0, ctx->own_dof1, ctx->d_partial, ctx->d_ticket, omu);
    if (ctx->nranks > 1) {
        // CHANGE and MAXU are max-reduced, GRAY summed; slots are laid out CHANGE, GRAY, MAXU
        ESIMP_TRY(comm_allreduce(ctx, ctx->d_scal + SC_CHANGE, 1, true));
        ESIMP_TRY(comm_allreduce(ctx, ctx->d_scal + SC_GRAY, 1, false));
        ESIMP_TRY(comm_allreduce(ctx, ctx->d_scal + SC_MAXU, 1, true));
    }
    return ESIMP_OK;
}

// ---------------------------------------------------------------------------- Dirichlet fix-up
// out[d] = m * in[d] on prescribed dofs: the (I-P) m term of apply!
__global__ void k_fixup(const long long* __restrict__ fixed, long long n, double m, const double* __restrict__ in,
                        double* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        long long d = fixed[i];
        out[d] = m * in[d];
    }
}
int fixup_lv(esimp_ctx* ctx, const long long* fixed, long long n_fixed, double m, const double* in, double* out) {
    if (n_fixed == 0) return ESIMP_OK;
    int blocks = (int)((n_fixed + 255) / 256);
    if (blocks > RED_BLOCKS) blocks = RED_BLOCKS;
    k_fixup<<<blocks, 256, 0, ctx->stream>>>(fixed, n_fixed, m, in, out);
    ctx->launches++;
    ESIMP_CUDA(cudaGetLastError());
    return ESIMP_OK;
}
int vec_fixup_dirichlet(esimp_ctx* ctx, const double* in, double* out) {
    return fixup_lv(ctx, (const long long*)ctx->d_fixed, ctx->n_fixed, ctx->diag_value, in, out);
}

// ---------------------------------------------------------------------------- PCG kernels
// r = f - q;  p = Dinv r;  rz = <r, Dinv r>;  rr = <r,r>;  ff = <f,f>
__global__ void k_pcg_init(const double* __restrict__ f, const double* __restrict__ q, const double* __restrict__ dinv,
                           double* __restrict__ r, double* __restrict__ p, long long i0, long long i1, double* partial,
                           unsigned int* ticket, RedOut<3> out) {
    double acc[3] = {0.0, 0.0, 0.0};
    for (long long i = i0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < i1; i += (long long)gridDim.x * blockDim.x) {
        double fi = f[i], ri = fi - q[i], zi = dinv[i] * ri;
        r[i] = ri;
        p[i] = zi;
        acc[0] = fma(ri, zi, acc[0]);
        acc[1] = fma(ri, ri, acc[1]);
        acc[2] = fma(fi, fi, acc[2]);
    }
    grid_reduce<3, false>(acc, partial, ticket, out);
}
// Streaming helper: [i0, i1) with 128-bit accesses.  All vectors share one index space and their base
// pointers are 256-byte aligned (cudaMalloc), so one alignment peel serves every array.
template <typename F2, typename F1>
__device__ __forceinline__ void stream_pairs(long long i0, long long i1, F2 f2, F1 f1) {
    long long a0 = (i0 + 1) & ~1LL, a1 = i1 & ~1LL;
    if (a0 > i1) a0 = i1;
    if (a1 < a0) a1 = a0;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
    if (tid == 0 && a0 > i0) f1(i0);
    if (tid == nth - 1 && i1 > a1) f1(a1);
    const long long v1 = a1 >> 1;
    long long v = (a0 >> 1) + tid;
    for (; v + nth < v1; v += 2 * nth) {  // two independent 16-byte transactions per array in flight
        f2(v);
        f2(v + nth);
    }
    if (v < v1) f2(v);
}
#define D2(ptr) reinterpret_cast<double2*>(ptr)
#define CD2(ptr) reinterpret_cast<const double2*>(ptr)

// alpha = rz/pq;  r -= alpha q;  rz_new = <r, Dinv r>;  rr = <r,r>          (24 B read + 8 B written per dof)
__global__ void __launch_bounds__(256, 4) k_pcg_update(const double* __restrict__ q, const double* __restrict__ dinv,
                                                    double* __restrict__ r, long long i0, long long i1,
                                                    const double* __restrict__ s_rz, const double* __restrict__ s_pq,
                                                    double* partial, unsigned int* ticket, RedOut<2> out) {
    const double alpha = *s_rz / *s_pq;
    double acc[2] = {0.0, 0.0};
    stream_pairs(
        i0, i1,
        [&](long long v) {
            double2 qi = CD2(q)[v], di = CD2(dinv)[v], ri = D2(r)[v];
            ri.x = fma(-alpha, qi.x, ri.x);
            ri.y = fma(-alpha, qi.y, ri.y);
            D2(r)[v] = ri;
            acc[0] = fma(ri.x * di.x, ri.x, acc[0]);
            acc[0] = fma(ri.y * di.y, ri.y, acc[0]);
            acc[1] = fma(ri.x, ri.x, acc[1]);
            acc[1] = fma(ri.y, ri.y, acc[1]);
        },
        [&](long long i) {
            double ri = fma(-alpha, q[i], r[i]);
            r[i] = ri;
            acc[0] = fma(ri * dinv[i], ri, acc[0]);
            acc[1] = fma(ri, ri, acc[1]);
        });
    grid_reduce<2, false>(acc, partial, ticket, out);
}
// x += alpha p (deferred from the update: p is read here anyway);  beta = rz_new/rz_old;  p = Dinv r + beta p
//                                                                      (32 B read + 16 B written per dof)
__global__ void __launch_bounds__(256, 4) k_pcg_direction(const double* __restrict__ r, const double* __restrict__ dinv,
                                                       double* __restrict__ p, double* __restrict__ x, long long i0, long long i1,
                                                       const double* __restrict__ s_rz_new, const double* __restrict__ s_rz_old,
                                                       const double* __restrict__ s_pq) {
    const double beta = *s_rz_new / *s_rz_old, alpha = *s_rz_old / *s_pq;
    stream_pairs(
        i0, i1,
        [&](long long v) {
            double2 pi = D2(p)[v], xi = D2(x)[v], ri = CD2(r)[v], di = CD2(dinv)[v];
            xi.x = fma(alpha, pi.x, xi.x);
            xi.y = fma(alpha, pi.y, xi.y);
            pi.x = fma(beta, pi.x, di.x * ri.x);
            pi.y = fma(beta, pi.y, di.y * ri.y);
            D2(x)[v] = xi;
            D2(p)[v] = pi;
        },
        [&](long long i) {
            const double pi = p[i];
            x[i] = fma(alpha, pi, x[i]);
            p[i] = fma(beta, pi, dinv[i] * r[i]);
        });
}

// y = P K P x for an input that is already zero on prescribed dofs (halo exchange + operator kernel)
static int op_apply_raw(esimp_ctx* ctx, double* d_in, double* d_out, int dot_slot = -1) {
    if (ctx->structured) {
        if (ctx->nranks > 1) ESIMP_TRY(comm_halo_nodes(ctx, d_in));
        return hex8_matvec(ctx, d_in, d_out, ctx->prm.matvec_variant, dot_slot);
    }
    return um_matvec(ctx, d_in, d_out);
}
int op_apply(esimp_ctx* ctx, double* d_in, double* d_out);
// q = A p and <p,q> in d_scal[slot]: fused into the operator's store where the kernel supports it
static int op_apply_dot(esimp_ctx* ctx, double* d_in, double* d_out, int slot) {
    if (hex8_matvec_fuses_dot(ctx)) {
        ESIMP_TRY(op_apply_raw(ctx, d_in, d_out, slot));
        ESIMP_TRY(vec_fixup_dirichlet(ctx, d_in, d_out));  // in[d] = 0 there, so <in,out> is unaffected
        return allreduce_scalars(ctx, slot, 1, false);
    }
    ESIMP_TRY(op_apply(ctx, d_in, d_out));
    return vec_dot(ctx, d_in, d_out, slot);
}
// A x = P K P x + m (I-P) x for x with x[d] = 0 on prescribed dofs (every PCG vector is)
int op_apply(esimp_ctx* ctx, double* d_in, double* d_out) {
    ESIMP_TRY(op_apply_raw(ctx, d_in, d_out));
    return vec_fixup_dirichlet(ctx, d_in, d_out);
}
// same for an arbitrary x: the masked copy goes through d_tmp, the m (I-P) x term uses the original
int op_apply_general(esimp_ctx* ctx, double* d_in, double* d_tmp, double* d_out) {
    ESIMP_CUDA(cudaMemcpyAsync(d_tmp, d_in, sizeof(double) * ctx->ndof_loc, cudaMemcpyDeviceToDevice, ctx->stream));
    const double m = ctx->diag_value;
    ctx->diag_value = 0.0;
    int s = vec_fixup_dirichlet(ctx, d_tmp, d_tmp);  // tmp[d] = 0
    ctx->diag_value = m;
    ESIMP_TRY(s);
    ESIMP_TRY(op_apply_raw(ctx, d_tmp, d_out));
    return vec_fixup_dirichlet(ctx, d_in, d_out);
}

// Jacobi-PCG on A = P K P + m (I-P), b = P f.  Starts from ctx->d_u (zero or warm start); the
// prescribed entries of u, r, p stay exactly 0 because f and the initial u are masked.
int pcg_solve(esimp_ctx* ctx, int* iters, double* relres) {
    const long long i0 = ctx->own_dof0, i1 = ctx->own_dof1;
    double* S = ctx->d_scal;
    static const int RZ[2] = {SC_RZ0, SC_RZ1}, RR[2] = {SC_RR0, SC_RR1};
    // r = f - A u
    ESIMP_TRY(op_apply(ctx, ctx->d_u, ctx->d_q));
    {
        RedOut<3> out{{S + RZ[0], S + RR[0], S + SC_FF}};
        LAUNCH_STREAM(k_pcg_init, ctx->d_f, ctx->d_q, ctx->d_dinv, ctx->d_r, ctx->d_p, i0, i1, ctx->d_partial, ctx->d_ticket, out);
        if (ctx->nranks > 1) {
            ESIMP_TRY(comm_allreduce(ctx, S + RZ[0], 2, false));  // <r,z>, <r,r> adjacent
            ESIMP_TRY(comm_allreduce(ctx, S + SC_FF, 1, false));
        }
    }
    ESIMP_TRY(fetch_scalars(ctx));
    const double ff = ctx->h_scal[SC_FF];
    ctx->f_norm2 = ff;
    const double tol2 = ctx->prm.pcg_rtol * ctx->prm.pcg_rtol * ff;
    double rr = ctx->h_scal[RR[0]];
    int it = 0;
    const int every = ctx->prm.pcg_check_every > 0 ? ctx->prm.pcg_check_every : 1;
    int status = ESIMP_OK;
    if (ff == 0.0) {
        ESIMP_CUDA(cudaMemsetAsync(ctx->d_u, 0, sizeof(double) * ctx->ndof_loc, ctx->stream));
        rr = 0.0;
    }
    while (rr > tol2 && it < ctx->prm.pcg_maxit) {
        const int cur = it & 1, nxt = cur ^ 1;  // pair holding <r,z>, <r,r> of the current / next residual
        ESIMP_TRY(op_apply_dot(ctx, ctx->d_p, ctx->d_q, SC_PQ));
        {
            RedOut<2> out{{S + RZ[nxt], S + RR[nxt]}};
            LAUNCH_STREAM(k_pcg_update, ctx->d_q, ctx->d_dinv, ctx->d_r, i0, i1, S + RZ[cur], S + SC_PQ, ctx->d_partial, ctx->d_ticket, out);
            if (ctx->nranks > 1) ESIMP_TRY(comm_allreduce(ctx, S + RZ[nxt], 2, false));
        }
        LAUNCH_STREAM(k_pcg_direction, ctx->d_r, ctx->d_dinv, ctx->d_p, ctx->d_u, i0, i1, S + RZ[nxt], S + RZ[cur], S + SC_PQ);
        ++it;
        if (it % every == 0 || it >= ctx->prm.pcg_maxit) {
            ESIMP_TRY(fetch_scalars(ctx));
            rr = ctx->h_scal[RR[nxt]];
            if (!(rr == rr)) return esimp_fail(ctx, ESIMP_ERR_STATE, "PCG breakdown: residual is NaN");
        }
    }
    if (rr > tol2) status = ESIMP_WARN_PCG_MAXIT;
    if (iters) *iters = it;
    if (relres) *relres = ff > 0 ? sqrt(rr / ff) : 0.0;
    return status;
}
