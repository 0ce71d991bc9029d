// post_load.cu -- the density-dependent body force inside the loop and the post-solve stress fields.
//
// Replaces
//   apply_variable_density_volume_force!   FiniteElementAnalysis.jl:486-526, called on every assembly at
//                                          Optimization.jl:300-311 and :517-528 (f += rho_e * base * b * int N_a)
//   calculate_stresses_simp                FiniteElementAnalysis.jl:566-597 (per-cell Gauss-point stresses, the
//                                          von Mises value of the cell-averaged stress, its maximum and cell id)
//   calculate_von_mises_stresses           PostProcessing.jl:204-222 (von Mises of the FIRST Gauss point: the
//                                          exported "von_mises_stress" field)
//   calculate_element_energy               PostProcessing.jl:172-198 (1/2 sum_q sigma:eps dOmega)
// so that the export branches of the loop (Optimization.jl:407-477) read three cell arrays instead of
// downloading u and walking a Dict of tensors on the host.
//
// Structured grids: a node GATHERS the (at most 8) cells around it -- no atomics, deterministic -- and the
// physical shape-function gradients at the 8 Gauss points of the cuboid sit in constant memory.
// Imported meshes: colour-by-colour scatter for the load (like the EbE operator) and gradients rebuilt from
// the cell's node coordinates for the stresses.
#include <math.h>

#include <algorithm>

#include "ctx.cuh"
#include "stress_math.cuh"

__constant__ double c_gradN[8][8][3];  // [gauss point][node][d/dx_d], cuboid hx*hy*hz, VTK node order
static const esimp_ctx* g_grad_owner[64] = {};  // per device

void post_forget_ctx(const esimp_ctx* ctx) {
    std::lock_guard<std::mutex> lock(g_const_mutex);
    if (g_grad_owner[ctx->device & 63] == ctx) g_grad_owner[ctx->device & 63] = nullptr;
}

#define LAUNCH_STREAM(kern, ...)                                        \
    do {                                                                \
        kern<<<RED_BLOCKS, RED_THREADS, 0, ctx->stream>>>(__VA_ARGS__); \
        ctx->launches++;                                                \
        ESIMP_CUDA(cudaGetLastError());                                 \
    } while (0)

static int upload_gradients(esimp_ctx* ctx) {
    std::lock_guard<std::mutex> lock(g_const_mutex);
    if (g_grad_owner[ctx->device & 63] == ctx) return ESIMP_OK;
    double G[8][8][3];
    cuboid_gradient_table(ctx->hx, ctx->hy, ctx->hz, G);
    ESIMP_CUDA(cudaMemcpyToSymbol(c_gradN, G, sizeof(G)));
    g_grad_owner[ctx->device & 63] = ctx;
    return ESIMP_OK;
}

// ---------------------------------------------------------------------------- body force
// f(node, c) = fload(node, c) + b_c * V/8 * sum over the cells around the node of (rho_e * base >= 1e-6 ? rho_e * base : 0)
__global__ void k_body_force(GridDev g, int own_p0, int own_p1, const double* __restrict__ phys, double base, double bx, double by,
                             double bz, const double* __restrict__ fload, double* __restrict__ f) {
    const long long n = (long long)g.NPX * g.plane;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
        const int lp = (int)(idx / g.plane);
        const int rem = (int)(idx % g.plane);
        const int j = rem / g.nzp, k = rem % g.nzp;
        if (k > g.nz) continue;  // row pads
        double s = 0.0;
        if (lp >= own_p0 && lp < own_p1) {
            for (int dl = -1; dl <= 0; ++dl) {
                const int le = lp - g.e2p + dl, gi = g.ge0 + le;
                if (le < 0 || le >= g.NEX || gi < 0 || gi >= g.nx) continue;
                for (int dj = -1; dj <= 0; ++dj) {
                    const int jj = j + dj;
                    if (jj < 0 || jj >= g.ny) continue;
                    for (int dk = -1; dk <= 0; ++dk) {
                        const int kk = k + dk;
                        if (kk < 0 || kk >= g.nz) continue;
                        const double d = phys[g.cell(le, jj, kk)] * base;
                        if (!(d < 1e-6)) s += d;  // FiniteElementAnalysis.jl:495-497
                    }
                }
            }
        }
        const long long d0 = g.dof(lp, 0, j, k), d1 = g.dof(lp, 1, j, k), d2 = g.dof(lp, 2, j, k);
        f[d0] = (fload ? fload[d0] : 0.0) + bx * s;
        f[d1] = (fload ? fload[d1] : 0.0) + by * s;
        f[d2] = (fload ? fload[d2] : 0.0) + bz * s;
    }
}

template <int NPE>
__global__ void k_um_body(long long s0, long long s1, long long ncell, const int* __restrict__ order, const int* __restrict__ celldofs,
                          const double* __restrict__ nint, const double* __restrict__ dens, double base, double bx, double by,
                          double bz, double* __restrict__ f) {
    long long s = s0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= s1) return;
    const double d = dens[order[s]] * base;
    if (d < 1e-6) return;
#pragma unroll
    for (int a = 0; a < NPE; ++a) {
        const double w = nint[(long long)a * ncell + s] * d;
        f[celldofs[(long long)(3 * a + 0) * ncell + s]] += bx * w;  // no other cell of this colour touches these dofs
        f[celldofs[(long long)(3 * a + 1) * ncell + s]] += by * w;
        f[celldofs[(long long)(3 * a + 2) * ncell + s]] += bz * w;
    }
}

// d_out = (d_fload or 0) + body force of the densities d_dens * base
int body_force_into(esimp_ctx* ctx, const double* d_dens, double base, const double b[3], const double* d_fload, double* d_out) {
    if (ctx->structured) {
        const double w = ctx->hx * ctx->hy * ctx->hz / 8.0;  // int N_a dOmega on a cuboid
        LAUNCH_STREAM(k_body_force, ctx->g, ctx->own_p0, ctx->own_p1, d_dens, base, b[0] * w, b[1] * w, b[2] * w, d_fload, d_out);
        return ESIMP_OK;
    }
    const UnstructuredMesh& um = ctx->um;
    if (d_fload)
        ESIMP_CUDA(cudaMemcpyAsync(d_out, d_fload, sizeof(double) * um.n_dofs, cudaMemcpyDeviceToDevice, ctx->stream));
    else
        ESIMP_CUDA(cudaMemsetAsync(d_out, 0, sizeof(double) * um.n_dofs, ctx->stream));
    for (int c = 0; c < um.n_colors; ++c) {
        const long long s0 = um.color_ptr[c], s1 = um.color_ptr[c + 1];
        if (s1 == s0) continue;
        const unsigned nb = (unsigned)((s1 - s0 + 127) / 128);
        if (um.npe == 8)
            k_um_body<8><<<nb, 128, 0, ctx->stream>>>(s0, s1, um.n_cells, um.d_order, um.d_celldofs, um.d_nint, d_dens, base, b[0],
                                                      b[1], b[2], d_out);
        else
            k_um_body<4><<<nb, 128, 0, ctx->stream>>>(s0, s1, um.n_cells, um.d_order, um.d_celldofs, um.d_nint, d_dens, base, b[0],
                                                      b[1], b[2], d_out);
        ctx->launches++;
    }
    ESIMP_CUDA(cudaGetLastError());
    return ESIMP_OK;
}

// the loop's right-hand side for the current physical densities (Optimization.jl:300-314)
int body_force_refresh_rhs(esimp_ctx* ctx) {
    if (!ctx->body_on) return ESIMP_OK;
    return body_force_into(ctx, ctx->d_phys, ctx->body_base, ctx->body_b, ctx->d_fload, ctx->d_f);
}

// ---------------------------------------------------------------------------- stresses (arithmetic: stress_math.cuh)
__device__ __forceinline__ void load_cuboid_u(const GridDev& g, const double* __restrict__ u, int lp, int j, int k, double (&ue)[8][3]) {
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int c = 0; c < 3; ++c) ue[a][c] = u[g.dof(lp + hex_sx(a), c, j + hex_sy(a), k + hex_sz(a))];
}

// fields over the owned layers [le0, le1), device cell layout; max of vm_avg through the grid reduction
__global__ void __launch_bounds__(256) k_stress_fields(GridDev g, int le0, int le1, const double* __restrict__ u,
                                                       const double* __restrict__ E, double nu, double dvol, double* __restrict__ vm_q1,
                                                       double* __restrict__ vm_avg, double* __restrict__ energy, double* partial,
                                                       unsigned int* ticket, RedOut<1> out) {
    const long long n = (long long)(le1 - le0) * g.eplane;
    double acc[1] = {0.0};
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
        const int le = le0 + (int)(idx / g.eplane);
        const int rem = (int)(idx % g.eplane);
        const int j = rem / g.nz, k = rem % g.nz;
        double ue[8][3];
        load_cuboid_u(g, u, le + g.e2p, j, k, ue);
        const long long e = g.cell(le, j, k);
        double lam, mu;
        lame_of(E[e], nu, lam, mu);
        const CellStress r = cuboid_cell_stress(ue, c_gradN, lam, mu, dvol, nullptr);
        vm_q1[e] = r.vm_q1;
        vm_avg[e] = r.vm_avg;
        energy[e] = r.energy;
        acc[0] = fmax(acc[0], r.vm_avg);
    }
    grid_reduce<1, true>(acc, partial, ticket, out);
}

// smallest caller cell id whose vm_avg equals vmax (the reference keeps the FIRST maximum, :586): reduce max of -id
__global__ void __launch_bounds__(256) k_argmax_structured(GridDev g, int le0, int le1, const double* __restrict__ vm, const double* vmax,
                                                           double* partial, unsigned int* ticket, RedOut<1> out) {
    const long long n = (long long)(le1 - le0) * g.eplane;
    const double target = *vmax;
    double acc[1] = {-1.0e300};
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
        const int le = le0 + (int)(idx / g.eplane);
        const int rem = (int)(idx % g.eplane);
        const int j = rem / g.nz, k = rem % g.nz;
        if (vm[g.cell(le, j, k)] == target) {
            const long long id = (long long)(g.ge0 + le) + (long long)g.nx * (j + (long long)g.ny * k);
            acc[0] = fmax(acc[0], -(double)id);
        }
    }
    grid_reduce<1, true>(acc, partial, ticket, out);
}

__global__ void __launch_bounds__(256) k_argmax_flat(long long n, const double* __restrict__ vm, const double* vmax, double* partial,
                                                     unsigned int* ticket, RedOut<1> out) {
    const double target = *vmax;
    double acc[1] = {-1.0e300};
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x)
        if (vm[e] == target) acc[0] = fmax(acc[0], -(double)e);
    grid_reduce<1, true>(acc, partial, ticket, out);
}

template <int NPE>
__device__ __forceinline__ void load_um_cell(long long s, long long ncell, const int* __restrict__ celldofs, const double* __restrict__ xcell,
                                             const double* __restrict__ u, double (&x)[NPE][3], double (&ue)[NPE][3]) {
#pragma unroll
    for (int a = 0; a < NPE; ++a)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            x[a][c] = xcell[(long long)(3 * a + c) * ncell + s];
            ue[a][c] = u[celldofs[(long long)(3 * a + c) * ncell + s]];
        }
}

template <int NPE>
__global__ void __launch_bounds__(128) k_um_stress_fields(long long ncell, const int* __restrict__ order, const int* __restrict__ celldofs,
                                                          const double* __restrict__ xcell, const double* __restrict__ u,
                                                          const double* __restrict__ E, double nu, double* __restrict__ vm_q1,
                                                          double* __restrict__ vm_avg, double* __restrict__ energy, double* partial,
                                                          unsigned int* ticket, RedOut<1> out) {
    double acc[1] = {0.0};
    for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < ncell; s += (long long)gridDim.x * blockDim.x) {
        double x[NPE][3], ue[NPE][3];
        load_um_cell<NPE>(s, ncell, celldofs, xcell, u, x, ue);
        const int e = order[s];
        double lam, mu;
        lame_of(E[e], nu, lam, mu);
        CellStress r;
        if constexpr (NPE == 8)
            r = hex_cell_stress(x, ue, lam, mu, nullptr);
        else
            r = tet_cell_stress(x, ue, lam, mu, nullptr);
        vm_q1[e] = r.vm_q1;
        vm_avg[e] = r.vm_avg;
        energy[e] = r.energy;
        acc[0] = fmax(acc[0], r.vm_avg);
    }
    grid_reduce<1, true>(acc, partial, ticket, out);
}

// Gauss-point tensors of the caller's cells [id0, id0 + n): out[(id - id0)][q][6]; cells of other ranks stay 0
__global__ void __launch_bounds__(256) k_stress_tensors(GridDev g, int own_e0, int own_e1, long long id0, long long n,
                                                        const double* __restrict__ u, const double* __restrict__ E, double nu, double dvol,
                                                        double* __restrict__ out) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const long long id = id0 + t;
        const int i = (int)(id % g.nx);
        const long long r = id / g.nx;
        const int j = (int)(r % g.ny), k = (int)(r / g.ny);
        const int le = i - g.ge0;
        double* o = out + t * 48;
        if (le < own_e0 || le >= own_e1) {
            for (int q = 0; q < 48; ++q) o[q] = 0.0;
            continue;
        }
        double ue[8][3];
        load_cuboid_u(g, u, le + g.e2p, j, k, ue);
        double lam, mu;
        lame_of(E[g.cell(le, j, k)], nu, lam, mu);
        cuboid_cell_stress(ue, c_gradN, lam, mu, dvol, o);
    }
}

template <int NPE>
__global__ void __launch_bounds__(128) k_um_stress_tensors(long long ncell, long long id0, long long n, const int* __restrict__ order,
                                                           const int* __restrict__ celldofs, const double* __restrict__ xcell,
                                                           const double* __restrict__ u, const double* __restrict__ E, double nu,
                                                           double* __restrict__ out) {
    constexpr int NQ6 = NPE == 8 ? 48 : 6;
    for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < ncell; s += (long long)gridDim.x * blockDim.x) {
        const long long e = order[s];
        if (e < id0 || e >= id0 + n) continue;
        double x[NPE][3], ue[NPE][3];
        load_um_cell<NPE>(s, ncell, celldofs, xcell, u, x, ue);
        double lam, mu;
        lame_of(E[e], nu, lam, mu);
        if constexpr (NPE == 8)
            hex_cell_stress(x, ue, lam, mu, out + (e - id0) * NQ6);
        else
            tet_cell_stress(x, ue, lam, mu, out + (e - id0) * NQ6);
    }
}

// vm_q1 -> d_ctmp, vm_avg -> d_dcf, element energy -> d_rho_new (scratch of the loop between iterations);
// SC_TMP0 = max vm_avg, SC_TMP1 = -(first cell id attaining it)
int stress_fields(esimp_ctx* ctx, double* max_vm, int64_t* max_cell) {
    double* S = ctx->d_scal;
    RedOut<1> omax{{S + SC_TMP0}}, oarg{{S + SC_TMP1}};
    if (ctx->structured) {
        ESIMP_TRY(upload_gradients(ctx));
        if (ctx->nranks > 1) ESIMP_TRY(comm_halo_nodes(ctx, ctx->d_u));
        LAUNCH_STREAM(k_stress_fields, ctx->g, ctx->own_e0, ctx->own_e1, ctx->d_u, ctx->d_E, ctx->prm.nu,
                      ctx->hx * ctx->hy * ctx->hz / 8.0, ctx->d_ctmp, ctx->d_dcf, ctx->d_rho_new, ctx->d_partial, ctx->d_ticket, omax);
        ESIMP_TRY(allreduce_scalars(ctx, SC_TMP0, 1, true));
        LAUNCH_STREAM(k_argmax_structured, ctx->g, ctx->own_e0, ctx->own_e1, ctx->d_dcf, S + SC_TMP0, ctx->d_partial, ctx->d_ticket, oarg);
        ESIMP_TRY(allreduce_scalars(ctx, SC_TMP1, 1, true));
    } else {
        const UnstructuredMesh& um = ctx->um;
        if (um.npe == 8)
            k_um_stress_fields<8><<<RED_BLOCKS, 128, 0, ctx->stream>>>(um.n_cells, um.d_order, um.d_celldofs, um.d_xcell, ctx->d_u, ctx->d_E,
                                                                       ctx->prm.nu, ctx->d_ctmp, ctx->d_dcf, ctx->d_rho_new, ctx->d_partial,
                                                                       ctx->d_ticket, omax);
        else
            k_um_stress_fields<4><<<RED_BLOCKS, 128, 0, ctx->stream>>>(um.n_cells, um.d_order, um.d_celldofs, um.d_xcell, ctx->d_u, ctx->d_E,
                                                                       ctx->prm.nu, ctx->d_ctmp, ctx->d_dcf, ctx->d_rho_new, ctx->d_partial,
                                                                       ctx->d_ticket, omax);
        ctx->launches++;
        ESIMP_CUDA(cudaGetLastError());
        LAUNCH_STREAM(k_argmax_flat, (long long)um.n_cells, ctx->d_dcf, S + SC_TMP0, ctx->d_partial, ctx->d_ticket, oarg);
    }
    ESIMP_TRY(fetch_scalars(ctx));
    const double vmax = ctx->h_scal[SC_TMP0];
    if (max_vm) *max_vm = vmax;
    if (max_cell) *max_cell = vmax > 0.0 ? (int64_t)llround(-ctx->h_scal[SC_TMP1]) : -1;  // 0.0 start value, strict > (:569-589)
    return ESIMP_OK;
}

int stress_tensors(esimp_ctx* ctx, int64_t id0, int64_t n, double* d_out) {
    if (ctx->structured) {
        ESIMP_TRY(upload_gradients(ctx));
        if (ctx->nranks > 1) ESIMP_TRY(comm_halo_nodes(ctx, ctx->d_u));
        LAUNCH_STREAM(k_stress_tensors, ctx->g, ctx->own_e0, ctx->own_e1, (long long)id0, (long long)n, ctx->d_u, ctx->d_E, ctx->prm.nu,
                      ctx->hx * ctx->hy * ctx->hz / 8.0, d_out);
        return ESIMP_OK;
    }
    const UnstructuredMesh& um = ctx->um;
    if (um.npe == 8)
        k_um_stress_tensors<8><<<RED_BLOCKS, 128, 0, ctx->stream>>>(um.n_cells, (long long)id0, (long long)n, um.d_order, um.d_celldofs,
                                                                    um.d_xcell, ctx->d_u, ctx->d_E, ctx->prm.nu, d_out);
    else
        k_um_stress_tensors<4><<<RED_BLOCKS, 128, 0, ctx->stream>>>(um.n_cells, (long long)id0, (long long)n, um.d_order, um.d_celldofs,
                                                                    um.d_xcell, ctx->d_u, ctx->d_E, ctx->prm.nu, d_out);
    ctx->launches++;
    ESIMP_CUDA(cudaGetLastError());
    return ESIMP_OK;
}
