// unstructured.cu -- element-by-element gather-scatter operator with graph colouring, for imported
// hex8 / tet4 meshes (import_mesh -> Ferrite.Grid, MeshImport.jl:20).
//
// Replaces, for unstructured grids:
//   initialize_element_cache      FiniteElementAnalysis.jl:608-630  (unit matrices per cell, kept in HBM)
//   assemble_with_cache! + apply! FiniteElementAnalysis.jl:678-712  (y += E_e K_e u_e, colour by colour)
//   calculate_sensitivities!      SensitivityAnalysis.jl:31-80
//   FilterCache neighbour lists   FilterCommon.jl:61-98             (CSR weights on the device)
// Cells are greedily coloured so that no two cells of a colour share a node; one launch per colour
// scatters with plain stores (race free by construction, deterministic).  Per-cell data is stored
// colour-sorted and entry-major ([entry][cell]) so that a warp reads consecutive cells coalesced.
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <numeric>

#include "ctx.cuh"

template <int ND>
__global__ void __launch_bounds__(128) k_um_matvec(long long s0, long long s1, long long ncell, const int* __restrict__ order,
                                                  const int* __restrict__ celldofs, const double* __restrict__ ke,
                                                  const double* __restrict__ E, const uint8_t* __restrict__ mask,
                                                  const double* __restrict__ u, double* __restrict__ y) {
    long long s = s0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= s1) return;
    int dofs[ND];
    double ue[ND];
#pragma unroll
    for (int l = 0; l < ND; ++l) {
        dofs[l] = celldofs[(long long)l * ncell + s];
        ue[l] = mask[dofs[l]] ? 0.0 : u[dofs[l]];
    }
    const double Ee = E[order[s]];
#pragma unroll 1
    for (int r = 0; r < ND; ++r) {
        double acc = 0.0;
#pragma unroll
        for (int c = 0; c < ND; ++c) acc = fma(ke[((long long)r * ND + c) * ncell + s], ue[c], acc);
        int d = celldofs[(long long)r * ncell + s];
        y[d] += Ee * acc;  // no other cell of this colour touches dof d
    }
}

// The same operator for general hexahedra WITHOUT the stored element matrices: K_e u_e = sum over the 2x2x2 Gauss points of
// B' D B u_e detJ, rebuilt from the 8 node coordinates of the cell (FiniteElementAnalysis.jl:174-193 applied to u instead of
// assembled).  24 coordinates (192 B) per cell per apply instead of the 576 entries of K_e (4.6 KB): the streamed variant is
// bandwidth bound on exactly that array, this one is FP64 bound (~2.8 k flop per cell).  Reference gradients dN_a/dxi at the
// Gauss points come from constant memory.
__constant__ double c_um_dN[8][8][3];  // [gauss point][node][d/dxi_d] on the reference cube, VTK node order
__global__ void __launch_bounds__(128) k_um_matvec_hex_geom(long long s0, long long s1, long long ncell, const int* __restrict__ order,
                                                           const int* __restrict__ celldofs, const double* __restrict__ xcell,
                                                           const double* __restrict__ E, const uint8_t* __restrict__ mask, double lam0,
                                                           double mu0, const double* __restrict__ u, double* __restrict__ y) {
    long long s = s0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= s1) return;
    double x[8][3], ue[8][3], fe[8][3];
    int dofs[24];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int l = 3 * a + c;
            x[a][c] = xcell[(long long)l * ncell + s];
            dofs[l] = celldofs[(long long)l * ncell + s];
            ue[a][c] = mask[dofs[l]] ? 0.0 : u[dofs[l]];
            fe[a][c] = 0.0;
        }
#pragma unroll 1
    for (int q = 0; q < 8; ++q) {
        double J[3][3] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};
#pragma unroll
        for (int a = 0; a < 8; ++a)
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int d = 0; d < 3; ++d) J[c][d] = fma(x[a][c], c_um_dN[q][a][d], J[c][d]);
        const double det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
                           J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
        const double id = 1.0 / det;
        double Ji[3][3];
        Ji[0][0] = (J[1][1] * J[2][2] - J[1][2] * J[2][1]) * id;
        Ji[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * id;
        Ji[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * id;
        Ji[1][0] = (J[1][2] * J[2][0] - J[1][0] * J[2][2]) * id;
        Ji[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * id;
        Ji[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * id;
        Ji[2][0] = (J[1][0] * J[2][1] - J[1][1] * J[2][0]) * id;
        Ji[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * id;
        Ji[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * id;
        double g[8][3], H[3][3] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};  // g_a = grad N_a, H = grad u
#pragma unroll
        for (int a = 0; a < 8; ++a) {
#pragma unroll
            for (int c = 0; c < 3; ++c) g[a][c] = c_um_dN[q][a][0] * Ji[0][c] + c_um_dN[q][a][1] * Ji[1][c] + c_um_dN[q][a][2] * Ji[2][c];
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int d = 0; d < 3; ++d) H[c][d] = fma(ue[a][c], g[a][d], H[c][d]);
        }
        // sigma = lam tr(eps) I + 2 mu eps, times the quadrature weight |det J| (unit weights)
        const double w = fabs(det), tr = H[0][0] + H[1][1] + H[2][2];
        double S[3][3];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int d = 0; d < 3; ++d) S[c][d] = w * (mu0 * (H[c][d] + H[d][c]) + (c == d ? lam0 * tr : 0.0));
#pragma unroll
        for (int a = 0; a < 8; ++a)
#pragma unroll
            for (int c = 0; c < 3; ++c) fe[a][c] += S[c][0] * g[a][0] + S[c][1] * g[a][1] + S[c][2] * g[a][2];
    }
    const double Ee = E[order[s]];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int c = 0; c < 3; ++c) y[dofs[3 * a + c]] += Ee * fe[a][c];  // no other cell of this colour touches these dofs
}

template <int ND>
__global__ void k_um_diag(long long s0, long long s1, long long ncell, const int* __restrict__ order,
                          const int* __restrict__ celldofs, const double* __restrict__ ke, const double* __restrict__ E,
                          double* __restrict__ diag) {
    long long s = s0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= s1) return;
    const double Ee = E[order[s]];
    for (int r = 0; r < ND; ++r) diag[celldofs[(long long)r * ncell + s]] += Ee * ke[((long long)r * ND + r) * ncell + s];
}

__global__ void k_um_diag_finish(long long n, const uint8_t* __restrict__ mask, double m, double* __restrict__ dinv) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dinv[i] = mask[i] ? 1.0 / m : 1.0 / dinv[i];
}

template <int ND>
__global__ void k_um_sens(long long ncell, const int* __restrict__ order, const int* __restrict__ celldofs,
                          const double* __restrict__ ke, const double* __restrict__ u, const double* __restrict__ rho,
                          const double* __restrict__ E, double pexp, double dE, double* __restrict__ dc, double* partial,
                          unsigned int* ticket, RedOut<1> out) {
    double acc[1] = {0.0};
    for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < ncell; s += (long long)gridDim.x * blockDim.x) {
        double ue[ND];
#pragma unroll
        for (int l = 0; l < ND; ++l) ue[l] = u[celldofs[(long long)l * ncell + s]];
        double q = 0.0;
#pragma unroll 1
        for (int r = 0; r < ND; ++r) {
            double a = 0.0;
#pragma unroll
            for (int c = 0; c < ND; ++c) a = fma(ke[((long long)r * ND + c) * ncell + s], ue[c], a);
            q = fma(ue[r], a, q);
        }
        int e = order[s];
        double r = rho[e];
        double rpow = (pexp == 3.0) ? r * r : pow(r, pexp - 1.0);
        dc[e] = -(pexp * rpow * dE) * q;
        acc[0] += 0.5 * E[e] * q;
    }
    grid_reduce<1, false>(acc, partial, ticket, out);
}

void um_free(esimp_ctx* ctx) {
    UnstructuredMesh& um = ctx->um;
    cudaFree(um.d_order);
    cudaFree(um.d_celldofs);
    cudaFree(um.d_ke);
    cudaFree(um.d_vol);
    cudaFree(um.d_rowptr);
    cudaFree(um.d_col);
    cudaFree(um.d_w);
    cudaFree(um.d_D);
    cudaFree(um.d_nint);
    cudaFree(um.d_xcell);
    um = UnstructuredMesh();
}

// int N_a dOmega with the 2x2x2 rule of the cell values (the fe_body weights of FiniteElementAnalysis.jl:503-514)
static void hex_shape_integrals(const double* x, double* out8) {
    static const double HEXREF[8][3] = {{-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1},
                                        {-1, -1, 1},  {1, -1, 1},  {1, 1, 1},  {-1, 1, 1}};
    const double gp = 1.0 / sqrt(3.0);
    for (int a = 0; a < 8; ++a) out8[a] = 0.0;
    for (int q = 0; q < 8; ++q) {
        const double xi[3] = {(q & 1) ? gp : -gp, (q & 2) ? gp : -gp, (q & 4) ? gp : -gp};
        double J[3][3] = {{0}}, N[8];
        for (int a = 0; a < 8; ++a) {
            double f0 = 1 + HEXREF[a][0] * xi[0], f1 = 1 + HEXREF[a][1] * xi[1], f2 = 1 + HEXREF[a][2] * xi[2];
            double dN[3] = {0.125 * HEXREF[a][0] * f1 * f2, 0.125 * f0 * HEXREF[a][1] * f2, 0.125 * f0 * f1 * HEXREF[a][2]};
            N[a] = 0.125 * f0 * f1 * f2;
            for (int c = 0; c < 3; ++c)
                for (int d = 0; d < 3; ++d) J[c][d] += x[3 * a + c] * dN[d];
        }
        double det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
                     J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
        for (int a = 0; a < 8; ++a) out8[a] += N[a] * fabs(det);
    }
}

static double hex_volume(const double* x) {
    // order-3 Gauss rule on det J (calculate_element_volumes, FiniteElementAnalysis.jl:754-771)
    static const double HEXREF[8][3] = {{-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1},
                                        {-1, -1, 1},  {1, -1, 1},  {1, 1, 1},  {-1, 1, 1}};
    const double pts[3] = {-sqrt(0.6), 0.0, sqrt(0.6)}, wts[3] = {5.0 / 9.0, 8.0 / 9.0, 5.0 / 9.0};
    double vol = 0.0;
    for (int a3 = 0; a3 < 3; ++a3)
        for (int a2 = 0; a2 < 3; ++a2)
            for (int a1 = 0; a1 < 3; ++a1) {
                double xi[3] = {pts[a1], pts[a2], pts[a3]};
                double J[3][3] = {{0}};
                for (int a = 0; a < 8; ++a) {
                    double f0 = 1 + HEXREF[a][0] * xi[0], f1 = 1 + HEXREF[a][1] * xi[1], f2 = 1 + HEXREF[a][2] * xi[2];
                    double dN[3] = {0.125 * HEXREF[a][0] * f1 * f2, 0.125 * f0 * HEXREF[a][1] * f2, 0.125 * f0 * f1 * HEXREF[a][2]};
                    for (int c = 0; c < 3; ++c)
                        for (int d = 0; d < 3; ++d) J[c][d] += x[3 * a + c] * dN[d];
                }
                double det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
                             J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
                vol += det * wts[a1] * wts[a2] * wts[a3];
            }
    return vol;
}

int um_build(esimp_ctx* ctx, int celltype, int64_t n_cells, int64_t n_nodes, const double* xyz, const int64_t* conn,
             const int64_t* celldofs) {
    if (celltype != ESIMP_CELL_HEX8 && celltype != ESIMP_CELL_TET4) return esimp_fail(ctx, ESIMP_ERR_ARG, "celltype must be 8 or 4");
    um_free(ctx);
    UnstructuredMesh& um = ctx->um;
    const int npe = celltype, nd = 3 * npe;
    um.npe = npe;
    um.n_cells = n_cells;
    um.n_nodes = n_nodes;
    int64_t maxdof = -1;
    for (int64_t i = 0; i < n_cells * nd; ++i) maxdof = std::max(maxdof, celldofs[i]);
    um.n_dofs = maxdof + 1;
    // ---- greedy colouring over the node -> cell adjacency
    std::vector<int64_t> nptr(n_nodes + 1, 0);
    for (int64_t i = 0; i < n_cells * npe; ++i) nptr[conn[i] + 1]++;
    for (int64_t n = 0; n < n_nodes; ++n) nptr[n + 1] += nptr[n];
    std::vector<int32_t> ncell(nptr[n_nodes]);
    {
        std::vector<int64_t> fill(nptr.begin(), nptr.end() - 1);
        for (int64_t e = 0; e < n_cells; ++e)
            for (int a = 0; a < npe; ++a) ncell[fill[conn[e * npe + a]]++] = (int32_t)e;
    }
    std::vector<int> color(n_cells, -1);
    std::vector<int64_t> stamp(256, -1);
    int ncol = 0;
    for (int64_t e = 0; e < n_cells; ++e) {
        for (int a = 0; a < npe; ++a) {
            int64_t n = conn[e * npe + a];
            for (int64_t q = nptr[n]; q < nptr[n + 1]; ++q) {
                int c = color[ncell[q]];
                if (c >= 0) {
                    if (c >= (int)stamp.size()) stamp.resize(2 * c + 2, -1);
                    stamp[c] = e;
                }
            }
        }
        int c = 0;
        while (c < (int)stamp.size() && stamp[c] == e) ++c;
        if (c >= (int)stamp.size()) stamp.resize(2 * c + 2, -1);
        color[e] = c;
        ncol = std::max(ncol, c + 1);
    }
    um.n_colors = ncol;
    std::vector<int32_t> order(n_cells);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return color[a] < color[b]; });
    um.color_ptr.assign(ncol + 1, 0);
    for (int64_t e = 0; e < n_cells; ++e) um.color_ptr[color[e] + 1]++;
    for (int c = 0; c < ncol; ++c) um.color_ptr[c + 1] += um.color_ptr[c];
    // ---- per-cell data, colour-sorted, entry-major
    const double nu = ctx->prm.nu;
    const double lam = nu / ((1 + nu) * (1 - 2 * nu)), mu = 1.0 / (2 * (1 + nu));
    std::vector<int32_t> cd((size_t)nd * n_cells);
    std::vector<double> ke((size_t)nd * nd * n_cells), vol(n_cells), nint((size_t)npe * n_cells), xcell((size_t)nd * n_cells);
    std::vector<double> kel(nd * nd), x(3 * npe);
    for (int64_t s = 0; s < n_cells; ++s) {
        int64_t e = order[s];
        for (int a = 0; a < npe; ++a)
            for (int c = 0; c < 3; ++c) x[3 * a + c] = xyz[3 * conn[e * npe + a] + c];
        if (npe == 8)
            hex8_element_matrix(x.data(), lam, mu, kel.data());
        else
            tet4_element_matrix(x.data(), lam, mu, kel.data());
        for (int l = 0; l < nd; ++l) cd[(size_t)l * n_cells + s] = (int32_t)celldofs[e * nd + l];
        for (int q = 0; q < nd * nd; ++q) ke[(size_t)q * n_cells + s] = kel[q];
        for (int l = 0; l < nd; ++l) xcell[(size_t)l * n_cells + s] = x[l];
        if (npe == 8) {
            vol[e] = hex_volume(x.data());
            double ni[8];
            hex_shape_integrals(x.data(), ni);
            for (int a = 0; a < 8; ++a) nint[(size_t)a * n_cells + s] = ni[a];
        } else {
            double J[3][3];
            for (int c = 0; c < 3; ++c)
                for (int d = 0; d < 3; ++d) J[c][d] = x[3 * (d + 1) + c] - x[c];
            double det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
                         J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
            vol[e] = fabs(det) / 6.0;
            for (int a = 0; a < 4; ++a) nint[(size_t)a * n_cells + s] = vol[e] / 4.0;  // linear N_a: exact with the order-2 rule
        }
    }
    ESIMP_CUDA(cudaMalloc(&um.d_order, sizeof(int32_t) * n_cells));
    ESIMP_CUDA(cudaMalloc(&um.d_celldofs, sizeof(int32_t) * cd.size()));
    ESIMP_CUDA(cudaMalloc(&um.d_ke, sizeof(double) * ke.size()));
    ESIMP_CUDA(cudaMalloc(&um.d_vol, sizeof(double) * n_cells));
    ESIMP_CUDA(cudaMemcpy(um.d_order, order.data(), sizeof(int32_t) * n_cells, cudaMemcpyHostToDevice));
    ESIMP_CUDA(cudaMemcpy(um.d_celldofs, cd.data(), sizeof(int32_t) * cd.size(), cudaMemcpyHostToDevice));
    ESIMP_CUDA(cudaMemcpy(um.d_ke, ke.data(), sizeof(double) * ke.size(), cudaMemcpyHostToDevice));
    ESIMP_CUDA(cudaMemcpy(um.d_vol, vol.data(), sizeof(double) * n_cells, cudaMemcpyHostToDevice));
    ESIMP_CUDA(cudaMalloc(&um.d_nint, sizeof(double) * nint.size()));
    ESIMP_CUDA(cudaMalloc(&um.d_xcell, sizeof(double) * xcell.size()));
    ESIMP_CUDA(cudaMemcpy(um.d_nint, nint.data(), sizeof(double) * nint.size(), cudaMemcpyHostToDevice));
    ESIMP_CUDA(cudaMemcpy(um.d_xcell, xcell.data(), sizeof(double) * xcell.size(), cudaMemcpyHostToDevice));
    return ESIMP_OK;
}

static const esimp_ctx* g_um_dn_owner[64] = {};  // per device
static int um_upload_reference_gradients(esimp_ctx* ctx) {
    std::lock_guard<std::mutex> lock(g_const_mutex);
    if (g_um_dn_owner[ctx->device & 63] == ctx) return ESIMP_OK;
    static const double R[8][3] = {{-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1}, {-1, -1, 1}, {1, -1, 1}, {1, 1, 1}, {-1, 1, 1}};
    double dN[8][8][3];
    const double gp = 1.0 / sqrt(3.0);
    for (int q = 0; q < 8; ++q) {  // QuadratureRule{RefHexahedron}(2): first coordinate fastest
        const double xi[3] = {(q & 1) ? gp : -gp, (q & 2) ? gp : -gp, (q & 4) ? gp : -gp};
        for (int a = 0; a < 8; ++a) {
            const double f0 = 1 + R[a][0] * xi[0], f1 = 1 + R[a][1] * xi[1], f2 = 1 + R[a][2] * xi[2];
            dN[q][a][0] = 0.125 * R[a][0] * f1 * f2;
            dN[q][a][1] = 0.125 * f0 * R[a][1] * f2;
            dN[q][a][2] = 0.125 * f0 * f1 * R[a][2];
        }
    }
    ESIMP_CUDA(cudaMemcpyToSymbolAsync(c_um_dN, dN, sizeof(dN), 0, cudaMemcpyHostToDevice, ctx->stream));
    ESIMP_CUDA(cudaStreamSynchronize(ctx->stream));  // dN is a stack array
    g_um_dn_owner[ctx->device & 63] = ctx;
    return ESIMP_OK;
}
void um_forget_ctx(const esimp_ctx* ctx) {
    std::lock_guard<std::mutex> lock(g_const_mutex);
    if (g_um_dn_owner[ctx->device & 63] == ctx) g_um_dn_owner[ctx->device & 63] = nullptr;
}

template <int ND>
static int um_matvec_t(esimp_ctx* ctx, const double* d_in, double* d_out) {
    const UnstructuredMesh& um = ctx->um;
    // hexahedra: K_e u_e from the node coordinates (ESIMP_UM_STREAM_KE=1 keeps the stored-matrix kernel for comparison)
    static const bool stream_ke = getenv("ESIMP_UM_STREAM_KE") && atoi(getenv("ESIMP_UM_STREAM_KE")) == 1;
    const bool geom = ND == 24 && !stream_ke;
    if (geom) ESIMP_TRY(um_upload_reference_gradients(ctx));
    const double nu = ctx->prm.nu, lam0 = nu / ((1 + nu) * (1 - 2 * nu)), mu0 = 1.0 / (2 * (1 + nu));
    PROF_BEGIN(ctx, PROF_UM_MATVEC, true)
    ESIMP_CUDA(cudaMemsetAsync(d_out, 0, sizeof(double) * um.n_dofs, ctx->stream));
    for (int c = 0; c < um.n_colors; ++c) {
        long long s0 = um.color_ptr[c], s1 = um.color_ptr[c + 1];
        if (s1 == s0) continue;
        if (geom)
            k_um_matvec_hex_geom<<<(unsigned)((s1 - s0 + 127) / 128), 128, 0, ctx->stream>>>(s0, s1, um.n_cells, um.d_order, um.d_celldofs,
                                                                                             um.d_xcell, ctx->d_E, ctx->d_mask, lam0, mu0, d_in, d_out);
        else
            k_um_matvec<ND><<<(unsigned)((s1 - s0 + 127) / 128), 128, 0, ctx->stream>>>(s0, s1, um.n_cells, um.d_order, um.d_celldofs,
                                                                                        um.d_ke, ctx->d_E, ctx->d_mask, d_in, d_out);
        ctx->launches++;
    }
    PROF_END(ctx)
    ESIMP_CUDA(cudaGetLastError());
    return ESIMP_OK;
}
int um_matvec(esimp_ctx* ctx, const double* d_in, double* d_out) {
    return ctx->um.npe == 8 ? um_matvec_t<24>(ctx, d_in, d_out) : um_matvec_t<12>(ctx, d_in, d_out);
}

int um_diagonal(esimp_ctx* ctx) {
    const UnstructuredMesh& um = ctx->um;
    ESIMP_CUDA(cudaMemsetAsync(ctx->d_dinv, 0, sizeof(double) * um.n_dofs, ctx->stream));
    for (int c = 0; c < um.n_colors; ++c) {
        long long s0 = um.color_ptr[c], s1 = um.color_ptr[c + 1];
        if (s1 == s0) continue;
        unsigned nb = (unsigned)((s1 - s0 + 127) / 128);
        if (um.npe == 8)
            k_um_diag<24><<<nb, 128, 0, ctx->stream>>>(s0, s1, um.n_cells, um.d_order, um.d_celldofs, um.d_ke, ctx->d_E, ctx->d_dinv);
        else
            k_um_diag<12><<<nb, 128, 0, ctx->stream>>>(s0, s1, um.n_cells, um.d_order, um.d_celldofs, um.d_ke, ctx->d_E, ctx->d_dinv);
        ctx->launches++;
    }
    ESIMP_CUDA(cudaGetLastError());
    ESIMP_TRY(reduce_sum_scaled(ctx, ctx->d_dinv, nullptr, 1.0, 0, um.n_dofs, SC_DIAGSUM));
    ESIMP_TRY(fetch_scalars(ctx));
    ctx->diag_value = ctx->h_scal[SC_DIAGSUM] / (double)um.n_dofs;
    k_um_diag_finish<<<RED_BLOCKS, RED_THREADS, 0, ctx->stream>>>(um.n_dofs, ctx->d_mask, ctx->diag_value, ctx->d_dinv);
    ctx->launches++;
    ESIMP_CUDA(cudaGetLastError());
    return ESIMP_OK;
}

int um_sens_energy(esimp_ctx* ctx) {
    const UnstructuredMesh& um = ctx->um;
    RedOut<1> out{{ctx->d_scal + SC_ENERGY}};
    const double dE = ctx->prm.E0 - ctx->prm.Emin;
    if (um.npe == 8)
        k_um_sens<24><<<RED_BLOCKS, 128, 0, ctx->stream>>>(um.n_cells, um.d_order, um.d_celldofs, um.d_ke, ctx->d_u, ctx->d_phys,
                                                           ctx->d_E, ctx->prm.p, dE, ctx->d_dc, ctx->d_partial, ctx->d_ticket, out);
    else
        k_um_sens<12><<<RED_BLOCKS, 128, 0, ctx->stream>>>(um.n_cells, um.d_order, um.d_celldofs, um.d_ke, ctx->d_u, ctx->d_phys,
                                                           ctx->d_E, ctx->prm.p, dE, ctx->d_dc, ctx->d_partial, ctx->d_ticket, out);
    ctx->launches++;
    ESIMP_CUDA(cudaGetLastError());
    return ESIMP_OK;
}

// FilterCache (FilterCommon.jl:40-98) as CSR on the device: weights max(0, R - |c_j - c_i|) and the
// density-filter denominators D_i are computed once here instead of on every apply.
int um_filter_build(esimp_ctx* ctx, const int64_t* rowptr, const int64_t* col, const double* centers, const double* volumes,
                    double R) {
    UnstructuredMesh& um = ctx->um;
    const int64_t n = um.n_cells, nnz = rowptr[n];
    std::vector<double> w(nnz), D(n, 0.0);
    std::vector<int32_t> c32(nnz);
    for (int64_t i = 0; i < n; ++i)
        for (int64_t q = rowptr[i]; q < rowptr[i + 1]; ++q) {
            int64_t j = col[q];
            double dx = centers[3 * j] - centers[3 * i], dy = centers[3 * j + 1] - centers[3 * i + 1],
                   dz = centers[3 * j + 2] - centers[3 * i + 2];
            double wt = std::max(0.0, R - sqrt(dx * dx + dy * dy + dz * dz));
            w[q] = wt;
            c32[q] = (int32_t)j;
            if (wt > 0.0) D[i] += wt * volumes[j];
        }
    cudaFree(um.d_rowptr);
    cudaFree(um.d_col);
    cudaFree(um.d_w);
    cudaFree(um.d_D);
    um.nnz = nnz;
    ESIMP_CUDA(cudaMalloc(&um.d_rowptr, sizeof(int64_t) * (n + 1)));
    ESIMP_CUDA(cudaMalloc(&um.d_col, sizeof(int32_t) * std::max<int64_t>(nnz, 1)));
    ESIMP_CUDA(cudaMalloc(&um.d_w, sizeof(double) * std::max<int64_t>(nnz, 1)));
    ESIMP_CUDA(cudaMalloc(&um.d_D, sizeof(double) * n));
    ESIMP_CUDA(cudaMemcpy(um.d_rowptr, rowptr, sizeof(int64_t) * (n + 1), cudaMemcpyHostToDevice));
    ESIMP_CUDA(cudaMemcpy(um.d_col, c32.data(), sizeof(int32_t) * nnz, cudaMemcpyHostToDevice));
    ESIMP_CUDA(cudaMemcpy(um.d_w, w.data(), sizeof(double) * nnz, cudaMemcpyHostToDevice));
    ESIMP_CUDA(cudaMemcpy(um.d_D, D.data(), sizeof(double) * n, cudaMemcpyHostToDevice));
    ESIMP_CUDA(cudaMemcpy(um.d_vol, volumes, sizeof(double) * n, cudaMemcpyHostToDevice));
    ctx->filter_ready = true;
    return ESIMP_OK;
}
