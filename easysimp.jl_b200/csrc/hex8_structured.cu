// hex8_structured.cu -- matrix-free hex8 operator on structured grids (sm_100a)
//
// Replaces, for grids made by generate_grid(Hexahedron, ...):
//   * assemble_with_cache! (FiniteElementAnalysis.jl:678-712): K = sum_e E(rho_e)/E0 * K_unit
//   * Ferrite apply!(K,f,ch) (Optimization.jl:581-583): zero rows/cols of prescribed dofs
//   * the K*u products inside the solve (Optimization.jl:317) and the energy (:320)
//   * calculate_sensitivities! (SensitivityAnalysis.jl:31-80): u_e' K0 u_e per cell
// by y = sum_e E_e G_e' K0 G_e u applied without ever forming K.
//
// The fused kernel works in the trilinear *mode basis*: per displacement component the 8 nodal
// values of a cell are transformed by a 2x2x2 Walsh-Hadamard butterfly into the coefficients of
// {1, xi, eta, zeta, xi.eta, eta.zeta, xi.zeta, xi.eta.zeta}.  For a cuboid the reflection
// symmetry Z2^3 makes K~ = T'K0T/64 block diagonal: 8 parity classes, one (component, mode)
// pair per component in each class, 45 non-zeros instead of 576 (pattern below, checked on the
// host for every (hx,hy,hz,nu)).  A thread owns one (y,z) column of cells and marches along x:
// the 2-D transform of a node plane is shared by the two cells either side of it, and the
// x-direction butterfly of the output is shared the same way, so a cell costs 162 FP64 issue
// slots instead of 600 and each u value is read from HBM once.
#include <cuda.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <type_traits>

#include <atomic>

#include "ctx.cuh"

// ---------------------------------------------------------------------------- constants
__constant__ double c_K0[24 * 24];  // dense unit-modulus element matrix, local dof 3a+c
__constant__ double c_B[8][3][3];   // mode-space blocks per parity class
__constant__ float c_Bf[8][3][3];   // the same in single precision (FP32 multigrid cycle, multigrid32.cu)
__constant__ double c_kdiag[3];     // K0[3a+c,3a+c] (same for every node a of a cuboid)
static const esimp_ctx* g_const_owner[64] = {};  // per device
static const esimp_ctx* g_const_owner_f32[64] = {};  // per device  // owner of c_Bf (uploaded only when the FP32 cycle is used)

#include "fused_kernel.cuh"
#include "sweep_kernel.cuh"
#include "sweep1_kernel.cuh"


static const double HEXREF[8][3] = {{-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1},
                                    {-1, -1, 1},  {1, -1, 1},  {1, 1, 1},  {-1, 1, 1}};

// Element matrix from shape-function gradients:
// ke[(a,c),(b,d)] += dV (lam g_a[c] g_b[d] + mu (delta_cd g_a.g_b + g_a[d] g_b[c]))
// == (sym grad N_i) : (lam tr(eps_j) I + 2 mu eps_j), FiniteElementAnalysis.jl:181-190.
static void add_gradients(int n, const double (*g)[3], double lam, double mu, double dV, double* ke) {
    const int nd = 3 * n;
    for (int a = 0; a < n; ++a)
        for (int b = 0; b < n; ++b) {
            double gg = g[a][0] * g[b][0] + g[a][1] * g[b][1] + g[a][2] * g[b][2];
            for (int c = 0; c < 3; ++c)
                for (int d = 0; d < 3; ++d) {
                    double v = lam * g[a][c] * g[b][d] + mu * g[a][d] * g[b][c] + (c == d ? mu * gg : 0.0);
                    ke[(3 * a + c) * nd + 3 * b + d] += dV * v;
                }
        }
}

static double inv3(const double J[3][3], double Ji[3][3]) {
    double det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
                 J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
    double id = 1.0 / det;
    Ji[0][0] = (J[1][1] * J[2][2] - J[1][2] * J[2][1]) * id;
    Ji[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * id;
    Ji[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * id;
    Ji[1][0] = (J[1][2] * J[2][0] - J[1][0] * J[2][2]) * id;
    Ji[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * id;
    Ji[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * id;
    Ji[2][0] = (J[1][0] * J[2][1] - J[1][1] * J[2][0]) * id;
    Ji[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * id;
    Ji[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * id;
    return det;
}

// 8-node hexahedron, 2x2x2 Gauss (QuadratureRule{RefHexahedron}(2), FiniteElementAnalysis.jl:138)
void hex8_element_matrix(const double* xyz, double lam, double mu, double* ke) {
    memset(ke, 0, sizeof(double) * 24 * 24);
    const double gp = 1.0 / sqrt(3.0);
    for (int qz = 0; qz < 2; ++qz)
        for (int qy = 0; qy < 2; ++qy)
            for (int qx = 0; qx < 2; ++qx) {
                double xi[3] = {qx ? gp : -gp, qy ? gp : -gp, qz ? gp : -gp};
                double dN[8][3];
                for (int a = 0; a < 8; ++a) {
                    double f0 = 1 + HEXREF[a][0] * xi[0], f1 = 1 + HEXREF[a][1] * xi[1], f2 = 1 + HEXREF[a][2] * xi[2];
                    dN[a][0] = 0.125 * HEXREF[a][0] * f1 * f2;
                    dN[a][1] = 0.125 * f0 * HEXREF[a][1] * f2;
                    dN[a][2] = 0.125 * f0 * f1 * HEXREF[a][2];
                }
                double J[3][3] = {{0}}, Ji[3][3];
                for (int a = 0; a < 8; ++a)
                    for (int c = 0; c < 3; ++c)
                        for (int d = 0; d < 3; ++d) J[c][d] += xyz[3 * a + c] * dN[a][d];
                double det = inv3(J, Ji);
                double g[8][3];
                for (int a = 0; a < 8; ++a)
                    for (int c = 0; c < 3; ++c) g[a][c] = dN[a][0] * Ji[0][c] + dN[a][1] * Ji[1][c] + dN[a][2] * Ji[2][c];
                add_gradients(8, g, lam, mu, fabs(det), ke);
            }
}

// 4-node tetrahedron: constant gradients, K = V B'DB (weights of any rule sum to 1/6)
void tet4_element_matrix(const double* xyz, double lam, double mu, double* ke) {
    memset(ke, 0, sizeof(double) * 12 * 12);
    const double dN[4][3] = {{-1, -1, -1}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    double J[3][3] = {{0}}, Ji[3][3];
    for (int a = 0; a < 4; ++a)
        for (int c = 0; c < 3; ++c)
            for (int d = 0; d < 3; ++d) J[c][d] += xyz[3 * a + c] * dN[a][d];
    double det = inv3(J, Ji);
    double g[4][3];
    for (int a = 0; a < 4; ++a)
        for (int c = 0; c < 3; ++c) g[a][c] = dN[a][0] * Ji[0][c] + dN[a][1] * Ji[1][c] + dN[a][2] * Ji[2][c];
    add_gradients(4, g, lam, mu, fabs(det) / 6.0, ke);
}

// a destroyed context must not be mistaken for the owner of the constants by a later one at the same address
void hex8_forget_ctx(const esimp_ctx* ctx) {
    std::lock_guard<std::mutex> lock(g_const_mutex);
    if (g_const_owner[ctx->device & 63] == ctx) g_const_owner[ctx->device & 63] = nullptr;
    if (g_const_owner_f32[ctx->device & 63] == ctx) g_const_owner_f32[ctx->device & 63] = nullptr;
}

// Build K0 (unit modulus), its mode-space blocks and the nodal diagonal for the context's cuboid.
int hex8_upload_constants(esimp_ctx* ctx) {
    std::lock_guard<std::mutex> lock(g_const_mutex);
    if (!ctx->k0_valid) {
        const double nu = ctx->prm.nu;
        const double lam = nu / ((1 + nu) * (1 - 2 * nu)), mu = 1.0 / (2 * (1 + nu));  // SensitivityAnalysis.jl:44-45
        double xyz[24];
        for (int a = 0; a < 8; ++a) {
            xyz[3 * a + 0] = 0.5 * (1 + HEXREF[a][0]) * ctx->hx;
            xyz[3 * a + 1] = 0.5 * (1 + HEXREF[a][1]) * ctx->hy;
            xyz[3 * a + 2] = 0.5 * (1 + HEXREF[a][2]) * ctx->hz;
        }
        hex8_element_matrix(xyz, lam, mu, ctx->K0);
        // K~ = T' K0 T / 64 with T[3a+c][3m+c] = sx^bx sy^by sz^bz, m = bx + 2 by + 4 bz
        double H[8][8];
        for (int a = 0; a < 8; ++a)
            for (int m = 0; m < 8; ++m)
                H[a][m] = ((m & 1) ? HEXREF[a][0] : 1.0) * ((m & 2) ? HEXREF[a][1] : 1.0) * ((m & 4) ? HEXREF[a][2] : 1.0);
        double Kt[24][24];
        double kmax = 0;
        for (int m1 = 0; m1 < 8; ++m1)
            for (int c = 0; c < 3; ++c)
                for (int m2 = 0; m2 < 8; ++m2)
                    for (int d = 0; d < 3; ++d) {
                        double s = 0;
                        for (int a = 0; a < 8; ++a)
                            for (int b = 0; b < 8; ++b) s += H[a][m1] * ctx->K0[(3 * a + c) * 24 + 3 * b + d] * H[b][m2];
                        Kt[3 * m1 + c][3 * m2 + d] = s / 64.0;
                        kmax = fmax(kmax, fabs(s / 64.0));
                    }
        for (int i = 0; i < 24; ++i)
            for (int j = 0; j < 24; ++j) {
                int m1 = i / 3, c = i % 3, m2 = j / 3, d = j % 3;
                int q1 = m1 ^ (1 << c), q2 = m2 ^ (1 << d);
                bool structural = (q1 == q2) && pat(q1, c, d);
                if (!structural && fabs(Kt[i][j]) > 1e-12 * kmax)
                    return esimp_fail(ctx, ESIMP_ERR_STATE, "mode-space K0 violates the cuboid parity pattern");
            }
        for (int q = 0; q < 8; ++q)
            for (int c = 0; c < 3; ++c)
                for (int d = 0; d < 3; ++d) ctx->B[q][c][d] = pat(q, c, d) ? Kt[3 * (q ^ (1 << c)) + c][3 * (q ^ (1 << d)) + d] : 0.0;
        for (int c = 0; c < 3; ++c) ctx->kdiag[c] = ctx->K0[c * 24 + c];
        ctx->k0_valid = true;
        g_const_owner[ctx->device & 63] = nullptr;
    }
    if (g_const_owner[ctx->device & 63] != ctx) {
        ESIMP_CUDA(cudaMemcpyToSymbolAsync(c_K0, ctx->K0, sizeof(ctx->K0), 0, cudaMemcpyHostToDevice, ctx->stream));
        ESIMP_CUDA(cudaMemcpyToSymbolAsync(c_B, ctx->B, sizeof(ctx->B), 0, cudaMemcpyHostToDevice, ctx->stream));
        ESIMP_CUDA(cudaMemcpyToSymbolAsync(c_kdiag, ctx->kdiag, sizeof(ctx->kdiag), 0, cudaMemcpyHostToDevice, ctx->stream));
        g_const_owner[ctx->device & 63] = ctx;
    }
    return ESIMP_OK;
}

// ---------------------------------------------------------------------------- V1: dense + atomics
// One thread per cell, dense 24x24 from constant memory, atomicAdd into a zeroed y.
// Debug / cross-check variant (ESIMP_MATVEC_DENSE_ATOMIC); ~3.5x the FP64 work of the fused kernel.
__global__ void __launch_bounds__(128) k_matvec_dense_atomic(GridDev g, int le0, int le1, const double* __restrict__ E,
                                                           const uint8_t* __restrict__ mask,
                                                           const double* __restrict__ u, double* __restrict__ y) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long n = (long long)(le1 - le0) * g.eplane;
    if (idx >= n) return;
    int le = le0 + (int)(idx / g.eplane);
    int rem = (int)(idx % g.eplane);
    int j = rem / g.nz, k = rem % g.nz;
    const int sx[8] = {0, 1, 1, 0, 0, 1, 1, 0}, sy[8] = {0, 0, 1, 1, 0, 0, 1, 1}, sz[8] = {0, 0, 0, 0, 1, 1, 1, 1};
    double ue[24];
    long long di[8];
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        int lp = le + g.e2p + sx[a];
        di[a] = g.dof(lp, 0, j + sy[a], k + sz[a]);
        uint8_t m = mask[g.node(lp, j + sy[a], k + sz[a])];
#pragma unroll
        for (int c = 0; c < 3; ++c) ue[3 * a + c] = ((m >> c) & 1) ? 0.0 : u[di[a] + c * g.plane];
    }
    double Ee = E[g.cell(le, j, k)];
#pragma unroll 1
    for (int r = 0; r < 24; ++r) {
        double acc = 0.0;
#pragma unroll
        for (int cidx = 0; cidx < 24; ++cidx) acc = fma(c_K0[r * 24 + cidx], ue[cidx], acc);
        atomicAdd(&y[di[r / 3] + (r % 3) * g.plane], Ee * acc);
    }
}

// zero the tile-perimeter rows / columns of y that the ATOMIC variant accumulates into
template <int TY, int TZ>
__global__ void k_zero_perimeter(GridDev g, int ps0, int ps1, double* __restrict__ y) {
    long long n = (long long)(ps1 - ps0) * 3 * g.plane;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        int k = (int)(i % g.nzp);
        int j = (int)((i / g.nzp) % (g.ny + 1));
        if ((j % TY) == 0 || (k % TZ) == 0) y[(long long)ps0 * 3 * g.plane + i] = 0.0;
    }
}

// vw supplies the geometry; E_t is the modulus array in the kernel's arithmetic type (vw.E for T = double)
template <typename T, int TY, int TZ, bool ATOMIC, int MINB, int MODE>
static int launch_fused_tt(esimp_ctx* ctx, const OpView& vw, const T* E_t, bool fine, const T* d_in, T* d_out, int dot_slot,
                           const FuseArgsT<T>& fa) {
    constexpr bool DOT = MODE == FUSE_DOT || MODE == FUSE_SMOOTH_DOT;
    const GridDev& g = vw.g;
    constexpr int SY = ATOMIC ? TY : TY - 1, SZ = ATOMIC ? TZ : TZ - 1;
    const int nty = ATOMIC ? (g.ny + SY - 1) / SY : (g.ny + 1 + SY - 1) / SY;
    const int ntz = ATOMIC ? (g.nz + SZ - 1) / SZ : (g.nz + 1 + SZ - 1) / SZ;
    const int nplanes = vw.own_p1 - vw.own_p0;
    // x chunks: enough blocks for >= ~6 waves of 148 SMs x MINB CTAs, chunk >= 8 planes
    int chunk = nplanes;
    const long long tiles = (long long)nty * ntz;
    while (chunk > 8 && tiles * ((nplanes + chunk - 1) / chunk) < 148LL * MINB * 6) chunk = (chunk + 1) / 2;
    if (ctx->fused_chunk > 0) chunk = ctx->fused_chunk;
    const int nchunk = (nplanes + chunk - 1) / chunk;
    const size_t smem = sizeof(FusedSmem<T, TY, TZ, (MODE == FUSE_SMOOTH || MODE == FUSE_SMOOTH_DOT) ? 3 : (MODE == FUSE_RESID ? 1 : 0)>);
    static std::atomic<unsigned long long> attr_devices{0};  // the attribute is per device: one bit per device this process has used
    auto kern = k_matvec_fused<T, TY, TZ, ATOMIC, MINB, MODE>;
    const unsigned long long dev_bit = 1ull << (ctx->device & 63);
    if (!(attr_devices.load() & dev_bit)) {
        ESIMP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_devices.fetch_or(dev_bit);
    }
    if constexpr (ATOMIC && std::is_same<T, double>::value) {
        k_zero_perimeter<TY, TZ><<<RED_BLOCKS, RED_THREADS, 0, ctx->stream>>>(g, vw.own_p0, vw.own_p1, d_out);
        ctx->launches++;
    }
    dim3 grid(ntz, nty, nchunk), block(TY * TZ);
    const long long nblk = (long long)ntz * nty * nchunk;
    if (DOT && nblk > ctx->partial_mv_cap) {
        cudaFree(ctx->d_partial_mv);
        ESIMP_CUDA(cudaMalloc(&ctx->d_partial_mv, sizeof(double) * nblk));
        ctx->partial_mv_cap = nblk;
    }
    RedOut<1> dot_out{{ctx->d_scal + (DOT ? dot_slot : 0)}};
    // only fine-grid applies are timed (multigrid levels reuse this kernel on coarser grids)
    // every fine-grid launch is timed; tag 0 = plain / dot applies (the roofline number), 1 = fused smoother / residual modes
    const bool timed = ctx->prof_on >= (MODE <= FUSE_DOT ? 1 : 2) && fine && 2 * ctx->prof_n + 1 < (int)ctx->prof_ev.size();
    if (timed) ctx->prof_tag[ctx->prof_n] = MODE <= FUSE_DOT ? PROF_OP_PLAIN : (MODE == FUSE_RESID ? PROF_OP_RESID : PROF_OP_SMOOTH);
    if (timed) cudaEventRecord(ctx->prof_ev[2 * ctx->prof_n], ctx->stream);
    kern<<<grid, block, smem, ctx->stream>>>(g, vw.mv_e0, vw.mv_e1, vw.own_p0, vw.own_p1, chunk, E_t, d_in, d_out,
                                             ctx->d_partial_mv, ctx->d_ticket + 2, dot_out, fa);
    if (timed) cudaEventRecord(ctx->prof_ev[2 * ctx->prof_n++ + 1], ctx->stream);
    ctx->launches++;
    ESIMP_CUDA(cudaGetLastError());
    return ESIMP_OK;
}


// ---------------------------------------------------------------------------- the z-sweeping TMA kernel (sweep_kernel.cuh)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}
// rank-3 FP64 tensor map: dims / strides in elements (innermost first), no swizzle, out-of-bounds elements read as zero
static int make_map3(esimp_ctx* ctx, CUtensorMap* tm, const void* base, cuuint64_t d0, cuuint64_t d1, cuuint64_t d2, cuuint64_t s1, cuuint64_t s2,
                     cuuint32_t b0, cuuint32_t b1, cuuint32_t b2) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return esimp_fail(ctx, ESIMP_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
    cuuint64_t dim[3] = {d0, d1, d2}, stride[2] = {s1 * 8, s2 * 8};
    cuuint32_t box[3] = {b0, b1, b2}, es[3] = {1, 1, 1};
    // L2 promotion: the boxes are 256- / 272-byte row pieces that start 64-byte aligned, so a 256-byte promotion granule fetches
    // up to twice the bytes from DRAM (ncu r2: 1.73 GB read per plain apply against 1.28 GB of box bytes; the smoother mode with
    // its three operand boxes per plane went from 1.45 ms to 1.13 ms with 64-byte granules, 1.30 / 1.31 ms with 128 / none).
    // ESIMP_TMA_PROMO = 0 .. 3 selects none / 64 / 128 / 256 bytes; default 64
    static const int promo = getenv("ESIMP_TMA_PROMO") ? atoi(getenv("ESIMP_TMA_PROMO")) : 1;
    const CUtensorMapL2promotion l2p = promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                                     : promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                     : promo == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<void*>(base), dim, stride, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, l2p, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return esimp_fail(ctx, ESIMP_ERR_CUDA, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
    return ESIMP_OK;
}
// the sweep kernel needs 16-byte strides (even nz for the modulus rows; the node rows are padded) and enough work per CTA
bool hex8_sweep_eligible(const esimp_ctx* ctx, const OpView& vw) {
    static const int env = getenv("ESIMP_SWEEP") ? atoi(getenv("ESIMP_SWEEP")) : 1;
    if (!env) return false;
    const int v = ctx->prm.matvec_variant;
    if (v != ESIMP_MATVEC_AUTO && v != ESIMP_MATVEC_SWEEP) return false;
    const GridDev& g = vw.g;
    if ((g.nz % 2) != 0 || (g.nzp % 2) != 0) return false;  // TMA strides are multiples of 16 bytes
    if (v == ESIMP_MATVEC_SWEEP) return true;                // asked for by name: any size (tests)
    return g.nz >= 64 && g.ny >= 30 && vw.own_p1 - vw.own_p0 >= 8;
}
template <int MODE, int RING, int NW1>
static auto sweep_kernel_ptr() {
    if constexpr (NW1 == 0)
        return k_matvec_sweep<MODE, RING>;
    else
        return k_matvec_sweep1<MODE, RING, NW1>;
}
// NW1 = 0: k_matvec_sweep (two cells per thread, 8 consumer warps); NW1 = 16 | 12: k_matvec_sweep1 (one cell per thread, NW1 consumer warps)
template <int MODE, int RING, int NW1 = 0>
static int launch_sweep_r(esimp_ctx* ctx, const OpView& vw, bool fine, const double* d_in, double* d_out, int dot_slot, const FuseArgs& fa) {
    using TR = SweepTraits<MODE>;
    using Smem = std::conditional_t<NW1 == 0, SweepSmem<TR::NF, RING>, Sweep1Smem<TR::NF, RING, NW1 == 0 ? 16 : NW1>>;
    constexpr int SW_CROWS = NW1 == 0 ? ::SW_CROWS : NW1, SW_ROWS = SW_CROWS + 1, SW_OROWS = SW_CROWS - 1, SW_CARRY_ROW = SW_CROWS * 3;
    constexpr int SW_THREADS = NW1 == 0 ? ::SW_THREADS : 32 * (NW1 + S1_PW);
    const GridDev& g = vw.g;
    const int nplanes = vw.own_p1 - vw.own_p0;
    const int nty = (g.ny + 1 + SW_CROWS - 2) / (SW_CROWS - 1);
    // x chunks: one CTA per SM; pick the chunk count whose waves x (planes per chunk + restart cost) is smallest
    const int smem_limit = 232448 - 1024;
    const int max_chunk = std::min(nplanes, (int)((smem_limit - (int)sizeof(Smem)) / (SW_CARRY_ROW * 8)) - 3);
    int chunk = std::max(1, max_chunk);
    {
        double best = 1e300;
        for (int nchunk = 1; nchunk <= nplanes; ++nchunk) {
            const int c = (nplanes + nchunk - 1) / nchunk;
            if (c > max_chunk) continue;
            const int nc = (nplanes + c - 1) / c;
            const double waves = ceil((double)nty * nc / 148.0);
            const double cost = waves * (c + 3.0);
            if (cost < best - 1e-9) {
                best = cost;
                chunk = c;
            }
            if (c <= 4) break;
        }
    }
    if (ctx->fused_chunk > 0) chunk = std::min(ctx->fused_chunk, std::max(1, max_chunk));
    const int nchunk = (nplanes + chunk - 1) / chunk;
    const size_t smem = sizeof(Smem) + (size_t)(chunk + 3) * SW_CARRY_ROW * 8;
    static std::atomic<unsigned long long> attr_devices{0};
    auto kern = sweep_kernel_ptr<MODE, RING, NW1>();
    const unsigned long long dev_bit = 1ull << (ctx->device & 63);
    if (!(attr_devices.load() & dev_bit)) {
        ESIMP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_limit));
        attr_devices.fetch_or(dev_bit);
    }
    SweepArgs a{};
    a.g = g;
    a.le0 = vw.mv_e0;
    a.le1 = vw.mv_e1;
    a.ps0 = vw.own_p0;
    a.ps1 = vw.own_p1;
    a.chunk = chunk;
    a.y = d_out;
    a.d_out = fa.d;
    a.c1 = fa.c1;
    a.c2 = fa.c2;
    a.dmode = fa.c1 != 0.0 ? fa.dmode : FUSE_D_NONE;
    a.write_d = fa.write_d;
    static const int pfd = getenv("ESIMP_SWEEP_PFD") ? atoi(getenv("ESIMP_SWEEP_PFD")) : 0;
    a.pf_dist = pfd;
    const long long nblk = (long long)nty * nchunk;
    if (TR::DOT && nblk > ctx->partial_mv_cap) {
        cudaFree(ctx->d_partial_mv);
        ESIMP_CUDA(cudaMalloc(&ctx->d_partial_mv, sizeof(double) * nblk));
        ctx->partial_mv_cap = nblk;
    }
    a.partial = ctx->d_partial_mv;
    a.ticket = ctx->d_ticket + 2;
    a.dot_out = RedOut<1>{{ctx->d_scal + (TR::DOT ? dot_slot : 0)}};
    memcpy(a.blk.B, ctx->B, sizeof(ctx->B));
    alignas(64) CUtensorMap tmU, tmE, tmOp[3], tmOpx[3];
    const cuuint64_t nz1 = g.nz + 1, ny1 = g.ny + 1, n3 = 3ull * g.NPX;
    ESIMP_TRY(make_map3(ctx, &tmU, d_in, nz1, ny1, n3, g.nzp, g.plane, SW_BOXZ, SW_ROWS, 3));
    ESIMP_TRY(make_map3(ctx, &tmE, vw.E, g.nz, g.ny, g.NEX, g.nz, g.eplane, 32, SW_CROWS, 1));
    for (int b = 0; b < 3; ++b) tmOp[b] = tmOpx[b] = tmU;
    const double* ops[3] = {fa.r, fa.sinv, fa.dmode == FUSE_D_FROM_ZPREV ? fa.zprev : fa.d};
    for (int b = 0; b < TR::NF; ++b) {  // r | r, Dinv, d: exactly the owned nodes of a segment, and the extra column nz
        const bool scalar = TR::NF > 1 && b == 1;  // D^-1 travels as one double per node (sw_dinv): planes of [ny+1][nzp], one per x plane
        ESIMP_TRY(make_map3(ctx, &tmOp[b], ops[b], nz1, ny1, scalar ? (cuuint64_t)g.NPX : n3, g.nzp, g.plane, 32, SW_OROWS, scalar ? 1 : 3));
        ESIMP_TRY(make_map3(ctx, &tmOpx[b], ops[b], nz1, ny1, scalar ? (cuuint64_t)g.NPX : n3, g.nzp, g.plane, 2, SW_OROWS, scalar ? 1 : 3));
    }
    for (int c = 0; c < 3; ++c) a.ik[c] = 1.0 / ctx->kdiag[c];
    const bool timed = ctx->prof_on >= (MODE <= FUSE_DOT ? 1 : 2) && fine && 2 * ctx->prof_n + 1 < (int)ctx->prof_ev.size();
    if (timed) ctx->prof_tag[ctx->prof_n] = MODE <= FUSE_DOT ? PROF_OP_PLAIN : (MODE == FUSE_RESID ? PROF_OP_RESID : PROF_OP_SMOOTH);
    if (timed) cudaEventRecord(ctx->prof_ev[2 * ctx->prof_n], ctx->stream);
    kern<<<dim3(nty, nchunk), SW_THREADS, smem, ctx->stream>>>(a, tmU, tmE, tmOp[0], tmOp[1], tmOp[2], tmOpx[0], tmOpx[1], tmOpx[2]);
    if (timed) cudaEventRecord(ctx->prof_ev[2 * ctx->prof_n++ + 1], ctx->stream);
    ctx->launches++;
    ESIMP_CUDA(cudaGetLastError());
    return ESIMP_OK;
}

template <int MODE>
static int launch_sweep(esimp_ctx* ctx, const OpView& vw, bool fine, const double* d_in, double* d_out, int dot_slot, const FuseArgs& fa) {
    static const int ring = getenv("ESIMP_SWEEP_RING") ? atoi(getenv("ESIMP_SWEEP_RING")) : 0;  // tuning: depth of the input ring
    // ESIMP_SWEEP_CPT1=12: one cell per thread with 12 consumer warps (sweep1_kernel.cuh) instead of two cells per thread with 8.  An
    // experiment that LOST on the B200 (profiles/r2_one_cell_per_thread_sweep_isolated.log, 512x256x256, ms per launch, plain / +dot /
    // smoother / residual / smoother+dot): 0.66 / 0.68 / 1.18 / 0.76 / 1.25 against 0.56 / 0.63 / 0.96 / 0.67 / 1.02 for the default kernel;
    // with 16 consumer warps (112 registers, spills) 0.74 / 1.07 / 2.04 / 1.12 / 2.29.  Parity is green (same log set); it stays as an
    // opt-in so that the measurement can be repeated.  ESIMP_SWEEP_CPT1_MODES: bit mask of the store modes that take it (default all)
    static const int cpt1 = getenv("ESIMP_SWEEP_CPT1") ? atoi(getenv("ESIMP_SWEEP_CPT1")) : 0;
    static const int cpt1_modes = getenv("ESIMP_SWEEP_CPT1_MODES") ? atoi(getenv("ESIMP_SWEEP_CPT1_MODES")) : 31;
    if (cpt1 == 12 && ((cpt1_modes >> MODE) & 1)) return launch_sweep_r<MODE, SweepTraits<MODE>::RING, 12>(ctx, vw, fine, d_in, d_out, dot_slot, fa);
    if (SweepTraits<MODE>::NF == 0) {
        if (ring == 4) return launch_sweep_r<MODE, 4>(ctx, vw, fine, d_in, d_out, dot_slot, fa);
        if (ring == 6) return launch_sweep_r<MODE, 6>(ctx, vw, fine, d_in, d_out, dot_slot, fa);
        if (ring == 7) return launch_sweep_r<MODE, 7>(ctx, vw, fine, d_in, d_out, dot_slot, fa);
    }
    return launch_sweep_r<MODE, SweepTraits<MODE>::RING>(ctx, vw, fine, d_in, d_out, dot_slot, fa);
}

// the FP64 operator: modulus from the view, fine grid = the context's own modulus array
template <int TY, int TZ, bool ATOMIC, int MINB, int MODE>
static int launch_fused_t(esimp_ctx* ctx, const OpView& vw, const double* d_in, double* d_out, int dot_slot, const FuseArgs& fa = FuseArgs()) {
    return launch_fused_tt<double, TY, TZ, ATOMIC, MINB, MODE>(ctx, vw, vw.E, vw.E == ctx->d_E, d_in, d_out, dot_slot, fa);
}

template <int TY, int TZ, bool ATOMIC, int MINB>
static int launch_fused(esimp_ctx* ctx, const OpView& vw, const double* d_in, double* d_out, int dot_slot) {
    return dot_slot >= 0 ? launch_fused_t<TY, TZ, ATOMIC, MINB, FUSE_DOT>(ctx, vw, d_in, d_out, dot_slot)
                         : launch_fused_t<TY, TZ, ATOMIC, MINB, FUSE_PLAIN>(ctx, vw, d_in, d_out, -1);
}

bool hex8_matvec_fuses_dot(const esimp_ctx* ctx) {
    return ctx->structured && ctx->prm.matvec_variant != ESIMP_MATVEC_DENSE_ATOMIC;
}

// y = P K P u (prescribed rows are fixed up by vec_fixup_dirichlet).  dot_slot >= 0: also leaves <u, y> over the
// owned dofs in d_scal[dot_slot] (fused into the operator's store; structured fused variants only)
int hex8_matvec_view(esimp_ctx* ctx, const OpView& vw, const double* d_in, double* d_out, int variant, int dot_slot) {
    ESIMP_TRY(hex8_upload_constants(ctx));
    const GridDev& g = vw.g;
    // 16x16 is the default: 8x16 (4 CTAs/SM) is 7 % faster in isolation but not inside the power-capped solve
    if ((variant == ESIMP_MATVEC_AUTO || variant == ESIMP_MATVEC_SWEEP) && hex8_sweep_eligible(ctx, vw)) {
        const bool fine = vw.E == ctx->d_E;
        return dot_slot >= 0 ? launch_sweep<FUSE_DOT>(ctx, vw, fine, d_in, d_out, dot_slot, FuseArgs())
                             : launch_sweep<FUSE_PLAIN>(ctx, vw, fine, d_in, d_out, -1, FuseArgs());
    }
    if (variant == ESIMP_MATVEC_AUTO || variant == ESIMP_MATVEC_SWEEP) variant = ESIMP_MATVEC_FUSED;
    if (variant == ESIMP_MATVEC_DENSE_ATOMIC) {
        ESIMP_CUDA(cudaMemsetAsync(d_out, 0, sizeof(double) * 3 * g.NPX * g.plane, ctx->stream));
        long long n = (long long)(vw.mv_e1 - vw.mv_e0) * g.eplane;
        k_matvec_dense_atomic<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(g, vw.mv_e0, vw.mv_e1, vw.E, vw.mask, d_in, d_out);
        ctx->launches++;
        ESIMP_CUDA(cudaGetLastError());
        return ESIMP_OK;
    }
    if (variant == ESIMP_MATVEC_FUSED_ATOMIC) return launch_fused<8, 32, true, 2>(ctx, vw, d_in, d_out, dot_slot);
    if (variant == ESIMP_MATVEC_FUSED_8x32) return launch_fused<8, 32, false, 2>(ctx, vw, d_in, d_out, dot_slot);
    if (variant == ESIMP_MATVEC_FUSED_8x16) return launch_fused<8, 16, false, 4>(ctx, vw, d_in, d_out, dot_slot);
    // owner-computes tiles: 16x16 wastes less on the perimeter than 8x32
    return launch_fused<16, 16, false, 2>(ctx, vw, d_in, d_out, dot_slot);
}
// multigrid smoother step / residual fused into the operator's store (16x16 owner tiles: measured faster than 8x16 for these modes)
int hex8_smooth_view(esimp_ctx* ctx, const OpView& vw, const double* z_in, double* z_out, const double* r, const double* dinv,
                     double* d, double c1, double c2, int dot_slot, int dmode, int write_d, const double* zprev, const double* sinv) {
    ESIMP_TRY(hex8_upload_constants(ctx));
    FuseArgs fa;
    fa.r = r;
    fa.dinv = dinv;
    fa.sinv = sinv;
    fa.d = d;
    fa.c1 = c1;
    fa.c2 = c2;
    fa.dmode = dmode;
    fa.write_d = write_d;
    fa.zprev = zprev;
    // smoother modes: with 256-byte TMA promotion the sweep kernel's operand boxes over-fetched (1.39 - 1.45 ms against 1.29 ms for the
    // round-1 kernel at 512x256x256); with 64-byte granules it is the faster one (inside the SIMP step: 0.90 ms against 1.15 ms per
    // smoother launch, profiles/README.md).  ESIMP_SWEEP_SMOOTH=0 goes back to the round-1 kernel for these modes.
    static const int sweep_smooth = getenv("ESIMP_SWEEP_SMOOTH") ? atoi(getenv("ESIMP_SWEEP_SMOOTH")) : 1;
    if (sinv && (sweep_smooth || ctx->prm.matvec_variant == ESIMP_MATVEC_SWEEP) && hex8_sweep_eligible(ctx, vw)) {
        const bool fine = vw.E == ctx->d_E;
        return dot_slot >= 0 ? launch_sweep<FUSE_SMOOTH_DOT>(ctx, vw, fine, z_in, z_out, dot_slot, fa)
                             : launch_sweep<FUSE_SMOOTH>(ctx, vw, fine, z_in, z_out, -1, fa);
    }
    if (dot_slot >= 0) return launch_fused_t<16, 16, false, 2, FUSE_SMOOTH_DOT>(ctx, vw, z_in, z_out, dot_slot, fa);
    return launch_fused_t<16, 16, false, 2, FUSE_SMOOTH>(ctx, vw, z_in, z_out, -1, fa);
}
int hex8_resid_view(esimp_ctx* ctx, const OpView& vw, const double* z_in, double* res_out, const double* r) {
    ESIMP_TRY(hex8_upload_constants(ctx));
    FuseArgs fa;
    fa.r = r;
    if (hex8_sweep_eligible(ctx, vw)) return launch_sweep<FUSE_RESID>(ctx, vw, vw.E == ctx->d_E, z_in, res_out, -1, fa);
    return launch_fused_t<16, 16, false, 2, FUSE_RESID>(ctx, vw, z_in, res_out, -1, fa);
}
// ---- single-precision instances for the FP32 multigrid cycle (multigrid32.cu): same kernel, T = float
static int upload_constants_f32(esimp_ctx* ctx) {
    const bool stale = !ctx->k0_valid;  // hex8_upload_constants recomputes B then
    ESIMP_TRY(hex8_upload_constants(ctx));
    std::lock_guard<std::mutex> lock(g_const_mutex);
    if (stale || g_const_owner_f32[ctx->device & 63] != ctx) {
        float Bf[8][3][3];
        for (int q = 0; q < 8; ++q)
            for (int c = 0; c < 3; ++c)
                for (int d = 0; d < 3; ++d) Bf[q][c][d] = (float)ctx->B[q][c][d];
        // stream ordered like the FP64 constants; a pageable source is staged before the call returns, so a local is fine
        ESIMP_CUDA(cudaMemcpyToSymbolAsync(c_Bf, Bf, sizeof(Bf), 0, cudaMemcpyHostToDevice, ctx->stream));
        g_const_owner_f32[ctx->device & 63] = ctx;
    }
    return ESIMP_OK;
}
int hex8_smooth_view_f32(esimp_ctx* ctx, const OpView& vw, const float* E, const float* z_in, float* z_out, const float* r,
                         const float* dinv, float* d, float c1, float c2) {
    ESIMP_TRY(upload_constants_f32(ctx));
    FuseArgsT<float> fa;
    fa.r = r;
    fa.dinv = dinv;
    fa.d = d;
    fa.c1 = c1;
    fa.c2 = c2;
    // tuning knob for round 2: 3 CTAs of 256 threads per SM (<= 85 registers) instead of 2
    static const bool three = getenv("ESIMP_F32_CTAS") && atoi(getenv("ESIMP_F32_CTAS")) == 3;
    if (three) return launch_fused_tt<float, 16, 16, false, 3, FUSE_SMOOTH>(ctx, vw, E, vw.g.nx == ctx->g.nx, z_in, z_out, -1, fa);
    return launch_fused_tt<float, 16, 16, false, 2, FUSE_SMOOTH>(ctx, vw, E, vw.g.nx == ctx->g.nx, z_in, z_out, -1, fa);
}
int hex8_resid_view_f32(esimp_ctx* ctx, const OpView& vw, const float* E, const float* z_in, float* res_out, const float* r) {
    ESIMP_TRY(upload_constants_f32(ctx));
    FuseArgsT<float> fa;
    fa.r = r;
    static const bool three = getenv("ESIMP_F32_CTAS") && atoi(getenv("ESIMP_F32_CTAS")) == 3;
    if (three) return launch_fused_tt<float, 16, 16, false, 3, FUSE_RESID>(ctx, vw, E, vw.g.nx == ctx->g.nx, z_in, res_out, -1, fa);
    return launch_fused_tt<float, 16, 16, false, 2, FUSE_RESID>(ctx, vw, E, vw.g.nx == ctx->g.nx, z_in, res_out, -1, fa);
}
OpView fine_view(const esimp_ctx* ctx) {
    OpView v;
    v.g = ctx->g;
    v.own_p0 = ctx->own_p0;
    v.own_p1 = ctx->own_p1;
    v.mv_e0 = ctx->mv_e0;
    v.mv_e1 = ctx->mv_e1;
    v.E = ctx->d_E;
    v.mask = ctx->d_mask;
    return v;
}
int hex8_matvec(esimp_ctx* ctx, const double* d_in, double* d_out, int variant, int dot_slot) {
    return hex8_matvec_view(ctx, fine_view(ctx), d_in, d_out, variant, dot_slot);
}

// ---------------------------------------------------------------------------- Jacobi diagonal
// diag(K)[node,c] = K0cc[c] * sum of E over the <=8 cells around the node; prescribed dofs get
// diag_value (apply! puts m = mean|diag K| there).  Pass 1 accumulates sum(diag) for m.
__global__ void k_diag(GridDev g, int ps0, int ps1, int NEX, const double* __restrict__ E, double* __restrict__ diag,
                       double* partial, unsigned int* ticket, RedOut<1> out) {
    long long n = (long long)(ps1 - ps0) * g.plane;
    double acc[1] = {0.0};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        int k = (int)(i % g.nzp);
        if (k > g.nz) continue;  // row pads: no node (their diag / dinv entries stay zero)
        int j = (int)((i / g.nzp) % (g.ny + 1));
        int lp = ps0 + (int)(i / g.plane);
        double s = 0.0;
#pragma unroll
        for (int dx = -1; dx <= 0; ++dx) {
            int le = lp - g.e2p + dx;
            int ge = g.ge0 + le;
            if (le < 0 || le >= NEX || ge < 0 || ge >= g.nx) continue;
#pragma unroll
            for (int dy = -1; dy <= 0; ++dy) {
                int jj = j + dy;
                if (jj < 0 || jj >= g.ny) continue;
#pragma unroll
                for (int dz = -1; dz <= 0; ++dz) {
                    int kk = k + dz;
                    if (kk < 0 || kk >= g.nz) continue;
                    s += E[g.cell(le, jj, kk)];
                }
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            double d = c_kdiag[c] * s;
            diag[g.dof(lp, c, j, k)] = d;
            acc[0] += d;
        }
    }
    grid_reduce<1, false>(acc, partial, ticket, out);
}

__global__ void k_diag_invert(GridDev g, int ps0, int ps1, const uint8_t* __restrict__ mask, double m,
                              double* __restrict__ dinv) {
    long long n = (long long)(ps1 - ps0) * 3 * g.plane;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        long long d = (long long)ps0 * 3 * g.plane + i;
        int lp = (int)(d / (3 * g.plane));
        int c = (int)((d / g.plane) % 3);
        long long nn = (long long)lp * g.plane + d % g.plane;
        bool fixed = ((mask[nn] >> c) & 1) || (int)(d % g.nzp) > g.nz;  // row pads: dinv = 0 like prescribed dofs
        dinv[d] = fixed ? 0.0 : 1.0 / dinv[d];  // 0: prescribed dofs never move (r = 0 there; the fused smoother relies on it)
    }
}

int hex8_diagonal_view(esimp_ctx* ctx, const OpView& vw, double* dinv, long long n_dofs_global, double* diag_value) {
    ESIMP_TRY(hex8_upload_constants(ctx));
    RedOut<1> out{{ctx->d_scal + SC_DIAGSUM}};
    k_diag<<<RED_BLOCKS, RED_THREADS, 0, ctx->stream>>>(vw.g, vw.own_p0, vw.own_p1, vw.g.NEX, vw.E, dinv, ctx->d_partial, ctx->d_ticket, out);
    ctx->launches++;
    ESIMP_CUDA(cudaGetLastError());
    ESIMP_TRY(allreduce_scalars(ctx, SC_DIAGSUM, 1, false));
    ESIMP_TRY(fetch_scalars(ctx));
    *diag_value = ctx->h_scal[SC_DIAGSUM] / (double)n_dofs_global;  // mean |diag K| (all entries > 0)
    k_diag_invert<<<RED_BLOCKS, RED_THREADS, 0, ctx->stream>>>(vw.g, vw.own_p0, vw.own_p1, vw.mask, *diag_value, dinv);
    ctx->launches++;
    ESIMP_CUDA(cudaGetLastError());
    return ESIMP_OK;
}
// sinv[node] = 1 / (sum of E around the node) = dinv[node, c] * K0cc[c] for any component c that is not prescribed, with the node's
// prescribed-component mask in the three lowest mantissa bits (sweep_kernel.cuh, sw_dinv); pads and fully prescribed nodes: 0 | mask
__global__ void k_node_sinv(GridDev g, int ps0, int ps1, const uint8_t* __restrict__ mask, const double* __restrict__ dinv,
                            double* __restrict__ sinv) {
    const long long n = (long long)(ps1 - ps0) * g.plane;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int lp = ps0 + (int)(i / g.plane);
        const long long in_plane = i % g.plane;
        const long long nn = (long long)lp * g.plane + in_plane;
        const bool pad = (int)(in_plane % g.nzp) > g.nz;
        const unsigned m = pad ? 7u : (mask[nn] & 7u);
        double s = 0.0;
#pragma unroll
        for (int c = 2; c >= 0; --c)
            if (!((m >> c) & 1)) s = dinv[((long long)lp * 3 + c) * g.plane + in_plane] * c_kdiag[c];
        sinv[nn] = __longlong_as_double((__double_as_longlong(s) & ~7ll) | (long long)m);
    }
}
int hex8_node_scalar_dinv(esimp_ctx* ctx, const OpView& vw, const double* dinv, double* sinv) {
    ESIMP_TRY(hex8_upload_constants(ctx));
    k_node_sinv<<<RED_BLOCKS, RED_THREADS, 0, ctx->stream>>>(vw.g, vw.own_p0, vw.own_p1, vw.mask, dinv, sinv);
    ctx->launches++;
    ESIMP_CUDA(cudaGetLastError());
    return ESIMP_OK;
}
int hex8_diagonal(esimp_ctx* ctx) {
    return hex8_diagonal_view(ctx, fine_view(ctx), ctx->d_dinv, ctx->n_dofs_global, &ctx->diag_value);
}

// ---------------------------------------------------------------------------- sensitivities + energy
// One thread per owned cell: q_e = u_e' K0 u_e = v' K~ v in mode space (45 products),
// dc_e = -p rho^(p-1) (E0-Emin) q_e (SensitivityAnalysis.jl:74-78), energy += 0.5 E_e q_e.
// u is used as is (the reference does not mask it; after a solve u = 0 on prescribed dofs).
__global__ void __launch_bounds__(256) k_sens_energy(GridDev g, int le0, int le1, const double* __restrict__ u,
                                                     const double* __restrict__ rho,
                                                     const double* __restrict__ E, double pexp, double dE,
                                                     double* __restrict__ dc, double* partial, unsigned int* ticket,
                                                     RedOut<1> out) {
    long long n = (long long)(le1 - le0) * g.eplane;
    double acc[1] = {0.0};
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
        int le = le0 + (int)(idx / g.eplane);
        int rem = (int)(idx % g.eplane);
        int j = rem / g.nz, k = rem % g.nz;
        int lp = le + g.e2p;
        double v[3][8];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            double F[2][4];
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                double uu[2][2];
#pragma unroll
                for (int a = 0; a < 2; ++a)
#pragma unroll
                    for (int b = 0; b < 2; ++b) {
                        uu[a][b] = u[g.dof(lp + s, c, j + a, k + b)];
                    }
                fwd2d(uu[0][0], uu[1][0], uu[0][1], uu[1][1], F[s]);
            }
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                v[c][2 * b] = F[1][b] + F[0][b];
                v[c][2 * b + 1] = F[1][b] - F[0][b];
            }
        }
        double w[3][8];
        apply_blocks(v, w);
        double q = 0.0;
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int m = 0; m < 8; ++m)
                if (!(m == 0)) q = fma(v[c][m], w[c][m], q);
        long long e = g.cell(le, j, k);
        double r = rho[e];
        double rpow = (pexp == 3.0) ? r * r : pow(r, pexp - 1.0);
        dc[e] = -(pexp * rpow * dE) * q;
        acc[0] += 0.5 * E[e] * q;
    }
    grid_reduce<1, false>(acc, partial, ticket, out);
}

int hex8_sens_energy(esimp_ctx* ctx) {
    ESIMP_TRY(hex8_upload_constants(ctx));
    RedOut<1> out{{ctx->d_scal + SC_ENERGY}};
    PROF_BEGIN(ctx, PROF_SENS, true)
    k_sens_energy<<<RED_BLOCKS, 256, 0, ctx->stream>>>(ctx->g, ctx->own_e0, ctx->own_e1, ctx->d_u, ctx->d_phys,
                                                       ctx->d_E, ctx->prm.p, ctx->prm.E0 - ctx->prm.Emin, ctx->d_dc,
                                                       ctx->d_partial, ctx->d_ticket, out);
    PROF_END(ctx)
    ctx->launches++;
    ESIMP_CUDA(cudaGetLastError());
    return allreduce_scalars(ctx, SC_ENERGY, 1, false);
}
