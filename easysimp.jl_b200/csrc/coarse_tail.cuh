// coarse_tail.cuh -- the bottom of the multigrid V-cycle as ONE kernel on one thread-block cluster (sm_100a).
//
// Below ~10 000 nodes a level is pure latency: every smoother step, residual, restriction and prolongation is a launch of a few
// microseconds (r2 launch list at 512x256x256: 66 launches of 7.2 us on average per CG iteration, 0.47 ms out of 7 ms, the same on
// every rank count -- the part of the V-cycle that does not strong-scale).  This kernel walks the levels l0 .. last of the cycle
// (32x16x16 cells and smaller at 512x256x256) inside one launch: 8 CTAs of one cluster, one thread per node, the phases of
// mg_vcycle / mg_smooth (csrc/multigrid.cu) separated by hardware cluster barriers instead of kernel boundaries.
//
// The arithmetic is the one the level-by-level path does: Chebyshev-Jacobi steps z' = z + c1 d + c2 Dinv (r - A z) with the same
// coefficients and the same "previous direction = z or z - z_prev" rule, 27-point full-weighting restriction, trilinear
// prolongation, Dirichlet dofs held at zero through Dinv = 0 / the level masks.  Only the operator is evaluated differently: a node
// gathers E_e K0[row of the node, :] u_e from its <= 8 cells with the dense 24 x 24 unit matrix in shared memory (the mode-space
// x-march of the big kernels needs planes of thousands of nodes to pay).  Sums are taken in another order than in the fused
// kernels, so the preconditioner differs in the last bits -- like it does between the sweep and the round-1 kernel.
// Needs levels that are whole on this rank (one rank, or replicated levels).
#pragma once
#include <cooperative_groups.h>

constexpr int CT_MAX_LEVELS = 8;
constexpr int CT_CTAS = 8;        // portable cluster size
constexpr int CT_THREADS = 512;
constexpr long long CT_MAX_NODES = 10000;  // largest level the tail takes

struct CoarseTailLevel {
    int nx, ny, nz, nzp;          // cells, node-row pitch
    const double* E;              // [nx][ny][nz]
    const uint8_t* mask;          // [nx+1][ny+1][nzp], bit c = dof c prescribed
    const double* dinv;           // node-vector layout [plane][c][j][nzp]
    double *r, *za, *zb, *q;      // rhs, correction (current / other buffer), residual
    double theta, delta;          // Chebyshev interval of this level
};
// doubles of shared memory a CTA needs to stage its slab (+ one plane either side) of a level
inline long long coarse_tail_slab_doubles(int nx, int ny, int nz) {
    const long long planes = (nx + 1 + CT_CTAS - 1) / CT_CTAS + 2;
    return planes * 3 * (ny + 1) * (nz + 1);
}
struct CoarseTailArgs {
    int nlev, nu, coarse_its;
    const double* K0;             // 24 x 24, local dof 3a + c, Ferrite node order
    CoarseTailLevel lv[CT_MAX_LEVELS];
};

namespace coarse_tail {
namespace cg = cooperative_groups;

struct Lv {
    int nx, ny, nz, nzp, nnodes;
    long long plane;
    __device__ __forceinline__ long long dof(int i, int c, int j, int k) const { return ((long long)(i * 3 + c) * (ny + 1) + j) * nzp + k; }
    __device__ __forceinline__ long long node(int i, int j, int k) const { return ((long long)i * (ny + 1) + j) * nzp + k; }
};
__device__ __forceinline__ Lv make_lv(const CoarseTailLevel& L) {
    Lv g;
    g.nx = L.nx; g.ny = L.ny; g.nz = L.nz; g.nzp = L.nzp;
    g.nnodes = (L.nx + 1) * (L.ny + 1) * (L.nz + 1);
    g.plane = (long long)(L.ny + 1) * L.nzp;
    return g;
}
// node n (0 .. nnodes) -> (i, j, k), k fastest
__device__ __forceinline__ void node_ijk(const Lv& g, int n, int& i, int& j, int& k) {
    k = n % (g.nz + 1);
    const int t = n / (g.nz + 1);
    j = t % (g.ny + 1);
    i = t / (g.ny + 1);
}

// this CTA's planes of a level: [p0, p1) of the nx + 1 node planes, split evenly over the cluster
__device__ __forceinline__ void slab(const Lv& g, int& p0, int& p1) {
    p0 = (int)((long long)(g.nx + 1) * blockIdx.x / gridDim.x);
    p1 = (int)((long long)(g.nx + 1) * (blockIdx.x + 1) / gridDim.x);
}
// copy planes p0 - 1 .. p1 of node vector u into shared memory (compact rows, no pitch pad): one coalesced read of every value
// instead of 24 gathers per node from L2 (the data was written by other CTAs one cluster barrier ago, so L1 starts cold)
__device__ __forceinline__ void stage(const Lv& g, const double* u, double* sb, int p0, int p1) {
    const int nk = g.nz + 1, nyz = (g.ny + 1) * nk;
    const int lo = max(p0 - 1, 0), hi = min(p1 + 1, g.nx + 1);
    const int total = (hi - lo) * 3 * nyz;
    for (int t = threadIdx.x; t < total; t += blockDim.x) {
        const int k = t % nk, j = (t / nk) % (g.ny + 1), pc = t / nyz;
        const int c = pc % 3, pl = lo + pc / 3;
        sb[((pl - (p0 - 1)) * 3 + c) * nyz + j * nk + k] = u[g.dof(pl, c, j, k)];
    }
}
// (A u) at node (i, j, k): sum over the cells around the node of E_e K0[3a + c, :] u_e, u from the staged slab
__device__ __forceinline__ void apply_node(const Lv& g, const double* __restrict__ E, const double* sK0, const double* sb, int p0, int i, int j,
                                           int k, double (&y)[3]) {
    const int nk = g.nz + 1, nyz = (g.ny + 1) * nk;
    y[0] = y[1] = y[2] = 0.0;
#pragma unroll 1
    for (int cell = 0; cell < 8; ++cell) {
        const int di = cell & 1, dj = (cell >> 1) & 1, dk = (cell >> 2) & 1;  // the node is corner (di, dj, dk) of this cell
        const int ci = i - di, cj = j - dj, ck = k - dk;
        if (ci < 0 || ci >= g.nx || cj < 0 || cj >= g.ny || ck < 0 || ck >= g.nz) continue;
        const double Ee = E[((long long)ci * g.ny + cj) * g.nz + ck];
        // Ferrite hexahedron corners: a -> (sx, sy, sz) = (0,0,0) (1,0,0) (1,1,0) (0,1,0) (0,0,1) (1,0,1) (1,1,1) (0,1,1)
        const int a = (dk << 2) | (dj ? (di ? 2 : 3) : (di ? 1 : 0));
        double acc[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const int sx = (b == 1 || b == 2 || b == 5 || b == 6) ? 1 : 0, sy = (b == 2 || b == 3 || b == 6 || b == 7) ? 1 : 0, sz = b >> 2;
            const double* up = sb + ((ci + sx - (p0 - 1)) * 3) * nyz + (cj + sy) * nk + (ck + sz);
            const double u0 = up[0], u1 = up[nyz], u2 = up[2 * nyz];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const double* row = sK0 + (3 * a + c) * 24 + 3 * b;
                acc[c] = fma(row[0], u0, fma(row[1], u1, fma(row[2], u2, acc[c])));
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) y[c] = fma(Ee, acc[c], y[c]);
    }
}

// one Chebyshev step on every node: zout = z + c1 dprev + c2 Dinv (r - A z); dmode: 0 no previous direction, 1 dprev = z,
// 2 dprev = z - zout_old (zout holds the iterate before z on entry)
__device__ __forceinline__ void smooth_step(const Lv& g, const CoarseTailLevel& L, const double* sK0, double* sb, const double* z, double* zout,
                                            double c1, double c2, int dmode) {
    int p0, p1;
    slab(g, p0, p1);
    stage(g, z, sb, p0, p1);
    __syncthreads();
    const int nk = g.nz + 1, nyz = (g.ny + 1) * nk;
    for (int n = threadIdx.x; n < (p1 - p0) * nyz; n += blockDim.x) {
        const int k = n % nk, j = (n / nk) % (g.ny + 1), i = p0 + n / nyz;
        double Az[3];
        apply_node(g, L.E, sK0, sb, p0, i, j, k, Az);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const long long d = g.dof(i, c, j, k);
            const double zin = sb[((i - (p0 - 1)) * 3 + c) * nyz + j * nk + k];
            const double dprev = dmode == 1 ? zin : (dmode == 2 ? zin - zout[d] : 0.0);
            zout[d] = zin + fma(c1, dprev, c2 * L.dinv[d] * (L.r[d] - Az[c]));
        }
    }
}

// k Chebyshev steps for A z = r on one level; returns with the result in *pz (the buffers are swapped after every step that
// involves the operator, exactly like mg_smooth does on the host side)
__device__ __forceinline__ void smooth(cg::cluster_group& cl, const Lv& g, const CoarseTailLevel& L, const double* sK0, double* sb, double*& z, double*& zo,
                                       int ksteps, bool zero_guess, int tid, int nthr) {
    const double sigma = L.theta / L.delta;
    double rho = 1.0 / sigma;
    int done = 0;
    if (zero_guess) {
        for (int n = tid; n < g.nnodes; n += nthr) {
            int i, j, k;
            node_ijk(g, n, i, j, k);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const long long d = g.dof(i, c, j, k);
                z[d] = L.dinv[d] * L.r[d] * (1.0 / L.theta);
            }
        }
        done = 1;
        cl.sync();
    }
    for (int s = done; s < ksteps; ++s) {
        double c1, c2;
        if (s == 0) {
            c1 = 0.0;
            c2 = 1.0 / L.theta;
        } else {
            const double rho_new = 1.0 / (2.0 * sigma - rho);
            c1 = rho_new * rho;
            c2 = 2.0 * rho_new / L.delta;
            rho = rho_new;
        }
        const int dmode = s == 0 ? 0 : (s == 1 && zero_guess ? 1 : 2);
        smooth_step(g, L, sK0, sb, z, zo, c1, c2, dmode);
        double* t = z;
        z = zo;
        zo = t;
        cl.sync();
    }
}
}  // namespace coarse_tail

__global__ void __cluster_dims__(CT_CTAS, 1, 1) __launch_bounds__(CT_THREADS, 1) k_coarse_tail(const __grid_constant__ CoarseTailArgs a) {
    using namespace coarse_tail;
    cg::cluster_group cl = cg::this_cluster();
    __shared__ double sK0[576];
    extern __shared__ double sb[];  // staged slab of the vector the operator is applied to
    for (int t = threadIdx.x; t < 576; t += blockDim.x) sK0[t] = a.K0[t];
    __syncthreads();
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
    double* z[CT_MAX_LEVELS];
    double* zo[CT_MAX_LEVELS];
#pragma unroll
    for (int l = 0; l < CT_MAX_LEVELS; ++l) {
        z[l] = l < a.nlev ? a.lv[l].za : nullptr;
        zo[l] = l < a.nlev ? a.lv[l].zb : nullptr;
    }
    // ---- downward leg
    for (int l = 0; l < a.nlev; ++l) {
        const CoarseTailLevel& L = a.lv[l];
        const Lv g = make_lv(L);
        if (l == a.nlev - 1) {
            smooth(cl, g, L, sK0, sb, z[l], zo[l], a.coarse_its, true, tid, nthr);
            break;
        }
        smooth(cl, g, L, sK0, sb, z[l], zo[l], a.nu, true, tid, nthr);
        // residual q = r - A z (zero on prescribed dofs)
        {
            int p0, p1;
            slab(g, p0, p1);
            stage(g, z[l], sb, p0, p1);
            __syncthreads();
            const int nk = g.nz + 1, nyz = (g.ny + 1) * nk;
            for (int n = threadIdx.x; n < (p1 - p0) * nyz; n += blockDim.x) {
                const int k = n % nk, j = (n / nk) % (g.ny + 1), i = p0 + n / nyz;
                double Az[3];
                apply_node(g, L.E, sK0, sb, p0, i, j, k, Az);
                const uint8_t m = L.mask[g.node(i, j, k)];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const long long d = g.dof(i, c, j, k);
                    L.q[d] = ((m >> c) & 1) ? 0.0 : L.r[d] - Az[c];
                }
            }
        }
        cl.sync();
        // restriction onto level l + 1: 27-point full weighting, zero on prescribed coarse dofs
        const CoarseTailLevel& C = a.lv[l + 1];
        const Lv gc = make_lv(C);
        for (int n = tid; n < gc.nnodes; n += nthr) {
            int I, J, K;
            node_ijk(gc, n, I, J, K);
            const uint8_t m = C.mask[gc.node(I, J, K)];
            double acc[3] = {0.0, 0.0, 0.0};
            for (int dx = -1; dx <= 1; ++dx) {
                const int fi = 2 * I + dx;
                if (fi < 0 || fi > g.nx) continue;
                for (int dy = -1; dy <= 1; ++dy) {
                    const int fj = 2 * J + dy;
                    if (fj < 0 || fj > g.ny) continue;
                    for (int dz = -1; dz <= 1; ++dz) {
                        const int fk = 2 * K + dz;
                        if (fk < 0 || fk > g.nz) continue;
                        const double w = (dx ? 0.5 : 1.0) * (dy ? 0.5 : 1.0) * (dz ? 0.5 : 1.0);
                        const long long d0 = g.dof(fi, 0, fj, fk);
#pragma unroll
                        for (int c = 0; c < 3; ++c) acc[c] = fma(w, L.q[d0 + c * g.plane], acc[c]);
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) C.r[gc.dof(I, c, J, K)] = ((m >> c) & 1) ? 0.0 : acc[c];
        }
        cl.sync();
    }
    // ---- upward leg
    for (int l = a.nlev - 2; l >= 0; --l) {
        const CoarseTailLevel& L = a.lv[l];
        const Lv g = make_lv(L);
        const Lv gc = make_lv(a.lv[l + 1]);
        const double* zc = z[l + 1];
        double* zf = z[l];
        // z_l += P z_{l+1}: the parents of fine index i along an axis are i >> 1 and (i + 1) >> 1, weight 1/2 each
        for (int n = tid; n < g.nnodes; n += nthr) {
            int i, j, k;
            node_ijk(g, n, i, j, k);
            const uint8_t m = L.mask[g.node(i, j, k)];
            double acc[3] = {0.0, 0.0, 0.0};
#pragma unroll
            for (int p = 0; p < 8; ++p) {
                const long long d0 = gc.dof((i + (p & 1)) >> 1, 0, (j + ((p >> 1) & 1)) >> 1, (k + (p >> 2)) >> 1);
#pragma unroll
                for (int c = 0; c < 3; ++c) acc[c] += zc[d0 + c * gc.plane];
            }
#pragma unroll
            for (int c = 0; c < 3; ++c)
                if (!((m >> c) & 1)) zf[g.dof(i, c, j, k)] += 0.125 * acc[c];
        }
        cl.sync();
        smooth(cl, g, L, sK0, sb, z[l], zo[l], a.nu, false, tid, nthr);
    }
}
