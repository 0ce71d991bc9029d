// multigrid.cu -- geometric-multigrid preconditioner for the FP64 CG solve ("next" row).
//
// The reference solves K u = f with a fresh CHOLMOD factorisation every iteration (Optimization.jl:317);
// the matrix-free replacement is CG, and with a Jacobi preconditioner a 512-cell-long cantilever needs
// ~5 000 iterations per solve.  This file adds a V-cycle preconditioner built ONLY from pieces that
// already exist for the fine grid:
//   * level l is the same structured grid with every dimension halved l times; its operator is the SAME
//     fused hex8 kernel (hex8_structured.cu) with a coarse modulus field.  For cuboids K0(2h) = 2 K0(h), so
//     the factor is folded into the modulus: E'_{l+1} = 2 * mean_8(E'_l) = 1/4 * sum of the 8 children
//     (rediscretisation; equals the Galerkin operator P'AP wherever E is constant inside a coarse cell);
//   * P = trilinear interpolation, R = P' (full weighting), Dirichlet masks injected from the fine grid;
//   * smoother = Chebyshev-accelerated Jacobi on [lmax/mg_ratio, lmax], lmax(D^-1 A) from a few power
//     iterations per level and SIMP iteration; pre- and post-smoothing use the same polynomial, so the
//     V-cycle is symmetric positive definite and plain PCG applies;
//   * coarsest level: mg_coarse_its Chebyshev iterations.
//   * with the default operator variant the Chebyshev update z' = z + c1 d + c2 Dinv (r - A z), the residual
//     r - A z and the <r, z> of the outer CG are fused into the store phase of the operator kernel
//     (FUSE_SMOOTH / FUSE_RESID / FUSE_SMOOTH_DOT modes); k_cheb_step / k_residual serve the other variants.
// x-slab partition: level l of rank r is the slab [ex0/2^l, ex1/2^l); the same one-plane node halo and
// one-layer cell halo (comm.cu) are used on every level.  Levels with at most mg_replicate_cells cells are
// REPLICATED on every rank instead (one allreduce of the restricted residual at the transition, no
// communication below): deep levels are latency bound, not bandwidth bound.
// Prototype + convergence study: scripts/gmg_prototype.py (27-58 iterations where Jacobi-PCG needs 500-1200).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>

#include <chrono>

#include <atomic>

#include "ctx.cuh"
#include "coarse_tail.cuh"

#define LAUNCH_STREAM(kern, ...)                                        \
    do {                                                                \
        kern<<<RED_BLOCKS, RED_THREADS, 0, ctx->stream>>>(__VA_ARGS__); \
        ctx->launches++;                                                \
        ESIMP_CUDA(cudaGetLastError());                                 \
    } while (0)

// ---------------------------------------------------------------------------- kernels
__global__ void k_restrict_E(GridDev gf, GridDev gc, int ce0, int ce1, const double* __restrict__ Ef, double* __restrict__ Ec) {
    long long n = (long long)(ce1 - ce0) * gc.eplane;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
        int K = (int)(idx % gc.nz);
        long long t = idx / gc.nz;
        int J = (int)(t % gc.ny);
        int Le = ce0 + (int)(t / gc.ny);
        int lef = 2 * (gc.ge0 + Le) - gf.ge0;
        double s = 0.0;
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b)
#pragma unroll
                for (int c = 0; c < 2; ++c) s += Ef[gf.cell(lef + a, 2 * J + b, 2 * K + c)];
        Ec[gc.cell(Le, J, K)] = 0.25 * s;
    }
}

// r_c = P' res_f on the owned coarse nodes (27-point full weighting), 0 on prescribed coarse dofs
// coarse planes a thread marches through (one extra fine plane is read per chunk): long chunks where bandwidth counts, single
// planes on the small levels where only the latency of the launch does
static inline int restrict_chunk(long long coarse_dofs) { return coarse_dofs >= 400000 ? 8 : 1; }
__global__ void __launch_bounds__(256) k_restrict(GridDev gf, GridDev gc, int cp0, int cp1, int chunk, const double* __restrict__ resf,
                                                  const uint8_t* __restrict__ maskc, double* __restrict__ rc) {
    // One thread per coarse (J, K) position and component; it MARCHES along x through a chunk of coarse planes (blockIdx.y walks the
    // (chunk, component) pairs), so every fine plane is read once: the y-z weighted sum of the odd fine plane 2I+1 is carried from
    // coarse plane I to I+1 (6 loads of 16 bytes per coarse dof instead of 9; the round-2 launch list had the 9-load version at
    // 0.44 ms = 0.32 of the HBM roofline on the fine grid).  The x dimension of the grid covers the (J, K) positions of a plane incl.
    // the row pads (32-bit index arithmetic only), lanes along K.  Along z the three fine values 2K-1, 2K, 2K+1 come from ONE aligned
    // 16-byte load (2K, 2K+1) plus the second half of the left neighbour lane's pair (the pad entries behind a fine row are zero, so
    // 2K+1 past the last node contributes nothing).
    const int nJK = (gc.ny + 1) * gc.nzp;
    const int nJK_round = (nJK + 31) / 32 * 32;  // whole warps: the shuffles need every lane
    const int lane = threadIdx.x & 31;
    const int nchunk = (cp1 - cp0 + chunk - 1) / chunk;
    for (int pc = blockIdx.y; pc < nchunk * 3; pc += gridDim.y) {
        const int c = pc % 3;
        const int L0 = cp0 + (pc / 3) * chunk, L1 = min(L0 + chunk, cp1);
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nJK_round; i += gridDim.x * blockDim.x) {
            const bool in_range = i < nJK;
            const int id = in_range ? i : nJK - 1;
            const int J = id / gc.nzp, K = id - J * gc.nzp;
            const bool node = in_range && K <= gc.nz;     // pad entries of the coarse rows stay zero
            const bool left_in_lane = lane > 0 && K > 0;  // lane - 1 holds (J, K - 1) of the same row
            const int kf = 2 * (K <= gc.nz ? K : gc.nz);  // pad threads read a valid address, their result is dropped
            // y-z weighted sum of local fine plane lpf (global Gf) around (2J, 2K); 0 for a plane outside the grid / the slab
            auto yz_sum = [&](int Gf, int lpf) -> double {
                const bool xok = !(Gf < 0 || Gf > gf.nx || lpf < 0 || lpf >= gf.NPX);
                const double* __restrict__ fplane = resf + (long long)((xok ? lpf : 0) * 3 + c) * gf.plane;
                double acc = 0.0;
#pragma unroll
                for (int dy = -1; dy <= 1; ++dy) {
                    const int jf = 2 * J + dy;
                    const bool ok = xok && jf >= 0 && jf <= gf.ny;
                    const double* row = fplane + (ok ? jf : 2 * J) * gf.nzp;
                    double2 v = make_double2(0.0, 0.0);
                    if (ok) v = *reinterpret_cast<const double2*>(row + kf);
                    double left = __shfl_up_sync(0xffffffffu, v.y, 1);
                    if (!left_in_lane) left = (ok && K > 0) ? row[kf - 1] : 0.0;
                    acc = fma(dy ? 0.5 : 1.0, v.x + 0.5 * (v.y + left), acc);
                }
                return acc;
            };
            const int Gf0 = 2 * (gc.gp0 + L0), lpf0 = Gf0 - gf.gp0;
            double carry = yz_sum(Gf0 - 1, lpf0 - 1);
            for (int Lp = L0; Lp < L1; ++Lp) {
                const int Gf = 2 * (gc.gp0 + Lp), lpf = Gf - gf.gp0;
                const double mid = yz_sum(Gf, lpf), hi = yz_sum(Gf + 1, lpf + 1);
                const double acc = fma(0.5, carry, fma(0.5, hi, mid));
                carry = hi;
                if (node) {
                    const uint8_t m = maskc[(long long)Lp * gc.plane + id];
                    rc[(long long)(Lp * 3 + c) * gc.plane + id] = ((m >> c) & 1) ? 0.0 : acc;
                }
            }
        }
    }
}

// z_f += P z_c on the owned fine nodes (trilinear interpolation), prescribed fine dofs untouched.
// One thread per coarse (J,K) position of a (plane, component): it loads its 2x2 coarse parents (x 1 or 2 coarse
// planes) once and updates the 2x2 patch of fine nodes (2J+a, 2K+b) -- 4-8 loads for 4 outputs instead of 8 each.
// Along each axis the parents of fine index i are i>>1 and (i+1)>>1 with weight 1/2 each (the same node if i is even).
__global__ void __launch_bounds__(256) k_prolong_add(GridDev gf, GridDev gc, int fp0, int fp1, const double* __restrict__ zc,
                                                     const uint8_t* __restrict__ maskf, double* __restrict__ zf) {
    const int nJ = gc.ny + 1, nK = gc.nz + 1, nJK = nJ * nK, nzf = gf.nzp, nzc = gc.nzp;  // nK counts, nzf / nzc are row pitches
    for (int pc = blockIdx.y; pc < (fp1 - fp0) * 3; pc += gridDim.y) {
        const int lp = fp0 + pc / 3, c = pc % 3;
        const int Gf = gf.gp0 + lp;
        const bool two = Gf & 1;  // odd fine plane: average of two coarse planes
        const double* __restrict__ c0 = zc + (long long)(((Gf >> 1) - gc.gp0) * 3 + c) * gc.plane;
        const double* __restrict__ c1 = zc + (long long)((((Gf + 1) >> 1) - gc.gp0) * 3 + c) * gc.plane;
        double* __restrict__ zrow = zf + (long long)(lp * 3 + c) * gf.plane;
        const uint8_t* __restrict__ mrow = maskf + (long long)lp * gf.plane;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nJK; i += gridDim.x * blockDim.x) {
            const int J = i / nK, K = i - J * nK;
            const bool hasJ = J < gc.ny, hasK = K < gc.nz;  // the last coarse row / column only owns the even fine index
            const int o00 = J * nzc + K, o01 = o00 + (hasK ? 1 : 0), o10 = o00 + (hasJ ? nzc : 0), o11 = o10 + (hasK ? 1 : 0);
            double v00 = c0[o00], v01 = c0[o01], v10 = c0[o10], v11 = c0[o11];
            if (two) {
                v00 = 0.5 * (v00 + c1[o00]);
                v01 = 0.5 * (v01 + c1[o01]);
                v10 = 0.5 * (v10 + c1[o10]);
                v11 = 0.5 * (v11 + c1[o11]);
            }
            // the fine pair (2K, 2K+1) of a row is one aligned 16-byte read-modify-write (the row pitch is a multiple of 8 doubles;
            // behind the last node, 2K+1 is a pad entry: it gets += 0 and stays zero)
            const int f00 = (2 * J) * nzf + 2 * K;
            {
                double2* p = reinterpret_cast<double2*>(zrow + f00);
                double2 z2 = *p;
                if (!((mrow[f00] >> c) & 1)) z2.x += v00;
                if (hasK && !((mrow[f00 + 1] >> c) & 1)) z2.y += 0.5 * (v00 + v01);
                *p = z2;
            }
            if (hasJ) {
                const int f10 = f00 + nzf;
                double2* p = reinterpret_cast<double2*>(zrow + f10);
                double2 z2 = *p;
                if (!((mrow[f10] >> c) & 1)) z2.x += 0.5 * (v00 + v10);
                if (hasK && !((mrow[f10 + 1] >> c) & 1)) z2.y += 0.25 * ((v00 + v01) + (v10 + v11));
                *p = z2;
            }
        }
    }
}

// Chebyshev-Jacobi: first step from a zero guess: d = Dinv r / theta, z = d
// (d = nullptr: the fused smoother takes the first direction from z itself, FUSE_D_IS_Z, and d is never stored)
__global__ void k_cheb_first(const double* __restrict__ r, const double* __restrict__ dinv, double inv_theta, double* __restrict__ d,
                             double* __restrict__ z, long long i0, long long i1) {
    for (long long i = i0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < i1; i += (long long)gridDim.x * blockDim.x) {
        double v = dinv[i] * r[i] * inv_theta;
        if (d) d[i] = v;
        z[i] = v;
    }
}
// general step with q = A z:  d = c1 d + c2 Dinv (r - q);  z += d      (c1 = 0 restarts the recurrence)
__global__ void k_cheb_step(const double* __restrict__ r, const double* __restrict__ q, const double* __restrict__ dinv, double c1,
                            double c2, double* __restrict__ d, double* __restrict__ z, long long i0, long long i1) {
    for (long long i = i0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < i1; i += (long long)gridDim.x * blockDim.x) {
        double v = c2 * dinv[i] * (r[i] - q[i]);
        if (c1 != 0.0) v = fma(c1, d[i], v);
        d[i] = v;
        z[i] += v;
    }
}
// q = r - q
__global__ void k_residual(const double* __restrict__ r, double* __restrict__ q, long long i0, long long i1) {
    for (long long i = i0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < i1; i += (long long)gridDim.x * blockDim.x)
        q[i] = r[i] - q[i];
}
// power iteration helpers
__global__ void k_pi_init(GridDev g, const uint8_t* __restrict__ mask, double* __restrict__ x, long long i0, long long i1) {
    for (long long i = i0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < i1; i += (long long)gridDim.x * blockDim.x) {
        unsigned long long h = (unsigned long long)i * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull;
        h ^= h >> 29;
        h *= 0xBF58476D1CE4E5B9ull;
        h ^= h >> 32;
        int c = (int)((i / g.plane) % 3);
        long long nn = (i / (3 * g.plane)) * g.plane + i % g.plane;
        double v = (double)(h >> 11) * (1.0 / 9007199254740992.0) - 0.5;
        x[i] = (((mask[nn] >> c) & 1) || (int)(i % g.nzp) > g.nz) ? 0.0 : v;  // prescribed dofs and row pads stay zero
    }
}
__global__ void k_pi_scale(const double* __restrict__ q, const double* __restrict__ dinv, double* __restrict__ y, long long i0,
                           long long i1, double* partial, unsigned int* ticket, RedOut<1> out) {
    double acc[1] = {0.0};
    for (long long i = i0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < i1; i += (long long)gridDim.x * blockDim.x) {
        double v = dinv[i] * q[i];
        y[i] = v;
        acc[0] = fma(v, v, acc[0]);
    }
    grid_reduce<1, false>(acc, partial, ticket, out);
}
__global__ void k_pi_normalize(const double* __restrict__ y, const double* __restrict__ nrm2, double* __restrict__ x, long long i0,
                               long long i1) {
    const double s = rsqrt(*nrm2);
    for (long long i = i0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < i1; i += (long long)gridDim.x * blockDim.x)
        x[i] = y[i] * s;
}
// PCG pieces
__global__ void k_mg_init(const double* __restrict__ f, const double* __restrict__ q, double* __restrict__ r, long long i0, long long i1,
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
__global__ void k_mg_update(const double* __restrict__ q, double* __restrict__ r, long long i0, long long i1,
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
// same update plus the first (zero-guess) Chebyshev step of the level-0 pre-smoother on the new residual:
// d = z = Dinv r / theta  (saves the separate k_cheb_first pass over r and Dinv)
__global__ void k_mg_update_first(const double* __restrict__ q, double* __restrict__ r, const double* __restrict__ dinv,
                                  double inv_theta, double* __restrict__ d, double* __restrict__ z, long long i0, long long i1,
                                  const double* __restrict__ s_rz, const double* __restrict__ s_pq, double* partial,
                                  unsigned int* ticket, RedOut<1> out) {
    const double alpha = *s_rz / *s_pq;
    double acc[1] = {0.0};
    for (long long i = i0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < i1; i += (long long)gridDim.x * blockDim.x) {
        double ri = fma(-alpha, q[i], r[i]);
        r[i] = ri;
        acc[0] = fma(ri, ri, acc[0]);
        double v = dinv[i] * ri * inv_theta;
        if (d) d[i] = v;  // nullptr with the fused smoother (FUSE_D_IS_Z): 40 B/dof instead of 48
        z[i] = v;
    }
    grid_reduce<1, false>(acc, partial, ticket, out);
}
// x += alpha p;  p = z + beta p   (beta = 0 on the first call: p = z)
__global__ void k_mg_direction(const double* __restrict__ z, double* __restrict__ p, double* __restrict__ x, long long i0, long long i1,
                               const double* __restrict__ s_rz_new, const double* __restrict__ s_rz_old,
                               const double* __restrict__ s_pq, int mode) {
    // mode 0: both updates; 1: first call (p = z); 2: converged, only the deferred x update
    const double beta = mode == 0 ? *s_rz_new / *s_rz_old : 0.0, alpha = mode == 1 ? 0.0 : *s_rz_old / *s_pq;
    for (long long i = i0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < i1; i += (long long)gridDim.x * blockDim.x) {
        const double pi = p[i];
        if (mode != 1) x[i] = fma(alpha, pi, x[i]);
        if (mode != 2) p[i] = mode == 1 ? z[i] : fma(beta, pi, z[i]);
    }
}

// ---------------------------------------------------------------------------- hierarchy
void mg_free(esimp_ctx* ctx) {
    for (size_t l = 0; l < ctx->mg.size(); ++l) {
        MGLevel& L = ctx->mg[l];
        cudaFree(L.d_z);
        cudaFree(L.d_sinv);
        cudaFree(L.d_z2);
        cudaFree(L.d_d);
        cudaFree(L.d_pi);
        float* f32[] = {L.f_E, L.f_dinv, L.f_r, L.f_z, L.f_z2, L.f_d, L.f_q};
        for (float* p : f32) cudaFree(p);
        if (l > 0) {
            cudaFree(L.d_E);
            cudaFree(L.d_mask);
            cudaFree(L.d_fixed);
            cudaFree(L.d_dinv);
            cudaFree(L.d_r);
            cudaFree(L.d_q);
        }
    }
    ctx->mg.clear();
}

static int mg_build_levels(esimp_ctx* ctx) {
    mg_free(ctx);
    const int R = ctx->nranks, r = ctx->rank;
    int nx = ctx->g.nx, ny = ctx->g.ny, nz = ctx->g.nz;
    // level 0 aliases the fine grid
    {
        MGLevel L;
        L.vw = fine_view(ctx);
        L.has_left = ctx->has_left;
        L.has_right = ctx->has_right;
        L.ndof_loc = ctx->ndof_loc;
        L.ncell_loc = ctx->ncell_loc;
        L.own_dof0 = ctx->own_dof0;
        L.own_dof1 = ctx->own_dof1;
        L.own_cell0 = ctx->own_cell0;
        L.own_cell1 = ctx->own_cell1;
        L.n_dofs_global = ctx->n_dofs_global;
        L.d_E = ctx->d_E;
        L.d_mask = ctx->d_mask;
        L.d_fixed = (long long*)ctx->d_fixed;
        L.n_fixed = ctx->n_fixed;
        L.d_dinv = ctx->d_dinv;
        L.d_r = ctx->d_r;
        L.d_q = ctx->d_q;
        ESIMP_CUDA(cudaMalloc(&L.d_z, sizeof(double) * L.ndof_loc));
        ESIMP_CUDA(cudaMalloc(&L.d_d, sizeof(double) * L.ndof_loc));
        ESIMP_CUDA(cudaMalloc(&L.d_pi, sizeof(double) * L.ndof_loc));
        ESIMP_CUDA(cudaMalloc(&L.d_z2, sizeof(double) * L.ndof_loc));
        ESIMP_CUDA(cudaMemset(L.d_z2, 0, sizeof(double) * L.ndof_loc));
        ESIMP_CUDA(cudaMemset(L.d_z, 0, sizeof(double) * L.ndof_loc));
        ESIMP_CUDA(cudaMemset(L.d_d, 0, sizeof(double) * L.ndof_loc));
        ctx->mg.push_back(L);
    }
    bool replicated = false;
    for (int l = 1; l < 14; ++l) {
        if ((nx & 1) || (ny & 1) || (nz & 1)) break;
        const int cx = nx / 2, cy = ny / 2, cz = nz / 2;
        // Slabs: a level stays distributed while every rank keeps >= mg_min_layers whole cell layers and the level is
        // big enough to be worth its halo exchanges; below that the level (and every coarser one) is REPLICATED: each
        // rank holds the whole small grid, the restricted residual is summed with one allreduce, and nothing under
        // that level communicates (the deep levels of a slab hierarchy are pure NCCL latency otherwise).
        if (R > 1 && !replicated) {
            const bool can_split = cx % R == 0 && cx / R >= ctx->mg_min_layers;
            if (!can_split || (long long)cx * cy * cz <= ctx->mg_replicate_cells) {
                if (!ctx->mg_replicate) {
                    if (!can_split) break;
                } else {
                    replicated = true;
                    // a LARGE level replicated only because its layers do not divide by the rank count: correct, but every rank then
                    // smooths the whole level and the residual sum is a big allreduce -- say so once (nx a multiple of 2^k * ranks avoids it)
                    if (!can_split && (long long)cx * cy * cz > ctx->mg_replicate_cells && r == 0)
                        fprintf(stderr,
                                "[esimp] multigrid level %d (%dx%dx%d cells) cannot be cut into %d x-slabs of whole layers: it and all coarser levels "
                                "are replicated on every GPU (slower; choose nx as a multiple of %d)\n",
                                l, cx, cy, cz, R, 2 * R << (l - 1));
                }
            }
        }
        nx = cx;
        ny = cy;
        nz = cz;
        const bool rp = replicated;
        const int ex0 = rp ? 0 : (int)((long long)nx * r / R), ex1 = rp ? nx : (int)((long long)nx * (r + 1) / R);
        const int left = (!rp && ctx->has_left) ? 1 : 0, right = (!rp && ctx->has_right) ? 1 : 0;
        MGLevel L;
        L.replicated = rp;
        GridDev& g = L.vw.g;
        g.nx = nx;
        g.ny = ny;
        g.nz = nz;
        g.gp0 = ex0 - left;
        g.NPX = (ex1 - ex0) + 1 + left;
        g.ge0 = ex0 - left;
        g.NEX = (ex1 - ex0) + left + right;
        g.e2p = g.ge0 - g.gp0;
        g.nzp = esimp_node_pitch(nz);
        g.plane = (long long)(ny + 1) * g.nzp;
        g.eplane = (long long)ny * nz;
        L.vw.own_p0 = left;
        L.vw.own_p1 = g.NPX - right;
        L.vw.mv_e0 = 0;  // owned layers + the redundant left interface layer
        L.vw.mv_e1 = left + (ex1 - ex0);
        L.has_left = left != 0;
        L.has_right = right != 0;
        L.ndof_loc = 3LL * g.NPX * g.plane;
        L.ncell_loc = (long long)g.NEX * g.eplane;
        L.own_dof0 = 3LL * L.vw.own_p0 * g.plane;
        L.own_dof1 = 3LL * L.vw.own_p1 * g.plane;
        L.own_cell0 = (long long)left * g.eplane;
        L.own_cell1 = (long long)(left + ex1 - ex0) * g.eplane;
        L.n_dofs_global = 3LL * (nx + 1) * (ny + 1) * (nz + 1);
        // injected Dirichlet mask: coarse node (I,J,K) is fine node (I,J,K) << l
        std::vector<uint8_t> mask((size_t)g.NPX * g.plane, 0);
        std::vector<long long> fixed;
        const long long fnx = ctx->g.nx + 1, fny = ctx->g.ny + 1;
        const int sh = (1 << l) - 1;
        for (int64_t q : ctx->prescribed_nc) {
            int c = (int)(q % 3);
            int64_t nref = q / 3;
            int i = (int)(nref % fnx), j = (int)((nref / fnx) % fny), k = (int)(nref / (fnx * fny));
            if ((i & sh) || (j & sh) || (k & sh)) continue;
            int lp = (i >> l) - g.gp0;
            if (lp < 0 || lp >= g.NPX) continue;
            uint8_t& m = mask[(size_t)g.node(lp, j >> l, k >> l)];
            if (!((m >> c) & 1) && lp >= L.vw.own_p0 && lp < L.vw.own_p1) fixed.push_back(g.dof(lp, c, j >> l, k >> l));
            m |= (uint8_t)(1 << c);
        }
        L.n_fixed = (long long)fixed.size();
        ESIMP_CUDA(cudaMalloc(&L.d_mask, mask.size()));
        ESIMP_CUDA(cudaMemcpy(L.d_mask, mask.data(), mask.size(), cudaMemcpyHostToDevice));
        ESIMP_CUDA(cudaMalloc(&L.d_fixed, sizeof(long long) * std::max<long long>(L.n_fixed, 1)));
        if (L.n_fixed) ESIMP_CUDA(cudaMemcpy(L.d_fixed, fixed.data(), sizeof(long long) * L.n_fixed, cudaMemcpyHostToDevice));
        ESIMP_CUDA(cudaMalloc(&L.d_E, sizeof(double) * L.ncell_loc));
        double** vecs[] = {&L.d_dinv, &L.d_r, &L.d_z, &L.d_z2, &L.d_d, &L.d_q, &L.d_pi};
        for (double** v : vecs) {
            ESIMP_CUDA(cudaMalloc(v, sizeof(double) * L.ndof_loc));
            ESIMP_CUDA(cudaMemset(*v, 0, sizeof(double) * L.ndof_loc));
        }
        L.vw.E = L.d_E;
        L.vw.mask = L.d_mask;
        ctx->mg.push_back(L);
    }
    return ESIMP_OK;
}

// While a replicated level is processed the context behaves like a single-rank one: every halo exchange and
// allreduce is skipped (all ranks execute the same deterministic kernels on the same data).
struct LocalScope {
    esimp_ctx* c;
    int saved;
    LocalScope(esimp_ctx* ctx, bool on) : c(ctx), saved(ctx->nranks) {
        if (on) c->nranks = 1;
    }
    ~LocalScope() { c->nranks = saved; }
};

static int mg_apply(esimp_ctx* ctx, MGLevel& L, double* in, double* out, int dot_slot = -1) {
    if (ctx->nranks > 1) ESIMP_TRY(comm_halo_nodes_lv(ctx, L.vw.g, L.vw.own_p0, L.vw.own_p1, L.has_left, L.has_right, in));
    ESIMP_TRY(hex8_matvec_view(ctx, L.vw, in, out, ctx->prm.matvec_variant, dot_slot));
    return fixup_lv(ctx, L.d_fixed, L.n_fixed, L.diag_value, in, out);
}

// lmax(D^-1 A) by power iteration (x in d_z, A x in d_q, D^-1 A x in d_d)
static int mg_estimate_lmax(esimp_ctx* ctx, MGLevel& L, int iters) {
    if (L.pi_valid) {  // warm start from the dominant vector of the previous SIMP iteration: a few steps suffice
        ESIMP_CUDA(cudaMemcpyAsync(L.d_z, L.d_pi, sizeof(double) * L.ndof_loc, cudaMemcpyDeviceToDevice, ctx->stream));
        iters = 2;
    } else {
        LAUNCH_STREAM(k_pi_init, L.vw.g, L.d_mask, L.d_z, L.own_dof0, L.own_dof1);
    }
    RedOut<1> out{{ctx->d_scal + SC_TMP0}};
    double lam = 4.0;
    for (int it = 0; it < iters; ++it) {
        ESIMP_TRY(mg_apply(ctx, L, L.d_z, L.d_q));
        LAUNCH_STREAM(k_pi_scale, L.d_q, L.d_dinv, L.d_d, L.own_dof0, L.own_dof1, ctx->d_partial, ctx->d_ticket, out);
        ESIMP_TRY(allreduce_scalars(ctx, SC_TMP0, 1, false));
        LAUNCH_STREAM(k_pi_normalize, L.d_d, ctx->d_scal + SC_TMP0, L.d_z, L.own_dof0, L.own_dof1);
    }
    // one more product with the normalised vector: ||D^-1 A x|| / ||x|| with ||x|| = 1
    ESIMP_TRY(mg_apply(ctx, L, L.d_z, L.d_q));
    LAUNCH_STREAM(k_pi_scale, L.d_q, L.d_dinv, L.d_d, L.own_dof0, L.own_dof1, ctx->d_partial, ctx->d_ticket, out);
    ESIMP_TRY(allreduce_scalars(ctx, SC_TMP0, 1, false));
    ESIMP_CUDA(cudaMemcpyAsync(L.d_pi, L.d_z, sizeof(double) * L.ndof_loc, cudaMemcpyDeviceToDevice, ctx->stream));
    L.pi_valid = true;
    ESIMP_TRY(fetch_scalars(ctx));
    lam = sqrt(ctx->h_scal[SC_TMP0]);
    if (!(lam > 0.5) || !(lam < 100.0)) {
        lam = 4.0;
        L.pi_valid = false;
    }
    L.lmax = 1.1 * lam;
    return ESIMP_OK;
}

static void mg_plan_tail(esimp_ctx* ctx);
int mg_setup(esimp_ctx* ctx) {
    if (!ctx->structured) return esimp_fail(ctx, ESIMP_ERR_STATE, "the multigrid preconditioner needs a structured grid");
    if (ctx->mg.empty() || ctx->mg_bc_version != ctx->bc_version) {
        ESIMP_TRY(mg_build_levels(ctx));
        ctx->mg_bc_version = ctx->bc_version;
        mg_plan_tail(ctx);
    }
    ctx->mg[0].diag_value = ctx->diag_value;
    auto node_scalar = [&](MGLevel& L) -> int {  // D^-1 in the form the sweep kernel's smoother modes read (levels it serves only)
        if (!hex8_sweep_eligible(ctx, L.vw)) return ESIMP_OK;
        if (!L.d_sinv) {
            ESIMP_CUDA(cudaMalloc(&L.d_sinv, sizeof(double) * L.vw.g.NPX * L.vw.g.plane));
            ESIMP_CUDA(cudaMemsetAsync(L.d_sinv, 0, sizeof(double) * L.vw.g.NPX * L.vw.g.plane, ctx->stream));
        }
        return hex8_node_scalar_dinv(ctx, L.vw, L.d_dinv, L.d_sinv);
    };
    ESIMP_TRY(node_scalar(ctx->mg[0]));
    for (size_t l = 1; l < ctx->mg.size(); ++l) {
        MGLevel &F = ctx->mg[l - 1], &C = ctx->mg[l];
        if (C.replicated && !F.replicated) {
            // distributed -> replicated: a rank restricts the coarse layers I whose first child layer 2I it owns (the
            // second child may be its right ghost layer), the rest of the full-grid array is summed in by an allreduce
            const int fe0 = F.vw.g.ge0 + (int)(F.own_cell0 / F.vw.g.eplane), fe1 = F.vw.g.ge0 + (int)(F.own_cell1 / F.vw.g.eplane);
            ESIMP_CUDA(cudaMemsetAsync(C.d_E, 0, sizeof(double) * C.ncell_loc, ctx->stream));
            LAUNCH_STREAM(k_restrict_E, F.vw.g, C.vw.g, (fe0 + 1) / 2, (fe1 + 1) / 2, F.d_E, C.d_E);
            ESIMP_TRY(comm_allreduce(ctx, C.d_E, C.ncell_loc, false));
        } else {
            LocalScope ls(ctx, C.replicated);
            const int left = C.has_left ? 1 : 0;
            LAUNCH_STREAM(k_restrict_E, F.vw.g, C.vw.g, left, left + (int)((C.own_cell1 - C.own_cell0) / C.vw.g.eplane), F.d_E, C.d_E);
            if (ctx->nranks > 1)
                ESIMP_TRY(comm_halo_cells_lv(ctx, C.vw.g, left, left + (int)((C.own_cell1 - C.own_cell0) / C.vw.g.eplane), 1,
                                             C.has_left, C.has_right, C.d_E));
        }
        LocalScope ls(ctx, C.replicated);
        ESIMP_TRY(hex8_diagonal_view(ctx, C.vw, C.d_dinv, C.n_dofs_global, &C.diag_value));
        ESIMP_TRY(node_scalar(C));
    }
    for (auto& L : ctx->mg) {
        LocalScope ls(ctx, L.replicated);
        ESIMP_TRY(mg_estimate_lmax(ctx, L, 8));
    }
    if (getenv("ESIMP_MG_DEBUG"))
        for (size_t l = 0; l < ctx->mg.size(); ++l) {
            const MGLevel& L = ctx->mg[l];
            fprintf(stderr, "[mg rank %d] level %zu%s: %dx%dx%d NPX %d NEX %d own_p [%d,%d) mv_e [%d,%d) n_fixed %lld diag_value %.6e lmax %.4f\n", ctx->rank,
                    l, L.replicated ? " (replicated)" : "", L.vw.g.nx, L.vw.g.ny, L.vw.g.nz, L.vw.g.NPX, L.vw.g.NEX, L.vw.own_p0, L.vw.own_p1, L.vw.mv_e0, L.vw.mv_e1, L.n_fixed,
                    L.diag_value, L.lmax);
        }
    return ESIMP_OK;
}

// the operator kernels fuse the Chebyshev update / the residual into their store phase
static bool mg_fused(const esimp_ctx* ctx) {
    const int v = ctx->prm.matvec_variant;
    return v == ESIMP_MATVEC_AUTO || v == ESIMP_MATVEC_FUSED || v == ESIMP_MATVEC_SWEEP;  // SWEEP by name: every store mode of the sweep kernel (tests)
}
// k Chebyshev-Jacobi iterations on level L for A z = r
// first_done: the zero-guess first step was already applied by the caller (k_mg_update_first)
// dot_slot >= 0: the last fused step also leaves <r, z> in d_scal[dot_slot]; *dot_done reports whether it did
static int mg_smooth(esimp_ctx* ctx, MGLevel& L, int k, bool zero_guess, bool first_done = false, int dot_slot = -1,
                     bool* dot_done = nullptr) {
    const double lmax = L.lmax, lmin = lmax / ctx->mg_ratio;
    const double theta = 0.5 * (lmax + lmin), delta = 0.5 * (lmax - lmin), sigma = theta / delta;
    double rho = 1.0 / sigma;
    int done = 0;
    const bool fused = mg_fused(ctx);
    if (zero_guess) {
        if (!first_done) LAUNCH_STREAM(k_cheb_first, L.d_r, L.d_dinv, 1.0 / theta, fused ? nullptr : L.d_d, L.d_z, L.own_dof0, L.own_dof1);
        done = 1;
    }
    for (int i = done; i < k; ++i) {
        if (!fused) ESIMP_TRY(mg_apply(ctx, L, L.d_z, L.d_q));
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
        if (fused) {
            // Chebyshev update fused into the operator's store: z' = z + c1 d + c2 Dinv (r - A z)
            if (ctx->nranks > 1) ESIMP_TRY(comm_halo_nodes_lv(ctx, L.vw.g, L.vw.own_p0, L.vw.own_p1, L.has_left, L.has_right, L.d_z));
            const bool with_dot = dot_slot >= 0 && i == k - 1;
            // the previous direction is never a stored vector: z itself right after the zero-guess first step, z - z_prev later
            // (z_prev = the input of the previous step = d_z2, which this step overwrites in place)
            const int dmode = i == 0 ? FUSE_D_NONE : (i == 1 && zero_guess ? FUSE_D_IS_Z : FUSE_D_FROM_ZPREV);
            ESIMP_TRY(hex8_smooth_view(ctx, L.vw, L.d_z, L.d_z2, L.d_r, L.d_dinv, L.d_d, c1, c2, with_dot ? dot_slot : -1, dmode, 0, L.d_z2, L.d_sinv));
            if (with_dot && dot_done) *dot_done = true;
            std::swap(L.d_z, L.d_z2);
        } else {
            LAUNCH_STREAM(k_cheb_step, L.d_r, L.d_q, L.d_dinv, c1, c2, L.d_d, L.d_z, L.own_dof0, L.own_dof1);
        }
    }
    return ESIMP_OK;
}


// which levels the single-kernel tail of the V-cycle takes: the longest run of trailing levels that are whole on this rank, small
// enough for one cluster and at most CT_MAX_LEVELS long; never the fine grid
static void mg_plan_tail(esimp_ctx* ctx) {
    ctx->mg_tail_start = 0;
    if (!ctx->mg_tail) return;
    const int n = (int)ctx->mg.size();
    int t = n;
    while (t - 1 >= 1 && n - (t - 1) <= CT_MAX_LEVELS) {
        const MGLevel& L = ctx->mg[t - 1];
        const GridDev& g = L.vw.g;
        const bool whole = (ctx->nranks == 1 || L.replicated) && g.gp0 == 0 && g.ge0 == 0 && g.e2p == 0 && g.NPX == g.nx + 1 && g.NEX == g.nx;
        if (!whole || (long long)(g.nx + 1) * (g.ny + 1) * (g.nz + 1) > std::min(CT_MAX_NODES, ctx->mg_tail_nodes)) break;
        if (coarse_tail_slab_doubles(g.nx, g.ny, g.nz) * 8 > 200 * 1024) break;
        --t;
    }
    if (t < n) ctx->mg_tail_start = t;
}
// z_l = V(r_l) for l = mg_tail_start, all levels below included, in one launch
static int mg_tail_cycle(esimp_ctx* ctx) {
    const int t = ctx->mg_tail_start, n = (int)ctx->mg.size();
    if (!ctx->d_K0) {
        ESIMP_CUDA(cudaMalloc(&ctx->d_K0, sizeof(double) * 576));
        ESIMP_CUDA(cudaMemcpyAsync(ctx->d_K0, ctx->K0, sizeof(double) * 576, cudaMemcpyHostToDevice, ctx->stream));
    }
    CoarseTailArgs a{};
    a.nlev = n - t;
    a.nu = ctx->mg_nu;
    a.coarse_its = ctx->mg_coarse_its;
    a.K0 = ctx->d_K0;
    for (int l = t; l < n; ++l) {
        MGLevel& L = ctx->mg[l];
        CoarseTailLevel& T = a.lv[l - t];
        T.nx = L.vw.g.nx;
        T.ny = L.vw.g.ny;
        T.nz = L.vw.g.nz;
        T.nzp = L.vw.g.nzp;
        T.E = L.d_E;
        T.mask = L.d_mask;
        T.dinv = L.d_dinv;
        T.r = L.d_r;
        T.za = L.d_z;
        T.zb = L.d_z2;
        T.q = L.d_q;
        const double lmax = L.lmax, lmin = lmax / ctx->mg_ratio;
        T.theta = 0.5 * (lmax + lmin);
        T.delta = 0.5 * (lmax - lmin);
    }
    long long sdbl = 0;
    for (int l = t; l < n; ++l) sdbl = std::max(sdbl, coarse_tail_slab_doubles(ctx->mg[l].vw.g.nx, ctx->mg[l].vw.g.ny, ctx->mg[l].vw.g.nz));
    static std::atomic<unsigned long long> attr_devices{0};
    const unsigned long long dev_bit = 1ull << (ctx->device & 63);
    if (!(attr_devices.load() & dev_bit)) {
        ESIMP_CUDA(cudaFuncSetAttribute(k_coarse_tail, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_devices.fetch_or(dev_bit);
    }
    k_coarse_tail<<<CT_CTAS, CT_THREADS, (size_t)sdbl * 8, ctx->stream>>>(a);
    ctx->launches++;
    ESIMP_CUDA(cudaGetLastError());
    // where the kernel left z of its first level: one buffer swap per smoother step that applies the operator
    const int swaps = a.nlev == 1 ? ctx->mg_coarse_its - 1 : (ctx->mg_nu - 1) + ctx->mg_nu;
    if (swaps & 1) std::swap(ctx->mg[t].d_z, ctx->mg[t].d_z2);
    return ESIMP_OK;
}

// z_l = V(r_l).  Level 0 only: first_done / dot_slot / dot_done as in mg_smooth.
static int mg_vcycle(esimp_ctx* ctx, int l, bool first_done = false, int dot_slot = -1, bool* dot_done = nullptr) {
    MGLevel& L = ctx->mg[l];
    if (l > 0 && l == ctx->mg_tail_start) return mg_tail_cycle(ctx);
    if (l == (int)ctx->mg.size() - 1)
        return mg_smooth(ctx, L, l == 0 ? ctx->mg_nu : ctx->mg_coarse_its, true, first_done, dot_slot, dot_done);
    MGLevel& C = ctx->mg[l + 1];
    ESIMP_TRY(mg_smooth(ctx, L, ctx->mg_nu, true, first_done));
    if (ctx->prm.matvec_variant == ESIMP_MATVEC_AUTO || ctx->prm.matvec_variant == ESIMP_MATVEC_FUSED ||
        ctx->prm.matvec_variant == ESIMP_MATVEC_SWEEP) {
        // residual fused into the operator's store: q = r - A z (prescribed rows are zeroed by the fix-up: z = 0 there)
        if (ctx->nranks > 1) ESIMP_TRY(comm_halo_nodes_lv(ctx, L.vw.g, L.vw.own_p0, L.vw.own_p1, L.has_left, L.has_right, L.d_z));
        ESIMP_TRY(hex8_resid_view(ctx, L.vw, L.d_z, L.d_q, L.d_r));
        ESIMP_TRY(fixup_lv(ctx, L.d_fixed, L.n_fixed, L.diag_value, L.d_z, L.d_q));
    } else {
        ESIMP_TRY(mg_apply(ctx, L, L.d_z, L.d_q));
        LAUNCH_STREAM(k_residual, L.d_r, L.d_q, L.own_dof0, L.own_dof1);
    }
    // PROF_MG_COARSE: from the exchange of the fine residual to the end of the prolongation, one pair per V-cycle
    const bool time_coarse = l == 0 && ctx->prof_on >= 2 && !ctx->prof_coarse_pending && ctx->prof_coarse_ev[0];
    if (time_coarse) cudaEventRecord(ctx->prof_coarse_ev[0], ctx->stream);
    if (ctx->nranks > 1) ESIMP_TRY(comm_halo_nodes_lv(ctx, L.vw.g, L.vw.own_p0, L.vw.own_p1, L.has_left, L.has_right, L.d_q));
    if (C.replicated && !L.replicated) {
        // distributed -> replicated: restrict onto the coarse planes I whose fine plane 2I this rank owns (their
        // neighbours 2I-1, 2I+1 are owned or ghost planes), sum the full coarse vector over the ranks, then run the
        // rest of the cycle without any communication; the prolongation back reads the local full coarse grid
        const int fp0g = L.vw.g.gp0 + L.vw.own_p0, fp1g = L.vw.g.gp0 + L.vw.own_p1;  // owned fine planes, global
        ESIMP_CUDA(cudaMemsetAsync(C.d_r, 0, sizeof(double) * C.ndof_loc, ctx->stream));
        {
            const int cp0 = (fp0g + 1) / 2, cp1 = (fp1g - 1) / 2 + 1;
            const int nJK = (C.vw.g.ny + 1) * C.vw.g.nzp;
            const int chunk = restrict_chunk((long long)nJK * 3 * (cp1 - cp0));
            dim3 grid((unsigned)std::min(64, (nJK + 255) / 256), (unsigned)std::max(1, std::min(65535, (cp1 - cp0 + chunk - 1) / chunk * 3)));
            k_restrict<<<grid, 256, 0, ctx->stream>>>(L.vw.g, C.vw.g, cp0, cp1, chunk, L.d_q, C.d_mask, C.d_r);
            ctx->launches++;
            ESIMP_CUDA(cudaGetLastError());
        }
        ESIMP_TRY(comm_allreduce(ctx, C.d_r, C.ndof_loc, false));
        LocalScope ls(ctx, true);
        ESIMP_TRY(mg_vcycle(ctx, l + 1));
    } else {
        {
            PROF_BEGIN(ctx, PROF_RESTRICT, l == 0)
            const int nJK = (C.vw.g.ny + 1) * C.vw.g.nzp;
            const int chunk = restrict_chunk((long long)nJK * 3 * (C.vw.own_p1 - C.vw.own_p0));
            dim3 grid((unsigned)std::min(64, (nJK + 255) / 256),
                      (unsigned)std::max(1, std::min(65535, (C.vw.own_p1 - C.vw.own_p0 + chunk - 1) / chunk * 3)));
            k_restrict<<<grid, 256, 0, ctx->stream>>>(L.vw.g, C.vw.g, C.vw.own_p0, C.vw.own_p1, chunk, L.d_q, C.d_mask, C.d_r);
            ctx->launches++;
            ESIMP_CUDA(cudaGetLastError());
            PROF_END(ctx)
        }
        ESIMP_TRY(mg_vcycle(ctx, l + 1));
        if (ctx->nranks > 1) ESIMP_TRY(comm_halo_nodes_lv(ctx, C.vw.g, C.vw.own_p0, C.vw.own_p1, C.has_left, C.has_right, C.d_z));
    }
    {
        const int nyz = (C.vw.g.ny + 1) * (C.vw.g.nz + 1);  // one thread per coarse (J,K) position
        dim3 grid((unsigned)std::min(64, (nyz + 255) / 256), (unsigned)std::min(65535, (L.vw.own_p1 - L.vw.own_p0) * 3));
        PROF_BEGIN(ctx, PROF_PROLONG, l == 0)
        k_prolong_add<<<grid, 256, 0, ctx->stream>>>(L.vw.g, C.vw.g, L.vw.own_p0, L.vw.own_p1, C.d_z, L.d_mask, L.d_z);
        PROF_END(ctx)
        ctx->launches++;
        ESIMP_CUDA(cudaGetLastError());
    }
    if (time_coarse) {
        cudaEventRecord(ctx->prof_coarse_ev[1], ctx->stream);
        ctx->prof_coarse_pending = true;
    }
    return mg_smooth(ctx, L, ctx->mg_nu, false, false, dot_slot, dot_done);
}

__global__ void k_dot2(const double* __restrict__ a, const double* __restrict__ b, long long i0, long long i1, double* partial,
                       unsigned int* ticket, RedOut<1> out) {
    double acc[1] = {0.0};
    for (long long i = i0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < i1; i += (long long)gridDim.x * blockDim.x)
        acc[0] = fma(a[i], b[i], acc[0]);
    grid_reduce<1, false>(acc, partial, ticket, out);
}

// CG preconditioned with one V-cycle per iteration.  Same contract as pcg_solve (vector_ops.cu).
struct HostPcgClock {  // PROF_HOST_PCG / PROF_HOST_WAIT: wall time of the solve and of the blocking scalar fetches
    esimp_ctx* ctx;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    double wait_ms = 0.0;
    int fetches = 0;
    explicit HostPcgClock(esimp_ctx* c) : ctx(c) {}
    ~HostPcgClock() {
        if (ctx->prof_on < 2) return;
        ctx->prof_ms_cat[PROF_HOST_PCG] += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        ctx->prof_cnt_cat[PROF_HOST_PCG] += 1;
        ctx->prof_ms_cat[PROF_HOST_WAIT] += wait_ms;
        ctx->prof_cnt_cat[PROF_HOST_WAIT] += fetches;
    }
};
int pcg_solve_mg(esimp_ctx* ctx, int* iters, double* relres) {
    if (ctx->mg.empty()) return esimp_fail(ctx, ESIMP_ERR_STATE, "multigrid hierarchy not built");
    HostPcgClock clock(ctx);
    MGLevel& L0 = ctx->mg[0];
    const long long i0 = ctx->own_dof0, i1 = ctx->own_dof1;
    double* S = ctx->d_scal;
    static const int RZ[2] = {SC_RZ0, SC_RZ1};
    ESIMP_TRY(mg_apply(ctx, L0, ctx->d_u, ctx->d_q));
    {
        RedOut<2> out{{S + SC_RR0, S + SC_FF}};
        LAUNCH_STREAM(k_mg_init, ctx->d_f, ctx->d_q, ctx->d_r, i0, i1, ctx->d_partial, ctx->d_ticket, out);
        ESIMP_TRY(allreduce_scalars(ctx, SC_RR0, 1, false));
        ESIMP_TRY(allreduce_scalars(ctx, SC_FF, 1, false));
    }
    ESIMP_TRY(fetch_scalars(ctx));
    const double ff = ctx->h_scal[SC_FF];
    ctx->f_norm2 = ff;
    const double tol2 = ctx->prm.pcg_rtol * ctx->prm.pcg_rtol * ff;
    double rr = ctx->h_scal[SC_RR0];
    int it = 0, status = ESIMP_OK;
    const bool lag = ctx->mg_lag > 0;  // opt-in (ESIMP_PCG_LAG=1): measured +2 % per SIMP step at N = 1, +1.7 % at N = 2
    if (ff == 0.0) {
        ESIMP_CUDA(cudaMemsetAsync(ctx->d_u, 0, sizeof(double) * ctx->ndof_loc, ctx->stream));
        rr = 0.0;
    }
    if (rr > tol2) {
        ESIMP_TRY(mg_vcycle(ctx, 0));  // z = M^-1 r
        RedOut<1> out{{S + RZ[0]}};
        LAUNCH_STREAM(k_dot2, ctx->d_r, L0.d_z, i0, i1, ctx->d_partial, ctx->d_ticket, out);
        ESIMP_TRY(allreduce_scalars(ctx, RZ[0], 1, false));
        LAUNCH_STREAM(k_mg_direction, L0.d_z, ctx->d_p, ctx->d_u, i0, i1, S + RZ[0], S + RZ[0], S + SC_PQ, 1);
    }
    while (rr > tol2 && it < ctx->prm.pcg_maxit) {
        const int cur = it & 1, nxt = cur ^ 1;
        // q = A p, <p,q>.  The V-cycle reuses d_q as scratch, so q is consumed by the update right away.
        if (hex8_matvec_fuses_dot(ctx)) {
            ESIMP_TRY(mg_apply(ctx, L0, ctx->d_p, ctx->d_q, SC_PQ));
            ESIMP_TRY(allreduce_scalars(ctx, SC_PQ, 1, false));
        } else {
            ESIMP_TRY(mg_apply(ctx, L0, ctx->d_p, ctx->d_q));
            ESIMP_TRY(vec_dot(ctx, ctx->d_p, ctx->d_q, SC_PQ));
        }
        {
            // r -= alpha q, <r,r>, and the first Chebyshev step of the coming V-cycle (d = z = Dinv r / theta)
            const double lmax0 = L0.lmax, theta0 = 0.5 * (lmax0 + lmax0 / ctx->mg_ratio);
            RedOut<1> out{{S + SC_RR0}};
{
                PROF_BEGIN(ctx, PROF_MG_UPDATE_FIRST, true)
                LAUNCH_STREAM(k_mg_update_first, ctx->d_q, ctx->d_r, L0.d_dinv, 1.0 / theta0, mg_fused(ctx) ? nullptr : L0.d_d, L0.d_z, i0, i1, S + RZ[cur], S + SC_PQ,
                          ctx->d_partial, ctx->d_ticket, out);
                PROF_END(ctx)
            }
            ESIMP_TRY(allreduce_scalars(ctx, SC_RR0, 1, false));
        }
        ++it;
        if (lag && it < ctx->prm.pcg_maxit) {
            // lagged poll: leave <r,r> of this iteration in the pinned ring and look at the previous iteration's instead
            double* ring = ctx->h_scal + N_SCAL;
            ESIMP_CUDA(cudaMemcpyAsync(ring + (it & 1), S + SC_RR0, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
            ESIMP_CUDA(cudaEventRecord(ctx->lag_ev[it & 1], ctx->stream));
            if (it >= 2) {
                const auto w0 = std::chrono::steady_clock::now();
                ESIMP_CUDA(cudaEventSynchronize(ctx->lag_ev[(it - 1) & 1]));
                clock.wait_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - w0).count();
                clock.fetches++;
                const double rr_prev = ring[(it - 1) & 1];
                if (!(rr_prev == rr_prev)) return esimp_fail(ctx, ESIMP_ERR_STATE, "PCG breakdown: residual is NaN");
                if (rr_prev <= tol2) {
                    // iteration it - 1 had converged: the direction kernel of that pass has already applied x += alpha_{it-1} p, and
                    // the update of this surplus iteration is deferred and never applied, so x is the iterate of iteration it - 1
                    rr = rr_prev;
                    --it;
                    break;
                }
                rr = rr_prev;  // not converged: rr only has to stay above tol2 for the loop condition
            }
        } else {
            {
                const auto w0 = std::chrono::steady_clock::now();
                ESIMP_TRY(fetch_scalars(ctx));
                clock.wait_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - w0).count();
                clock.fetches++;
            }
            rr = ctx->h_scal[SC_RR0];
            if (getenv("ESIMP_MG_DEBUG") && ctx->rank == 0)
                fprintf(stderr, "[mg] it %d rr %.6e pq %.6e rz %.6e\n", it, rr, ctx->h_scal[SC_PQ], ctx->h_scal[RZ[cur]]);
            if (!(rr == rr)) return esimp_fail(ctx, ESIMP_ERR_STATE, "PCG breakdown: residual is NaN");
            if (rr <= tol2 || it >= ctx->prm.pcg_maxit) {  // only the deferred x += alpha p is left to do
                LAUNCH_STREAM(k_mg_direction, L0.d_z, ctx->d_p, ctx->d_u, i0, i1, S + RZ[cur], S + RZ[cur], S + SC_PQ, 2);
                break;
            }
        }
        bool dot_done = false;
        ESIMP_TRY(mg_vcycle(ctx, 0, true, RZ[nxt], &dot_done));  // <r,z> comes out of the last fused smoother step
        if (!dot_done) {
            RedOut<1> out{{S + RZ[nxt]}};
            LAUNCH_STREAM(k_dot2, ctx->d_r, L0.d_z, i0, i1, ctx->d_partial, ctx->d_ticket, out);
        }
        ESIMP_TRY(allreduce_scalars(ctx, RZ[nxt], 1, false));
        {
            PROF_BEGIN(ctx, PROF_MG_DIRECTION, true)
            LAUNCH_STREAM(k_mg_direction, L0.d_z, ctx->d_p, ctx->d_u, i0, i1, S + RZ[nxt], S + RZ[cur], S + SC_PQ, 0);
            PROF_END(ctx)
        }
    }
    if (rr > tol2) status = ESIMP_WARN_PCG_MAXIT;
    if (iters) *iters = it;
    if (relres) *relres = ff > 0 ? sqrt(rr / ff) : 0.0;
    return status;
}
