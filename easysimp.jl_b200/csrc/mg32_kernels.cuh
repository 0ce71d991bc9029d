// mg32_kernels.cuh -- the single-precision vector / transfer kernels of the FP32 V-cycle (multigrid32.cu) and the
// FP64 pieces of the outer CG that consume its result.  In a header so that the host self-test
// (selftest_kernel_host.cpp) compiles and runs the same source on the CPU.
// Prerequisites at the point of inclusion: GridDev, RedOut<>, grid_reduce<>() (ctx.cuh / grid_dev.h).
#pragma once

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
        int K = (int)(idx % gc.nzp);
        if (K > gc.nz) continue;  // pad entries of the node rows stay zero
        long long t = idx / gc.nzp;
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
    const int nJ = gc.ny + 1, nK = gc.nz + 1, nJK = nJ * nK, nzf = gf.nzp, nzc = gc.nzp;  // nK counts, nzf / nzc are row pitches
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
            const int o00 = J * nzc + K, o01 = o00 + (hasK ? 1 : 0), o10 = o00 + (hasJ ? nzc : 0), o11 = o10 + (hasK ? 1 : 0);
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

