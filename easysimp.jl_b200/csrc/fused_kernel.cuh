// fused_kernel.cuh -- the fused x-marching hex8 operator kernel and its butterflies (see hex8_structured.cu for the
// derivation).  Kept in a header so that the SAME source is compiled by nvcc into libeasysimp_b200.so and, with the
// shims of selftest_kernel_host.cpp (threads of a block as host threads, cp.async as plain copies), by the host compiler
// for the CPU self-test of the -m "not gpu" suite.
// Prerequisites at the point of inclusion: GridDev, RedOut<>, grid_reduce_idx<>() (ctx.cuh / grid_dev.h) and the
// mode-space blocks c_B (double) and c_Bf (float).
#pragma once

// structural non-zeros of the class blocks (class q = mode ^ (1<<component))
__device__ __host__ constexpr bool pat(int q, int c, int d) {
    // q=0,7: full; q=1,2,4: the component whose mode is the constant (translation) drops out;
    // q=3,5,6: the component whose mode is xi.eta.zeta decouples from the other two.
    return (q == 0 || q == 7)   ? true
           : (q == 1)           ? (c != 0 && d != 0)
           : (q == 2)           ? (c != 1 && d != 1)
           : (q == 4)           ? (c != 2 && d != 2)
           : (q == 3)           ? ((c == 2) == (d == 2))
           : (q == 5)           ? ((c == 1) == (d == 1))
                                : ((c == 0) == (d == 0));
}

// ---------------------------------------------------------------------------- butterflies
// forward 2-D transform of the 4 corner values of a (y,z) cell face:
// F[by + 2 bz] = sum_{sy,sz} sy^by sz^bz u(sy,sz);  uYZ: first index = y offset, second = z offset
template <typename T>
__device__ __forceinline__ void fwd2d(T u00, T u10, T u01, T u11, T (&F)[4]) {
    T a0 = u10 + u00, a1 = u10 - u00, b0 = u11 + u01, b1 = u11 - u01;
    F[0] = b0 + a0;
    F[1] = b1 + a1;
    F[2] = b0 - a0;
    F[3] = b1 - a1;
}
// inverse: y(sy,sz) = sum_b sy^by sz^bz A[b]
template <typename T>
__device__ __forceinline__ void inv2d(const T (&A)[4], T& y00, T& y10, T& y01, T& y11) {
    T pm0 = A[0] - A[2], pm1 = A[1] - A[3], pp0 = A[0] + A[2], pp1 = A[1] + A[3];
    y00 = pm0 - pm1;
    y10 = pm0 + pm1;
    y01 = pp0 - pp1;
    y11 = pp0 + pp1;
}

template <typename T>
__device__ __forceinline__ T mode_block(int q, int c, int d);
template <>
__device__ __forceinline__ double mode_block<double>(int q, int c, int d) {
    return c_B[q][c][d];
}
template <>
__device__ __forceinline__ float mode_block<float>(int q, int c, int d) {
    return c_Bf[q][c][d];
}
// w = K~ v for one cell; v, w indexed [c][m], m = bx + 2 (by + 2 bz)
template <typename T>
__device__ __forceinline__ void apply_blocks(const T (&v)[3][8], T (&w)[3][8]) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            T acc = T(0);
            bool first = true;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                if (pat(q, c, d)) {
                    if (first) {
                        acc = mode_block<T>(q, c, d) * v[d][q ^ (1 << d)];
                        first = false;
                    } else {
                        acc = fma(mode_block<T>(q, c, d), v[d][q ^ (1 << d)], acc);
                    }
                }
            }
            w[c][q ^ (1 << c)] = acc;
        }
    }
}

// ---------------------------------------------------------------------------- V2: fused x-march
// Block = TZ lanes (z, fastest axis) x TY rows (y).  Thread (ty,tz) owns the cell column
// (Y0+ty, Z0+tz) and marches along x through [Lstart, Lend).  Per step it
//   1. reads the 4 nodes of the new plane from the shared input tile (3 comps) -> Fn (12 values)
//   2. v = E (Fn +- Fp)  (E folded into the x butterfly), w = K~ v  (45 FMAs, constant operands)
//   3. Gp = w0 + w1 -> next plane, Gm = w0 - w1 -> this plane; A = Gp(previous cell) + Gm
//   4. inverse 2-D butterfly -> 4 corner contributions; corner (0,0) stays in registers, the other
//      three go through shared memory to the threads that own those nodes
//   5. one __syncthreads per plane; node totals are stored (owner tiles) or red.add'ed (atomic tiles)
// The INPUT MUST BE ZERO ON PRESCRIBED DOFS (op_apply guarantees it): no mask is read here, so the
// plane prefetched two steps ahead sits in registers untouched until it is staged (ncu r1: the
// in-kernel mask select made every prefetch a long-scoreboard stall).
// Shared output tiles carry a zero border (row 0 / column 0) so that the gather needs no guards.
// ATOMIC=false: tiles overlap by one cell row/column, a tile owns the nodes interior to it and the
//   perimeter cells are computed twice: deterministic, no zero-fill, ~(1-1/TY)(1-1/TZ) efficient.
// ATOMIC=true : tiles do not overlap; perimeter nodes are combined with atomicAdd into a y whose
//   perimeter rows/columns were zeroed first (ncu r1: 2x slower than owner tiles on B200).
template <typename T, int TY, int TZ, int NF = 0>
struct FusedSmem {
    T in[4][3][TY + 1][TZ + 1];      // u of a node plane: 4-deep cp.async ring
    T E[4][TY][TZ];                  // modulus of the cell layer consumed together with that plane
    T out[2][3][3][TY + 1][TZ + 1];  // corner contributions c01, c10, c11 at [ty+1][tz+1]; borders stay 0
    T fz[NF > 0 ? NF : 1][3][TY][TZ];  // fused-update operands (r | r, Dinv, d) of the plane the next flush completes;
                                            // every thread only touches its own slot, so one buffer and no barrier suffice
};

#ifndef ESIMP_HOST_EMULATION
// 8-byte asynchronous global->shared copy (LDGSTS); !valid zero-fills the destination
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc, bool valid) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    int sz = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc, bool valid) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    int sz = valid ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}
#else
// host emulation (selftest_kernel_host.cpp): either the copy happens at once, or -- deferred mode -- it is queued per
// thread and only lands when cp_async_wait<N> retires its group, the LATEST moment the hardware may complete it.  A
// kernel that is right in both modes reads no asynchronously copied datum before its wait_group + barrier.
inline void cp_async8(void* smem_dst, const void* gsrc, bool valid) { emu_cp_async(smem_dst, gsrc, 8, valid); }
inline void cp_async4(void* smem_dst, const void* gsrc, bool valid) { emu_cp_async(smem_dst, gsrc, 4, valid); }
inline void cp_async_commit() { emu_cp_commit(); }
template <int N>
inline void cp_async_wait() {
    emu_cp_wait(N);
}
#endif
// one scalar of the kernel's arithmetic type
__device__ __forceinline__ void cp_async_scalar(double* smem_dst, const double* gsrc, bool valid) { cp_async8(smem_dst, gsrc, valid); }
__device__ __forceinline__ void cp_async_scalar(float* smem_dst, const float* gsrc, bool valid) { cp_async4(smem_dst, gsrc, valid); }

// MODE selects what the store phase does with s = (A u)[dof] on the nodes this thread owns:
//   FUSE_PLAIN  y = s                                  FUSE_DOT     y = s and <u, y> (PCG)
//   FUSE_SMOOTH one Chebyshev-Jacobi step of the multigrid smoother with u = z:
//               v = c1 d + c2 Dinv (r - s);  d = v;  y = z + v      (Dinv is 0 on prescribed dofs)
//               The previous direction d need not be a stored vector (FuseArgsT::dmode): after the first step from a zero
//               guess d IS the iterate z (FUSE_D_IS_Z), later it is z - z_prev (FUSE_D_FROM_ZPREV: z_prev is the input
//               buffer of the previous step, which this step overwrites in place -- every thread reads its own node of
//               z_prev before it stores there); write_d = 0 then drops the store of d as well: 32-40 B/dof instead of 48
//   FUSE_RESID  y = r - s                              (residual that is restricted to the next level)
// r, Dinv, d of the plane are prefetched into registers at the start of the step (before the FP64 work), so
// their latency is covered; the operator is FP64-bound, the extra bytes ride for free.
enum { FUSE_PLAIN = 0, FUSE_DOT = 1, FUSE_SMOOTH = 2, FUSE_RESID = 3, FUSE_SMOOTH_DOT = 4 };  // 4: FUSE_SMOOTH and <r, y>
template <typename T>
struct FuseArgsT {
    const T* r = nullptr;
    const T* dinv = nullptr;
    T* d = nullptr;
    T c1 = T(0), c2 = T(0);
    int dmode = FUSE_D_BUFFER;   // where the previous Chebyshev direction comes from (ignored when c1 == 0)
    int write_d = 1;             // store the new direction into d
    const T* zprev = nullptr;    // FUSE_D_FROM_ZPREV: the iterate before z (may alias the output vector)
    const T* sinv = nullptr;     // sweep kernel only: D^-1 as one double per node (sweep_kernel.cuh, sw_dinv); this kernel reads dinv
};
using FuseArgs = FuseArgsT<double>;
// T = double: the operator of the FP64 solve.  T = float: the same kernel for the single-precision multigrid cycle
// (multigrid32.cu); the fused dot products always accumulate in double.
template <typename T, int TY, int TZ, bool ATOMIC, int MINB, int MODE>
__global__ void __launch_bounds__(TY* TZ, MINB) k_matvec_fused(GridDev g, int le0, int le1, int ps0, int ps1, int chunk,
                                                               const T* __restrict__ E, const T* __restrict__ u,
                                                               T* __restrict__ y, double* partial, unsigned int* ticket,
                                                               RedOut<1> dot_out, FuseArgsT<T> fa) {
    constexpr bool DOT = MODE == FUSE_DOT || MODE == FUSE_SMOOTH_DOT;
    constexpr bool SMOOTH = MODE == FUSE_SMOOTH || MODE == FUSE_SMOOTH_DOT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NF = SMOOTH ? 3 : (MODE == FUSE_RESID ? 1 : 0);
    // previous Chebyshev direction: a stored vector, z - z_prev, or (neither flag, c1 != 0) the iterate itself; c1 == 0: unused
    const bool d_buffer = SMOOTH && fa.c1 != T(0) && fa.dmode == FUSE_D_BUFFER, d_zprev = SMOOTH && fa.c1 != T(0) && fa.dmode == FUSE_D_FROM_ZPREV;
    FusedSmem<T, TY, TZ, NF>& sm = *reinterpret_cast<FusedSmem<T, TY, TZ, NF>*>(smem_raw);
    const int tid = threadIdx.x;
    const int tz = tid % TZ, ty = tid / TZ;
    constexpr int SY = ATOMIC ? TY : TY - 1, SZ = ATOMIC ? TZ : TZ - 1;
    const int Y0 = blockIdx.y * SY - (ATOMIC ? 0 : 1), Z0 = blockIdx.x * SZ - (ATOMIC ? 0 : 1);
    const int ey = Y0 + ty, ez = Z0 + tz;  // my cell column == my node (low corner)
    const bool cell_ok = ey >= 0 && ey < g.ny && ez >= 0 && ez < g.nz;
    const bool node_ok = ey >= 0 && ey <= g.ny && ez >= 0 && ez <= g.nz;
    // halo node of the input tile handled by this thread (row TY: tid 0..TZ, column TZ: the next TY)
    const bool has_halo = tid < TZ + 1 + TY;
    const int hj = tid <= TZ ? TY : tid - TZ - 1, hk = tid <= TZ ? tid : TZ;
    const int hy = Y0 + hj, hz = Z0 + hk;
    const bool halo_ok = has_halo && hy >= 0 && hy <= g.ny && hz >= 0 && hz <= g.nz;

    // node planes this block owns and the cell layers that touch them
    const int P0 = ps0 + blockIdx.z * chunk, P1 = min(P0 + chunk, ps1);
    const int Lstart = max(P0 - g.e2p - 1, le0), Lend = min(P1 - g.e2p, le1);
    double pq[1] = {0.0};  // <in, out> over the nodes this thread stores (fused PCG dot product)
    if (Lstart < Lend) {   // always true for a valid slab (every owned plane touches a cell layer)
    const int nsteps = Lend - Lstart;
    const int pl0 = Lstart + g.e2p;  // first node plane read
    const long long plane = g.plane, plane3 = 3 * g.plane;

    // running pointers (strength-reduced addressing: one add per plane)
    const T* pu = u + (long long)pl0 * plane3 + (node_ok ? (long long)ey * g.nzp + ez : 0);
    const T* ph = u + (long long)pl0 * plane3 + (halo_ok ? (long long)hy * g.nzp + hz : 0);
    const T* pE = E + (long long)Lstart * g.eplane + (cell_ok ? (long long)ey * g.nz + ez : 0);
    T* py = y + (long long)pl0 * plane3 + (node_ok ? (long long)ey * g.nzp + ez : 0);
    T* pyh = y + (long long)pl0 * plane3 + (halo_ok ? (long long)hy * g.nzp + hz : 0);
    const bool own_store = node_ok && ty >= 1 && tz >= 1;
    long long foff = (long long)pl0 * plane3 + (node_ok ? (long long)ey * g.nzp + ez : 0);  // same offset as py

    // zero the output tiles once (their borders are never written again)
    for (int i = tid; i < 2 * 9 * (TY + 1) * (TZ + 1); i += TY * TZ) (&sm.out[0][0][0][0][0])[i] = T(0);

    // issue the asynchronous copies of plane pl0+k (and of the modulus of layer Lstart+k-1, which is
    // consumed in the same step) into ring slot k&3; always commits one group so that wait_group counts
    auto issue_plane = [&](int k) {
        if (k <= nsteps) {
            const int buf = k & 3;
#pragma unroll
            for (int c = 0; c < 3; ++c) cp_async_scalar(&sm.in[buf][c][ty][tz], pu + c * plane, node_ok);
            if (has_halo) {
#pragma unroll
                for (int c = 0; c < 3; ++c) cp_async_scalar(&sm.in[buf][c][hj][hk], ph + c * plane, halo_ok);
            }
            if (k >= 1) {
                cp_async_scalar(&sm.E[buf][ty][tz], pE, cell_ok);
                pE += g.eplane;
            }
            pu += plane3;
            ph += plane3;
        }
        cp_async_commit();
    };
    auto face_modes = [&](int buf, T (&F)[3][4]) {
#pragma unroll
        for (int c = 0; c < 3; ++c)
            fwd2d(sm.in[buf][c][ty][tz], sm.in[buf][c][ty + 1][tz], sm.in[buf][c][ty][tz + 1], sm.in[buf][c][ty + 1][tz + 1], F[c]);
    };
    auto put_corners = [&](int buf, int c, T y01, T y10, T y11) {
        sm.out[buf][0][c][ty + 1][tz + 1] = y01;  // -> node (ty, tz+1)
        sm.out[buf][1][c][ty + 1][tz + 1] = y10;  // -> node (ty+1, tz)
        sm.out[buf][2][c][ty + 1][tz + 1] = y11;  // -> node (ty+1, tz+1)
    };
    // combine the corner contributions of the plane py points at, write the node totals, advance py
    auto flush_plane = [&](int buf, int ring, bool store, const T (&c00)[3]) {
        if (store) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                T s = c00[c] + sm.out[buf][0][c][ty + 1][tz] + sm.out[buf][1][c][ty][tz + 1] + sm.out[buf][2][c][ty][tz];
                if (own_store) {
                    if (SMOOTH) {
                        const T rv = sm.fz[0][c][ty][tz];
                        const T zin = sm.in[ring][c][ty][tz];  // the input plane is still in its ring slot
                        const T dprev = d_buffer ? sm.fz[2 % NF][c][ty][tz] : (d_zprev ? zin - sm.fz[2 % NF][c][ty][tz] : zin);
                        const T v = fma(fa.c1, dprev, fa.c2 * sm.fz[1 % NF][c][ty][tz] * (rv - s));
                        if (fa.write_d) (fa.d + foff)[c * plane] = v;
                        const T znew = zin + v;
                        py[c * plane] = znew;
                        if (MODE == FUSE_SMOOTH_DOT) pq[0] = fma((double)rv, (double)znew, pq[0]);  // <r, z>: the PCG scalar of the next iteration
                    } else if (MODE == FUSE_RESID) {
                        py[c * plane] = sm.fz[0][c][ty][tz] - s;
                    } else {
                        py[c * plane] = s;
                    }
                    if (MODE == FUSE_DOT) pq[0] = fma((double)s, (double)sm.in[ring][c][ty][tz], pq[0]);
                } else if (ATOMIC && node_ok) {
                    atomicAdd(&py[c * plane], s);
                }
                if (ATOMIC && halo_ok) {
                    // top row (hj == TY): cells (TY-1, hk) and (TY-1, hk-1); right column: (hj, TZ-1), (hj-1, TZ-1)
                    T h = (hj == TY) ? ((hk < TZ ? sm.out[buf][1][c][TY][hk + 1] : T(0)) + sm.out[buf][2][c][TY][hk])
                                          : (sm.out[buf][0][c][hj + 1][TZ] + sm.out[buf][2][c][hj][TZ]);
                    atomicAdd(&pyh[c * plane], h);
                }
            }
        }
        py += plane3;
        pyh += plane3;
        foff += plane3;
    };
    // asynchronous copy of the fused-update operands of the plane the NEXT flush completes (joins the commit group
    // of the following issue_plane; FUSE_SMOOTH / FUSE_RESID wait for all groups before the flush)
    auto issue_fuse = [&](bool store) {
        if (MODE >= FUSE_SMOOTH && store && own_store) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                cp_async_scalar(&sm.fz[0][c][ty][tz], fa.r + foff + c * plane, true);
                if (SMOOTH) {
                    cp_async_scalar(&sm.fz[1 % NF][c][ty][tz], fa.dinv + foff + c * plane, true);
                    if (d_buffer || d_zprev) cp_async_scalar(&sm.fz[2 % NF][c][ty][tz], (d_zprev ? fa.zprev : fa.d) + foff + c * plane, true);
                }
            }
        }
    };
    // ---- prologue: planes pl0 .. pl0+3 in flight, the first two landed
    issue_fuse(pl0 >= P0 && pl0 < P1);  // operands of the first plane a flush completes (group of plane 0)
    issue_plane(0);
    issue_plane(1);
    issue_plane(2);
    cp_async_wait<1>();
    __syncthreads();
    T Fp[3][4], Gpp[3][4];
    face_modes(0, Fp);
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int b = 0; b < 4; ++b) Gpp[c][b] = T(0);
    issue_plane(3);

    int lp = pl0;  // plane the current step completes
    // one marching step; BUF = it & 1 is a compile-time constant (loop unrolled by two below) so that
    // every shared-memory access is [thread base + immediate]
    auto step = [&](auto bufc, int it) {
        constexpr int BUF = decltype(bufc)::value;
        const int ring = (it + 1) & 3;  // slot of plane pl0+it+1 and of this layer's modulus
        const T Ee = sm.E[ring][ty][tz];
        T Fn[3][4];
        face_modes(ring, Fn);
        // x butterfly with the modulus folded in: v0 = E (Fn + Fp), v1 = E (Fn - Fp)
        T v[3][8], w[3][8];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                T ep = Ee * Fp[c][b];
                v[c][2 * b] = fma(Ee, Fn[c][b], ep);
                v[c][2 * b + 1] = fma(Ee, Fn[c][b], -ep);
            }
        apply_blocks(v, w);
        T c00[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            T A[4];
#pragma unroll
            // b = 0: w[c][0] is the translation mode, structurally zero (written out so that no 0 -+ x is issued)
            A[0] = Gpp[c][0] - w[c][1];
            Gpp[c][0] = w[c][1];
            Fp[c][0] = Fn[c][0];
#pragma unroll
            for (int b = 1; b < 4; ++b) {
                A[b] = Gpp[c][b] + (w[c][2 * b] - w[c][2 * b + 1]);
                Gpp[c][b] = w[c][2 * b] + w[c][2 * b + 1];
                Fp[c][b] = Fn[c][b];
            }
            T y10, y01, y11;
            inv2d(A, c00[c], y10, y01, y11);
            put_corners(BUF, c, y01, y10, y11);
        }
        if (MODE >= FUSE_SMOOTH)
            cp_async_wait<0>();  // also the operand tile issued after the previous flush
        else
            cp_async_wait<1>();  // plane pl0+it+2 has landed (only the newest group may be pending)
        __syncthreads();
        flush_plane(BUF, it & 3, lp >= P0 && lp < P1, c00);
        ++lp;
        issue_fuse(lp >= P0 && lp < P1);  // operands of the next plane, after this thread has consumed its slot
        issue_plane(it + 4);              // into the slot of plane pl0+it, which nobody reads any more
    };
    int it = 0;
#pragma unroll 1
    for (; it + 1 < nsteps; it += 2) {
        step(std::integral_constant<int, 0>{}, it);
        step(std::integral_constant<int, 1>{}, it + 1);
    }
    if (it < nsteps) step(std::integral_constant<int, 0>{}, it);
    // ---- epilogue: the last plane only receives Gp of the last cell
    {
        T c00[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            T y10, y01, y11;
            inv2d(Gpp[c], c00[c], y10, y01, y11);
            put_corners(nsteps & 1, c, y01, y10, y11);
        }
        cp_async_wait<0>();
        __syncthreads();
        flush_plane(nsteps & 1, nsteps & 3, lp >= P0 && lp < P1, c00);
    }
    }  // Lstart < Lend
    if (DOT) {
        const unsigned int nblk = gridDim.x * gridDim.y * gridDim.z;
        const unsigned int bid = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
        grid_reduce_idx<1, false>(pq, partial, ticket, dot_out, bid, nblk);
    }
}

