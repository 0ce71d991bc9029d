// sweep_kernel.cuh -- the z-sweeping, TMA-fed, dataflow-synchronised version of the fused hex8 operator (sm_100a).
//
// Same arithmetic as fused_kernel.cuh (mode-space blocks, x-march with the modulus folded into the x butterfly), different
// machine mapping, built from what ncu said about the round-1 kernel (FP64 pipe 58 % busy, 1.27x redundant FP64 work on tile
// perimeters, 287 non-FP64 issue slots per 324 FP64 ones, one CTA barrier per plane):
//   * a CTA owns 15 node rows (16 cell rows: 8 warps, TWO stacked cells per thread) x ALL of z x one x chunk.  z is swept
//     in segments of 32 cell columns (lanes); the z-high face contribution of lane 31 is carried to lane 0 of the next
//     segment through shared memory, so z costs no recomputed perimeter at all (tile redundancy 16/15, not (16/15)^2);
//   * two cells per thread share the z butterflies of their common node row, the y exchange of that row stays in
//     registers, and every per-step overhead (barrier waits, constant loads, addressing) is paid once per two cells;
//   * node planes, modulus tiles and the fused-update operands arrive by TMA (cp.async.bulk.tensor 3-D boxes issued by ONE
//     thread of a producer warp, mbarrier complete_tx) -- no per-thread LDGSTS, no address arithmetic, out-of-range rows /
//     columns / cells are zero-filled by the TMA unit (E = 0 kills every contribution of a cell that does not exist);
//   * no CTA barrier in the loop: z neighbours exchange through warp shuffles, y neighbours (adjacent warps) through a
//     double-buffered shared row with a full / empty mbarrier pair per warp, the input ring through full / empty mbarriers;
//   * the 45 mode-space constants are kernel parameters (constant bank 0: direct DFMA operands, and no process-global
//     __constant__ state shared between contexts).
// Needs the padded node-row pitch of grid_dev.h (16-byte TMA strides) and an even nz (modulus rows); hex8_structured.cu
// falls back to k_matvec_fused otherwise (coarse multigrid levels).
// Compiled for the host as well (selftest_kernel_host.cpp: threads of a block as host threads, mbarriers / TMA / shuffles
// emulated) so that the -m "not gpu" suite runs this very source against the oracle.
#pragma once

constexpr int SW_NW = 8;                // consumer warps; a thread owns the two stacked cells (rows 2w, 2w + 1) of a z column
constexpr int SW_CROWS = 2 * SW_NW;     // cell rows of a tile; a tile owns SW_CROWS - 1 node rows
constexpr int SW_ROWS = SW_CROWS + 1;   // node rows of an input box
constexpr int SW_BOXZ = 34;             // node columns of a box: 32 + the z-high neighbour + 1 (16-byte multiple)
constexpr int SW_PW = 4;                // producer warps: one whole warpgroup (setmaxnreg works per warpgroup); one thread of it issues
constexpr int SW_THREADS = 32 * (SW_NW + SW_PW);
constexpr int SW_IN_BOX = 3 * SW_ROWS * SW_BOXZ;                 // doubles in an input box {34, 17, 3}
constexpr int SW_IN_SLOT = (SW_IN_BOX * 8 + 127) / 128 * 16;     // slot pitch in doubles: TMA destinations are 128-byte aligned
constexpr int SW_OROWS = SW_CROWS - 1;                           // node rows a tile owns = rows of an operand box
constexpr int SW_OP_BOX = 3 * SW_OROWS * 32;                     // operand box {32, 15, 3}: exactly the owned nodes of a segment, 11520 B = 90 * 128
constexpr int SW_OPX_BOX = 3 * SW_OROWS * 2;                     // ... and {2, 15, 3} for the extra node column nz of the last segment (720 B)
constexpr int SW_OPX_SLOT = 96;                                  // its slot pitch in doubles (768 B = 6 * 128)
constexpr int SW_E_BOX = SW_CROWS * 32;                          // modulus box {32, 16, 1}

struct SweepBlocks {
    double B[8][3][3];
};

#ifndef ESIMP_HOST_EMULATION
typedef unsigned long long SwBar;
typedef CUtensorMap SwTensorMap;
__device__ __forceinline__ unsigned sw_smem(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sw_bar_init(SwBar* b, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sw_smem(b)), "r"(count)); }
__device__ __forceinline__ void sw_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void sw_bar_expect_tx(SwBar* b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sw_smem(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sw_bar_arrive(SwBar* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(sw_smem(b)) : "memory"); }
__device__ __forceinline__ void sw_bar_wait(SwBar* b, unsigned parity) {
    asm volatile(
        "{\n.reg .pred p;\nSW_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra SW_DONE_%=;\nbra SW_WAIT_%=;\nSW_DONE_%=:\n}\n" ::"r"(sw_smem(b)),
        "r"(parity)
        : "memory");
}
// the same on precomputed 32-bit shared addresses (hot loop: no generic -> shared conversion, no address arithmetic)
__device__ __forceinline__ void sw_bar_arrive_a(unsigned b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b) : "memory"); }
__device__ __forceinline__ void sw_bar_wait_a(unsigned b, unsigned parity) {
    asm volatile(
        "{\n.reg .pred p;\nSW_WAITA_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra SW_DONEA_%=;\nbra SW_WAITA_%=;\nSW_DONEA_%=:\n}\n" ::"r"(b),
        "r"(parity)
        : "memory");
}
typedef unsigned SwBarAddr;
__device__ __forceinline__ SwBarAddr sw_bar_addr(const SwBar* b) { return sw_smem(b); }
__device__ __forceinline__ SwBarAddr sw_bar_at(SwBarAddr base, int index) { return base + 8u * (unsigned)index; }
__device__ __forceinline__ void sw_tma_3d(void* dst, const SwTensorMap* tm, int c0, int c1, int c2, SwBar* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(sw_smem(dst)),
                 "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(sw_smem(bar))
                 : "memory");
}
// pull a box into L2 only (no shared-memory destination, no barrier): issued a few planes ahead of the real load, so that the
// DRAM latency under load is not bounded by the depth of the shared-memory ring
__device__ __forceinline__ void sw_tma_prefetch_3d(const SwTensorMap* tm, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(tm), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ double sw_shfl_up1(double v) {
    const int lo = __shfl_up_sync(0xffffffffu, __double2loint(v), 1), hi = __shfl_up_sync(0xffffffffu, __double2hiint(v), 1);
    return __hiloint2double(hi, lo);
}
// keep a thread-invariant value in its register: the compiler must not rematerialise it from %tid at every use
#define SW_PIN32(x) asm volatile("" : "+r"(x))
#define SW_PIN64(x) asm volatile("" : "+l"(x))
__device__ __forceinline__ void sw_syncwarp() { __syncwarp(); }
__device__ __forceinline__ long long sw_double_as_ll(double v) { return __double_as_longlong(v); }
__device__ __forceinline__ double sw_ll_as_double(long long v) { return __longlong_as_double(v); }
// register reallocation between the producer warpgroup and the two consumer warpgroups (168 per thread at launch)
__device__ __forceinline__ void sw_regs_producer() { asm volatile("setmaxnreg.dec.sync.aligned.u32 40;\n"); }
__device__ __forceinline__ void sw_regs_consumer() { asm volatile("setmaxnreg.inc.sync.aligned.u32 232;\n"); }
__device__ __forceinline__ void sw_consumer_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(32 * SW_NW) : "memory"); }
#define SW_SMEM_DECL extern __shared__ __align__(128) unsigned char sw_smem_raw[]
#endif

template <int NF, int RING>
struct SweepSmem {
    // XE: 2-slot y exchange with an explicit empty barrier, the shared memory it frees spent on a third operand slot.  Measured
    // at 512x256x256 (r2): smoother 1.39 -> 1.35 ms but residual 0.73 -> 0.80 ms (shorter x chunks, one more wait per plane), so
    // it stays off: the exchange ring is indexed like the input ring and needs no empty barrier, operands are double-buffered
    static constexpr bool XE = false;
    static constexpr int XN = XE ? 2 : RING, FN = NF > 0 ? (XE ? 3 : 2) : 1;
    alignas(128) double in[RING][SW_IN_SLOT];                             // node-plane boxes [c][17 rows][34]
    alignas(128) double E[RING][SW_E_BOX];                                // modulus boxes [16 rows][32]
    alignas(128) double fz[FN][NF > 0 ? NF : 1][NF > 0 ? SW_OP_BOX : 16];     // r | r, Dinv, d boxes [c][15 owned rows][32]
    alignas(128) double fzx[FN][NF > 0 ? NF : 1][NF > 0 ? SW_OPX_SLOT : 16];  // the same for node column nz: [c][15 rows][2]
    // y-high mode coefficients (bz = 0, 1) of warp w's upper cell for node row 2w + 2 (warp w + 1).  One slot per input-ring
    // slot, indexed like the plane it belongs to: warp w + 1 has read the slot before it releases that plane's input slot, and
    // warp w cannot be RING planes ahead of a plane that is not released, so no "empty" barrier is needed for this exchange
    double xch[XN][SW_NW][3][2][32];
    SwBar full[RING], empty[RING], ffull[FN], fempty[FN], xfull[XN][SW_NW], xempty[XN][SW_NW];
};
// the carry rows follow the struct: carry[ci][row][3] = z-high face sums of lane 31 (per node row, component), read by
// lane 0 one segment later; row ci = 0 .. nsteps (output planes of the chunk)
constexpr int SW_CARRY_ROW = SW_CROWS * 3;
// Store-phase operands read from shared memory BEFORE the y stage instead of next to their use (the consumer warps issue in order and
// nothing independent is left at the tail of a step, so a load placed there exposes its whole latency): bit 0 = u for <u, y>
// (FUSE_DOT), bit 1 = r / Dinv / z / z_prev of the lower node row (smoother modes), bit 2 = of the middle row as well.
// Measured inside the SIMP step at 512x256x256 (r2 call 23): 0 -> 3.068 it/s, 1 -> 3.075 (dot launch 0.628 -> 0.612 ms), 3 -> 3.110
// (smoother launch 0.838 -> 0.815 ms), 7 -> 2.852 (register spills).  Also tried and dropped: <u, y> gathered cell by cell in mode space
// (no second read of u, but 12 more live doubles: spills, 0.88 ms) and one accumulator per component (no change)
#ifndef SW_PRELOAD
#define SW_PRELOAD 3
#endif

template <int MODE>
struct SweepTraits {
    static constexpr bool DOT = MODE == FUSE_DOT || MODE == FUSE_SMOOTH_DOT;
    static constexpr bool SMOOTH = MODE == FUSE_SMOOTH || MODE == FUSE_SMOOTH_DOT;
    static constexpr int NF = SMOOTH ? 3 : (MODE == FUSE_RESID ? 1 : 0);
    static constexpr int RING = NF > 0 ? 4 : 5;
};

struct SweepArgs {
    GridDev g;
    int le0, le1, ps0, ps1, chunk;
    double* y;        // result vector (owned nodes of the owned planes)
    double* d_out;    // FUSE_SMOOTH: the Chebyshev direction is updated in place (read through the d tensor map)
    double c1, c2;
    double ik[3];     // FUSE_SMOOTH: 1 / K0[3a+c, 3a+c] -- Dinv[node, c] = ik[c] * (1 / sum of E around the node), see sw_dinv
    int dmode;        // FUSE_SMOOTH: where the previous Chebyshev direction comes from (FUSE_D_*; FUSE_D_NONE when c1 == 0)
    int write_d;      // FUSE_SMOOTH: store the new direction into d_out
    int pf_dist;      // producer: planes by which the L2 prefetch runs ahead of the shared-memory loads (0 = none)
    double* partial;
    unsigned int* ticket;
    RedOut<1> dot_out;
    SweepBlocks blk;
};

// w = K~ v with the blocks taken from kernel parameters
__device__ __forceinline__ void sw_apply_blocks(const SweepBlocks& blk, const double (&v)[3][8], double (&w)[3][8]) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            double acc = 0.0;
            bool first = true;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                if (pat(q, c, d)) {
                    if (first) {
                        acc = blk.B[q][c][d] * v[d][q ^ (1 << d)];
                        first = false;
                    } else {
                        acc = fma(blk.B[q][c][d], v[d][q ^ (1 << d)], acc);
                    }
                }
            }
            w[c][q ^ (1 << c)] = acc;
        }
    }
}

// The smoother's D^-1 is ONE double per node, not three: on a cuboid diag K[node, c] = K0[3a+c, 3a+c] * (sum of E over the cells
// around the node), so the operand box carries s = 1 / sum E and the three lowest mantissa bits of s say which components of the
// node are prescribed (Dinv = 0 there; 1e-15 relative on s is far below what a preconditioner notices).  8 B per node instead of
// 24: 5.3 B/dof less traffic in a kernel that runs at 80 % of the DRAM bandwidth.
__device__ __forceinline__ double sw_dinv(double sraw, int c, double ikc) {
    const long long bits = sw_double_as_ll(sraw);
    return ((bits >> c) & 1) ? 0.0 : sw_ll_as_double(bits & ~7ll) * ikc;
}

// one cell, one x step: Fn = y-z modes of the new plane from the z butterflies (zs, zd) of its two node rows; v = E (Fn +- Fp),
// w = K~ v, A = contributions to the node plane this step completes (y-z mode space), Gpp = contributions to the next one
__device__ __forceinline__ void sw_cell_step(const SweepBlocks& blk, const double (&zs0)[3], const double (&zd0)[3], const double (&zs1)[3],
                                             const double (&zd1)[3], double Ee, const double (&Fp)[3][4], double (&Fn)[3][4],
                                             double (&Gpp)[3][4], double (&A)[3][4]) {
    double v[3][8], wv[3][8];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        Fn[c][0] = zs1[c] + zs0[c];
        Fn[c][1] = zs1[c] - zs0[c];
        Fn[c][2] = zd1[c] + zd0[c];
        Fn[c][3] = zd1[c] - zd0[c];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const double ep = Ee * Fp[c][b];
            v[c][2 * b] = fma(Ee, Fn[c][b], ep);
            v[c][2 * b + 1] = fma(Ee, Fn[c][b], -ep);
        }
    }
    sw_apply_blocks(blk, v, wv);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        A[c][0] = Gpp[c][0] - wv[c][1];  // wv[c][0] is the translation mode, structurally zero
        Gpp[c][0] = wv[c][1];
#pragma unroll
        for (int b = 1; b < 4; ++b) {
            A[c][b] = Gpp[c][b] + (wv[c][2 * b] - wv[c][2 * b + 1]);
            Gpp[c][b] = wv[c][2 * b] + wv[c][2 * b + 1];
        }
    }
}

template <int MODE, int RING>
__global__ void __launch_bounds__(SW_THREADS, 1)
    k_matvec_sweep(const __grid_constant__ SweepArgs a, const __grid_constant__ SwTensorMap tmU, const __grid_constant__ SwTensorMap tmE,
                   const __grid_constant__ SwTensorMap tmR, const __grid_constant__ SwTensorMap tmDinv, const __grid_constant__ SwTensorMap tmD,
                   const __grid_constant__ SwTensorMap tmRx, const __grid_constant__ SwTensorMap tmDinvx, const __grid_constant__ SwTensorMap tmDx) {
    using TR = SweepTraits<MODE>;
    constexpr int NF = TR::NF, NFS = NF > 0 ? NF : 1;
    SW_SMEM_DECL;
    using SM = SweepSmem<NF, RING>;
    SM& sm = *reinterpret_cast<SM*>(sw_smem_raw);
    double* const carry = reinterpret_cast<double*>(sw_smem_raw + sizeof(SweepSmem<NF, RING>));
    const GridDev& g = a.g;
    int w = (int)(threadIdx.x >> 5) - SW_PW, lane = threadIdx.x & 31;  // w < 0: producer warpgroup, w = 0 .. 7: consumer warps
    SW_PIN32(w);
    SW_PIN32(lane);
    const int Y0 = blockIdx.x * (SW_CROWS - 1) - 1;
    const int P0 = a.ps0 + blockIdx.y * a.chunk, P1 = min(P0 + a.chunk, a.ps1);
    const int Lstart = max(P0 - g.e2p - 1, a.le0), Lend = min(P1 - g.e2p, a.le1);
    const int nsteps = Lend - Lstart, pl0 = Lstart + g.e2p;
    const int nseg = (g.nz + 31) / 32;
    const bool extra = (g.nz & 31) == 0;  // node column nz is nobody's own column: lane 31 of the last segment takes it along
    double pq[1] = {0.0};

    if (threadIdx.x == 0) {
        for (int s = 0; s < RING; ++s) {
            sw_bar_init(&sm.full[s], 1);
            sw_bar_init(&sm.empty[s], SW_NW);
        }
        for (int s = 0; s < SM::XN; ++s)
            for (int r = 0; r < SW_NW; ++r) {
                sw_bar_init(&sm.xfull[s][r], 1);
                sw_bar_init(&sm.xempty[s][r], 1);
            }
        for (int s = 0; s < SM::FN; ++s) {
            sw_bar_init(&sm.ffull[s], 1);
            sw_bar_init(&sm.fempty[s], SW_NW);
        }
        sw_fence_init();
    }
    __syncthreads();

    if (nsteps > 0) {
        if (w < 0) {
            // ------------------------------------------------------------------ producer: one thread feeds the rings
            sw_regs_producer();
            if (w == -SW_PW && lane == 0) {
                int slot = 0, round = 0;  // input ring position; round counts completed trips through the ring
                int t = 0;                // operand boxes issued
                auto issue_operands = [&](int s, int lp) {
                    if (NF > 0 && lp >= P0 && lp < P1) {
                        const int fs = t % SM::FN;
                        if (t >= SM::FN) sw_bar_wait(&sm.fempty[fs], ((t / SM::FN) - 1) & 1);
                        const int nb = NF == 1 ? 1 : ((a.dmode == FUSE_D_BUFFER || a.dmode == FUSE_D_FROM_ZPREV) ? 3 : 2);  // third box: d or z_prev
                        const bool xcol = extra && s == nseg - 1;  // the last segment also brings the operands of node column nz
                        // box 1 of the smoother modes is the node scalar of D^-1 (sw_dinv): one plane {32, 15, 1} at plane coordinate lp, a
                        // third of the bytes of the three-component boxes {32, 15, 3} at 3 lp
                        const int full = SW_OP_BOX + (xcol ? SW_OPX_BOX : 0);
                        sw_bar_expect_tx(&sm.ffull[fs], (unsigned)((NF == 1 ? full : (nb - 1) * full + full / 3) * 8));
                        const SwTensorMap* maps[6] = {&tmR, &tmDinv, &tmD, &tmRx, &tmDinvx, &tmDx};
                        for (int b = 0; b < nb; ++b) {
                            const int pz = (NF > 1 && b == 1) ? lp : 3 * lp;
                            sw_tma_3d(&sm.fz[fs][b % NFS][0], maps[b], 32 * s, Y0 + 1, pz, &sm.ffull[fs]);
                            if (xcol) sw_tma_3d(&sm.fzx[fs][b % NFS][0], maps[3 + b], g.nz, Y0 + 1, pz, &sm.ffull[fs]);
                        }
                        ++t;
                    }
                };
                const int per_seg = nsteps + 1, total = nseg * per_seg;
                auto prefetch_l2 = [&](int L) {  // the L-th plane this CTA loads (segments are swept one after the other)
                    if (L < total) {
                        const int ps = L / per_seg, pj = L - ps * per_seg;
                        sw_tma_prefetch_3d(&tmU, 32 * ps, Y0, 3 * (pl0 + pj));
                        if (pj >= 1) sw_tma_prefetch_3d(&tmE, 32 * ps, Y0, Lstart + pj - 1);
                    }
                };
                if (a.pf_dist > 0)
                    for (int L = RING; L < RING + a.pf_dist; ++L) prefetch_l2(L);
                for (int s = 0; s < nseg; ++s) {
                    for (int j = 0; j <= nsteps; ++j) {
                        if (a.pf_dist > 0) prefetch_l2(s * per_seg + j + RING + a.pf_dist);
                        if (round > 0) sw_bar_wait(&sm.empty[slot], (round - 1) & 1);
                        sw_bar_expect_tx(&sm.full[slot], (unsigned)((SW_IN_BOX + (j >= 1 ? SW_E_BOX : 0)) * 8));
                        sw_tma_3d(&sm.in[slot][0], &tmU, 32 * s, Y0, 3 * (pl0 + j), &sm.full[slot]);
                        if (j >= 1) sw_tma_3d(&sm.E[slot][0], &tmE, 32 * s, Y0, Lstart + j - 1, &sm.full[slot]);
                        if (++slot == RING) {
                            slot = 0;
                            ++round;
                        }
                        if (j >= 1) issue_operands(s, pl0 + j - 1);  // operands of the plane that step j - 1 completes
                    }
                    issue_operands(s, pl0 + nsteps);  // ... and of the plane the epilogue completes
                }
            }
        } else {
            // ------------------------------------------------------------------ consumers: warp w = cell rows Y0 + 2w, Y0 + 2w + 1
            sw_regs_consumer();
            int slot = 0, par = 0;      // input ring position and the parity its full barrier completes next
            int pslot = 0, ppar = 0;    // slot / parity of the previous plane (released after that plane's store)
            int t = 0, fs = 0, fpar = 0;  // operand boxes consumed; slot and parity of the next one
            int xcount = 0;             // y exchanges done (fused modes: 2-slot exchange with an empty barrier)
            const int ey0 = Y0 + 2 * w;  // node row of my lower cell's low corner; I store rows ey0 and ey0 + 1
            // thread-invariant values, computed once and pinned (the compiler would rematerialise them from %tid at every use)
            int row_ok0 = (w >= 1 && ey0 <= g.ny) ? 1 : 0;  // row 0 of a tile belongs to the tile below
            int row_ok1 = (ey0 + 1 <= g.ny) ? 1 : 0;
            int is_top = w == SW_NW - 1 ? 1 : 0, is_bottom = w == 0 ? 1 : 0;
            int thr_box = (2 * w) * SW_BOXZ + lane;   // my lower node inside an input / operand box plane
            int thr_e = (2 * w) * 32 + lane;          // my lower cell inside a modulus box
            SW_PIN32(row_ok0);
            SW_PIN32(row_ok1);
            SW_PIN32(is_top);
            SW_PIN32(is_bottom);
            SW_PIN32(thr_box);
            SW_PIN32(thr_e);
            const double* const in_thr = &sm.in[0][0] + thr_box;
            const double* const e_thr = &sm.E[0][0] + thr_e;
            const double* const fz_thr = &sm.fz[0][0][0] + (2 * w - 1) * 32 + lane;   // box row of node row r of mine: 2w - 1 + r
            const double* const fzx_thr = &sm.fzx[0][0][0] + (2 * w - 1) * 2;
            double* const xch_out = &sm.xch[0][w][0][0][0] + lane;                       // what I leave for the warp above
            const double* const xch_in = &sm.xch[0][w > 0 ? w - 1 : 0][0][0][0] + lane;  // what the warp below left for me
            constexpr int XCH_SLOT = SW_NW * 3 * 2 * 32;
            double* const carry_thr = carry + (2 * w) * 3;
            SwBarAddr b_full = sw_bar_addr(&sm.full[0]), b_empty = sw_bar_addr(&sm.empty[0]);
            SwBarAddr b_ffull = sw_bar_addr(&sm.ffull[0]), b_fempty = sw_bar_addr(&sm.fempty[0]);
            SwBarAddr b_xfull_me = sw_bar_addr(&sm.xfull[0][w]), b_xfull_lo = sw_bar_addr(&sm.xfull[0][w > 0 ? w - 1 : 0]);
            SwBarAddr b_xempty_me = sw_bar_addr(&sm.xempty[0][w]), b_xempty_lo = sw_bar_addr(&sm.xempty[0][w > 0 ? w - 1 : 0]);
            SW_PIN32(b_full);
            SW_PIN32(b_empty);
            SW_PIN32(b_xfull_me);
            SW_PIN32(b_xfull_lo);
            long long plane8 = g.plane, row8 = g.nzp;  // element strides of a component plane / a node row
            SW_PIN64(plane8);
            SW_PIN64(row8);
            // running pointers to (plane lp, component 0, row ey0, my column) of the result (and of the direction vector the smoother
            // updates): advanced by one node plane per step, so a store costs two adds instead of a full index computation
            double* yp = a.y;
            double* dp = a.d_out;

            // one output plane: AA / AB = y-z mode coefficients of the contributions of my two cells to node plane lp
            auto finish_plane = [&](const double (&AA)[3][4], const double (&AB)[3][4], int s, int lp, int ci) {
                {
                    // a plane outside [P0, P1) (the first one of a chunk that starts with the redundant layer) goes through the same
                    // exchange so that the slot / parity bookkeeping stays in step with the input ring; only its stores are skipped
                    const bool plane_ok = lp >= P0 && lp < P1;  // CTA-uniform
                    constexpr bool PRE_DOT = MODE == FUSE_DOT && (SW_PRELOAD & 1) != 0;
                    constexpr bool PRE_SM = TR::SMOOTH && (SW_PRELOAD & 6) != 0;
                    constexpr bool PRE_RS = MODE == FUSE_RESID && (SW_PRELOAD & 2) != 0;  // residual mode: r of both node rows (6 doubles)
                    constexpr int PRE_NROWS = (PRE_DOT || PRE_RS) ? 2 : (PRE_SM ? ((SW_PRELOAD & 4) ? 2 : 1) : 0);  // node rows r < PRE_NROWS are preloaded
                    double pu[2][3], pr[2][3], pz2[2][3], psn[2];  // psn: the node scalar of D^-1, one per node row
                    const bool last_seg = s == nseg - 1;
                    const bool ext = extra && last_seg && lane == 31;
                    const bool col_ok = plane_ok && 32 * s + lane <= g.nz;
                    constexpr bool XE = SM::XE;
                    const int xs = XE ? (xcount & 1) : pslot;
                    const int xpar = XE ? ((xcount >> 1) & 1) : ppar;
                    const double* uin = in_thr + pslot * SW_IN_SLOT;  // my nodes of plane lp (its slot is still held)
                    const double* fzp = fz_thr + fs * (NFS * SW_OP_BOX);
                    const double* fzxp = fzx_thr + fs * (NFS * SW_OPX_SLOT);
                    double* cr = carry_thr + ci * SW_CARRY_ROW;
                    // store one node: r = 0 / 1 my lower / upper node row, col = 0 my column, 1 the extra column nz
                    auto store_node = [&](int r, int c, double s_val, int col) {
                        const long long off = r * row8 + c * plane8 + col;
                        // operands of this node: my own column from the {32, 15, 3} box, the extra column from the {2, 15, 3} one
                        const double* op = col ? fzxp + c * SW_OROWS * 2 + r * 2 : fzp + c * SW_OROWS * 32 + r * 32;
                        const int ostride = col ? SW_OPX_SLOT : SW_OP_BOX;
                        const bool pre = col == 0 && r < PRE_NROWS;  // operands of this node were read before the y stage
                        if (TR::SMOOTH) {
                            const double rv = pre && PRE_SM ? pr[r][c] : op[0];
                            const double sraw = pre && PRE_SM ? psn[r] : (col ? fzxp[(1 % NFS) * SW_OPX_SLOT + r * 2] : fzp[(1 % NFS) * SW_OP_BOX + r * 32]);
                            const double dinv = sw_dinv(sraw, c, a.ik[c]);
                            const double zin = pre && PRE_SM ? pu[r][c] : uin[c * SW_ROWS * SW_BOXZ + r * SW_BOXZ + col];
                            const double third = pre && PRE_SM ? pz2[r][c] : op[(2 % NFS) * ostride];
                            const double dprev = a.dmode == FUSE_D_BUFFER ? third : a.dmode == FUSE_D_FROM_ZPREV ? zin - third : zin;  // FUSE_D_IS_Z / unused
                            const double v = fma(a.c1, dprev, a.c2 * dinv * (rv - s_val));
                            if (a.write_d) dp[off] = v;
                            const double znew = zin + v;
                            yp[off] = znew;
                            if (MODE == FUSE_SMOOTH_DOT) pq[0] = fma(rv, znew, pq[0]);
                        } else if (MODE == FUSE_RESID) {
                            yp[off] = (pre && PRE_RS ? pr[r][c] : op[0]) - s_val;
                        } else {
                            yp[off] = s_val;
                            if (MODE == FUSE_DOT) pq[0] = fma(s_val, pre && PRE_DOT ? pu[r][c] : uin[c * SW_ROWS * SW_BOXZ + r * SW_BOXZ + col], pq[0]);
                        }
                    };
                    // z stage of one node row: m[c][0 / 1] = z-even / z-odd coefficient (complete over y)
                    auto z_stage_store = [&](int r, const double (&m)[3][2], bool row_ok) {
                        double high[3], hin[3];
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            high[c] = m[c][0] + m[c][1];  // -> node column k + 1
                            hin[c] = sw_shfl_up1(high[c]);
                        }
                        if (lane == 0) {  // the z-low neighbour of lane 0 is lane 31 of the previous segment (zeros in segment 0)
#pragma unroll
                            for (int c = 0; c < 3; ++c) hin[c] = cr[r * 3 + c];
                        }
                        sw_syncwarp();  // lane 0 has read the carry of the previous segment before lane 31 overwrites it
                        if (lane == 31 && !last_seg) {
#pragma unroll
                            for (int c = 0; c < 3; ++c) cr[r * 3 + c] = high[c];
                        }
                        if (row_ok && col_ok) {
#pragma unroll
                            for (int c = 0; c < 3; ++c) store_node(r, c, (m[c][0] - m[c][1]) + hin[c], 0);
                        }
                        if (ext && row_ok && plane_ok) {
#pragma unroll
                            for (int c = 0; c < 3; ++c) store_node(r, c, high[c], 1);
                        }
                    };
                    if (NF > 0 && plane_ok) sw_bar_wait_a(sw_bar_at(b_ffull, fs), fpar);
                    if (PRE_DOT || PRE_SM || PRE_RS) {
#pragma unroll
                        for (int r = 0; r < PRE_NROWS; ++r)
#pragma unroll
                            for (int c = 0; c < 3; ++c) {
                                if (!PRE_RS) pu[r][c] = uin[c * SW_ROWS * SW_BOXZ + r * SW_BOXZ];
                                if (PRE_RS) pr[r][c] = fzp[c * SW_OROWS * 32 + r * 32];
                                if (PRE_SM) {
                                    const double* op = fzp + c * SW_OROWS * 32 + r * 32;
                                    pr[r][c] = op[0];
                                    if (c == 0) psn[r] = fzp[(1 % NFS) * SW_OP_BOX + r * 32];
                                    pz2[r][c] = op[(2 % NFS) * SW_OP_BOX];
                                }
                            }
                    }
                    // ---- y stage inside the thread: the middle node row is complete, the upper one goes to the warp above
                    double m1[3][2];
#pragma unroll
                    for (int c = 0; c < 3; ++c)
#pragma unroll
                        for (int bz = 0; bz < 2; ++bz) m1[c][bz] = (AA[c][2 * bz] + AA[c][2 * bz + 1]) + (AB[c][2 * bz] - AB[c][2 * bz + 1]);
                    if (!is_top) {
                        if (XE && xcount >= 2) sw_bar_wait_a(sw_bar_at(b_xempty_me, xs * SW_NW), ((xcount >> 1) - 1) & 1);
                        double* xo = xch_out + xs * XCH_SLOT;
#pragma unroll
                        for (int c = 0; c < 3; ++c)
#pragma unroll
                            for (int bz = 0; bz < 2; ++bz) xo[(c * 2 + bz) * 32] = AB[c][2 * bz] + AB[c][2 * bz + 1];
                        sw_syncwarp();
                        if (lane == 0) sw_bar_arrive_a(sw_bar_at(b_xfull_me, xs * SW_NW));
                    }
                    z_stage_store(1, m1, row_ok1 != 0);
                    // ---- lower node row: add what the warp below left for it
                    if (!is_bottom) {
                        sw_bar_wait_a(sw_bar_at(b_xfull_lo, xs * SW_NW), xpar);
                        const double* xi = xch_in + xs * XCH_SLOT;
                        double m0[3][2];
#pragma unroll
                        for (int c = 0; c < 3; ++c)
#pragma unroll
                            for (int bz = 0; bz < 2; ++bz) m0[c][bz] = (AA[c][2 * bz] - AA[c][2 * bz + 1]) + xi[(c * 2 + bz) * 32];
                        if (XE) {
                            sw_syncwarp();
                            if (lane == 0) sw_bar_arrive_a(sw_bar_at(b_xempty_lo, xs * SW_NW));
                        }
                        z_stage_store(0, m0, row_ok0 != 0);
                    }
                    if (NF > 0 && plane_ok) {
                        sw_syncwarp();
                        if (lane == 0) sw_bar_arrive_a(sw_bar_at(b_fempty, fs));
                        ++t;
                        if (++fs == SM::FN) {
                            fs = 0;
                            fpar ^= 1;
                        }
                    }
                    ++xcount;
                }
                yp += 3 * plane8;
                if (TR::SMOOTH) dp += 3 * plane8;
            };
            auto advance_ring = [&]() {
                pslot = slot;
                ppar = par;
                if (++slot == RING) {
                    slot = 0;
                    par ^= 1;
                }
            };
            auto release = [&](int sl) {
                sw_syncwarp();
                if (lane == 0) sw_bar_arrive_a(sw_bar_at(b_empty, sl));
            };
            // z butterflies of my three node rows of the plane in ring slot sl
            auto load_rows = [&](int sl, double (&zs)[3][3], double (&zd)[3][3]) {
                const double* p = in_thr + sl * SW_IN_SLOT;
#pragma unroll
                for (int c = 0; c < 3; ++c)
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        const double u0 = p[c * SW_ROWS * SW_BOXZ + r * SW_BOXZ], u1 = p[c * SW_ROWS * SW_BOXZ + r * SW_BOXZ + 1];
                        zs[r][c] = u1 + u0;
                        zd[r][c] = u1 - u0;
                    }
            };
            // the carry of segment 0 is zero
            for (int i = w * 32 + lane; i < (nsteps + 1) * SW_CARRY_ROW; i += 32 * SW_NW) carry[i] = 0.0;
            sw_consumer_barrier();  // consumer warps only

            double GA[3][4], GB[3][4];  // contributions of my two cells to the NEXT node plane (x-high half)
            int lp = pl0;
            int seg = 0;
            // one x step: FpX = modes of the previous plane (input), FnX = modes of the new plane (output)
            auto step = [&](const double (&FpA)[3][4], const double (&FpB)[3][4], double (&FnA)[3][4], double (&FnB)[3][4], int i) {
                sw_bar_wait_a(sw_bar_at(b_full, slot), par);
                double zs[3][3], zd[3][3];
                load_rows(slot, zs, zd);
                const double EA = e_thr[slot * SW_E_BOX], EB = e_thr[slot * SW_E_BOX + 32];
                double AA[3][4], AB[3][4];
                sw_cell_step(a.blk, zs[0], zd[0], zs[1], zd[1], EA, FpA, FnA, GA, AA);
                sw_cell_step(a.blk, zs[1], zd[1], zs[2], zd[2], EB, FpB, FnB, GB, AB);
                const int held = pslot;  // plane lp: its input slot is released once the plane is stored
                finish_plane(AA, AB, seg, lp, i);
                release(held);
                advance_ring();
                ++lp;
            };
            for (seg = 0; seg < nseg; ++seg) {
                // ---- first plane of the segment: only its y-z transform is needed
                double F0A[3][4], F0B[3][4], F1A[3][4], F1B[3][4];
                sw_bar_wait_a(sw_bar_at(b_full, slot), par);
                {
                    double zs[3][3], zd[3][3];
                    load_rows(slot, zs, zd);
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        F0A[c][0] = zs[1][c] + zs[0][c];
                        F0A[c][1] = zs[1][c] - zs[0][c];
                        F0A[c][2] = zd[1][c] + zd[0][c];
                        F0A[c][3] = zd[1][c] - zd[0][c];
                        F0B[c][0] = zs[2][c] + zs[1][c];
                        F0B[c][1] = zs[2][c] - zs[1][c];
                        F0B[c][2] = zd[2][c] + zd[1][c];
                        F0B[c][3] = zd[2][c] - zd[1][c];
                    }
                }
                advance_ring();
#pragma unroll
                for (int c = 0; c < 3; ++c)
#pragma unroll
                    for (int b = 0; b < 4; ++b) GA[c][b] = GB[c][b] = 0.0;
                lp = pl0;
                {
                    const long long off0 = (long long)pl0 * 3 * plane8 + (long long)ey0 * row8 + 32 * seg + lane;
                    yp = a.y + off0;
                    if (TR::SMOOTH) dp = a.d_out + off0;
                }
                int i = 0;
#pragma unroll 1
                for (; i + 1 < nsteps; i += 2) {
                    step(F0A, F0B, F1A, F1B, i);
                    step(F1A, F1B, F0A, F0B, i + 1);
                }
                if (i < nsteps) step(F0A, F0B, F1A, F1B, i);
                // ---- last plane of the chunk: only the x-high half of the last cell layer
                finish_plane(GA, GB, seg, lp, nsteps);
                release(pslot);
            }
        }
    }
    if (TR::DOT) {
        const unsigned int nblk = gridDim.x * gridDim.y;
        const unsigned int bid = blockIdx.x + gridDim.x * blockIdx.y;
        grid_reduce_idx<1, false>(pq, a.partial, a.ticket, a.dot_out, bid, nblk);
    }
}
