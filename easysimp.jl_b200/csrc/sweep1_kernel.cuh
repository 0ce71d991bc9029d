// sweep1_kernel.cuh -- the sweep kernel (sweep_kernel.cuh) with ONE cell per thread and NW consumer warps per CTA.
//
// Why: ncu on the two-cells-per-thread kernel (profiles/r2_final_ncu_operator_modes.json, plain mode) shows 2.1 resident warps per
// scheduler issuing 0.54 instructions per cycle, each warp 3.9 cycles per instruction (1.28 of them fixed-latency waits, 0.73 on mbarrier /
// shared-memory scoreboards): the FP64 pipe is 61 % busy and neither it nor DRAM (45 %) is the limit -- latency hiding is.  The two-cell
// thread needs 232 registers, so 8 consumer warps are all an SM can hold.  A one-cell thread runs in 112 registers (16 consumer warps next to a
// producer warpgroup that keeps 24) or 160 (12 consumer warps): 3 - 4 warps per scheduler at the price of one y exchange and one set of barrier
// operations per cell instead of per two cells, and z butterflies that are no longer shared between stacked cells.
//
// Same tile geometry idea (a CTA owns NW - 1 node rows x all of z x one x chunk, z swept in 32-lane segments, carry rows through shared
// memory), same TMA boxes with NW + 1 / NW / NW - 1 rows, same store-phase modes, same arithmetic (sw_cell_step).  The y exchange between
// adjacent warps always uses the two-slot ring with an explicit "empty" barrier (16 warps x RING slots would not fit next to the operand
// boxes).  Opt-in: ESIMP_SWEEP_CPT1 = 16 | 12 (hex8_structured.cu); compiled for the host like its sibling (selftest_kernel_host.cpp).
#pragma once

constexpr int S1_PW = 4;  // producer warps: one whole warpgroup (setmaxnreg works per warpgroup; the register file is allocated in units of
                          // four warps anyway, so a lone producer warp would cost the same); one thread of it issues
#ifndef ESIMP_HOST_EMULATION
template <int NW>
__device__ __forceinline__ void sw1_consumer_barrier() {
    asm volatile("bar.sync 1, %0;" ::"n"(32 * NW) : "memory");
}
// register reallocation: the CTA is launched with 65536 / threads registers per thread (96 for 16 + 4 warps, 128 for 12 + 4); the producer
// warpgroup keeps 24, the consumers take what that frees: 112 (16 warps) / 160 (12 warps)
__device__ __forceinline__ void sw1_regs_producer() { asm volatile("setmaxnreg.dec.sync.aligned.u32 24;\n"); }
template <int NW>
__device__ __forceinline__ void sw1_regs_consumer() {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(NW == 16 ? 112 : 160));
}
#endif

template <int NW>
struct S1Geo {
    static constexpr int CROWS = NW;        // cell rows of a tile (warp w = cell row Y0 + w)
    static constexpr int ROWS = NW + 1;     // node rows of an input box
    static constexpr int OROWS = NW - 1;    // node rows a tile owns = rows of an operand box
    static constexpr int THREADS = 32 * (NW + S1_PW);  // warps 0 .. 3: the producer warpgroup
    static constexpr int IN_BOX = 3 * ROWS * SW_BOXZ;
    static constexpr int IN_SLOT = (IN_BOX * 8 + 127) / 128 * 16;
    static constexpr int OP_BOX = 3 * OROWS * 32;   // bytes: a multiple of 128 for NW = 16 (11520) and NW = 12 (8448)
    static constexpr int OPX_BOX = 3 * OROWS * 2;
    static constexpr int E_BOX = CROWS * 32;
    static constexpr int CARRY_ROW = CROWS * 3;
    static_assert((OP_BOX * 8) % 128 == 0 && (E_BOX * 8) % 128 == 0 && OPX_BOX <= SW_OPX_SLOT, "TMA destinations are 128-byte aligned");
};

template <int NF, int RING, int NW>
struct Sweep1Smem {
    using GEO = S1Geo<NW>;
    static constexpr int FN = NF > 0 ? 2 : 1;
    alignas(128) double in[RING][GEO::IN_SLOT];
    alignas(128) double E[RING][GEO::E_BOX];
    alignas(128) double fz[FN][NF > 0 ? NF : 1][NF > 0 ? GEO::OP_BOX : 16];
    alignas(128) double fzx[FN][NF > 0 ? NF : 1][NF > 0 ? SW_OPX_SLOT : 16];
    double xch[2][NW][3][2][32];  // y-high mode coefficients (bz = 0, 1) of warp w's cell for node row w + 1 (warp w + 1)
    SwBar full[RING], empty[RING], ffull[FN], fempty[FN], xfull[2][NW], xempty[2][NW];
};

template <int MODE, int RING, int NW>
__global__ void __launch_bounds__(S1Geo<NW>::THREADS, 1)
    k_matvec_sweep1(const __grid_constant__ SweepArgs a, const __grid_constant__ SwTensorMap tmU, const __grid_constant__ SwTensorMap tmE,
                    const __grid_constant__ SwTensorMap tmR, const __grid_constant__ SwTensorMap tmDinv, const __grid_constant__ SwTensorMap tmD,
                    const __grid_constant__ SwTensorMap tmRx, const __grid_constant__ SwTensorMap tmDinvx, const __grid_constant__ SwTensorMap tmDx) {
    using TR = SweepTraits<MODE>;
    using GEO = S1Geo<NW>;
    constexpr int NF = TR::NF, NFS = NF > 0 ? NF : 1;
    SW_SMEM_DECL;
    using SM = Sweep1Smem<NF, RING, NW>;
    SM& sm = *reinterpret_cast<SM*>(sw_smem_raw);
    double* const carry = reinterpret_cast<double*>(sw_smem_raw + sizeof(SM));
    const GridDev& g = a.g;
    int w = (int)(threadIdx.x >> 5) - S1_PW, lane = threadIdx.x & 31;  // w < 0: producer warpgroup, w = 0 .. NW - 1: consumer warps
    SW_PIN32(w);
    SW_PIN32(lane);
    const int Y0 = blockIdx.x * (GEO::CROWS - 1) - 1;
    const int P0 = a.ps0 + blockIdx.y * a.chunk, P1 = min(P0 + a.chunk, a.ps1);
    const int Lstart = max(P0 - g.e2p - 1, a.le0), Lend = min(P1 - g.e2p, a.le1);
    const int nsteps = Lend - Lstart, pl0 = Lstart + g.e2p;
    const int nseg = (g.nz + 31) / 32;
    const bool extra = (g.nz & 31) == 0;  // node column nz is nobody's own column: lane 31 of the last segment takes it along
    double pq[1] = {0.0};

    if (threadIdx.x == 0) {
        for (int s = 0; s < RING; ++s) {
            sw_bar_init(&sm.full[s], 1);
            sw_bar_init(&sm.empty[s], NW);
        }
        for (int s = 0; s < 2; ++s)
            for (int r = 0; r < NW; ++r) {
                sw_bar_init(&sm.xfull[s][r], 1);
                sw_bar_init(&sm.xempty[s][r], 1);
            }
        for (int s = 0; s < SM::FN; ++s) {
            sw_bar_init(&sm.ffull[s], 1);
            sw_bar_init(&sm.fempty[s], NW);
        }
        sw_fence_init();
    }
    __syncthreads();

    if (nsteps > 0) {
        if (w < 0) {
            // ------------------------------------------------------------------ producer: one thread feeds the rings
            sw1_regs_producer();
            if (w == -S1_PW && lane == 0) {
                int slot = 0, round = 0;  // input ring position; round counts completed trips through the ring
                int t = 0;                // operand boxes issued
                auto issue_operands = [&](int s, int lp) {
                    if (NF > 0 && lp >= P0 && lp < P1) {
                        const int fs = t % SM::FN;
                        if (t >= SM::FN) sw_bar_wait(&sm.fempty[fs], ((t / SM::FN) - 1) & 1);
                        const int nb = NF == 1 ? 1 : ((a.dmode == FUSE_D_BUFFER || a.dmode == FUSE_D_FROM_ZPREV) ? 3 : 2);  // third box: d or z_prev
                        const bool xcol = extra && s == nseg - 1;  // the last segment also brings the operands of node column nz
                        const int full = GEO::OP_BOX + (xcol ? GEO::OPX_BOX : 0);
                        sw_bar_expect_tx(&sm.ffull[fs], (unsigned)((NF == 1 ? full : (nb - 1) * full + full / 3) * 8));
                        const SwTensorMap* maps[6] = {&tmR, &tmDinv, &tmD, &tmRx, &tmDinvx, &tmDx};
                        for (int b = 0; b < nb; ++b) {
                            const int pz = (NF > 1 && b == 1) ? lp : 3 * lp;  // box 1 of the smoother modes: node scalar of D^-1 (sw_dinv)
                            sw_tma_3d(&sm.fz[fs][b % NFS][0], maps[b], 32 * s, Y0 + 1, pz, &sm.ffull[fs]);
                            if (xcol) sw_tma_3d(&sm.fzx[fs][b % NFS][0], maps[3 + b], g.nz, Y0 + 1, pz, &sm.ffull[fs]);
                        }
                        ++t;
                    }
                };
                const int per_seg = nsteps + 1, total = nseg * per_seg;
                auto prefetch_l2 = [&](int L) {
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
                        sw_bar_expect_tx(&sm.full[slot], (unsigned)((GEO::IN_BOX + (j >= 1 ? GEO::E_BOX : 0)) * 8));
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
            // ------------------------------------------------------------------ consumers: warp w = cell row Y0 + w, stores node row Y0 + w
            sw1_regs_consumer<NW>();
            int slot = 0, par = 0;      // input ring position and the parity its full barrier completes next
            int pslot = 0;              // slot of the previous plane (released after that plane's store)
            int fs = 0, fpar = 0;       // slot and parity of the next operand box
            int xcount = 0;             // y exchanges done
            const int ey0 = Y0 + w;     // node row of my cell's low corner: the row I store
            int row_ok = (w >= 1 && ey0 <= g.ny) ? 1 : 0;  // row 0 of a tile belongs to the tile below
            int is_top = w == NW - 1 ? 1 : 0, is_bottom = w == 0 ? 1 : 0;
            int thr_box = w * SW_BOXZ + lane;  // my lower node inside an input box plane
            int thr_e = w * 32 + lane;         // my cell inside a modulus box
            SW_PIN32(row_ok);
            SW_PIN32(is_top);
            SW_PIN32(is_bottom);
            SW_PIN32(thr_box);
            SW_PIN32(thr_e);
            const double* const in_thr = &sm.in[0][0] + thr_box;
            const double* const e_thr = &sm.E[0][0] + thr_e;
            const double* const fz_thr = &sm.fz[0][0][0] + (w - 1) * 32 + lane;  // operand-box row of my node row: w - 1 (never read by warp 0)
            const double* const fzx_thr = &sm.fzx[0][0][0] + (w - 1) * 2;
            double* const xch_out = &sm.xch[0][w][0][0][0] + lane;                       // what I leave for the warp above
            const double* const xch_in = &sm.xch[0][w > 0 ? w - 1 : 0][0][0][0] + lane;  // what the warp below left for me
            constexpr int XCH_SLOT = NW * 3 * 2 * 32;
            double* const carry_thr = carry + w * 3;
            SwBarAddr b_full = sw_bar_addr(&sm.full[0]), b_empty = sw_bar_addr(&sm.empty[0]);
            SwBarAddr b_ffull = sw_bar_addr(&sm.ffull[0]), b_fempty = sw_bar_addr(&sm.fempty[0]);
            SwBarAddr b_xfull_me = sw_bar_addr(&sm.xfull[0][w]), b_xfull_lo = sw_bar_addr(&sm.xfull[0][w > 0 ? w - 1 : 0]);
            SwBarAddr b_xempty_me = sw_bar_addr(&sm.xempty[0][w]), b_xempty_lo = sw_bar_addr(&sm.xempty[0][w > 0 ? w - 1 : 0]);
            SW_PIN32(b_full);
            SW_PIN32(b_empty);
            long long plane8 = g.plane;
            SW_PIN64(plane8);
            double* yp = a.y;      // (plane lp, component 0, row ey0, my column) of the result, advanced by one node plane per step
            double* dp = a.d_out;  // ... and of the direction vector the smoother updates

            // one output plane: A = y-z mode coefficients of my cell's contributions to node plane lp
            auto finish_plane = [&](const double (&A)[3][4], int s, int lp, int ci) {
                {
                    const bool plane_ok = lp >= P0 && lp < P1;  // CTA-uniform; a plane outside goes through the same exchange, only its stores are skipped
                    const bool last_seg = s == nseg - 1;
                    const bool ext = extra && last_seg && lane == 31;
                    const bool col_ok = plane_ok && 32 * s + lane <= g.nz;
                    const int xs = xcount & 1, xpar = (xcount >> 1) & 1;
                    const double* uin = in_thr + pslot * GEO::IN_SLOT;  // my node of plane lp (its slot is still held)
                    const double* fzp = fz_thr + fs * (NFS * GEO::OP_BOX);
                    const double* fzxp = fzx_thr + fs * (NFS * SW_OPX_SLOT);
                    double* cr = carry_thr + ci * GEO::CARRY_ROW;
                    double pu[3], pr[3], pz2[3], psn = 0.0;  // store-phase operands of my node, read before the exchange wait
                    if (NF > 0 && plane_ok) sw_bar_wait_a(sw_bar_at(b_ffull, fs), fpar);
                    if (!is_bottom) {
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            if (TR::DOT || TR::SMOOTH) pu[c] = uin[c * GEO::ROWS * SW_BOXZ];
                            if (NF > 0) pr[c] = fzp[c * GEO::OROWS * 32];
                            if (TR::SMOOTH) pz2[c] = fzp[(2 % NFS) * GEO::OP_BOX + c * GEO::OROWS * 32];
                        }
                        if (TR::SMOOTH) psn = fzp[(1 % NFS) * GEO::OP_BOX];
                    }
                    // store one component of my node: col = 0 my column, 1 the extra column nz (operands read in place: rare)
                    auto store_node = [&](int c, double s_val, int col) {
                        const long long off = c * plane8 + col;
                        const double* op = col ? fzxp + c * GEO::OROWS * 2 : fzp + c * GEO::OROWS * 32;
                        const int ostride = col ? SW_OPX_SLOT : GEO::OP_BOX;
                        if (TR::SMOOTH) {
                            const double rv = col ? op[0] : pr[c];
                            const double sraw = col ? fzxp[(1 % NFS) * SW_OPX_SLOT] : psn;
                            const double dinv = sw_dinv(sraw, c, a.ik[c]);
                            const double zin = col ? uin[c * GEO::ROWS * SW_BOXZ + 1] : pu[c];
                            const double third = col ? op[(2 % NFS) * ostride] : pz2[c];
                            const double dprev = a.dmode == FUSE_D_BUFFER ? third : a.dmode == FUSE_D_FROM_ZPREV ? zin - third : zin;  // FUSE_D_IS_Z / unused
                            const double v = fma(a.c1, dprev, a.c2 * dinv * (rv - s_val));
                            if (a.write_d) dp[off] = v;
                            const double znew = zin + v;
                            yp[off] = znew;
                            if (MODE == FUSE_SMOOTH_DOT) pq[0] = fma(rv, znew, pq[0]);
                        } else if (MODE == FUSE_RESID) {
                            yp[off] = (col ? op[0] : pr[c]) - s_val;
                        } else {
                            yp[off] = s_val;
                            if (MODE == FUSE_DOT) pq[0] = fma(s_val, col ? uin[c * GEO::ROWS * SW_BOXZ + 1] : pu[c], pq[0]);
                        }
                    };
                    // ---- y stage: the upper node row of my cell goes to the warp above, the lower one is mine
                    if (!is_top) {
                        if (xcount >= 2) sw_bar_wait_a(sw_bar_at(b_xempty_me, xs * NW), ((xcount >> 1) - 1) & 1);
                        double* xo = xch_out + xs * XCH_SLOT;
#pragma unroll
                        for (int c = 0; c < 3; ++c)
#pragma unroll
                            for (int bz = 0; bz < 2; ++bz) xo[(c * 2 + bz) * 32] = A[c][2 * bz] + A[c][2 * bz + 1];
                        sw_syncwarp();
                        if (lane == 0) sw_bar_arrive_a(sw_bar_at(b_xfull_me, xs * NW));
                    }
                    if (!is_bottom) {
                        sw_bar_wait_a(sw_bar_at(b_xfull_lo, xs * NW), xpar);
                        const double* xi = xch_in + xs * XCH_SLOT;
                        double m[3][2];  // z-even / z-odd coefficient of my node row, complete over y
#pragma unroll
                        for (int c = 0; c < 3; ++c)
#pragma unroll
                            for (int bz = 0; bz < 2; ++bz) m[c][bz] = (A[c][2 * bz] - A[c][2 * bz + 1]) + xi[(c * 2 + bz) * 32];
                        sw_syncwarp();
                        if (lane == 0) sw_bar_arrive_a(sw_bar_at(b_xempty_lo, xs * NW));
                        // ---- z stage
                        double high[3], hin[3];
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            high[c] = m[c][0] + m[c][1];  // -> node column k + 1
                            hin[c] = sw_shfl_up1(high[c]);
                        }
                        if (lane == 0) {  // the z-low neighbour of lane 0 is lane 31 of the previous segment (zeros in segment 0)
#pragma unroll
                            for (int c = 0; c < 3; ++c) hin[c] = cr[c];
                        }
                        sw_syncwarp();  // lane 0 has read the carry of the previous segment before lane 31 overwrites it
                        if (lane == 31 && !last_seg) {
#pragma unroll
                            for (int c = 0; c < 3; ++c) cr[c] = high[c];
                        }
                        if (row_ok && col_ok) {
#pragma unroll
                            for (int c = 0; c < 3; ++c) store_node(c, (m[c][0] - m[c][1]) + hin[c], 0);
                        }
                        if (ext && row_ok && plane_ok) {
#pragma unroll
                            for (int c = 0; c < 3; ++c) store_node(c, high[c], 1);
                        }
                    }
                    if (NF > 0 && plane_ok) {
                        sw_syncwarp();
                        if (lane == 0) sw_bar_arrive_a(sw_bar_at(b_fempty, fs));
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
                if (++slot == RING) {
                    slot = 0;
                    par ^= 1;
                }
            };
            auto release = [&](int sl) {
                sw_syncwarp();
                if (lane == 0) sw_bar_arrive_a(sw_bar_at(b_empty, sl));
            };
            // z butterflies of my two node rows of the plane in ring slot sl
            auto load_rows = [&](int sl, double (&zs)[2][3], double (&zd)[2][3]) {
                const double* p = in_thr + sl * GEO::IN_SLOT;
#pragma unroll
                for (int c = 0; c < 3; ++c)
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        const double u0 = p[c * GEO::ROWS * SW_BOXZ + r * SW_BOXZ], u1 = p[c * GEO::ROWS * SW_BOXZ + r * SW_BOXZ + 1];
                        zs[r][c] = u1 + u0;
                        zd[r][c] = u1 - u0;
                    }
            };
            // the carry of segment 0 is zero
            for (int i = w * 32 + lane; i < (nsteps + 1) * GEO::CARRY_ROW; i += 32 * NW) carry[i] = 0.0;
            sw1_consumer_barrier<NW>();  // consumer warps only

            double G[3][4];  // contributions of my cell to the NEXT node plane (x-high half)
            int lp = pl0;
            int seg = 0;
            auto step = [&](const double (&Fp)[3][4], double (&Fn)[3][4], int i) {
                sw_bar_wait_a(sw_bar_at(b_full, slot), par);
                double zs[2][3], zd[2][3];
                load_rows(slot, zs, zd);
                const double Ee = e_thr[slot * GEO::E_BOX];
                double A[3][4];
                sw_cell_step(a.blk, zs[0], zd[0], zs[1], zd[1], Ee, Fp, Fn, G, A);
                const int held = pslot;  // plane lp: its input slot is released once the plane is stored
                finish_plane(A, seg, lp, i);
                release(held);
                advance_ring();
                ++lp;
            };
            for (seg = 0; seg < nseg; ++seg) {
                // ---- first plane of the segment: only its y-z transform is needed
                double F0[3][4], F1[3][4];
                sw_bar_wait_a(sw_bar_at(b_full, slot), par);
                {
                    double zs[2][3], zd[2][3];
                    load_rows(slot, zs, zd);
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        F0[c][0] = zs[1][c] + zs[0][c];
                        F0[c][1] = zs[1][c] - zs[0][c];
                        F0[c][2] = zd[1][c] + zd[0][c];
                        F0[c][3] = zd[1][c] - zd[0][c];
                    }
                }
                advance_ring();
#pragma unroll
                for (int c = 0; c < 3; ++c)
#pragma unroll
                    for (int b = 0; b < 4; ++b) G[c][b] = 0.0;
                lp = pl0;
                {
                    const long long off0 = (long long)pl0 * 3 * plane8 + (long long)ey0 * g.nzp + 32 * seg + lane;
                    yp = a.y + off0;
                    if (TR::SMOOTH) dp = a.d_out + off0;
                }
                int i = 0;
#pragma unroll 1
                for (; i + 1 < nsteps; i += 2) {
                    step(F0, F1, i);
                    step(F1, F0, i + 1);
                }
                if (i < nsteps) step(F0, F1, i);
                // ---- last plane of the chunk: only the x-high half of the last cell layer
                finish_plane(G, seg, lp, nsteps);
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
