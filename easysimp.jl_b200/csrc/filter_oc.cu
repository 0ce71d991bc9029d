// filter_oc.cu -- SIMP modulus, FilterCache applies (stencil on structured grids, CSR otherwise)
// and the optimality-criteria update.
//
// Replaces
//   create_simp_material_model        FiniteElementAnalysis.jl:100-112   (E = Emin + (E0-Emin) rho^p)
//   apply_sensitivity_filter_cached!  SensitivityFilter.jl:33-67
//   apply_density_filter_cached!      DensityFilter.jl:30-60
//   apply_density_filter_chain_rule_cached!  DensityFilter.jl:77-117
//   optimality_criteria_update        OptimalityCriteria.jl:69-146
// On a structured grid the KD-tree neighbour lists of create_filter_cache (FilterCommon.jl:61-98)
// are a fixed stencil {(di,dj,dk): |offset| <= R}; weights max(0, R - dist) sit in constant memory
// and the truncated stencils at the domain boundary give per-cell denominators.
#include <math.h>

#include <algorithm>

#include "ctx.cuh"

constexpr int MAX_STENCIL = 1100;
__constant__ int c_st_di[MAX_STENCIL], c_st_dj[MAX_STENCIL], c_st_dk[MAX_STENCIL];
__constant__ double c_st_w[MAX_STENCIL];
static const esimp_ctx* g_stencil_owner[64] = {};  // per device

#define LAUNCH_STREAM(kern, ...)                                        \
    do {                                                                \
        kern<<<RED_BLOCKS, RED_THREADS, 0, ctx->stream>>>(__VA_ARGS__); \
        ctx->launches++;                                                \
        ESIMP_CUDA(cudaGetLastError());                                 \
    } while (0)

// ---------------------------------------------------------------------------- modulus
__global__ void k_modulus(const double* __restrict__ rho, double* __restrict__ E, long long n, double E0, double Emin, double p) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double r = rho[i];
        double rp = (p == 3.0) ? r * r * r : pow(r, p);
        E[i] = Emin + (E0 - Emin) * rp;
    }
}
int simp_modulus(esimp_ctx* ctx) {
    LAUNCH_STREAM(k_modulus, ctx->d_phys, ctx->d_E, (long long)ctx->ncell_loc, ctx->prm.E0, ctx->prm.Emin, ctx->prm.p);
    return ESIMP_OK;
}

// ---------------------------------------------------------------------------- stencil filter
void filter_forget_ctx(const esimp_ctx* ctx) {
    std::lock_guard<std::mutex> lock(g_const_mutex);
    if (g_stencil_owner[ctx->device & 63] == ctx) g_stencil_owner[ctx->device & 63] = nullptr;
}
static int upload_stencil(esimp_ctx* ctx) {
    std::lock_guard<std::mutex> lock(g_const_mutex);
    if (g_stencil_owner[ctx->device & 63] == ctx) return ESIMP_OK;
    std::vector<int> di(ctx->st_n), dj(ctx->st_n), dk(ctx->st_n);
    for (int s = 0; s < ctx->st_n; ++s) {
        di[s] = ctx->st_off[3 * s];
        dj[s] = ctx->st_off[3 * s + 1];
        dk[s] = ctx->st_off[3 * s + 2];
    }
    ESIMP_CUDA(cudaMemcpyToSymbol(c_st_di, di.data(), sizeof(int) * ctx->st_n));
    ESIMP_CUDA(cudaMemcpyToSymbol(c_st_dj, dj.data(), sizeof(int) * ctx->st_n));
    ESIMP_CUDA(cudaMemcpyToSymbol(c_st_dk, dk.data(), sizeof(int) * ctx->st_n));
    ESIMP_CUDA(cudaMemcpyToSymbol(c_st_w, ctx->st_w.data(), sizeof(double) * ctx->st_n));
    g_stencil_owner[ctx->device & 63] = ctx;
    return ESIMP_OK;
}

// D_i = sum_j w_ij V_j over the in-domain stencil (denominator of the density filter); computed for
// every stored cell (ghost layers too) from GLOBAL grid bounds, so it needs no halo.
__global__ void k_filter_D(GridDev g, int nst, double vol, double* __restrict__ D, long long n) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        int k = (int)(e % g.nz);
        long long t = e / g.nz;
        int j = (int)(t % g.ny);
        int gi = g.ge0 + (int)(t / g.ny);
        double den = 0.0;
        for (int s = 0; s < nst; ++s) {
            int ii = gi + c_st_di[s], jj = j + c_st_dj[s], kk = k + c_st_dk[s];
            if (ii < 0 || ii >= g.nx || jj < 0 || jj >= g.ny || kk < 0 || kk >= g.nz) continue;
            den += c_st_w[s] * vol;
        }
        D[e] = den;
    }
}

int filter_build_stencil(esimp_ctx* ctx) {
    const double R = ctx->prm.filter_radius;
    const int ri = (int)floor(R / ctx->hx + 1e-9), rj = (int)floor(R / ctx->hy + 1e-9), rk = (int)floor(R / ctx->hz + 1e-9);
    ctx->st_off.clear();
    ctx->st_w.clear();
    // neighbour order: ascending caller cell id (x fastest) like the sorted lists of the oracle
    for (int dk = -rk; dk <= rk; ++dk)
        for (int dj = -rj; dj <= rj; ++dj)
            for (int di = -ri; di <= ri; ++di) {
                double dx = di * ctx->hx, dy = dj * ctx->hy, dz = dk * ctx->hz;
                double w = R - sqrt(dx * dx + dy * dy + dz * dz);
                if (w > 0.0) {  // weight > 0 only (SensitivityFilter.jl:52, DensityFilter.jl:48)
                    ctx->st_off.push_back(di);
                    ctx->st_off.push_back(dj);
                    ctx->st_off.push_back(dk);
                    ctx->st_w.push_back(w);
                }
            }
    ctx->st_n = (int)ctx->st_w.size();
    if (ctx->st_n > MAX_STENCIL) return esimp_fail(ctx, ESIMP_ERR_ARG, "filter radius too large for the stencil table");
    ctx->st_halo = 0;
    for (int s = 0; s < ctx->st_n; ++s) ctx->st_halo = std::max(ctx->st_halo, abs(ctx->st_off[3 * s]));
    if (ctx->nranks > 1 && ctx->st_halo > ctx->halo_e)
        return esimp_fail(ctx, ESIMP_ERR_STATE, "filter radius needs more ghost layers than the slab holds");
    {
        std::lock_guard<std::mutex> lock(g_const_mutex);
        g_stencil_owner[ctx->device & 63] = nullptr;
    }
    ESIMP_TRY(upload_stencil(ctx));
    if (!ctx->d_filtD) ESIMP_CUDA(cudaMalloc(&ctx->d_filtD, sizeof(double) * ctx->ncell_loc));
    LAUNCH_STREAM(k_filter_D, ctx->g, ctx->st_n, ctx->hx * ctx->hy * ctx->hz, ctx->d_filtD, (long long)ctx->ncell_loc);
    ctx->filter_ready = true;
    return ESIMP_OK;
}

// mode 0: sensitivity filter, 1: density filter, 2: density filter transpose
template <int MODE>
__global__ void __launch_bounds__(256) k_filter_stencil(GridDev g, int le0, int le1, int nst, double vol,
                                                        const double* __restrict__ a, const double* __restrict__ b,
                                                        const double* __restrict__ D, double* __restrict__ out) {
    // One thread per cell, blockIdx.y = one (x-layer, y-row) pair so that the cell triple comes without any
    // division; bounds are three unsigned compares per stencil entry and the per-entry 1/vol of the
    // sensitivity filter is one multiplication by the reciprocal (exact for the unit cells of the
    // reference's generated grids, within one ulp otherwise: the parity tests hold 1e-10).
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= g.nz) return;
    const int nrow = (le1 - le0) * g.ny;
    const double ivol = 1.0 / vol;
    for (int row = blockIdx.y; row < nrow; row += gridDim.y) {
        const int lrow = row / g.ny;
        const int j = row - lrow * g.ny, le = le0 + lrow;
        const int gi = g.ge0 + le;
        const long long e = g.cell(le, j, k);
        const double* __restrict__ ae = a + e;
        const double* __restrict__ be = (MODE == 0) ? b + e : nullptr;
        const double* __restrict__ De = (MODE == 2) ? D + e : nullptr;
        double num = 0.0, den = 0.0;
        for (int s = 0; s < nst; ++s) {
            const int di = c_st_di[s], dj = c_st_dj[s], dk = c_st_dk[s];
            if ((unsigned)(gi + di) >= (unsigned)g.nx || (unsigned)(j + dj) >= (unsigned)g.ny ||
                (unsigned)(k + dk) >= (unsigned)g.nz)
                continue;
            const long long off = (long long)di * g.eplane + (dj * g.nz + dk);
            const double w = c_st_w[s];
            if (MODE == 0) {  // a = rho (design), b = dc
                num += w * ae[off] * be[off] * ivol;
                den += w;
            } else if (MODE == 1) {  // a = rho
                double wv = w * vol;
                num += wv * ae[off];
                den += wv;
            } else {  // a = s ; out_e = sum_i (w_ie V_e / D_i) s_i
                double Di = De[off];
                if (Di > 1e-12) num += (w * vol / Di) * ae[off];
            }
        }
        double r;
        if (MODE == 0) {
            double rho_safe = fmax(1e-3, ae[0]);
            r = den > 1e-12 ? num / (rho_safe / vol * den) : be[0];
        } else if (MODE == 1) {
            r = den > 1e-12 ? num / den : ae[0];
        } else {
            r = num;
        }
        out[e] = r;
    }
}

// CSR variants (unstructured meshes)
template <int MODE>
__global__ void k_filter_csr(long long n, const long long* __restrict__ rowptr, const int* __restrict__ col,
                             const double* __restrict__ w, const double* __restrict__ vol, const double* __restrict__ a,
                             const double* __restrict__ b, const double* __restrict__ D, double* __restrict__ out) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        double num = 0.0, den = 0.0;
        for (long long q = rowptr[e]; q < rowptr[e + 1]; ++q) {
            double wq = w[q];
            if (!(wq > 0.0)) continue;
            int nb = col[q];
            if (MODE == 0) {
                num += wq * a[nb] * b[nb] / vol[nb];
                den += wq;
            } else if (MODE == 1) {
                double wv = wq * vol[nb];
                num += wv * a[nb];
                den += wv;
            } else {  // symmetric neighbour relation: w_ie = w_ei
                double Di = D[nb];
                if (Di > 1e-12) num += (wq * vol[e] / Di) * a[nb];
            }
        }
        double r;
        if (MODE == 0) {
            double rho_safe = fmax(1e-3, a[e]);
            r = den > 1e-12 ? num / (rho_safe / vol[e] * den) : b[e];
        } else if (MODE == 1) {
            r = den > 1e-12 ? num / den : a[e];
        } else {
            r = num;
        }
        out[e] = r;
    }
}

template <int MODE>
static int filter_apply(esimp_ctx* ctx, const double* a, const double* b, double* out) {
    if (!ctx->filter_ready) return esimp_fail(ctx, ESIMP_ERR_STATE, "filter not built");
    if (ctx->structured) {
        ESIMP_TRY(upload_stencil(ctx));
        auto kern = k_filter_stencil<MODE>;
        PROF_BEGIN(ctx, PROF_FILTER, true)
        const int bx = ctx->g.nz >= 192 ? 256 : (ctx->g.nz >= 96 ? 128 : (ctx->g.nz >= 48 ? 64 : 32));
        const long long nrow = (long long)(ctx->own_e1 - ctx->own_e0) * ctx->g.ny;
        const dim3 grid((unsigned)((ctx->g.nz + bx - 1) / bx), (unsigned)std::max(1LL, std::min(nrow, 65535LL)));
        kern<<<grid, bx, 0, ctx->stream>>>(ctx->g, ctx->own_e0, ctx->own_e1, ctx->st_n, ctx->hx * ctx->hy * ctx->hz, a, b,
                                           ctx->d_filtD, out);
        ctx->launches++;
        ESIMP_CUDA(cudaGetLastError());
        PROF_END(ctx)
    } else {
        const UnstructuredMesh& um = ctx->um;
        auto kern = k_filter_csr<MODE>;
        LAUNCH_STREAM(kern, (long long)um.n_cells, (const long long*)um.d_rowptr, um.d_col, um.d_w, um.d_vol, a, b,
                      um.d_D, out);
    }
    return ESIMP_OK;
}
int filter_sens_apply(esimp_ctx* ctx, const double* rho, const double* dc, double* out) { return filter_apply<0>(ctx, rho, dc, out); }
int filter_dens_apply(esimp_ctx* ctx, const double* rho, double* out) { return filter_apply<1>(ctx, rho, nullptr, out); }
int filter_densT_apply(esimp_ctx* ctx, const double* s, double* out) { return filter_apply<2>(ctx, s, nullptr, out); }

// ---------------------------------------------------------------------------- optimality criteria
// One bisection step (OptimalityCriteria.jl:102-126): candidate for every stored cell (ghost layers
// too, so a following density filter needs no halo exchange) and, when no density filter is active,
// the volume sum over owned cells with a warp-shuffle reduction.
__global__ void k_oc_candidate(const double* __restrict__ rho, const double* __restrict__ dc, const double* __restrict__ dv,
                               const double* __restrict__ vol, double vol_uniform, double lam, double move, double damping,
                               double* __restrict__ rn, long long n, long long o0, long long o1, double* partial,
                               unsigned int* ticket, RedOut<1> out) {
    double acc[1] = {0.0};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double x = rho[i];
        double Be = fabs(dc[i]) / (lam * dv[i]);
        double cand = x * ((damping == 0.5) ? sqrt(Be) : pow(Be, damping));
        double v = fmax(1e-3, fmax(x - move, fmin(1.0, fmin(x + move, cand))));
        rn[i] = v;
        if (i >= o0 && i < o1) acc[0] += v * (vol ? vol[i] : vol_uniform);
    }
    grid_reduce<1, false>(acc, partial, ticket, out);
}

// ---- multi-section form of the same bisection (no density filter inside the loop) ----------------------
// cand_i(lambda) = rho_i (|dc_i| / (lambda dV_i))^damping = t_i * lambda^-damping with t_i hoisted out of the loop.
// One pass evaluates the volume for the 2^OC_LEVELS - 1 midpoints the next OC_LEVELS bisection steps can visit
// (16 B/cell per pass instead of 32 B/cell per step); the host then walks exactly the reference's bisection
// (same midpoints 0.5 (lo + hi), same |dV| < 1e-6 break, same update of lo / hi).
constexpr int OC_LEVELS = 4, OC_M = (1 << OC_LEVELS) - 1;
struct OcScales {
    double s[OC_M];  // lambda_m^-damping
};
__global__ void k_oc_prepare(const double* __restrict__ rho, const double* __restrict__ dc, const double* __restrict__ dv,
                             double damping, double* __restrict__ t, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double b = fabs(dc[i]) / dv[i];
        t[i] = rho[i] * ((damping == 0.5) ? sqrt(b) : pow(b, damping));
    }
}
__global__ void __launch_bounds__(256) k_oc_volumes(const double* __restrict__ rho, const double* __restrict__ t,
                                                    const double* __restrict__ vol, double vol_uniform, OcScales sc, double move,
                                                    long long o0, long long o1, double* partial, unsigned int* ticket,
                                                    RedOut<OC_M> out) {
    double acc[OC_M];
#pragma unroll
    for (int m = 0; m < OC_M; ++m) acc[m] = 0.0;
    for (long long i = o0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < o1; i += (long long)gridDim.x * blockDim.x) {
        const double x = rho[i], ti = t[i], v = vol ? vol[i] : vol_uniform;
        const double lo = fmax(1e-3, x - move), hi = fmin(1.0, x + move);
#pragma unroll
        for (int m = 0; m < OC_M; ++m) acc[m] = fma(fmax(lo, fmin(hi, ti * sc.s[m])), v, acc[m]);
    }
    grid_reduce<OC_M, false>(acc, partial, ticket, out);
}
__global__ void k_oc_commit(const double* __restrict__ rho, const double* __restrict__ t, double s, double move,
                            double* __restrict__ rn, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double x = rho[i];
        rn[i] = fmax(fmax(1e-3, x - move), fmin(fmin(1.0, x + move), t[i] * s));
    }
}

static int oc_update_multisection(esimp_ctx* ctx, const double* rho, const double* dc, const double* dv, double* rho_new,
                                  double* lambda, int* steps, double target) {
    const double* vol = ctx->unstructured ? ctx->um.d_vol : nullptr;
    const double vuni = ctx->hx * ctx->hy * ctx->hz, damping = ctx->prm.damping, move = ctx->prm.move_limit;
    double* t = ctx->d_ctmp;
    LAUNCH_STREAM(k_oc_prepare, rho, dc, dv, damping, t, (long long)ctx->ncell_loc);
    double lo = 1e-9, hi = 1e9, lam = NAN;
    int it = 0, status = ESIMP_WARN_OC_MAXIT;
    bool done = false;
    RedOut<OC_M> out;
    for (int m = 0; m < OC_M; ++m) out.p[m] = ctx->d_scal + SC_OC0 + m;
    while (!done && it < 200) {
        // heap-ordered tree of the midpoints reachable in the next OC_LEVELS steps: node 0 = mid(lo,hi),
        // children of node n: 2n+1 = mid(lo_n, mid_n) (taken when err <= 0), 2n+2 = mid(mid_n, hi_n) (err > 0)
        double nlo[OC_M], nhi[OC_M], mid[OC_M];
        OcScales sc;
        nlo[0] = lo;
        nhi[0] = hi;
        for (int n = 0; n < OC_M; ++n) {
            mid[n] = 0.5 * (nlo[n] + nhi[n]);
            sc.s[n] = (damping == 0.5) ? 1.0 / sqrt(mid[n]) : pow(mid[n], -damping);
            if (2 * n + 2 < OC_M) {
                nlo[2 * n + 1] = nlo[n];
                nhi[2 * n + 1] = mid[n];
                nlo[2 * n + 2] = mid[n];
                nhi[2 * n + 2] = nhi[n];
            }
        }
        {
            PROF_BEGIN(ctx, PROF_OC_VOLUMES, true)
            LAUNCH_STREAM(k_oc_volumes, rho, t, vol, vuni, sc, move, (long long)ctx->own_cell0, (long long)ctx->own_cell1, ctx->d_partial,
                          ctx->d_ticket, out);
            PROF_END(ctx)
        }
        ESIMP_TRY(allreduce_scalars(ctx, SC_OC0, OC_M, false));
        ESIMP_TRY(fetch_scalars(ctx));
        int n = 0;
        for (int lev = 0; lev < OC_LEVELS && it < 200; ++lev) {
            ++it;
            lam = mid[n];
            double err = ctx->h_scal[SC_OC0 + n] - target;
            if (fabs(err) < 1e-6) {
                status = ESIMP_OK;
                done = true;
                break;
            }
            if (err > 0) {
                lo = lam;
                n = 2 * n + 2;
            } else {
                hi = lam;
                n = 2 * n + 1;
            }
        }
    }
    const double s = (damping == 0.5) ? 1.0 / sqrt(lam) : pow(lam, -damping);
    LAUNCH_STREAM(k_oc_commit, rho, t, s, move, rho_new, (long long)ctx->ncell_loc);
    if (lambda) *lambda = lam;
    if (steps) *steps = it;
    return status;
}

int oc_update(esimp_ctx* ctx, const double* rho, const double* dc, const double* dv, double* rho_new, double* lambda, int* steps) {
    const bool use_df = ctx->prm.filter_type == ESIMP_FILTER_DENSITY;
    const double* vol = ctx->unstructured ? ctx->um.d_vol : nullptr;
    const double vuni = ctx->hx * ctx->hy * ctx->hz;
    // total volume and target (OptimalityCriteria.jl:84)
    double total_volume;
    if (ctx->unstructured) {
        ESIMP_TRY(reduce_sum_scaled(ctx, ctx->um.d_vol, nullptr, 1.0, 0, ctx->ncell_loc, SC_TMP0));
        ESIMP_TRY(fetch_scalars(ctx));
        total_volume = ctx->h_scal[SC_TMP0];
    } else {
        total_volume = vuni * (double)ctx->n_cells_global;
    }
    const double target = ctx->prm.volume_fraction * total_volume;
    if (!use_df) return oc_update_multisection(ctx, rho, dc, dv, rho_new, lambda, steps, target);
    // density filter: the filter sits inside the bisection (OptimalityCriteria.jl:120-122), one step per pass
    double lo = 1e-9, hi = 1e9, lam = NAN;
    int it = 0, status = ESIMP_WARN_OC_MAXIT;
    RedOut<1> out{{ctx->d_scal + SC_VOL}};
    for (it = 1; it <= 200; ++it) {
        lam = 0.5 * (lo + hi);
        LAUNCH_STREAM(k_oc_candidate, rho, dc, dv, vol, vuni, lam, ctx->prm.move_limit, ctx->prm.damping, rho_new,
                      (long long)ctx->ncell_loc, (long long)ctx->own_cell0, (long long)ctx->own_cell1, ctx->d_partial, ctx->d_ticket,
                      out);
        if (use_df) {
            ESIMP_TRY(filter_dens_apply(ctx, rho_new, ctx->d_ctmp));
            ESIMP_TRY(reduce_sum_scaled(ctx, ctx->d_ctmp, vol, vol ? 1.0 : vuni, ctx->own_cell0, ctx->own_cell1, SC_VOL));
        } else {
            ESIMP_TRY(allreduce_scalars(ctx, SC_VOL, 1, false));
        }
        ESIMP_TRY(fetch_scalars(ctx));
        double err = ctx->h_scal[SC_VOL] - target;
        if (fabs(err) < 1e-6) {
            status = ESIMP_OK;
            break;
        }
        if (err > 0)
            lo = lam;
        else
            hi = lam;
    }
    if (it > 200) it = 200;
    if (lambda) *lambda = lam;
    if (steps) *steps = it;
    return status;
}
