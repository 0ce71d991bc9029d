/*
 * cpu_simp.c -- CPU restatement (plain C + OpenMP) of the reference's per-iteration SIMP path on
 * structured hex grids.  TEST / BASELINE INFRASTRUCTURE ONLY: used by tests/ (checked against
 * oracle/simp_oracle.py) and by bench.py's cpu_baseline / `--impl reference` legs.  The product
 * (libeasysimp_b200.so) never links or calls it.
 *
 * PARITY UNPINNED by the reference's own tests (none assert numbers); pinned through
 * simp_oracle.py's known-answer tests (tests/test_oracle.py, tests/test_cpu_port.py).
 *
 * What it follows (reference file:line):
 *   cell loop  y += (E(rho_e)/E0) K_unit u_e      assemble_with_cache!, FiniteElementAnalysis.jl:678-712
 *                                                 (same scale-and-scatter, applied to u instead of to K)
 *   Dirichlet  zero rows/cols, diag = m           Ferrite apply!, Optimization.jl:581-583
 *   solve      Jacobi-PCG                         stands in for cholesky(Symmetric(K,:L)) \ f, Optimization.jl:317,
 *                                                 which cannot exist at 100 M dofs (K alone ~100 GB)
 *   dc_e       -p rho^(p-1)(E0-Emin) u_e'K0u_e    SensitivityAnalysis.jl:74-78
 *   filter     Sigmund sensitivity filter         SensitivityFilter.jl:33-67
 *   OC step    candidate + volume                 OptimalityCriteria.jl:102-126
 * Numbering is the caller's native one: node = i + (nx+1)(j + (ny+1)k), dof = 3 node + c,
 * cell = i + nx (j + ny k)  (Ferrite generate_grid order).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static const int SX[8] = {0, 1, 1, 0, 0, 1, 1, 0}, SY[8] = {0, 0, 1, 1, 0, 0, 1, 1}, SZ[8] = {0, 0, 0, 0, 1, 1, 1, 1};

int cpu_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void cpu_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

static inline int64_t node_id(int nx, int ny, int i, int j, int k) { return i + (int64_t)(nx + 1) * (j + (int64_t)(ny + 1) * k); }

/* y = P K(E) P u + m (I-P) u ; fixed: 1 byte per dof */
void cpu_hex8_matvec(int nx, int ny, int nz, const double* K0, const double* E, const uint8_t* fixed, double m,
                     const double* u, double* y) {
    const int64_t ndof = 3 * (int64_t)(nx + 1) * (ny + 1) * (nz + 1);
#pragma omp parallel for schedule(static)
    for (int64_t d = 0; d < ndof; ++d) y[d] = 0.0;
    for (int color = 0; color < 2; ++color) { /* layers k and k+2 share no node: race-free scatter */
#pragma omp parallel for schedule(dynamic, 1)
        for (int k = color; k < nz; k += 2) {
            for (int j = 0; j < ny; ++j)
                for (int i = 0; i < nx; ++i) {
                    int64_t nd[8];
                    double ue[24], ye[24];
                    for (int a = 0; a < 8; ++a) {
                        nd[a] = 3 * node_id(nx, ny, i + SX[a], j + SY[a], k + SZ[a]);
                        for (int c = 0; c < 3; ++c) ue[3 * a + c] = fixed[nd[a] + c] ? 0.0 : u[nd[a] + c];
                    }
                    const double Ee = E[i + (int64_t)nx * (j + (int64_t)ny * k)];
                    for (int r = 0; r < 24; ++r) {
                        double acc = 0.0;
                        for (int c = 0; c < 24; ++c) acc += K0[r * 24 + c] * ue[c];
                        ye[r] = Ee * acc;
                    }
                    for (int a = 0; a < 8; ++a)
                        for (int c = 0; c < 3; ++c) y[nd[a] + c] += ye[3 * a + c];
                }
        }
    }
#pragma omp parallel for schedule(static)
    for (int64_t d = 0; d < ndof; ++d)
        if (fixed[d]) y[d] = m * u[d];
}

/* diag(K) before BCs (sum over cells of E K0[r,r]) ; returns mean */
double cpu_hex8_diagonal(int nx, int ny, int nz, const double* K0, const double* E, double* diag) {
    const int64_t ndof = 3 * (int64_t)(nx + 1) * (ny + 1) * (nz + 1);
    memset(diag, 0, sizeof(double) * ndof);
    for (int color = 0; color < 2; ++color) {
#pragma omp parallel for schedule(dynamic, 1)
        for (int k = color; k < nz; k += 2)
            for (int j = 0; j < ny; ++j)
                for (int i = 0; i < nx; ++i) {
                    const double Ee = E[i + (int64_t)nx * (j + (int64_t)ny * k)];
                    for (int a = 0; a < 8; ++a) {
                        int64_t n3 = 3 * node_id(nx, ny, i + SX[a], j + SY[a], k + SZ[a]);
                        for (int c = 0; c < 3; ++c) diag[n3 + c] += Ee * K0[(3 * a + c) * 25];
                    }
                }
    }
    double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static)
    for (int64_t d = 0; d < ndof; ++d) s += diag[d];
    return s / (double)ndof;
}

static double dotp(const double* a, const double* b, int64_t n) {
    double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static)
    for (int64_t i = 0; i < n; ++i) s += a[i] * b[i];
    return s;
}

/* Jacobi-PCG on the BC-modified operator; x is the initial guess in / solution out.  work: 4*ndof
 * doubles.  Runs until ||r|| <= rtol ||f|| or maxit iterations; returns iterations done and *relres. */
int cpu_pcg(int nx, int ny, int nz, const double* K0, const double* E, const uint8_t* fixed, const double* f,
            double* x, double rtol, int maxit, double* work, double* relres) {
    const int64_t n = 3 * (int64_t)(nx + 1) * (ny + 1) * (nz + 1);
    double *r = work, *p = work + n, *q = work + 2 * n, *dinv = work + 3 * n;
    const double m = cpu_hex8_diagonal(nx, ny, nz, K0, E, dinv);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) dinv[i] = fixed[i] ? 1.0 / m : 1.0 / dinv[i];
    cpu_hex8_matvec(nx, ny, nz, K0, E, fixed, m, x, q);
    double ff = 0.0, rz = 0.0, rr = 0.0;
#pragma omp parallel for reduction(+ : ff, rz, rr) schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        double fi = fixed[i] ? 0.0 : f[i];
        if (fixed[i]) x[i] = 0.0;
        r[i] = fi - (fixed[i] ? 0.0 : q[i]);
        p[i] = dinv[i] * r[i];
        ff += fi * fi;
        rz += r[i] * p[i];
        rr += r[i] * r[i];
    }
    int it = 0;
    while (it < maxit && rr > rtol * rtol * ff) {
        cpu_hex8_matvec(nx, ny, nz, K0, E, fixed, m, p, q);
        const double alpha = rz / dotp(p, q, n);
        double rz_new = 0.0;
        rr = 0.0;
#pragma omp parallel for reduction(+ : rz_new, rr) schedule(static)
        for (int64_t i = 0; i < n; ++i) {
            x[i] += alpha * p[i];
            r[i] -= alpha * q[i];
            rz_new += r[i] * dinv[i] * r[i];
            rr += r[i] * r[i];
        }
        const double beta = rz_new / rz;
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; ++i) p[i] = dinv[i] * r[i] + beta * p[i];
        rz = rz_new;
        ++it;
    }
    if (relres) *relres = ff > 0 ? sqrt(rr / ff) : 0.0;
    return it;
}

/* dc_e and energy = 0.5 sum E_e u_e'K0u_e */
double cpu_sensitivities(int nx, int ny, int nz, const double* K0, const double* rho, const double* E, const double* u,
                         double E0, double Emin, double p, double* dc) {
    double energy = 0.0;
#pragma omp parallel for reduction(+ : energy) schedule(static)
    for (int k = 0; k < nz; ++k)
        for (int j = 0; j < ny; ++j)
            for (int i = 0; i < nx; ++i) {
                double ue[24];
                for (int a = 0; a < 8; ++a) {
                    int64_t n3 = 3 * node_id(nx, ny, i + SX[a], j + SY[a], k + SZ[a]);
                    for (int c = 0; c < 3; ++c) ue[3 * a + c] = u[n3 + c];
                }
                double q = 0.0;
                for (int r = 0; r < 24; ++r) {
                    double acc = 0.0;
                    for (int c = 0; c < 24; ++c) acc += K0[r * 24 + c] * ue[c];
                    q += ue[r] * acc;
                }
                int64_t e = i + (int64_t)nx * (j + (int64_t)ny * k);
                dc[e] = -p * pow(rho[e], p - 1.0) * (E0 - Emin) * q;
                energy += 0.5 * E[e] * q;
            }
    return energy;
}

/* Sigmund sensitivity filter on the structured grid: neighbours within R of the cell centre */
void cpu_filter_sens(int nx, int ny, int nz, double hx, double hy, double hz, double R, const double* rho, const double* dc,
                     double* out) {
    const int ri = (int)floor(R / hx + 1e-9), rj = (int)floor(R / hy + 1e-9), rk = (int)floor(R / hz + 1e-9);
    const double vol = hx * hy * hz;
#pragma omp parallel for schedule(static)
    for (int k = 0; k < nz; ++k)
        for (int j = 0; j < ny; ++j)
            for (int i = 0; i < nx; ++i) {
                double num = 0.0, den = 0.0;
                for (int dk = -rk; dk <= rk; ++dk)
                    for (int dj = -rj; dj <= rj; ++dj)
                        for (int di = -ri; di <= ri; ++di) {
                            int ii = i + di, jj = j + dj, kk = k + dk;
                            if (ii < 0 || ii >= nx || jj < 0 || jj >= ny || kk < 0 || kk >= nz) continue;
                            double dx = di * hx, dy = dj * hy, dz = dk * hz;
                            double w = R - sqrt(dx * dx + dy * dy + dz * dz);
                            if (w > 0.0) {
                                int64_t nb = ii + (int64_t)nx * (jj + (int64_t)ny * kk);
                                num += w * rho[nb] * dc[nb] / vol;
                                den += w;
                            }
                        }
                int64_t e = i + (int64_t)nx * (j + (int64_t)ny * k);
                double rs = rho[e] > 1e-3 ? rho[e] : 1e-3;
                out[e] = den > 1e-12 ? num / (rs / vol * den) : dc[e];
            }
}

/* one OC bisection step: candidate densities for lambda, returns dot(new, V) (uniform V) */
double cpu_oc_step(int64_t n, const double* rho, const double* dc, const double* dv, double lam, double move, double damping,
                   double vol, double* rho_new) {
    double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        double Be = fabs(dc[i]) / (lam * dv[i]);
        double cand = rho[i] * pow(Be, damping);
        double v = fmax(1e-3, fmax(rho[i] - move, fmin(1.0, fmin(rho[i] + move, cand))));
        rho_new[i] = v;
        s += v * vol;
    }
    return s;
}

void cpu_modulus(int64_t n, const double* rho, double E0, double Emin, double p, double* E) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) E[i] = Emin + (E0 - Emin) * pow(rho[i], p);
}
