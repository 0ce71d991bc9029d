// stress_math.cuh -- per-cell stress arithmetic shared by the kernels of post_load.cu and by the host
// self-test (selftest_host.cu, built for the CPU so that the -m "not gpu" suite checks this very source
// against the oracle).  Restates function_gradient / symmetric / constitutive_relation / dev as used by
// calculate_stress_at_quadrature_points and calculate_stresses_simp (FiniteElementAnalysis.jl:537-597).
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define ESIMP_HD __host__ __device__ __forceinline__
#else
#define ESIMP_HD inline
#endif

// reference-cell sign of VTK hexahedron node a along x, y, z (0 = low face, 1 = high face)
ESIMP_HD int hex_sx(int a) { return ((a + 1) >> 1) & 1; }
ESIMP_HD int hex_sy(int a) { return (a >> 1) & 1; }
ESIMP_HD int hex_sz(int a) { return a >> 2; }


// physical shape-function gradients at the 8 Gauss points of a cuboid hx*hy*hz
// (QuadratureRule{RefHexahedron}(2): first coordinate fastest), G[q][a][d]
inline void cuboid_gradient_table(double hx, double hy, double hz, double (&G)[8][8][3]) {
    const double gp = 1.0 / sqrt(3.0);
    const double h[3] = {hx, hy, hz};
    for (int q = 0; q < 8; ++q) {
        const double xi[3] = {(q & 1) ? gp : -gp, (q & 2) ? gp : -gp, (q & 4) ? gp : -gp};
        for (int a = 0; a < 8; ++a) {
            const double s[3] = {hex_sx(a) ? 1.0 : -1.0, hex_sy(a) ? 1.0 : -1.0, hex_sz(a) ? 1.0 : -1.0};
            const double f[3] = {1.0 + s[0] * xi[0], 1.0 + s[1] * xi[1], 1.0 + s[2] * xi[2]};
            G[q][a][0] = 0.125 * s[0] * f[1] * f[2] * (2.0 / h[0]);
            G[q][a][1] = 0.125 * f[0] * s[1] * f[2] * (2.0 / h[1]);
            G[q][a][2] = 0.125 * f[0] * f[1] * s[2] * (2.0 / h[2]);
        }
    }
}

// Stress at one quadrature point from the nodal displacements ue[a][c] and the physical gradients gr[a][d].
// sig in SymmetricTensor{2,3} storage order (11, 21, 31, 22, 32, 33); returns sigma : eps.
template <int NPE, typename GR>
ESIMP_HD double qp_stress(const double (&ue)[NPE][3], const GR& gr, double lam, double mu, double (&sig)[6]) {
    double G[3][3] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};  // G[c][d] = d u_c / d x_d (function_gradient, :543)
#pragma unroll
    for (int a = 0; a < NPE; ++a)
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int d = 0; d < 3; ++d) G[c][d] = fma(ue[a][c], gr(a, d), G[c][d]);
    const double exx = G[0][0], eyy = G[1][1], ezz = G[2][2];
    const double exy = 0.5 * (G[0][1] + G[1][0]), exz = 0.5 * (G[0][2] + G[2][0]), eyz = 0.5 * (G[1][2] + G[2][1]);
    const double ltr = lam * (exx + eyy + ezz), m2 = 2.0 * mu;
    sig[0] = ltr + m2 * exx;  // constitutive_relation, FiniteElementAnalysis.jl:63-65
    sig[1] = m2 * exy;
    sig[2] = m2 * exz;
    sig[3] = ltr + m2 * eyy;
    sig[4] = m2 * eyz;
    sig[5] = ltr + m2 * ezz;
    return sig[0] * exx + sig[3] * eyy + sig[5] * ezz + 2.0 * (sig[1] * exy + sig[2] * exz + sig[4] * eyz);
}

ESIMP_HD double von_mises6(const double (&s)[6]) {
    const double m = (s[0] + s[3] + s[5]) / 3.0;
    const double a = s[0] - m, b = s[3] - m, c = s[5] - m;
    return sqrt(1.5 * (a * a + b * b + c * c + 2.0 * (s[1] * s[1] + s[2] * s[2] + s[4] * s[4])));
}

struct GradTable {  // cuboid: gradients of one Gauss point from a [8][3] table (constant memory on the device)
    const double (*t)[3];
    ESIMP_HD double operator()(int a, int d) const { return t[a][d]; }
};
template <int NPE>
struct GradLocal {
    double g[NPE][3];
    ESIMP_HD double operator()(int a, int d) const { return g[a][d]; }
};

// physical gradients and |det J| of a general hexahedron at reference point xi
ESIMP_HD double hex_gradients(const double (&x)[8][3], const double (&xi)[3], GradLocal<8>& out) {
    double dN[8][3], J[3][3] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        const double s0 = hex_sx(a) ? 1.0 : -1.0, s1 = hex_sy(a) ? 1.0 : -1.0, s2 = hex_sz(a) ? 1.0 : -1.0;
        const double f0 = 1.0 + s0 * xi[0], f1 = 1.0 + s1 * xi[1], f2 = 1.0 + s2 * xi[2];
        dN[a][0] = 0.125 * s0 * f1 * f2;
        dN[a][1] = 0.125 * f0 * s1 * f2;
        dN[a][2] = 0.125 * f0 * f1 * s2;
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int d = 0; d < 3; ++d) J[c][d] = fma(x[a][c], dN[a][d], J[c][d]);  // J[c][d] = d x_c / d xi_d
    }
    const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1], c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2],
                 c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
    const double det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
    const double id = 1.0 / det;
    double Ji[3][3];  // inverse of J
    Ji[0][0] = c00 * id;
    Ji[1][0] = c01 * id;
    Ji[2][0] = c02 * id;
    Ji[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * id;
    Ji[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * id;
    Ji[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * id;
    Ji[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * id;
    Ji[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * id;
    Ji[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * id;
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int d = 0; d < 3; ++d) out.g[a][d] = dN[a][0] * Ji[0][d] + dN[a][1] * Ji[1][d] + dN[a][2] * Ji[2][d];
    return fabs(det);
}

// constant gradients and volume of a linear tetrahedron
ESIMP_HD double tet_gradients(const double (&x)[4][3], GradLocal<4>& out) {
    double J[3][3];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int d = 0; d < 3; ++d) J[c][d] = x[d + 1][c] - x[0][c];
    const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1], c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2],
                 c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
    const double det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
    const double id = 1.0 / det;
    double Ji[3][3];
    Ji[0][0] = c00 * id;
    Ji[1][0] = c01 * id;
    Ji[2][0] = c02 * id;
    Ji[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * id;
    Ji[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * id;
    Ji[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * id;
    Ji[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * id;
    Ji[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * id;
    Ji[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * id;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        out.g[1][d] = Ji[0][d];  // dN/dxi = [[-1,-1,-1],[1,0,0],[0,1,0],[0,0,1]]
        out.g[2][d] = Ji[1][d];
        out.g[3][d] = Ji[2][d];
        out.g[0][d] = -(Ji[0][d] + Ji[1][d] + Ji[2][d]);
    }
    return fabs(det) / 6.0;
}

struct CellStress {
    double vm_q1, vm_avg, energy;
};

// all Gauss points of one cuboid cell, gradients from table[q][a][d]; sig_out (48 doubles, [q][6]) optional
ESIMP_HD CellStress cuboid_cell_stress(const double (&ue)[8][3], const double (*table)[8][3], double lam, double mu, double dvol,
                                       double* sig_out) {
    double avg[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0}, W = 0.0, vm1 = 0.0;
#pragma unroll 1
    for (int q = 0; q < 8; ++q) {
        double sig[6];
        W += qp_stress<8>(ue, GradTable{table[q]}, lam, mu, sig) * dvol;
        if (q == 0) vm1 = von_mises6(sig);
#pragma unroll
        for (int t = 0; t < 6; ++t) {
            avg[t] += sig[t];
            if (sig_out) sig_out[q * 6 + t] = sig[t];
        }
    }
#pragma unroll
    for (int t = 0; t < 6; ++t) avg[t] *= 0.125;
    return CellStress{vm1, von_mises6(avg), 0.5 * W};
}

ESIMP_HD CellStress hex_cell_stress(const double (&x)[8][3], const double (&ue)[8][3], double lam, double mu,
                                                       double* sig_out) {
    const double gp = 0.57735026918962576451;  // 1/sqrt(3)
    double avg[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0}, W = 0.0, vm1 = 0.0;
#pragma unroll 1
    for (int q = 0; q < 8; ++q) {
        const double xi[3] = {(q & 1) ? gp : -gp, (q & 2) ? gp : -gp, (q & 4) ? gp : -gp};
        GradLocal<8> gr;
        const double dvol = hex_gradients(x, xi, gr);
        double sig[6];
        W += qp_stress<8>(ue, gr, lam, mu, sig) * dvol;
        if (q == 0) vm1 = von_mises6(sig);
#pragma unroll
        for (int t = 0; t < 6; ++t) {
            avg[t] += sig[t];
            if (sig_out) sig_out[q * 6 + t] = sig[t];
        }
    }
#pragma unroll
    for (int t = 0; t < 6; ++t) avg[t] *= 0.125;
    return CellStress{vm1, von_mises6(avg), 0.5 * W};
}

ESIMP_HD CellStress tet_cell_stress(const double (&x)[4][3], const double (&ue)[4][3], double lam, double mu,
                                                       double* sig_out) {
    GradLocal<4> gr;
    const double vol = tet_gradients(x, gr);
    double sig[6];
    const double W = qp_stress<4>(ue, gr, lam, mu, sig) * vol;
    if (sig_out)
#pragma unroll
        for (int t = 0; t < 6; ++t) sig_out[t] = sig[t];
    const double vm = von_mises6(sig);
    return CellStress{vm, vm, 0.5 * W};
}

ESIMP_HD void lame_of(double E, double nu, double& lam, double& mu) {
    lam = E * nu / ((1.0 + nu) * (1.0 - 2.0 * nu));  // compute_lame_parameters, FiniteElementAnalysis.jl:52-56
    mu = E / (2.0 * (1.0 + nu));
}

