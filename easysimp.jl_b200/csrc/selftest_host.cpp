// selftest_host.cpp -- CPU build of the per-cell stress arithmetic in stress_math.cuh (the same source the
// kernels of post_load.cu compile), so that the -m "not gpu" suite can hold it against the oracle without a
// device.  TEST INFRASTRUCTURE: not linked into libeasysimp_b200.so, nothing in the product calls it.
#include "stress_math.cuh"

extern "C" {

// general cell from its node coordinates: npe = 8 (hex, VTK order) or 4 (tet); out3 = vm(q1), vm(avg), energy
int selftest_cell_stress(int npe, const double* xyz, const double* ue_flat, double E, double nu, double* sig_out, double* out3) {
    double lam, mu;
    lame_of(E, nu, lam, mu);
    CellStress r;
    if (npe == 8) {
        double x[8][3], ue[8][3];
        for (int a = 0; a < 8; ++a)
            for (int c = 0; c < 3; ++c) {
                x[a][c] = xyz[3 * a + c];
                ue[a][c] = ue_flat[3 * a + c];
            }
        r = hex_cell_stress(x, ue, lam, mu, sig_out);
    } else if (npe == 4) {
        double x[4][3], ue[4][3];
        for (int a = 0; a < 4; ++a)
            for (int c = 0; c < 3; ++c) {
                x[a][c] = xyz[3 * a + c];
                ue[a][c] = ue_flat[3 * a + c];
            }
        r = tet_cell_stress(x, ue, lam, mu, sig_out);
    } else {
        return -1;
    }
    out3[0] = r.vm_q1;
    out3[1] = r.vm_avg;
    out3[2] = r.energy;
    return 0;
}

// cuboid cell through the gradient table the structured kernels keep in constant memory
int selftest_cuboid_stress(double hx, double hy, double hz, const double* ue_flat, double E, double nu, double* sig_out, double* out3) {
    double G[8][8][3], ue[8][3], lam, mu;
    cuboid_gradient_table(hx, hy, hz, G);
    lame_of(E, nu, lam, mu);
    for (int a = 0; a < 8; ++a)
        for (int c = 0; c < 3; ++c) ue[a][c] = ue_flat[3 * a + c];
    const CellStress r = cuboid_cell_stress(ue, G, lam, mu, hx * hy * hz / 8.0, sig_out);
    out3[0] = r.vm_q1;
    out3[1] = r.vm_avg;
    out3[2] = r.energy;
    return 0;
}
}
