// grid_dev.h -- slab geometry and the reduction-output handle, free of CUDA includes so that the host self-test
// (selftest_kernel_host.cpp) compiles the very same definitions as the kernels.
#pragma once

#ifndef ESIMP_HD
#ifdef __CUDACC__
#define ESIMP_HD __host__ __device__ __forceinline__
#else
#define ESIMP_HD inline
#endif
#endif

// geometry of the slab this rank holds (all ints fit 32 bit; flat indices are 64 bit)
struct GridDev {
    int nx, ny, nz;        // global element counts
    int NPX;               // local node planes stored
    int NEX;               // local element layers stored
    int gp0;               // global x index of local node plane 0
    int ge0;               // global x index of local element layer 0
    int e2p;               // local plane of the low face of local layer le = le + e2p
    long long plane;       // (ny+1)*(nz+1)
    long long eplane;      // ny*nz
    ESIMP_HD long long dof(int lp, int c, int j, int k) const {
        return ((long long)(lp * 3 + c) * (ny + 1) + j) * (nz + 1) + k;
    }
    ESIMP_HD long long node(int lp, int j, int k) const {
        return ((long long)lp * (ny + 1) + j) * (nz + 1) + k;
    }
    ESIMP_HD long long cell(int le, int j, int k) const {
        return ((long long)le * ny + j) * nz + k;
    }
};

// where a grid reduction leaves its NV results (device pointers)
template <int NV>
struct RedOut {
    double* p[NV];
};
