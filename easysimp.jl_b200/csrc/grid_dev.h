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
    int nzp;               // pitch of a node row in doubles: nz + 1 rounded up to a multiple of 8, so that every row starts on a
                           // 64-byte DRAM atom and TMA boxes (16-byte strides) can address the node vectors; the pad entries of
                           // every node vector are zero and stay zero (no kernel writes them, flat vector kernels map 0 to 0)
    long long plane;       // (ny+1)*nzp
    long long eplane;      // ny*nz
    ESIMP_HD long long dof(int lp, int c, int j, int k) const {
        return ((long long)(lp * 3 + c) * (ny + 1) + j) * nzp + k;
    }
    ESIMP_HD long long node(int lp, int j, int k) const {
        return ((long long)lp * (ny + 1) + j) * nzp + k;
    }
    ESIMP_HD long long cell(int le, int j, int k) const {
        return ((long long)le * ny + j) * nz + k;
    }
};

// node-row pitch for nz cells per row
ESIMP_HD int esimp_node_pitch(int nz) { return (nz + 1 + 7) & ~7; }

// where the fused Chebyshev step of the operator kernels takes the previous direction d from (FuseArgsT::dmode)
enum { FUSE_D_NONE = 0, FUSE_D_BUFFER = 1, FUSE_D_IS_Z = 2, FUSE_D_FROM_ZPREV = 3 };

// where a grid reduction leaves its NV results (device pointers)
template <int NV>
struct RedOut {
    double* p[NV];
};
