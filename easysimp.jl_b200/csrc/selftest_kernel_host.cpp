// selftest_kernel_host.cpp -- runs the fused hex8 operator kernel (fused_kernel.cuh, the very source nvcc compiles into
// libeasysimp_b200.so) on the CPU: every thread of a CUDA block is a host thread, __syncthreads is a barrier,
// cp.async copies complete at once, blocks run one after the other.  TEST INFRASTRUCTURE for the -m "not gpu"
// suite (tests/test_kernel_emulation.py): not linked into the product library, nothing in the product calls it.
// It checks the kernel's LOGIC (tiling, ring indexing, butterflies, store modes, float instantiation); it cannot
// prove the absence of races that depend on the real completion order of asynchronous copies.
#define ESIMP_HOST_EMULATION 1
#include <algorithm>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <thread>
#include <type_traits>
#include <vector>

using std::fma;
using std::max;
using std::min;

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__
#define __align__(n)

struct EmuDim3 {
    unsigned x = 1, y = 1, z = 1;
};
static thread_local EmuDim3 threadIdx;
static EmuDim3 blockIdx, blockDim, gridDim;
alignas(16) unsigned char smem_raw[256 * 1024];  // the block's dynamic shared memory (blocks run sequentially)
static std::barrier<>* g_barrier = nullptr;
static std::mutex g_mutex;
inline void __syncthreads() { g_barrier->arrive_and_wait(); }
template <typename T>
inline void atomicAdd(T* p, T v) {
    std::lock_guard<std::mutex> lock(g_mutex);
    *p += v;
}

double c_B[8][3][3];
float c_Bf[8][3][3];

// ---- cp.async model: immediate (copies land at issue) or deferred (they land when wait_group retires their group)
static bool g_deferred_copies = false;
struct EmuCopy {
    void* dst;
    const void* src;
    int bytes;
    bool valid;
};
static thread_local std::vector<std::vector<EmuCopy>> t_groups;  // committed groups, oldest first
static thread_local std::vector<EmuCopy> t_open;                 // copies issued since the last commit
inline void emu_apply(const EmuCopy& c) {
    if (c.valid)
        memcpy(c.dst, c.src, c.bytes);
    else
        memset(c.dst, 0, c.bytes);
}
inline void emu_cp_async(void* dst, const void* src, int bytes, bool valid) {
    EmuCopy c{dst, src, bytes, valid};
    if (g_deferred_copies)
        t_open.push_back(c);
    else
        emu_apply(c);
}
inline void emu_cp_commit() {
    if (!g_deferred_copies) return;
    t_groups.push_back(t_open);
    t_open.clear();
}
inline void emu_cp_wait(int n_pending_allowed) {  // cp.async.wait_group N: all but the N newest groups are complete
    if (!g_deferred_copies) return;
    while ((int)t_groups.size() > n_pending_allowed) {
        for (const EmuCopy& c : t_groups.front()) emu_apply(c);
        t_groups.erase(t_groups.begin());
    }
}

#include "grid_dev.h"

// the fused dot product: every thread of every block contributes its partial sum
template <int NV, bool IS_MAX>
inline void grid_reduce_idx(double (&v)[NV], double*, unsigned int*, const RedOut<NV>& out, unsigned int, unsigned int) {
    std::lock_guard<std::mutex> lock(g_mutex);
    for (int i = 0; i < NV; ++i) *out.p[i] = IS_MAX ? std::max(*out.p[i], v[i]) : *out.p[i] + v[i];
}

template <int NV, bool IS_MAX>
inline void grid_reduce(double (&v)[NV], double* partial, unsigned int* ticket, const RedOut<NV>& out) {
    grid_reduce_idx<NV, IS_MAX>(v, partial, ticket, out, 0u, 1u);
}

#include "fused_kernel.cuh"
#include "mg32_kernels.cuh"

namespace {

template <typename T, int MODE, int TY = 16, int TZ = 16>
void run_grid(const GridDev& g, int chunk, const T* E, const T* u, T* y, double* dot, const FuseArgsT<T>& fa) {
    constexpr int SY = TY - 1, SZ = TZ - 1;
    const int nty = (g.ny + 1 + SY - 1) / SY, ntz = (g.nz + 1 + SZ - 1) / SZ;
    const int nplanes = g.NPX;
    if (chunk <= 0) chunk = nplanes;
    const int nchunk = (nplanes + chunk - 1) / chunk;
    gridDim = EmuDim3{(unsigned)ntz, (unsigned)nty, (unsigned)nchunk};
    blockDim = EmuDim3{(unsigned)(TY * TZ), 1, 1};
    RedOut<1> out{{dot}};
    for (int bz = 0; bz < nchunk; ++bz)
        for (int by = 0; by < nty; ++by)
            for (int bx = 0; bx < ntz; ++bx) {
                blockIdx = EmuDim3{(unsigned)bx, (unsigned)by, (unsigned)bz};
                std::barrier<> barrier(TY * TZ);
                g_barrier = &barrier;
                std::vector<std::thread> threads;
                threads.reserve(TY * TZ);
                for (int t = 0; t < TY * TZ; ++t)
                    threads.emplace_back([&, t] {
                        threadIdx = EmuDim3{(unsigned)t, 0, 0};
                        k_matvec_fused<T, TY, TZ, false, 2, MODE>(g, 0, g.NEX, 0, g.NPX, chunk, E, u, y, nullptr, nullptr, out, fa);
                    });
                for (auto& th : threads) th.join();
            }
}

template <typename T>
int run_mode(int mode, const GridDev& g, int chunk, const T* E, const T* u, T* y, double* dot, const FuseArgsT<T>& fa) {
    switch (mode) {
        case FUSE_PLAIN: run_grid<T, FUSE_PLAIN>(g, chunk, E, u, y, dot, fa); return 0;
        case FUSE_DOT: run_grid<T, FUSE_DOT>(g, chunk, E, u, y, dot, fa); return 0;
        case FUSE_SMOOTH: run_grid<T, FUSE_SMOOTH>(g, chunk, E, u, y, dot, fa); return 0;
        case FUSE_RESID: run_grid<T, FUSE_RESID>(g, chunk, E, u, y, dot, fa); return 0;
        case FUSE_SMOOTH_DOT: run_grid<T, FUSE_SMOOTH_DOT>(g, chunk, E, u, y, dot, fa); return 0;
    }
    return -1;
}

GridDev level_grid(int nx, int ny, int nz) {
    GridDev g;
    g.nx = nx;
    g.ny = ny;
    g.nz = nz;
    g.NPX = nx + 1;
    g.NEX = nx;
    g.gp0 = g.ge0 = g.e2p = 0;
    g.nzp = nz + 1;  // the emulation keeps the unpadded pitch: the tests build their device-layout arrays with numpy reshapes
    g.plane = (long long)(ny + 1) * g.nzp;
    g.eplane = (long long)ny * nz;
    return g;
}

// the grid-stride kernels of mg32_kernels.cuh run as ONE thread of a 1x1 grid: it strides over every element
void single_thread() {
    gridDim = blockDim = EmuDim3{1, 1, 1};
    blockIdx = EmuDim3{0, 0, 0};
    threadIdx = EmuDim3{0, 0, 0};
}

}  // namespace

extern "C" {

// ---- the FP32 V-cycle's vector / transfer kernels (mg32_kernels.cuh), single rank, device layouts
void emu_k32_restrict(int fnx, int fny, int fnz, const float* resf, const uint8_t* maskc, float* rc) {
    single_thread();
    GridDev gf = level_grid(fnx, fny, fnz), gc = level_grid(fnx / 2, fny / 2, fnz / 2);
    k32_restrict(gf, gc, 0, gc.NPX, resf, maskc, rc);
}
void emu_k32_prolong_add(int fnx, int fny, int fnz, const float* zc, const uint8_t* maskf, float* zf) {
    single_thread();
    GridDev gf = level_grid(fnx, fny, fnz), gc = level_grid(fnx / 2, fny / 2, fnz / 2);
    k32_prolong_add(gf, gc, 0, gf.NPX, zc, maskf, zf);
}
void emu_k32_cheb_first(long long n, const float* r, const float* dinv, float inv_theta, float* d, float* z) {
    single_thread();
    k32_cheb_first(r, dinv, inv_theta, d, z, 0, n);
}
void emu_k32_enter(long long n, const double* r, double inv_s, const float* dinv, float inv_theta, float* r32, float* d, float* z) {
    single_thread();
    k32_enter(r, inv_s, dinv, inv_theta, r32, d, z, 0, n);
}
void emu_k32_fixup(const long long* fixed, long long n, float m, const float* in, float* out) {
    single_thread();
    k32_fixup(fixed, n, m, in, out);
}
void emu_k32_convert(long long n, const double* src, float* dst) {
    single_thread();
    k32_convert(src, dst, n);
}
double emu_k32_dot_rz(long long n, const double* r, const float* z, double s) {
    single_thread();
    double acc = 0.0;
    RedOut<1> out{{&acc}};
    k32_dot_rz(r, z, s, 0, n, nullptr, nullptr, out);
    return acc;
}
double emu_k32_update(long long n, const double* q, double* r, double rz, double pq) {
    single_thread();
    double acc = 0.0;
    RedOut<1> out{{&acc}};
    k32_update(q, r, 0, n, &rz, &pq, nullptr, nullptr, out);
    return acc;
}
void emu_k32_init(long long n, const double* f, const double* q, double* r, double* rr_ff) {
    single_thread();
    rr_ff[0] = rr_ff[1] = 0.0;
    RedOut<2> out{{rr_ff, rr_ff + 1}};
    k32_init(f, q, r, 0, n, nullptr, nullptr, out);
}
void emu_k32_direction(long long n, const float* z, double s, double* p, double* x, double rz_new, double rz_old, double pq, int mode) {
    single_thread();
    k32_direction(z, s, p, x, 0, n, &rz_new, &rz_old, &pq, mode);
}

// plain operator with the other owner-computes tile shapes the library can select (ESIMP_MATVEC_FUSED_8x16 / _8x32)
int emu_fused_tile(int ty, int tz, int nx, int ny, int nz, int chunk, const double* blocks, const double* E, const double* u, double* y) {
    GridDev g;
    g.nx = nx;
    g.ny = ny;
    g.nz = nz;
    g.NPX = nx + 1;
    g.NEX = nx;
    g.gp0 = g.ge0 = g.e2p = 0;
    g.nzp = nz + 1;  // the emulation keeps the unpadded pitch: the tests build their device-layout arrays with numpy reshapes
    g.plane = (long long)(ny + 1) * g.nzp;
    g.eplane = (long long)ny * nz;
    for (int q = 0; q < 8; ++q)
        for (int c = 0; c < 3; ++c)
            for (int e = 0; e < 3; ++e) c_B[q][c][e] = blocks[(q * 3 + c) * 3 + e];
    double dummy = 0.0;
    FuseArgsT<double> fa;
    if (ty == 8 && tz == 16) {
        run_grid<double, FUSE_PLAIN, 8, 16>(g, chunk, E, u, y, &dummy, fa);
    } else if (ty == 8 && tz == 32) {
        run_grid<double, FUSE_PLAIN, 8, 32>(g, chunk, E, u, y, &dummy, fa);
    } else if (ty == 16 && tz == 16) {
        run_grid<double, FUSE_PLAIN, 16, 16>(g, chunk, E, u, y, &dummy, fa);
    } else {
        return -1;
    }
    return 0;
}


int emu_pat(int q, int c, int d) { return pat(q, c, d) ? 1 : 0; }
void emu_set_deferred_copies(int on) { g_deferred_copies = on != 0; }

// plain FP64 operator on ONE x-slab of a partitioned grid: the caller passes the slab geometry esimp_set_grid_structured
// builds (gp0, NPX, ge0, NEX, owned planes [ps0, ps1), visited layers [le0, le1)) and the slab's local arrays
int emu_fused_slab(int nx, int ny, int nz, int gp0, int npx, int ge0, int nex, int ps0, int ps1, int le0, int le1, int chunk,
                   const double* blocks, const double* E_loc, const double* u_loc, double* y_loc) {
    GridDev g;
    g.nx = nx;
    g.ny = ny;
    g.nz = nz;
    g.NPX = npx;
    g.NEX = nex;
    g.gp0 = gp0;
    g.ge0 = ge0;
    g.e2p = ge0 - gp0;
    g.nzp = nz + 1;  // the emulation keeps the unpadded pitch: the tests build their device-layout arrays with numpy reshapes
    g.plane = (long long)(ny + 1) * g.nzp;
    g.eplane = (long long)ny * nz;
    for (int q = 0; q < 8; ++q)
        for (int c = 0; c < 3; ++c)
            for (int e = 0; e < 3; ++e) c_B[q][c][e] = blocks[(q * 3 + c) * 3 + e];
    constexpr int TY = 16, TZ = 16, SY = TY - 1, SZ = TZ - 1;
    const int nty = (ny + 1 + SY - 1) / SY, ntz = (nz + 1 + SZ - 1) / SZ;
    const int nplanes = ps1 - ps0;
    if (chunk <= 0) chunk = nplanes;
    const int nchunk = (nplanes + chunk - 1) / chunk;
    gridDim = EmuDim3{(unsigned)ntz, (unsigned)nty, (unsigned)nchunk};
    blockDim = EmuDim3{(unsigned)(TY * TZ), 1, 1};
    double dummy = 0.0;
    RedOut<1> out{{&dummy}};
    FuseArgsT<double> fa;
    for (int bz = 0; bz < nchunk; ++bz)
        for (int by = 0; by < nty; ++by)
            for (int bx = 0; bx < ntz; ++bx) {
                blockIdx = EmuDim3{(unsigned)bx, (unsigned)by, (unsigned)bz};
                std::barrier<> barrier(TY * TZ);
                g_barrier = &barrier;
                std::vector<std::thread> threads;
                threads.reserve(TY * TZ);
                for (int t = 0; t < TY * TZ; ++t)
                    threads.emplace_back([&, t] {
                        threadIdx = EmuDim3{(unsigned)t, 0, 0};
                        k_matvec_fused<double, TY, TZ, false, 2, FUSE_PLAIN>(g, le0, le1, ps0, ps1, chunk, E_loc, u_loc, y_loc, nullptr, nullptr,
                                                                             out, fa);
                    });
                for (auto& th : threads) th.join();
            }
    return 0;
}

// One launch of the fused kernel on an nx x ny x nz single-rank grid.  All vectors in the library's device layout
// (node vectors [plane][component][y][z], cell array [layer][y][z]); `blocks` = the 8x3x3 mode-space blocks of the
// cuboid; r / dinv / d are the smoother operands (d is updated in place), unused ones may be NULL.
int emu_fused(int is_float, int mode, int nx, int ny, int nz, int chunk, const double* blocks, const double* E, const double* u,
              const double* r, const double* dinv, double* d, double c1, double c2, double* y, double* dot) {
    GridDev g;
    g.nx = nx;
    g.ny = ny;
    g.nz = nz;
    g.NPX = nx + 1;
    g.NEX = nx;
    g.gp0 = g.ge0 = g.e2p = 0;
    g.nzp = nz + 1;  // the emulation keeps the unpadded pitch: the tests build their device-layout arrays with numpy reshapes
    g.plane = (long long)(ny + 1) * g.nzp;
    g.eplane = (long long)ny * nz;
    const long long nd = 3 * g.NPX * g.plane, nc = (long long)g.NEX * g.eplane;
    for (int q = 0; q < 8; ++q)
        for (int c = 0; c < 3; ++c)
            for (int e = 0; e < 3; ++e) {
                c_B[q][c][e] = blocks[(q * 3 + c) * 3 + e];
                c_Bf[q][c][e] = (float)c_B[q][c][e];
            }
    if (dot) *dot = 0.0;
    double dummy = 0.0;
    if (!is_float) {
        FuseArgsT<double> fa;
        fa.r = r;
        fa.dinv = dinv;
        fa.d = d;
        fa.c1 = c1;
        fa.c2 = c2;
        return run_mode<double>(mode, g, chunk, E, u, y, dot ? dot : &dummy, fa);
    }
    std::vector<float> Ef(E, E + nc), uf(u, u + nd), yf(nd), rf, df, dinvf;
    for (long long i = 0; i < nd; ++i) yf[i] = (float)y[i];
    FuseArgsT<float> fa;
    if (r) {
        rf.assign(r, r + nd);
        fa.r = rf.data();
    }
    if (dinv) {
        dinvf.assign(dinv, dinv + nd);
        fa.dinv = dinvf.data();
    }
    if (d) {
        df.assign(d, d + nd);
        fa.d = df.data();
    }
    fa.c1 = (float)c1;
    fa.c2 = (float)c2;
    int st = run_mode<float>(mode, g, chunk, Ef.data(), uf.data(), yf.data(), dot ? dot : &dummy, fa);
    for (long long i = 0; i < nd; ++i) y[i] = yf[i];
    if (d)
        for (long long i = 0; i < nd; ++i) d[i] = df[i];
    return st;
}
}
