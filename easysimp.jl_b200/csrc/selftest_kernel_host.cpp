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
// fused smoother step: source of the previous direction (FUSE_D_* of grid_dev.h: 1 = the d vector), store of the new one, z_prev
static int g_dmode = 1, g_write_d = 1;
static const double* g_zprev = nullptr;
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

// ---------------------------------------------------------------------------------------------------------------------
// emulation of what sweep_kernel.cuh uses beyond plain CUDA C++: mbarriers (phase / pending arrivals / transaction bytes),
// 3-D TMA boxes with out-of-bounds zero fill, warp shuffles and warp barriers.  Two TMA models: the box lands at issue,
// or a DMA thread delivers it a little later (a consumer that reads before its mbarrier wait then sees stale data).
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <memory>
struct SwBar {
    std::atomic<unsigned> phase{0};
    int count = 0, pending = 0;
    long long tx = 0;
};
static std::mutex g_bar_mutex;
static void emu_bar_check(SwBar* b) {
    if (b->pending == 0 && b->tx == 0) {
        b->pending = b->count;
        b->phase.fetch_add(1, std::memory_order_release);
    }
}
inline void sw_bar_init(SwBar* b, int count) {
    new (b) SwBar();
    b->count = b->pending = count;
}
inline void sw_fence_init() {}
inline void sw_bar_arrive(SwBar* b) {
    std::lock_guard<std::mutex> lock(g_bar_mutex);
    b->pending--;
    emu_bar_check(b);
}
inline void sw_bar_expect_tx(SwBar* b, unsigned bytes) {
    std::lock_guard<std::mutex> lock(g_bar_mutex);
    b->tx += bytes;
    b->pending--;
    emu_bar_check(b);
}
inline void emu_bar_complete_tx(SwBar* b, long long bytes) {
    std::lock_guard<std::mutex> lock(g_bar_mutex);
    b->tx -= bytes;
    emu_bar_check(b);
}
inline void sw_bar_wait(SwBar* b, unsigned parity) {
    while ((b->phase.load(std::memory_order_acquire) & 1u) == (parity & 1u)) std::this_thread::yield();
}
typedef SwBar* SwBarAddr;
inline SwBarAddr sw_bar_addr(const SwBar* b) { return const_cast<SwBar*>(b); }
inline SwBarAddr sw_bar_at(SwBarAddr base, int index) { return base + index; }
inline void sw_bar_arrive_a(SwBarAddr b) { sw_bar_arrive(b); }
inline void sw_bar_wait_a(SwBarAddr b, unsigned parity) { sw_bar_wait(b, parity); }

struct SwTensorMap {  // what cuTensorMapEncodeTiled would hold for a rank-3 FP64 tensor (strides in elements)
    const double* base = nullptr;
    long long dim[3] = {0, 0, 0}, stride[3] = {1, 0, 0};
    int box[3] = {0, 0, 0};
};
struct EmuTmaJob {
    double* dst;
    SwTensorMap tm;  // by value: the kernel's by-value parameter lives on the issuing host thread's stack
    int c[3];
    SwBar* bar;
};
static bool g_tma_deferred = false;
static std::mutex g_tma_mutex;
static std::deque<EmuTmaJob> g_tma_queue;
static void emu_tma_run(const EmuTmaJob& j) {
    const SwTensorMap& t = j.tm;
    long long n = 0;
    for (int z = 0; z < t.box[2]; ++z)
        for (int y = 0; y < t.box[1]; ++y)
            for (int x = 0; x < t.box[0]; ++x, ++n) {
                const long long gx = j.c[0] + x, gy = j.c[1] + y, gz = j.c[2] + z;
                const bool in = gx >= 0 && gx < t.dim[0] && gy >= 0 && gy < t.dim[1] && gz >= 0 && gz < t.dim[2];
                j.dst[n] = in ? t.base[gx * t.stride[0] + gy * t.stride[1] + gz * t.stride[2]] : 0.0;
            }
    emu_bar_complete_tx(j.bar, n * 8);
}
inline void sw_tma_3d(void* dst, const SwTensorMap* tm, int c0, int c1, int c2, SwBar* bar) {
    EmuTmaJob j{(double*)dst, *tm, {c0, c1, c2}, bar};
    if (!g_tma_deferred) {
        emu_tma_run(j);
    } else {
        std::lock_guard<std::mutex> lock(g_tma_mutex);
        g_tma_queue.push_back(j);
    }
}
inline void sw_tma_prefetch_3d(const SwTensorMap*, int, int, int) {}
// warp-level primitives: one std::barrier and one exchange row per warp of the block
static std::vector<std::unique_ptr<std::barrier<>>> g_warp_barrier;
static double g_shfl_row[64][32];
static std::barrier<>* g_consumer_barrier = nullptr;
inline void sw_syncwarp() { g_warp_barrier[threadIdx.x >> 5]->arrive_and_wait(); }
inline double sw_shfl_up1(double v) {
    const int wp = threadIdx.x >> 5, ln = threadIdx.x & 31;
    g_shfl_row[wp][ln] = v;
    g_warp_barrier[wp]->arrive_and_wait();
    const double r = ln > 0 ? g_shfl_row[wp][ln - 1] : v;
    g_warp_barrier[wp]->arrive_and_wait();
    return r;
}
inline long long sw_double_as_ll(double v) {
    long long b;
    memcpy(&b, &v, 8);
    return b;
}
inline double sw_ll_as_double(long long b) {
    double v;
    memcpy(&v, &b, 8);
    return v;
}
inline void sw_regs_producer() {}
inline void sw_regs_consumer() {}
inline void sw_consumer_barrier() { g_consumer_barrier->arrive_and_wait(); }
#define SW_PIN32(x) (void)(x)
#define SW_PIN64(x) (void)(x)
#define __grid_constant__
alignas(128) unsigned char sw_smem_raw[512 * 1024];
#define SW_SMEM_DECL
#include "sweep_kernel.cuh"
template <int NW>
inline void sw1_consumer_barrier() { g_consumer_barrier->arrive_and_wait(); }
inline void sw1_regs_producer() {}
template <int NW>
inline void sw1_regs_consumer() {}
#include "sweep1_kernel.cuh"

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
        fa.dmode = g_dmode;
        fa.write_d = g_write_d;
        fa.zprev = g_zprev;
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

void emu_set_deferred_tma(int on) { g_tma_deferred = on != 0; }
// how the fused smoother step of both kernels gets its previous direction (FUSE_D_*), whether it stores the new one, and the
// previous iterate for FUSE_D_FROM_ZPREV (device layout of the call that follows; may alias its output vector)
void emu_set_smoother_direction(int dmode, int write_d, const double* zprev) {
    g_dmode = dmode;
    g_write_d = write_d;
    g_zprev = zprev;
}

static double g_ik[3] = {1.0, 1.0, 1.0};
static int g_sweep_nw = 0;  // 0: the two-cells-per-thread kernel (sweep_kernel.cuh); 12 / 16: sweep1_kernel.cuh with that many consumer warps
void emu_sweep_set_variant(int nw) { g_sweep_nw = nw; }
// 1 / K0[3a+c, 3a+c] for the sweep kernel's smoother modes: their `dinv` argument is the NODE scalar array [npx][ny+1][nzp]
// (s = 1 / sum E with the prescribed-component mask in the three lowest mantissa bits), Dinv[node, c] = ik[c] * s
void emu_sweep_set_ik(double k0, double k1, double k2) {
    g_ik[0] = k0;
    g_ik[1] = k1;
    g_ik[2] = k2;
}

// One launch of the sweep kernel (sweep_kernel.cuh) on ONE x-slab (or the whole grid: gp0 = ge0 = 0, npx = nx + 1, nex = nx,
// [ps0, ps1) = [0, nx + 1), [le0, le1) = [0, nx)).  Vectors in the library's device layout with node-row pitch nzp.
int emu_sweep(int mode, int nx, int ny, int nz, int nzp, int gp0, int npx, int ge0, int nex, int ps0, int ps1, int le0, int le1, int chunk,
              const double* blocks, const double* E, const double* u, const double* r, const double* dinv, double* d, double c1, double c2,
              double* y, double* dot) {
    SweepArgs a{};
    GridDev& g = a.g;
    g.nx = nx;
    g.ny = ny;
    g.nz = nz;
    g.NPX = npx;
    g.NEX = nex;
    g.gp0 = gp0;
    g.ge0 = ge0;
    g.e2p = ge0 - gp0;
    g.nzp = nzp;
    g.plane = (long long)(ny + 1) * nzp;
    g.eplane = (long long)ny * nz;
    a.le0 = le0;
    a.le1 = le1;
    a.ps0 = ps0;
    a.ps1 = ps1;
    const int nplanes = ps1 - ps0;
    a.chunk = chunk > 0 ? chunk : nplanes;
    a.y = y;
    a.d_out = d;
    a.c1 = c1;
    a.c2 = c2;
    a.dmode = c1 != 0.0 ? g_dmode : FUSE_D_NONE;
    a.write_d = g_write_d;
    for (int c = 0; c < 3; ++c) a.ik[c] = g_ik[c];
    double dummy = 0.0;
    if (dot) *dot = 0.0;
    a.dot_out = RedOut<1>{{dot ? dot : &dummy}};
    for (int q = 0; q < 8; ++q)
        for (int c = 0; c < 3; ++c)
            for (int e = 0; e < 3; ++e) a.blk.B[q][c][e] = blocks[(q * 3 + c) * 3 + e];
    auto node_map = [&](const double* base, int bz_, int by_) {
        SwTensorMap t;
        t.base = base;
        t.dim[0] = nz + 1;
        t.dim[1] = ny + 1;
        t.dim[2] = 3LL * npx;
        t.stride[0] = 1;
        t.stride[1] = nzp;
        t.stride[2] = g.plane;
        t.box[0] = bz_;
        t.box[1] = by_;
        t.box[2] = 3;
        return t;
    };
    auto scalar_map = [&](const double* base, int bz_, int by_) {  // one double per node: [npx][ny+1][nzp]
        SwTensorMap t = node_map(base, bz_, by_);
        t.dim[2] = npx;
        t.box[2] = 1;
        return t;
    };
    const int nw1 = g_sweep_nw;  // geometry of the kernel that runs
    const int crows = nw1 ? nw1 : SW_CROWS, rows = crows + 1, orows = crows - 1, nthreads = nw1 ? 32 * (nw1 + 4) : SW_THREADS, ncons = nw1 ? nw1 : SW_NW;
    if (nw1 != 0 && nw1 != 12 && nw1 != 16) return -1;
    SwTensorMap tmU = node_map(u, SW_BOXZ, rows), tmR = node_map(r, 32, orows), tmDinv = scalar_map(dinv, 32, orows),
                tmD = node_map(g_dmode == FUSE_D_FROM_ZPREV ? g_zprev : d, 32, orows), tmRx = node_map(r, 2, orows), tmDinvx = scalar_map(dinv, 2, orows),
                tmDx = node_map(g_dmode == FUSE_D_FROM_ZPREV ? g_zprev : d, 2, orows), tmE;
    tmE.base = E;
    tmE.dim[0] = nz;
    tmE.dim[1] = ny;
    tmE.dim[2] = nex;
    tmE.stride[0] = 1;
    tmE.stride[1] = nz;
    tmE.stride[2] = g.eplane;
    tmE.box[0] = 32;
    tmE.box[1] = crows;
    tmE.box[2] = 1;
    const int nty = (ny + 1 + crows - 2) / (crows - 1), nchunk = (nplanes + a.chunk - 1) / a.chunk;
    gridDim = EmuDim3{(unsigned)nty, (unsigned)nchunk, 1};
    blockDim = EmuDim3{(unsigned)nthreads, 1, 1};
    g_warp_barrier.clear();
    for (int wi = 0; wi < nthreads / 32; ++wi) g_warp_barrier.emplace_back(new std::barrier<>(32));
    for (int by = 0; by < nchunk; ++by)
        for (int bx = 0; bx < nty; ++bx) {
            blockIdx = EmuDim3{(unsigned)bx, (unsigned)by, 0};
            std::barrier<> barrier(nthreads), cbarrier(32 * ncons);
            g_barrier = &barrier;
            g_consumer_barrier = &cbarrier;
            std::atomic<bool> done{false};
            std::thread dma([&] {  // deferred TMA model: boxes land a little after they were issued, in order
                while (true) {
                    EmuTmaJob j;
                    bool have = false;
                    {
                        std::lock_guard<std::mutex> lock(g_tma_mutex);
                        if (!g_tma_queue.empty()) {
                            j = g_tma_queue.front();
                            g_tma_queue.pop_front();
                            have = true;
                        }
                    }
                    if (have) {
                        std::this_thread::sleep_for(std::chrono::microseconds(200));
                        emu_tma_run(j);
                    } else if (done.load()) {
                        break;
                    } else {
                        std::this_thread::yield();
                    }
                }
            });
            std::vector<std::thread> threads;
            threads.reserve(nthreads);
            for (int t = 0; t < nthreads; ++t)
                threads.emplace_back([&, t] {
                    threadIdx = EmuDim3{(unsigned)t, 0, 0};
#define EMU_SW1(M, N) k_matvec_sweep1<M, SweepTraits<M>::RING, N>(a, tmU, tmE, tmR, tmDinv, tmD, tmRx, tmDinvx, tmDx)
#define EMU_SW1_MODES(N)                                  \
    switch (mode) {                                       \
        case FUSE_PLAIN: EMU_SW1(FUSE_PLAIN, N); break;   \
        case FUSE_DOT: EMU_SW1(FUSE_DOT, N); break;       \
        case FUSE_SMOOTH: EMU_SW1(FUSE_SMOOTH, N); break; \
        case FUSE_RESID: EMU_SW1(FUSE_RESID, N); break;   \
        case FUSE_SMOOTH_DOT: EMU_SW1(FUSE_SMOOTH_DOT, N); break; \
    }
                    if (nw1 == 16) {
                        EMU_SW1_MODES(16)
                        return;
                    }
                    if (nw1 == 12) {
                        EMU_SW1_MODES(12)
                        return;
                    }
                    switch (mode) {
                        case FUSE_PLAIN: k_matvec_sweep<FUSE_PLAIN, SweepTraits<FUSE_PLAIN>::RING>(a, tmU, tmE, tmR, tmDinv, tmD, tmRx, tmDinvx, tmDx); break;
                        case FUSE_DOT: k_matvec_sweep<FUSE_DOT, SweepTraits<FUSE_DOT>::RING>(a, tmU, tmE, tmR, tmDinv, tmD, tmRx, tmDinvx, tmDx); break;
                        case FUSE_SMOOTH: k_matvec_sweep<FUSE_SMOOTH, SweepTraits<FUSE_SMOOTH>::RING>(a, tmU, tmE, tmR, tmDinv, tmD, tmRx, tmDinvx, tmDx); break;
                        case FUSE_RESID: k_matvec_sweep<FUSE_RESID, SweepTraits<FUSE_RESID>::RING>(a, tmU, tmE, tmR, tmDinv, tmD, tmRx, tmDinvx, tmDx); break;
                        case FUSE_SMOOTH_DOT: k_matvec_sweep<FUSE_SMOOTH_DOT, SweepTraits<FUSE_SMOOTH_DOT>::RING>(a, tmU, tmE, tmR, tmDinv, tmD, tmRx, tmDinvx, tmDx); break;
                    }
                });
            for (auto& th : threads) th.join();
            done.store(true);
            dma.join();
        }
    return 0;
}
}
