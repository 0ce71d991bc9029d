// ctx.cuh -- context, slab geometry and shared device helpers of libeasysimp_b200.so
//
// Data layout in HBM (structured grids), chosen for the x-slab partition and for coalesced
// lane-along-z access; it is NOT the reference's layout (Ferrite numbers dofs by first
// encounter in cell order, FiniteElementAnalysis.jl:257-279) -- api.cu permutes once at the
// boundary:
//   node vectors   v[lp][c][j][k]   lp = local x node plane, c = component, j = y, k = z (fastest)
//                  -> one x-plane (all 3 components) is ONE contiguous block of 3*plane doubles,
//                     so the halo exchanged with a neighbour slab needs no packing and the dofs a
//                     rank owns are one contiguous range (dot products run over it directly).
//   cell arrays    a[le][j][k]      le = local x element layer
//   node mask      m[lp][j][k]      uint8, bit c set = dof (node, c) is prescribed
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/easysimp_b200.h"
#include "grid_dev.h"

#define ESIMP_CUDA(call)                                                                         \
    do {                                                                                         \
        cudaError_t _e = (call);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            return esimp_fail(ctx, ESIMP_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e)); \
        }                                                                                        \
    } while (0)
#define ESIMP_TRY(expr)              \
    do {                             \
        int _s = (expr);             \
        if (_s < 0) return _s;       \
    } while (0)

// what the structured operator needs to know about one grid level (fine grid or a multigrid level)
struct OpView {
    GridDev g;
    int own_p0, own_p1, mv_e0, mv_e1;
    const double* E;
    const uint8_t* mask;
};

// one level of the geometric multigrid hierarchy (level 0 aliases the context's fine-grid arrays)
struct MGLevel {
    OpView vw{};
    bool has_left = false, has_right = false;
    long long ndof_loc = 0, ncell_loc = 0, own_dof0 = 0, own_dof1 = 0, own_cell0 = 0, own_cell1 = 0, n_dofs_global = 0;
    double* d_E = nullptr;
    uint8_t* d_mask = nullptr;
    long long* d_fixed = nullptr;
    long long n_fixed = 0;
    double* d_dinv = nullptr;
    double* d_sinv = nullptr;  // node scalar form of d_dinv for the sweep kernel's smoother modes (hex8_node_scalar_dinv)
    double diag_value = 1.0, lmax = 4.0;
    double *d_r = nullptr, *d_z = nullptr, *d_d = nullptr, *d_q = nullptr;  // rhs, correction, Chebyshev direction, A z
    double* d_z2 = nullptr;  // second buffer of the correction: the fused smoother writes z' while neighbours still read z
    double* d_pi = nullptr;  // power-iteration vector, kept between SIMP iterations (warm start of the lmax estimate)
    bool pi_valid = false;
    bool owns = false;
    bool replicated = false;  // multi-rank: every rank holds the whole (small) grid of this level, no communication
    // single-precision mirror of the level for the FP32 V-cycle (multigrid32.cu; allocated on first use)
    float *f_E = nullptr, *f_dinv = nullptr, *f_r = nullptr, *f_z = nullptr, *f_z2 = nullptr, *f_d = nullptr, *f_q = nullptr;
};

struct UnstructuredMesh {
    int npe = 0;            // 8 or 4
    int64_t n_cells = 0, n_nodes = 0, n_dofs = 0;
    int n_colors = 0;
    std::vector<int64_t> color_ptr;  // cells of colour c are order[color_ptr[c] .. color_ptr[c+1])
    int32_t* d_order = nullptr;      // cell ids sorted by colour
    int32_t* d_celldofs = nullptr;   // [3*npe][n_cells] (dof-major: coalesced over cells)
    double* d_ke = nullptr;          // [(3npe)^2][n_cells] unit-modulus element matrices
    double* d_vol = nullptr;         // element volumes
    // CSR filter
    int64_t nnz = 0;
    int64_t* d_rowptr = nullptr;
    int32_t* d_col = nullptr;
    double* d_w = nullptr;           // max(0, R - dist)
    double* d_D = nullptr;           // D_i = sum_j w_ij V_j
    // body force / stress recovery (post_load.cu), colour-sorted and entry-major like d_ke
    double* d_nint = nullptr;        // [npe][n_cells] int N_a dOmega (2x2x2 Gauss on hexes, V/4 on tets)
    double* d_xcell = nullptr;       // [3*npe][n_cells] node coordinates of each cell
};

// __constant__ blocks have one copy per DEVICE; who uploaded one last is tracked per device under this lock, so that contexts on
// different GPUs of one process (one host thread each) do not invalidate or race each other's uploads.  Two contexts on the SAME
// device with different cuboids still share the block: drive them one after the other (every entry point drains its stream).
#include <mutex>
inline std::mutex g_const_mutex;

struct NcclApi;  // comm.cu

struct esimp_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    esimp_params prm{};
    int64_t launches = 0;

    // ---- distributed
    int rank = 0, nranks = 1;
    void* nccl_comm = nullptr;
    // peer-memory halo exchange of node planes (comm.cu): every rank owns one mailbox (2 sides x 2 parities x one fine node plane)
    // and one flag block, both mapped into the x-neighbours through CUDA IPC; a single kernel pushes the boundary planes over NVLink,
    // signals, waits for the neighbours' signal and unpacks.  p2p = 0: not set up (NCCL send / recv), 1: active, -1: unavailable
    int p2p = 0;
    bool p2p_same_process = false;  // neighbours are contexts of this process (esimp_multi): raw peer pointers, nothing to close
    double* p2p_mail = nullptr;               // [side from: 0 left, 1 right][parity][p2p_plane_cap]
    unsigned long long* p2p_flags = nullptr;  // [0] epoch written by the left neighbour, [1] by the right one, [2] CTA ticket, [3] timeouts
    double* p2p_peer_mail[2] = {nullptr, nullptr};               // the left / right neighbour's mailbox (IPC mappings)
    unsigned long long* p2p_peer_flags[2] = {nullptr, nullptr};  // ... and flag block
    long long p2p_plane_cap = 0;                                  // doubles per mailbox slot
    unsigned long long p2p_epoch[2] = {0, 0};                     // exchanges done with the left / right neighbour

    // ---- mesh kind
    bool structured = false, unstructured = false;
    GridDev g{};
    double hx = 1, hy = 1, hz = 1;
    int own_p0 = 0, own_p1 = 0;  // owned local node planes
    int own_e0 = 0, own_e1 = 0;  // owned local element layers
    int mv_e0 = 0, mv_e1 = 0;    // element layers the operator visits (owned + redundant left layer)
    int halo_e = 0;              // ghost element layers per interior side
    int fused_chunk = 0;         // >0: force the x-chunk length of the fused operator (tuning)
    bool has_left = false, has_right = false;
    int64_t n_dofs_global = 0, n_cells_global = 0;
    int64_t ndof_loc = 0, ncell_loc = 0;  // stored (incl. ghosts)
    int64_t own_dof0 = 0, own_dof1 = 0;   // owned dof range in local storage
    int64_t own_cell0 = 0, own_cell1 = 0; // owned cell range
    int64_t* d_dof_of_node = nullptr;     // caller numbering (global, 3*n_nodes) or null
    UnstructuredMesh um;
    std::vector<MGLevel> mg;            // geometric multigrid hierarchy (precond == GMG)
    std::vector<int64_t> prescribed_nc;  // prescribed (3*node_ref + comp), kept for the coarse-level masks
    int bc_version = 0, mg_bc_version = -1;
    int mg_replicate = 1;                   // multi-rank: replicate the deep levels instead of stopping / exchanging halos
    long long mg_replicate_cells = 100000;  // levels with at most this many cells (global) are replicated (64x32x32 at C5)
    int mg_min_layers = 1;  // multi-rank: stop coarsening when a rank would hold fewer cell layers than this
    int mg_nu = 2, mg_coarse_its = 16;  // tuned on 512x256x256 (scripts/mg_sweep.sh): nu 2, ratio 10 is the fastest
    double mg_ratio = 10.0;
    // GMG-CG convergence poll: 0 / -1 = read <r,r> back after every iteration (the host waits for the device once per iteration),
    // 1 = lagged: iteration k is enqueued before <r,r> of iteration k-1 is looked at, so the host runs one iteration ahead.  The
    // iterate returned is the same (the deferred x update of the surplus iteration is never applied).  Opt-in (ESIMP_PCG_LAG=1):
    // on 1 and 2 B200s the fine-grid kernels are long enough for the host to keep ahead and the lagged poll measured 2 % slower
    int mg_lag = -1;
    // levels >= mg_tail_start run inside ONE cluster kernel (coarse_tail.cuh); 0 = none.  ESIMP_MG_TAIL=0 switches the tail off
    int mg_tail = 1, mg_tail_start = 0;
    // largest level (nodes) the tail takes (ESIMP_MG_TAIL_NODES).  10 000 would also take the 32x16x16 level of a 512x256x256 hierarchy, but the
    // dense node-by-node operator of the tail on 8 SMs is slower there than nine launches of the fused kernel on 148 (ncu: 304 us per
    // V-cycle for the tail with that level, SIMP step 331.4 ms against 329.3 ms without it)
    long long mg_tail_nodes = 2000;
    double* d_K0 = nullptr;  // dense 24 x 24 unit element matrix for the tail kernel
    cudaEvent_t lag_ev[2]{};

    // element matrices (host copies; uploaded to __constant__ on demand)
    double K0[24 * 24];
    double B[8][3][3];
    double kdiag[3];
    bool k0_valid = false;
    double diag_value = 1.0;  // m = mean |diag K|

    // stencil filter
    std::vector<int> st_off;   // 3*n
    std::vector<double> st_w;
    int st_n = 0, st_halo = 0;
    double* d_filtD = nullptr;  // D_i per stored cell (structured)
    bool filter_ready = false;

    // ---- device state
    uint8_t* d_mask = nullptr;  // per node (structured) / per dof (unstructured)
    int64_t* d_fixed = nullptr; // prescribed dofs in device numbering
    int64_t n_fixed = 0;
    double *d_f = nullptr, *d_u = nullptr, *d_r = nullptr, *d_p = nullptr, *d_q = nullptr, *d_dinv = nullptr;
    double *d_rho = nullptr, *d_phys = nullptr, *d_E = nullptr, *d_dc = nullptr, *d_dcf = nullptr, *d_dv = nullptr,
           *d_rho_new = nullptr, *d_rho_old = nullptr, *d_ctmp = nullptr;
    double* d_stage = nullptr;  // staging for global host arrays
    int64_t stage_cap = 0;
    // reductions
    double* d_partial = nullptr;    // [16][RED_BLOCKS]
    double* d_partial_mv = nullptr; // one partial per CTA of the fused operator (fused <p,q>)
    long long partial_mv_cap = 0;
    unsigned int* d_ticket = nullptr;
    double* d_scal = nullptr;       // device scalars (see SC_* in vector_ops.cu)
    double* h_scal = nullptr;       // pinned mirror
    bool density_set = false, load_set = false, design_set = false;
    // variable-density body force (Optimization.jl:300-311): f = fload + body(rho_phys * body_base) on every assembly
    bool body_on = false;
    double body_b[3] = {0.0, 0.0, 0.0}, body_base = 0.0;
    double* d_fload = nullptr;  // density-independent part of f (allocated when the body force is switched on)
    double f_norm2 = 0.0;
    cudaEvent_t ev[6]{};
    // in-region timing of the operator kernel (esimp_profile_*): event pairs recorded on ctx->stream
    int prof_on = 0;  // 0 off, 1 = event pairs around the plain / dot operator launches only (cheap enough for a timed region), 2 = every category
    std::vector<cudaEvent_t> prof_ev;
    std::vector<uint8_t> prof_tag;
    int prof_n = 0;
    double prof_ms_cat[24] = {0};    // summed kernel time per category (PROF_* below)
    int64_t prof_cnt_cat[24] = {0};
    cudaEvent_t prof_coarse_ev[2]{};  // one V-cycle's levels >= 1 (PROF_MG_COARSE): own pair, it encloses other categories
    bool prof_coarse_pending = false;
};
int prof_collect(esimp_ctx* ctx);  // vector_ops.cu: fold finished event pairs into prof_ms_cat / prof_cnt_cat
// kernel categories of the in-region timing (fine grid only; multigrid levels reuse the kernels on coarser grids)
enum { PROF_OP_PLAIN = 0, PROF_OP_SMOOTH = 1, PROF_OP_RESID = 2, PROF_FILTER = 3, PROF_SENS = 4, PROF_OC_VOLUMES = 5, PROF_PROLONG = 6,
       PROF_RESTRICT = 7, PROF_MG_UPDATE_FIRST = 8, PROF_MG_DIRECTION = 9, PROF_UM_MATVEC = 10,
       // slab partition: neighbour plane exchange, scalar allreduces, the level-sum allreduce at the distributed -> replicated
       // transition, everything below level 0 of one V-cycle (restriction to prolongation), and two HOST clocks of the PCG loop
       // (ms blocked in the per-iteration scalar fetch; ms of the whole solve)
       PROF_COMM_HALO = 11, PROF_COMM_ALLREDUCE = 12, PROF_COMM_LEVELSUM = 13, PROF_MG_COARSE = 14, PROF_HOST_PCG = 15, PROF_HOST_WAIT = 16, PROF_NCAT = 17 };
// bracket the launches between PROF_BEGIN and PROF_END with an event pair on the context's stream while profiling is on
#define PROF_BEGIN(ctx, cat, cond)                                                                         \
    const bool _prof_timed = (ctx)->prof_on >= ((cat) == PROF_OP_PLAIN ? 1 : 2) && (cond) && 2 * (ctx)->prof_n + 1 < (int)(ctx)->prof_ev.size(); \
    if (_prof_timed) {                                                                                     \
        (ctx)->prof_tag[(ctx)->prof_n] = (cat);                                                            \
        cudaEventRecord((ctx)->prof_ev[2 * (ctx)->prof_n], (ctx)->stream);                                 \
    }
#define PROF_END(ctx) \
    if (_prof_timed) cudaEventRecord((ctx)->prof_ev[2 * (ctx)->prof_n++ + 1], (ctx)->stream);

int esimp_fail(esimp_ctx* ctx, int code, const std::string& msg);

// ------------------------------------------------------------------ launch geometry helpers
constexpr int RED_BLOCKS = 148 * 8;  // persistent grid-stride grids: 8 CTAs of 256 per SM (full occupancy)
constexpr int RED_THREADS = 256;
constexpr int N_SCAL = 32;

// device scalar slots
enum {
    // PCG: <p,q>; two ping-pong pairs (<r,z>, <r,r>) so that one 2-double allreduce serves both; <f,f>
    SC_PQ = 0, SC_RZ0 = 1, SC_RR0 = 2, SC_RZ1 = 3, SC_RR1 = 4, SC_FF = 5, SC_VOL = 6, SC_ENERGY = 7, SC_CHANGE = 8,
    SC_GRAY = 9, SC_MAXU = 10, SC_DIAGSUM = 11, SC_TMP0 = 12, SC_TMP1 = 13, SC_TMP2 = 14,
    SC_OC0 = 16  // .. SC_OC0 + 14: volumes of the 15 candidate multipliers of one OC multi-section pass
};

// ------------------------------------------------------------------ module entry points
// hex8_structured.cu
int hex8_upload_constants(esimp_ctx* ctx);
void hex8_forget_ctx(const esimp_ctx* ctx);
void filter_forget_ctx(const esimp_ctx* ctx);
void post_forget_ctx(const esimp_ctx* ctx);
void um_forget_ctx(const esimp_ctx* ctx);
// post_load.cu
int body_force_into(esimp_ctx* ctx, const double* d_dens, double base, const double b[3], const double* d_fload, double* d_out);
int body_force_refresh_rhs(esimp_ctx* ctx);  // d_f = d_fload + body(d_phys) when the body force is on
int stress_fields(esimp_ctx* ctx, double* max_vm, int64_t* max_cell);  // d_ctmp = vm(q1), d_dcf = vm(avg), d_rho_new = energy
int stress_tensors(esimp_ctx* ctx, int64_t id0, int64_t n, double* d_out);
int hex8_matvec(esimp_ctx* ctx, const double* d_in, double* d_out, int variant, int dot_slot = -1);
bool hex8_matvec_fuses_dot(const esimp_ctx* ctx);
OpView fine_view(const esimp_ctx* ctx);
int hex8_matvec_view(esimp_ctx* ctx, const OpView& vw, const double* d_in, double* d_out, int variant, int dot_slot);
int hex8_diagonal_view(esimp_ctx* ctx, const OpView& vw, double* dinv, long long n_dofs_global, double* diag_value);
// dmode (FUSE_D_* of fused_kernel.cuh): 1 = the previous direction is read from d, 2 = it is z_in itself (first step after a zero guess),
// 3 = it is z_in - zprev (zprev may alias z_out); write_d = 0 drops the store of the new direction
int hex8_smooth_view(esimp_ctx* ctx, const OpView& vw, const double* z_in, double* z_out, const double* r, const double* dinv,
                     double* d, double c1, double c2, int dot_slot = -1, int dmode = 1, int write_d = 1, const double* zprev = nullptr,
                     const double* sinv = nullptr);
// D^-1 of a view as ONE double per node ([NPX][ny+1][nzp]): 1 / (sum of E around the node) with the prescribed-component mask in the
// three lowest mantissa bits -- what the sweep kernel's smoother modes read instead of the three-component dinv (8 instead of 24 B/node)
int hex8_node_scalar_dinv(esimp_ctx* ctx, const OpView& vw, const double* dinv, double* sinv);
int hex8_resid_view(esimp_ctx* ctx, const OpView& vw, const double* z_in, double* res_out, const double* r);
bool hex8_sweep_eligible(const esimp_ctx* ctx, const OpView& vw);
int hex8_smooth_view_f32(esimp_ctx* ctx, const OpView& vw, const float* E, const float* z_in, float* z_out, const float* r,
                         const float* dinv, float* d, float c1, float c2);
int hex8_resid_view_f32(esimp_ctx* ctx, const OpView& vw, const float* E, const float* z_in, float* res_out, const float* r);
// multigrid32.cu: FP64 CG preconditioned by the V-cycle evaluated in FP32 (precond == ESIMP_PRECOND_GMG_FP32)
int pcg_solve_mg32(esimp_ctx* ctx, int* iters, double* relres);
// multigrid.cu
int mg_setup(esimp_ctx* ctx);          // build / refresh the hierarchy for the current E and mask
void mg_free(esimp_ctx* ctx);
int pcg_solve_mg(esimp_ctx* ctx, int* iters, double* relres);
int fixup_lv(esimp_ctx* ctx, const long long* fixed, long long n_fixed, double m, const double* in, double* out);
int comm_halo_nodes_lv(esimp_ctx* ctx, const GridDev& g, int own_p0, int own_p1, bool has_left, bool has_right, double* v);
int comm_halo_cells_lv(esimp_ctx* ctx, const GridDev& g, int own_e0, int own_e1, int depth, bool has_left, bool has_right, double* a);
int hex8_diagonal(esimp_ctx* ctx);                 // fills d_dinv with 1/diag, sets diag_value
int hex8_sens_energy(esimp_ctx* ctx);              // d_dc (owned), SC_ENERGY
// vector_ops.cu
int vec_dot(esimp_ctx* ctx, const double* a, const double* b, int slot);
int vec_fixup_dirichlet(esimp_ctx* ctx, const double* in, double* out);
int pcg_solve(esimp_ctx* ctx, int* iters, double* relres);
int vec_gather_to_device(esimp_ctx* ctx, const double* d_src_global, double* d_dst_local);
int vec_scatter_from_device(esimp_ctx* ctx, const double* d_src_local, double* d_dst_global);
int cell_gather_to_device(esimp_ctx* ctx, const double* d_src_global, double* d_dst_local);
int cell_scatter_from_device(esimp_ctx* ctx, const double* d_src_local, double* d_dst_global);
int reduce_sum_scaled(esimp_ctx* ctx, const double* a, const double* w_or_null, double scale, int64_t i0, int64_t i1, int slot);
int reduce_metrics(esimp_ctx* ctx);                // change / grayness / max|u|
int fetch_scalars(esimp_ctx* ctx);                 // d_scal -> h_scal (sync)
int allreduce_scalars(esimp_ctx* ctx, int slot0, int n, bool is_max);
// filter_oc.cu
int simp_modulus(esimp_ctx* ctx);                  // d_phys -> d_E (all stored layers)
int filter_build_stencil(esimp_ctx* ctx);
int filter_sens_apply(esimp_ctx* ctx, const double* rho, const double* dc, double* out);
int filter_dens_apply(esimp_ctx* ctx, const double* rho, double* out);
int filter_densT_apply(esimp_ctx* ctx, const double* s, double* out);
int oc_update(esimp_ctx* ctx, const double* rho, const double* dc, const double* dv, double* rho_new, double* lambda,
              int* steps);
// unstructured.cu
int um_build(esimp_ctx* ctx, int celltype, int64_t n_cells, int64_t n_nodes, const double* xyz, const int64_t* conn,
             const int64_t* celldofs);
int um_matvec(esimp_ctx* ctx, const double* d_in, double* d_out);
int um_diagonal(esimp_ctx* ctx);
int um_sens_energy(esimp_ctx* ctx);
int um_filter_build(esimp_ctx* ctx, const int64_t* rowptr, const int64_t* col, const double* centers,
                    const double* volumes, double R);
void um_free(esimp_ctx* ctx);
void hex8_element_matrix(const double* xyz8, double lam, double mu, double* ke24);
void tet4_element_matrix(const double* xyz4, double lam, double mu, double* ke12);
// comm.cu
int comm_unique_id(void* id128, std::string& err);
int comm_init(esimp_ctx* ctx, int rank, int nranks, const void* id128);
int comm_p2p_setup(esimp_ctx* ctx, long long plane_doubles);  // after the slab is known; falls back to NCCL when IPC is unavailable
void comm_destroy(esimp_ctx* ctx);
int comm_allreduce(esimp_ctx* ctx, double* d_buf, long long n, bool is_max);
int comm_allreduce_f32(esimp_ctx* ctx, float* d_buf, long long n);
int comm_halo_nodes_lv_f32(esimp_ctx* ctx, const GridDev& g, int own_p0, int own_p1, bool has_left, bool has_right, float* v);
int comm_halo_nodes(esimp_ctx* ctx, double* d_vec);            // one-plane u halo, both directions
int comm_halo_cells(esimp_ctx* ctx, double* d_arr);            // halo_e-layer cell halo, both directions

// ------------------------------------------------------------------ device reduction helpers
#ifdef __CUDACC__
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Deterministic grid reduction of NV values: warp shuffle -> smem -> per-block partial ->
// the last block to finish (ticket) folds the partials in a fixed order and writes out[i].
// partial: [NV][gridDim.x]; all threads of all blocks must call it.
template <int NV, bool IS_MAX>
__device__ __forceinline__ void grid_reduce_idx(double (&v)[NV], double* partial, unsigned int* ticket, const RedOut<NV>& out,
                                                unsigned int bid, unsigned int nblk);
template <int NV, bool IS_MAX>
__device__ __forceinline__ void grid_reduce(double (&v)[NV], double* partial, unsigned int* ticket, const RedOut<NV>& out) {
    grid_reduce_idx<NV, IS_MAX>(v, partial, ticket, out, blockIdx.x, gridDim.x);
}
// partial: [NV][nblk]; bid = linear CTA id; every thread of every CTA must call it
template <int NV, bool IS_MAX>
__device__ __forceinline__ void grid_reduce_idx(double (&v)[NV], double* partial, unsigned int* ticket, const RedOut<NV>& out,
                                                unsigned int bid, unsigned int nblk) {
    __shared__ double s_red[NV][32];
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double w = IS_MAX ? warp_max(v[i]) : warp_sum(v[i]);
        if (lane == 0) s_red[i][warp] = w;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            double w = lane < nwarp ? s_red[i][lane] : (IS_MAX ? -1.0e300 : 0.0);
            w = IS_MAX ? warp_max(w) : warp_sum(w);
            if (lane == 0) partial[(size_t)i * nblk + bid] = w;
        }
    }
    if (threadIdx.x == 0) {
        __threadfence();
        unsigned int t = atomicAdd(ticket, 1u);
        s_last = (t == nblk - 1);
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            double acc = IS_MAX ? -1.0e300 : 0.0;
            for (unsigned int b = threadIdx.x; b < nblk; b += blockDim.x) {
                double w = __ldcg(&partial[(size_t)i * nblk + b]);
                acc = IS_MAX ? fmax(acc, w) : acc + w;
            }
            acc = IS_MAX ? warp_max(acc) : warp_sum(acc);
            __syncthreads();
            if (lane == 0) s_red[i][warp] = acc;
            __syncthreads();
            if (warp == 0) {
                double w = lane < nwarp ? s_red[i][lane] : (IS_MAX ? -1.0e300 : 0.0);
                w = IS_MAX ? warp_max(w) : warp_sum(w);
                if (lane == 0) *out.p[i] = w;
            }
        }
        if (threadIdx.x == 0) *ticket = 0u;
    }
}
#endif
