// comm.cu -- NCCL plumbing for the x-slab partition (one process per GPU).
//
// The reference has no communication layer at all (single Julia process).  The slab partition adds
// exactly three exchanges:
//   * comm_halo_nodes : one node plane (3*plane doubles, contiguous in our layout) to each x-neighbour
//                       before every operator apply
//   * comm_halo_cells : halo_e cell layers of rho / dc for the filter stencils
//   * comm_allreduce  : the PCG dot products, OC volume, energy, change / max|u|
// libnccl.so.2 is dlopen'ed lazily (the copy torch already loaded is reused when the host process
// imported torch first), so single-GPU users never need NCCL installed.
#include <dlfcn.h>
#include <unistd.h>
#include <nccl.h>

#include "ctx.cuh"

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;

static int load_nccl(std::string& err) {
    if (g_nccl.lib) return ESIMP_OK;
    const char* env = getenv("ESIMP_NCCL_LIB");
    const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
    void* lib = nullptr;
    for (const char* n : names) {
        if (!n) continue;
        lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (lib) break;
    }
    if (!lib) {
        err = std::string("cannot dlopen libnccl.so.2: ") + dlerror();
        return ESIMP_ERR_NCCL;
    }
#define SYM(field, name)                                               \
    g_nccl.field = (decltype(g_nccl.field))dlsym(lib, name);           \
    if (!g_nccl.field) {                                               \
        err = std::string("libnccl is missing symbol ") + name;        \
        return ESIMP_ERR_NCCL;                                         \
    }
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(AllReduce, "ncclAllReduce")
    SYM(AllGather, "ncclAllGather")
    SYM(Send, "ncclSend")
    SYM(Recv, "ncclRecv")
    SYM(GroupStart, "ncclGroupStart")
    SYM(GroupEnd, "ncclGroupEnd")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    g_nccl.lib = lib;
    return ESIMP_OK;
}

// inside ncclGroupStart / ncclGroupEnd: remember the first failure and carry on, so that the group is always closed
#define ESIMP_NCCL_IN_GROUP(call)                                   \
    do {                                                            \
        ncclResult_t _r = (call);                                   \
        if (_r != ncclSuccess && _grp_err == ncclSuccess) {         \
            _grp_err = _r;                                          \
            _grp_what = #call;                                      \
        }                                                           \
    } while (0)
#define ESIMP_NCCL_GROUP_BEGIN()           \
    ncclResult_t _grp_err = ncclSuccess;   \
    const char* _grp_what = "";            \
    ESIMP_NCCL(g_nccl.GroupStart())
#define ESIMP_NCCL_GROUP_END()                                                                                        \
    do {                                                                                                              \
        ncclResult_t _e = g_nccl.GroupEnd();                                                                          \
        if (_grp_err != ncclSuccess) return esimp_fail(ctx, ESIMP_ERR_NCCL, std::string(_grp_what) + ": " + g_nccl.GetErrorString(_grp_err)); \
        if (_e != ncclSuccess) return esimp_fail(ctx, ESIMP_ERR_NCCL, std::string("ncclGroupEnd: ") + g_nccl.GetErrorString(_e));            \
    } while (0)
#define ESIMP_NCCL(call)                                                                                         \
    do {                                                                                                         \
        ncclResult_t _r = (call);                                                                                \
        if (_r != ncclSuccess) return esimp_fail(ctx, ESIMP_ERR_NCCL, std::string(#call) + ": " + g_nccl.GetErrorString(_r)); \
    } while (0)

int comm_unique_id(void* id128, std::string& err) {
    int s = load_nccl(err);
    if (s < 0) return s;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclResult_t r = g_nccl.GetUniqueId((ncclUniqueId*)id128);
    if (r != ncclSuccess) {
        err = std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(r);
        return ESIMP_ERR_NCCL;
    }
    return ESIMP_OK;
}

int comm_init(esimp_ctx* ctx, int rank, int nranks, const void* id128) {
    if (nranks < 1 || rank < 0 || rank >= nranks) return esimp_fail(ctx, ESIMP_ERR_ARG, "bad rank / nranks");
    ctx->rank = rank;
    ctx->nranks = nranks;
    if (nranks == 1) return ESIMP_OK;
    std::string err;
    if (load_nccl(err) < 0) return esimp_fail(ctx, ESIMP_ERR_NCCL, err);
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t comm;
    ESIMP_NCCL(g_nccl.CommInitRank(&comm, nranks, id, rank));
    ctx->nccl_comm = comm;
    return ESIMP_OK;
}

void comm_destroy(esimp_ctx* ctx) {
    for (int s = 0; s < 2; ++s) {
        if (ctx->p2p_peer_mail[s] && !ctx->p2p_same_process) cudaIpcCloseMemHandle(ctx->p2p_peer_mail[s]);
        if (ctx->p2p_peer_flags[s] && !ctx->p2p_same_process) cudaIpcCloseMemHandle(ctx->p2p_peer_flags[s]);
        ctx->p2p_peer_mail[s] = nullptr;
        ctx->p2p_peer_flags[s] = nullptr;
    }
    cudaFree(ctx->p2p_mail);
    cudaFree(ctx->p2p_flags);
    ctx->p2p_mail = nullptr;
    ctx->p2p_flags = nullptr;
    ctx->p2p = 0;
    if (ctx->nccl_comm && g_nccl.lib) g_nccl.CommDestroy((ncclComm_t)ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
}

int comm_allreduce(esimp_ctx* ctx, double* d_buf, long long n, bool is_max) {
    if (ctx->nranks <= 1) return ESIMP_OK;
    PROF_BEGIN(ctx, n <= 16 ? PROF_COMM_ALLREDUCE : PROF_COMM_LEVELSUM, true)
    ESIMP_NCCL(g_nccl.AllReduce(d_buf, d_buf, (size_t)n, ncclDouble, is_max ? ncclMax : ncclSum, (ncclComm_t)ctx->nccl_comm,
                                ctx->stream));
    PROF_END(ctx)
    return ESIMP_OK;
}


// ---------------------------------------------------------------------------- node-plane halo over peer memory (NVLink / NVSwitch)
// One kernel per exchange instead of an ncclSend / ncclRecv group (25 us per exchange on 2 B200s, ~800 exchanges per SIMP step):
//   1. every CTA copies its share of my first / last owned plane into the neighbours' mailboxes (16-byte stores over NVLink);
//   2. the last CTA to finish (ticket) publishes the exchange number in the neighbours' flag words (system-scope release);
//   3. every CTA waits until the neighbours' exchange number has arrived in MY flag words (system-scope acquire) and copies its
//      share of my mailbox into my ghost planes.
// Mailbox slots alternate with the parity of the exchange number: a neighbour can be at most one exchange ahead (it cannot finish
// exchange e + 1 before I have pushed e + 1, which I do after unpacking e), so two slots suffice.  A rank that waits longer than
// ~2 s gives up, counts the timeout in flags[3] and leaves the ghost plane as it is (the solve then fails its residual test
// instead of hanging the GPU).
constexpr int P2P_THREADS = 256;
static int p2p_ctas() {
    static const int n = getenv("ESIMP_P2P_CTAS") ? std::max(1, std::min(148, atoi(getenv("ESIMP_P2P_CTAS")))) : 96;
    return n;
}
__device__ __forceinline__ unsigned long long p2p_ld_acquire(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void p2p_st_release(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__global__ void __launch_bounds__(P2P_THREADS) k_halo_p2p(double* v, long long np, long long send_l, long long send_r, long long ghost_l, long long ghost_r,
                                                          double* peer_mail_l, double* peer_mail_r, unsigned long long* peer_flag_l,
                                                          unsigned long long* peer_flag_r, const double* my_mail_from_l, const double* my_mail_from_r,
                                                          unsigned long long* my_flags, unsigned long long epoch_l, unsigned long long epoch_r) {
    const long long n2 = np / 2;  // planes are multiples of 8 doubles (row pitch), 16-byte aligned
    const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
    if (peer_mail_l) {
        const double2* src = reinterpret_cast<const double2*>(v + send_l);
        double2* dst = reinterpret_cast<double2*>(peer_mail_l);
        for (long long i = i0; i < n2; i += stride) dst[i] = src[i];
    }
    if (peer_mail_r) {
        const double2* src = reinterpret_cast<const double2*>(v + send_r);
        double2* dst = reinterpret_cast<double2*>(peer_mail_r);
        for (long long i = i0; i < n2; i += stride) dst[i] = src[i];
    }
    __threadfence_system();
    __syncthreads();
    __shared__ bool s_last;
    if (threadIdx.x == 0) {
        const unsigned long long t = atomicAdd(&my_flags[2], 1ull);
        s_last = t == gridDim.x - 1;
    }
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        if (peer_flag_l) p2p_st_release(peer_flag_l, epoch_l);
        if (peer_flag_r) p2p_st_release(peer_flag_r, epoch_r);
        my_flags[2] = 0ull;  // ticket for the next exchange (stream order: no other exchange kernel of mine is in flight)
    }
    // wait for the neighbours' planes
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        bool ok = true;
        while (ok && ((peer_mail_l && p2p_ld_acquire(&my_flags[0]) < epoch_l) || (peer_mail_r && p2p_ld_acquire(&my_flags[1]) < epoch_r))) {
            __nanosleep(20);
            if (clock64() - t0 > 4000000000ll) ok = false;
        }
        if (!ok) atomicAdd(&my_flags[3], 1ull);
    }
    __syncthreads();
    if (peer_mail_l) {
        const double2* src = reinterpret_cast<const double2*>(my_mail_from_l);
        double2* dst = reinterpret_cast<double2*>(v + ghost_l);
        for (long long i = i0; i < n2; i += stride) dst[i] = __ldcv(&src[i]);
    }
    if (peer_mail_r) {
        const double2* src = reinterpret_cast<const double2*>(my_mail_from_r);
        double2* dst = reinterpret_cast<double2*>(v + ghost_r);
        for (long long i = i0; i < n2; i += stride) dst[i] = __ldcv(&src[i]);
    }
}

// Allocate the mailbox + flags, exchange their IPC handles through the NCCL communicator and map the neighbours'.  Collective.
int comm_p2p_setup(esimp_ctx* ctx, long long plane_doubles) {
    if (ctx->nranks <= 1 || ctx->p2p != 0) return ESIMP_OK;
    // default: on from 4 ranks.  On 2 B200s an exchange costs 24.0 us with 96 CTAs (26.9 us with 32) against 25.9 us for the NCCL group
    // (what both pay is the launch and the wait for the neighbour, not the 1.6 MB transfer); on 8 B200s 23.9 us against 31.8 us, the
    // SIMP step 108.8 ms against 113.3 ms with bit-identical compliance histories (profiles/r2_bench_c5_n8_*.json).
    // ESIMP_P2P_HALO = 0 / 1 forces NCCL / peer memory
    const int enabled = getenv("ESIMP_P2P_HALO") ? atoi(getenv("ESIMP_P2P_HALO")) : (ctx->nranks >= 4 ? 1 : 0);
    ctx->p2p = -1;
    if (!enabled) return ESIMP_OK;
    ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
    const int R = ctx->nranks, r = ctx->rank;
    ctx->p2p_plane_cap = (plane_doubles + 1) & ~1ll;
    struct Rec {
        cudaIpcMemHandle_t mail, flags;
        int ok, device;
        long long pid;                  // ranks of ONE process (esimp_multi) share an address space: plain peer access, no IPC
        double* mail_ptr;
        unsigned long long* flags_ptr;
    };
    Rec mine{};
    mine.ok = 1;
    mine.device = ctx->device;
    mine.pid = (long long)getpid();
    if (cudaMalloc(&ctx->p2p_mail, sizeof(double) * 4 * ctx->p2p_plane_cap) != cudaSuccess ||
        cudaMalloc(&ctx->p2p_flags, sizeof(unsigned long long) * 8) != cudaSuccess)
        mine.ok = 0;
    if (mine.ok) {
        cudaMemset(ctx->p2p_flags, 0, sizeof(unsigned long long) * 8);
        mine.mail_ptr = ctx->p2p_mail;
        mine.flags_ptr = ctx->p2p_flags;
        if (cudaIpcGetMemHandle(&mine.mail, ctx->p2p_mail) != cudaSuccess || cudaIpcGetMemHandle(&mine.flags, ctx->p2p_flags) != cudaSuccess) mine.ok = 0;
    }
    cudaGetLastError();
    // all records to all ranks (bytes through ncclAllGather on a scratch device buffer)
    unsigned char* d_all = nullptr;
    ESIMP_CUDA(cudaMalloc(&d_all, sizeof(Rec) * R));
    ESIMP_CUDA(cudaMemcpyAsync(d_all + sizeof(Rec) * r, &mine, sizeof(Rec), cudaMemcpyHostToDevice, ctx->stream));
    ESIMP_NCCL(g_nccl.AllGather(d_all + sizeof(Rec) * r, d_all, sizeof(Rec), ncclChar, comm, ctx->stream));
    std::vector<Rec> all(R);
    ESIMP_CUDA(cudaMemcpyAsync(all.data(), d_all, sizeof(Rec) * R, cudaMemcpyDeviceToHost, ctx->stream));
    ESIMP_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(d_all);
    bool ok = true;
    for (int q = 0; q < R; ++q) ok = ok && all[q].ok;
    // map the neighbours (every rank must reach the same verdict, so failures are agreed on with one more tiny allreduce)
    int local_fail = 0;
    if (ok) {
        for (int s = 0; s < 2 && !local_fail; ++s) {
            const int q = s == 0 ? r - 1 : r + 1;
            if (q < 0 || q >= R) continue;
            int can = 0;
            if (all[q].device != ctx->device) cudaDeviceCanAccessPeer(&can, ctx->device, all[q].device);
            if (can && all[q].pid == mine.pid) {
                const cudaError_t e = cudaDeviceEnablePeerAccess(all[q].device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) local_fail = 1;
                cudaGetLastError();
                ctx->p2p_peer_mail[s] = all[q].mail_ptr;
                ctx->p2p_peer_flags[s] = all[q].flags_ptr;
                ctx->p2p_same_process = true;
                continue;
            }
            if (!can || cudaIpcOpenMemHandle((void**)&ctx->p2p_peer_mail[s], all[q].mail, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
                cudaIpcOpenMemHandle((void**)&ctx->p2p_peer_flags[s], all[q].flags, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess)
                local_fail = 1;
        }
        cudaGetLastError();
    } else {
        local_fail = 1;
    }
    double* d_f = ctx->d_scal + SC_TMP2;
    const double lf = (double)local_fail;
    ESIMP_CUDA(cudaMemcpyAsync(d_f, &lf, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    ESIMP_NCCL(g_nccl.AllReduce(d_f, d_f, 1, ncclDouble, ncclMax, comm, ctx->stream));
    double any_fail = 1.0;
    ESIMP_CUDA(cudaMemcpyAsync(&any_fail, d_f, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    ESIMP_CUDA(cudaStreamSynchronize(ctx->stream));
    if (any_fail == 0.0) ctx->p2p = 1;
    else if (getenv("ESIMP_P2P_DEBUG")) fprintf(stderr, "[esimp rank %d] peer-memory halo unavailable (local_fail %d): NCCL send / recv stays\n", r, local_fail);
    return ESIMP_OK;
}

static int halo_nodes_p2p(esimp_ctx* ctx, const GridDev& g, int own_p0, int own_p1, bool has_left, bool has_right, double* v) {
    const long long np = 3 * g.plane, cap = ctx->p2p_plane_cap;
    if (has_left) ctx->p2p_epoch[0]++;
    if (has_right) ctx->p2p_epoch[1]++;
    const unsigned long long el = ctx->p2p_epoch[0], er = ctx->p2p_epoch[1];
    // slot layout of a mailbox: [from left | from right][parity][cap].  I am the RIGHT neighbour of my left neighbour, so my plane goes
    // into its "from right" half, and into the "from left" half of my right neighbour
    double* peer_l = has_left ? ctx->p2p_peer_mail[0] + (2 + (el & 1)) * cap : nullptr;
    double* peer_r = has_right ? ctx->p2p_peer_mail[1] + (0 + (er & 1)) * cap : nullptr;
    unsigned long long* pf_l = has_left ? ctx->p2p_peer_flags[0] + 1 : nullptr;   // the left neighbour's "written by the right one"
    unsigned long long* pf_r = has_right ? ctx->p2p_peer_flags[1] + 0 : nullptr;  // the right neighbour's "written by the left one"
    PROF_BEGIN(ctx, PROF_COMM_HALO, true)
    k_halo_p2p<<<p2p_ctas(), P2P_THREADS, 0, ctx->stream>>>(v, np, (long long)own_p0 * np, (long long)(own_p1 - 1) * np, 0ll, (long long)(g.NPX - 1) * np, peer_l,
                                                          peer_r, pf_l, pf_r, ctx->p2p_mail + (0 + (el & 1)) * cap, ctx->p2p_mail + (2 + (er & 1)) * cap,
                                                          ctx->p2p_flags, el, er);
    PROF_END(ctx)
    ctx->launches++;
    ESIMP_CUDA(cudaGetLastError());
    return ESIMP_OK;
}

// Send my first / last OWNED node plane to the left / right neighbour, receive their planes into my
// ghost planes.  A plane is 3*plane contiguous doubles ([lp][c][y][z] layout): no packing.
int comm_halo_nodes_lv(esimp_ctx* ctx, const GridDev& g, int own_p0, int own_p1, bool has_left, bool has_right, double* v) {
    if (ctx->nranks <= 1) return ESIMP_OK;
    if (ctx->p2p == 1 && 3 * g.plane <= ctx->p2p_plane_cap) return halo_nodes_p2p(ctx, g, own_p0, own_p1, has_left, has_right, v);
    ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
    const size_t np = (size_t)3 * g.plane;
    PROF_BEGIN(ctx, PROF_COMM_HALO, true)
    ESIMP_NCCL_GROUP_BEGIN();
    if (has_left) {
        ESIMP_NCCL_IN_GROUP(g_nccl.Send(v + (size_t)own_p0 * np, np, ncclDouble, ctx->rank - 1, comm, ctx->stream));
        ESIMP_NCCL_IN_GROUP(g_nccl.Recv(v, np, ncclDouble, ctx->rank - 1, comm, ctx->stream));
    }
    if (has_right) {
        ESIMP_NCCL_IN_GROUP(g_nccl.Send(v + (size_t)(own_p1 - 1) * np, np, ncclDouble, ctx->rank + 1, comm, ctx->stream));
        ESIMP_NCCL_IN_GROUP(g_nccl.Recv(v + (size_t)(g.NPX - 1) * np, np, ncclDouble, ctx->rank + 1, comm, ctx->stream));
    }
    ESIMP_NCCL_GROUP_END();
    PROF_END(ctx)
    return ESIMP_OK;
}
int comm_allreduce_f32(esimp_ctx* ctx, float* d_buf, long long n) {
    if (ctx->nranks <= 1) return ESIMP_OK;
    ESIMP_NCCL(g_nccl.AllReduce(d_buf, d_buf, (size_t)n, ncclFloat, ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    return ESIMP_OK;
}
int comm_halo_nodes_lv_f32(esimp_ctx* ctx, const GridDev& g, int own_p0, int own_p1, bool has_left, bool has_right, float* v) {
    if (ctx->nranks <= 1) return ESIMP_OK;
    ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
    const size_t np = (size_t)3 * g.plane;
    ESIMP_NCCL_GROUP_BEGIN();
    if (has_left) {
        ESIMP_NCCL_IN_GROUP(g_nccl.Send(v + (size_t)own_p0 * np, np, ncclFloat, ctx->rank - 1, comm, ctx->stream));
        ESIMP_NCCL_IN_GROUP(g_nccl.Recv(v, np, ncclFloat, ctx->rank - 1, comm, ctx->stream));
    }
    if (has_right) {
        ESIMP_NCCL_IN_GROUP(g_nccl.Send(v + (size_t)(own_p1 - 1) * np, np, ncclFloat, ctx->rank + 1, comm, ctx->stream));
        ESIMP_NCCL_IN_GROUP(g_nccl.Recv(v + (size_t)(g.NPX - 1) * np, np, ncclFloat, ctx->rank + 1, comm, ctx->stream));
    }
    ESIMP_NCCL_GROUP_END();
    return ESIMP_OK;
}
int comm_halo_nodes(esimp_ctx* ctx, double* v) {
    return comm_halo_nodes_lv(ctx, ctx->g, ctx->own_p0, ctx->own_p1, ctx->has_left, ctx->has_right, v);
}

// `depth` cell layers each way (rho, dc for the filter; E for the redundant interface layer)
int comm_halo_cells_lv(esimp_ctx* ctx, const GridDev& g, int own_e0, int own_e1, int depth, bool has_left, bool has_right, double* a) {
    if (ctx->nranks <= 1) return ESIMP_OK;
    ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
    const size_t nl = (size_t)depth * g.eplane;
    ESIMP_NCCL_GROUP_BEGIN();
    if (has_left) {
        ESIMP_NCCL_IN_GROUP(g_nccl.Send(a + (size_t)own_e0 * g.eplane, nl, ncclDouble, ctx->rank - 1, comm, ctx->stream));
        ESIMP_NCCL_IN_GROUP(g_nccl.Recv(a + (size_t)(own_e0 - depth) * g.eplane, nl, ncclDouble, ctx->rank - 1, comm, ctx->stream));
    }
    if (has_right) {
        ESIMP_NCCL_IN_GROUP(g_nccl.Send(a + (size_t)(own_e1 - depth) * g.eplane, nl, ncclDouble, ctx->rank + 1, comm, ctx->stream));
        ESIMP_NCCL_IN_GROUP(g_nccl.Recv(a + (size_t)own_e1 * g.eplane, nl, ncclDouble, ctx->rank + 1, comm, ctx->stream));
    }
    ESIMP_NCCL_GROUP_END();
    return ESIMP_OK;
}
int comm_halo_cells(esimp_ctx* ctx, double* a) {
    return comm_halo_cells_lv(ctx, ctx->g, ctx->own_e0, ctx->own_e1, ctx->halo_e, ctx->has_left, ctx->has_right, a);
}
