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
#include <nccl.h>

#include "ctx.cuh"

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
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

// Send my first / last OWNED node plane to the left / right neighbour, receive their planes into my
// ghost planes.  A plane is 3*plane contiguous doubles ([lp][c][y][z] layout): no packing.
int comm_halo_nodes_lv(esimp_ctx* ctx, const GridDev& g, int own_p0, int own_p1, bool has_left, bool has_right, double* v) {
    if (ctx->nranks <= 1) return ESIMP_OK;
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
