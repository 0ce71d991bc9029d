// api.cu -- the C ABI (include/easysimp_b200.h): context lifecycle, host<->device marshalling, the
// stage calls used by the parity tests and the fused SIMP iteration (Optimization.jl:271-379).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "ctx.cuh"

int op_apply(esimp_ctx* ctx, double* d_in, double* d_out);  // vector_ops.cu
int op_apply_general(esimp_ctx* ctx, double* d_in, double* d_tmp, double* d_out);

static std::string g_create_error;

int esimp_fail(esimp_ctx* ctx, int code, const std::string& msg) {
    if (ctx)
        ctx->err = msg;
    else
        g_create_error = msg;
    return code;
}

#define CHECK_CTX()                                  \
    if (!ctx) return ESIMP_ERR_ARG;                  \
    if (cudaSetDevice(ctx->device) != cudaSuccess) return esimp_fail(ctx, ESIMP_ERR_CUDA, "cudaSetDevice failed")
#define NEED(cond, msg) \
    if (!(cond)) return esimp_fail(ctx, ESIMP_ERR_STATE, msg)

template <typename T>
static int dev_alloc(esimp_ctx* ctx, T** p, int64_t n, bool zero = true) {
    if (*p) cudaFree(*p);
    *p = nullptr;
    ESIMP_CUDA(cudaMalloc(p, sizeof(T) * std::max<int64_t>(n, 1)));
    if (zero) ESIMP_CUDA(cudaMemset(*p, 0, sizeof(T) * std::max<int64_t>(n, 1)));
    return ESIMP_OK;
}
static int ensure_stage(esimp_ctx* ctx, int64_t n) {
    if (ctx->stage_cap >= n) return ESIMP_OK;
    ESIMP_TRY(dev_alloc(ctx, &ctx->d_stage, n, false));
    ctx->stage_cap = n;
    return ESIMP_OK;
}
static void free_mesh_state(esimp_ctx* ctx) {
    void* ptrs[] = {ctx->d_dof_of_node, ctx->d_filtD, ctx->d_mask, ctx->d_fixed, ctx->d_f, ctx->d_u, ctx->d_r, ctx->d_p,
                    ctx->d_q, ctx->d_dinv, ctx->d_rho, ctx->d_phys, ctx->d_E, ctx->d_dc, ctx->d_dcf, ctx->d_dv,
                    ctx->d_rho_new, ctx->d_rho_old, ctx->d_ctmp, ctx->d_stage, ctx->d_fload};
    for (void* p : ptrs) cudaFree(p);
    ctx->d_dof_of_node = nullptr;
    ctx->d_filtD = nullptr;
    ctx->d_mask = nullptr;
    ctx->d_fixed = nullptr;
    ctx->d_f = ctx->d_u = ctx->d_r = ctx->d_p = ctx->d_q = ctx->d_dinv = nullptr;
    ctx->d_rho = ctx->d_phys = ctx->d_E = ctx->d_dc = ctx->d_dcf = ctx->d_dv = ctx->d_rho_new = ctx->d_rho_old = ctx->d_ctmp = nullptr;
    ctx->d_stage = nullptr;
    ctx->d_fload = nullptr;
    ctx->body_on = false;
    post_forget_ctx(ctx);
    um_forget_ctx(ctx);
    ctx->stage_cap = 0;
    ctx->n_fixed = 0;
    um_free(ctx);
    mg_free(ctx);
    ctx->prescribed_nc.clear();
    ctx->structured = ctx->unstructured = false;
    ctx->filter_ready = ctx->density_set = ctx->load_set = ctx->design_set = false;
    ctx->k0_valid = false;
}
static int alloc_state(esimp_ctx* ctx) {
    double** nodev[] = {&ctx->d_f, &ctx->d_u, &ctx->d_r, &ctx->d_p, &ctx->d_q, &ctx->d_dinv};
    for (double** p : nodev) ESIMP_TRY(dev_alloc(ctx, p, ctx->ndof_loc));
    double** cellv[] = {&ctx->d_rho, &ctx->d_phys, &ctx->d_E, &ctx->d_dc, &ctx->d_dcf, &ctx->d_dv, &ctx->d_rho_new, &ctx->d_rho_old, &ctx->d_ctmp};
    for (double** p : cellv) ESIMP_TRY(dev_alloc(ctx, p, ctx->ncell_loc));
    return ESIMP_OK;
}

extern "C" {

int esimp_version(void) { return ESIMP_VERSION; }

int esimp_default_params(esimp_params* p) {
    if (!p) return ESIMP_ERR_ARG;
    memset(p, 0, sizeof(*p));
    p->E0 = 1.0;  // Optimization.jl:87-102
    p->Emin = 1e-9;
    p->nu = 0.3;
    p->p = 3.0;
    p->volume_fraction = 0.5;
    p->move_limit = 0.2;
    p->damping = 0.5;
    p->filter_radius = 1.5;
    p->filter_type = ESIMP_FILTER_SENSITIVITY;
    p->pcg_maxit = 200000;
    p->pcg_rtol = 1e-8;
    p->pcg_check_every = 25;
    p->matvec_variant = ESIMP_MATVEC_AUTO;
    p->warm_start = 1;
    p->precond = ESIMP_PRECOND_JACOBI;
    return ESIMP_OK;
}

int esimp_create(esimp_ctx** out, int device) {
    if (!out) return ESIMP_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return esimp_fail(nullptr, ESIMP_ERR_NODEVICE, "no CUDA device: libeasysimp_b200 has no CPU fallback");
    if (device < 0 && cudaGetDevice(&device) != cudaSuccess) return esimp_fail(nullptr, ESIMP_ERR_CUDA, "cudaGetDevice failed");
    if (device >= ndev) return esimp_fail(nullptr, ESIMP_ERR_ARG, "device index out of range");
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return esimp_fail(nullptr, ESIMP_ERR_CUDA, "cudaGetDeviceProperties failed");
    if (prop.major != 10)
        return esimp_fail(nullptr, ESIMP_ERR_NODEVICE,
                          std::string("device is sm_") + std::to_string(prop.major * 10 + prop.minor) + "; this library is built for sm_100a only");
    esimp_ctx* ctx = new esimp_ctx();
    ctx->device = device;
    esimp_default_params(&ctx->prm);
    if (const char* e = getenv("ESIMP_FUSED_CHUNK")) ctx->fused_chunk = atoi(e);
    if (const char* e = getenv("ESIMP_MG_NU")) ctx->mg_nu = std::max(1, atoi(e));
    if (const char* e = getenv("ESIMP_MG_COARSE_ITS")) ctx->mg_coarse_its = std::max(1, atoi(e));
    if (const char* e = getenv("ESIMP_MG_RATIO")) ctx->mg_ratio = std::max(1.5, atof(e));
    if (const char* e = getenv("ESIMP_MG_MIN_LAYERS")) ctx->mg_min_layers = std::max(1, atoi(e));
    if (const char* e = getenv("ESIMP_MG_REPLICATE")) ctx->mg_replicate = atoi(e);
    if (const char* e = getenv("ESIMP_MG_REPLICATE_CELLS")) ctx->mg_replicate_cells = atoll(e);
    if (const char* e = getenv("ESIMP_PCG_LAG")) ctx->mg_lag = atoi(e);
    if (const char* e = getenv("ESIMP_MG_TAIL")) ctx->mg_tail = atoi(e);
    if (const char* e = getenv("ESIMP_MG_TAIL_NODES")) ctx->mg_tail_nodes = atoll(e);
    auto bail = [&](const char* what) {
        std::string msg = std::string(what) + ": " + cudaGetErrorString(cudaGetLastError());
        esimp_destroy(ctx);  // frees whatever was created so far
        return esimp_fail(nullptr, ESIMP_ERR_CUDA, msg);
    };
    if (cudaSetDevice(device) != cudaSuccess) return bail("cudaSetDevice");
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) return bail("cudaStreamCreate");
    if (cudaMalloc(&ctx->d_partial, sizeof(double) * 16 * RED_BLOCKS) != cudaSuccess) return bail("cudaMalloc");
    if (cudaMalloc(&ctx->d_ticket, sizeof(unsigned int) * 8) != cudaSuccess) return bail("cudaMalloc");
    if (cudaMalloc(&ctx->d_scal, sizeof(double) * N_SCAL) != cudaSuccess) return bail("cudaMalloc");
    if (cudaMallocHost(&ctx->h_scal, sizeof(double) * (N_SCAL + 2)) != cudaSuccess) return bail("cudaMallocHost");
    cudaMemset(ctx->d_ticket, 0, sizeof(unsigned int) * 8);
    cudaMemset(ctx->d_scal, 0, sizeof(double) * N_SCAL);
    for (auto& e : ctx->ev) cudaEventCreate(&e);
    for (auto& e : ctx->lag_ev) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    *out = ctx;
    return ESIMP_OK;
}

int esimp_destroy(esimp_ctx* ctx) {
    if (!ctx) return ESIMP_OK;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    comm_destroy(ctx);
    free_mesh_state(ctx);
    hex8_forget_ctx(ctx);
    filter_forget_ctx(ctx);
    cudaFree(ctx->d_partial);
    cudaFree(ctx->d_partial_mv);
    cudaFree(ctx->d_ticket);
    cudaFree(ctx->d_scal);
    cudaFree(ctx->d_K0);
    cudaFreeHost(ctx->h_scal);
    for (auto& e : ctx->ev)
        if (e) cudaEventDestroy(e);
    for (auto& e : ctx->prof_ev) cudaEventDestroy(e);
    for (auto& e : ctx->lag_ev)
        if (e) cudaEventDestroy(e);
    for (auto& e : ctx->prof_coarse_ev)
        if (e) cudaEventDestroy(e);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return ESIMP_OK;
}

const char* esimp_last_error(const esimp_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int esimp_comm_unique_id(void* id128) {
    if (!id128) return ESIMP_ERR_ARG;
    std::string err;
    int s = comm_unique_id(id128, err);
    if (s < 0) g_create_error = err;
    return s;
}
int esimp_comm_init(esimp_ctx* ctx, int rank, int nranks, const void* id128) {
    CHECK_CTX();
    NEED(!ctx->structured && !ctx->unstructured, "esimp_comm_init must precede the mesh call");
    return comm_init(ctx, rank, nranks, id128);
}

int64_t esimp_n_dofs(const esimp_ctx* ctx) { return ctx ? ctx->n_dofs_global : 0; }
int64_t esimp_n_cells(const esimp_ctx* ctx) { return ctx ? ctx->n_cells_global : 0; }
int64_t esimp_launch_count(const esimp_ctx* ctx) { return ctx ? ctx->launches : 0; }

int esimp_set_params(esimp_ctx* ctx, const esimp_params* p) {
    CHECK_CTX();
    if (!p) return esimp_fail(ctx, ESIMP_ERR_ARG, "params is NULL");
    if (!(p->nu > -1.0 && p->nu < 0.5) || !(p->E0 > 0) || !(p->p >= 1.0)) return esimp_fail(ctx, ESIMP_ERR_ARG, "invalid material parameters");
    bool nu_changed = p->nu != ctx->prm.nu;
    if (p->precond != ESIMP_PRECOND_JACOBI && p->precond != ESIMP_PRECOND_GMG && p->precond != ESIMP_PRECOND_GMG_FP32)
        return esimp_fail(ctx, ESIMP_ERR_ARG, "unknown preconditioner");
    // The FP32 V-cycle is experimental: on its first GPU run (round 2, profiles/r2_gmg_fp32_first_gpu_run.log) it solved every case to
    // the FP64 answer but needed 285 instead of 184 CG iterations on an anisotropic grid, so it is not a supported option.
    if (p->precond == ESIMP_PRECOND_GMG_FP32) {
        const char* e = getenv("ESIMP_EXPERIMENTAL_GMG_FP32");
        if (!e || atoi(e) != 1)
            return esimp_fail(ctx, ESIMP_ERR_ARG, "ESIMP_PRECOND_GMG_FP32 is experimental (set ESIMP_EXPERIMENTAL_GMG_FP32=1 to try it)");
    }
    bool r_changed = p->filter_radius != ctx->prm.filter_radius;
    ctx->prm = *p;
    if (nu_changed) {
        if (ctx->unstructured) return esimp_fail(ctx, ESIMP_ERR_STATE, "set nu before esimp_set_mesh_unstructured");
        ctx->k0_valid = false;
    }
    if (r_changed && ctx->structured) ctx->filter_ready = false;
    ctx->density_set = false;  // E depends on E0, Emin, p
    return ESIMP_OK;
}

int esimp_set_grid_structured(esimp_ctx* ctx, int nx, int ny, int nz, double hx, double hy, double hz, const int64_t* dof_of_node) {
    CHECK_CTX();
    if (nx < 1 || ny < 1 || nz < 1 || !(hx > 0) || !(hy > 0) || !(hz > 0)) return esimp_fail(ctx, ESIMP_ERR_ARG, "bad grid dimensions");
    free_mesh_state(ctx);
    const int R = ctx->nranks, r = ctx->rank;
    const int ex0 = (int)((int64_t)nx * r / R), ex1 = (int)((int64_t)nx * (r + 1) / R);
    ctx->has_left = r > 0;
    ctx->has_right = r < R - 1;
    ctx->halo_e = R > 1 ? std::max(1, (int)floor(ctx->prm.filter_radius / hx + 1e-9)) : 0;
    // the thinnest slab of the partition (floor(nx / R) layers), not this rank's own: every rank takes the same decision, so a
    // grid that is too small for the rank count fails on all ranks together instead of leaving the others inside a collective
    if (R > 1 && nx / R < ctx->halo_e) return esimp_fail(ctx, ESIMP_ERR_ARG, "slab thinner than the filter halo");
    const int left = ctx->has_left ? 1 : 0, right = ctx->has_right ? 1 : 0;
    const int gl = ctx->has_left ? ctx->halo_e : 0, gr = ctx->has_right ? ctx->halo_e : 0;
    GridDev& g = ctx->g;
    g.nx = nx;
    g.ny = ny;
    g.nz = nz;
    g.gp0 = ex0 - left;
    g.NPX = (ex1 - ex0) + 1 + left;
    g.ge0 = ex0 - gl;
    g.NEX = (ex1 - ex0) + gl + gr;
    g.e2p = g.ge0 - g.gp0;
    g.nzp = esimp_node_pitch(nz);
    g.plane = (long long)(ny + 1) * g.nzp;
    g.eplane = (long long)ny * nz;
    ctx->hx = hx;
    ctx->hy = hy;
    ctx->hz = hz;
    ctx->own_p0 = left;
    ctx->own_p1 = g.NPX - right;
    ctx->own_e0 = gl;
    ctx->own_e1 = gl + (ex1 - ex0);
    ctx->mv_e0 = ctx->own_e0 - left;
    ctx->mv_e1 = ctx->own_e1;
    ctx->n_dofs_global = 3LL * (nx + 1) * (ny + 1) * (nz + 1);
    ctx->n_cells_global = (int64_t)nx * ny * nz;
    ctx->ndof_loc = 3LL * g.NPX * g.plane;
    ctx->ncell_loc = (int64_t)g.NEX * g.eplane;
    ctx->own_dof0 = 3LL * ctx->own_p0 * g.plane;
    ctx->own_dof1 = 3LL * ctx->own_p1 * g.plane;
    ctx->own_cell0 = (int64_t)ctx->own_e0 * g.eplane;
    ctx->own_cell1 = (int64_t)ctx->own_e1 * g.eplane;
    ctx->structured = true;
    ESIMP_TRY(alloc_state(ctx));
    ESIMP_TRY(dev_alloc(ctx, &ctx->d_mask, (int64_t)g.NPX * g.plane));
    if (dof_of_node) {
        ESIMP_TRY(dev_alloc(ctx, &ctx->d_dof_of_node, ctx->n_dofs_global, false));
        ESIMP_CUDA(cudaMemcpy(ctx->d_dof_of_node, dof_of_node, sizeof(int64_t) * ctx->n_dofs_global, cudaMemcpyHostToDevice));
        // keep a host inverse for esimp_set_dirichlet
    }
    if (ctx->nranks > 1) ESIMP_TRY(comm_p2p_setup(ctx, 3 * g.plane));  // collective: node-plane halos over peer memory when IPC works
    return ESIMP_OK;
}

int esimp_set_mesh_unstructured(esimp_ctx* ctx, int celltype, int64_t n_cells, int64_t n_nodes, const double* xyz, const int64_t* conn,
                                const int64_t* celldofs) {
    CHECK_CTX();
    if (!xyz || !conn || !celldofs || n_cells < 1 || n_nodes < 1) return esimp_fail(ctx, ESIMP_ERR_ARG, "bad mesh arguments");
    NEED(ctx->nranks == 1, "unstructured meshes run as replicas only (no slab partition)");
    free_mesh_state(ctx);
    ESIMP_TRY(um_build(ctx, celltype, n_cells, n_nodes, xyz, conn, celldofs));
    ctx->unstructured = true;
    ctx->n_dofs_global = ctx->ndof_loc = ctx->um.n_dofs;
    ctx->n_cells_global = ctx->ncell_loc = n_cells;
    ctx->own_dof0 = 0;
    ctx->own_dof1 = ctx->ndof_loc;
    ctx->own_cell0 = 0;
    ctx->own_cell1 = n_cells;
    ESIMP_TRY(alloc_state(ctx));
    ESIMP_TRY(dev_alloc(ctx, &ctx->d_mask, ctx->ndof_loc));
    return ESIMP_OK;
}

int esimp_set_dirichlet(esimp_ctx* ctx, const int64_t* dofs, int64_t n) {
    CHECK_CTX();
    NEED(ctx->structured || ctx->unstructured, "no mesh");
    if (n < 0 || (n > 0 && !dofs)) return esimp_fail(ctx, ESIMP_ERR_ARG, "bad dirichlet arguments");
    std::vector<int64_t> fixed;
    if (ctx->unstructured) {
        std::vector<uint8_t> mask(ctx->ndof_loc, 0);
        for (int64_t i = 0; i < n; ++i) {
            if (dofs[i] < 0 || dofs[i] >= ctx->ndof_loc) return esimp_fail(ctx, ESIMP_ERR_ARG, "prescribed dof out of range");
            if (!mask[dofs[i]]) fixed.push_back(dofs[i]);
            mask[dofs[i]] = 1;
        }
        ESIMP_CUDA(cudaMemcpy(ctx->d_mask, mask.data(), mask.size(), cudaMemcpyHostToDevice));
    } else {
        const GridDev& g = ctx->g;
        ctx->prescribed_nc.clear();
        std::vector<uint8_t> mask((size_t)g.NPX * g.plane, 0);
        std::vector<int64_t> inv;  // caller dof -> 3*node + c
        if (ctx->d_dof_of_node) {
            std::vector<int64_t> don(ctx->n_dofs_global);
            ESIMP_CUDA(cudaMemcpy(don.data(), ctx->d_dof_of_node, sizeof(int64_t) * don.size(), cudaMemcpyDeviceToHost));
            inv.assign(ctx->n_dofs_global, -1);
            for (int64_t q = 0; q < ctx->n_dofs_global; ++q)
                if (don[q] >= 0 && don[q] < ctx->n_dofs_global) inv[don[q]] = q;
        }
        for (int64_t i = 0; i < n; ++i) {
            int64_t d = dofs[i];
            if (d < 0 || d >= ctx->n_dofs_global) return esimp_fail(ctx, ESIMP_ERR_ARG, "prescribed dof out of range");
            int64_t q = inv.empty() ? d : inv[d];
            if (q < 0) return esimp_fail(ctx, ESIMP_ERR_ARG, "prescribed dof not in dof_of_node");
            ctx->prescribed_nc.push_back(q);
            int c = (int)(q % 3);
            int64_t nref = q / 3;
            int gi = (int)(nref % (g.nx + 1));
            int j = (int)((nref / (g.nx + 1)) % (g.ny + 1));
            int k = (int)(nref / ((int64_t)(g.nx + 1) * (g.ny + 1)));
            int lp = gi - g.gp0;
            if (lp < 0 || lp >= g.NPX) continue;
            uint8_t& m = mask[(size_t)g.node(lp, j, k)];
            if (!((m >> c) & 1) && lp >= ctx->own_p0 && lp < ctx->own_p1) fixed.push_back(g.dof(lp, c, j, k));
            m |= (uint8_t)(1 << c);
        }
        ESIMP_CUDA(cudaMemcpy(ctx->d_mask, mask.data(), mask.size(), cudaMemcpyHostToDevice));
    }
    ctx->bc_version++;
    ctx->n_fixed = (int64_t)fixed.size();
    ESIMP_TRY(dev_alloc(ctx, &ctx->d_fixed, ctx->n_fixed, false));
    if (ctx->n_fixed) ESIMP_CUDA(cudaMemcpy(ctx->d_fixed, fixed.data(), sizeof(int64_t) * ctx->n_fixed, cudaMemcpyHostToDevice));
    ctx->density_set = false;  // the Jacobi diagonal depends on the mask: esimp_set_density must follow
    return ESIMP_OK;
}

static int upload_dofs(esimp_ctx* ctx, const double* h, double* d_local) {
    ESIMP_TRY(ensure_stage(ctx, ctx->n_dofs_global));
    ESIMP_CUDA(cudaMemcpyAsync(ctx->d_stage, h, sizeof(double) * ctx->n_dofs_global, cudaMemcpyHostToDevice, ctx->stream));
    return vec_gather_to_device(ctx, ctx->d_stage, d_local);
}
static int download_dofs(esimp_ctx* ctx, const double* d_local, double* h) {
    ESIMP_TRY(ensure_stage(ctx, ctx->n_dofs_global));
    if (ctx->nranks > 1) ESIMP_CUDA(cudaMemsetAsync(ctx->d_stage, 0, sizeof(double) * ctx->n_dofs_global, ctx->stream));
    ESIMP_TRY(vec_scatter_from_device(ctx, d_local, ctx->d_stage));
    ESIMP_CUDA(cudaMemcpyAsync(h, ctx->d_stage, sizeof(double) * ctx->n_dofs_global, cudaMemcpyDeviceToHost, ctx->stream));
    ESIMP_CUDA(cudaStreamSynchronize(ctx->stream));
    return ESIMP_OK;
}
static int upload_cells(esimp_ctx* ctx, const double* h, double* d_local) {
    ESIMP_TRY(ensure_stage(ctx, ctx->n_cells_global));
    ESIMP_CUDA(cudaMemcpyAsync(ctx->d_stage, h, sizeof(double) * ctx->n_cells_global, cudaMemcpyHostToDevice, ctx->stream));
    return cell_gather_to_device(ctx, ctx->d_stage, d_local);
}
static int download_cells(esimp_ctx* ctx, const double* d_local, double* h) {
    ESIMP_TRY(ensure_stage(ctx, ctx->n_cells_global));
    if (ctx->nranks > 1) ESIMP_CUDA(cudaMemsetAsync(ctx->d_stage, 0, sizeof(double) * ctx->n_cells_global, ctx->stream));
    ESIMP_TRY(cell_scatter_from_device(ctx, d_local, ctx->d_stage));
    ESIMP_CUDA(cudaMemcpyAsync(h, ctx->d_stage, sizeof(double) * ctx->n_cells_global, cudaMemcpyDeviceToHost, ctx->stream));
    ESIMP_CUDA(cudaStreamSynchronize(ctx->stream));
    return ESIMP_OK;
}

// Multi-rank structured grids: only the x-slab of a cell array moves.  The caller's numbering is x fastest (cell = i + nx (j + ny k)),
// so a slab of w layers is ny*nz runs of w doubles with pitch nx: one cudaMemcpy2DAsync each way instead of the whole array.
__global__ void k_gather_cells_slab(GridDev g, int l0, int w, const double* __restrict__ src, double* __restrict__ dst) {
    const long long n = (long long)w * g.eplane;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(i % g.nz);
        const long long t = i / g.nz;
        const int j = (int)(t % g.ny), l = (int)(t / g.ny);
        dst[g.cell(l0 + l, j, k)] = src[l + (long long)w * (j + (long long)g.ny * k)];
    }
}
__global__ void k_scatter_cells_slab(GridDev g, int l0, int w, const double* __restrict__ src, double* __restrict__ dst) {
    const long long n = (long long)w * g.eplane;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(i % g.nz);
        const long long t = i / g.nz;
        const int j = (int)(t % g.ny), l = (int)(t / g.ny);
        dst[l + (long long)w * (j + (long long)g.ny * k)] = src[g.cell(l0 + l, j, k)];
    }
}
// host array -> the stored layers [l0, l0 + w) of a local cell array
static int upload_cells_slab(esimp_ctx* ctx, const double* h, double* d_local, int l0, int w) {
    const GridDev& g = ctx->g;
    ESIMP_TRY(ensure_stage(ctx, (int64_t)w * g.eplane));
    ESIMP_CUDA(cudaMemcpy2DAsync(ctx->d_stage, sizeof(double) * w, h + g.ge0 + l0, sizeof(double) * g.nx, sizeof(double) * w, (size_t)g.eplane,
                                 cudaMemcpyHostToDevice, ctx->stream));
    k_gather_cells_slab<<<RED_BLOCKS, RED_THREADS, 0, ctx->stream>>>(g, l0, w, ctx->d_stage, d_local);
    ctx->launches++;
    ESIMP_CUDA(cudaGetLastError());
    return ESIMP_OK;
}
// the OWNED layers of a local cell array -> the matching entries of the host array (nothing else is written)
static int download_cells_slab(esimp_ctx* ctx, const double* d_local, double* h) {
    const GridDev& g = ctx->g;
    const int l0 = ctx->own_e0, w = ctx->own_e1 - ctx->own_e0;
    ESIMP_TRY(ensure_stage(ctx, (int64_t)w * g.eplane));
    k_scatter_cells_slab<<<RED_BLOCKS, RED_THREADS, 0, ctx->stream>>>(g, l0, w, d_local, ctx->d_stage);
    ctx->launches++;
    ESIMP_CUDA(cudaGetLastError());
    ESIMP_CUDA(cudaMemcpy2DAsync(h + g.ge0 + l0, sizeof(double) * g.nx, ctx->d_stage, sizeof(double) * w, sizeof(double) * w, (size_t)g.eplane,
                                 cudaMemcpyDeviceToHost, ctx->stream));
    ESIMP_CUDA(cudaStreamSynchronize(ctx->stream));
    return ESIMP_OK;
}

int esimp_set_load(esimp_ctx* ctx, const double* f) {
    CHECK_CTX();
    NEED(ctx->structured || ctx->unstructured, "no mesh");
    if (!f) return esimp_fail(ctx, ESIMP_ERR_ARG, "f is NULL");
    ESIMP_TRY(upload_dofs(ctx, f, ctx->d_f));
    if (ctx->body_on) {  // f is rebuilt from this density-independent part on every assembly
        ESIMP_CUDA(cudaMemcpyAsync(ctx->d_fload, ctx->d_f, sizeof(double) * ctx->ndof_loc, cudaMemcpyDeviceToDevice, ctx->stream));
        ctx->density_set = false;
    }
    ESIMP_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->load_set = true;
    return ESIMP_OK;
}

int esimp_set_body_force(esimp_ctx* ctx, const double* accel3, double base_density) {
    CHECK_CTX();
    NEED(ctx->structured || ctx->unstructured, "no mesh");
    const size_t bytes = sizeof(double) * ctx->ndof_loc;
    if (!accel3) {  // off: f goes back to the density-independent loads
        if (ctx->body_on) ESIMP_CUDA(cudaMemcpyAsync(ctx->d_f, ctx->d_fload, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
        ctx->body_on = false;
        ESIMP_CUDA(cudaStreamSynchronize(ctx->stream));
        return ESIMP_OK;
    }
    if (!ctx->body_on) {
        if (!ctx->d_fload) ESIMP_TRY(dev_alloc(ctx, &ctx->d_fload, ctx->ndof_loc));
        if (ctx->load_set)
            ESIMP_CUDA(cudaMemcpyAsync(ctx->d_fload, ctx->d_f, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
        else
            ESIMP_CUDA(cudaMemsetAsync(ctx->d_fload, 0, bytes, ctx->stream));
    }
    for (int c = 0; c < 3; ++c) ctx->body_b[c] = accel3[c];
    ctx->body_base = base_density;
    ctx->body_on = true;
    ctx->load_set = true;      // a body force alone is a valid right-hand side
    ctx->density_set = false;  // f now depends on the density: esimp_set_density / esimp_iterate rebuild it
    ESIMP_CUDA(cudaStreamSynchronize(ctx->stream));
    return ESIMP_OK;
}

int esimp_body_force(esimp_ctx* ctx, const double* accel3, const double* density_data, double* f_body) {
    CHECK_CTX();
    NEED(ctx->structured || ctx->unstructured, "no mesh");
    if (!accel3 || !density_data || !f_body) return esimp_fail(ctx, ESIMP_ERR_ARG, "NULL argument");
    // stage call: borrows d_ctmp / d_q as scratch
    ESIMP_TRY(upload_cells(ctx, density_data, ctx->d_ctmp));
    ESIMP_TRY(body_force_into(ctx, ctx->d_ctmp, 1.0, accel3, nullptr, ctx->d_q));
    return download_dofs(ctx, ctx->d_q, f_body);
}

int esimp_filter_build_structured(esimp_ctx* ctx) {
    CHECK_CTX();
    NEED(ctx->structured, "structured grid required");
    return filter_build_stencil(ctx);
}
int esimp_filter_build_csr(esimp_ctx* ctx, const int64_t* rowptr, const int64_t* col, const double* centers, const double* volumes,
                           double R) {
    CHECK_CTX();
    NEED(ctx->unstructured, "unstructured mesh required");
    if (!rowptr || !col || !centers || !volumes) return esimp_fail(ctx, ESIMP_ERR_ARG, "NULL filter argument");
    ctx->prm.filter_radius = R;
    return um_filter_build(ctx, rowptr, col, centers, volumes, R);
}

// E from d_phys (all stored layers), Jacobi diagonal, diag_value: everything the operator needs
static int prepare_operator(esimp_ctx* ctx) {
    ESIMP_TRY(simp_modulus(ctx));
    if (ctx->structured)
        ESIMP_TRY(hex8_diagonal(ctx));
    else
        ESIMP_TRY(um_diagonal(ctx));
    if ((ctx->prm.precond == ESIMP_PRECOND_GMG || ctx->prm.precond == ESIMP_PRECOND_GMG_FP32) && ctx->structured) ESIMP_TRY(mg_setup(ctx));
    ESIMP_TRY(body_force_refresh_rhs(ctx));  // Optimization.jl:300-311
    ctx->density_set = true;
    return ESIMP_OK;
}

int esimp_set_density(esimp_ctx* ctx, const double* rho_phys) {
    CHECK_CTX();
    NEED(ctx->structured || ctx->unstructured, "no mesh");
    if (!rho_phys) return esimp_fail(ctx, ESIMP_ERR_ARG, "rho is NULL");
    ESIMP_TRY(upload_cells(ctx, rho_phys, ctx->d_phys));
    return prepare_operator(ctx);
}

int esimp_matvec(esimp_ctx* ctx, const double* u, double* y) {
    CHECK_CTX();
    NEED(ctx->density_set, "esimp_set_density first");
    if (!u || !y) return esimp_fail(ctx, ESIMP_ERR_ARG, "NULL vector");
    ESIMP_TRY(upload_dofs(ctx, u, ctx->d_p));
    ESIMP_TRY(op_apply_general(ctx, ctx->d_p, ctx->d_r, ctx->d_q));
    return download_dofs(ctx, ctx->d_q, y);
}

int esimp_diagonal(esimp_ctx* ctx, double* d) {
    CHECK_CTX();
    NEED(ctx->density_set, "esimp_set_density first");
    if (!d) return esimp_fail(ctx, ESIMP_ERR_ARG, "NULL vector");
    ESIMP_TRY(download_dofs(ctx, ctx->d_dinv, d));
    for (int64_t i = 0; i < ctx->n_dofs_global; ++i)
        d[i] = d[i] != 0.0 ? 1.0 / d[i] : (ctx->structured ? ctx->diag_value : 0.0);  // structured: Dinv is stored as 0 on prescribed dofs
    return ESIMP_OK;
}

// zero f and u on prescribed dofs (apply!: f[d] = 0; homogeneous Dirichlet => u[d] = 0)
static int mask_rhs_and_guess(esimp_ctx* ctx) {
    const double m = ctx->diag_value;
    ctx->diag_value = 0.0;
    int s = vec_fixup_dirichlet(ctx, ctx->d_f, ctx->d_f);
    if (s >= 0) s = vec_fixup_dirichlet(ctx, ctx->d_u, ctx->d_u);
    ctx->diag_value = m;
    return s;
}

static int solve_internal(esimp_ctx* ctx, int* iters, double* relres) {
    NEED(ctx->density_set, "esimp_set_density first");
    NEED(ctx->load_set, "esimp_set_load first");
    if (!ctx->prm.warm_start) ESIMP_CUDA(cudaMemsetAsync(ctx->d_u, 0, sizeof(double) * ctx->ndof_loc, ctx->stream));
    ESIMP_TRY(mask_rhs_and_guess(ctx));
    if (ctx->prm.precond == ESIMP_PRECOND_GMG && ctx->structured) return pcg_solve_mg(ctx, iters, relres);
    if (ctx->prm.precond == ESIMP_PRECOND_GMG_FP32 && ctx->structured) return pcg_solve_mg32(ctx, iters, relres);
    return pcg_solve(ctx, iters, relres);
}

int esimp_solve(esimp_ctx* ctx, double* u_out, int* iters, double* relres) {
    CHECK_CTX();
    int st = solve_internal(ctx, iters, relres);
    if (st < 0) return st;
    if (u_out) ESIMP_TRY(download_dofs(ctx, ctx->d_u, u_out));
    return st;
}

static int sens_internal(esimp_ctx* ctx) {
    if (ctx->structured) {
        if (ctx->nranks > 1) ESIMP_TRY(comm_halo_nodes(ctx, ctx->d_u));
        return hex8_sens_energy(ctx);
    }
    return um_sens_energy(ctx);
}

int esimp_sensitivities(esimp_ctx* ctx, double* dc, double* energy) {
    CHECK_CTX();
    NEED(ctx->density_set, "esimp_set_density first");
    ESIMP_TRY(sens_internal(ctx));
    ESIMP_TRY(fetch_scalars(ctx));
    if (energy) *energy = ctx->h_scal[SC_ENERGY];
    if (dc) ESIMP_TRY(download_cells(ctx, ctx->d_dc, dc));
    return ESIMP_OK;
}

int esimp_filter_sens(esimp_ctx* ctx, const double* rho, const double* dc, double* out) {
    CHECK_CTX();
    if (!rho || !dc || !out) return esimp_fail(ctx, ESIMP_ERR_ARG, "NULL array");
    ESIMP_TRY(upload_cells(ctx, rho, ctx->d_rho_old));
    ESIMP_TRY(upload_cells(ctx, dc, ctx->d_rho_new));
    ESIMP_TRY(filter_sens_apply(ctx, ctx->d_rho_old, ctx->d_rho_new, ctx->d_ctmp));
    return download_cells(ctx, ctx->d_ctmp, out);
}
int esimp_filter_dens(esimp_ctx* ctx, const double* rho, double* out) {
    CHECK_CTX();
    if (!rho || !out) return esimp_fail(ctx, ESIMP_ERR_ARG, "NULL array");
    ESIMP_TRY(upload_cells(ctx, rho, ctx->d_rho_old));
    ESIMP_TRY(filter_dens_apply(ctx, ctx->d_rho_old, ctx->d_ctmp));
    return download_cells(ctx, ctx->d_ctmp, out);
}
int esimp_filter_dens_T(esimp_ctx* ctx, const double* s, double* out) {
    CHECK_CTX();
    if (!s || !out) return esimp_fail(ctx, ESIMP_ERR_ARG, "NULL array");
    ESIMP_TRY(upload_cells(ctx, s, ctx->d_rho_old));
    ESIMP_TRY(filter_densT_apply(ctx, ctx->d_rho_old, ctx->d_ctmp));
    return download_cells(ctx, ctx->d_ctmp, out);
}

int esimp_oc(esimp_ctx* ctx, const double* rho, const double* dc, const double* dv, double* rho_new, double* lambda, int* steps) {
    CHECK_CTX();
    if (!rho || !dc || !dv || !rho_new) return esimp_fail(ctx, ESIMP_ERR_ARG, "NULL array");
    // stage call: borrows rho_old / dcf / dc as scratch, leaves rho, dv of the loop state untouched
    double* d_dvtmp = nullptr;
    ESIMP_TRY(dev_alloc(ctx, &d_dvtmp, ctx->ncell_loc, false));
    int st = upload_cells(ctx, rho, ctx->d_rho_old);
    if (st >= 0) st = upload_cells(ctx, dc, ctx->d_dcf);
    if (st >= 0) st = upload_cells(ctx, dv, d_dvtmp);
    if (st >= 0) st = oc_update(ctx, ctx->d_rho_old, ctx->d_dcf, d_dvtmp, ctx->d_rho_new, lambda, steps);
    int st2 = st >= 0 ? download_cells(ctx, ctx->d_rho_new, rho_new) : st;
    cudaFree(d_dvtmp);
    return st2 < 0 ? st2 : st;
}

__global__ void k_fill(double* a, double v, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) a[i] = v;
}
__global__ void k_scale_by(const double* __restrict__ vol, double s, double* __restrict__ out, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = vol[i] * s;
}

int esimp_init_design(esimp_ctx* ctx, const double* rho) {
    CHECK_CTX();
    NEED(ctx->structured || ctx->unstructured, "no mesh");
    NEED(ctx->filter_ready, "build the filter first");
    const long long n = ctx->ncell_loc;
    if (rho) {
        ESIMP_TRY(upload_cells(ctx, rho, ctx->d_rho));
    } else {
        k_fill<<<RED_BLOCKS, RED_THREADS, 0, ctx->stream>>>(ctx->d_rho, ctx->prm.volume_fraction, n);  // Optimization.jl:222
        ctx->launches++;
    }
    // volume sensitivities V/Vtot, chain-rule transformed once for the density filter (Optimization.jl:242-248)
    if (ctx->structured) {
        k_fill<<<RED_BLOCKS, RED_THREADS, 0, ctx->stream>>>(ctx->d_ctmp, 1.0 / (double)ctx->n_cells_global, n);
        ctx->launches++;
    } else {
        ESIMP_TRY(reduce_sum_scaled(ctx, ctx->um.d_vol, nullptr, 1.0, 0, n, SC_TMP0));
        ESIMP_TRY(fetch_scalars(ctx));
        k_scale_by<<<RED_BLOCKS, RED_THREADS, 0, ctx->stream>>>(ctx->um.d_vol, 1.0 / ctx->h_scal[SC_TMP0], ctx->d_ctmp, n);
        ctx->launches++;
    }
    if (ctx->prm.filter_type == ESIMP_FILTER_DENSITY) {
        ESIMP_TRY(filter_densT_apply(ctx, ctx->d_ctmp, ctx->d_dv));
        if (ctx->nranks > 1) ESIMP_TRY(comm_halo_cells(ctx, ctx->d_dv));
    } else {
        ESIMP_CUDA(cudaMemcpyAsync(ctx->d_dv, ctx->d_ctmp, sizeof(double) * n, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    ESIMP_CUDA(cudaMemsetAsync(ctx->d_u, 0, sizeof(double) * ctx->ndof_loc, ctx->stream));
    ESIMP_CUDA(cudaGetLastError());
    ctx->design_set = true;
    return ESIMP_OK;
}

// physical densities from the design (density filter or identity), then E + diagonal
static int physical_from_design(esimp_ctx* ctx) {
    const size_t bytes = sizeof(double) * ctx->ncell_loc;
    if (ctx->prm.filter_type == ESIMP_FILTER_DENSITY) {
        ESIMP_TRY(filter_dens_apply(ctx, ctx->d_rho, ctx->d_phys));  // Optimization.jl:280-282
        if (ctx->nranks > 1) ESIMP_TRY(comm_halo_cells(ctx, ctx->d_phys));
    } else {
        ESIMP_CUDA(cudaMemcpyAsync(ctx->d_phys, ctx->d_rho, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    return prepare_operator(ctx);
}

int esimp_iterate(esimp_ctx* ctx, esimp_iter_stats* out) {
    CHECK_CTX();
    NEED(ctx->design_set, "esimp_init_design first");
    NEED(ctx->load_set, "esimp_set_load first");
    const bool use_df = ctx->prm.filter_type == ESIMP_FILTER_DENSITY;
    const size_t cbytes = sizeof(double) * ctx->ncell_loc;
    int status = 0;
    cudaEventRecord(ctx->ev[0], ctx->stream);
    ESIMP_CUDA(cudaMemcpyAsync(ctx->d_rho_old, ctx->d_rho, cbytes, cudaMemcpyDeviceToDevice, ctx->stream));  // :273
    ESIMP_TRY(physical_from_design(ctx));
    cudaEventRecord(ctx->ev[1], ctx->stream);
    int pcg_iters = 0;
    double relres = 0.0;
    int st = solve_internal(ctx, &pcg_iters, &relres);  // :317
    if (st < 0) return st;
    status |= st;
    cudaEventRecord(ctx->ev[2], ctx->stream);
    ESIMP_TRY(sens_internal(ctx));  // :320 energy, :328 dc
    {
        const double* vol = ctx->unstructured ? ctx->um.d_vol : nullptr;
        ESIMP_TRY(reduce_sum_scaled(ctx, ctx->d_phys, vol, vol ? 1.0 : ctx->hx * ctx->hy * ctx->hz, ctx->own_cell0, ctx->own_cell1, SC_VOL));  // :321
    }
    ESIMP_TRY(fetch_scalars(ctx));
    const double energy = ctx->h_scal[SC_ENERGY], volume = ctx->h_scal[SC_VOL];
    if (ctx->nranks > 1) ESIMP_TRY(comm_halo_cells(ctx, ctx->d_dc));
    if (use_df)
        ESIMP_TRY(filter_densT_apply(ctx, ctx->d_dc, ctx->d_dcf));  // :344
    else
        ESIMP_TRY(filter_sens_apply(ctx, ctx->d_rho, ctx->d_dc, ctx->d_dcf));  // :351
    if (ctx->nranks > 1) ESIMP_TRY(comm_halo_cells(ctx, ctx->d_dcf));
    cudaEventRecord(ctx->ev[3], ctx->stream);
    double lambda = 0.0;
    int oc_steps = 0;
    st = oc_update(ctx, ctx->d_rho, ctx->d_dcf, ctx->d_dv, ctx->d_rho_new, &lambda, &oc_steps);  // :360
    if (st < 0) return st;
    status |= st;
    std::swap(ctx->d_rho, ctx->d_rho_new);
    ESIMP_TRY(reduce_metrics(ctx));  // :374-379
    cudaEventRecord(ctx->ev[4], ctx->stream);
    ESIMP_TRY(fetch_scalars(ctx));
    if (out) {
        memset(out, 0, sizeof(*out));
        out->energy = energy;
        out->volume = volume;
        out->change = ctx->h_scal[SC_CHANGE];
        out->lambda = lambda;
        out->grayness = ctx->h_scal[SC_GRAY] / (double)ctx->n_cells_global;
        out->max_u = ctx->h_scal[SC_MAXU];
        out->pcg_relres = relres;
        out->pcg_iters = pcg_iters;
        out->oc_steps = oc_steps;
        out->status = status;
        float ms;
        cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
        out->ms_filter = ms;
        cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2]);
        out->ms_solve = ms;
        cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]);
        out->ms_sens = ms;
        cudaEventElapsedTime(&ms, ctx->ev[3], ctx->ev[4]);
        out->ms_oc = ms;
        cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[4]);
        out->ms_total = ms;
    }
    return status;
}

int esimp_final_analysis(esimp_ctx* ctx, double* energy, double* volume, int* pcg_iters) {
    CHECK_CTX();
    NEED(ctx->design_set, "esimp_init_design first");
    ESIMP_TRY(physical_from_design(ctx));  // Optimization.jl:498-515
    double relres;
    int st = solve_internal(ctx, pcg_iters, &relres);  // :532
    if (st < 0) return st;
    ESIMP_TRY(sens_internal(ctx));
    const double* vol = ctx->unstructured ? ctx->um.d_vol : nullptr;
    ESIMP_TRY(reduce_sum_scaled(ctx, ctx->d_phys, vol, vol ? 1.0 : ctx->hx * ctx->hy * ctx->hz, ctx->own_cell0, ctx->own_cell1, SC_VOL));
    ESIMP_TRY(fetch_scalars(ctx));
    if (energy) *energy = ctx->h_scal[SC_ENERGY];  // :534
    if (volume) *volume = ctx->h_scal[SC_VOL];     // :535
    return st;
}

int esimp_stress_nq(const esimp_ctx* ctx) { return !ctx ? 0 : ctx->structured ? 8 : ctx->unstructured ? (ctx->um.npe == 8 ? 8 : 1) : 0; }

int esimp_stress_fields(esimp_ctx* ctx, double* von_mises_q1, double* von_mises_avg, double* element_energy, double* max_von_mises,
                        int64_t* max_cell) {
    CHECK_CTX();
    NEED(ctx->density_set, "esimp_set_density / esimp_iterate first");
    ESIMP_TRY(stress_fields(ctx, max_von_mises, max_cell));
    if (von_mises_q1) ESIMP_TRY(download_cells(ctx, ctx->d_ctmp, von_mises_q1));
    if (von_mises_avg) ESIMP_TRY(download_cells(ctx, ctx->d_dcf, von_mises_avg));
    if (element_energy) ESIMP_TRY(download_cells(ctx, ctx->d_rho_new, element_energy));
    return ESIMP_OK;
}

int esimp_stress_tensors(esimp_ctx* ctx, int64_t cell0, int64_t n, double* sigma) {
    CHECK_CTX();
    NEED(ctx->density_set, "esimp_set_density / esimp_iterate first");
    if (!sigma) return esimp_fail(ctx, ESIMP_ERR_ARG, "NULL array");
    if (cell0 < 0 || n < 0 || cell0 + n > ctx->n_cells_global) return esimp_fail(ctx, ESIMP_ERR_ARG, "cell range outside the mesh");
    if (n == 0) return ESIMP_OK;
    const int64_t len = n * esimp_stress_nq(ctx) * 6;
    double* d_out = nullptr;
    ESIMP_CUDA(cudaMalloc(&d_out, sizeof(double) * len));
    int st = stress_tensors(ctx, cell0, n, d_out);
    if (st >= 0 && cudaMemcpyAsync(sigma, d_out, sizeof(double) * len, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)
        st = esimp_fail(ctx, ESIMP_ERR_CUDA, "stress tensor download failed");
    cudaStreamSynchronize(ctx->stream);
    cudaFree(d_out);
    return st;
}

int esimp_get_design(esimp_ctx* ctx, double* rho) {
    CHECK_CTX();
    if (!rho) return esimp_fail(ctx, ESIMP_ERR_ARG, "NULL array");
    // multi-rank: only the cells this rank owns are written (the entries of other ranks' slabs stay untouched)
    if (ctx->nranks > 1 && ctx->structured) return download_cells_slab(ctx, ctx->d_rho, rho);
    return download_cells(ctx, ctx->d_rho, rho);
}
int esimp_get_physical_density(esimp_ctx* ctx, double* rho) {
    CHECK_CTX();
    if (!rho) return esimp_fail(ctx, ESIMP_ERR_ARG, "NULL array");
    return download_cells(ctx, ctx->d_phys, rho);
}
int esimp_get_displacement(esimp_ctx* ctx, double* u) {
    CHECK_CTX();
    if (!u) return esimp_fail(ctx, ESIMP_ERR_ARG, "NULL array");
    return download_dofs(ctx, ctx->d_u, u);
}
int esimp_set_displacement(esimp_ctx* ctx, const double* u) {
    CHECK_CTX();
    if (!u) return esimp_fail(ctx, ESIMP_ERR_ARG, "NULL array");
    ESIMP_TRY(upload_dofs(ctx, u, ctx->d_u));
    ESIMP_CUDA(cudaStreamSynchronize(ctx->stream));
    return ESIMP_OK;
}

int esimp_set_design(esimp_ctx* ctx, const double* rho) {
    CHECK_CTX();
    NEED(ctx->design_set, "esimp_init_design first");
    if (!rho) return esimp_fail(ctx, ESIMP_ERR_ARG, "NULL array");
    // multi-rank: the caller's array is authoritative only on the cells this rank owns (esimp_get_design fills
    // nothing else): only that slab is uploaded, the ghost layers are refreshed from their owners
    if (ctx->nranks > 1 && ctx->structured) {
        ESIMP_TRY(upload_cells_slab(ctx, rho, ctx->d_rho, ctx->own_e0, ctx->own_e1 - ctx->own_e0));
        ESIMP_TRY(comm_halo_cells(ctx, ctx->d_rho));
    } else {
        ESIMP_TRY(upload_cells(ctx, rho, ctx->d_rho));
    }
    ESIMP_CUDA(cudaStreamSynchronize(ctx->stream));
    return ESIMP_OK;
}

int esimp_profile_enable(esimp_ctx* ctx, int on) {
    CHECK_CTX();
    ESIMP_CUDA(cudaStreamSynchronize(ctx->stream));
    ESIMP_TRY(prof_collect(ctx));
    if (on && ctx->prof_ev.empty()) {
        ctx->prof_ev.resize(2 * 4096);
        ctx->prof_tag.assign(4096, 0);
        for (auto& e : ctx->prof_ev) ESIMP_CUDA(cudaEventCreate(&e));
        for (auto& e : ctx->prof_coarse_ev) ESIMP_CUDA(cudaEventCreate(&e));
    }
    ctx->prof_on = on < 0 ? 0 : (on > 2 ? 2 : on);
    for (int c = 0; c < PROF_NCAT; ++c) {
        ctx->prof_ms_cat[c] = 0.0;
        ctx->prof_cnt_cat[c] = 0;
    }
    return ESIMP_OK;
}
// 16 B per owned dof (read u, write y) + 8 B per owned cell (read E); the row pads of the device layout are not counted
static double algorithmic_bytes_per_apply(const esimp_ctx* ctx) {
    double dofs = (double)(ctx->own_dof1 - ctx->own_dof0);
    if (ctx->structured) dofs = 3.0 * (ctx->own_p1 - ctx->own_p0) * (double)(ctx->g.ny + 1) * (double)(ctx->g.nz + 1);
    return 16.0 * dofs + 8.0 * (double)(ctx->own_cell1 - ctx->own_cell0);
}
int esimp_profile_matvec(esimp_ctx* ctx, double* ms_sum, int64_t* count, double* bytes_per_apply) {
    CHECK_CTX();
    ESIMP_CUDA(cudaStreamSynchronize(ctx->stream));
    ESIMP_TRY(prof_collect(ctx));
    if (ms_sum) *ms_sum = ctx->prof_ms_cat[PROF_OP_PLAIN];
    if (count) *count = ctx->prof_cnt_cat[PROF_OP_PLAIN];
    if (bytes_per_apply)
        *bytes_per_apply = algorithmic_bytes_per_apply(ctx);
    return ESIMP_OK;
}

int esimp_profile_operator_all(esimp_ctx* ctx, double* ms_sum_all, int64_t* count_all) {
    CHECK_CTX();
    ESIMP_CUDA(cudaStreamSynchronize(ctx->stream));
    ESIMP_TRY(prof_collect(ctx));
    if (ms_sum_all) *ms_sum_all = ctx->prof_ms_cat[PROF_OP_PLAIN] + ctx->prof_ms_cat[PROF_OP_SMOOTH] + ctx->prof_ms_cat[PROF_OP_RESID];
    if (count_all) *count_all = ctx->prof_cnt_cat[PROF_OP_PLAIN] + ctx->prof_cnt_cat[PROF_OP_SMOOTH] + ctx->prof_cnt_cat[PROF_OP_RESID];
    return ESIMP_OK;
}

int esimp_profile_kernels(esimp_ctx* ctx, int n, double* ms_sum, int64_t* count) {
    CHECK_CTX();
    if (n < 0 || n > PROF_NCAT || !ms_sum || !count) return esimp_fail(ctx, ESIMP_ERR_ARG, "bad arguments");
    ESIMP_CUDA(cudaStreamSynchronize(ctx->stream));
    ESIMP_TRY(prof_collect(ctx));
    for (int c = 0; c < n; ++c) {
        ms_sum[c] = ctx->prof_ms_cat[c];
        count[c] = ctx->prof_cnt_cat[c];
    }
    return ESIMP_OK;
}

int esimp_bench_matvec(esimp_ctx* ctx, int reps, double* ms_avg, double* bytes_per_apply) {
    CHECK_CTX();
    NEED(ctx->density_set, "esimp_set_density first");
    if (reps < 1) return esimp_fail(ctx, ESIMP_ERR_ARG, "reps < 1");
    ESIMP_CUDA(cudaMemcpyAsync(ctx->d_p, ctx->d_f, sizeof(double) * ctx->ndof_loc, cudaMemcpyDeviceToDevice, ctx->stream));
    {
        const double m = ctx->diag_value;
        ctx->diag_value = 0.0;
        int s = vec_fixup_dirichlet(ctx, ctx->d_p, ctx->d_p);  // operator input must vanish on prescribed dofs
        ctx->diag_value = m;
        ESIMP_TRY(s);
    }
    ESIMP_TRY(op_apply(ctx, ctx->d_p, ctx->d_q));  // warm-up
    cudaEventRecord(ctx->ev[0], ctx->stream);
    for (int i = 0; i < reps; ++i) ESIMP_TRY(op_apply(ctx, ctx->d_p, ctx->d_q));
    cudaEventRecord(ctx->ev[1], ctx->stream);
    ESIMP_CUDA(cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
    if (ms_avg) *ms_avg = ms / reps;
    if (bytes_per_apply)
        *bytes_per_apply = algorithmic_bytes_per_apply(ctx);
    return ESIMP_OK;
}

int esimp_bench_operator_mode(esimp_ctx* ctx, int mode, int reps, double* ms_avg, double* bytes_per_apply) {
    CHECK_CTX();
    NEED(ctx->structured && ctx->density_set, "structured grid and esimp_set_density first");
    if (reps < 1 || mode < 0 || mode > 4) return esimp_fail(ctx, ESIMP_ERR_ARG, "bad mode / reps");
    ESIMP_CUDA(cudaMemcpyAsync(ctx->d_p, ctx->d_f, sizeof(double) * ctx->ndof_loc, cudaMemcpyDeviceToDevice, ctx->stream));
    {
        const double m = ctx->diag_value;
        ctx->diag_value = 0.0;
        int s = vec_fixup_dirichlet(ctx, ctx->d_p, ctx->d_p);  // operator input must vanish on prescribed dofs
        ctx->diag_value = m;
        ESIMP_TRY(s);
    }
    ESIMP_CUDA(cudaMemcpyAsync(ctx->d_r, ctx->d_p, sizeof(double) * ctx->ndof_loc, cudaMemcpyDeviceToDevice, ctx->stream));
    ESIMP_CUDA(cudaMemsetAsync(ctx->d_u, 0, sizeof(double) * ctx->ndof_loc, ctx->stream));
    const OpView vw = fine_view(ctx);
    double* d_sinv = nullptr;  // the smoother modes of the sweep kernel read D^-1 as one double per node
    if ((mode == 2 || mode == 4) && hex8_sweep_eligible(ctx, vw)) {
        ESIMP_CUDA(cudaMalloc(&d_sinv, sizeof(double) * ctx->g.NPX * ctx->g.plane));
        ESIMP_CUDA(cudaMemsetAsync(d_sinv, 0, sizeof(double) * ctx->g.NPX * ctx->g.plane, ctx->stream));
        int s = hex8_node_scalar_dinv(ctx, vw, ctx->d_dinv, d_sinv);
        if (s < 0) {
            cudaFree(d_sinv);
            return s;
        }
    }
    struct FreeOnExit {
        double* p;
        cudaStream_t st;
        ~FreeOnExit() {
            if (p) {
                cudaStreamSynchronize(st);
                cudaFree(p);
            }
        }
    } free_sinv{d_sinv, ctx->stream};
    auto run = [&]() -> int {
        switch (mode) {
            case 0: return hex8_matvec_view(ctx, vw, ctx->d_p, ctx->d_q, ctx->prm.matvec_variant, -1);
            case 1: return hex8_matvec_view(ctx, vw, ctx->d_p, ctx->d_q, ctx->prm.matvec_variant, SC_PQ);
            case 2: return hex8_smooth_view(ctx, vw, ctx->d_p, ctx->d_q, ctx->d_r, ctx->d_dinv, ctx->d_u, 0.3, 0.7, -1, 1, 1, nullptr, d_sinv);
            case 3: return hex8_resid_view(ctx, vw, ctx->d_p, ctx->d_q, ctx->d_r);
            default: return hex8_smooth_view(ctx, vw, ctx->d_p, ctx->d_q, ctx->d_r, ctx->d_dinv, ctx->d_u, 0.3, 0.7, SC_PQ, 1, 1, nullptr, d_sinv);
        }
    };
    ESIMP_TRY(run());  // warm-up
    cudaEventRecord(ctx->ev[0], ctx->stream);
    for (int i = 0; i < reps; ++i) ESIMP_TRY(run());
    cudaEventRecord(ctx->ev[1], ctx->stream);
    ESIMP_CUDA(cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
    if (ms_avg) *ms_avg = ms / reps;
    if (bytes_per_apply) {
        const double dofs = 3.0 * (ctx->own_p1 - ctx->own_p0) * (double)(ctx->g.ny + 1) * (double)(ctx->g.nz + 1);
        // smoother modes: z, r, d read, z', d written (40 B/dof) + D^-1 (8 B per node = 2.67 B/dof on the sweep kernel, 8 B/dof on the round-1 kernel)
        const double per_dof = mode <= 1 ? 16.0 : (mode == 3 ? 24.0 : (d_sinv ? 40.0 + 8.0 / 3.0 : 48.0));
        *bytes_per_apply = per_dof * dofs + 8.0 * (double)(ctx->own_cell1 - ctx->own_cell0);
    }
    ESIMP_CUDA(cudaMemsetAsync(ctx->d_u, 0, sizeof(double) * ctx->ndof_loc, ctx->stream));
    return ESIMP_OK;
}

}  // extern "C"
