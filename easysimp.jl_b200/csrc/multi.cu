// multi.cu -- one host process, N GPUs: the handle a single Julia session drives ("multi-GPU is internal, invisible
// to the caller").  The reference is one process (Optimization.jl:178); a caller that cannot be launched once per GPU creates an
// esimp_multi instead of an esimp_ctx and makes the same calls with the same GLOBAL arrays.  Inside: one esimp_ctx per device (the
// very contexts the one-process-per-GPU path uses, x-slab partition, NCCL communicator initialised from N threads of this process),
// and every call fans out to one host thread per GPU, because the per-rank entry points contain collectives and stream
// synchronisations that all ranks have to reach together.
#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "ctx.cuh"

struct esimp_multi {
    std::vector<esimp_ctx*> ctx;
    std::vector<int> dev;
    std::string err;
};

static std::string g_multi_create_error;

// run fn(rank) on every context concurrently; the most severe status wins (errors < 0 before warnings > 0)
template <typename F>
static int run_all(esimp_multi* m, F fn) {
    const int n = (int)m->ctx.size();
    std::vector<int> st(n, 0);
    std::vector<std::thread> th;
    th.reserve(n);
    for (int r = 0; r < n; ++r)
        th.emplace_back([&, r] {
            cudaSetDevice(m->dev[r]);
            st[r] = fn(r);
        });
    for (auto& t : th) t.join();
    int worst = 0;
    for (int r = 0; r < n; ++r) {
        if (st[r] < 0 && (worst >= 0 || st[r] < worst)) {
            worst = st[r];
            const char* e = esimp_last_error(m->ctx[r]);
            m->err = "rank " + std::to_string(r) + " (device " + std::to_string(m->dev[r]) + "): " + (e ? e : "");
        } else if (st[r] > 0 && worst >= 0) {
            worst |= st[r];
        }
    }
    return worst;
}

extern "C" {

int esimp_multi_create(esimp_multi** out, const int* devices, int n_devices) {
    if (!out || n_devices < 1) return ESIMP_ERR_ARG;
    *out = nullptr;
    esimp_multi* m = new esimp_multi();
    auto fail = [&](int code, const std::string& msg) {
        g_multi_create_error = msg;
        for (esimp_ctx* c : m->ctx) esimp_destroy(c);
        delete m;
        return code;
    };
    for (int r = 0; r < n_devices; ++r) {
        const int d = devices ? devices[r] : r;
        esimp_ctx* c = nullptr;
        int s = esimp_create(&c, d);
        if (s < 0) {
            const char* e = esimp_last_error(nullptr);
            return fail(s, std::string("esimp_create on device ") + std::to_string(d) + ": " + (e ? e : ""));
        }
        m->ctx.push_back(c);
        m->dev.push_back(d);
    }
    if (n_devices > 1) {
        unsigned char id[128];
        int s = esimp_comm_unique_id(id);
        if (s < 0) return fail(s, "ncclGetUniqueId failed (libnccl.so.2 not loadable?)");
        s = run_all(m, [&](int r) { return esimp_comm_init(m->ctx[r], r, n_devices, id); });  // ncclCommInitRank blocks until all ranks joined
        if (s < 0) return fail(s, "esimp_comm_init: " + m->err);
    }
    *out = m;
    return ESIMP_OK;
}

int esimp_multi_destroy(esimp_multi* m) {
    if (!m) return ESIMP_OK;
    run_all(m, [&](int r) { return esimp_destroy(m->ctx[r]); });
    delete m;
    return ESIMP_OK;
}

const char* esimp_multi_last_error(const esimp_multi* m) { return m ? m->err.c_str() : g_multi_create_error.c_str(); }
int esimp_multi_n_gpus(const esimp_multi* m) { return m ? (int)m->ctx.size() : 0; }
esimp_ctx* esimp_multi_rank_ctx(esimp_multi* m, int rank) { return (m && rank >= 0 && rank < (int)m->ctx.size()) ? m->ctx[rank] : nullptr; }
int64_t esimp_multi_n_dofs(const esimp_multi* m) { return m && !m->ctx.empty() ? esimp_n_dofs(m->ctx[0]) : 0; }
int64_t esimp_multi_n_cells(const esimp_multi* m) { return m && !m->ctx.empty() ? esimp_n_cells(m->ctx[0]) : 0; }

#define MULTI_CHECK() \
    if (!m || m->ctx.empty()) return ESIMP_ERR_ARG

int esimp_multi_set_params(esimp_multi* m, const esimp_params* p) {
    MULTI_CHECK();
    return run_all(m, [&](int r) { return esimp_set_params(m->ctx[r], p); });
}
int esimp_multi_set_grid_structured(esimp_multi* m, int nx, int ny, int nz, double hx, double hy, double hz, const int64_t* dof_of_node) {
    MULTI_CHECK();
    return run_all(m, [&](int r) { return esimp_set_grid_structured(m->ctx[r], nx, ny, nz, hx, hy, hz, dof_of_node); });
}
int esimp_multi_set_dirichlet(esimp_multi* m, const int64_t* dofs, int64_t n) {
    MULTI_CHECK();
    return run_all(m, [&](int r) { return esimp_set_dirichlet(m->ctx[r], dofs, n); });
}
int esimp_multi_set_load(esimp_multi* m, const double* f) {
    MULTI_CHECK();
    return run_all(m, [&](int r) { return esimp_set_load(m->ctx[r], f); });
}
int esimp_multi_set_body_force(esimp_multi* m, const double* accel3, double base_density) {
    MULTI_CHECK();
    return run_all(m, [&](int r) { return esimp_set_body_force(m->ctx[r], accel3, base_density); });
}
int esimp_multi_filter_build_structured(esimp_multi* m) {
    MULTI_CHECK();
    return run_all(m, [&](int r) { return esimp_filter_build_structured(m->ctx[r]); });
}
int esimp_multi_init_design(esimp_multi* m, const double* rho) {
    MULTI_CHECK();
    return run_all(m, [&](int r) { return esimp_init_design(m->ctx[r], rho); });
}
int esimp_multi_set_design(esimp_multi* m, const double* rho) {
    MULTI_CHECK();
    return run_all(m, [&](int r) { return esimp_set_design(m->ctx[r], rho); });  // every rank reads its own slab of the one array
}
// the scalars are allreduced inside the library: identical on every rank; the phase timers are the slowest rank's
int esimp_multi_iterate(esimp_multi* m, esimp_iter_stats* out) {
    MULTI_CHECK();
    std::vector<esimp_iter_stats> st(m->ctx.size());
    const int s = run_all(m, [&](int r) { return esimp_iterate(m->ctx[r], &st[r]); });
    if (out) {
        *out = st[0];
        for (const auto& q : st) {
            out->ms_filter = std::max(out->ms_filter, q.ms_filter);
            out->ms_solve = std::max(out->ms_solve, q.ms_solve);
            out->ms_sens = std::max(out->ms_sens, q.ms_sens);
            out->ms_oc = std::max(out->ms_oc, q.ms_oc);
            out->ms_total = std::max(out->ms_total, q.ms_total);
            out->status |= q.status;
        }
    }
    return s;
}
int esimp_multi_final_analysis(esimp_multi* m, double* energy, double* volume, int* pcg_iters) {
    MULTI_CHECK();
    const int n = (int)m->ctx.size();
    std::vector<double> e(n), v(n);
    std::vector<int> it(n);
    const int s = run_all(m, [&](int r) { return esimp_final_analysis(m->ctx[r], &e[r], &v[r], &it[r]); });
    if (energy) *energy = e[0];
    if (volume) *volume = v[0];
    if (pcg_iters) *pcg_iters = it[0];
    return s;
}
// design variables: every rank writes the cells of its own slab into the one array (disjoint runs)
int esimp_multi_get_design(esimp_multi* m, double* rho) {
    MULTI_CHECK();
    return run_all(m, [&](int r) { return esimp_get_design(m->ctx[r], rho); });
}
// the other getters fill zeros outside a rank's slab: gather rank by rank through one scratch array and add
static int gather_sum(esimp_multi* m, double* out, int64_t n, int (*get)(esimp_ctx*, double*)) {
    cudaSetDevice(m->dev[0]);
    int s = get(m->ctx[0], out);
    if (s < 0 || m->ctx.size() == 1) {
        if (s < 0) m->err = std::string("rank 0: ") + esimp_last_error(m->ctx[0]);
        return s;
    }
    std::vector<double> tmp((size_t)n);
    for (size_t r = 1; r < m->ctx.size(); ++r) {
        cudaSetDevice(m->dev[r]);
        s = get(m->ctx[r], tmp.data());
        if (s < 0) {
            m->err = "rank " + std::to_string(r) + ": " + esimp_last_error(m->ctx[r]);
            return s;
        }
        for (int64_t i = 0; i < n; ++i) out[i] += tmp[(size_t)i];
    }
    return ESIMP_OK;
}
int esimp_multi_get_physical_density(esimp_multi* m, double* rho) {
    MULTI_CHECK();
    if (!rho) return ESIMP_ERR_ARG;
    return gather_sum(m, rho, esimp_n_cells(m->ctx[0]), esimp_get_physical_density);
}
int esimp_multi_get_displacement(esimp_multi* m, double* u) {
    MULTI_CHECK();
    if (!u) return ESIMP_ERR_ARG;
    return gather_sum(m, u, esimp_n_dofs(m->ctx[0]), esimp_get_displacement);
}

}  // extern "C"
