// extern "C" entry points of libfiksi_b200.so (declared in include/fiksi_b200.h).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <atomic>
#include <cstring>
#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/fiksi_b200.h"
#include "lm_kernels.cuh"
#include "lm_sketch.cuh"
#include "multifrontal.cuh"
#include "single_pass.hpp"
#include "sparse_path.cuh"
#include "symbolic.hpp"

namespace {

thread_local std::string g_error;

int fail(int code, const std::string& msg) {
    g_error = msg;
    return code;
}
int cuda_fail(cudaError_t e, const char* what) {
    g_error = std::string(what) + ": " + cudaGetErrorString(e);
    return e == cudaErrorMemoryAllocation ? FK_ERR_OOM : FK_ERR_CUDA;
}
#define CU(call)                                           \
    do {                                                   \
        cudaError_t e_ = (call);                           \
        if (e_ != cudaSuccess) return cuda_fail(e_, #call); \
    } while (0)

int usable_devices() {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

// Packs the topology tables into one device allocation and fills the DevProgram view.
struct DeviceProgram {
    int device = -1;
    void* buf = nullptr;
    fk::DevProgram view{};
    const uint32_t *jcolptr = nullptr, *jrow = nullptr;  // CSC of the Jacobian (L-BFGS)
    std::unique_ptr<fk::SkProgram> sketch;               // parameter block of the sketch-per-thread LM kernel
    ~DeviceProgram() {
        if (buf) {
            cudaSetDevice(device);
            cudaFree(buf);
        }
    }
};

template <class T>
size_t reserve(size_t& off, size_t count) {
    off = (off + 15) & ~size_t(15);
    size_t at = off;
    off += std::max<size_t>(count, 1) * sizeof(T);
    return at;
}

int upload_program(const fk::Topology& t, int device, DeviceProgram& out) {
    CU(cudaSetDevice(device));
    struct Item { const void* src; size_t bytes; size_t at; };
    std::vector<Item> items;
    size_t off = 0;
    auto add = [&](const auto& vec) {
        using T = typename std::decay<decltype(vec)>::type::value_type;
        size_t at = reserve<T>(off, vec.size());
        items.push_back({vec.data(), vec.size() * sizeof(T), at});
        return at;
    };
    const fk::Topology::Tables& tb = t.tab;
    std::vector<uint32_t> jcolptr(t.n_free + 1), jrow(t.jac_nnz);  // CSC of the Jacobian without damping rows (L-BFGS gradient)
    for (uint32_t c = 0; c <= t.n_free; c++) jcolptr[c] = t.aug_colptr[c] - c;
    for (uint32_t c = 0; c < t.n_free; c++)
        for (uint32_t q = t.aug_colptr[c]; q + 1 < t.aug_colptr[c + 1]; q++) jrow[q - c] = t.aug_rowidx[q];
    const size_t o_jcp = add(jcolptr), o_jrow = add(jrow);
    size_t o_free = add(t.free_vars), o_rh = add(tb.row_hdr), o_rs = add(tb.row_slots),
           o_af = add(tb.a_flags), o_ao = add(tb.a_ops), o_ad = add(tb.a_dst),
           o_gf = add(tb.g_flags), o_go = add(tb.g_ops), o_gd = add(tb.g_dst),
           o_fs = add(tb.f_steps), o_fo = add(tb.f_ops), o_ss = add(tb.s_steps), o_so = add(tb.s_ops),
           o_bs = add(tb.b_steps), o_bo = add(tb.b_ops);
    const size_t o_sk = (t.sk.ok && fk::sk_fits(t.sk.entries, t.sk.tab.size())) ? add(t.sk.tab) : SIZE_MAX;
    std::vector<unsigned char> host(off + 16, 0);
    for (const Item& it : items)
        if (it.bytes) std::memcpy(host.data() + it.at, it.src, it.bytes);
    CU(cudaMalloc(&out.buf, host.size()));
    out.device = device;
    CU(cudaMemcpy(out.buf, host.data(), host.size(), cudaMemcpyHostToDevice));
    unsigned char* b = (unsigned char*)out.buf;
    fk::DevProgram& v = out.view;
    v.n_vars = t.n_vars; v.n_expr = t.n_expr; v.n = t.n_free; v.m = t.n_rows;
    v.jnnz = t.jac_nnz; v.lnnz = (uint32_t)t.l_rowidx.size(); v.tile = t.tile; v.uniform_kind = tb.uniform_kind;
    v.eval_rounds = tb.eval_rounds; v.a_nsteps = tb.a_nsteps; v.g_nsteps = tb.g_nsteps;
    v.f_nsteps = tb.f_nsteps; v.s_nsteps = tb.s_nsteps; v.b_nsteps = tb.b_nsteps;
    v.free_vars = (const uint32_t*)(b + o_free);
    v.row_hdr = (const uint32_t*)(b + o_rh); v.row_slots = (const uint2*)(b + o_rs);
    v.a_flags = (const uint32_t*)(b + o_af); v.a_ops = (const uint32_t*)(b + o_ao); v.a_dst = (const uint32_t*)(b + o_ad);
    v.g_flags = (const uint32_t*)(b + o_gf); v.g_ops = (const uint32_t*)(b + o_go); v.g_dst = (const uint32_t*)(b + o_gd);
    v.f_steps = (const uint32_t*)(b + o_fs); v.f_ops = (const uint2*)(b + o_fo);
    v.s_steps = (const uint2*)(b + o_ss); v.s_ops = (const uint32_t*)(b + o_so);
    v.b_steps = (const uint2*)(b + o_bs); v.b_ops = (const uint32_t*)(b + o_bo);
    out.jcolptr = (const uint32_t*)(b + o_jcp); out.jrow = (const uint32_t*)(b + o_jrow);
    v.sketch_prog = nullptr;
    if (o_sk != SIZE_MAX) {
        const fk::Topology::SketchTables& k = t.sk;
        out.sketch.reset(new fk::SkProgram());
        fk::SkProgram& p = *out.sketch;
        std::memset(&p, 0, sizeof(p));
        p.n_vars = t.n_vars; p.n_expr = t.n_expr; p.n = t.n_free; p.m = t.n_rows; p.lnnz = (uint32_t)t.l_rowidx.size(); p.entries = k.entries;
        p.xa = k.xa; p.xb = k.xb; p.w = k.w; p.f = k.f; p.fx = k.fx; p.pr = k.pr; p.nfix = k.nfix; p.npar = k.npar;
        p.off_free = k.off_free; p.off_fix = k.off_fix; p.off_par = k.off_par;
        p.off_eval = k.off_eval; p.off_factor = k.off_factor; p.off_back = k.off_back;
        p.tab_words = (uint32_t)k.tab.size();
        p.tab = (const uint32_t*)(b + o_sk);
        v.sketch_prog = out.sketch.get();
    }
    return FK_OK;
}

}  // namespace

struct fk_batch_plan;
// Per-device staging pipeline of fk_batch_solve: kStreams plans + streams, reused across calls.
struct DevicePipeline {
// (eight chunks in flight: a chunk's kernel lasts one LM solve however small the chunk is, so with three the copies of
// small chunks could not keep the device busy; bench truss end to end 56.2 -> 59.0 M sketches/s with 16 chunks.  Swept on
// B200 with tools/e2e_sweep.py: 4 / 6 / 8 / 16 / 32 streams x 8..32 chunks all end between 1.10 and 1.19 ms per call of
// 65,536 trusses against 0.905 ms for the kernel alone)
#ifndef FK_PIPELINE_STREAMS
#define FK_PIPELINE_STREAMS 8
#endif
    static constexpr uint32_t kStreams = FK_PIPELINE_STREAMS;
    std::mutex mu;
    int device = -1;
    uint32_t chunk = 0;
    fk_batch_plan* plans[kStreams] = {};
    cudaStream_t streams[kStreams] = {};
    // small requests (one sketch, a handful): one pinned staging block, one device block, one copy each way
    static constexpr size_t kSmallBytes = 64 * 1024;
    unsigned char* small_h = nullptr;
    unsigned char* small_d = nullptr;
    cudaStream_t small_stream = nullptr;
    cudaEvent_t shared_param_ev = nullptr;  // fk_batch_system_solve with one parameter row for all sketches: its upload, once per call
    // fk_batch_system_solve_begin / _wait: up to two calls in flight.  A call's token picks (by parity) the events recorded behind its
    // last chunk on every stream and its copy of the shared parameter row.
    uint64_t next_token = 1;
    cudaEvent_t done_ev[2][kStreams] = {};
    double* d_shared_param[2] = {nullptr, nullptr};
    uint32_t shared_param_cap = 0;
    void release();
    ~DevicePipeline() { release(); }
};

// Device tables and grow-only device buffers of fk_batch_analyze for one (topology, device).
struct AnalyzePool {
    int device = -1;
    std::mutex mu;
    uint8_t* d_kind = nullptr;
    uint32_t* d_slot = nullptr;
    double *d_vars = nullptr, *d_param = nullptr;
    uint8_t* d_out = nullptr;
    uint32_t capacity = 0;
    cudaStream_t stream = nullptr;
    ~AnalyzePool() {
        if (device >= 0) cudaSetDevice(device);
        cudaFree(d_kind); cudaFree(d_slot); cudaFree(d_vars); cudaFree(d_param); cudaFree(d_out);
        if (stream) cudaStreamDestroy(stream);
    }
};

// Device tables of assemble::solve's pre-processing for one (topology, device): expression kinds, the variables that
// draw from the solve's generator (ascending) and the draws themselves (fiksi/src/rand.rs:24-39, two per variable).
struct PrepareTables {
    int device = -1;
    std::vector<uint32_t> perturb_vars;
    uint32_t seed = 0;
    uint8_t* d_kind = nullptr;
    uint32_t* d_perturb = nullptr;
    double* d_draws = nullptr;
    uint32_t* d_free_draw = nullptr;  // sketch kernel's raw mode: draw index per free column / per fixed variable the rows read
    uint32_t* d_fix_draw = nullptr;
    ~PrepareTables() {
        if (device >= 0) cudaSetDevice(device);
        cudaFree(d_kind); cudaFree(d_perturb); cudaFree(d_draws); cudaFree(d_free_draw); cudaFree(d_fix_draw);
    }
};

struct fk_topology {
    fk::Topology t;
    std::mutex mu;
    std::map<int, std::unique_ptr<DeviceProgram>> programs;
    std::map<int, std::unique_ptr<DevicePipeline>> pipelines;
    std::map<int, std::unique_ptr<fk::SparseSolver>> sparse;  // path 2: one solver per device
    std::mutex sparse_mu;               // a SparseSolver owns one set of work vectors and CUDA graphs: one solve at a time
    std::map<int, std::unique_ptr<struct AnalyzePool>> analyze_pools;  // fk_batch_analyze: tables + grow-only buffers
    std::map<int, std::unique_ptr<struct PrepareTables>> prepare_tables;  // fk_batch_system_solve: kinds, perturbation list, draws
    std::unique_ptr<fk_topology> latency_twin;  // same topology with 32 lanes per sketch: single-system solves
    bool twin_tried = false;
    std::mutex sp_mu;                   // SinglePass plan: one sub-topology per strongly connected set
    bool sp_planned = false;
    std::vector<fk_topology*> sp_subs;
    ~fk_topology();

    int sparse_for(int device, fk::SparseSolver** out, std::string* err) {
        std::lock_guard<std::mutex> lock(mu);
        auto& p = sparse[device];
        if (!p) {
            std::unique_ptr<fk::SparseSolver> s(new fk::SparseSolver());
            int rc = s->init(t, device, err);
            if (rc != FK_OK) return rc;
            p = std::move(s);
        }
        *out = p.get();
        return FK_OK;
    }

    // This topology with op tables for 32 lanes per system (single-system solves and heterogeneous batches): itself when
    // its batch-oriented lane count is already 32, else a twin that shares the symbolic analysis and re-packs the tables.
    fk_topology* lanes32() {
        if (t.path != 0 || t.tile == 32) return this;
        std::lock_guard<std::mutex> lock(mu);
        if (!twin_tried) {
            twin_tried = true;
            static const bool off = std::getenv("FK_NO_LATENCY_TWIN") != nullptr;
            if (!off) {
                std::unique_ptr<fk_topology> tw(new (std::nothrow) fk_topology());
                if (tw) {
                    tw->t = t;
                    tw->t.tile = 32;
                    if (tw->t.build_tables() == FK_OK) latency_twin = std::move(tw);
                }
            }
        }
        return latency_twin ? latency_twin.get() : this;
    }

    DevicePipeline* pipeline_for(int device) {
        std::lock_guard<std::mutex> lock(mu);
        auto& p = pipelines[device];
        if (!p) {
            p.reset(new DevicePipeline());
            p->device = device;
        }
        return p.get();
    }

    int program_for(int device, const fk::DevProgram** out, const DeviceProgram** full = nullptr) {
        std::lock_guard<std::mutex> lock(mu);
        auto it = programs.find(device);
        if (it == programs.end()) {
            std::unique_ptr<DeviceProgram> p(new DeviceProgram());
            int rc = upload_program(t, device, *p);
            if (rc != FK_OK) return rc;
            it = programs.emplace(device, std::move(p)).first;
        }
        *out = &it->second->view;
        if (full) *full = it->second.get();
        return FK_OK;
    }
};

struct fk_batch_plan {
    fk_topology* topo = nullptr;
    int device = 0;
    uint32_t capacity = 0, n = 0;
    const fk::DevProgram* prog = nullptr;
    const DeviceProgram* full = nullptr;
    double *d_vars = nullptr, *d_params = nullptr, *d_out = nullptr, *d_er = nullptr, *d_ej = nullptr;
    fk_report* d_rep = nullptr;
    double *d_raw_vars = nullptr, *d_raw_param = nullptr, *d_scales = nullptr;  // fk_batch_system_solve: unscaled input, scale per sketch
    uint64_t launches = 0;
    ~fk_batch_plan() {
        cudaSetDevice(device);
        cudaFree(d_vars); cudaFree(d_params); cudaFree(d_out); cudaFree(d_er); cudaFree(d_ej); cudaFree(d_rep);
        cudaFree(d_raw_vars); cudaFree(d_raw_param); cudaFree(d_scales);
    }
};

fk_topology::~fk_topology() {
    for (fk_topology* s : sp_subs) delete s;
}

void DevicePipeline::release() {
    if (device >= 0) cudaSetDevice(device);
    for (uint32_t s = 0; s < kStreams; s++) {
        if (streams[s]) cudaStreamDestroy(streams[s]);
        streams[s] = nullptr;
        delete plans[s];
        plans[s] = nullptr;
    }
    chunk = 0;
    if (small_h) cudaFreeHost(small_h);
    if (small_d) cudaFree(small_d);
    if (small_stream) cudaStreamDestroy(small_stream);
    small_h = small_d = nullptr;
    small_stream = nullptr;
    if (shared_param_ev) cudaEventDestroy(shared_param_ev);
    shared_param_ev = nullptr;
    for (int a = 0; a < 2; a++) {
        for (uint32_t k = 0; k < kStreams; k++) {
            if (done_ev[a][k]) cudaEventDestroy(done_ev[a][k]);
            done_ev[a][k] = nullptr;
        }
        if (d_shared_param[a]) cudaFree(d_shared_param[a]);
        d_shared_param[a] = nullptr;
    }
    shared_param_cap = 0;
}

extern "C" {

int fk_version(void) { return 100; }
const char* fk_last_error(void) { return g_error.c_str(); }
int fk_device_count(void) { return usable_devices(); }

int fk_topology_create(const fk_problem* problem, fk_topology** out) {
    if (!problem || !out) return fail(FK_ERR_INVALID, "null argument");
    std::unique_ptr<fk_topology> t(new (std::nothrow) fk_topology());
    if (!t) return fail(FK_ERR_OOM, "host allocation failed");
    int rc;
    try {
        rc = t->t.build(*problem);
    } catch (const std::bad_alloc&) {
        return fail(FK_ERR_OOM, "host allocation failed in symbolic analysis");
    }
    if (rc != FK_OK) return fail(rc, t->t.error);
    *out = t.release();
    return FK_OK;
}

void fk_topology_destroy(fk_topology* topo) { delete topo; }

int fk_topology_info_get(const fk_topology* topo, fk_topology_info* info) {
    if (!topo || !info) return fail(FK_ERR_INVALID, "null argument");
    topo->t.fill_info(info);
    return FK_OK;
}

int fk_topology_symbolic(const fk_topology* topo, uint32_t* aug_colptr, uint32_t* aug_rowidx, int32_t* colamd_perm,
                         int32_t* etree_parent, uint32_t* r_colptr, uint32_t* r_rowidx) {
    if (!topo) return fail(FK_ERR_INVALID, "null topology");
    const fk::Topology& t = topo->t;
    auto copy = [](auto* dst, const auto& v) {
        if (dst && !v.empty()) std::memcpy(dst, v.data(), v.size() * sizeof(v[0]));
    };
    copy(aug_colptr, t.aug_colptr); copy(aug_rowidx, t.aug_rowidx); copy(colamd_perm, t.perm);
    copy(etree_parent, t.parent); copy(r_colptr, t.r_colptr); copy(r_rowidx, t.r_rowidx);
    return FK_OK;
}

int fk_symbolic(const fk_problem* problem, uint32_t* aug_colptr, uint32_t* aug_rowidx, int32_t* colamd_perm,
                int32_t* etree_parent, uint32_t* r_colptr, uint32_t* r_rowidx) {
    fk_topology* t = nullptr;
    int rc = fk_topology_create(problem, &t);
    if (rc != FK_OK) return rc;
    rc = fk_topology_symbolic(t, aug_colptr, aug_rowidx, colamd_perm, etree_parent, r_colptr, r_rowidx);
    fk_topology_destroy(t);
    return rc;
}

int fk_topology_supernodal(const fk_topology* topo, fk_supernodal_info* info, uint32_t* sn_first, uint32_t* front,
                           int32_t* sn_parent, uint32_t* rows, uint32_t* rel, uint8_t* big, uint32_t* level, uint32_t* tasks,
                           uint32_t* launches) {
    if (!topo || !info) return fail(FK_ERR_INVALID, "null argument");
    fk::Multifrontal mf;
    std::string err;
    try {
        if (mf.build_symbolic(topo->t, &err) != cudaSuccess) return fail(FK_ERR_TOO_LARGE, err);
    } catch (const std::bad_alloc&) {
        return fail(FK_ERR_OOM, "host allocation failed in the supernodal analysis");
    }
    const fk::MfSymbolic& y = mf.sym;
    const auto seq = mf.factor_launches();
    info->n_supernodes = y.S; info->n_small_subtrees = y.nsub; info->n_big = y.nbig; info->n_levels = y.nlevels;
    info->max_front = y.max_front; info->n_tasks = (uint32_t)(y.tasks.size() / 4); info->n_launches = (uint32_t)seq.size(); info->pad0 = 0;
    info->rows_total = y.rows.size(); info->rel_total = y.rel.size(); info->panel_doubles = y.pan_total; info->update_doubles = y.upd_total;
    auto copy = [](auto* dst, const auto& v) {
        if (dst && !v.empty()) std::memcpy(dst, v.data(), v.size() * sizeof(v[0]));
    };
    if (sn_first) {
        copy(sn_first, y.c0);
        sn_first[y.S] = y.n;
    }
    copy(front, y.f); copy(sn_parent, y.sparent); copy(rows, y.rows); copy(rel, y.rel); copy(big, y.big); copy(level, y.level);
    copy(tasks, y.tasks);
    if (launches)
        for (size_t k = 0; k < seq.size(); k++) {
            launches[3 * k] = (uint32_t)seq[k].kind; launches[3 * k + 1] = seq[k].first; launches[3 * k + 2] = seq[k].count;
        }
    return FK_OK;
}

// ---- Decomposer::SinglePass on a uniform batch ---------------------------------------------------------
// The plan (one sub-topology per strongly connected set) is built once per topology and cached.  Variables
// and parameters of a shard stay resident on its device for the whole pass: per set one batched LM launch over
// all sketches and one scatter of the solved values back into the variable rows (assemble/mod.rs:201-208).
static int single_pass_plan_for(fk_topology* topo, const std::vector<fk_topology*>** out) {
    std::lock_guard<std::mutex> lock(topo->sp_mu);
    if (!topo->sp_planned) {
        const fk::Topology& t = topo->t;
        std::vector<std::vector<uint32_t>> var_exprs(t.n_vars), expr_vars(t.n_expr);
        for (uint32_t e = 0; e < t.n_expr; e++) {
            uint32_t sv[8];
            const int a = fk::expand_slots(t.kind[e], &t.idx[4 * (size_t)e], sv);
            expr_vars[e].assign(sv, sv + a);
            for (int k = 0; k < a; k++) var_exprs[sv[k]].push_back(e);
        }
        fk::SinglePassPlanner planner(var_exprs, expr_vars);
        const std::vector<fk::SinglePassStep> plan = planner.plan(t.free_vars);
        std::vector<fk_topology*> subs;
        for (const fk::SinglePassStep& step : plan) {
            fk_problem p{};
            p.n_vars = t.n_vars; p.n_expr = t.n_expr; p.kind = t.kind.data(); p.idx = t.idx.data();
            p.n_free = (uint32_t)step.free_variables.size(); p.free_vars = step.free_variables.data();
            p.n_rows = (uint32_t)step.expressions.size(); p.rows = step.expressions.data();
            fk_topology* sub = nullptr;
            int rc = fk_topology_create(&p, &sub);
            if (rc == FK_OK && sub->t.path == 2) {
                fk_topology_destroy(sub);
                rc = fail(FK_ERR_TOO_LARGE, "a strongly connected set needs the global sparse path; use fk_system_solve_opts");
            }
            if (rc != FK_OK) {
                for (fk_topology* s : subs) fk_topology_destroy(s);
                return rc;
            }
            subs.push_back(sub);
        }
        topo->sp_subs = std::move(subs);
        topo->sp_planned = true;
    }
    *out = &topo->sp_subs;
    return FK_OK;
}

static int single_pass_device_range(fk_topology* topo, const std::vector<fk_topology*>& subs, int device, uint32_t lo, uint32_t hi,
                                    double* vars, const double* param, fk_report* reports, std::string* err) {
    const fk::Topology& t = topo->t;
    const size_t steps = subs.size();
    uint32_t max_free = 1;
    for (const fk_topology* s : subs) max_free = std::max(max_free, s->t.n_free);
    int rc = FK_OK;
    double *d_vars = nullptr, *d_params = nullptr, *d_out = nullptr;
    fk_report *d_rep = nullptr, *d_rep_t = nullptr;
    cudaStream_t stream = nullptr;
    auto done = [&](int code) {
        if (stream) cudaStreamDestroy(stream);
        cudaFree(d_vars); cudaFree(d_params); cudaFree(d_out); cudaFree(d_rep); cudaFree(d_rep_t);
        if (code != FK_OK && err) *err = g_error;
        return code;
    };
#define SP_CU(call)                                              \
    do {                                                         \
        cudaError_t e_ = (call);                                 \
        if (e_ != cudaSuccess) return done(cuda_fail(e_, #call)); \
    } while (0)
    SP_CU(cudaSetDevice(device));
    // bound the resident set of one chunk to ~8 GiB
    const size_t per_sketch = sizeof(double) * ((size_t)t.n_vars + t.n_expr + max_free) + 2 * sizeof(fk_report) * std::max<size_t>(1, steps);
    const uint32_t chunk = (uint32_t)std::min<uint64_t>(hi - lo, std::max<uint64_t>(1, (8ull << 30) / per_sketch));
    SP_CU(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    SP_CU(cudaMalloc(&d_vars, sizeof(double) * std::max<size_t>(1, (size_t)chunk * t.n_vars)));
    SP_CU(cudaMalloc(&d_params, sizeof(double) * std::max<size_t>(1, (size_t)chunk * t.n_expr)));
    SP_CU(cudaMalloc(&d_out, sizeof(double) * (size_t)chunk * max_free));
    SP_CU(cudaMalloc(&d_rep, sizeof(fk_report) * std::max<size_t>(1, (size_t)chunk * steps)));
    if (reports) SP_CU(cudaMalloc(&d_rep_t, sizeof(fk_report) * std::max<size_t>(1, (size_t)chunk * steps)));
    std::vector<const fk::DevProgram*> progs(steps, nullptr);
    for (size_t st = 0; st < steps; st++) {
        rc = subs[st]->program_for(device, &progs[st]);
        if (rc != FK_OK) return done(rc);
    }
    for (uint32_t at = lo; at < hi; at += chunk) {
        const uint32_t cnt = std::min(chunk, hi - at);
        SP_CU(cudaMemcpyAsync(d_vars, vars + (size_t)at * t.n_vars, sizeof(double) * (size_t)cnt * t.n_vars, cudaMemcpyHostToDevice, stream));
        if (t.n_expr) SP_CU(cudaMemcpyAsync(d_params, param + (size_t)at * t.n_expr, sizeof(double) * (size_t)cnt * t.n_expr, cudaMemcpyHostToDevice, stream));
        for (size_t st = 0; st < steps; st++) {
            int e = fk::launch_batch_lm(*progs[st], cnt, d_vars, d_params, d_out, d_rep + st * (size_t)cnt, stream);
            if (e == 0) e = fk::launch_scatter_free(progs[st]->free_vars, progs[st]->n, t.n_vars, cnt, d_out, d_vars, stream);
            if (e != 0) return done(cuda_fail((cudaError_t)e, "launch SinglePass set"));
        }
        SP_CU(cudaMemcpyAsync(vars + (size_t)at * t.n_vars, d_vars, sizeof(double) * (size_t)cnt * t.n_vars, cudaMemcpyDeviceToHost, stream));
        if (reports && steps) {
            const int e = fk::launch_transpose_reports(d_rep, d_rep_t, cnt, (uint32_t)steps, stream);
            if (e != 0) return done(cuda_fail((cudaError_t)e, "launch fk_transpose_reports_kernel"));
            SP_CU(cudaMemcpyAsync(reports + (size_t)at * steps, d_rep_t, sizeof(fk_report) * (size_t)cnt * steps, cudaMemcpyDeviceToHost, stream));
        }
        SP_CU(cudaStreamSynchronize(stream));
    }
#undef SP_CU
    return done(FK_OK);
}

int fk_batch_solve_single_pass(const fk_topology* topo_c, uint32_t n, double* vars, const double* param, fk_report* reports,
                               uint32_t* n_steps, int n_gpus) {
    fk_topology* topo = const_cast<fk_topology*>(topo_c);
    if (!topo || (n && (!vars || (!param && topo->t.n_expr)))) return fail(FK_ERR_INVALID, "null argument");
    const std::vector<fk_topology*>* subs = nullptr;
    int rc = single_pass_plan_for(topo, &subs);
    if (rc != FK_OK) return rc;
    if (n_steps) *n_steps = (uint32_t)subs->size();
    if (n == 0) return FK_OK;
    const int ndev = usable_devices();
    if (ndev == 0) return fail(FK_ERR_NO_DEVICE, "no CUDA device visible; fiksi_b200 has no CPU fallback");
    if (n_gpus <= 0 || n_gpus > ndev) n_gpus = ndev;
    if (n_gpus == 1) {
        int cur = 0;
        if (cudaGetDevice(&cur) != cudaSuccess) cur = 0;
        return single_pass_device_range(topo, *subs, cur, 0, n, vars, param, reports, nullptr);
    }
    std::vector<int> rcs(n_gpus, FK_OK);
    std::vector<std::string> errs(n_gpus);
    std::vector<std::thread> pool;
    for (int g = 0; g < n_gpus; g++) {
        const uint32_t lo = (uint32_t)((uint64_t)n * g / n_gpus), hi = (uint32_t)((uint64_t)n * (g + 1) / n_gpus);
        pool.emplace_back([=, &rcs, &errs]() {
            if (hi > lo) rcs[g] = single_pass_device_range(topo, *subs, g, lo, hi, vars, param, reports, &errs[g]);
        });
    }
    for (auto& th : pool) th.join();
    for (int g = 0; g < n_gpus; g++)
        if (rcs[g] != FK_OK) return fail(rcs[g], errs[g]);
    return FK_OK;
}

// ---- L-BFGS on a uniform batch -------------------------------------------------------------------
static int run_device_range(fk_topology* topo, int device, uint32_t lo, uint32_t hi, const double* vars, const double* param,
                            double* free_out, fk_report* reports, std::string* err, int optimizer = 0, uint64_t* token = nullptr);

int fk_batch_solve_lbfgs(const fk_topology* topo_c, int device, uint32_t n, const double* vars, const double* param, double* free_out,
                         fk_report* reports) {
    if (!topo_c) return fail(FK_ERR_INVALID, "null topology");
    if (n == 0) return FK_OK;
    if (!vars || !free_out || !reports || (!param && topo_c->t.n_expr)) return fail(FK_ERR_INVALID, "null buffer");
    if (topo_c->t.path != 0) return fail(FK_ERR_TOO_LARGE, "L-BFGS is built for the shared-memory tile paths only");
    const int ndev = usable_devices();
    if (ndev == 0) return fail(FK_ERR_NO_DEVICE, "no CUDA device visible; fiksi_b200 has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(FK_ERR_INVALID, "device index out of range");
    // same pooled plans / streams / pinned staging as the LM entry (no allocation per call)
    return run_device_range(const_cast<fk_topology*>(topo_c), device, 0, n, vars, param, free_out, reports, nullptr, 1);
}

// ---- System::analyze on a batch -----------------------------------------------------------------
int fk_batch_analyze(const fk_topology* topo_c, int device, uint32_t n, const double* vars, const double* param, uint8_t* independent) {
    fk_topology* topo = const_cast<fk_topology*>(topo_c);
    if (!topo || (n && (!vars || !independent || (!param && topo->t.n_expr)))) return fail(FK_ERR_INVALID, "null argument");
    const int ndev = usable_devices();
    if (ndev == 0) return fail(FK_ERR_NO_DEVICE, "no CUDA device visible; fiksi_b200 has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(FK_ERR_INVALID, "device index out of range");
    const fk::Topology& t = topo->t;
    if (n == 0 || t.n_expr == 0) return FK_OK;
    AnalyzePool* pool;
    {
        std::lock_guard<std::mutex> lock(topo->mu);
        auto& p = topo->analyze_pools[device];
        if (!p) p.reset(new AnalyzePool());
        pool = p.get();
    }
    std::lock_guard<std::mutex> lock(pool->mu);
    CU(cudaSetDevice(device));
    if (pool->device < 0) {  // tables: once per (topology, device)
        std::vector<uint32_t> slot_var((size_t)t.n_expr * 8, 0);
        for (uint32_t e = 0; e < t.n_expr; e++) {
            uint32_t sv[8];
            const int a = fk::expand_slots(t.kind[e], &t.idx[4 * (size_t)e], sv);
            for (int k = 0; k < a; k++) slot_var[(size_t)e * 8 + k] = sv[k];
        }
        pool->device = device;
        CU(cudaMalloc(&pool->d_kind, t.n_expr));
        CU(cudaMalloc(&pool->d_slot, slot_var.size() * sizeof(uint32_t)));
        CU(cudaStreamCreateWithFlags(&pool->stream, cudaStreamNonBlocking));
        CU(cudaMemcpyAsync(pool->d_kind, t.kind.data(), t.n_expr, cudaMemcpyHostToDevice, pool->stream));
        CU(cudaMemcpyAsync(pool->d_slot, slot_var.data(), slot_var.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, pool->stream));
        CU(cudaStreamSynchronize(pool->stream));  // slot_var leaves scope
    }
    if (pool->capacity < n) {  // grow-only
        cudaFree(pool->d_vars); cudaFree(pool->d_param); cudaFree(pool->d_out);
        pool->d_vars = pool->d_param = nullptr; pool->d_out = nullptr; pool->capacity = 0;
        const uint32_t cap = std::max<uint32_t>(n, 1024);
        CU(cudaMalloc(&pool->d_vars, sizeof(double) * (size_t)cap * std::max<uint32_t>(t.n_vars, 1)));
        CU(cudaMalloc(&pool->d_param, sizeof(double) * (size_t)cap * t.n_expr));
        CU(cudaMalloc(&pool->d_out, (size_t)cap * t.n_expr));
        pool->capacity = cap;
    }
    CU(cudaMemcpyAsync(pool->d_vars, vars, sizeof(double) * (size_t)n * t.n_vars, cudaMemcpyHostToDevice, pool->stream));
    CU(cudaMemcpyAsync(pool->d_param, param, sizeof(double) * (size_t)n * t.n_expr, cudaMemcpyHostToDevice, pool->stream));
    const int e = fk::launch_batch_analyze(t.n_vars, t.n_expr, pool->d_kind, pool->d_slot, n, pool->d_vars, pool->d_param, pool->d_out, pool->stream);
    if (e == (int)cudaErrorInvalidConfiguration)
        return fail(FK_ERR_TOO_LARGE, "the expression x variable matrix of one sketch does not fit shared memory");
    if (e != 0) return cuda_fail((cudaError_t)e, "launch fk_batch_analyze_kernel");
    CU(cudaMemcpyAsync(independent, pool->d_out, (size_t)n * t.n_expr, cudaMemcpyDeviceToHost, pool->stream));
    CU(cudaStreamSynchronize(pool->stream));
    return FK_OK;
}

// ---- device-resident batch plan -----------------------------------------------------------------
int fk_batch_plan_create(const fk_topology* topo_c, uint32_t capacity, int device, fk_batch_plan** out) {
    fk_topology* topo = const_cast<fk_topology*>(topo_c);
    if (!topo || !out || capacity == 0) return fail(FK_ERR_INVALID, "null topology / zero capacity");
    int ndev = usable_devices();
    if (ndev == 0) return fail(FK_ERR_NO_DEVICE, "no CUDA device visible; fiksi_b200 has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(FK_ERR_INVALID, "device index out of range");
    if (topo->t.path == 2)
        return fail(FK_ERR_TOO_LARGE, "topology needs the global sparse path; use fk_lm_solve");
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(FK_ERR_NO_DEVICE, "device is not sm_100 class; kernels are built for sm_100a only");
    std::unique_ptr<fk_batch_plan> p(new fk_batch_plan());
    p->topo = topo; p->device = device; p->capacity = capacity;
    int rc = topo->program_for(device, &p->prog, &p->full);
    if (rc != FK_OK) return rc;
    const fk::Topology& t = topo->t;
    CU(cudaMalloc(&p->d_vars, sizeof(double) * std::max<size_t>(1, (size_t)capacity * t.n_vars)));
    CU(cudaMalloc(&p->d_params, sizeof(double) * std::max<size_t>(1, (size_t)capacity * t.n_expr)));
    CU(cudaMalloc(&p->d_out, sizeof(double) * std::max<size_t>(1, (size_t)capacity * t.n_free)));
    CU(cudaMalloc(&p->d_rep, sizeof(fk_report) * (size_t)capacity));
    *out = p.release();
    return FK_OK;
}

void fk_batch_plan_destroy(fk_batch_plan* plan) { delete plan; }

int fk_batch_plan_upload(fk_batch_plan* plan, uint32_t n, const double* vars, const double* param, void* stream) {
    if (!plan || n > plan->capacity || (n && (!vars || (!param && plan->topo->t.n_expr))))
        return fail(FK_ERR_INVALID, "bad upload arguments");
    const fk::Topology& t = plan->topo->t;
    CU(cudaSetDevice(plan->device));
    plan->n = n;
    if (n && t.n_vars) CU(cudaMemcpyAsync(plan->d_vars, vars, sizeof(double) * (size_t)n * t.n_vars, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    if (n && t.n_expr) CU(cudaMemcpyAsync(plan->d_params, param, sizeof(double) * (size_t)n * t.n_expr, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return FK_OK;
}

int fk_batch_plan_run(fk_batch_plan* plan, void* stream) {
    if (!plan) return fail(FK_ERR_INVALID, "null plan");
    CU(cudaSetDevice(plan->device));
    int e = fk::launch_batch_lm(*plan->prog, plan->n, plan->d_vars, plan->d_params, plan->d_out,
                                plan->d_rep, stream);
    if (e != 0) return cuda_fail((cudaError_t)e, "launch fk_batch_lm_kernel");
    if (plan->n) plan->launches++;
    return FK_OK;
}

int fk_batch_plan_run_lbfgs(fk_batch_plan* plan, void* stream) {
    if (!plan) return fail(FK_ERR_INVALID, "null plan");
    if (plan->topo->t.path != 0) return fail(FK_ERR_TOO_LARGE, "L-BFGS is built for the shared-memory tile paths only");
    CU(cudaSetDevice(plan->device));
    const int e = fk::launch_batch_lbfgs(*plan->prog, plan->full->jcolptr, plan->full->jrow, plan->n, plan->d_vars, plan->d_params,
                                         plan->d_out, plan->d_rep, stream);
    if (e == (int)cudaErrorInvalidConfiguration) return fail(FK_ERR_TOO_LARGE, "the L-BFGS state of one sketch does not fit the tile path");
    if (e != 0) return cuda_fail((cudaError_t)e, "launch fk_batch_lbfgs_kernel");
    if (plan->n) plan->launches++;
    return FK_OK;
}

int fk_batch_plan_download(fk_batch_plan* plan, double* free_out, fk_report* reports, void* stream) {
    if (!plan) return fail(FK_ERR_INVALID, "null plan");
    const fk::Topology& t = plan->topo->t;
    CU(cudaSetDevice(plan->device));
    if (free_out && plan->n && t.n_free)
        CU(cudaMemcpyAsync(free_out, plan->d_out, sizeof(double) * (size_t)plan->n * t.n_free, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    if (reports && plan->n)
        CU(cudaMemcpyAsync(reports, plan->d_rep, sizeof(fk_report) * (size_t)plan->n, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return FK_OK;
}

int fk_batch_plan_device_ptrs(fk_batch_plan* plan, void** vars, void** param, void** free_out, void** reports) {
    if (!plan) return fail(FK_ERR_INVALID, "null plan");
    if (vars) *vars = plan->d_vars;
    if (param) *param = plan->d_params;
    if (free_out) *free_out = plan->d_out;
    if (reports) *reports = plan->d_rep;
    return FK_OK;
}

uint64_t fk_batch_plan_launches(const fk_batch_plan* plan) { return plan ? plan->launches : 0; }

int fk_batch_plan_sync(fk_batch_plan* plan) {
    if (!plan) return fail(FK_ERR_INVALID, "null plan");
    CU(cudaSetDevice(plan->device));
    CU(cudaDeviceSynchronize());
    return FK_OK;
}

int fk_batch_plan_eval(fk_batch_plan* plan, int mode, void* stream) {
    if (!plan) return fail(FK_ERR_INVALID, "null plan");
    const fk::Topology& t = plan->topo->t;
    CU(cudaSetDevice(plan->device));
    if (!plan->d_er) CU(cudaMalloc(&plan->d_er, sizeof(double) * std::max<size_t>(1, (size_t)plan->capacity * t.n_rows)));
    if (!plan->d_ej && mode == 0) CU(cudaMalloc(&plan->d_ej, sizeof(double) * std::max<size_t>(1, (size_t)plan->capacity * t.jac_nnz)));
    int e = fk::launch_batch_eval(*plan->prog, plan->n, plan->d_vars, plan->d_params, plan->d_er, plan->d_ej, mode, stream);
    if (e != 0) return cuda_fail((cudaError_t)e, "launch fk_batch_eval_kernel");
    if (plan->n) plan->launches++;
    return FK_OK;
}

int fk_batch_plan_eval_download(fk_batch_plan* plan, double* out_r, double* out_j, void* stream) {
    if (!plan) return fail(FK_ERR_INVALID, "null plan");
    const fk::Topology& t = plan->topo->t;
    CU(cudaSetDevice(plan->device));
    if (out_r && plan->d_er && plan->n)
        CU(cudaMemcpyAsync(out_r, plan->d_er, sizeof(double) * (size_t)plan->n * t.n_rows, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    if (out_j && plan->d_ej && plan->n)
        CU(cudaMemcpyAsync(out_j, plan->d_ej, sizeof(double) * (size_t)plan->n * t.jac_nnz, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return FK_OK;
}

int fk_fp64_peak_tflops(int device, double* out) {
    if (!out) return fail(FK_ERR_INVALID, "null argument");
    if (usable_devices() == 0) return fail(FK_ERR_NO_DEVICE, "no CUDA device visible");
    CU(cudaSetDevice(device));
    double v = 0.0;
    int e = fk::measure_fp64_peak(&v);
    if (e != 0) return cuda_fail((cudaError_t)e, "fp64 peak microbenchmark");
    *out = v;
    return FK_OK;
}

void fk_set_lm_kernel(int choice) { fk::lm_kernel_choice().store(choice < 0 || choice > 3 ? -1 : choice); }
int fk_get_lm_kernel(void) { return fk::lm_kernel_choice().load(); }
int fk_topology_sketch_kernel_info(const fk_topology* topo, int* available, uint32_t* state_doubles, uint32_t* table_words) {
    if (!topo) return fail(FK_ERR_INVALID, "null topology");
    const fk::Topology::SketchTables& k = topo->t.sk;
    if (available) *available = (k.ok && fk::sk_fits(k.entries, k.tab.size())) ? 1 : 0;
    if (state_doubles) *state_doubles = k.entries;
    if (table_words) *table_words = (uint32_t)k.tab.size();
    return FK_OK;
}

int fk_topology_batch_kernel(const fk_topology* topo, uint32_t n_sketches) {
    if (!topo) return fail(FK_ERR_INVALID, "null topology");
    const fk::Topology::SketchTables& k = topo->t.sk;
    fk::SkProgram probe{};
    probe.entries = k.entries;
    probe.tab_words = (uint32_t)k.tab.size();
    fk::DevProgram p{};
    p.sketch_prog = (k.ok && fk::sk_fits(k.entries, k.tab.size())) ? &probe : nullptr;
    if (!fk::batch_lm_uses_sketch_kernel(p, n_sketches)) return 0;
    return fk::sketch_kernel_is_pair(probe, n_sketches) ? 2 : 1;
}

void* fk_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
void fk_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

// Chunk of the host-buffer pipelines: about an eighth of the request (copies of one chunk overlap the kernels of the
// others), at least 2,048 sketches, and whole waves of the sketch-per-thread kernel when that is the kernel in use.
static uint32_t pipeline_chunk(fk_topology* topo, int device, uint32_t total, uint32_t n_chunks) {
    uint32_t chunk = std::min(total, std::max<uint32_t>(2048, (total + n_chunks - 1) / n_chunks));
    const fk::DevProgram* prog = nullptr;
    if (topo->program_for(device, &prog) == FK_OK && fk::batch_lm_uses_sketch_kernel(*prog, total)) {
        static int sms[64] = {0};
        if (device >= 0 && device < 64 && sms[device] == 0) {
            int v = 0;
            if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) == cudaSuccess) sms[device] = v;
        }
        const uint32_t wave = fk::sketch_kernel_wave(*prog->sketch_prog, (device >= 0 && device < 64 && sms[device]) ? sms[device] : 148);
        if (wave && chunk > wave / 2) chunk = std::min(total, ((chunk + wave - 1) / wave) * wave);
    }
    return chunk;
}

// ---- host-buffer batch: shard by sketch over the devices, pipeline chunks per device ----------------
static int launch_optimizer(const fk::DevProgram& prog, const DeviceProgram& full, int optimizer, uint32_t n, const double* vars,
                            const double* params, double* out, fk_report* reps, cudaStream_t st) {
    if (optimizer == 1) return fk::launch_batch_lbfgs(prog, full.jcolptr, full.jrow, n, vars, params, out, reps, st);
    return fk::launch_batch_lm(prog, n, vars, params, out, reps, st);
}

// Small request (one sketch, a handful): the caller's (pageable) arrays go through ONE pinned staging block -- one
// copy in, the kernel, one copy out, one synchronisation -- instead of four pageable copies on the chunk pipeline.
// (Enqueue and finish are separate so that a caller can overlap host work with the request.)
struct SmallRequest {
    DevicePipeline* pl = nullptr;
    std::unique_lock<std::mutex> lock;  // the pipeline's staging block is ours until finish()
    size_t in_al = 0, out_x = 0, out_r = 0;
};
static bool small_fits(const fk::Topology& t, uint32_t total) {
    const size_t in_v = sizeof(double) * (size_t)total * t.n_vars, in_p = sizeof(double) * (size_t)total * t.n_expr;
    const size_t out_x = sizeof(double) * (size_t)total * t.n_free, out_r = sizeof(fk_report) * (size_t)total;
    return ((in_v + in_p + 15) & ~size_t(15)) + out_x + out_r <= DevicePipeline::kSmallBytes;
}
static int small_enqueue(fk_topology* topo, int device, uint32_t lo, uint32_t hi, const double* vars, const double* param, int optimizer,
                         SmallRequest& rq, std::string* err) {
    const fk::Topology& t = topo->t;
    const uint32_t total = hi - lo;
    DevicePipeline* pl = topo->pipeline_for(device);
    rq.lock = std::unique_lock<std::mutex>(pl->mu);
    rq.pl = pl;
    const size_t in_v = sizeof(double) * (size_t)total * t.n_vars, in_p = sizeof(double) * (size_t)total * t.n_expr;
    rq.out_x = sizeof(double) * (size_t)total * t.n_free;
    rq.out_r = sizeof(fk_report) * (size_t)total;
    rq.in_al = (in_v + in_p + 15) & ~size_t(15);
    const fk::DevProgram* prog = nullptr;
    const DeviceProgram* full = nullptr;
    int rc = topo->program_for(device, &prog, &full);
    auto give_up = [&](int code) { if (err) *err = g_error; rq.lock.unlock(); rq.pl = nullptr; return code; };
    if (rc != FK_OK) return give_up(rc);
    auto bad = [&](cudaError_t e, const char* what) { return give_up(cuda_fail(e, what)); };
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return bad(e, "cudaSetDevice");
    if (!pl->small_h || !pl->small_d || !pl->small_stream) {
        // all three or none: a half-initialised block would poison every later call on this topology
        if (pl->small_h) cudaFreeHost(pl->small_h);
        if (pl->small_d) cudaFree(pl->small_d);
        if (pl->small_stream) cudaStreamDestroy(pl->small_stream);
        pl->small_h = pl->small_d = nullptr;
        pl->small_stream = nullptr;
        if ((e = cudaMallocHost((void**)&pl->small_h, DevicePipeline::kSmallBytes)) != cudaSuccess) { pl->small_h = nullptr; return bad(e, "cudaMallocHost"); }
        if ((e = cudaMalloc((void**)&pl->small_d, DevicePipeline::kSmallBytes)) != cudaSuccess) {
            cudaFreeHost(pl->small_h); pl->small_h = pl->small_d = nullptr;
            return bad(e, "cudaMalloc");
        }
        if ((e = cudaStreamCreateWithFlags(&pl->small_stream, cudaStreamNonBlocking)) != cudaSuccess) {
            cudaFreeHost(pl->small_h); cudaFree(pl->small_d); pl->small_h = pl->small_d = nullptr; pl->small_stream = nullptr;
            return bad(e, "cudaStreamCreate");
        }
    }
    std::memcpy(pl->small_h, vars + (size_t)lo * t.n_vars, in_v);
    if (in_p) std::memcpy(pl->small_h + in_v, param + (size_t)lo * t.n_expr, in_p);
    // A few kilobytes: the kernel reads its inputs from, and writes its results to, the pinned block itself (pinned host
    // memory is mapped into the device's address space under unified addressing; the kernel touches every input and output
    // once) -- two DMA set-ups less on a path that is all latency.  FK_NO_ZERO_COPY=1 keeps the two copies (A/B knob).
    static const bool no_zero_copy = std::getenv("FK_NO_ZERO_COPY") != nullptr;
    static const bool can_map = [] {
        int dev = 0, ok = 0;
        return cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&ok, cudaDevAttrCanUseHostPointerForRegisteredMem, dev) == cudaSuccess && ok != 0;
    }();
    const bool zero_copy = !no_zero_copy && can_map && rq.in_al + rq.out_x + rq.out_r <= 8 * 1024;
    unsigned char* base = zero_copy ? pl->small_h : pl->small_d;
    if (!zero_copy && (e = cudaMemcpyAsync(pl->small_d, pl->small_h, in_v + in_p, cudaMemcpyHostToDevice, pl->small_stream)) != cudaSuccess) return bad(e, "cudaMemcpyAsync");
    const int le = launch_optimizer(*prog, *full, optimizer, total, (const double*)base, (const double*)(base + in_v),
                                    (double*)(base + rq.in_al), (fk_report*)(base + rq.in_al + rq.out_x), pl->small_stream);
    if (le == (int)cudaErrorInvalidConfiguration && optimizer == 1) { fail(FK_ERR_TOO_LARGE, "the L-BFGS state of one sketch does not fit the tile path"); return give_up(FK_ERR_TOO_LARGE); }
    if (le != 0) return bad((cudaError_t)le, "launch batched solve kernel");
    if (!zero_copy && (e = cudaMemcpyAsync(pl->small_h + rq.in_al, pl->small_d + rq.in_al, rq.out_x + rq.out_r, cudaMemcpyDeviceToHost, pl->small_stream)) != cudaSuccess) return bad(e, "cudaMemcpyAsync");
    return FK_OK;
}
static int small_finish(const fk::Topology& t, uint32_t lo, SmallRequest& rq, double* free_out, fk_report* reports, std::string* err) {
    if (!rq.pl) return FK_OK;
    DevicePipeline* pl = rq.pl;
    int rc = FK_OK;
    const cudaError_t e = cudaStreamSynchronize(pl->small_stream);
    if (e != cudaSuccess) {
        rc = cuda_fail(e, "batch kernel / copy failed");
        if (err) *err = g_error;
    } else {
        std::memcpy(free_out + (size_t)lo * t.n_free, pl->small_h + rq.in_al, rq.out_x);
        if (reports) std::memcpy(reports + lo, pl->small_h + rq.in_al + rq.out_x, rq.out_r);
    }
    rq.lock.unlock();
    rq.pl = nullptr;
    return rc;
}

// token != nullptr: return once every chunk is enqueued; *token names the call for fk_batch_solve_device_wait (0: already complete).
static int run_device_range(fk_topology* topo, int device, uint32_t lo, uint32_t hi, const double* vars, const double* param,
                            double* free_out, fk_report* reports, std::string* err, int optimizer, uint64_t* token) {
    const fk::Topology& t = topo->t;
    const uint32_t total = hi - lo;
    if (token) *token = 0;
    if (total == 0) return FK_OK;
    if (small_fits(t, total)) {
        SmallRequest rq;
        int rc = small_enqueue(topo, device, lo, hi, vars, param, optimizer, rq, err);
        if (rc == FK_OK) rc = small_finish(t, lo, rq, free_out, reports, err);
        return rc;
    }
    DevicePipeline* pl = topo->pipeline_for(device);
    std::lock_guard<std::mutex> lock(pl->mu);
    constexpr uint32_t kStreams = DevicePipeline::kStreams;
    // chunks small enough to overlap H2D / kernel / D2H, large enough to fill the machine
    static const uint32_t n_chunks = [] {
        const char* e = std::getenv("FK_E2E_CHUNKS");  // tuning knob
        const int v = e ? std::atoi(e) : 0;
        return (uint32_t)(v >= 1 && v <= 256 ? v : 8);
    }();
    uint32_t chunk = optimizer == 0 ? pipeline_chunk(topo, device, total, n_chunks)
                                    : std::min(total, std::max<uint32_t>(2048, (total + n_chunks - 1) / n_chunks));
    int rc = FK_OK;
    if (pl->chunk < chunk) {
        pl->release();
        for (uint32_t s = 0; s < kStreams && rc == FK_OK; s++) {
            rc = fk_batch_plan_create(topo, chunk, device, &pl->plans[s]);
            if (rc == FK_OK && cudaStreamCreateWithFlags(&pl->streams[s], cudaStreamNonBlocking) != cudaSuccess)
                rc = fail(FK_ERR_CUDA, "cudaStreamCreate failed");
        }
        if (rc == FK_OK) pl->chunk = chunk;
        else pl->release();
    } else {
        if (pl->chunk < total) chunk = std::min(chunk, pl->chunk);  // (plans sized by an earlier, smaller request)
    }
    // FK_E2E_TRACE=1 (debug): CUDA-event timeline of the chunks of one call to stderr (upload done / kernel done / download done)
    static const bool trace = std::getenv("FK_E2E_TRACE") != nullptr;
    std::vector<cudaEvent_t> tev;
    cudaEvent_t t_start = nullptr;
    auto stamp = [&](cudaStream_t st) {
        if (!trace) return;
        cudaEvent_t ev;
        cudaEventCreate(&ev);
        cudaEventRecord(ev, st);
        tev.push_back(ev);
    };
    if (trace) {
        cudaEventCreate(&t_start);
        cudaEventRecord(t_start, pl->streams[0]);
    }
    uint32_t s = 0;
    for (uint32_t at = lo; at < hi && rc == FK_OK; at += chunk, s = (s + 1) % kStreams) {
        uint32_t cnt = std::min(chunk, hi - at);
        // a plan's buffers are reused only after its previous chunk has fully drained
        if (cudaStreamSynchronize(pl->streams[s]) != cudaSuccess) { rc = fail(FK_ERR_CUDA, "stream sync failed"); break; }
        rc = fk_batch_plan_upload(pl->plans[s], cnt, vars + (size_t)at * t.n_vars, param ? param + (size_t)at * t.n_expr : nullptr, pl->streams[s]);
        stamp(pl->streams[s]);
        if (rc == FK_OK) rc = optimizer == 1 ? fk_batch_plan_run_lbfgs(pl->plans[s], pl->streams[s]) : fk_batch_plan_run(pl->plans[s], pl->streams[s]);
        stamp(pl->streams[s]);
        if (rc == FK_OK) rc = fk_batch_plan_download(pl->plans[s], free_out + (size_t)at * t.n_free, reports ? reports + at : nullptr, pl->streams[s]);
        stamp(pl->streams[s]);
    }
    if (token && !trace && rc == FK_OK) {
        const uint64_t my_token = pl->next_token++;
        const int slot = (int)(my_token & 1u);
        for (uint32_t k = 0; k < kStreams; k++) {
            if (!pl->streams[k]) continue;
            if (!pl->done_ev[slot][k] && cudaEventCreateWithFlags(&pl->done_ev[slot][k], cudaEventDisableTiming) != cudaSuccess)
                return fail(FK_ERR_CUDA, "cudaEventCreate failed");
            if (cudaEventRecord(pl->done_ev[slot][k], pl->streams[k]) != cudaSuccess) return fail(FK_ERR_CUDA, "cudaEventRecord failed");
        }
        *token = my_token;
        return FK_OK;
    }
    for (uint32_t k = 0; k < kStreams; k++)
        if (pl->streams[k] && cudaStreamSynchronize(pl->streams[k]) != cudaSuccess && rc == FK_OK)
            rc = cuda_fail(cudaGetLastError(), "batch kernel / copy failed");
    if (trace) {
        for (size_t c = 0; c + 2 < tev.size(); c += 3) {
            float a = 0, b = 0, d = 0;
            cudaEventElapsedTime(&a, t_start, tev[c]);
            cudaEventElapsedTime(&b, t_start, tev[c + 1]);
            cudaEventElapsedTime(&d, t_start, tev[c + 2]);
            fprintf(stderr, "[e2e trace] chunk %2zu: upload done %7.1f us, kernel done %7.1f us, download done %7.1f us\n", c / 3, a * 1e3f, b * 1e3f, d * 1e3f);
        }
        for (cudaEvent_t ev : tev) cudaEventDestroy(ev);
        cudaEventDestroy(t_start);
    }
    if (rc != FK_OK && err) *err = g_error;
    return rc;
}

// ---- System::solve level batch: scale, perturbation and write-back on the device -----------------------------
static int prepare_tables_for(fk_topology* topo, int device, const fk_prepare_opts& o, PrepareTables** out) {
    const fk::Topology& t = topo->t;
    std::lock_guard<std::mutex> lock(topo->mu);
    auto& p = topo->prepare_tables[device];
    std::vector<uint32_t> list;
    if (o.flags & FK_PREP_PERTURB) {
        if (o.perturb_vars) list.assign(o.perturb_vars, o.perturb_vars + o.n_perturb);
        else list = t.free_vars;
        for (size_t q = 0; q < list.size(); q++)
            if (list[q] >= t.n_vars || (q && list[q] <= list[q - 1]))
                return fail(FK_ERR_INVALID, "perturbed variables must be in range, ascending and distinct");
    }
    if (p && p->perturb_vars == list && p->seed == o.seed) { *out = p.get(); return FK_OK; }
    std::unique_ptr<PrepareTables> q(new PrepareTables());
    CU(cudaSetDevice(device));
    q->device = device; q->perturb_vars = list; q->seed = o.seed;
    std::vector<double> draws(2 * list.size() + 1, 0.0);
    uint32_t st = o.seed;
    for (size_t k = 0; k < 2 * list.size(); k++) {  // rand.rs:24-39
        st = st * 1664525u + 1013904223u;
        draws[k] = (1.0 / 4294967295.0) * (double)st;
    }
    CU(cudaMalloc(&q->d_kind, std::max<uint32_t>(t.n_expr, 1)));
    CU(cudaMalloc(&q->d_perturb, sizeof(uint32_t) * std::max<size_t>(list.size(), 1)));
    CU(cudaMalloc(&q->d_draws, sizeof(double) * draws.size()));
    if (t.n_expr) CU(cudaMemcpy(q->d_kind, t.kind.data(), t.n_expr, cudaMemcpyHostToDevice));
    if (!list.empty()) CU(cudaMemcpy(q->d_perturb, list.data(), sizeof(uint32_t) * list.size(), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(q->d_draws, draws.data(), sizeof(double) * draws.size(), cudaMemcpyHostToDevice));
    {
        std::vector<uint32_t> where(t.n_vars, 0xFFFFFFFFu), free_draw(std::max<uint32_t>(t.n_free, 1), 0xFFFFFFFFu), fix_draw(std::max<uint32_t>(t.sk.nfix, 1), 0xFFFFFFFFu);
        for (size_t j = 0; j < list.size(); j++) where[list[j]] = (uint32_t)j;
        for (uint32_t c = 0; c < t.n_free; c++) free_draw[c] = where[t.free_vars[c]];
        if (t.sk.ok)
            for (uint32_t i = 0; i < t.sk.nfix; i++) fix_draw[i] = where[t.sk.tab[t.sk.off_fix + i]];
        CU(cudaMalloc(&q->d_free_draw, sizeof(uint32_t) * free_draw.size()));
        CU(cudaMalloc(&q->d_fix_draw, sizeof(uint32_t) * fix_draw.size()));
        CU(cudaMemcpy(q->d_free_draw, free_draw.data(), sizeof(uint32_t) * free_draw.size(), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(q->d_fix_draw, fix_draw.data(), sizeof(uint32_t) * fix_draw.size(), cudaMemcpyHostToDevice));
    }
    p = std::move(q);
    *out = p.get();
    return FK_OK;
}

// token == nullptr: the call returns when the results are in the caller's buffers.  Otherwise it returns once every chunk is
// enqueued (the host still waits, chunk by chunk, for a free plan: at most kStreams chunks are in flight) and *token names the
// call for fk_batch_system_solve_wait.
static int system_solve_enqueue(const fk_topology* topo_c, int device, uint32_t n, const double* raw_vars, const double* raw_param,
                                const fk_prepare_opts* opts, double* free_out, double* scales_out, fk_report* reports, uint64_t* token) {
    fk_topology* topo = const_cast<fk_topology*>(topo_c);
    if (!topo || !opts) return fail(FK_ERR_INVALID, "null argument");
    if (token) *token = 0;
    if (n == 0) return FK_OK;
    const fk::Topology& t = topo->t;
    if (!raw_vars || !free_out || (!raw_param && t.n_expr)) return fail(FK_ERR_INVALID, "null buffer");
    if (t.path == 2) return fail(FK_ERR_TOO_LARGE, "topology needs the global sparse path; use fk_system_solve");
    const int ndev = usable_devices();
    if (ndev == 0) return fail(FK_ERR_NO_DEVICE, "no CUDA device visible; fiksi_b200 has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(FK_ERR_INVALID, "device index out of range");
    PrepareTables* pt = nullptr;
    int rc = prepare_tables_for(topo, device, *opts, &pt);
    if (rc != FK_OK) return rc;
    const bool shared = (opts->flags & FK_PREP_SHARED_PARAM) != 0;
    DevicePipeline* pl = topo->pipeline_for(device);
    std::lock_guard<std::mutex> lock(pl->mu);
    constexpr uint32_t kStreams = DevicePipeline::kStreams;
    static const uint32_t n_chunks = [] {
        const char* e = std::getenv("FK_E2E_CHUNKS");  // tuning knob
        const int v = e ? std::atoi(e) : 0;
        return (uint32_t)(v >= 1 && v <= 256 ? v : 16);
    }();
    uint32_t chunk = pipeline_chunk(topo, device, n, n_chunks);
    if (pl->chunk < chunk) {
        pl->release();
        for (uint32_t s = 0; s < kStreams && rc == FK_OK; s++) {
            rc = fk_batch_plan_create(topo, chunk, device, &pl->plans[s]);
            if (rc == FK_OK && cudaStreamCreateWithFlags(&pl->streams[s], cudaStreamNonBlocking) != cudaSuccess) rc = fail(FK_ERR_CUDA, "cudaStreamCreate failed");
        }
        if (rc == FK_OK) pl->chunk = chunk;
        else { pl->release(); return rc; }
    }
    CU(cudaSetDevice(device));
    for (uint32_t s = 0; s < kStreams; s++) {  // unscaled input / scale buffers of the plans, on first use
        fk_batch_plan* p = pl->plans[s];
        if (!p->d_raw_vars) CU(cudaMalloc(&p->d_raw_vars, sizeof(double) * std::max<size_t>(1, (size_t)p->capacity * t.n_vars)));
        if (!p->d_raw_param) CU(cudaMalloc(&p->d_raw_param, sizeof(double) * std::max<size_t>(1, (size_t)p->capacity * t.n_expr)));
        if (!p->d_scales) CU(cudaMalloc(&p->d_scales, sizeof(double) * p->capacity));
    }
    static const bool no_fuse = std::getenv("FK_NO_FUSED_PREPARE") != nullptr;  // A/B knob
    static const uint32_t use_streams = [] {
        const char* e = std::getenv("FK_E2E_STREAMS");  // tuning knob: chunks in flight (at most kStreams)
        const int v = e ? std::atoi(e) : 0;
        return (uint32_t)(v >= 1 && v <= (int)DevicePipeline::kStreams ? v : (int)DevicePipeline::kStreams);
    }();
    // FK_E2E_TRACE=1 (debug): CUDA-event timeline of the chunks of one call to stderr (upload done / kernel done / download done)
    static const bool trace = std::getenv("FK_E2E_TRACE") != nullptr;
    std::vector<cudaEvent_t> tev;
    cudaEvent_t t_start = nullptr;
    if (trace) {
        cudaEventCreate(&t_start);
        cudaEventRecord(t_start, pl->streams[0]);
    }
    // One parameter row for all sketches: uploaded once per call, every stream waits for it (a 300-byte copy per chunk kept the
    // H2D engine ~10 us per chunk, which the first chunks of a call -- the ones the device is waiting for -- paid in full).
    const uint64_t my_token = pl->next_token++;
    const int slot = (int)(my_token & 1u);
    const double* d_shared_param = nullptr;
    if (shared && t.n_expr) {
        if (!pl->shared_param_ev) CU(cudaEventCreateWithFlags(&pl->shared_param_ev, cudaEventDisableTiming));
        if (pl->shared_param_cap < t.n_expr) {  // (two rows: the kernels of the previous call may still read theirs)
            for (int a = 0; a < 2; a++) {
                if (pl->d_shared_param[a]) CU(cudaFree(pl->d_shared_param[a]));
                pl->d_shared_param[a] = nullptr;
                CU(cudaMalloc(&pl->d_shared_param[a], sizeof(double) * t.n_expr));
            }
            pl->shared_param_cap = t.n_expr;
        }
        d_shared_param = pl->d_shared_param[slot];
        CU(cudaMemcpyAsync(pl->d_shared_param[slot], raw_param, sizeof(double) * t.n_expr, cudaMemcpyHostToDevice, pl->streams[0]));
        CU(cudaEventRecord(pl->shared_param_ev, pl->streams[0]));
        for (uint32_t k = 1; k < use_streams; k++) CU(cudaStreamWaitEvent(pl->streams[k], pl->shared_param_ev, 0));
    }
    // (Chunks that taper at both ends of a call -- a quarter and a half of the regular size first, halving down to 1,024 sketches at the
    // end -- bring the first kernel forward from 65 to 24 us and change nothing: 1,108 against 1,092 us per call.  The last kernel ends
    // ~1,070 us after the call started either way; the chunked kernels together take ~1,040 us where one launch over the resident
    // batch takes 905.)
    uint32_t s = 0;
    for (uint32_t at = 0; at < n && rc == FK_OK; at += chunk, s = (s + 1) % use_streams) {
        const uint32_t cnt = std::min(chunk, n - at);
        fk_batch_plan* p = pl->plans[s];
        cudaStream_t st = pl->streams[s];
        const bool fused = !no_fuse && fk::batch_lm_uses_sketch_kernel(*p->prog, cnt) && fk::sk_raw_fits(*p->prog->sketch_prog, shared);
        CU(cudaStreamSynchronize(st));  // the plan's buffers are reused only after its previous chunk has drained
        p->n = cnt;
        if (t.n_vars) CU(cudaMemcpyAsync(p->d_raw_vars, raw_vars + (size_t)at * t.n_vars, sizeof(double) * (size_t)cnt * t.n_vars, cudaMemcpyHostToDevice, st));
        if (t.n_expr && !shared)
            CU(cudaMemcpyAsync(p->d_raw_param, raw_param + (size_t)at * t.n_expr, sizeof(double) * (size_t)cnt * t.n_expr, cudaMemcpyHostToDevice, st));
        const double* d_param = shared ? d_shared_param : p->d_raw_param;
        if (trace) { cudaEvent_t ev; cudaEventCreate(&ev); cudaEventRecord(ev, st); tev.push_back(ev); }
        int e;
        if (fused) {  // the sketch-per-thread kernel scales, perturbs and writes back itself (SkRaw)
            fk::SkRaw raw{};
            raw.raw_vars = p->d_raw_vars; raw.raw_param = d_param; raw.shared_param = shared ? 1u : 0u;
            raw.kinds = pt->d_kind; raw.free_draw = pt->d_free_draw; raw.fix_draw = pt->d_fix_draw; raw.draws = pt->d_draws;
            raw.scales = p->d_scales;
            e = fk::launch_batch_lm_sketch(*p->prog->sketch_prog, cnt, nullptr, nullptr, p->d_out, p->d_rep, st, &raw);
            p->launches += 1;
        } else {
            e = fk::launch_batch_prepare(cnt, t.n_vars, t.n_expr, pt->d_kind, shared, (uint32_t)pt->perturb_vars.size(), pt->d_perturb, pt->d_draws,
                                         p->d_raw_vars, d_param, p->d_vars, p->d_params, p->d_scales, st);
            if (e == 0) e = fk::launch_batch_lm(*p->prog, cnt, p->d_vars, p->d_params, p->d_out, p->d_rep, st);
            if (e == 0) e = fk::launch_batch_unscale(cnt, t.n_free, p->d_scales, p->d_out, st);
            p->launches += 3;
        }
        if (e != 0) return cuda_fail((cudaError_t)e, "launch fk_batch_system_solve kernels");
        if (trace) { cudaEvent_t ev; cudaEventCreate(&ev); cudaEventRecord(ev, st); tev.push_back(ev); }
        if (t.n_free) CU(cudaMemcpyAsync(free_out + (size_t)at * t.n_free, p->d_out, sizeof(double) * (size_t)cnt * t.n_free, cudaMemcpyDeviceToHost, st));
        if (scales_out) CU(cudaMemcpyAsync(scales_out + at, p->d_scales, sizeof(double) * cnt, cudaMemcpyDeviceToHost, st));
        if (reports) CU(cudaMemcpyAsync(reports + at, p->d_rep, sizeof(fk_report) * (size_t)cnt, cudaMemcpyDeviceToHost, st));
        if (trace) { cudaEvent_t ev; cudaEventCreate(&ev); cudaEventRecord(ev, st); tev.push_back(ev); }
    }
    if (token && !trace) {
        for (uint32_t k = 0; k < kStreams; k++) {
            if (!pl->streams[k]) continue;
            if (!pl->done_ev[slot][k]) CU(cudaEventCreateWithFlags(&pl->done_ev[slot][k], cudaEventDisableTiming));
            CU(cudaEventRecord(pl->done_ev[slot][k], pl->streams[k]));
        }
        *token = my_token;
        return rc;
    }
    for (uint32_t k = 0; k < kStreams; k++)
        if (pl->streams[k]) CU(cudaStreamSynchronize(pl->streams[k]));
    if (token) *token = my_token;  // (tracing: the call has completed; waiting for it is a no-op)
    if (trace) {
        for (size_t c = 0; c + 2 < tev.size() + 0 && c < tev.size(); c += 3) {
            float a = 0, b = 0, d = 0;
            cudaEventElapsedTime(&a, t_start, tev[c]);
            cudaEventElapsedTime(&b, t_start, tev[c + 1]);
            cudaEventElapsedTime(&d, t_start, tev[c + 2]);
            fprintf(stderr, "[e2e trace] chunk %2zu: upload done %7.1f us, kernel done %7.1f us, download done %7.1f us\n", c / 3, a * 1e3f, b * 1e3f, d * 1e3f);
        }
        for (cudaEvent_t ev : tev) cudaEventDestroy(ev);
        cudaEventDestroy(t_start);
    }
    return rc;
}

int fk_batch_system_solve(const fk_topology* topo, int device, uint32_t n, const double* raw_vars, const double* raw_param,
                          const fk_prepare_opts* opts, double* free_out, double* scales_out, fk_report* reports) {
    return system_solve_enqueue(topo, device, n, raw_vars, raw_param, opts, free_out, scales_out, reports, nullptr);
}

int fk_batch_system_solve_begin(const fk_topology* topo, int device, uint32_t n, const double* raw_vars, const double* raw_param,
                                const fk_prepare_opts* opts, double* free_out, double* scales_out, fk_report* reports, uint64_t* token) {
    if (!token) return fail(FK_ERR_INVALID, "null token");
    return system_solve_enqueue(topo, device, n, raw_vars, raw_param, opts, free_out, scales_out, reports, token);
}

int fk_batch_system_solve_wait(const fk_topology* topo_c, int device, uint64_t token) {
    fk_topology* topo = const_cast<fk_topology*>(topo_c);
    if (!topo) return fail(FK_ERR_INVALID, "null topology");
    if (token == 0) return FK_OK;  // (an empty call)
    const int ndev = usable_devices();
    if (device < 0 || device >= ndev) return fail(FK_ERR_INVALID, "device index out of range");
    DevicePipeline* pl = topo->pipeline_for(device);
    cudaEvent_t ev[DevicePipeline::kStreams];
    {
        std::lock_guard<std::mutex> lock(pl->mu);
        if (token >= pl->next_token) return fail(FK_ERR_INVALID, "unknown token");
        // (a token older than the last two calls: its slot now holds a later call's events, recorded behind this call's work on
        // the same streams, so waiting for those covers it)
        for (uint32_t k = 0; k < DevicePipeline::kStreams; k++) ev[k] = pl->done_ev[token & 1u][k];
    }
    for (uint32_t k = 0; k < DevicePipeline::kStreams; k++)
        if (ev[k]) CU(cudaEventSynchronize(ev[k]));
    return FK_OK;
}

int fk_batch_solve_device(const fk_topology* topo_c, int device, uint32_t n, const double* vars, const double* param,
                          double* free_out, fk_report* reports) {
    fk_topology* topo = const_cast<fk_topology*>(topo_c);
    if (!topo) return fail(FK_ERR_INVALID, "null topology");
    if (n == 0) return FK_OK;
    if (!vars || !free_out || (!param && topo->t.n_expr)) return fail(FK_ERR_INVALID, "null buffer");
    int ndev = usable_devices();
    if (ndev == 0) return fail(FK_ERR_NO_DEVICE, "no CUDA device visible; fiksi_b200 has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(FK_ERR_INVALID, "device index out of range");
    return run_device_range(topo, device, 0, n, vars, param, free_out, reports, nullptr);
}

int fk_batch_solve_device_begin(const fk_topology* topo_c, int device, uint32_t n, const double* vars, const double* param,
                                double* free_out, fk_report* reports, uint64_t* token) {
    fk_topology* topo = const_cast<fk_topology*>(topo_c);
    if (!topo || !token) return fail(FK_ERR_INVALID, "null topology / token");
    *token = 0;
    if (n == 0) return FK_OK;
    if (!vars || !free_out || (!param && topo->t.n_expr)) return fail(FK_ERR_INVALID, "null buffer");
    int ndev = usable_devices();
    if (ndev == 0) return fail(FK_ERR_NO_DEVICE, "no CUDA device visible; fiksi_b200 has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(FK_ERR_INVALID, "device index out of range");
    return run_device_range(topo, device, 0, n, vars, param, free_out, reports, nullptr, 0, token);
}

int fk_batch_solve_device_wait(const fk_topology* topo, int device, uint64_t token) { return fk_batch_system_solve_wait(topo, device, token); }

int fk_batch_solve(const fk_topology* topo_c, uint32_t n, const double* vars, const double* param, double* free_out,
                   fk_report* reports, int n_gpus) {
    fk_topology* topo = const_cast<fk_topology*>(topo_c);
    if (!topo) return fail(FK_ERR_INVALID, "null topology");
    if (n == 0) return FK_OK;
    if (!vars || !free_out || (!param && topo->t.n_expr)) return fail(FK_ERR_INVALID, "null buffer");
    int ndev = usable_devices();
    if (ndev == 0) return fail(FK_ERR_NO_DEVICE, "no CUDA device visible; fiksi_b200 has no CPU fallback");
    if (n_gpus <= 0 || n_gpus > ndev) n_gpus = ndev;
    if (n_gpus == 1) {
        int cur = 0;  // one process per GPU: honour the caller's current device
        if (cudaGetDevice(&cur) != cudaSuccess) cur = 0;
        return run_device_range(topo, cur, 0, n, vars, param, free_out, reports, nullptr);
    }
    std::vector<int> rcs(n_gpus, FK_OK);
    std::vector<std::string> errs(n_gpus);
    std::vector<std::thread> pool;
    for (int g = 0; g < n_gpus; g++) {
        uint32_t lo = (uint32_t)((uint64_t)n * g / n_gpus), hi = (uint32_t)((uint64_t)n * (g + 1) / n_gpus);
        pool.emplace_back([=, &rcs, &errs]() { rcs[g] = run_device_range(topo, g, lo, hi, vars, param, free_out, reports, &errs[g]); });
    }
    for (auto& th : pool) th.join();
    for (int g = 0; g < n_gpus; g++)
        if (rcs[g] != FK_OK) return fail(rcs[g], errs[g]);
    return FK_OK;
}

// ---- single system with a cached topology (any path) --------------------------------------------------
int fk_topology_lm_solve(fk_topology* topo, const double* vars, const double* param, double* free_values,
                         fk_report* report) {
    if (!topo || !free_values || (!vars && topo->t.n_vars)) return fail(FK_ERR_INVALID, "null argument");
    int ndev = usable_devices();
    if (ndev == 0) return fail(FK_ERR_NO_DEVICE, "no CUDA device visible; fiksi_b200 has no CPU fallback");
    int cur = 0;
    if (cudaGetDevice(&cur) != cudaSuccess) cur = 0;
    const fk::Topology& t = topo->t;
    if (t.path == 2) {
        fk::SparseSolver* sp = nullptr;
        std::string err;
        int rc = topo->sparse_for(cur, &sp, &err);
        if (rc != FK_OK) return fail(rc, err);
        // the caller's free values are the starting point (== `variables` of lm.rs:21)
        if (!param && t.n_expr) return fail(FK_ERR_INVALID, "null parameter array");
        std::lock_guard<std::mutex> lock(topo->sparse_mu);
        rc = sp->solve(vars, param, free_values, report, &err);
        return rc == FK_OK ? FK_OK : fail(rc, err);
    }
    std::vector<double> v(vars, vars + t.n_vars), out(std::max<uint32_t>(1, t.n_free));
    for (uint32_t f = 0; f < t.n_free; f++) v[t.free_vars[f]] = free_values[f];
    std::vector<double> zero;
    if (!param && t.n_expr) zero.assign(t.n_expr, 0.0);
    // One system is latency bound (one warp walks the whole LM loop): the lane count chosen for batch throughput
    // (4-16 by footprint) leaves lanes of that warp idle.  A twin of the topology with all 32 lanes, built on the
    // first single-system solve, serves these calls (same operations per entry, same results; tools/lat_probe.py:
    // 144 -> 89 us on the mixed-primitive sketch).
    fk_topology* exec = topo->lanes32();
    int rc = run_device_range(exec, cur, 0, 1, v.data(), param ? param : zero.data(), out.data(), report, nullptr);
    if (rc == FK_OK && t.n_free) std::memcpy(free_values, out.data(), sizeof(double) * t.n_free);
    return rc;
}

int fk_topology_eval(fk_topology* topo, const double* vars, const double* param, const double* free_values,
                     double* out_r, double* out_j, int repeats, float* ms_per_eval) {
    if (!topo || !vars || !free_values) return fail(FK_ERR_INVALID, "null argument");
    if (usable_devices() == 0) return fail(FK_ERR_NO_DEVICE, "no CUDA device visible; fiksi_b200 has no CPU fallback");
    if (topo->t.path != 2) return fail(FK_ERR_INVALID, "fk_topology_eval serves the global sparse path; use fk_batch_plan_eval");
    int cur = 0;
    if (cudaGetDevice(&cur) != cudaSuccess) cur = 0;
    fk::SparseSolver* sp = nullptr;
    std::string err;
    int rc = topo->sparse_for(cur, &sp, &err);
    if (rc == FK_OK) {
        std::lock_guard<std::mutex> lock(topo->sparse_mu);
        rc = sp->eval_once(vars, param, free_values, out_r, out_j, repeats, ms_per_eval, &err);
    }
    return rc == FK_OK ? FK_OK : fail(rc, err);
}

int fk_topology_last_timing(fk_topology* topo, float* out8) {
    if (!topo || !out8) return fail(FK_ERR_INVALID, "null argument");
    std::memset(out8, 0, 8 * sizeof(float));
    std::lock_guard<std::mutex> lock(topo->mu);
    for (auto& kv : topo->sparse) {
        const auto& l = kv.second->last;
        out8[0] = l.eval_ms; out8[1] = l.assemble_ms; out8[2] = l.factor_ms; out8[3] = l.tri_ms;
        out8[4] = (float)l.evals; out8[5] = (float)l.factors; out8[6] = l.fwd_ms; out8[7] = l.bwd_ms;
    }
    return FK_OK;
}

// ---- general entry points ------------------------------------------------------------------------------
// "Symbolic analysis once per topology" for the callers that hand over flattened problems (fk_lm_solve,
// fk_lm_solve_batch, fk_system_solve*): a process-wide LRU cache of fk_topology objects keyed by the structural
// content of the problem (sizes, kinds, indices, free set, row list).  The reference repeats the analysis on every
// LM call (fiksi/src/solve/lm.rs:98-104); an interactive solver re-solves the same sketch over and over.  Entries
// are shared_ptrs: an evicted topology lives until the solves that use it have returned.
namespace {

struct TopologyCache {
    struct Entry { uint64_t sig; std::shared_ptr<fk_topology> topo; };
    std::mutex mu;
    std::list<Entry> lru;  // front = most recently used
    std::unordered_multimap<uint64_t, std::list<Entry>::iterator> index;  // signature -> entry
    void drop(std::list<Entry>::iterator it) {
        auto range = index.equal_range(it->sig);
        for (auto q = range.first; q != range.second; ++q)
            if (q->second == it) { index.erase(q); break; }
        lru.erase(it);
    }
    size_t capacity = [] {
        const char* e = std::getenv("FK_TOPOLOGY_CACHE");
        const int v = e ? std::atoi(e) : 1024;
        return (size_t)std::max(0, v);
    }();
    uint64_t hits = 0, misses = 0;
};
TopologyCache& topology_cache() {
    static TopologyCache* c = new TopologyCache();  // never destroyed: CUDA may be gone by static-destruction time
    return *c;
}

uint64_t structural_signature(const fk_problem& p) {
    uint64_t sig = 1469598103934665603ull;
    auto mixb = [&](const void* d, size_t bytes) {
        const unsigned char* c = (const unsigned char*)d;
        for (size_t k = 0; k < bytes; k++) sig = (sig ^ c[k]) * 1099511628211ull;
    };
    const uint32_t hdr[4] = {p.n_vars, p.n_expr, p.n_free, p.n_rows};
    mixb(hdr, sizeof hdr);
    if (p.n_expr) { mixb(p.kind, p.n_expr); mixb(p.idx, 16 * (size_t)p.n_expr); }
    if (p.n_free) mixb(p.free_vars, 4 * (size_t)p.n_free);
    if (p.n_rows) mixb(p.rows, 4 * (size_t)p.n_rows);
    return sig;
}
bool same_structure(const fk::Topology& t, const fk_problem& p) {
    return t.n_vars == p.n_vars && t.n_expr == p.n_expr && t.n_free == p.n_free && t.n_rows == p.n_rows &&
           (!p.n_expr || (!std::memcmp(t.kind.data(), p.kind, p.n_expr) && !std::memcmp(t.idx.data(), p.idx, 16 * (size_t)p.n_expr))) &&
           (!p.n_free || !std::memcmp(t.free_vars.data(), p.free_vars, 4 * (size_t)p.n_free)) &&
           (!p.n_rows || !std::memcmp(t.rows.data(), p.rows, 4 * (size_t)p.n_rows));
}
bool problem_arrays_ok(const fk_problem& p) {
    return !((p.n_vars && !p.vars) || (p.n_expr && (!p.kind || !p.idx)) || (p.n_free && !p.free_vars) || (p.n_rows && !p.rows));
}

std::shared_ptr<fk_topology> cache_lookup(uint64_t sig, const fk_problem& p) {
    TopologyCache& c = topology_cache();
    std::lock_guard<std::mutex> lock(c.mu);
    auto range = c.index.equal_range(sig);
    for (auto q = range.first; q != range.second; ++q)
        if (same_structure(q->second->topo->t, p)) {
            c.lru.splice(c.lru.begin(), c.lru, q->second);  // (list iterators stay valid)
            c.hits++;
            return c.lru.front().topo;
        }
    c.misses++;
    return nullptr;
}
void cache_insert(uint64_t sig, const std::shared_ptr<fk_topology>& topo) {
    TopologyCache& c = topology_cache();
    std::lock_guard<std::mutex> lock(c.mu);
    if (c.capacity == 0) return;
    c.lru.push_front({sig, topo});
    c.index.emplace(sig, c.lru.begin());
    while (c.lru.size() > c.capacity) c.drop(std::prev(c.lru.end()));
    // large single systems keep hundreds of megabytes of factor storage on the device: at most four of them stay
    if (topo->t.path == 2) {
        size_t large = 0;
        for (auto it = c.lru.begin(); it != c.lru.end();) {
            auto cur = it++;
            if (cur->topo->t.path == 2 && ++large > 4) c.drop(cur);
        }
    }
}

}  // namespace

void fk_topology_cache_configure(uint32_t capacity) {
    TopologyCache& c = topology_cache();
    std::lock_guard<std::mutex> lock(c.mu);
    c.capacity = capacity;
    while (c.lru.size() > c.capacity) c.drop(std::prev(c.lru.end()));
}
void fk_topology_cache_clear(void) {
    TopologyCache& c = topology_cache();
    std::lock_guard<std::mutex> lock(c.mu);
    c.index.clear();
    c.lru.clear();
}
void fk_topology_cache_stats(uint64_t* hits, uint64_t* misses, uint32_t* entries) {
    TopologyCache& c = topology_cache();
    std::lock_guard<std::mutex> lock(c.mu);
    if (hits) *hits = c.hits;
    if (misses) *misses = c.misses;
    if (entries) *entries = (uint32_t)c.lru.size();
}

// Staging of the heterogeneous-batch kernel, one per device: pinned host block + device block (grow-only), one stream.
struct HeteroPool {
    std::mutex mu;
    int device = -1;
    unsigned char *h = nullptr, *d = nullptr;
    size_t cap = 0;
    cudaStream_t stream = nullptr;
};
static HeteroPool* hetero_pool_for(int device) {
    static std::mutex m;
    static std::map<int, HeteroPool*> pools;  // never destroyed (CUDA may be gone at static-destruction time)
    std::lock_guard<std::mutex> lock(m);
    HeteroPool*& p = pools[device];
    if (!p) {
        p = new HeteroPool();
        p->device = device;
    }
    return p;
}

// Solves jobs (group index, member) of small path-0 groups with ONE launch of fk_hetero_lm_kernel on `device`.
struct HeteroItem { uint32_t group, problem; };
static int hetero_solve(int device, const std::vector<fk_topology*>& topos32, const std::vector<HeteroItem>& items,
                        const fk_problem* const* problems, double* const* free_values, fk_report* reports) {
    HeteroPool* pool = hetero_pool_for(device);
    std::lock_guard<std::mutex> lock(pool->mu);
    CU(cudaSetDevice(device));
    const size_t ng = topos32.size(), nj = items.size();
    std::vector<fk::DevProgram> progs(ng);
    uint32_t max_state = 1;
    for (size_t g = 0; g < ng; g++) {
        const fk::DevProgram* v = nullptr;
        int rc = topos32[g]->program_for(device, &v);
        if (rc != FK_OK) return rc;
        progs[g] = *v;
        progs[g].sketch_prog = nullptr;
        max_state = std::max(max_state, fk::lm_smem_doubles(v->n, v->m, v->jnnz, v->lnnz));
    }
    // layout of the staging block: [programs][jobs][inputs: vars row + param row per job] | [outputs][reports]
    auto align = [](size_t x) { return (x + 255) & ~size_t(255); };
    std::vector<fk::HeteroJob> jobs(nj);
    size_t in_doubles = 0, out_doubles = 0;
    for (size_t k = 0; k < nj; k++) {
        const fk::Topology& t = topos32[items[k].group]->t;
        jobs[k].prog = items[k].group; jobs[k].pad = 0;
        jobs[k].vars_off = in_doubles; in_doubles += t.n_vars;
        jobs[k].param_off = in_doubles; in_doubles += t.n_expr;
        jobs[k].out_off = out_doubles; out_doubles += t.n_free;
    }
    const size_t o_progs = 0, o_jobs = align(o_progs + ng * sizeof(fk::DevProgram)), o_in = align(o_jobs + nj * sizeof(fk::HeteroJob));
    const size_t o_out = align(o_in + in_doubles * sizeof(double)), o_rep = align(o_out + out_doubles * sizeof(double));
    const size_t total = align(o_rep + nj * sizeof(fk_report));
    if (pool->cap < total) {
        if (pool->h) cudaFreeHost(pool->h);
        if (pool->d) cudaFree(pool->d);
        pool->h = pool->d = nullptr; pool->cap = 0;
        const size_t cap = std::max<size_t>(total + total / 2, 1 << 20);
        CU(cudaMallocHost((void**)&pool->h, cap));
        cudaError_t e = cudaMalloc((void**)&pool->d, cap);
        if (e != cudaSuccess) { cudaFreeHost(pool->h); pool->h = nullptr; return cuda_fail(e, "cudaMalloc"); }
        pool->cap = cap;
    }
    if (!pool->stream) CU(cudaStreamCreateWithFlags(&pool->stream, cudaStreamNonBlocking));
    std::memcpy(pool->h + o_progs, progs.data(), ng * sizeof(fk::DevProgram));
    std::memcpy(pool->h + o_jobs, jobs.data(), nj * sizeof(fk::HeteroJob));
    double* in = reinterpret_cast<double*>(pool->h + o_in);
    for (size_t k = 0; k < nj; k++) {
        const fk::Topology& t = topos32[items[k].group]->t;
        const fk_problem* p = problems[items[k].problem];
        double* v = in + jobs[k].vars_off;
        if (t.n_vars) std::memcpy(v, p->vars, sizeof(double) * t.n_vars);
        for (uint32_t f = 0; f < t.n_free; f++) v[t.free_vars[f]] = free_values[items[k].problem][f];
        double* q = in + jobs[k].param_off;
        if (t.n_expr) {
            if (p->param) std::memcpy(q, p->param, sizeof(double) * t.n_expr);
            else std::fill(q, q + t.n_expr, 0.0);
        }
    }
    CU(cudaMemcpyAsync(pool->d, pool->h, o_out, cudaMemcpyHostToDevice, pool->stream));
    const int e = fk::launch_hetero_lm(reinterpret_cast<const fk::DevProgram*>(pool->d + o_progs), reinterpret_cast<const fk::HeteroJob*>(pool->d + o_jobs),
                                       (uint32_t)nj, max_state, reinterpret_cast<const double*>(pool->d + o_in), reinterpret_cast<double*>(pool->d + o_out),
                                       reinterpret_cast<fk_report*>(pool->d + o_rep), pool->stream);
    if (e != 0) return cuda_fail((cudaError_t)e, "launch fk_hetero_lm_kernel");
    CU(cudaMemcpyAsync(pool->h + o_out, pool->d + o_out, total - o_out, cudaMemcpyDeviceToHost, pool->stream));
    CU(cudaStreamSynchronize(pool->stream));
    const double* out = reinterpret_cast<const double*>(pool->h + o_out);
    const fk_report* reps = reinterpret_cast<const fk_report*>(pool->h + o_rep);
    for (size_t k = 0; k < nj; k++) {
        const fk::Topology& t = topos32[items[k].group]->t;
        if (t.n_free) std::memcpy(free_values[items[k].problem], out + jobs[k].out_off, sizeof(double) * t.n_free);
        if (reports) reports[items[k].problem] = reps[k];
    }
    return FK_OK;
}

int fk_lm_solve_batch(uint32_t n, const fk_problem* const* problems, double* const* free_values, fk_report* reports,
                      int n_gpus) {
    if (n == 0) return FK_OK;
    if (!problems || !free_values) return fail(FK_ERR_INVALID, "null argument");
    const int ndev = usable_devices();
    if (ndev == 0) return fail(FK_ERR_NO_DEVICE, "no CUDA device visible; fiksi_b200 has no CPU fallback");
    if (n_gpus <= 0 || n_gpus > ndev) n_gpus = ndev;
    // ---- group the problems by topology -----------------------------------------------------------------
    struct Group {
        uint64_t sig = 0;
        std::shared_ptr<fk_topology> topo;
        std::vector<uint32_t> members;
        int rc = FK_OK;
        std::string err;
        // staging of the group's uniform batch
        std::vector<double> vars, param, out;
        std::vector<fk_report> reps;
        int device = 0;
        bool done = false;  // solved by the heterogeneous launch
    };
    std::vector<Group> groups;
    std::multimap<uint64_t, size_t> by_sig;
    for (uint32_t i = 0; i < n; i++) {
        const fk_problem* p = problems[i];
        if (!p || !free_values[i]) return fail(FK_ERR_INVALID, "null problem in batch");
        if (!problem_arrays_ok(*p)) return fail(FK_ERR_INVALID, "null array in fk_problem");
        const uint64_t sig = structural_signature(*p);
        size_t gi = SIZE_MAX;
        auto range = by_sig.equal_range(sig);
        for (auto it = range.first; it != range.second; ++it) {
            const fk_problem& q = *problems[groups[it->second].members[0]];
            const fk_problem& pp = *p;
            if (q.n_vars == pp.n_vars && q.n_expr == pp.n_expr && q.n_free == pp.n_free && q.n_rows == pp.n_rows &&
                (!pp.n_expr || (!std::memcmp(q.kind, pp.kind, pp.n_expr) && !std::memcmp(q.idx, pp.idx, 16 * (size_t)pp.n_expr))) &&
                (!pp.n_free || !std::memcmp(q.free_vars, pp.free_vars, 4 * (size_t)pp.n_free)) &&
                (!pp.n_rows || !std::memcmp(q.rows, pp.rows, 4 * (size_t)pp.n_rows))) {
                gi = it->second;
                break;
            }
        }
        if (gi == SIZE_MAX) {
            groups.emplace_back();
            groups.back().sig = sig;
            gi = groups.size() - 1;
            by_sig.emplace(sig, gi);
        }
        groups[gi].members.push_back(i);
    }
    // ---- topologies: cache first, the misses are analysed on host threads ------------------------------------
    std::vector<size_t> missing;
    for (size_t g = 0; g < groups.size(); g++) {
        groups[g].topo = cache_lookup(groups[g].sig, *problems[groups[g].members[0]]);
        if (!groups[g].topo) missing.push_back(g);
    }
    if (!missing.empty()) {
        auto build_one = [&](size_t g) {
            fk_topology* t = nullptr;
            groups[g].rc = fk_topology_create(problems[groups[g].members[0]], &t);
            if (groups[g].rc != FK_OK) groups[g].err = g_error;
            else groups[g].topo.reset(t, fk_topology_destroy);
        };
        const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        const size_t workers = std::min<size_t>(missing.size(), hw);
        if (workers <= 1) {
            for (size_t g : missing) build_one(g);
        } else {
            std::atomic<size_t> next{0};
            std::vector<std::thread> pool;
            for (size_t w = 0; w < workers; w++)
                pool.emplace_back([&] {
                    for (size_t k = next++; k < missing.size(); k = next++) build_one(missing[k]);
                });
            for (auto& th : pool) th.join();
        }
        for (size_t g : missing) {
            if (groups[g].rc != FK_OK) return fail(groups[g].rc, groups[g].err);
            cache_insert(groups[g].sig, groups[g].topo);
        }
    }
    int cur = 0;
    if (cudaGetDevice(&cur) != cudaSuccess) cur = 0;
    // ---- solve.  With several devices and at least as many groups, whole groups go to devices round robin; otherwise
    // every large group is sharded by sketch over the devices (fk_batch_solve).
    const bool groups_to_devices = n_gpus > 1 && groups.size() >= (size_t)n_gpus;
    size_t rr = 0;
    int first_rc = FK_OK;
    std::string first_err;
    auto note = [&](int rc, const std::string& e) { if (rc != FK_OK && first_rc == FK_OK) { first_rc = rc; first_err = e; } };
    // Several small groups: ONE launch of the heterogeneous kernel for all of them (one warp per system) instead of one
    // launch, two copies and a synchronisation per topology.  Groups of 64 systems and more keep the uniform batch kernels
    // (from there the sketch-per-thread kernel is the faster one).  (FK_NO_HETERO=1: A/B knob.)
    {
        static const bool no_hetero = std::getenv("FK_NO_HETERO") != nullptr;
        std::vector<size_t> small;
        for (size_t g = 0; g < groups.size(); g++) {
            const fk::Topology& t = groups[g].topo->t;
            if (t.path == 0 && groups[g].members.size() < 64 && fk::lm_smem_doubles(t.n_free, t.n_rows, t.jac_nnz, (uint32_t)t.l_rowidx.size()) * 8 <= 48 * 1024)
                small.push_back(g);
        }
        if (!no_hetero && small.size() >= 2) {
            std::vector<fk_topology*> topos32;
            std::vector<HeteroItem> items;
            for (size_t q = 0; q < small.size(); q++) {
                Group& gr = groups[small[q]];
                fk_topology* t32 = gr.topo->lanes32();
                if (t32->t.tile != 32) continue;  // no 32-lane tables (FK_NO_LATENCY_TWIN): the group keeps the uniform kernel
                for (uint32_t i : gr.members) items.push_back({(uint32_t)topos32.size(), i});
                topos32.push_back(t32);
                gr.done = true;
            }
            if (!items.empty() && n_gpus > 1 && items.size() >= 64u * (size_t)n_gpus) {
                // several devices: contiguous shares of the systems, one host thread and one launch per device
                std::vector<int> rcs(n_gpus, FK_OK);
                std::vector<std::string> errs(n_gpus);
                std::vector<std::thread> pool;
                for (int g = 0; g < n_gpus; g++)
                    pool.emplace_back([&, g] {
                        const size_t lo = items.size() * g / n_gpus, hi = items.size() * (g + 1) / n_gpus;
                        std::vector<HeteroItem> part(items.begin() + lo, items.begin() + hi);
                        rcs[g] = hetero_solve(g, topos32, part, problems, free_values, reports);
                        if (rcs[g] != FK_OK) errs[g] = g_error;
                    });
                for (auto& th : pool) th.join();
                for (int g = 0; g < n_gpus; g++) note(rcs[g], errs[g]);
            } else if (!items.empty()) {
                const int rc = hetero_solve(cur, topos32, items, problems, free_values, reports);
                if (rc != FK_OK) note(rc, g_error);
            }
        }
    }
    for (Group& gr : groups) {
        if (gr.done) continue;
        const fk::Topology& t = gr.topo->t;
        const size_t cnt = gr.members.size();
        if (t.path == 2) continue;  // large systems below
        gr.vars.resize(cnt * t.n_vars); gr.param.resize(cnt * std::max<uint32_t>(t.n_expr, 1)); gr.out.resize(cnt * std::max<uint32_t>(t.n_free, 1));
        gr.reps.resize(cnt);
        for (size_t k = 0; k < cnt; k++) {
            const fk_problem* p = problems[gr.members[k]];
            if (t.n_vars) std::memcpy(&gr.vars[k * t.n_vars], p->vars, sizeof(double) * t.n_vars);
            for (uint32_t f = 0; f < t.n_free; f++) gr.vars[k * t.n_vars + t.free_vars[f]] = free_values[gr.members[k]][f];
            if (t.n_expr) {
                if (p->param) std::memcpy(&gr.param[k * t.n_expr], p->param, sizeof(double) * t.n_expr);
                else std::fill(gr.param.begin() + k * t.n_expr, gr.param.begin() + (k + 1) * t.n_expr, 0.0);
            }
        }
        gr.device = groups_to_devices ? (int)(rr++ % (size_t)n_gpus) : cur;
    }
    for (Group& gr : groups) {
        if (gr.done) continue;
        const fk::Topology& t = gr.topo->t;
        const size_t cnt = gr.members.size();
        if (t.path == 2) {  // large systems: one after the other through the global sparse path
            for (uint32_t i : gr.members) {
                int rc = fk_topology_lm_solve(gr.topo.get(), problems[i]->vars, problems[i]->param, free_values[i], reports ? &reports[i] : nullptr);
                if (rc != FK_OK) note(rc, g_error);
            }
            continue;
        }
        int rc = FK_OK;
        std::string e;
        if (groups_to_devices || n_gpus == 1 || small_fits(t, (uint32_t)cnt)) {
            rc = run_device_range(gr.topo.get(), gr.device, 0, (uint32_t)cnt, gr.vars.data(), gr.param.data(), gr.out.data(), gr.reps.data(), &e, 0);
        } else {
            rc = fk_batch_solve(gr.topo.get(), (uint32_t)cnt, gr.vars.data(), gr.param.data(), gr.out.data(), gr.reps.data(), n_gpus);
            if (rc != FK_OK) e = g_error;
        }
        if (rc != FK_OK) { note(rc, e); continue; }
        for (size_t k = 0; k < cnt; k++) {
            std::memcpy(free_values[gr.members[k]], &gr.out[k * t.n_free], sizeof(double) * t.n_free);
            if (reports) reports[gr.members[k]] = gr.reps[k];
        }
    }
    if (cudaSetDevice(cur) != cudaSuccess) cudaGetLastError();
    return first_rc == FK_OK ? FK_OK : fail(first_rc, first_err);
}

int fk_lm_solve(const fk_problem* problem, double* free_values, fk_report* report) {
    const fk_problem* ps[1] = {problem};
    double* fv[1] = {free_values};
    return fk_lm_solve_batch(1, ps, fv, report, 1);
}

}  // extern "C"
