// Host side of the engine API (include/dtalite_b200.h, level B): owns the device image of the network,
// the column pool and the shortest-path-tree batches, and sequences the kernels of one assignment
// iteration on one CUDA stream.  Replaces the stage-1/2 loops of network_assignment
// (reference main_api.cpp:589-759).  No CPU compute path exists here: every entry point launches
// kernels and fails when CUDA is unavailable.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <future>
#include <map>
#include <set>
#include <numeric>
#include <chrono>
#include <thread>
#include <queue>
#include <tuple>

#include "dtl_internal.h"

using namespace dtl;

namespace {

thread_local std::string g_err;
void set_err(const char* fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
}

#define CK(call)                                                                                           \
    do {                                                                                                   \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess) {                                                                           \
            set_err("CUDA error %s at %s:%d: %s", cudaGetErrorName(e_), __FILE__, __LINE__, cudaGetErrorString(e_)); \
            return -1;                                                                                     \
        }                                                                                                  \
    } while (0)

// ---- NCCL through dlopen: single-GPU users never need libnccl -------------------------------------
struct Nccl {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;
    ncclResult_t (*CommGetAsyncError)(ncclComm_t, ncclResult_t*) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool load()
    {
        if (h) return true;
        const char* names[] = { "libnccl.so.2", "libnccl.so" };
        for (const char* n : names) {
            h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (h) break;
        }
        if (!h) { set_err("cannot load libnccl.so.2: %s", dlerror()); return false; }
        GetUniqueId = (decltype(GetUniqueId))dlsym(h, "ncclGetUniqueId");
        CommInitRank = (decltype(CommInitRank))dlsym(h, "ncclCommInitRank");
        AllReduce = (decltype(AllReduce))dlsym(h, "ncclAllReduce");
        CommDestroy = (decltype(CommDestroy))dlsym(h, "ncclCommDestroy");
        GetErrorString = (decltype(GetErrorString))dlsym(h, "ncclGetErrorString");
        CommAbort = (decltype(CommAbort))dlsym(h, "ncclCommAbort");
        CommGetAsyncError = (decltype(CommGetAsyncError))dlsym(h, "ncclCommGetAsyncError");
        if (!GetUniqueId || !CommInitRank || !AllReduce || !CommDestroy) { set_err("libnccl lacks a required symbol"); return false; }
        return true;
    }
} g_nccl;

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
};

} // namespace

struct dtl_engine {
    int dev = 0, sm_count = 148;
    cudaStream_t st = nullptr;
    // side stream of the assignment loop: the column -> link-volume scatter of iteration k (+ its all-reduce) only feeds the VDF
    // pass of iteration k+1, so it runs beside the shortest-path search of iteration k, on the SMs the search leaves free
    cudaStream_t st2 = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool side_pending = false, overlap = true;
    // pipelined iterations (dtl_iterations): the back-trace of iteration k runs beside the search of iteration k+1, whose trees go to a
    // second set of buffers; expect_side / defer_helpers tell the launch code that side-stream work will follow the launch
    bool expect_side = false, defer_helpers = false, pipelined = false;
    cudaEvent_t ev_k1[2] = { nullptr, nullptr }, ev_vdf = nullptr, ev_done = nullptr;
    struct TreeSet {
        unsigned long long* d_bl = nullptr; size_t bl_cap = 0;
        int2* d_bp = nullptr; size_t bp_cap = 0;
        unsigned long long* d_labels = nullptr; int* d_pred = nullptr;
        int* d_tie_nodes = nullptr; int* d_batch_flag = nullptr;
    } alt;
    bool alt_ready = false;
    // stream memory operations of the driver API (cuStreamWaitValue32 / cuStreamWriteValue32): the gate of the pipelined loop
    void *fn_wait_value = nullptr, *fn_write_value = nullptr;
    unsigned* d_go = nullptr;  // [0..1] per tree set: tag of the iteration whose search raised no flag (0 = it did); [2..3]: tag of the
                               // iteration whose back-trace may start (the next search is resident, or will not come before a redo)
    unsigned pipe_tag = 0;
    unsigned* next_started_word = nullptr; // handed to the next origin-bundle launch (BundleLaunch::started_word)
    unsigned next_started_tag = 0;
    int N = 0, L = 0, Z = 0, AT = 1, T = 1, B = 1;
    int rank = 0, world = 1;
    dtl_options opt{};
    float total_demand = 0;
    size_t dev_bytes = 0;
    std::vector<void*> allocs;

    // host copies
    std::vector<int> h_link_from, h_link_to, h_link_tag, h_node_zone, h_zone_centroid;
    std::vector<float> h_vot;
    // tasks owned by this rank, canonical (at, tau, zone) order
    std::vector<int> task_at, task_tau, task_zone;
    std::vector<int> task_cell_ptr; // [n_tasks+1] into d_task_cells
    int *d_task_origin = nullptr, *d_task_zone = nullptr, *d_task_cells = nullptr;
    // OD cells of a batch ordered by path length for the back-trace (sort_cells_by_path_length), once per batch
    int *d_path_len = nullptr, *d_task_cell_ptr = nullptr;
    std::set<int> sorted_batches; // first task of the batches whose cells are sorted
    unsigned char* d_task_skip = nullptr;
    // OD cells owned by this rank
    std::vector<int> cell_global; // local -> global OD index
    int n_cells = 0;
    int *d_od_dest = nullptr, *d_od_at = nullptr, *d_od_tau = nullptr, *d_od_task = nullptr;
    double* d_od_vol = nullptr;

    std::vector<ForwardStar> fs; // [AT*T], index at*T+tau
    // sensitivity-analysis stage (dtl_sa_*)
    std::vector<int> sa_rt_info;           // [AT] real_time_information of the agent types
    std::vector<int> sa_tau, sa_link;      // supply-side lane changes
    std::vector<double> sa_lanes;
    bool sa_mode = false;                  // trees_and_columns only visits the agent types with real-time information
    unsigned char *d_sa_decay_at = nullptr, *d_od_impact = nullptr;
    // "dms" scenario rows (dtl_sa_set_dms): message signs, information zones, and what follows from them in the SA stage
    std::vector<int> dms_tau, dms_link, dms_zone;
    std::vector<char> info_zone;                 // [Z]
    unsigned char *d_sa_decay_cell = nullptr, *d_sa_task_skip = nullptr; // [n_cells] / [n_tasks], see sa_set_dms_impl
    std::vector<int> h_od_at, h_od_tau, h_od_task, h_od_dzone; // [n_cells] host copies (route modification)
    std::vector<unsigned char> h_task_skip;     // [n_tasks] origin has no physical out-link (routing.h:428)
    std::vector<unsigned char> sa_slot_impact;  // [n_cells * cap] impacted_path_flag the route modification set (1 impacted, 3 diverted new column)
    std::vector<unsigned char> sa_cell_info;    // [n_cells] CColumnVector::information_type (1: a detour was looked for in this cell's columns)
    signed char* d_design_flag = nullptr;
    int* d_rt_info = nullptr;
    std::vector<std::vector<double>> sa_hist_time, sa_hist_volume; // per SA column-pool iteration: path time / volume of every slot
    // ODME (dtl_odme_*)
    bool odme_ready = false;
    double *d_obs = nullptr, *d_est_dev = nullptr, *d_sensor_dev = nullptr;
    int *d_ub = nullptr, *d_sensor_idx = nullptr;
    float *d_at_pce = nullptr, *d_at_occ = nullptr;
    std::vector<int> odme_sensor_idx;      // tau * L + link of the links with a count, in (link, tau) order
    std::vector<double> odme_obs;          // [T*L]
    std::vector<int> odme_ub;              // [T*L]
    std::vector<ForwardStar> bstar; // [T] stars of the backward search from a destination (dtl_backward_trees), built on first use
    int* d_link_to = nullptr;
    int** d_link_edge_ptrs = nullptr;
    Edge** d_edges_ptrs = nullptr;
    int *d_link_tag = nullptr, *d_link_from = nullptr;
    VdfParams vp{};
    LinkState ls{};
    double *d_pce = nullptr, *d_occ = nullptr, *d_toll = nullptr, *d_lr = nullptr;
    double *d_pce_uni = nullptr, *d_occ_uni = nullptr; // [T*AT]
    double2* d_link_cost = nullptr;                    // [T*AT*L] K2b workspace
    float* d_vot = nullptr;
    double *d_block_sums = nullptr, *d_scalars = nullptr, *d_volume_tmp = nullptr;
    double* h_scalars = nullptr; // [16] scalar read-backs
    double h_scalars_buf[16] = { 0 };
    ColumnPool pool{};
    int* d_error_flag = nullptr;
    double *d_least = nullptr, *d_od_sums = nullptr;

    // tree batch
    int batch_tasks = 0;
    unsigned long long* d_labels = nullptr;
    int* d_pred = nullptr;
    int* d_tie_nodes = nullptr;
    std::vector<int> h_tie_nodes;
    // keep_trees: trees of the whole last iteration
    unsigned long long* d_keep_labels = nullptr;
    int* d_keep_pred = nullptr;
    bool trees_valid = false;
    // sssp workspace
    int sssp_block = 256, sssp_grid = 0;
    QEntry* d_ws = nullptr;
    size_t ws_stride = 0;
    int cap_near = 0, cap_far = 0, tie_cap = 4096;
    // worst-case queues for the few trees whose frontier outgrows the regular ones (re-run path)
    QEntry* d_ws_big = nullptr;
    size_t ws_big_stride = 0;
    int cap_near_big = 0, cap_far_big = 0, big_grid = 8;
    int* d_ws_big_tie = nullptr;
    int* d_rerun_list = nullptr;
    int* d_tie_cand = nullptr;
    int* d_task_counter = nullptr;
    SsspStats* d_stats = nullptr;
    unsigned long long* d_counters = nullptr; // [0] link refs, [1] path nodes
    int *d_replay_tasks = nullptr, *d_replay_status = nullptr, *d_replay_next = nullptr;
    int replay_cap = 0;
    // origin-bundle kernel (kernels_sssp_bundle.cu): node-major label rows, per-CTA state words and far piles, bundle plans
    std::vector<int> h_task_origin;
    unsigned long long* d_bl = nullptr;
    size_t bl_cap = 0;          // elements
    int2* d_bfar = nullptr;
    int bundle_grid = 0, bundle_cap_far = 0;
    int* d_bundle_counter = nullptr;
    struct BundlePlan {
        int B = 0, n_bundles = 0; int* d_tasks = nullptr; int* d_slots = nullptr;
        std::vector<int> h_tasks;                         // [n_bundles * B] batch-local task of every slot (-1 = empty)
        int *d_dest_ptr = nullptr, *d_dest_idx = nullptr; // destination-bounded search: node * B + slot of every destination with demand
    };
    // destination-bounded search (dtl_set_bounded_search): on by default -- searches whose trees are only back-traced stop at the last
    // destination with demand; off (or DTL_COMPLETE_SEARCH=1): every search builds the reference's complete tree
    bool bounded = true, count_pass = false, in_loop = false; // in_loop: the trees of this launch are only back-traced
    unsigned* d_bundle_bound = nullptr;
    size_t bundle_bound_cap = 0;
    std::vector<int> h_task_cells, h_od_dest;             // [n_cells] cells grouped by task (creation order) / destination centroid of a cell
    std::map<std::tuple<int, int, int, int>, BundlePlan> bundle_plans; // (forward star, first task, tasks, B)
    int* d_bundle_tmp = nullptr;
    size_t bundle_tmp_cap = 0;
    int* d_bundle_work = nullptr; // fused label -> predecessor pass: finished-bundle list and tile counters
    size_t bundle_work_cap = 0;
    int2* d_bp = nullptr;       // node-major predecessor records of a bundle batch (slim label -> tree pass)
    size_t bp_cap = 0;
    int* d_batch_flag = nullptr;
    int last_sssp_variant = -1;
    double counted_edges_cached = -1;
    std::thread embed_thread;        // landmark embeddings of the forward stars (bundle planner), computed behind dtl_create
    std::vector<float> h_embed_all;  // embedding of the whole network when the zone partition needed one (world_size > 1)

    // timing
    std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> pending;
    std::vector<cudaEvent_t> free_events;
    double ms[8] = { 0 };
    long long launches[8] = { 0 };
    double ctr_base[10] = { 0 };
    double tasks_done = 0, reruns = 0;

    ncclComm_t comm = nullptr;
    std::string comm_key;     // key of `comm` in the process-wide table (dtl_comm_init)
    bool local_shard = false; // world > 1 without a communicator: the caller reduces the shards (tests)

    // Device memory: allocations up to 32 MB are carved out of 256 MB slabs (an engine makes ~80 of them; one cudaMalloc each
    // costs more than the copies that fill them), larger ones get their own cudaMalloc.  Everything is freed in dtl_destroy.
    char* slab = nullptr;
    size_t slab_used = 0, slab_cap = 0;
    template <class T>
    int alloc(T** p, size_t n)
    {
        size_t bytes = std::max<size_t>(n, 1) * sizeof(T);
        if (bytes <= ((size_t)32 << 20)) {
            const size_t need = (bytes + 255) & ~(size_t)255;
            if (!slab || slab_used + need > slab_cap) {
                void* q = nullptr;
                const size_t cap = (size_t)256 << 20;
                cudaError_t e = cudaMalloc(&q, cap);
                if (e != cudaSuccess) {
                    set_err("cudaMalloc of %zu bytes failed: %s", cap, cudaGetErrorString(e));
                    return -1;
                }
                allocs.push_back(q);
                slab = (char*)q; slab_used = 0; slab_cap = cap;
            }
            *p = (T*)(slab + slab_used);
            slab_used += need;
            dev_bytes += need;
            return 0;
        }
        void* q = nullptr;
        cudaError_t e = cudaMalloc(&q, bytes);
        if (e != cudaSuccess) {
            set_err("cudaMalloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
            return -1;
        }
        allocs.push_back(q);
        dev_bytes += bytes;
        *p = (T*)q;
        return 0;
    }
    template <class T>
    int upload(T** p, const T* src, size_t n)
    {
        if (alloc(p, n)) return -1;
        if (n) CK(cudaMemcpy(*p, src, n * sizeof(T), cudaMemcpyHostToDevice));
        return 0;
    }
    cudaEvent_t get_event()
    {
        if (!free_events.empty()) { cudaEvent_t e = free_events.back(); free_events.pop_back(); return e; }
        cudaEvent_t e;
        cudaEventCreate(&e);
        return e;
    }
    int phase_begin(int cls, cudaStream_t on = nullptr)
    {
        cudaEvent_t a = get_event(), b = get_event();
        cudaEventRecord(a, on ? on : st);
        pending.push_back({ cls, { a, b } });
        return (int)pending.size() - 1;
    }
    void phase_end(int handle, int n_launches = 1, cudaStream_t on = nullptr)
    {
        cudaEventRecord(pending[handle].second.second, on ? on : st);
        launches[pending[handle].first] += n_launches;
    }
    // the main stream waits for what was put on the side stream (no-op when nothing is pending)
    void join_side()
    {
        if (!side_pending) return;
        cudaStreamWaitEvent(st, ev_join, 0);
        side_pending = false;
    }
    void flip_trees()
    {
        std::swap(d_bl, alt.d_bl); std::swap(bl_cap, alt.bl_cap);
        std::swap(d_bp, alt.d_bp); std::swap(bp_cap, alt.bp_cap);
        std::swap(d_labels, alt.d_labels); std::swap(d_pred, alt.d_pred);
        std::swap(d_tie_nodes, alt.d_tie_nodes); std::swap(d_batch_flag, alt.d_batch_flag);
    }
    void resolve_timing()
    {
        static const bool gap_trace = getenv("DTL_GAP_TRACE") != nullptr; // idle time of the main stream between two searches
        if (gap_trace) {
            const std::pair<cudaEvent_t, cudaEvent_t>* prev = nullptr;
            double gap = 0, vdf = 0;
            int n = 0;
            for (auto& p : pending) {
                if (p.first == 0 && prev) { float t = 0; if (cudaEventElapsedTime(&t, p.second.first, p.second.second) == cudaSuccess) vdf += t; }
                if (p.first != 2) continue;
                float t = 0;
                if (prev && cudaEventElapsedTime(&t, prev->second, p.second.first) == cudaSuccess) { gap += t; ++n; }
                prev = &p.second;
            }
            if (n) fprintf(stderr, "main stream between two searches: %.3f ms on average over %d gaps (of which VDF pass %.3f ms)\n", gap / n, n, vdf / n);
        }
        for (auto& p : pending) {
            float t = 0;
            if (cudaEventElapsedTime(&t, p.second.first, p.second.second) == cudaSuccess) ms[p.first] += t;
            free_events.push_back(p.second.first);
            free_events.push_back(p.second.second);
        }
        pending.clear();
    }
};

namespace {

int fs_index(const dtl_engine* e, int at, int tau) { return at * e->T + tau; }

// Work partition of the reference (g_assign_computing_tasks_to_memory_blocks, main_api.cpp:201-275, and the
// loop nest main_api.cpp:667-697): which (agent type, period, zone) triples get a shortest-path tree.
void build_task_list(const dtl_problem* p, std::vector<int>& at_out, std::vector<int>& tau_out, std::vector<int>& zone_out)
{
    const int Z = p->n_zones, AT = p->n_agent_types, T = p->n_periods, B = std::max(1, p->memory_blocks);
    struct Net { int at, tau; std::vector<int> zones; };
    std::vector<Net> nets;
    for (int at = 0; at < AT; ++at)
        for (int tau = 0; tau < T; ++tau) {
            const size_t base = nets.size();
            for (int z = 0; z < Z; ++z) {
                if (z < B) nets.push_back(Net{ at, tau, { z } });                        // one network per block
                else if (p->origin_demand[z] > 0.001) nets[base + z % B].zones.push_back(z); // later zones only with demand
            }
        }
    const int n_nets = (int)nets.size();
    const int n_blocks = std::min(n_nets, B);
    std::vector<unsigned char> visited((size_t)AT * T * Z + 1, 0);
    for (int pid = 0; pid < n_blocks; ++pid)
        for (int blk = 0; blk < AT * T; ++blk) {
            const int ni = blk * B + pid;
            if (ni >= n_nets) continue;
            for (int z : nets[ni].zones) visited[((size_t)nets[ni].at * T + nets[ni].tau) * Z + z] = 1;
        }
    for (int at = 0; at < AT; ++at)
        for (int tau = 0; tau < T; ++tau)
            for (int z = 0; z < Z; ++z)
                if (visited[((size_t)at * T + tau) * Z + z]) { at_out.push_back(at); tau_out.push_back(tau); zone_out.push_back(z); }
}

// Host image of one forward star, built without any CUDA call (a worker thread builds it while the link arrays are uploaded).
struct FsHost {
    std::vector<int> row, link_edge, in_ptr;
    std::vector<Edge> edges;
    std::vector<int2> in_edges, rows, rin4;
    std::vector<int4> link_info;
};

// The star of a search direction: `tail` / `head` are link_from / link_to for the forward search (routing.h:257-352) and
// swapped for the backward search from a destination (routing.h:1009-1273: the in-links of a node in insertion order);
// `allowed` may be null (every link), `tag` may be null (no zone-tagged links in this star); `fftt` seeds the label window.
void build_star_host(dtl_engine* e, ForwardStar& f, const int* tail, const int* head, const int* tag, const unsigned char* allowed,
                     const double* fftt, FsHost& h)
{
    const int N = e->N, L = e->L;
    std::vector<int>& row = h.row;
    std::vector<int>& link_edge = h.link_edge;
    row.assign(N + 1, 0); link_edge.assign(L, -1);
    for (int l = 0; l < L; ++l)
        if (!allowed || allowed[l]) row[tail[l] + 1]++;
    for (int i = 0; i < N; ++i) row[i + 1] += row[i];
    const int E = row[N];
    std::vector<Edge>& edges = h.edges;
    edges.resize(std::max(E, 1));
    std::vector<int> cur(row.begin(), row.end() - 1);
    double cost_sum = 0;
    long long cost_n = 0;
    for (int l = 0; l < L; ++l) // out-links of a node in insertion order == ascending link seq no
        if (!allowed || allowed[l]) {
            const int ei = cur[tail[l]]++;
            edges[ei].to = head[l] | (tag && tag[l] >= 0 ? 0x80000000 : 0);
            edges[ei].link = l;
            edges[ei].cost = 0.0;
            link_edge[l] = ei;
            const double c = fftt[l];
            if (c > 1e-3) { cost_sum += c; ++cost_n; }
        }
    std::vector<int>& in_ptr = h.in_ptr;
    in_ptr.assign(N + 1, 0);
    for (int ei = 0; ei < E; ++ei) in_ptr[(edges[ei].to & 0x7fffffff) + 1]++;
    for (int i = 0; i < N; ++i) in_ptr[i + 1] += in_ptr[i];
    std::vector<int2>& in_edges = h.in_edges;
    in_edges.resize(std::max(E, 1));
    std::vector<int> icur(in_ptr.begin(), in_ptr.end() - 1);
    for (int u = 0; u < N; ++u)
        for (int ei = row[u]; ei < row[u + 1]; ++ei) in_edges[icur[edges[ei].to & 0x7fffffff]++] = make_int2(u, ei);
    f.n_edges = E;
    f.h_row_ptr = row;
    f.h_to.resize(E);
    for (int ei = 0; ei < E; ++ei) f.h_to[ei] = edges[ei].to & 0x7fffffff;
    const double factor = e->opt.sssp_delta_factor > 0 ? e->opt.sssp_delta_factor : 8.0;
    f.delta = factor * (cost_n ? cost_sum / (double)cost_n : 1.0);
    std::vector<int2>& rows = h.rows;
    rows.resize(std::max(N, 1));
    for (int i = 0; i < N; ++i) {
        // a node whose out-edges are all zone-tagged connectors (a centroid) is expanded only as the origin of its own
        // tree (routing.h:779-786); the kernel reads the origin's row from row_ptr, so every other lookup sees an empty row
        bool all_tagged = row[i + 1] > row[i];
        for (int ei = row[i]; ei < row[i + 1] && all_tagged; ++ei) all_tagged = edges[ei].to < 0;
        rows[i] = all_tagged ? make_int2(row[i], row[i]) : make_int2(row[i], row[i + 1]);
        for (int ei = row[i]; ei < row[i + 1] && !all_tagged; ++ei) f.open_tags = f.open_tags || edges[ei].to < 0;
    }
    h.link_info.resize(std::max(L, 1));
    for (int l = 0; l < L; ++l) {
        const int2 r = rows[head[l]];
        h.link_info[l] = make_int4(tail[l], r.x, r.y, 0);
    }
    h.rin4.assign((size_t)std::max(N, 1) * 4, make_int2(-1, 0));
    for (int v = 0; v < N; ++v) {
        const int n_in = in_ptr[v + 1] - in_ptr[v];
        for (int q = 0; q < std::min(n_in, 4); ++q) h.rin4[(size_t)v * 4 + q] = in_edges[in_ptr[v] + q];
        if (n_in > 4) h.rin4[(size_t)v * 4 + 3].y |= 0x40000000;
    }
}

void build_forward_star_host(dtl_engine* e, const dtl_problem* p, int at, int tau, FsHost& h)
{
    build_star_host(e, e->fs[fs_index(e, at, tau)], p->link_from, p->link_to, p->link_tag, p->allowed + ((size_t)tau * e->AT + at) * e->L,
                    p->fftt + (size_t)tau * e->L, h);
}

int upload_star(dtl_engine* e, ForwardStar& f, const FsHost& h)
{
    if (e->upload(&f.row_ptr, h.row.data(), h.row.size())) return -1;
    if (e->upload(&f.rows, h.rows.data(), h.rows.size())) return -1;
    if (e->upload(&f.link_info, h.link_info.data(), h.link_info.size())) return -1;
    if (e->upload(&f.edges, h.edges.data(), h.edges.size())) return -1;
    if (e->upload(&f.in_ptr, h.in_ptr.data(), h.in_ptr.size())) return -1;
    if (e->upload(&f.in_edges, h.in_edges.data(), h.in_edges.size())) return -1;
    if (e->upload(&f.rin4, h.rin4.data(), h.rin4.size())) return -1;
    if (e->upload(&f.link_edge, h.link_edge.data(), h.link_edge.size())) return -1;
    return 0;
}

int upload_forward_star(dtl_engine* e, int at, int tau, const FsHost& h) { return upload_star(e, e->fs[fs_index(e, at, tau)], h); }

// ---- multi-GPU failure handling -------------------------------------------------------------------------------------------
// Raw NCCL has no watchdog: a rank that leaves the iteration early (a CUDA error, an exhausted column arena) would leave its
// peers blocked in their next collective for ever.  Three measures: (1) the soft error flags are all-reduced (max) at the end
// of every iteration, so every rank stops at the same point with the same message; (2) a rank that fails hard aborts the
// communicator on its way out; (3) every wait on the stream of an engine with a communicator polls, watches
// ncclCommGetAsyncError and gives up -- aborting the communicator -- after DTL_NCCL_TIMEOUT_S seconds (default 120).
std::map<std::string, ncclComm_t>& comm_table()
{
    static std::map<std::string, ncclComm_t> comms;
    return comms;
}

void comm_abort(dtl_engine* e)
{
    if (!e->comm) return;
    if (g_nccl.CommAbort) g_nccl.CommAbort(e->comm);
    comm_table().erase(e->comm_key);
    e->comm = nullptr;
    e->local_shard = false; // later collectives of this engine fail with "dtl_comm_init was not called" instead of hanging
}

int sync_stream(dtl_engine* e)
{
    if (!e->comm) {
        CK(cudaStreamSynchronize(e->st));
        return 0;
    }
    static const double limit_s = getenv("DTL_NCCL_TIMEOUT_S") ? atof(getenv("DTL_NCCL_TIMEOUT_S")) : 120.0;
    const auto t0 = std::chrono::steady_clock::now();
    for (long spins = 0;; ++spins) {
        const cudaError_t q = cudaStreamQuery(e->st);
        if (q == cudaSuccess) return 0;
        if (q != cudaErrorNotReady) {
            set_err("CUDA error %s while waiting for the stream: %s", cudaGetErrorName(q), cudaGetErrorString(q));
            comm_abort(e);
            return -1;
        }
        if ((spins & 1023) == 1023) {
            ncclResult_t ar = ncclSuccess;
            if (g_nccl.CommGetAsyncError && g_nccl.CommGetAsyncError(e->comm, &ar) == ncclSuccess && ar != ncclSuccess && ar != ncclInProgress) {
                set_err("NCCL reported an asynchronous error (%s): a peer rank failed", g_nccl.GetErrorString ? g_nccl.GetErrorString(ar) : "?");
                comm_abort(e);
                return -1;
            }
            const double waited = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (waited > limit_s) {
                set_err("a collective did not complete within %.0f s (DTL_NCCL_TIMEOUT_S): a peer rank left the iteration", limit_s);
                comm_abort(e);
                return -1;
            }
            if (waited > 0.002) std::this_thread::sleep_for(std::chrono::microseconds(50));
        }
    }
}

int all_reduce(dtl_engine* e, void* buf, size_t count, ncclDataType_t dt, ncclRedOp_t op = ncclSum, cudaStream_t on = nullptr)
{
    if (e->world <= 1 || e->local_shard) return 0;
    if (!e->comm) { set_err("world_size > 1 but dtl_comm_init was not called (or the communicator was aborted after a failure)"); return -1; }
    if (!on) on = e->st;
    const int ph = e->phase_begin(6, on);
    ncclResult_t r = g_nccl.AllReduce(buf, buf, count, dt, op, e->comm, on);
    e->phase_end(ph, 1, on);
    if (r != ncclSuccess) { set_err("ncclAllReduce failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"); return -1; }
    return 0;
}

int do_vdf(dtl_engine* e, bool detail, bool write_cost, const double* volume_override)
{
    VdfLaunch a{};
    a.T = e->T; a.L = e->L; a.AT = e->AT;
    a.p = e->vp; a.s = e->ls;
    a.toll = e->d_toll; a.lr_price = e->d_lr; a.vot = e->d_vot;
    a.link_edge = e->d_link_edge_ptrs; a.edges = e->d_edges_ptrs;
    a.block_sums = e->d_block_sums; a.sys_tt = e->d_scalars + 0;
    a.volume_override = volume_override;
    a.detail = detail; a.write_cost = write_cost; a.error_flag = e->d_error_flag;
    e->join_side();
    const int ph = e->phase_begin(0);
    launch_vdf(a, e->st);
    e->phase_end(ph, 2);
    CK(cudaGetLastError());
    return 0;
}

// side: on the side stream, after everything the main stream holds so far; the main stream picks the result up with join_side()
// (after != null: after that event only -- the pipelined loop, whose main stream already holds the next search)
int do_scatter(dtl_engine* e, int decay_k, bool person, bool sa = false, bool side = false, cudaEvent_t after = nullptr)
{
    const size_t TL = (size_t)e->T * e->L;
    if (!after) e->join_side();
    cudaStream_t on = e->st;
    if (side && e->overlap && e->st2) {
        on = e->st2;
        if (after) {
            CK(cudaStreamWaitEvent(on, after, 0));
        } else {
            CK(cudaEventRecord(e->ev_fork, e->st));
            CK(cudaStreamWaitEvent(on, e->ev_fork, 0));
        }
    }
    const int ph = e->phase_begin(1, on);
    CK(cudaMemsetAsync(e->ls.pce_fix, 0, TL * sizeof(unsigned long long), on));
    if (person) CK(cudaMemsetAsync(e->ls.person_fix, 0, TL * (1 + e->AT) * sizeof(unsigned long long), on));
    ScatterLaunch a{};
    a.L = e->L; a.T = e->T; a.AT = e->AT; a.pool = e->pool;
    a.od_at = e->d_od_at; a.od_tau = e->d_od_tau; a.pce = e->d_pce; a.occ = e->d_occ;
    a.pce_uni = e->d_pce_uni; a.occ_uni = e->d_occ_uni;
    a.pce_fix = e->ls.pce_fix; a.person_fix = person ? e->ls.person_fix : nullptr;
    a.decay_k = decay_k; a.link_refs = e->d_counters + 0;
    a.decay_at = sa ? e->d_sa_decay_at : nullptr;
    a.decay_cell = sa ? e->d_sa_decay_cell : nullptr;
    launch_scatter(a, on);
    e->phase_end(ph, 1, on);
    CK(cudaGetLastError());
    if (all_reduce(e, e->ls.pce_fix, TL, ncclUint64, ncclSum, on)) return -1;
    if (person && all_reduce(e, e->ls.person_fix, TL * (1 + e->AT), ncclUint64, ncclSum, on)) return -1;
    if (on != e->st) {
        CK(cudaEventRecord(e->ev_join, on));
        e->side_pending = true;
    }
    return 0;
}


// ---- origin bundles (kernels_sssp_bundle.cu) ------------------------------------------------------------------------
// The bundle kernel shares one search between B origins; it pays off when the origins of a bundle are close to each other
// in the network.  Origins are grouped by a k-d split of a two-dimensional landmark embedding: hop distances from two
// far-apart landmark nodes (host breadth-first searches, once per forward star, a few milliseconds for 100 000 nodes).
static void host_bfs(const ForwardStar& f, int N, int src, std::vector<float>& dist)
{
    dist.assign(N, INFINITY);
    std::vector<int> q;
    q.reserve(N);
    dist[src] = 0.f;
    q.push_back(src);
    for (size_t h = 0; h < q.size(); ++h) {
        const int u = q[h];
        for (int ei = f.h_row_ptr[u]; ei < f.h_row_ptr[u + 1]; ++ei)
            if (dist[f.h_to[ei]] == INFINITY) { dist[f.h_to[ei]] = dist[u] + 1.f; q.push_back(f.h_to[ei]); }
    }
}

static int farthest(const std::vector<float>& d)
{
    int best = -1;
    for (int i = 0; i < (int)d.size(); ++i)
        if (std::isfinite(d[i]) && (best < 0 || d[i] > d[best])) best = i;
    return best;
}

static void ensure_embedding(const ForwardStar& f, int N)
{
    if (!f.h_embed.empty()) return;
    f.h_embed.assign((size_t)2 * N, 0.f);
    int s0 = -1;
    for (int i = 0; i < N && s0 < 0; ++i)
        if (f.h_row_ptr[i + 1] > f.h_row_ptr[i]) s0 = i;
    if (s0 < 0) return;
    std::vector<float> d0, dA, dC, dD;
    host_bfs(f, N, s0, d0);
    const int A = farthest(d0);
    host_bfs(f, N, A, dA);
    const int C = farthest(dA);
    host_bfs(f, N, C, dC);
    int D = -1; // far from both A and C: the third corner
    float best = -1.f;
    for (int i = 0; i < N; ++i)
        if (std::isfinite(dA[i]) && std::isfinite(dC[i]) && std::min(dA[i], dC[i]) > best) { best = std::min(dA[i], dC[i]); D = i; }
    if (D < 0) D = C;
    host_bfs(f, N, D, dD);
    float big = 1.f;
    for (int i = 0; i < N; ++i) {
        if (std::isfinite(dA[i])) big = std::max(big, dA[i]);
        if (std::isfinite(dD[i])) big = std::max(big, dD[i]);
    }
    for (int i = 0; i < N; ++i) {
        f.h_embed[2 * i] = std::isfinite(dA[i]) ? dA[i] : 2.f * big;
        f.h_embed[2 * i + 1] = std::isfinite(dD[i]) ? dD[i] : 2.f * big;
    }
}

// `per` <= B origins per bundle (a bundle still occupies B label slots per node; the slots beyond `per` stay empty)
static void kd_split(std::vector<int>& idx, int lo, int hi, int B, int per, const float* embed, const int* origin, std::vector<int>& out)
{
    const int n = hi - lo;
    if (n <= per) {
        for (int i = 0; i < B; ++i) out.push_back(i < n ? idx[lo + i] : -1);
        return;
    }
    const int nb = (n + per - 1) / per;
    const int left = (nb / 2) * per;
    float mn[2] = { INFINITY, INFINITY }, mx[2] = { -INFINITY, -INFINITY };
    for (int i = lo; i < hi; ++i)
        for (int d = 0; d < 2; ++d) {
            const float x = embed[2 * origin[idx[i]] + d];
            mn[d] = std::min(mn[d], x); mx[d] = std::max(mx[d], x);
        }
    const int dim = (mx[1] - mn[1]) > (mx[0] - mn[0]) ? 1 : 0;
    std::nth_element(idx.begin() + lo, idx.begin() + lo + left, idx.begin() + hi, [&](int x, int y) {
        const float cx = embed[2 * origin[x] + dim], cy = embed[2 * origin[y] + dim];
        return cx < cy || (cx == cy && x < y);
    });
    kd_split(idx, lo, lo + left, B, per, embed, origin, out);
    kd_split(idx, lo + left, hi, B, per, embed, origin, out);
}

// Origin zones -> ranks.  The reference deals zones to its memory blocks round-robin (zone % blocks, main_api.cpp:259); any
// partition gives the same link volumes here (integer sums), so the zones of a rank are chosen to be NEIGHBOURS instead: a
// recursive median split of the landmark embedding, parts equal to within one zone.  Neighbouring origins are what the
// origin-bundle search shares work between (9.1 vs 9.8 ms per 1 000-tree shard, 7.4 vs 10.9 ms per 250-tree shard measured
// in round 1).  DTL_SHARD=mod selects the round-robin rule.  Every rank computes the same partition from the same problem.
static void split_zones(std::vector<int>& idx, int lo, int hi, int r0, int r1, const float* embed, const int* centroid, int* owner)
{
    if (r1 - r0 <= 1) {
        for (int i = lo; i < hi; ++i) owner[idx[i]] = r0;
        return;
    }
    const int rl = (r1 - r0) / 2;
    const int left = (int)((long long)(hi - lo) * rl / (r1 - r0));
    float mn[2] = { INFINITY, INFINITY }, mx[2] = { -INFINITY, -INFINITY };
    for (int i = lo; i < hi; ++i)
        for (int d = 0; d < 2; ++d) {
            const float x = embed[2 * centroid[idx[i]] + d];
            mn[d] = std::min(mn[d], x); mx[d] = std::max(mx[d], x);
        }
    const int dim = (mx[1] - mn[1]) > (mx[0] - mn[0]) ? 1 : 0;
    std::nth_element(idx.begin() + lo, idx.begin() + lo + left, idx.begin() + hi, [&](int x, int y) {
        const float cx = embed[2 * centroid[x] + dim], cy = embed[2 * centroid[y] + dim];
        return cx < cy || (cx == cy && x < y);
    });
    split_zones(idx, lo, lo + left, r0, r0 + rl, embed, centroid, owner);
    split_zones(idx, lo + left, hi, r0 + rl, r1, embed, centroid, owner);
}

static std::vector<int> zone_owners(const dtl_problem* p, int world, std::vector<float>* embed_out = nullptr)
{
    const int Z = p->n_zones, N = p->n_nodes, L = p->n_links;
    std::vector<int> owner(Z, 0);
    if (world <= 1) return owner;
    const char* rule = getenv("DTL_SHARD");
    if (rule && !strcmp(rule, "mod")) {
        for (int z = 0; z < Z; ++z) owner[z] = z % world;
        return owner;
    }
    ForwardStar f; // every link, whatever the agent type: only the shape of the network matters here
    f.h_row_ptr.assign(N + 1, 0);
    for (int l = 0; l < L; ++l) f.h_row_ptr[p->link_from[l] + 1]++;
    for (int i = 0; i < N; ++i) f.h_row_ptr[i + 1] += f.h_row_ptr[i];
    f.h_to.resize(L);
    std::vector<int> cur(f.h_row_ptr.begin(), f.h_row_ptr.end() - 1);
    for (int l = 0; l < L; ++l) f.h_to[cur[p->link_from[l]]++] = p->link_to[l];
    ensure_embedding(f, N);
    std::vector<int> idx; // zones that can carry trees first, the others are dealt with them (they cost nothing)
    for (int z = 0; z < Z; ++z) idx.push_back(z);
    split_zones(idx, 0, Z, 0, world, f.h_embed.data(), p->zone_centroid, owner.data());
    if (embed_out) *embed_out = std::move(f.h_embed);
    return owner;
}

// bundles of the batch [0, n_tasks): returns a device array [n_bundles][B] of batch-local task indices
static int plan_bundles(dtl_engine* e, const ForwardStar& f, int n_tasks, const int* h_origin, int plan_key, int B,
                        const int** d_tasks, const int** d_slots, int* n_bundles)
{
    const int fsi = (int)(&f - e->fs.data());
    const auto key = std::make_tuple(fsi, plan_key, n_tasks, B);
    if (plan_key >= 0) {
        auto it = e->bundle_plans.find(key);
        if (it != e->bundle_plans.end()) {
            *d_tasks = it->second.d_tasks; *d_slots = it->second.d_slots; *n_bundles = it->second.n_bundles;
            return 0;
        }
    }
    if (e->embed_thread.joinable()) e->embed_thread.join();
    ensure_embedding(f, e->N);
    std::vector<int> idx(n_tasks), out;
    std::iota(idx.begin(), idx.end(), 0);
    if (getenv("DTL_BUNDLE_NO_CLUSTER")) { // experiment hook: bundles of consecutive tasks
        for (int i = 0; i < (n_tasks + B - 1) / B * B; ++i) out.push_back(i < n_tasks ? i : -1);
    } else {
        // full bundles: 143 bundles of 14 origins (one per SM) instead of 125 of 16 measured 8.1 against 7.7 ms per 2 000 trees --
        // more expansions in total, and a bundle does not get faster with fewer origins
        kd_split(idx, 0, n_tasks, B, B, f.h_embed.data(), h_origin, out);
    }
    const int nb = (int)out.size() / B;
    const size_t n_out = out.size();
    out.resize(n_out + (size_t)std::max(1, n_tasks), -1); // followed by the inverse map: task -> bundle * B + slot
    for (size_t i = 0; i < n_out; ++i)
        if (out[i] >= 0) out[n_out + out[i]] = (int)i;
    int* d = nullptr;
    if (plan_key >= 0) {
        if (e->alloc(&d, out.size())) return -1;
        dtl_engine::BundlePlan pl;
        pl.B = B; pl.n_bundles = nb; pl.d_tasks = d; pl.d_slots = d + n_out;
        pl.h_tasks.assign(out.begin(), out.begin() + n_out);
        e->bundle_plans[key] = pl;
    } else {
        if (e->bundle_tmp_cap < out.size()) {
            if (e->alloc(&e->d_bundle_tmp, out.size() * 2)) return -1;
            e->bundle_tmp_cap = out.size() * 2;
        }
        d = e->d_bundle_tmp;
    }
    CK(cudaMemcpyAsync(d, out.data(), sizeof(int) * out.size(), cudaMemcpyHostToDevice, e->st));
    if (sync_stream(e)) return -1; // `out` is pageable and local
    *d_tasks = d; *d_slots = d + n_out; *n_bundles = nb;
    return 0;
}

static int make_bundle_launch(dtl_engine* e, const ForwardStar& f, const SsspLaunch& a, const int* h_origin, int plan_key, bool slim,
                              BundleLaunch& b, const int** d_slots)
{
    const int n = a.n_tasks;
    int B = 16;
    if (n < 8 * e->sm_count) B = 8;     // fewer than 8 trees per SM: bundles of 8 so that every SM still gets one
    if (n < 2 * e->sm_count) B = 4;     // (only when the kernel is forced: the round-synchronous search)
    if (const char* env = getenv("DTL_BUNDLE_B")) B = atoi(env) == 4 ? 4 : atoi(env) == 8 ? 8 : 16;
    b.s = a;
    b.B = B;
    b.tagged = f.open_tags || getenv("DTL_BUNDLE_TAGS") ? 1 : 0; // the env hook forces the per-edge tag check (tests)
    b.lpt = 2; // origins per thread; 4 measured slower (fewer warps to hide the round trips)
    if (const char* env = getenv("DTL_BUNDLE_LPT")) b.lpt = atoi(env) == 4 && B >= 8 ? 4 : 2;
    const auto plan_t0 = std::chrono::steady_clock::now();
    if (plan_bundles(e, f, n, h_origin, plan_key, B, &b.bundle_tasks, d_slots, &b.n_bundles)) return -1;
    if (getenv("DTL_CREATE_TRACE"))
        fprintf(stderr, "bundle plan (%d trees, bundles of %d): %.3f ms\n", n, B,
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - plan_t0).count());
    b.block = 1024;
    if (const char* env = getenv("DTL_BUNDLE_BLOCK")) b.block = atoi(env) == 512 ? 512 : atoi(env) == 768 ? 768 : 1024;
    // shared memory of a CTA: two queue bits per node, the rest for the two near queues (12 bytes per entry)
    int ctas_per_sm = b.block <= 512 && b.lpt <= 2 ? 2 : 1;
    const size_t smem_cta = (size_t)220 * 1024 / ctas_per_sm;
    const size_t fixed = bundle_smem_bytes(e->N, 0);
    b.cap_near = fixed + 24 * 64 <= smem_cta ? (int)std::min<size_t>(8192, (smem_cta - fixed) / 24) : 0;
    if (b.cap_near == 0) { set_err("origin-bundle kernel: %d nodes do not fit the shared-memory queue bits", e->N); return -1; }
    if (const char* env = getenv("DTL_SSSP_QUEUE_CAP")) b.cap_near = std::min(b.cap_near, std::max(8, atoi(env))); // test hook for the re-run path
    // B >= 8: the search without round barriers inside a label window (ring-buffer near queue, kernels_sssp_bundle_async.cu) is
    // 20 % faster than the round-synchronous one at the bench configuration; DTL_BUNDLE_ASYNC=0 selects the latter
    b.async = 0;
    const char* async_env = getenv("DTL_BUNDLE_ASYNC");
    if (B >= 8 && !(async_env && atoi(async_env) == 0) && !getenv("DTL_BUNDLE_BLOCK") && !getenv("DTL_BUNDLE_LPT")) {
        b.async = 1;
        b.block = 768; b.lpt = 2; // 768 threads x 80 registers: no spills; 8.3 against 8.7 ms per 2 000 trees with 1 024 x 64
        ctas_per_sm = 1;
        if (const char* env = getenv("DTL_BUNDLE_ASYNC_BLOCK")) { // experiment hook: "<threads>" or "<threads>x<CTAs per SM>"
            int t = 1024, c = 1;
            sscanf(env, "%dx%d", &t, &c);
            const int ok[6][2] = { { 1024, 1 }, { 896, 1 }, { 832, 1 }, { 768, 1 }, { 512, 1 }, { 512, 2 } };
            for (const auto& o2 : ok)
                if (o2[0] == t && o2[1] == c) { b.block = t; ctas_per_sm = c; }
        }
        int ring = 1;
        while ((size_t)2 * ring * 12 + fixed <= (size_t)226 * 1024 / ctas_per_sm - 1024 * (ctas_per_sm > 1) && 2 * ring <= 16384) ring *= 2;
        if (ring < 1024) { set_err("origin-bundle kernel: %d nodes leave no room for the shared-memory ring", e->N); return -1; }
        b.cap_near = ring; // far more slots than the 128 tickets the groups of a CTA can hold: a ticket never meets the slot of an older lap
        b.near_limit = ring;
        if (const char* env = getenv("DTL_SSSP_QUEUE_CAP")) b.near_limit = std::min(ring, std::max(8, atoi(env))); // test hook for the re-run path
        // windows of 10.4 mean link costs (measured per 2 000 trees: 5 / 6 / 7 / 8 / 9 / 10.5 / 12 / 14 / 16 mean link costs ->
        // 8.17 / 8.04 / 7.99 / 7.93 / 7.79 / 7.80 / 7.87 / 7.99 / 8.15 ms: wider windows re-expand more nodes, narrower ones drain more tails)
        if (!(e->opt.sssp_delta_factor > 0)) b.s.delta *= 1.3;
    }
    const int max_grid = e->sm_count * ctas_per_sm;
    b.ctas_per_sm = ctas_per_sm;
    if (!e->d_bundle_counter && e->alloc(&e->d_bundle_counter, 1)) return -1;
    if (e->bundle_grid < std::min(max_grid, b.n_bundles)) { // far piles of every resident CTA (an earlier, smaller buffer stays until dtl_destroy)
        e->bundle_grid = max_grid;
        e->bundle_cap_far = 2 * e->N + 1024;
        if (e->alloc(&e->d_bfar, (size_t)e->bundle_grid * 2 * e->bundle_cap_far)) return -1;
    }
    const size_t need = (size_t)b.n_bundles * B * e->N;
    if (e->bl_cap < need) {
        if (e->alloc(&e->d_bl, need)) return -1; // the previous, smaller buffer stays allocated until dtl_destroy
        e->bl_cap = need;
    }
    b.slim = slim ? 1 : 0;
    if (slim) {
        if (e->bp_cap < need) {
            if (e->alloc(&e->d_bp, need)) return -1;
            e->bp_cap = need;
        }
        if (!e->d_batch_flag && e->alloc(&e->d_batch_flag, 1)) return -1;
    }
    b.fuse_pred = slim && b.async && !getenv("DTL_NO_FUSED_PRED") ? 1 : 0;
    if (b.fuse_pred) {
        const size_t n_work = 2 + 2 * (size_t)b.n_bundles;
        if (e->bundle_work_cap < n_work) {
            if (e->alloc(&e->d_bundle_work, n_work)) return -1;
            e->bundle_work_cap = n_work;
        }
        b.work = e->d_bundle_work;
    }
    b.bp = e->d_bp; b.batch_flag = e->d_batch_flag;
    b.bl = e->d_bl; b.far = e->d_bfar; b.cap_far = e->bundle_cap_far;
    if (getenv("DTL_CREATE_TRACE"))
        fprintf(stderr, "bundle launch set up (plan + buffers): %.3f ms\n",
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - plan_t0).count());
    b.dest_ptr = nullptr; b.dest_idx = nullptr; b.bundle_bound = nullptr;
    if (e->bounded && !e->count_pass && b.fuse_pred && plan_key >= 0 && b.block == 768 && ctas_per_sm == 1) {
        // destination-bounded search: the (node, slot) of every destination the origins of a bundle have demand to, per bundle
        const int fsi = (int)(&f - e->fs.data());
        auto it = e->bundle_plans.find(std::make_tuple(fsi, plan_key, n, B));
        if (it != e->bundle_plans.end()) {
            dtl_engine::BundlePlan& pl = it->second;
            if (!pl.d_dest_ptr) {
                std::vector<int> ptr(pl.n_bundles + 1, 0), idx;
                for (int bi = 0; bi < pl.n_bundles; ++bi) {
                    for (int sl = 0; sl < B; ++sl) {
                        const int t = pl.h_tasks[(size_t)bi * B + sl];
                        if (t < 0) continue;
                        const int tg = plan_key + t; // plan_key is the first task of the batch
                        for (int q = e->task_cell_ptr[tg]; q < e->task_cell_ptr[tg + 1]; ++q)
                            idx.push_back(e->h_od_dest[e->h_task_cells[q]] * B + sl);
                    }
                    ptr[bi + 1] = (int)idx.size();
                }
                if (idx.empty()) idx.push_back(0);
                if (e->upload(&pl.d_dest_ptr, ptr.data(), ptr.size()) || e->upload(&pl.d_dest_idx, idx.data(), idx.size())) return -1;
            }
            if (e->bundle_bound_cap < (size_t)pl.n_bundles) {
                if (e->alloc(&e->d_bundle_bound, (size_t)pl.n_bundles)) return -1;
                e->bundle_bound_cap = (size_t)pl.n_bundles;
            }
            b.dest_ptr = pl.d_dest_ptr; b.dest_idx = pl.d_dest_idx; b.bundle_bound = e->d_bundle_bound;
        }
    }
    b.bundle_counter = e->d_bundle_counter;
    b.started_word = b.fuse_pred ? e->next_started_word : nullptr;
    b.started_tag = e->next_started_tag;
    b.grid = std::max(1, std::min(std::min(max_grid, e->bundle_grid), b.n_bundles));
    // CTAs without a bundle run the predecessor pass -- unless the side stream holds work (the scatter of this iteration): then the
    // SMs without a bundle are left to it, and the predecessor pass is done by the searching CTAs as they finish
    if (b.fuse_pred && !e->side_pending && !e->expect_side && !getenv("DTL_BUNDLE_NO_HELPERS")) b.grid = std::max(b.grid, std::min(max_grid, e->bundle_grid));
    return 0;
}

// K1 has three kernels (DESIGN.md "K1").  Per launch, by trees per SM:
//   * origin bundles (kernels_sssp_bundle*.cu) from 4 trees per SM upwards: 16 origins per search from 8 trees per SM (every SM
//     gets a bundle), 8 below.  A bundle keeps one SM busy for 5-7 ms whatever it holds, so this needs enough trees to hand
//     every SM a bundle; measured on the config-5 grid per launch: 2 000 trees 8.3 ms, 1 000 trees 6.1 ms (bundles of 8;
//     8.7 ms with the global-queue kernel), 500 trees 5.7 ms (5.2 ms with shared-memory queues);
//   * shared-memory queues (kernels_sssp_smq.cu) below 4 trees per SM (an origin shard on one of 4 or 8 GPUs): one big CTA
//     per tree, a launch lasts as long as one or two trees;
//   * global-memory queues (kernels_sssp.cu): networks whose per-node queue bits do not fit in shared memory, explicit block
//     sizes, and the worst-case re-run of trees whose queues overflowed.
// DTL_SSSP_KERNEL=gq|smq|bundle forces one (tests, experiments).
void choose_sssp_variant(dtl_engine* e, SsspLaunch& a)
{
    const char* force = getenv("DTL_SSSP_KERNEL");
    const int per_sm = (a.n_run + e->sm_count - 1) / std::max(1, e->sm_count);
    int v = per_sm <= 4 && !e->opt.sssp_block_threads ? 1 : 0; // an explicit block size asks for the global-queue kernel
    if (!e->opt.sssp_block_threads && !a.task_list && a.n_run >= 4 * e->sm_count) v = 2;
    if (force && !strcmp(force, "gq")) v = 0;
    if (force && !strcmp(force, "smq")) v = 1;
    if (force && !strcmp(force, "bundle")) v = 2;
    if (v == 2 && bundle_smem_bytes(e->N, 64) > (size_t)110 * 1024) v = 0; // the per-node queue bits must fit in shared memory
    a.smq = v;
    if (v != 1) return;
    a.block = per_sm <= 1 ? 1024 : per_sm <= 2 ? 512 : 256;
    if (const char* env = getenv("DTL_SSSP_SMQ_BLOCK")) a.block = atoi(env) == 256 ? 256 : atoi(env) == 512 ? 512 : 1024;
    a.cap_smem = 2 * a.block;
    if (const char* env = getenv("DTL_SSSP_QUEUE_CAP")) a.cap_smem = std::min(a.cap_smem, std::max(32, atoi(env))); // test hook
    if (const char* env = getenv("DTL_SSSP_SMQ_DELTA")) a.delta *= atof(env); // experiment hook: wider label window in latency mode
    a.grid = std::min(std::min(e->sssp_grid, e->sm_count * (1024 / a.block)), std::max(1, a.n_run));
}

static void launch_sssp_variant(const SsspLaunch& a, cudaStream_t st)
{
    if (a.smq == 1) launch_sssp_smq(a, st);
    else launch_sssp(a, st);
}

SsspLaunch make_sssp_launch(dtl_engine* e, const ForwardStar& f, int n_tasks, const int* d_origin, const int* d_zone,
                            int tree_off)
{
    SsspLaunch a{};
    a.N = e->N; a.row_ptr = f.row_ptr; a.rows = f.rows; a.link_info = f.link_info; a.edges = f.edges; a.in_ptr = f.in_ptr; a.in_edges = f.in_edges; a.rin4 = f.rin4;
    a.link_tag = e->d_link_tag; a.link_from = f.d_tail ? f.d_tail : e->d_link_from;
    a.n_tasks = n_tasks; a.n_run = n_tasks; a.task_list = nullptr; a.task_origin = d_origin; a.task_zone = d_zone;
    a.labels = e->d_labels + (size_t)tree_off * e->N; a.pred_link = e->d_pred + (size_t)tree_off * e->N;
    a.tie_nodes = e->d_tie_nodes;
    a.delta = f.delta;
    a.ws = e->d_ws; a.ws_stride = e->ws_stride; a.cap_near = e->cap_near; a.cap_far = e->cap_far;
    a.far_trigger = e->cap_far / 2;
    a.tie_cand = e->d_tie_cand; a.tie_cap = e->tie_cap; a.task_counter = e->d_task_counter; a.stats = e->d_stats;
    a.grid = std::min(e->sssp_grid, std::max(1, n_tasks)); a.block = e->sssp_block;
    choose_sssp_variant(e, a);
    a.bd_cell_ptr = nullptr; a.bd_cells = nullptr; a.bd_dest = nullptr;
    if (e->bounded && e->in_loop && !e->count_pass && a.smq == 1 && !e->opt.keep_trees && e->d_task_cell_ptr) {
        const int t0 = (int)(d_origin - e->d_task_origin); // first task of the batch
        a.bd_cell_ptr = e->d_task_cell_ptr + t0; a.bd_cells = e->d_task_cells; a.bd_dest = e->d_od_dest;
    }
    return a;
}

// One batch of shortest-path trees.  sssp_launch_batch launches the search; sssp_complete_batch waits for it, recomputes the
// trees whose queues overflowed and replays tied origins.  With `slim` (the caller only back-traces the trees) an origin-bundle
// batch leaves its trees node-major (BundleLaunch::bp) and writes no per-tree arrays; `st.slim` tells whether that happened.
struct BatchState {
    SsspLaunch a{};
    BundleLaunch b{};
    const int* d_slots = nullptr;
    bool bundle = false, slim = false;
    int helper_grid = 0; // pipelined iterations: helper CTAs to launch once the side stream's scatter is enqueued
    bool backtrace_issued = false; // pipelined iterations: the node-major back-trace was already put on the main stream
};

int sssp_launch_batch(dtl_engine* e, const ForwardStar& f, int n_tasks, const int* d_origin, const int* d_zone, int tree_off,
                      const int* h_origin, int plan_key, bool want_slim, BatchState& bs)
{
    SsspLaunch& a = bs.a;
    bs.backtrace_issued = false; bs.helper_grid = 0; // (a BatchState is reused by the pipelined loop)
    a = make_sssp_launch(e, f, n_tasks, d_origin, d_zone, tree_off);
    e->last_sssp_variant = a.smq;
    bs.bundle = a.smq == 2;
    bs.slim = bs.bundle && want_slim && !getenv("DTL_NO_SLIM_TREES");
    if (bs.bundle) { // origin bundles: one search per B neighbouring origins, then labels -> trees
        BundleLaunch& b = bs.b;
        if (make_bundle_launch(e, f, a, h_origin, plan_key, bs.slim, b, &bs.d_slots)) return -1;
        const int ph = e->phase_begin(2);
        CK(cudaMemsetAsync(e->d_bundle_counter, 0, sizeof(int), e->st));
        if (bs.slim) CK(cudaMemsetAsync(e->d_batch_flag, 0, sizeof(int), e->st));
        if (b.fuse_pred) {
            CK(cudaMemsetAsync(b.work, 0, sizeof(int), e->st));
            CK(cudaMemsetAsync(b.work + 1, 0xff, sizeof(int) * b.n_bundles, e->st));
            CK(cudaMemsetAsync(b.work + 1 + b.n_bundles, 0, sizeof(int) * (b.n_bundles + 1), e->st));
        }
        // the search kernel got no CTAs beyond its bundles (make_bundle_launch: the side stream holds the scatter): when that is
        // through, the SMs without a bundle run helper CTAs for the label -> predecessor pass, like the search kernel's own idle
        // CTAs would have (7.7 against 8.1 ms per 2 000 trees with / without them)
        const int max_grid = e->sm_count * b.ctas_per_sm;
        const bool want_helpers = b.fuse_pred && b.grid < max_grid && !getenv("DTL_BUNDLE_NO_HELPERS");
        bs.helper_grid = want_helpers && e->defer_helpers ? max_grid - b.grid : 0;
        const bool helpers = want_helpers && e->side_pending && !e->defer_helpers;
        if (helpers) {
            CK(cudaEventRecord(e->ev_fork, e->st)); // after the resets of b.work, before the search kernel
            CK(cudaStreamWaitEvent(e->st2, e->ev_fork, 0));
        }
        launch_sssp_bundle(b, e->st);
        if (helpers) {
            launch_bundle_pred_helpers(b, max_grid - b.grid, e->st2);
            CK(cudaEventRecord(e->ev_join, e->st2));
        }
        if (b.async) e->last_sssp_variant = 3;
        e->phase_end(ph, 2);
    } else {
        const int ph = e->phase_begin(2);
        CK(cudaMemsetAsync(e->d_task_counter, 0, sizeof(int), e->st));
        launch_sssp_variant(a, e->st);
        e->phase_end(ph, 1);
    }
    CK(cudaGetLastError());
    return 0;
}

int sssp_complete_batch(dtl_engine* e, BatchState& bs, int n_tasks, bool count_edges)
{
    SsspLaunch& a = bs.a;
    e->h_tie_nodes.resize(n_tasks);
    CK(cudaMemcpyAsync(e->h_tie_nodes.data(), e->d_tie_nodes, sizeof(int) * n_tasks, cudaMemcpyDeviceToHost, e->st));
    if (sync_stream(e)) return -1;
    std::vector<int> rerun;
    for (int i = 0; i < n_tasks; ++i)
        if (e->h_tie_nodes[i] < 0) rerun.push_back(i);
    if (e->pipelined) {
        bool any = !rerun.empty();
        for (int i = 0; i < n_tasks && !any; ++i) any = e->h_tie_nodes[i] > 0;
        // the re-run and replay kernels share counters and workspaces with the NEXT iteration's search, which is already running on the
        // other stream (e->st2 here: the streams are swapped while an iteration is completed): let it finish first
        if (any) CK(cudaStreamSynchronize(e->st2));
    }
    if (!rerun.empty()) { // frontier outgrew the regular queues: recompute those trees with worst-case capacities
        if (!e->d_ws_big) {
            if (e->alloc(&e->d_ws_big, (size_t)e->big_grid * e->ws_big_stride)) return -1;
            if (e->alloc(&e->d_ws_big_tie, (size_t)e->big_grid * e->tie_cap)) return -1;
            if (e->alloc(&e->d_rerun_list, std::max(1, e->batch_tasks))) return -1;
        }
        CK(cudaMemcpyAsync(e->d_rerun_list, rerun.data(), sizeof(int) * rerun.size(), cudaMemcpyHostToDevice, e->st));
        SsspLaunch b = a;
        b.task_list = e->d_rerun_list; b.n_run = (int)rerun.size();
        b.ws = e->d_ws_big; b.ws_stride = e->ws_big_stride; b.cap_near = e->cap_near_big; b.cap_far = e->cap_far_big;
        b.tie_cand = e->d_ws_big_tie; b.grid = std::min(e->big_grid, b.n_run);
        b.smq = 0; b.block = e->sssp_block; // the worst-case bounds are the global-queue kernel's
        b.far_trigger = e->N + 32; // a compacted pile holds <= N live entries, a round adds <= E, the trigger may lag one round: never overflows
        const int ph2 = e->phase_begin(2);
        CK(cudaMemsetAsync(e->d_task_counter, 0, sizeof(int), e->st));
        launch_sssp_variant(b, e->st);
        e->phase_end(ph2, 1);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(e->h_tie_nodes.data(), e->d_tie_nodes, sizeof(int) * n_tasks, cudaMemcpyDeviceToHost, e->st));
        if (sync_stream(e)) return -1;
        e->reruns += (double)rerun.size();
        for (int i : rerun)
            if (e->h_tie_nodes[i] < 0) { set_err("SSSP work queue overflow with worst-case capacities (internal bound violated)"); return -1; }
    }
    std::vector<int> replay;
    for (int i = 0; i < n_tasks; ++i)
        if (e->h_tie_nodes[i] > 0) replay.push_back(i);
    for (size_t done = 0; done < replay.size();) {
        if (!e->d_replay_tasks) {
            e->replay_cap = 64;
            if (e->alloc(&e->d_replay_tasks, e->replay_cap)) return -1;
            if (e->alloc(&e->d_replay_status, (size_t)e->replay_cap * e->N)) return -1;
            if (e->alloc(&e->d_replay_next, (size_t)e->replay_cap * e->N)) return -1;
        }
        const int n = (int)std::min<size_t>(e->replay_cap, replay.size() - done);
        CK(cudaMemcpyAsync(e->d_replay_tasks, replay.data() + done, sizeof(int) * n, cudaMemcpyHostToDevice, e->st));
        const int ph3 = e->phase_begin(3);
        launch_sssp_replay(a, e->d_replay_tasks, n, e->d_replay_status, e->d_replay_next, e->st);
        e->phase_end(ph3, 1);
        CK(cudaGetLastError());
        if (sync_stream(e)) return -1;
        done += n;
    }
    if (count_edges) {
        launch_count_edges(a, e->st);
        CK(cudaGetLastError());
    }
    e->tasks_done += n_tasks;
    return 0;
}

// run the SSSP kernel for one batch (trees land in d_labels/d_pred), replay tied origins
int run_sssp_batch(dtl_engine* e, const ForwardStar& f, int n_tasks, const int* d_origin, const int* d_zone, bool count_edges,
                   int tree_off, const int* h_origin, int plan_key)
{
    BatchState bs;
    if (sssp_launch_batch(e, f, n_tasks, d_origin, d_zone, tree_off, h_origin, plan_key, false, bs)) return -1;
    return sssp_complete_batch(e, bs, n_tasks, count_edges);
}

int check_flags(dtl_engine* e, bool collective = false)
{
    int flag = 0;
    SsspStats s{};
    // inside the assignment loop every rank sees the worst flag of any rank and stops at the same point (integer max: exact)
    int* d_flag = e->d_error_flag;
    if (collective && e->pipelined && e->world > 1 && !e->local_shard) {
        // the next iteration's kernels are running on the other stream and may set bits: reduce a copy (the bits stay in d_error_flag
        // and are seen one iteration later at the latest)
        d_flag = reinterpret_cast<int*>(e->d_counters + 3);
        CK(cudaMemcpyAsync(d_flag, e->d_error_flag, sizeof(int), cudaMemcpyDeviceToDevice, e->st));
    }
    if (collective && all_reduce(e, d_flag, 1, ncclInt32, ncclMax)) return -1;
    CK(cudaMemcpyAsync(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, e->st));
    CK(cudaMemcpyAsync(&s, e->d_stats, sizeof(s), cudaMemcpyDeviceToHost, e->st));
    if (sync_stream(e)) return -1;
    if (flag & 1) { set_err("column arena exhausted (%lld path links): raise dtl_options.column_arena_bytes", e->pool.arena_cap); return -1; }
    if (flag & 2) { set_err("an OD cell needs more than %d columns: raise dtl_options.max_columns_per_od", e->pool.cap); return -1; }
    if (flag & 4) {
        set_err("a generalized link cost (travel time + penalty + LR price + toll / VOT * 60) is negative: not supported, the "
                "shortest-path kernels order labels by their FP64 bit pattern (reference routing.h:229-231 accepts it)");
        return -1;
    }
    (void)s; // queue overflows are handled per batch by re-running those trees with worst-case queues
    return 0;
}

} // namespace

// =================================================================================================
extern "C" {

const char* dtl_last_error(void) { return g_err.c_str(); }
const char* dtl_version(void) { return "dtalite_b200 0.1 (sm_100a)"; }

void dtl_default_options(dtl_options* o)
{
    memset(o, 0, sizeof(*o));
    o->device = -1;
    o->max_columns_per_od = 24;
    o->rank = 0;
    o->world_size = 1;
    o->sssp_delta_factor = 0;
    o->sssp_block_threads = 0;
    o->sssp_ctas_per_sm = 0;
    o->tree_batch_bytes = 0;
    o->column_arena_bytes = 0;
    o->keep_trees = 0;
}

void dtl_destroy(dtl_engine* e)
{
    if (!e) return;
    if (e->embed_thread.joinable()) e->embed_thread.join();
    cudaSetDevice(e->dev);
    if (e->st2) cudaStreamSynchronize(e->st2);
    if (e->st) cudaStreamSynchronize(e->st);
    e->resolve_timing();
    for (cudaEvent_t ev : e->free_events) cudaEventDestroy(ev);
    for (cudaEvent_t ev : { e->ev_k1[0], e->ev_k1[1], e->ev_vdf, e->ev_done })
        if (ev) cudaEventDestroy(ev);
    if (e->ev_fork) cudaEventDestroy(e->ev_fork);
    if (e->ev_join) cudaEventDestroy(e->ev_join);
    if (e->st2) cudaStreamDestroy(e->st2);
    // e->comm is shared and lives as long as the process (see dtl_comm_init)
    for (void* p : e->allocs) cudaFree(p);
    if (e->st) cudaStreamDestroy(e->st);
    delete e;
}
// DTL_CREATE_TRACE=1: wall-clock time between the sections of create_impl on stderr
#define DTL_TRACE(what)                                                                                          \
    do {                                                                                                         \
        if (trace_on) {                                                                                          \
            const auto now_ = std::chrono::steady_clock::now();                                                 \
            fprintf(stderr, "dtl_create: %8.3f ms before \"%s\"\n", std::chrono::duration<double, std::milli>(now_ - trace_t).count(), what); \
            trace_t = now_;                                                                                      \
        }                                                                                                        \
    } while (0)


static int create_impl(dtl_engine* e, const dtl_problem* p, const dtl_options* opt)
{
    const bool trace_on = getenv("DTL_CREATE_TRACE") != nullptr;
    auto trace_t = std::chrono::steady_clock::now();
    dtl_options o;
    if (opt) o = *opt; else dtl_default_options(&o);
    e->opt = o;
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0) {
        set_err("no CUDA device available (%s): this engine has no CPU fallback", cudaGetErrorString(ce));
        return -1;
    }
    if (o.device >= 0) CK(cudaSetDevice(o.device));
    CK(cudaGetDevice(&e->dev));
    DTL_TRACE("device selected");
    CK(cudaDeviceGetAttribute(&e->sm_count, cudaDevAttrMultiProcessorCount, e->dev)); // (cudaGetDeviceProperties costs milliseconds)
    {
        // the main stream (shortest-path search) gets its CTAs placed before the side stream's
        int pr_least = 0, pr_greatest = 0;
        CK(cudaDeviceGetStreamPriorityRange(&pr_least, &pr_greatest));
        CK(cudaStreamCreateWithPriority(&e->st, cudaStreamNonBlocking, pr_greatest));
        e->overlap = !getenv("DTL_NO_OVERLAP");
        e->bounded = getenv("DTL_COMPLETE_SEARCH") == nullptr; // the API is dtl_set_bounded_search
        {
            CK(cudaStreamCreateWithPriority(&e->st2, cudaStreamNonBlocking, pr_least));
            CK(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming));
        }
    }
    DTL_TRACE("stream");
    e->h_scalars = e->h_scalars_buf; // pageable: a 16-byte read-back per iteration does not pay for pinning memory (2.7 ms)

    e->N = p->n_nodes; e->L = p->n_links; e->Z = p->n_zones; e->AT = p->n_agent_types; e->T = p->n_periods;
    e->B = std::max(1, p->memory_blocks);
    e->rank = o.rank; e->world = std::max(1, o.world_size);
    e->total_demand = p->total_demand_volume;
    const int N = e->N, L = e->L, Z = e->Z, AT = e->AT, T = e->T;
    if (N <= 0 || L < 0 || Z <= 0 || AT <= 0 || T <= 0) { set_err("empty problem"); return -1; }
    if (e->rank < 0 || e->rank >= e->world) { set_err("rank %d outside world of %d", e->rank, e->world); return -1; }
    e->h_link_from.assign(p->link_from, p->link_from + L);
    e->h_link_to.assign(p->link_to, p->link_to + L);
    e->h_link_tag.assign(p->link_tag, p->link_tag + L);
    e->h_node_zone.assign(p->node_zone, p->node_zone + N);
    e->h_zone_centroid.assign(p->zone_centroid, p->zone_centroid + Z);
    e->h_vot.assign(p->vot, p->vot + AT);
    for (int l = 0; l < L; ++l)
        if (p->link_from[l] < 0 || p->link_from[l] >= N || p->link_to[l] < 0 || p->link_to[l] >= N) {
            set_err("link %d has an endpoint outside [0, %d)", l, N);
            return -1;
        }

    for (int z = 0; z < Z; ++z)
        if (p->zone_centroid[z] < 0 || p->zone_centroid[z] >= N) { set_err("zone %d: centroid node outside [0, %d)", z, N); return -1; }
    for (int l = 0; l < L; ++l)
        if (p->link_tag[l] >= Z) { set_err("link %d: outgoing-connector zone tag %d outside [0, %d)", l, p->link_tag[l], Z); return -1; }
    // the host images of the forward stars are built by a worker thread while this thread partitions the demand and uploads
    // the link arrays (no CUDA call and no error path inside the worker)
    e->fs.resize((size_t)AT * T);
    std::vector<FsHost> fsh((size_t)AT * T);
    // (the same thread goes on to the bundle planner's landmark embeddings -- three or four breadth-first searches per forward
    // star -- and is joined by the first bundle plan; the zone partition of a multi-GPU run computes its own embedding on this
    // thread first, so the worker only reads e->h_embed_all after `owners_ready`)
    std::promise<void> stars_built, owners_ready;
    std::future<void> stars_future = stars_built.get_future();
    std::shared_future<void> owners_future = owners_ready.get_future().share();
    bool owners_signalled = false;
    e->embed_thread = std::thread([e, p, AT, T, L, &fsh, &stars_built, owners_future]() {
        for (int at = 0; at < AT; ++at)
            for (int tau = 0; tau < T; ++tau) build_forward_star_host(e, p, at, tau, fsh[fs_index(e, at, tau)]);
        stars_built.set_value();
        owners_future.wait();
        for (ForwardStar& f : e->fs) {
            if (!e->h_embed_all.empty() && f.n_edges == L) f.h_embed = e->h_embed_all; // same graph as the zone partition's
            else ensure_embedding(f, e->N);
        }
    });
    // an early return must not take `fsh` and the promises away from under the worker: it touches them only until it has
    // signalled `stars_built` (afterwards it works on the engine alone, and dtl_destroy joins it)
    bool stars_waited = false;
    struct Joiner {
        std::promise<void>& pr; bool& owners_done; std::future<void>& stars; bool& stars_done;
        ~Joiner() { if (!owners_done) pr.set_value(); if (!stars_done) stars.wait(); }
    } fs_joiner{ owners_ready, owners_signalled, stars_future, stars_waited };
    DTL_TRACE("tasks of this rank: origin zones are sha");
    // ---- tasks and OD cells of this rank ----
    // Origin zones are sharded over the ranks in spatial blocks (zone_owners; the reference deals zone % blocks); cells whose
    // origin has no task never get columns in the reference either.  The partition (with more than one rank: four breadth-first
    // searches), the task list and the cell arrays (two passes over the OD rows: count per task, then fill) are put together by
    // a second worker thread while this thread uploads the link arrays; they are uploaded when it is done.
    std::vector<int> owner, task_of, origin, od_dest, od_at, od_tau, od_task, task_cells, od_dzone;
    std::vector<unsigned char> skip;
    std::vector<double> od_vol;
    int bad_od_row = -1, n_tasks = 0;
    std::thread cells_thread([&]() {
        std::vector<int> g_at, g_tau, g_zone;
        build_task_list(p, g_at, g_tau, g_zone);
        owner = zone_owners(p, e->world, &e->h_embed_all);
        owners_ready.set_value();
        owners_signalled = true;
        task_of.assign((size_t)AT * T * Z, -1); // (at,tau,zone) -> local task
        for (size_t i = 0; i < g_at.size(); ++i)
            if (owner[g_zone[i]] == e->rank) {
                task_of[((long long)g_at[i] * T + g_tau[i]) * Z + g_zone[i]] = (int)e->task_at.size();
                e->task_at.push_back(g_at[i]); e->task_tau.push_back(g_tau[i]); e->task_zone.push_back(g_zone[i]);
            }
        n_tasks = (int)e->task_at.size();
        origin.assign(n_tasks, 0);
        skip.assign(n_tasks, 0);
        std::vector<int> n_out(N, 0);
        for (int l = 0; l < L; ++l) n_out[p->link_from[l]]++;
        for (int i = 0; i < n_tasks; ++i) {
            origin[i] = p->zone_centroid[e->task_zone[i]];
            skip[i] = n_out[origin[i]] == 0; // routing.h:428
        }
        e->h_task_origin = origin;
        e->task_cell_ptr.assign(n_tasks + 1, 0);
        std::vector<int> task_i_of(p->n_od, -1);
        size_t n_keep = 0;
        for (int i = 0; i < p->n_od; ++i) {
            if (p->od_o[i] < 0 || p->od_o[i] >= Z || p->od_d[i] < 0 || p->od_d[i] >= Z || p->od_at[i] < 0 || p->od_at[i] >= AT ||
                p->od_tau[i] < 0 || p->od_tau[i] >= T) {
                bad_od_row = i;
                return;
            }
            if (owner[p->od_o[i]] != e->rank) continue;
            const int t = task_of[((size_t)p->od_at[i] * T + p->od_tau[i]) * Z + p->od_o[i]];
            if (t < 0) continue;
            task_i_of[i] = t;
            e->task_cell_ptr[t + 1]++;
            ++n_keep;
        }
        for (int t = 0; t < n_tasks; ++t) e->task_cell_ptr[t + 1] += e->task_cell_ptr[t];
        std::vector<int> cur(e->task_cell_ptr.begin(), e->task_cell_ptr.end() - 1);
        e->cell_global.resize(n_keep); od_dest.resize(n_keep); od_at.resize(n_keep); od_tau.resize(n_keep); od_task.resize(n_keep);
        od_vol.resize(n_keep); task_cells.resize(n_keep); od_dzone.resize(n_keep);
        int c = 0;
        for (int i = 0; i < p->n_od; ++i) {
            const int t = task_i_of[i];
            if (t < 0) continue;
            e->cell_global[c] = i;
            od_dest[c] = p->zone_centroid[p->od_d[i]];
            od_at[c] = p->od_at[i]; od_tau[c] = p->od_tau[i]; od_task[c] = t; od_vol[c] = p->od_vol[i];
            od_dzone[c] = p->od_d[i];
            task_cells[cur[t]++] = c;
            ++c;
        }
    });
    struct ThreadJoiner { std::thread& t; ~ThreadJoiner() { if (t.joinable()) t.join(); } } cells_joiner{ cells_thread };

    DTL_TRACE("network image");
    // ---- network image: device arrays are allocated here, the copies (about sixty megabytes of pageable host memory on the
    // config-5 grid, ~30 arrays) are issued by three threads -- a pageable copy is bounded by the host-side staging of the
    // calling thread, not by the link ----
    struct CopyJob { void* dst; const void* src; size_t bytes; };
    std::vector<CopyJob> jobs;
    auto up = [&](auto** dst, const auto* src, size_t n) -> int {
        if (e->alloc(dst, n)) return -1;
        if (n) jobs.push_back(CopyJob{ (void*)*dst, (const void*)src, n * sizeof(**dst) });
        return 0;
    };
    if (up(&e->d_link_tag, p->link_tag, (size_t)L)) return -1;
    if (up(&e->d_link_from, p->link_from, (size_t)L)) return -1;
    const size_t TL = (size_t)T * L, TAL = (size_t)T * AT * L;
#define UPV(field) if (up(&e->vp.field, p->field, TL)) return -1
    UPV(fftt); UPV(cap); UPV(alpha); UPV(beta); UPV(vf); UPV(vc); UPV(nlanes); UPV(cap_lane); UPV(qdf); UPV(preload);
    UPV(penalty); UPV(Q_alpha); UPV(Q_beta); UPV(Q_cd); UPV(Q_n); UPV(Q_cp); UPV(Q_s); UPV(t2); UPV(L_h); UPV(start_h);
    UPV(end_h); UPV(plf); UPV(cycle_length); UPV(red_time); UPV(sat_flow); UPV(vdf_type);
#undef UPV
    if (up(&e->d_pce, p->pce, TAL)) return -1;
    if (up(&e->d_occ, p->occ, TAL)) return -1;
    if (up(&e->d_toll, p->toll, TAL)) return -1;
    if (up(&e->d_lr, p->lr_price, TAL)) return -1;
    {
        std::atomic<size_t> next{ 0 };
        std::atomic<int> failed{ 0 };
        auto worker = [&]() {
            if (cudaSetDevice(e->dev) != cudaSuccess) { failed = 1; return; }
            for (size_t j; (j = next.fetch_add(1)) < jobs.size();)
                if (cudaMemcpy(jobs[j].dst, jobs[j].src, jobs[j].bytes, cudaMemcpyHostToDevice) != cudaSuccess) failed = 1;
        };
        std::thread h1(worker), h2(worker);
        worker();
        h1.join(); h2.join();
        if (failed) { set_err("uploading the network image failed: %s", cudaGetErrorString(cudaGetLastError())); return -1; }
    }
    { // per (tau, at): the PCE / occupancy when all links share it (K3 then skips the gather), NaN otherwise
        std::vector<double> pu((size_t)T * AT), ou((size_t)T * AT);
        for (int ta = 0; ta < T * AT; ++ta) {
            const double* a0 = p->pce + (size_t)ta * L;
            const double* o0 = p->occ + (size_t)ta * L;
            bool ps = true, os = true;
            for (int l = 1; l < L && (ps || os); ++l) { ps = ps && a0[l] == a0[0]; os = os && o0[l] == o0[0]; }
            pu[ta] = ps && L > 0 ? a0[0] : std::nan("");
            ou[ta] = os && L > 0 ? o0[0] : std::nan("");
        }
        if (e->upload(&e->d_pce_uni, pu.data(), pu.size())) return -1;
        if (e->upload(&e->d_occ_uni, ou.data(), ou.size())) return -1;
        if (e->alloc(&e->d_link_cost, TAL)) return -1;
    }
    if (e->upload(&e->d_vot, p->vot, AT)) return -1;
    if (e->alloc(&e->ls.pce_fix, TL)) return -1;
    if (e->alloc(&e->ls.person_fix, TL * (1 + AT))) return -1;
    double** st_fields[] = { &e->ls.tt, &e->ls.link_volume, &e->ls.doc, &e->ls.voc, &e->ls.P, &e->ls.vt2,
                             &e->ls.avg_queue_speed, &e->ls.avg_speed_bpr, &e->ls.Q_mu, &e->ls.Q_gamma, &e->ls.t0,
                             &e->ls.t3, &e->ls.severe_P };
    for (double** f : st_fields) {
        if (e->alloc(f, TL)) return -1;
        CK(cudaMemset(*f, 0, TL * sizeof(double)));
    }
    CK(cudaMemset(e->ls.pce_fix, 0, TL * 8));
    CK(cudaMemset(e->ls.person_fix, 0, TL * (1 + AT) * 8));
    if (e->alloc(&e->d_block_sums, vdf_blocks((int)TL) + 1)) return -1;
    if (e->alloc(&e->d_scalars, 16)) return -1;
    if (e->alloc(&e->d_volume_tmp, TL)) return -1;
    DTL_TRACE("link arrays uploaded");
    cells_thread.join();
    if (bad_od_row >= 0) { set_err("OD row %d: zone, agent type or period index out of range", bad_od_row); return -1; }
    if (e->upload(&e->d_task_origin, origin.data(), n_tasks)) return -1;
    if (e->upload(&e->d_task_zone, e->task_zone.data(), n_tasks)) return -1;
    if (e->upload(&e->d_task_skip, skip.data(), n_tasks)) return -1;
    e->h_task_skip = skip;
    e->n_cells = (int)e->cell_global.size();
    if (e->upload(&e->d_od_dest, od_dest.data(), e->n_cells)) return -1;
    if (e->upload(&e->d_od_at, od_at.data(), e->n_cells)) return -1;
    if (e->upload(&e->d_od_tau, od_tau.data(), e->n_cells)) return -1;
    if (e->upload(&e->d_od_task, od_task.data(), e->n_cells)) return -1;
    e->h_od_at = od_at; e->h_od_tau = od_tau; e->h_od_task = od_task; e->h_od_dzone = od_dzone;
    e->h_task_cells = task_cells; e->h_od_dest = od_dest;
    if (e->upload(&e->d_od_vol, od_vol.data(), e->n_cells)) return -1;
    if (e->upload(&e->d_task_cells, task_cells.data(), task_cells.size())) return -1;
    if (!getenv("DTL_NO_CELL_SORT")) {
        if (e->alloc(&e->d_path_len, std::max<size_t>(1, task_cells.size()))) return -1;
        if (e->upload(&e->d_task_cell_ptr, e->task_cell_ptr.data(), e->task_cell_ptr.size())) return -1;
    }
    DTL_TRACE("OD cells uploaded");

    stars_future.wait();
    stars_waited = true;
    DTL_TRACE("forward stars built (worker thread)");
    int maxE = 0;
    for (int at = 0; at < AT; ++at)
        for (int tau = 0; tau < T; ++tau) {
            if (upload_forward_star(e, at, tau, fsh[fs_index(e, at, tau)])) return -1;
            maxE = std::max(maxE, e->fs[fs_index(e, at, tau)].n_edges);
        }
    fsh.clear();
    std::vector<int*> le(AT * T);
    std::vector<Edge*> ed(AT * T);
    for (int i = 0; i < AT * T; ++i) { le[i] = e->fs[i].link_edge; ed[i] = e->fs[i].edges; }
    if (e->upload(&e->d_link_edge_ptrs, le.data(), le.size())) return -1;
    if (e->upload(&e->d_edges_ptrs, ed.data(), ed.size())) return -1;
    DTL_TRACE("forward stars");

    DTL_TRACE("column pool");
    // ---- column pool ----
    size_t free_b = 0, total_b = 0;
    CK(cudaMemGetInfo(&free_b, &total_b));
    DTL_TRACE("cudaMemGetInfo");
    e->pool.n_od = e->n_cells;
    e->pool.cap = std::max(1, o.max_columns_per_od);
    const size_t slots = (size_t)e->n_cells * e->pool.cap;
    if (e->alloc(&e->pool.count, e->n_cells)) return -1;
    CK(cudaMemset(e->pool.count, 0, sizeof(int) * std::max(1, e->n_cells)));
    if (e->alloc(&e->pool.node_sum, slots)) return -1;
    if (e->alloc(&e->pool.n_links, slots)) return -1;
    if (e->alloc(&e->pool.n_nodes, slots)) return -1;
    if (e->alloc(&e->pool.offset, slots)) return -1;
    if (e->alloc(&e->pool.volume, slots)) return -1;
    if (e->alloc(&e->pool.cost, slots)) return -1;
    if (e->alloc(&e->pool.time, slots)) return -1;
    if (e->alloc(&e->pool.arena_used, 1)) return -1;
    CK(cudaMemset(e->pool.arena_used, 0, 8));
    if (e->alloc(&e->d_error_flag, 1)) return -1;
    CK(cudaMemset(e->d_error_flag, 0, 4));
    if (e->alloc(&e->d_least, e->n_cells)) return -1;
    CK(cudaMemset(e->d_least, 0, sizeof(double) * std::max(1, e->n_cells)));
    if (e->alloc(&e->d_od_sums, (size_t)e->n_cells * 4)) return -1;
    if (e->alloc(&e->d_stats, 1)) return -1;
    CK(cudaMemset(e->d_stats, 0, sizeof(SsspStats)));
    if (e->alloc(&e->d_counters, 4)) return -1;
    CK(cudaMemset(e->d_counters, 0, 32));
    if (e->alloc(&e->d_task_counter, 1)) return -1;

    DTL_TRACE("tree batch + SSSP workspace");
    // ---- tree batch + SSSP workspace ----
    const size_t per_tree = (size_t)N * 12;
    size_t tree_budget = o.tree_batch_bytes ? o.tree_batch_bytes : std::min<size_t>((size_t)24 << 30, free_b / 6);
    e->batch_tasks = (int)std::max<size_t>(1, std::min<size_t>(std::max(1, n_tasks), tree_budget / per_tree));
    if (o.keep_trees) e->batch_tasks = std::max(1, n_tasks);
    if (e->alloc(&e->d_labels, (size_t)e->batch_tasks * N)) return -1;
    if (e->alloc(&e->d_pred, (size_t)e->batch_tasks * N)) return -1;
    if (e->alloc(&e->d_tie_nodes, e->batch_tasks)) return -1;
    e->sssp_block = o.sssp_block_threads ? o.sssp_block_threads : 128;
    if (e->sssp_block != 64 && e->sssp_block != 128 && e->sssp_block != 256 && e->sssp_block != 512 && e->sssp_block != 1024) e->sssp_block = 128;
    int occ = 0, regs = 0;
    sssp_occupancy(e->sssp_block, &occ, &regs);
    if (occ <= 0) { set_err("SSSP kernel cannot be resident on this device (built for sm_100a)"); return -1; }
    int per_sm = o.sssp_ctas_per_sm > 0 ? std::min(o.sssp_ctas_per_sm, occ) : occ;
    // worst-case bounds (DESIGN.md "K1 queues"): a round pushes at most E entries, a compacted far pile holds at most
    // N live entries, and the compaction trigger may lag one round.  Regular CTAs get much smaller queues; a tree that
    // outgrows them is re-run with the bounds.
    e->cap_near_big = std::max(N, maxE) + 32;
    e->cap_far_big = N + 2 * maxE + 64;
    e->ws_big_stride = 2 * (size_t)e->cap_near_big + 2 * (size_t)e->cap_far_big + 4;
    int qcap = std::max(4096, N / 8);
    if (const char* env = getenv("DTL_SSSP_QUEUE_CAP")) qcap = std::max(4, atoi(env)); // test hook for the re-run path
    e->cap_near = std::min(e->cap_near_big, qcap);
    e->cap_far = std::min(e->cap_far_big, 4 * qcap);
    e->ws_stride = 2 * (size_t)e->cap_near + 2 * (size_t)e->cap_far + 4;
    const size_t ws_budget = std::min<size_t>((size_t)16 << 30, free_b / 4);
    int grid = e->sm_count * per_sm;
    grid = (int)std::min<size_t>(grid, std::max<size_t>(1, ws_budget / (e->ws_stride * sizeof(QEntry))));
    grid = std::min(grid, std::max(1, e->batch_tasks));
    e->sssp_grid = std::max(1, grid);
    if (e->alloc(&e->d_ws, (size_t)e->sssp_grid * e->ws_stride)) return -1;
    if (e->alloc(&e->d_tie_cand, (size_t)e->sssp_grid * e->tie_cap)) return -1;

    DTL_TRACE("path-link arena: what is left, bounded b");
    // ---- path-link arena: what is left, bounded by the option ----
    CK(cudaMemGetInfo(&free_b, &total_b));
    size_t arena_bytes = o.column_arena_bytes ? o.column_arena_bytes : std::min<size_t>((size_t)6 << 30, free_b / 4);
    // a column never has more links than nodes: enough for every slot is an upper bound worth respecting on small inputs
    const size_t hard = (slots + 1) * (size_t)(((size_t)N + 3) & ~(size_t)3) * sizeof(int);
    arena_bytes = std::min(arena_bytes, hard);
    e->pool.arena_cap = (long long)(arena_bytes / sizeof(int));
    if (e->alloc(&e->pool.arena, (size_t)e->pool.arena_cap)) return -1;
    CK(cudaDeviceSynchronize());
    DTL_TRACE("end");
    return 0;
}

dtl_engine* dtl_create(const dtl_problem* p, const dtl_options* opt)
{
    if (!p) { set_err("null problem"); return nullptr; }
    dtl_engine* e = new dtl_engine();
    if (create_impl(e, p, opt)) {
        std::string keep = g_err;
        dtl_destroy(e);
        g_err = keep;
        return nullptr;
    }
    return e;
}

size_t dtl_device_bytes(const dtl_engine* e) { return e ? e->dev_bytes : 0; }
int dtl_sssp_kernel_used(const dtl_engine* e) { return e ? e->last_sssp_variant : -1; }
int dtl_plan_bundles(const int* row_ptr, const int* head, int n_nodes, const int* origins, int n, int B, int* out, int capacity)
{
    if (!row_ptr || !head || !origins || n_nodes <= 0 || n < 0 || (B != 2 && B != 4 && B != 8 && B != 16)) { set_err("bad arguments"); return -1; }
    for (int i = 0; i < n; ++i)
        if (origins[i] < 0 || origins[i] >= n_nodes) { set_err("origin %d out of range", origins[i]); return -1; }
    ForwardStar f;
    f.h_row_ptr.assign(row_ptr, row_ptr + n_nodes + 1);
    f.h_to.assign(head, head + row_ptr[n_nodes]);
    ensure_embedding(f, n_nodes);
    std::vector<int> idx(n), res;
    std::iota(idx.begin(), idx.end(), 0);
    kd_split(idx, 0, n, B, B, f.h_embed.data(), origins, res);
    for (size_t i = 0; i < res.size() && (int)i < capacity; ++i) out[i] = res[i];
    return (int)res.size() / B;
}
int dtl_num_tasks(const dtl_engine* e) { return (int)e->task_at.size(); }
int dtl_get_tasks(const dtl_engine* e, int* at, int* tau, int* zone)
{
    const size_t n = e->task_at.size();
    if (at) memcpy(at, e->task_at.data(), n * sizeof(int));
    if (tau) memcpy(tau, e->task_tau.data(), n * sizeof(int));
    if (zone) memcpy(zone, e->task_zone.data(), n * sizeof(int));
    return 0;
}

int dtl_plan_tasks(const dtl_problem* p, int rank, int world_size, int* at, int* tau, int* zone, int capacity)
{
    if (!p || world_size < 1 || rank < 0 || rank >= world_size) { set_err("bad arguments"); return -1; }
    std::vector<int> a, t, z;
    build_task_list(p, a, t, z);
    for (int i = 0; i < p->n_zones; ++i)
        if (p->zone_centroid[i] < 0 || p->zone_centroid[i] >= p->n_nodes) { set_err("zone %d: centroid node outside [0, %d)", i, p->n_nodes); return -1; }
    const std::vector<int> owner = zone_owners(p, world_size);
    int n = 0;
    for (size_t i = 0; i < a.size(); ++i)
        if (owner[z[i]] == rank) {
            if (n < capacity) {
                if (at) at[n] = a[i];
                if (tau) tau[n] = t[i];
                if (zone) zone[n] = z[i];
            }
            ++n;
        }
    return n;
}

int dtl_comm_get_unique_id(unsigned char id128[128])
{
    if (!g_nccl.load()) return -1;
    ncclUniqueId id;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclResult_t r = g_nccl.GetUniqueId(&id);
    if (r != ncclSuccess) { set_err("ncclGetUniqueId failed"); return -1; }
    memcpy(id128, &id, 128);
    return 0;
}

int dtl_comm_init(dtl_engine* e, const unsigned char id128[128])
{
    if (e->world <= 1) return 0;
    if (!id128) { e->local_shard = true; return 0; } // shard-only mode: no collective, the caller sums the shards
    if (!g_nccl.load()) return -1;
    CK(cudaSetDevice(e->dev));
    // communicators are process-lifetime objects keyed by their id: a second engine of the same run re-uses it
    std::map<std::string, ncclComm_t>& comms = comm_table();
    const std::string key((const char*)id128, 128);
    e->comm_key = key;
    auto it = comms.find(key);
    if (it != comms.end()) { e->comm = it->second; return 0; }
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ncclResult_t r = g_nccl.CommInitRank(&e->comm, e->world, id, e->rank);
    if (r != ncclSuccess) { set_err("ncclCommInitRank failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"); return -1; }
    comms[key] = e->comm;
    return 0;
}

// -------------------------------------------------------------------------------------------------
static int columns_of_batch(dtl_engine* e, BatchState& bs, int t0, int t1, int tree_off, int k, bool want_count);

static int trees_and_columns(dtl_engine* e, int k, bool want_count)
{
    const int n_tasks = (int)e->task_at.size();
    int t0 = 0;
    while (t0 < n_tasks) {
        // a batch never spans two forward stars
        const int fsi = fs_index(e, e->task_at[t0], e->task_tau[t0]);
        int t1 = t0;
        while (t1 < n_tasks && t1 - t0 < e->batch_tasks && fs_index(e, e->task_at[t1], e->task_tau[t1]) == fsi) ++t1;
        const int nb = t1 - t0;
        if (e->sa_mode) {                                                 // main_api.cpp:984-994
            const int rt = e->sa_rt_info[e->task_at[t0]];
            bool any = rt == 1;
            if (rt == 2) // message-sign users: only the origins that are information zones (the other trees of the batch are not back-traced)
                for (int t = t0; t < t1 && !any; ++t) any = !e->info_zone.empty() && e->info_zone[e->task_zone[t]];
            if (!any) { t0 = t1; continue; }
        }
        const int tree_off = e->opt.keep_trees ? t0 : 0; // keep_trees: the batch buffer holds every task
        // the loop only back-traces the trees: an origin-bundle batch may leave them node-major (not when the trees are kept:
        // dtl_get_tree reads the per-tree arrays)
        BatchState bs;
        e->in_loop = true;
        const int rc_l = sssp_launch_batch(e, e->fs[fsi], nb, e->d_task_origin + t0, e->d_task_zone + t0, tree_off, e->h_task_origin.data() + t0, t0,
                                           !e->opt.keep_trees, bs);
        e->in_loop = false;
        if (rc_l) return -1;
        if (columns_of_batch(e, bs, t0, t1, tree_off, k, want_count)) return -1;
        t0 = t1;
    }
    e->trees_valid = e->opt.keep_trees != 0;
    return 0;
}

// the tail of a batch: flags of the search (ties, abandoned bundles, overflowed queues -> redone / replayed), then the back-trace
static ColumnLaunch make_column_launch(dtl_engine* e, int t0, int t1, int tree_off, int k)
{
    ColumnLaunch c{};
    c.N = e->N; c.L = e->L; c.T = e->T; c.AT = e->AT; c.pool = e->pool;
    c.od_dest_node = e->d_od_dest; c.od_at = e->d_od_at; c.od_tau = e->d_od_tau; c.od_task = e->d_od_task;
    c.od_vol = e->d_od_vol; c.link_from = e->d_link_from;
    c.labels = e->d_labels + (size_t)tree_off * e->N; c.pred_link = e->d_pred + (size_t)tree_off * e->N;
    c.task_first = t0; c.task_count = t1 - t0;
    c.batch_cells = e->d_task_cells + e->task_cell_ptr[t0];
    c.n_batch_cells = e->task_cell_ptr[t1] - e->task_cell_ptr[t0];
    c.task_skip_backtrace = e->sa_mode && e->d_sa_task_skip ? e->d_sa_task_skip : e->d_task_skip;
    c.k = k; c.least_contrib = e->d_least; c.error_flag = e->d_error_flag; c.path_nodes = e->d_counters + 1;
    c.path_len = e->d_path_len;
    return c;
}

// back-trace on the node-major trees of a bundle batch (on e->st); the kernel does nothing when the label -> tree pass flagged a tie
// or an overflowed bundle, and the batch is then redone through the per-tree path (columns_of_batch)
static int issue_slim_backtrace(dtl_engine* e, BatchState& bs, int t0, int t1, int tree_off, int k)
{
    ColumnLaunch cs = make_column_launch(e, t0, t1, tree_off, k);
    cs.bl = bs.b.bl; cs.bp = bs.b.bp; cs.task_slot = bs.d_slots; cs.B = bs.b.B; cs.batch_flag = bs.b.batch_flag;
    e->join_side(); // the scatter's MSA decay of the column volumes comes before this iteration's share is added
    const int ph = e->phase_begin(4);
    launch_backtrace(cs, e->st);
    e->phase_end(ph, 1);
    CK(cudaGetLastError());
    bs.backtrace_issued = true;
    return 0;
}

static int columns_of_batch(dtl_engine* e, BatchState& bs, int t0, int t1, int tree_off, int k, bool want_count)
{
    const int nb = t1 - t0;
    {
        ColumnLaunch c = make_column_launch(e, t0, t1, tree_off, k);
        bool done = false;
        if (bs.slim) {
            if (!bs.backtrace_issued && issue_slim_backtrace(e, bs, t0, t1, tree_off, k)) return -1;
            int flag = 0;
            CK(cudaMemcpyAsync(&flag, e->d_batch_flag, sizeof(int), cudaMemcpyDeviceToHost, e->st));
            if (sync_stream(e)) return -1;
            if (!flag) {
                done = true;
                e->tasks_done += nb;
                if (want_count) { launch_count_edges_nm(bs.b, e->st); CK(cudaGetLastError()); } // counted on the node-major labels
                if (e->d_path_len && !e->sorted_batches.count(t0)) {
                    e->sorted_batches.insert(t0);
                    sort_cells_by_path_length(e->d_path_len, e->d_task_cells + e->task_cell_ptr[t0], e->d_task_cell_ptr + t0,
                                              e->task_cell_ptr[t0], nb, e->st);
                    CK(cudaGetLastError());
                }
            }
            else {
                bs.b.slim = 0; bs.slim = false;
                const int ph2 = e->phase_begin(2);
                launch_sssp_bundle_finish(bs.b, e->st);
                e->phase_end(ph2, 1);
                CK(cudaGetLastError());
            }
        }
        if (!done) {
            if (sssp_complete_batch(e, bs, nb, want_count)) return -1;
            e->join_side();
            const int ph = e->phase_begin(4);
            launch_backtrace(c, e->st);
            e->phase_end(ph, 1);
            CK(cudaGetLastError());
        }
    }
    return 0;
}

static void iteration_row(const dtl_engine* e, double sysTT, double least, double row[5]);

static int iteration_impl(dtl_engine* e, int k, double row[5])
{
    CK(cudaSetDevice(e->dev));
    const int ph7 = e->phase_begin(7);
    if (do_vdf(e, false, true, nullptr)) return -1;                    // main_api.cpp:606
    if (do_scatter(e, k, false, false, true)) return -1;               // main_api.cpp:618 (side stream: joined before the back-trace)
    const bool want_count = e->counted_edges_cached < 0;
    if (want_count) CK(cudaMemsetAsync(&e->d_stats->counted_edges, 0, 8, e->st));
    static const bool k1_timeline = getenv("DTL_K1_TIMELINE") != nullptr; // -DSSSP_PROFILE build: globaltimer stamps of the bundle search
    if (k1_timeline) {
        SsspStats z{};
        CK(cudaMemcpyAsync(&z, e->d_stats, sizeof(z), cudaMemcpyDeviceToHost, e->st));
        if (sync_stream(e)) return -1;
        z.stage[4] = ~0ULL; z.stage[7] = ~0ULL; z.stage[5] = 0; z.stage[6] = 0; z.stage[3] = 0;
        for (int j = 0; j < 8; ++j) z.prof[j] = 0;
        CK(cudaMemcpyAsync(e->d_stats, &z, sizeof(z), cudaMemcpyHostToDevice, e->st));
        if (sync_stream(e)) return -1;
    }
    e->count_pass = want_count; // (the edge count of the trees is that of complete trees)
    const int rc_tc = trees_and_columns(e, k, want_count);             // main_api.cpp:647-699
    e->count_pass = false;
    if (rc_tc) return -1;
    e->join_side();
    if (k1_timeline) {
        SsspStats z{};
        CK(cudaMemcpyAsync(&z, e->d_stats, sizeof(z), cudaMemcpyDeviceToHost, e->st));
        if (sync_stream(e)) return -1;
        if (z.stage[3] > 0)
            fprintf(stderr, "K1 bundle search, thread-0 cycles per bundle: init+seed %.4g, window minimum %.4g, window distribute %.4g, drain %.4g | "
                            "windows per bundle %.1f, far entries scanned per window %.0f\n",
                    (double)z.prof[1] / z.stage[3], (double)z.prof[2] / z.stage[3], (double)z.prof[3] / z.stage[3], (double)z.prof[4] / z.stage[3],
                    (double)z.prof[6] / z.stage[3], (double)z.prof[5] / std::max(1.0, (double)z.prof[6]));
        if (z.stage[6] > 0)
            fprintf(stderr, "K1 timeline, iteration %d: first search ended %.3f ms after the first CTA started, last search %.3f ms, last CTA left %.3f ms\n", k,
                    (double)(z.stage[7] - z.stage[4]) * 1e-6, (double)(z.stage[5] - z.stage[4]) * 1e-6, (double)(z.stage[6] - z.stage[4]) * 1e-6);
    }
    launch_sum(e->d_least, e->n_cells, 1, e->d_scalars + 1, e->st);
    CK(cudaGetLastError());
    if (all_reduce(e, e->d_scalars + 1, 1, ncclFloat64)) return -1;
    if (want_count) {
        // counted edges: reachability does not change between iterations, so count once
        unsigned long long ce = 0;
        CK(cudaMemcpyAsync(&ce, &e->d_stats->counted_edges, 8, cudaMemcpyDeviceToHost, e->st));
        if (sync_stream(e)) return -1;
        e->h_scalars[8] = (double)ce;
        CK(cudaMemcpyAsync(e->d_scalars + 2, e->h_scalars + 8, 8, cudaMemcpyHostToDevice, e->st));
        if (all_reduce(e, e->d_scalars + 2, 1, ncclFloat64)) return -1;
        CK(cudaMemcpyAsync(e->h_scalars + 8, e->d_scalars + 2, 8, cudaMemcpyDeviceToHost, e->st));
        if (sync_stream(e)) return -1;
        e->counted_edges_cached = e->h_scalars[8];
    }
    CK(cudaMemcpyAsync(e->h_scalars, e->d_scalars, 2 * sizeof(double), cudaMemcpyDeviceToHost, e->st));
    e->phase_end(ph7, 0);
    if (check_flags(e, true)) return -1;
    e->resolve_timing();
    if (row) iteration_row(e, e->h_scalars[0], e->h_scalars[1], row);
    return 0;
}

// ---- pipelined iterations --------------------------------------------------------------------------------------------------
// The back-trace of iteration k (and its sums, scalar all-reduce and flag checks) only feeds the scatter of iteration k+1, and the
// search of iteration k+1 only needs the link volumes the scatter of iteration k left -- so a run of iterations is issued as
//     main stream:  K4(k) K1(k)            K4(k+1) K1(k+1)                    K4(k+2) K1(k+2) ...
//     side stream:        K3(k) helpers(k)         K2a(k) sums(k) K3(k+1) helpers(k+1)        K2a(k+1) ...
// with the trees of consecutive iterations in two sets of buffers (dtl_engine::flip_trees).  The host checks the flags of iteration k
// (ties, abandoned bundles, errors) while K1(k+1) runs and only then issues K3(k+1), so a batch that has to be redone is redone
// before anything reads its columns.  Same kernels on the same data in an order that respects every dependency: the results are
// those of dtl_iteration called n times (tests/test_engine_gpu.py::test_pipelined_iterations_*).
static bool pipeline_possible(const dtl_engine* e)
{
    if (!e->overlap || !e->st2 || e->opt.keep_trees || e->sa_mode || e->counted_edges_cached < 0 || getenv("DTL_NO_PIPELINE")) return false;
    const int n_tasks = (int)e->task_at.size();
    if (n_tasks == 0 || n_tasks > e->batch_tasks) return false;
    const int fsi = e->task_at[0] * e->T + e->task_tau[0];
    for (int t = 1; t < n_tasks; ++t)
        if (e->task_at[t] * e->T + e->task_tau[t] != fsi) return false; // one batch per iteration only
    return true;
}

static void iteration_row(const dtl_engine* e, double sysTT, double least, double row[5])
{
    row[0] = sysTT;
    row[1] = least;
    row[2] = (sysTT - least) / std::max(0.00001, least);           // main_api.cpp:711
    const float denom = 1 > e->total_demand ? (float)1 : e->total_demand; // main_api.cpp:716
    row[3] = sysTT / denom;
    row[4] = e->counted_edges_cached;
}

static int iterations_impl(dtl_engine* e, int k0, int n, double* rows)
{
    CK(cudaSetDevice(e->dev));
    int k = k0;
    const int k_end = k0 + n;
    // one at a time until the pipeline can start (the first iteration of an engine counts the edges of its trees)
    while (k < k_end && !(k_end - k >= 2 && pipeline_possible(e))) {
        if (iteration_impl(e, k, rows ? rows + (size_t)(k - k0) * 5 : nullptr)) return -1;
        ++k;
    }
    if (k >= k_end) return 0;

    const int n_tasks = (int)e->task_at.size();
    const int fsi = fs_index(e, e->task_at[0], e->task_tau[0]);
    const ForwardStar& f = e->fs[fsi];
    if (!e->alt_ready) {
        for (auto& ev : e->ev_k1) CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&e->ev_vdf, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&e->ev_done, cudaEventDisableTiming));
        // the second set of tree buffers: node-major labels / predecessor records are allocated by the launch that needs them; per-tree
        // arrays only when the searches write them (the single-origin kernels: an origin shard of a multi-GPU run) -- a bundle batch
        // touches them only when it is redone, and that happens with the main stream drained of anything that uses them
        SsspLaunch probe = make_sssp_launch(e, f, n_tasks, e->d_task_origin, e->d_task_zone, 0);
        const bool slim = probe.smq == 2 && !getenv("DTL_NO_SLIM_TREES");
        if (slim) {
            e->alt.d_labels = e->d_labels; e->alt.d_pred = e->d_pred;
        } else {
            if (e->alloc(&e->alt.d_labels, (size_t)e->batch_tasks * e->N)) return -1;
            if (e->alloc(&e->alt.d_pred, (size_t)e->batch_tasks * e->N)) return -1;
        }
        if (e->alloc(&e->alt.d_tie_nodes, e->batch_tasks)) return -1;
        if (e->alloc(&e->d_go, 4)) return -1;
        CK(cudaMemset(e->d_go, 0, 4 * sizeof(unsigned)));
        cudaDriverEntryPointQueryResult q1, q2;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &e->fn_wait_value, cudaEnableDefault, &q1) != cudaSuccess || q1 != cudaDriverEntryPointSuccess)
            e->fn_wait_value = nullptr;
        if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &e->fn_write_value, cudaEnableDefault, &q2) != cudaSuccess || q2 != cudaDriverEntryPointSuccess)
            e->fn_write_value = nullptr;
        cudaGetLastError();
        e->alt_ready = true;
    }
    struct Guard { // whatever happens, the engine leaves this function in its one-iteration-at-a-time state
        dtl_engine* e; cudaStream_t st, st2; bool ok;
        ~Guard()
        {
            e->st = st; e->st2 = st2; e->pipelined = e->expect_side = e->defer_helpers = false;
            e->next_started_word = nullptr;
            if (!ok && e->fn_write_value && e->d_go) {
                // a failure may have left a stream parked at a gate that nobody will open any more: open them all (the latest tag is
                // the only one a wait can still be parked on), so that dtl_destroy can drain the streams
                typedef int (*stream_value_fn)(cudaStream_t, unsigned long long, unsigned, unsigned);
                for (int i = 0; i < 4; ++i) ((stream_value_fn)e->fn_write_value)(st2, (unsigned long long)(uintptr_t)(e->d_go + i), e->pipe_tag, 0u);
            }
        }
    } guard{ e, e->st, e->st2, false };
    e->pipelined = e->expect_side = e->defer_helpers = true;
    const int ph7 = e->phase_begin(7);
    BatchState bs[2];
    auto helpers_and_join = [&](BatchState& b) -> int { // on the side stream, behind the scatter
        if (b.helper_grid > 0) {
            launch_bundle_pred_helpers(b.b, b.helper_grid, e->st2);
            CK(cudaGetLastError());
            CK(cudaEventRecord(e->ev_join, e->st2));
        }
        return 0;
    };
    auto vdf_and_search = [&](int kk, BatchState& b, int slot) -> int { // on the main stream: K4(kk), K1(kk)
        if (do_vdf(e, false, true, nullptr)) return -1;                                        // main_api.cpp:606 (joins the side stream)
        CK(cudaMemcpyAsync(e->d_scalars + 10 + slot, e->d_scalars, sizeof(double), cudaMemcpyDeviceToDevice, e->st)); // system TT of kk
        CK(cudaEventRecord(e->ev_vdf, e->st));
        e->in_loop = true;
        const int rc_l = sssp_launch_batch(e, f, n_tasks, e->d_task_origin, e->d_task_zone, 0, e->h_task_origin.data(), 0, true, b);
        e->in_loop = false;
        if (rc_l) return -1;
        CK(cudaEventRecord(e->ev_k1[slot], e->st));
        (void)kk;
        return 0;
    };
    // prologue: K4(k), K1(k) on the main stream, K3(k) + helpers on the side stream
    if (vdf_and_search(k, bs[0], 0)) return -1;
    if (do_scatter(e, k, false, false, true, e->ev_vdf)) return -1;                            // main_api.cpp:618
    if (helpers_and_join(bs[0])) return -1;
    // The slow paths of a batch (an order-dependent tie, an abandoned bundle, overflowed queues) re-read the link costs of iteration k,
    // which K4(k+1) rewrites.  So everything of iteration k+1 is issued behind a GATE on the main stream: a one-block kernel behind
    // K1(k) writes the iteration's tag into a device word when the search raised no flag (0 otherwise) and the stream waits for that
    // value (cuStreamWaitValue32) -- no host round trip between two searches (measured: 0.29-0.44 ms of an 8.7 ms iteration with the
    // host in between).  When a flag was raised the main stream stays parked at the gate while the side stream redoes the batch with
    // the costs of iteration k, and the write that ends the completion of iteration k (cuStreamWriteValue32) opens it.  Without stream
    // memory operations the host looks at the flags itself before it issues iteration k+1.
    typedef int (*stream_value_fn)(cudaStream_t, unsigned long long, unsigned, unsigned);
    const stream_value_fn wait_value = (stream_value_fn)e->fn_wait_value, write_value = (stream_value_fn)e->fn_write_value;
    const bool gated = wait_value && write_value && e->d_go && !getenv("DTL_PIPELINE_HOST_GATE");
    std::vector<int> h_flags;
    for (int cur = 0; k < k_end; ++k, cur ^= 1) {
        const bool last = k + 1 == k_end;
        const unsigned tag = ++e->pipe_tag ? e->pipe_tag : ++e->pipe_tag;
        bool clean = false;
        // An origin-bundle search leaves the side stream only the SMs without a bundle, and there the scatter alone lasts most of a
        // search: the node-major back-trace then stays on the MAIN stream, between K1(k) and the gate (it does nothing when the
        // search raised a flag), and only its sums and checks go to the side stream.  (With it on the side stream the side
        // stream is as long as the search: 8.80 ms per iteration against 9.03 without the pipeline.)
        // (bundles of 16 = 8 and more trees per SM, the whole demand on one GPU: measured 8.82 / 8.80 ms per iteration with the back-trace
        // on the main / side stream; a rank of 2, bundles of 8: 6.12 / 5.79 ms -- there the side stream has room)
        const bool bt_on_main = gated && bs[cur].slim && bs[cur].b.B >= 16 && !getenv("DTL_PIPELINE_BT_ON_SIDE");
        if (bt_on_main) {
            if (issue_slim_backtrace(e, bs[cur], 0, n_tasks, 0, k)) return -1;
            CK(cudaEventRecord(e->ev_k1[cur], e->st)); // what the side stream waits for is now the back-trace
        }
        bool side_waits = false; // the back-trace of iteration k waits until the search of iteration k+1 is resident
        if (!last && gated) {
            // (an origin-bundle search needs a whole SM per CTA: started at the end of K1(k), the back-trace's CTAs sit on every SM
            // when K1(k+1) arrives, and the VDF pass in between takes 0.24 instead of 0.04 ms)
            side_waits = !bt_on_main && bs[cur].bundle && bs[cur].b.fuse_pred && !getenv("DTL_PIPELINE_NO_START_ORDER");
            launch_set_go(bs[cur].slim ? e->d_batch_flag : e->d_tie_nodes, bs[cur].slim ? 1 : n_tasks, e->d_go + cur, e->d_go + 2 + cur, tag,
                          side_waits ? 0 : 1, e->st);
            CK(cudaGetLastError());
            e->next_started_word = side_waits ? e->d_go + 2 + cur : nullptr;
            e->next_started_tag = tag;
            if (wait_value(e->st, (unsigned long long)(uintptr_t)(e->d_go + cur), tag, 1u /* CU_STREAM_WAIT_VALUE_EQ */) != 0) {
                set_err("cuStreamWaitValue32 failed");
                return -1;
            }
            clean = true; // as far as issuing goes: what follows on the main stream runs only when the gate opens
        } else if (!last) {
            CK(cudaEventSynchronize(e->ev_k1[cur]));
            if (bs[cur].slim) {
                h_flags.assign(1, 0);
                CK(cudaMemcpy(h_flags.data(), e->d_batch_flag, sizeof(int), cudaMemcpyDeviceToHost));
            } else {
                h_flags.assign(n_tasks, 0);
                CK(cudaMemcpy(h_flags.data(), e->d_tie_nodes, sizeof(int) * n_tasks, cudaMemcpyDeviceToHost));
            }
            clean = true;
            for (int v : h_flags) clean = clean && v == 0;
        }
        if (!last && clean) {
            e->flip_trees();
            const int rc = vdf_and_search(k + 1, bs[cur ^ 1], cur ^ 1);
            e->flip_trees();
            e->next_started_word = nullptr;
            if (rc) return -1;
            if (side_waits && !bs[cur ^ 1].b.started_word) side_waits = false; // the next launch is not one that announces itself
        }
        // ---- iteration k is completed on the side stream (the helper code below issues on e->st)
        std::swap(e->st, e->st2);
        e->side_pending = false;
        // re-run / replay kernels share counters with a search of iteration k+1 that is already running: only possible with the host
        // gate (behind the device gate nothing of iteration k+1 runs while a batch is redone)
        e->pipelined = clean && !gated;
        int rc = 0;
        do {
            if (cudaStreamWaitEvent(e->st, e->ev_k1[cur], 0) != cudaSuccess) { set_err("cudaStreamWaitEvent failed"); rc = -1; break; }
            if (side_waits && wait_value(e->st, (unsigned long long)(uintptr_t)(e->d_go + 2 + cur), tag, 1u) != 0) {
                set_err("cuStreamWaitValue32 failed"); rc = -1; break;
            }
            if ((rc = columns_of_batch(e, bs[cur], 0, n_tasks, 0, k, false))) break;           // main_api.cpp:647-699
            if (!last && gated && write_value(e->st, (unsigned long long)(uintptr_t)(e->d_go + cur), tag, 0u) != 0) {
                set_err("cuStreamWriteValue32 failed"); rc = -1; break;
            }
            launch_sum(e->d_least, e->n_cells, 1, e->d_scalars + 1, e->st);
            if ((rc = all_reduce(e, e->d_scalars + 1, 1, ncclFloat64))) break;
            if (cudaMemcpyAsync(e->h_scalars + 1, e->d_scalars + 1, sizeof(double), cudaMemcpyDeviceToHost, e->st) != cudaSuccess ||
                cudaMemcpyAsync(e->h_scalars, e->d_scalars + 10 + cur, sizeof(double), cudaMemcpyDeviceToHost, e->st) != cudaSuccess) {
                set_err("cudaMemcpyAsync failed"); rc = -1; break;
            }
            rc = check_flags(e, true);
        } while (0);
        std::swap(e->st, e->st2);
        e->pipelined = true;
        if (rc) {
            // leave nothing parked at a gate that will never open
            if (gated && !last) {
                write_value(e->st2, (unsigned long long)(uintptr_t)(e->d_go + 2 + cur), tag, 0u);
                write_value(e->st2, (unsigned long long)(uintptr_t)(e->d_go + cur), tag, 0u);
            }
            return -1;
        }
        if (rows) iteration_row(e, e->h_scalars[0], e->h_scalars[1], rows + (size_t)(k - k0) * 5);
        if (!last && !clean) { // the batch was redone with the costs of iteration k: only now may they be rewritten
            CK(cudaEventRecord(e->ev_done, e->st2));
            CK(cudaStreamWaitEvent(e->st, e->ev_done, 0));
            e->flip_trees();
            const int rc2 = vdf_and_search(k + 1, bs[cur ^ 1], cur ^ 1);
            e->flip_trees();
            if (rc2) return -1;
        }
        if (!last) {
            if (do_scatter(e, k + 1, false, false, true, e->ev_vdf)) return -1;                // behind K2a(k) on the side stream, after K4(k+1)
            if (helpers_and_join(bs[cur ^ 1])) return -1;
            e->flip_trees(); // the set of iteration k+1 is the current one from here on
        }
    }
    CK(cudaEventRecord(e->ev_done, e->st2));
    CK(cudaStreamWaitEvent(e->st, e->ev_done, 0));
    e->side_pending = false;
    e->phase_end(ph7, 0);
    guard.ok = true;
    if (sync_stream(e)) return -1;
    e->resolve_timing();
    e->trees_valid = false;
    return 0;
}

// a rank that fails inside the loop aborts the communicator on its way out, so that its peers' collectives end (see comm_abort)
int dtl_iteration(dtl_engine* e, int k, double row[5])
{
    const int rc = iteration_impl(e, k, row);
    if (rc) comm_abort(e);
    return rc;
}

int dtl_iterations(dtl_engine* e, int k0, int n, double* rows)
{
    const int rc = iterations_impl(e, k0, n, rows);
    if (rc) comm_abort(e);
    return rc;
}

static int column_update_impl(dtl_engine* e, int n, double row[4], bool sa = false)
{
    CK(cudaSetDevice(e->dev));
    const int ph7 = e->phase_begin(7);
    // SA stage: the link state this call leaves is what the writers print after the stage, so person volumes and the detail
    // outputs of the VDF are kept up to date as well
    if (do_scatter(e, -1, sa, sa)) return -1;                          // assignment.h:575
    if (do_vdf(e, sa, true, nullptr)) return -1;                       // assignment.h:593
    PoolUpdateLaunch a{};
    a.L = e->L; a.T = e->T; a.AT = e->AT; a.pool = e->pool;
    a.od_at = e->d_od_at; a.od_tau = e->d_od_tau; a.od_vol = e->d_od_vol;
    a.tt = e->ls.tt; a.penalty = e->vp.penalty; a.toll = e->d_toll; a.vot = e->d_vot; a.n = n; a.od_sums = e->d_od_sums;
    a.link_cost = e->d_link_cost;
    a.sa = sa ? 1 : 0; a.design_flag = e->d_design_flag; a.rt_info = e->d_rt_info; a.od_impact = e->d_od_impact;
    const int ph5 = e->phase_begin(5);
    launch_pool_update(a, e->st);
    for (int j = 0; j < 4; ++j) launch_sum(e->d_od_sums + j, e->n_cells, 4, e->d_scalars + 4 + j, e->st);
    e->phase_end(ph5, 5);
    CK(cudaGetLastError());
    if (all_reduce(e, e->d_scalars + 4, 4, ncclFloat64)) return -1;
    CK(cudaMemcpyAsync(e->h_scalars + 4, e->d_scalars + 4, 4 * sizeof(double), cudaMemcpyDeviceToHost, e->st));
    e->phase_end(ph7, 0);
    if (sync_stream(e)) return -1;
    e->resolve_timing();
    if (row) {
        const double gap = e->h_scalars[4], cost = e->h_scalars[5], time = e->h_scalars[6], dem = e->h_scalars[7];
        row[0] = time / std::max(0.001, dem);                          // assignment.h:907
        row[1] = gap;
        row[2] = gap * 100.0 / std::max(0.00001, cost);                // assignment.h:911
        row[3] = dem;
    }
    return 0;
}

int dtl_column_update(dtl_engine* e, int n, double row[4])
{
    const int rc = column_update_impl(e, n, row);
    if (rc) comm_abort(e);
    return rc;
}

// ---- sensitivity-analysis stage: main_api.cpp:771-1042 -------------------------------------------------------------------
// What is covered: supply-side "sa" lane changes (scenario_API.h:152-160, applied at main_api.cpp:779-794), the column
// re-generation loop that re-routes the agent types with real-time information (:864-1034: K4, K3 with the per-agent-type decay
// rule, K1 + K2a for their origins only) and the column-pool iterations under the SA rules (:1040, assignment.h:643-846).
// Not covered: dms scenario rows / information zones / g_column_pool_route_modification, demand-side scenario factors.
static int sa_configure_impl(dtl_engine* e, const int* rt_info, int n_changes, const int* tau, const int* link, const double* lanes)
{
    CK(cudaSetDevice(e->dev));
    e->sa_rt_info.assign(rt_info, rt_info + e->AT);
    for (int i = 0; i < n_changes; ++i)
        if (tau[i] < 0 || tau[i] >= e->T || link[i] < 0 || link[i] >= e->L) { set_err("lane change %d: period or link out of range", i); return -1; }
    e->sa_tau.assign(tau, tau + n_changes); e->sa_link.assign(link, link + n_changes); e->sa_lanes.assign(lanes, lanes + n_changes);
    const size_t TL = (size_t)e->T * e->L;
    std::vector<signed char> flag(std::max<size_t>(1, TL), 0);
    for (int i = 0; i < n_changes; ++i) flag[(size_t)tau[i] * e->L + link[i]] = -1;
    std::vector<unsigned char> decay(std::max(1, e->AT));
    for (int at = 0; at < e->AT; ++at) decay[at] = rt_info[at] == 1;
    if (!e->d_design_flag) {
        if (e->alloc(&e->d_design_flag, TL) || e->alloc(&e->d_sa_decay_at, (size_t)e->AT) || e->alloc(&e->d_rt_info, (size_t)e->AT) ||
            e->alloc(&e->d_od_impact, (size_t)std::max(1, e->n_cells)))
            return -1;
    }
    CK(cudaMemcpy(e->d_design_flag, flag.data(), TL, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(e->d_sa_decay_at, decay.data(), (size_t)e->AT, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(e->d_rt_info, e->sa_rt_info.data(), sizeof(int) * e->AT, cudaMemcpyHostToDevice));
    CK(cudaMemset(e->d_od_impact, 0, (size_t)std::max(1, e->n_cells)));
    e->sa_hist_time.clear(); e->sa_hist_volume.clear();
    e->dms_tau.clear(); e->dms_link.clear(); e->dms_zone.clear(); e->info_zone.clear();
    e->sa_slot_impact.clear(); e->sa_cell_info.clear();
    e->d_sa_decay_cell = nullptr; e->d_sa_task_skip = nullptr; // (slab memory: the arrays of an earlier configuration stay allocated)
    return 0;
}

// "dms" rows of supply_side_scenario.csv (scenario_API.h:165-188), after dtl_sa_configure: the link becomes a message sign
// (network_design_flag 2), the zone of its to-node an information zone.  In the SA stage a message-sign user (real_time_information 2)
// is re-routed -- MSA decay of its columns, new shortest path -- only from an origin that is an information zone
// (assignment.h:160-170, main_api.cpp:990-994): that is a per-cell / per-task predicate, not a per-agent-type one.
static int sa_set_dms_impl(dtl_engine* e, int n, const int* tau, const int* link, const int* zone)
{
    CK(cudaSetDevice(e->dev));
    if (e->sa_rt_info.empty()) { set_err("dtl_sa_configure has not been called"); return -1; }
    if (e->world > 1 && !e->local_shard) {
        set_err("dms scenario rows need the whole column pool on one GPU (the route modification splices columns of other origins)");
        return -1;
    }
    for (int i = 0; i < n; ++i)
        if (tau[i] < 0 || tau[i] >= e->T || link[i] < 0 || link[i] >= e->L || zone[i] >= e->Z) { set_err("dms row %d: period, link or zone out of range", i); return -1; }
    e->dms_tau.assign(tau, tau + n); e->dms_link.assign(link, link + n); e->dms_zone.assign(zone, zone + n);
    e->info_zone.assign(std::max(1, e->Z), 0);
    for (int i = 0; i < n; ++i) {
        const signed char two = 2;
        CK(cudaMemcpy(e->d_design_flag + (size_t)tau[i] * e->L + link[i], &two, 1, cudaMemcpyHostToDevice));
        if (zone[i] >= 0) e->info_zone[zone[i]] = 1;
    }
    const int n_tasks = (int)e->task_at.size();
    std::vector<unsigned char> decay(std::max(1, e->n_cells), 0), skip(std::max(1, n_tasks), 0);
    for (int c = 0; c < e->n_cells; ++c) {
        const int rt = e->sa_rt_info[e->h_od_at[c]];
        decay[c] = rt == 1 || (rt == 2 && e->info_zone[e->task_zone[e->h_od_task[c]]]);
    }
    for (int t = 0; t < n_tasks; ++t)
        skip[t] = e->h_task_skip[t] || (e->sa_rt_info[e->task_at[t]] == 2 && !e->info_zone[e->task_zone[t]]);
    if (e->alloc(&e->d_sa_decay_cell, decay.size()) || e->alloc(&e->d_sa_task_skip, skip.size())) return -1;
    CK(cudaMemcpy(e->d_sa_decay_cell, decay.data(), decay.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(e->d_sa_task_skip, skip.data(), skip.size(), cudaMemcpyHostToDevice));
    return 0;
}

static int sa_begin_impl(dtl_engine* e, double* sys_tt)
{
    CK(cudaSetDevice(e->dev));
    if (e->sa_rt_info.empty()) { set_err("dtl_sa_configure has not been called"); return -1; }
    for (size_t i = 0; i < e->sa_link.size(); ++i) {                   // main_api.cpp:779-794
        const size_t a = (size_t)e->sa_tau[i] * e->L + e->sa_link[i];
        bool seen = false; // two rows for one (period, link): the flag is set once, the last row's lanes count
        for (size_t j = i + 1; j < e->sa_link.size(); ++j) seen = seen || (e->sa_tau[j] == e->sa_tau[i] && e->sa_link[j] == e->sa_link[i]);
        if (seen) continue;
        double nl = 0, cap = 0;
        CK(cudaMemcpy(&nl, e->vp.nlanes + a, 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(&cap, e->vp.cap + a, 8, cudaMemcpyDeviceToHost));
        const float old_lanes = (float)nl;                              // :786
        nl = (double)(float)e->sa_lanes[i];                             // :787 (sa_lanes_change is parsed as a float)
        cap *= nl / std::max(0.00001, (double)old_lanes);               // :788
        CK(cudaMemcpy(e->vp.nlanes + a, &nl, 8, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(e->vp.cap + a, &cap, 8, cudaMemcpyHostToDevice));
        if (nl < 0.01) { const double f = 1440; CK(cudaMemcpy(e->vp.fftt + a, &f, 8, cudaMemcpyHostToDevice)); } // :789-790
    }
    if (do_vdf(e, false, true, nullptr)) return -1;                    // :857
    CK(cudaMemcpyAsync(e->h_scalars, e->d_scalars, sizeof(double), cudaMemcpyDeviceToHost, e->st));
    if (sync_stream(e)) return -1;
    e->resolve_timing();
    if (sys_tt) *sys_tt = e->h_scalars[0];
    return 0;
}

static int sa_iteration_impl(dtl_engine* e, int it, double row[2])
{
    CK(cudaSetDevice(e->dev));
    if (e->sa_rt_info.empty()) { set_err("dtl_sa_configure has not been called"); return -1; }
    const int ph7 = e->phase_begin(7);
    launch_induced_delay_flags(e->d_design_flag, e->ls.doc, (size_t)e->T * e->L, e->st); // main_api.cpp:877-891 (DOC of the last VDF pass)
    CK(cudaGetLastError());
    // (person volumes and the detail outputs of the VDF are refreshed too: with S = 0 this is the state the writers print)
    if (do_vdf(e, true, true, nullptr)) return -1;                     // main_api.cpp:897
    if (do_scatter(e, it, true, true)) return -1;                      // :913
    e->sa_mode = true;
    const int rc = trees_and_columns(e, it, false);                    // :966-1018
    e->sa_mode = false;
    if (rc) return -1;
    CK(cudaMemcpyAsync(e->h_scalars, e->d_scalars, sizeof(double), cudaMemcpyDeviceToHost, e->st));
    e->phase_end(ph7, 0);
    if (check_flags(e, true)) return -1;
    e->resolve_timing();
    if (row) {
        const float denom = 1 > e->total_demand ? (float)1 : e->total_demand; // :1027
        row[0] = e->h_scalars[0];
        row[1] = e->h_scalars[0] / denom;
    }
    return 0;
}

struct PoolHost {
    std::vector<int> count, node_sum, n_links, n_nodes;
    std::vector<long long> offset;
    std::vector<double> volume, cost, time;
    std::vector<int> seq; // path_seq_no: the slot index unless the pool carries its own array
};
static int pool_to_host(dtl_engine* e, PoolHost& h)
{
    const size_t n = e->n_cells, s = n * e->pool.cap;
    h.count.resize(n); h.node_sum.resize(s); h.n_links.resize(s); h.n_nodes.resize(s); h.offset.resize(s);
    h.volume.resize(s); h.cost.resize(s); h.time.resize(s);
    if (!n) return 0;
    CK(cudaMemcpyAsync(h.count.data(), e->pool.count, n * 4, cudaMemcpyDeviceToHost, e->st));
    CK(cudaMemcpyAsync(h.node_sum.data(), e->pool.node_sum, s * 4, cudaMemcpyDeviceToHost, e->st));
    CK(cudaMemcpyAsync(h.n_links.data(), e->pool.n_links, s * 4, cudaMemcpyDeviceToHost, e->st));
    CK(cudaMemcpyAsync(h.n_nodes.data(), e->pool.n_nodes, s * 4, cudaMemcpyDeviceToHost, e->st));
    CK(cudaMemcpyAsync(h.offset.data(), e->pool.offset, s * 8, cudaMemcpyDeviceToHost, e->st));
    CK(cudaMemcpyAsync(h.volume.data(), e->pool.volume, s * 8, cudaMemcpyDeviceToHost, e->st));
    CK(cudaMemcpyAsync(h.cost.data(), e->pool.cost, s * 8, cudaMemcpyDeviceToHost, e->st));
    CK(cudaMemcpyAsync(h.time.data(), e->pool.time, s * 8, cudaMemcpyDeviceToHost, e->st));
    h.seq.resize(s);
    if (e->pool.seq) CK(cudaMemcpyAsync(h.seq.data(), e->pool.seq, s * 4, cudaMemcpyDeviceToHost, e->st));
    if (sync_stream(e)) return -1;
    if (!e->pool.seq)
        for (size_t i = 0; i < s; ++i) h.seq[i] = (int)(i % (size_t)e->pool.cap);
    return 0;
}

// g_column_pool_route_modification (assignment.h:1077-1383), stage II of the SA stage (main_api.cpp:1036): a column of a message-sign
// user that passes a message sign AND a link with a lane change is split -- a share exp(-0.1 t_base) / (exp(-0.1 t_base) + exp(-0.1 t_alt))
// stays, the rest goes to a new column: the old path up to the sign, then the first column of (information zone -> same destination)
// that avoids every impacted link and beats the non-diverted time.  The reference walks std::maps that it inserts into while it
// iterates, accumulates path times on the information zone's columns across ALL origins (assignment.h:1229: "+=", never reset) and
// numbers new columns with the map size before the insertion: an order-dependent serial pass over a few thousand columns, done once.
// It runs on the host over the downloaded column pool -- there is nothing here for a GPU -- and uploads what changed.  Float
// arithmetic where the reference uses float.  n_split: columns that were split.
static int sa_route_modification_impl(dtl_engine* e, int* n_split)
{
    CK(cudaSetDevice(e->dev));
    if (n_split) *n_split = 0;
    if (e->sa_rt_info.empty()) { set_err("dtl_sa_configure has not been called"); return -1; }
    if (e->dms_link.empty()) return 0;
    const int L = e->L, Z = e->Z, AT = e->AT, T = e->T, cap = e->pool.cap;
    const size_t TL = (size_t)T * L, slots = (size_t)e->n_cells * cap;
    PoolHost h;
    if (pool_to_host(e, h)) return -1;
    unsigned long long used = 0;
    CK(cudaMemcpy(&used, e->pool.arena_used, 8, cudaMemcpyDeviceToHost));
    used = std::min<unsigned long long>(used, (unsigned long long)e->pool.arena_cap);
    std::vector<int> arena((size_t)used + 4);
    CK(cudaMemcpy(arena.data(), e->pool.arena, sizeof(int) * (size_t)used, cudaMemcpyDeviceToHost));
    const size_t used0 = (size_t)used;
    std::vector<double> tt(TL);
    std::vector<signed char> flag(TL);
    std::vector<unsigned char> od_impact(std::max(1, e->n_cells));
    CK(cudaMemcpy(tt.data(), e->ls.tt, TL * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(flag.data(), e->d_design_flag, TL, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(od_impact.data(), e->d_od_impact, (size_t)e->n_cells, cudaMemcpyDeviceToHost));
    std::map<std::pair<int, int>, int> dms_zone_of; // (period, link) -> information zone (-1: not a zone of the model)
    for (size_t i = 0; i < e->dms_link.size(); ++i) dms_zone_of[{ e->dms_tau[i], e->dms_link[i] }] = e->dms_zone[i];
    std::map<std::tuple<int, int, int, int>, int> cell_of; // (origin zone, destination zone, agent type, period) -> cell
    for (int c = 0; c < e->n_cells; ++c) cell_of[std::make_tuple(e->task_zone[e->h_od_task[c]], e->h_od_dzone[c], e->h_od_at[c], e->h_od_tau[c])] = c;
    auto by_key = [&](int cell) { // the cell's slots in std::map order (ascending node_sum)
        std::vector<int> o(h.count[cell]);
        std::iota(o.begin(), o.end(), 0);
        const size_t base = (size_t)cell * cap;
        std::sort(o.begin(), o.end(), [&](int x, int y) { return h.node_sum[base + x] < h.node_sum[base + y]; });
        return o;
    };
    int splits = 0;
    std::vector<int> seq;
    e->sa_slot_impact.assign(std::max<size_t>(1, slots), 0);
    e->sa_cell_info.assign(std::max(1, e->n_cells), 0);
    for (int o = 0; o < Z; ++o)
        for (int d = 0; d < Z; ++d)
            for (int at = 0; at < AT; ++at) {
                if (e->sa_rt_info[at] != 2) continue;                                       // :1110
                for (int tau = 0; tau < T; ++tau) {
                    auto ci = cell_of.find(std::make_tuple(o, d, at, tau));
                    if (ci == cell_of.end()) continue;                                      // od_volume > 0, :1115
                    const int cell = ci->second;
                    const size_t base = (size_t)cell * cap;
                    const signed char* fl = flag.data() + (size_t)tau * L;
                    const double* t_link = tt.data() + (size_t)tau * L;
                    bool have_key = false;
                    int last_key = 0;
                    for (;;) { // the std::map walk: a column inserted with a larger key is met later in the same walk
                        int j = -1;
                        for (int q = 0; q < h.count[cell]; ++q) {
                            const int k = h.node_sum[base + q];
                            if ((!have_key || k > last_key) && (j < 0 || k < h.node_sum[base + j])) j = q;
                        }
                        if (j < 0) break;
                        have_key = true; last_key = h.node_sum[base + j];
                        if (h.volume[base + j] < 0.00001) continue;                         // :1136
                        bool passing_info = false, passing_cap = false;
                        int new_orig = -1;
                        seq.clear();
                        const int* links = arena.data() + h.offset[base + j];
                        for (int nl = 0; nl < h.n_links[base + j]; ++nl) {                  // :1141
                            const int l = links[nl];
                            if (!passing_info && fl[l] == 2) {                              // :1146-1150 (a message sign's to-node carries a zone id: the reader insists)
                                od_impact[cell] = 1;                                        // :1153-1154
                                const auto z = dms_zone_of.find({ tau, l });
                                if (z == dms_zone_of.end() || z->second < 0) continue;     // :1158-1159
                                passing_info = true; new_orig = z->second;
                                seq.assign(links, links + nl + 1);                          // :1166-1170
                            }
                            if (fl[l] == -1) { passing_cap = true; e->sa_slot_impact[base + j] = 1; } // :1176-1180
                        }
                        if (!(passing_cap && passing_info)) continue;                       // :1184
                        auto c2i = cell_of.find(std::make_tuple(new_orig, d, at, tau));
                        if (c2i == cell_of.end() || h.count[c2i->second] == 0) {            // :1203-1208 (the reference stops)
                            set_err("route modification: no column from the information zone %d to zone %d in the column pool", new_orig, d);
                            return -1;
                        }
                        const int cell2 = c2i->second;
                        const size_t base2 = (size_t)cell2 * cap;
                        e->sa_cell_info[cell2] = 1;                                         // :1360
                        const std::vector<int> order2 = by_key(cell2);
                        float non_div = 999999;                                             // :1223
                        auto diverts = [&](int q) {
                            const int* l2 = arena.data() + h.offset[base2 + q];
                            for (int nl2 = 1; nl2 < h.n_links[base2 + q]; ++nl2)
                                if (fl[l2[nl2]] <= -1) return false;                        // :1238 (incl. induced delay, -2)
                            return true;
                        };
                        for (int q : order2) {                                              // :1225-1250
                            const int* l2 = arena.data() + h.offset[base2 + q];
                            for (int nl2 = 1; nl2 < h.n_links[base2 + q]; ++nl2) h.time[base2 + q] += t_link[l2[nl2]]; // accumulates, :1229
                            if (!diverts(q)) {
                                const float abs_br = 3, ratio_br = 0.15;
                                if (h.time[base2 + q] < non_div - abs_br && h.time[base2 + q] < (non_div * (1 - ratio_br))) non_div = (float)h.time[base2 + q];
                            }
                        }
                        bool found = false;
                        for (int q : order2) {                                              // :1254-1287
                            if (found || !diverts(q) || !(h.time[base2 + q] < non_div)) continue;
                            const int* l2 = arena.data() + h.offset[base2 + q];
                            seq.insert(seq.end(), l2 + 1, l2 + h.n_links[base2 + q]);
                            found = true;
                        }
                        if (!found) continue;
                        float base_tt = 0, alt_tt = 0;                                      // optional detour, :1329-1356
                        for (int nl0 = 0; nl0 < h.n_links[base + j]; ++nl0) base_tt += t_link[links[nl0]];
                        for (size_t nl3 = 1; nl3 < seq.size(); ++nl3) alt_tt += t_link[seq[nl3]];
                        const float beta = -0.1;
                        const float prob_base = expf(beta * base_tt) / (expf(beta * base_tt) + expf(beta * alt_tt));
                        const float base_volume = (float)h.volume[base + j];
                        h.volume[base + j] = base_volume * prob_base;
                        int node_sum = 0;                                                   // g_add_new_column, :1020-1053
                        for (int l : seq) node_sum = (int)((unsigned)node_sum + (unsigned)l + (unsigned)e->h_link_to[l]);
                        const int size_before = h.count[cell];
                        int slot = -1;
                        for (int q = 0; q < size_before; ++q)
                            if (h.node_sum[base + q] == node_sum) slot = q;
                        if (slot < 0) {
                            if (size_before >= cap) { set_err("an OD cell needs more than %d columns: raise dtl_options.max_columns_per_od", cap); return -1; }
                            slot = size_before;
                            h.count[cell] = size_before + 1;
                            h.node_sum[base + slot] = node_sum;
                            h.cost[base + slot] = 0; h.time[base + slot] = 0;
                        }
                        h.seq[base + slot] = size_before; // "map[key].path_seq_no = map.size()": the size is taken before the entry is created
                        const size_t need = (seq.size() + 3) & ~(size_t)3; // columns stay 16-byte aligned and padded with -1
                        if ((long long)(used + need) > e->pool.arena_cap) { set_err("column arena exhausted (%lld path links): raise dtl_options.column_arena_bytes", e->pool.arena_cap); return -1; }
                        arena.resize((size_t)used + need + 4, -1);
                        std::copy(seq.begin(), seq.end(), arena.begin() + (size_t)used);
                        std::fill(arena.begin() + (size_t)used + seq.size(), arena.begin() + (size_t)used + need, -1);
                        h.offset[base + slot] = (long long)used;
                        used += need;
                        h.n_links[base + slot] = (int)seq.size();
                        h.n_nodes[base + slot] = (int)seq.size() + 1;
                        h.volume[base + slot] = base_volume * (1 - prob_base);
                        e->sa_slot_impact[base + slot] = 3;                                 // :1356
                        ++splits;
                    }
                }
            }
    // upload what changed
    if (!e->pool.seq && e->alloc(&e->pool.seq, std::max<size_t>(1, slots))) return -1;
    CK(cudaMemcpy(e->pool.seq, h.seq.data(), slots * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(e->pool.count, h.count.data(), (size_t)e->n_cells * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(e->pool.node_sum, h.node_sum.data(), slots * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(e->pool.n_links, h.n_links.data(), slots * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(e->pool.n_nodes, h.n_nodes.data(), slots * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(e->pool.offset, h.offset.data(), slots * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(e->pool.volume, h.volume.data(), slots * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(e->pool.cost, h.cost.data(), slots * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(e->pool.time, h.time.data(), slots * 8, cudaMemcpyHostToDevice));
    if ((size_t)used > used0) CK(cudaMemcpy(e->pool.arena + used0, arena.data() + used0, sizeof(int) * ((size_t)used - used0), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(e->pool.arena_used, &used, 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(e->d_od_impact, od_impact.data(), (size_t)e->n_cells, cudaMemcpyHostToDevice));
    if (n_split) *n_split = splits;
    return 0;
}

static int sa_column_update_impl(dtl_engine* e, int n, double row[4])
{
    if (e->sa_rt_info.empty()) { set_err("dtl_sa_configure has not been called"); return -1; }
    if (column_update_impl(e, n, row, true)) return -1;
    // path_time_per_iteration_SA_map / path_volume_per_iteration_SA_map (assignment.h:741, 896-900): what route_assignment.csv
    // prints per SA iteration
    const size_t s = (size_t)e->n_cells * e->pool.cap;
    e->sa_hist_time.emplace_back(s); e->sa_hist_volume.emplace_back(s);
    if (s) {
        CK(cudaMemcpy(e->sa_hist_time.back().data(), e->pool.time, s * 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(e->sa_hist_volume.back().data(), e->pool.volume, s * 8, cudaMemcpyDeviceToHost));
    }
    return 0;
}

// ---- ODME: Assignment::Demand_ODME, ODME.h:91-516 ------------------------------------------------------------------------
static OdmeLaunch make_odme_launch(dtl_engine* e)
{
    OdmeLaunch a{};
    a.L = e->L; a.T = e->T; a.AT = e->AT; a.pool = e->pool; a.od_at = e->d_od_at; a.od_tau = e->d_od_tau;
    a.tt = e->ls.tt; a.link_volume = e->ls.link_volume; a.at_pce = e->d_at_pce; a.at_occ = e->d_at_occ;
    a.pce_fix = e->ls.pce_fix; a.person_fix = e->ls.person_fix; a.od_sums = e->d_od_sums;
    a.obs = e->d_obs; a.ub = e->d_ub; a.est_dev = e->d_est_dev;
    a.n_sensors = (int)e->odme_sensor_idx.size(); a.sensor_idx = e->d_sensor_idx; a.sensor_dev = e->d_sensor_dev;
    return a;
}

static int odme_configure_impl(dtl_engine* e, const float* at_pce, const float* at_occ, int n, const int* tau, const int* link,
                               const float* count, const int* ub)
{
    CK(cudaSetDevice(e->dev));
    const size_t TL = (size_t)e->T * e->L;
    e->odme_obs.assign(std::max<size_t>(1, TL), 0.0);
    e->odme_ub.assign(std::max<size_t>(1, TL), 1);                     // CPeriod_VDF(): upper_bound_flag 1, VDF.h:60
    for (int i = 0; i < n; ++i) {                                      // ODME.h:156-175, rows in file order
        if (tau[i] < 0 || tau[i] >= e->T || link[i] < 0 || link[i] >= e->L) { set_err("measurement %d: period or link out of range", i); return -1; }
        const size_t a = (size_t)tau[i] * e->L + link[i];
        if (e->odme_obs[a] >= 1) {
            if (ub[i] == 0) { e->odme_obs[a] = count[i]; e->odme_ub[a] = ub[i]; } // an actual count overwrites, an upper bound does not
        } else {
            e->odme_obs[a] = count[i]; e->odme_ub[a] = ub[i];
        }
    }
    e->odme_sensor_idx.clear();
    for (int l = 0; l < e->L; ++l)                                     // the order of the summary sums, assignment.h:465-469
        for (int t = 0; t < e->T; ++t)
            if (e->odme_obs[(size_t)t * e->L + l] >= 1) e->odme_sensor_idx.push_back(t * e->L + l);
    if (!e->d_obs) {
        if (e->alloc(&e->d_obs, TL) || e->alloc(&e->d_est_dev, TL) || e->alloc(&e->d_ub, TL) || e->alloc(&e->d_at_pce, (size_t)e->AT) ||
            e->alloc(&e->d_at_occ, (size_t)e->AT))
            return -1;
    }
    if (e->alloc(&e->d_sensor_idx, e->odme_sensor_idx.size()) || e->alloc(&e->d_sensor_dev, e->odme_sensor_idx.size())) return -1;
    CK(cudaMemcpy(e->d_obs, e->odme_obs.data(), TL * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(e->d_ub, e->odme_ub.data(), TL * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(e->d_est_dev, 0, TL * 8));
    CK(cudaMemcpy(e->d_at_pce, at_pce, sizeof(float) * e->AT, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(e->d_at_occ, at_occ, sizeof(float) * e->AT, cudaMemcpyHostToDevice));
    if (!e->odme_sensor_idx.empty())
        CK(cudaMemcpy(e->d_sensor_idx, e->odme_sensor_idx.data(), sizeof(int) * e->odme_sensor_idx.size(), cudaMemcpyHostToDevice));
    e->odme_ready = true;
    return 0;
}

// g_reset_and_update_link_volume_based_on_ODME_columns, assignment.h:263-563
static int odme_scatter_impl(dtl_engine* e, int s, double row[8])
{
    (void)s;
    CK(cudaSetDevice(e->dev));
    if (!e->odme_ready) { set_err("dtl_odme_configure has not been called"); return -1; }
    const size_t TL = (size_t)e->T * e->L;
    const OdmeLaunch a = make_odme_launch(e);
    const int ph5 = e->phase_begin(5);
    launch_odme_stats(a, e->st);                                       // on the link times of the previous pass (:378-430)
    for (int j = 0; j < 3; ++j) launch_sum(e->d_od_sums + j, e->n_cells, 4, e->d_scalars + 4 + j, e->st);
    e->phase_end(ph5, 5);
    CK(cudaGetLastError());
    if (all_reduce(e, e->d_scalars + 4, 3, ncclFloat64)) return -1;
    const int ph1 = e->phase_begin(1);
    CK(cudaMemsetAsync(e->ls.pce_fix, 0, TL * sizeof(unsigned long long), e->st));
    CK(cudaMemsetAsync(e->ls.person_fix, 0, TL * (1 + e->AT) * sizeof(unsigned long long), e->st));
    launch_odme_scatter(a, e->st);                                     // :433-457
    e->phase_end(ph1, 1);
    CK(cudaGetLastError());
    if (all_reduce(e, e->ls.pce_fix, TL, ncclUint64)) return -1;
    if (all_reduce(e, e->ls.person_fix, TL * (1 + e->AT), ncclUint64)) return -1;
    if (do_vdf(e, true, true, nullptr)) return -1;                     // :467
    launch_odme_dev(a, e->st);                                         // :476
    CK(cudaGetLastError());
    const size_t ns = e->odme_sensor_idx.size();
    std::vector<double> dev(std::max<size_t>(1, ns));
    if (ns) CK(cudaMemcpyAsync(dev.data(), e->d_sensor_dev, ns * 8, cudaMemcpyDeviceToHost, e->st));
    CK(cudaMemcpyAsync(e->h_scalars + 4, e->d_scalars + 4, 3 * sizeof(double), cudaMemcpyDeviceToHost, e->st));
    if (sync_stream(e)) return -1;
    e->resolve_timing();
    // the summary sums of the reference: float accumulators, links in ascending order (:465-515)
    float total_gap = 0, sub_link = 0, sub_system = 0;
    int n_links = 0;
    for (size_t i = 0; i < ns; ++i) {
        const int idx = e->odme_sensor_idx[i];
        const double d = dev[i], obs = e->odme_obs[idx];
        if (e->odme_ub[idx] == 0 || d > 0) {
            total_gap += abs((int)d);                                  // :491: abs() binds to the int overload in the reference
            sub_link += fabs(d / obs);
            sub_system += d / obs;
        }
        n_links += 1;
    }
    if (row) {
        const double demand = e->h_scalars[4], time = e->h_scalars[5], ue = e->h_scalars[6];
        const int denom = std::max(1, n_links);
        row[0] = sub_link / denom;                                     // :560
        row[1] = sub_system / denom;                                   // :561
        row[2] = total_gap / denom;
        row[3] = time / std::max(0.1, demand);
        row[4] = ue / std::max(0.00001, demand);
        row[5] = ue / std::max(0.00001, time) * 100;
        row[6] = demand;
        row[7] = n_links;
    }
    return 0;
}

static int odme_update_impl(dtl_engine* e)
{
    CK(cudaSetDevice(e->dev));
    if (!e->odme_ready) { set_err("dtl_odme_configure has not been called"); return -1; }
    const int ph5 = e->phase_begin(5);
    launch_odme_update(make_odme_launch(e), e->st);                    // ODME.h:262-488
    e->phase_end(ph5, 1);
    CK(cudaGetLastError());
    if (sync_stream(e)) return -1;
    e->resolve_timing();
    return 0;
}

static int finalize_impl(dtl_engine* e, int K, double* sys_tt)
{
    (void)K;
    CK(cudaSetDevice(e->dev));
    if (do_scatter(e, -1, true)) return -1;                            // main_api.cpp:752
    if (do_vdf(e, true, true, nullptr)) return -1;                     // main_api.cpp:759
    CK(cudaMemcpyAsync(e->h_scalars, e->d_scalars, sizeof(double), cudaMemcpyDeviceToHost, e->st));
    if (sync_stream(e)) return -1;
    e->resolve_timing();
    if (sys_tt) *sys_tt = e->h_scalars[0];
    return 0;
}

int dtl_finalize(dtl_engine* e, int K, double* sys_tt)
{
    const int rc = finalize_impl(e, K, sys_tt);
    if (rc) comm_abort(e);
    return rc;
}

int dtl_odme_configure(dtl_engine* e, const float* at_pce, const float* at_occ, int n, const int* tau, const int* link, const float* count,
                       const int* upper_bound_flag)
{
    return odme_configure_impl(e, at_pce, at_occ, std::max(0, n), tau, link, count, upper_bound_flag);
}
int dtl_odme_scatter(dtl_engine* e, int s, double row[8])
{
    const int rc = odme_scatter_impl(e, s, row);
    if (rc) comm_abort(e);
    return rc;
}
int dtl_odme_update(dtl_engine* e)
{
    const int rc = odme_update_impl(e);
    if (rc) comm_abort(e);
    return rc;
}
// the ODME loop of the reference (ODME.h:240-260, 490-512): scatter, stop when the link MAPE has grown by more than 0.01 after
// five iterations, path-flow step; one more scatter at the end.  rows (may be NULL): [n_iterations + 1][8], the last row is
// the closing scatter; *n_done = iterations whose path-flow step ran
int dtl_odme_run(dtl_engine* e, int n_iterations, double* rows, int* n_done)
{
    double prev_gap = 9999999;
    int done = 0;
    for (int s = 0; s < n_iterations; ++s) {
        double row[8];
        if (dtl_odme_scatter(e, s, row)) return -1;
        if (rows) memcpy(rows + (size_t)s * 8, row, sizeof(row));
        const double gap_increase = row[0] - prev_gap;
        if (s >= 5 && gap_increase > 0.01) break;
        prev_gap = row[0];
        if (dtl_odme_update(e)) return -1;
        done = s + 1;
    }
    double row[8];
    if (dtl_odme_scatter(e, n_iterations, row)) return -1;
    if (rows) memcpy(rows + (size_t)n_iterations * 8, row, sizeof(row));
    if (n_done) *n_done = done;
    return 0;
}
// observed count, upper-bound flag and count deviation of every link of a period (host arrays of length L, may be NULL)
int dtl_odme_get(dtl_engine* e, int tau, double* obs_count, int* upper_bound_flag, double* est_count_dev)
{
    if (!e->odme_ready) { set_err("dtl_odme_configure has not been called"); return -1; }
    if (tau < 0 || tau >= e->T) { set_err("period %d out of range", tau); return -1; }
    CK(cudaSetDevice(e->dev));
    const size_t off = (size_t)tau * e->L;
    if (obs_count) memcpy(obs_count, e->odme_obs.data() + off, sizeof(double) * e->L);
    if (upper_bound_flag) memcpy(upper_bound_flag, e->odme_ub.data() + off, sizeof(int) * e->L);
    if (est_count_dev) CK(cudaMemcpy(est_count_dev, e->d_est_dev + off, sizeof(double) * e->L, cudaMemcpyDeviceToHost));
    return 0;
}

int dtl_sa_configure(dtl_engine* e, const int* rt_info, int n_changes, const int* sa_tau, const int* sa_link, const double* sa_lanes)
{
    return sa_configure_impl(e, rt_info, std::max(0, n_changes), sa_tau, sa_link, sa_lanes);
}
int dtl_sa_begin(dtl_engine* e, double* sys_tt)
{
    const int rc = sa_begin_impl(e, sys_tt);
    if (rc) comm_abort(e);
    return rc;
}
int dtl_sa_iteration(dtl_engine* e, int it, double row[2])
{
    const int rc = sa_iteration_impl(e, it, row);
    if (rc) comm_abort(e);
    return rc;
}
int dtl_sa_set_dms(dtl_engine* e, int n, const int* dms_tau, const int* dms_link, const int* dms_zone)
{
    return sa_set_dms_impl(e, n, dms_tau, dms_link, dms_zone);
}
// per column, in dtl_get_columns order: impacted_path_flag as the route modification left it, and the information_type of its cell
int dtl_sa_get_route_flags(dtl_engine* e, unsigned char* path_impact, unsigned char* info_type)
{
    CK(cudaSetDevice(e->dev));
    PoolHost h;
    if (pool_to_host(e, h)) return -1;
    long long c_out = 0;
    std::vector<int> order;
    for (int c = 0; c < e->n_cells; ++c) {
        const size_t base = (size_t)c * e->pool.cap;
        order.resize(h.count[c]);
        std::iota(order.begin(), order.end(), 0);
        std::sort(order.begin(), order.end(), [&](int x, int y) { return h.node_sum[base + x] < h.node_sum[base + y]; });
        for (int j : order) {
            if (path_impact) path_impact[c_out] = e->sa_slot_impact.size() > base + j ? e->sa_slot_impact[base + j] : 0;
            if (info_type) info_type[c_out] = e->sa_cell_info.size() > (size_t)c ? e->sa_cell_info[c] : 0;
            ++c_out;
        }
    }
    return 0;
}
int dtl_sa_route_modification(dtl_engine* e, int* n_split)
{
    const int rc = sa_route_modification_impl(e, n_split);
    if (rc) comm_abort(e);
    return rc;
}
int dtl_sa_column_update(dtl_engine* e, int n, double row[4])
{
    const int rc = sa_column_update_impl(e, n, row);
    if (rc) comm_abort(e);
    return rc;
}
int dtl_run_assignment(dtl_engine* e, int K, int CU, double* iter_rows, double* cu_rows)
{
    if (dtl_iterations(e, 0, K, iter_rows)) return -1;
    for (int n = 0; n < CU; ++n) {
        double row[4];
        if (dtl_column_update(e, n, row)) return -1;
        if (cu_rows) memcpy(cu_rows + (size_t)n * 4, row, sizeof(row));
    }
    return dtl_finalize(e, K, nullptr);
}

// -------------------------------------------------------------------------------------------------
static int fetch(dtl_engine* e, double* dst, const double* src, int tau)
{
    if (!dst) return 0;
    CK(cudaMemcpyAsync(dst, src + (size_t)tau * e->L, sizeof(double) * e->L, cudaMemcpyDeviceToHost, e->st));
    return 0;
}

static int fetch_fixed(dtl_engine* e, double* dst, const unsigned long long* src)
{
    std::vector<long long> tmp(e->L);
    CK(cudaMemcpyAsync(tmp.data(), src, sizeof(long long) * e->L, cudaMemcpyDeviceToHost, e->st));
    if (sync_stream(e)) return -1;
    for (int l = 0; l < e->L; ++l) dst[l] = (double)tmp[l] * (1.0 / DTL_FIX_SCALE);
    return 0;
}

int dtl_get_link_state(dtl_engine* e, int tau, double* tt, double* pce_volume, double* person_volume, double* link_volume,
                       double* doc, double* voc, double* P, double* vt2)
{
    if (tau < 0 || tau >= e->T) { set_err("period %d out of range", tau); return -1; }
    CK(cudaSetDevice(e->dev));
    if (fetch(e, tt, e->ls.tt, tau) || fetch(e, link_volume, e->ls.link_volume, tau) || fetch(e, doc, e->ls.doc, tau) ||
        fetch(e, voc, e->ls.voc, tau) || fetch(e, P, e->ls.P, tau) || fetch(e, vt2, e->ls.vt2, tau))
        return -1;
    if (sync_stream(e)) return -1;
    if (pce_volume && fetch_fixed(e, pce_volume, e->ls.pce_fix + (size_t)tau * e->L)) return -1;
    if (person_volume && fetch_fixed(e, person_volume, e->ls.person_fix + (size_t)tau * (1 + e->AT) * e->L)) return -1;
    return 0;
}

int dtl_get_volume_fixed(dtl_engine* e, int tau, long long* q)
{
    if (tau < 0 || tau >= e->T) { set_err("period %d out of range", tau); return -1; }
    CK(cudaSetDevice(e->dev));
    CK(cudaMemcpyAsync(q, e->ls.pce_fix + (size_t)tau * e->L, sizeof(long long) * e->L, cudaMemcpyDeviceToHost, e->st));
    if (sync_stream(e)) return -1;
    return 0;
}

int dtl_get_person_volume_at(dtl_engine* e, int tau, int at, double* out)
{
    if (tau < 0 || tau >= e->T || at < 0 || at >= e->AT) { set_err("index out of range"); return -1; }
    CK(cudaSetDevice(e->dev));
    return fetch_fixed(e, out, e->ls.person_fix + ((size_t)tau * (1 + e->AT) + 1 + at) * e->L);
}

int dtl_get_link_detail(dtl_engine* e, int tau, double* avg_queue_speed, double* avg_speed_bpr, double* Q_mu, double* Q_gamma,
                        double* t0, double* t3, double* severe)
{
    if (tau < 0 || tau >= e->T) { set_err("period %d out of range", tau); return -1; }
    CK(cudaSetDevice(e->dev));
    if (fetch(e, avg_queue_speed, e->ls.avg_queue_speed, tau) || fetch(e, avg_speed_bpr, e->ls.avg_speed_bpr, tau) ||
        fetch(e, Q_mu, e->ls.Q_mu, tau) || fetch(e, Q_gamma, e->ls.Q_gamma, tau) || fetch(e, t0, e->ls.t0, tau) ||
        fetch(e, t3, e->ls.t3, tau) || fetch(e, severe, e->ls.severe_P, tau))
        return -1;
    if (sync_stream(e)) return -1;
    return 0;
}

int dtl_get_model_speed(dtl_engine* e, int tau, int first_slot, int n_slots, float* speed)
{
    if (tau < 0 || tau >= e->T || n_slots <= 0) { set_err("bad arguments"); return -1; }
    CK(cudaSetDevice(e->dev));
    float* d = nullptr;
    const size_t n = (size_t)e->L * n_slots;
    CK(cudaMalloc(&d, std::max<size_t>(n, 1) * sizeof(float)));
    cudaMemsetAsync(d, 0, n * sizeof(float), e->st);
    launch_model_speed(e->T, e->L, tau, e->vp, e->ls, first_slot, n_slots, d, e->st);
    cudaError_t ce = cudaMemcpyAsync(speed, d, n * sizeof(float), cudaMemcpyDeviceToHost, e->st);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->st);
    cudaFree(d);
    CK(ce);
    return 0;
}

int dtl_get_gen_cost(dtl_engine* e, int at, int tau, double* cost)
{
    if (tau < 0 || tau >= e->T || at < 0 || at >= e->AT) { set_err("index out of range"); return -1; }
    CK(cudaSetDevice(e->dev));
    const ForwardStar& f = e->fs[fs_index(e, at, tau)];
    std::vector<Edge> ed(std::max(1, f.n_edges));
    CK(cudaMemcpyAsync(ed.data(), f.edges, sizeof(Edge) * f.n_edges, cudaMemcpyDeviceToHost, e->st));
    if (sync_stream(e)) return -1;
    for (int l = 0; l < e->L; ++l) cost[l] = 0.0;
    for (int i = 0; i < f.n_edges; ++i) cost[ed[i].link] = ed[i].cost;
    return 0;
}

// ---- column pool read-back ---------------------------------------------------------------------
long long dtl_num_columns(dtl_engine* e)
{
    if (cudaSetDevice(e->dev) != cudaSuccess) return -1;
    std::vector<int> cnt(std::max(1, e->n_cells));
    if (cudaMemcpy(cnt.data(), e->pool.count, sizeof(int) * e->n_cells, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    long long n = 0;
    for (int i = 0; i < e->n_cells; ++i) n += cnt[i];
    return n;
}

long long dtl_num_column_links(dtl_engine* e)
{
    PoolHost h;
    if (cudaSetDevice(e->dev) != cudaSuccess || pool_to_host(e, h)) return -1;
    long long n = 0;
    for (int c = 0; c < e->n_cells; ++c)
        for (int j = 0; j < h.count[c]; ++j) n += h.n_links[(size_t)c * e->pool.cap + j];
    return n;
}

int dtl_get_columns(dtl_engine* e, int* od_index, int* node_sum, int* seq_no, int* n_links, int* n_nodes, double* volume,
                    double* cost, double* time, int* links)
{
    CK(cudaSetDevice(e->dev));
    PoolHost h;
    if (pool_to_host(e, h)) return -1;
    unsigned long long used = 0;
    CK(cudaMemcpy(&used, e->pool.arena_used, 8, cudaMemcpyDeviceToHost));
    used = std::min<unsigned long long>(used, (unsigned long long)e->pool.arena_cap);
    std::vector<int> arena;
    if (links) {
        arena.resize(std::max<size_t>(1, (size_t)used));
        CK(cudaMemcpy(arena.data(), e->pool.arena, sizeof(int) * (size_t)used, cudaMemcpyDeviceToHost));
    }
    long long c_out = 0, l_out = 0;
    std::vector<int> order;
    for (int c = 0; c < e->n_cells; ++c) {
        const size_t base = (size_t)c * e->pool.cap;
        order.resize(h.count[c]);
        std::iota(order.begin(), order.end(), 0);
        std::sort(order.begin(), order.end(), [&](int x, int y) { return h.node_sum[base + x] < h.node_sum[base + y]; });
        for (int j : order) { // std::map order of the reference: ascending node_sum
            if (od_index) od_index[c_out] = e->cell_global[c];
            if (node_sum) node_sum[c_out] = h.node_sum[base + j];
            if (seq_no) seq_no[c_out] = h.seq[base + j]; // slots are filled in insertion order == path_seq_no (but see ColumnPool::seq)
            if (n_links) n_links[c_out] = h.n_links[base + j];
            if (n_nodes) n_nodes[c_out] = h.n_nodes[base + j];
            if (volume) volume[c_out] = h.volume[base + j];
            if (cost) cost[c_out] = h.cost[base + j];
            if (time) time[c_out] = h.time[base + j];
            if (links) {
                memcpy(links + l_out, arena.data() + h.offset[base + j], sizeof(int) * h.n_links[base + j]);
                l_out += h.n_links[base + j];
            }
            ++c_out;
        }
    }
    return 0;
}

// per-column records of the SA stage in dtl_get_columns order: od_impact = the cell's OD_impact_flag / at_od_impacted flag
// (assignment.h:699-713), hist_time / hist_volume [n_iterations][n_columns] = path travel time / volume after each
// dtl_sa_column_update (n_iterations = calls so far; may be NULL)
int dtl_sa_get_columns(dtl_engine* e, unsigned char* od_impact, int* n_iterations, double* hist_time, double* hist_volume)
{
    CK(cudaSetDevice(e->dev));
    PoolHost h;
    if (pool_to_host(e, h)) return -1;
    std::vector<unsigned char> imp(std::max(1, e->n_cells), 0);
    if (e->d_od_impact) CK(cudaMemcpy(imp.data(), e->d_od_impact, (size_t)e->n_cells, cudaMemcpyDeviceToHost));
    const size_t n_it = e->sa_hist_time.size();
    if (n_iterations) *n_iterations = (int)n_it;
    long long n_cols = 0;
    for (int c = 0; c < e->n_cells; ++c) n_cols += h.count[c];
    long long c_out = 0;
    std::vector<int> order;
    for (int c = 0; c < e->n_cells; ++c) {
        const size_t base = (size_t)c * e->pool.cap;
        order.resize(h.count[c]);
        std::iota(order.begin(), order.end(), 0);
        std::sort(order.begin(), order.end(), [&](int x, int y) { return h.node_sum[base + x] < h.node_sum[base + y]; });
        for (int j : order) {
            if (od_impact) od_impact[c_out] = imp[c];
            for (size_t q = 0; q < n_it; ++q) {
                if (hist_time) hist_time[q * n_cols + c_out] = e->sa_hist_time[q][base + j];
                if (hist_volume) hist_volume[q * n_cols + c_out] = e->sa_hist_volume[q][base + j];
            }
            ++c_out;
        }
    }
    return 0;
}

// ---- kernel-level entry points -----------------------------------------------------------------
// n one-to-all trees on star `f` from the nodes `origin` (zones: the zone whose tagged out-links the tree may use, per tree);
// h_tail[l] = the node a link leaves in the search direction (the predecessor node of a node reached over l)
static int trees_on_star(dtl_engine* e, ForwardStar& f, const double* cost, const std::vector<int>& origin, const int* zones, int n,
                         const std::vector<int>& h_tail, double* labels, int* pred_node, int* pred_link, int* tie_nodes, double stats[6])
{
    const int N = e->N;
    if (cost) { // known-answer mode: caller-supplied generalized costs
        std::vector<Edge> ed(std::max(1, f.n_edges));
        CK(cudaMemcpy(ed.data(), f.edges, sizeof(Edge) * f.n_edges, cudaMemcpyDeviceToHost));
        double sum = 0;
        long long cnt = 0;
        for (int i = 0; i < f.n_edges; ++i) {
            ed[i].cost = cost[ed[i].link];
            if (ed[i].cost > 1e-3) { sum += ed[i].cost; ++cnt; }
        }
        CK(cudaMemcpy(f.edges, ed.data(), sizeof(Edge) * f.n_edges, cudaMemcpyHostToDevice));
        const double factor = e->opt.sssp_delta_factor > 0 ? e->opt.sssp_delta_factor : 8.0;
        f.delta = factor * (cnt ? sum / (double)cnt : 1.0);
    }
    int *d_origin = nullptr, *d_zone = nullptr;
    CK(cudaMalloc(&d_origin, sizeof(int) * std::max(1, n)));
    CK(cudaMalloc(&d_zone, sizeof(int) * std::max(1, n)));
    int rc = 0;
    SsspStats s0{}, s1{};
    cudaMemcpy(&s0, e->d_stats, sizeof(s0), cudaMemcpyDeviceToHost);
    float total_ms = 0;
    std::vector<unsigned long long> hl;
    std::vector<int> hp;
    for (int b0 = 0; b0 < n && rc == 0; b0 += e->batch_tasks) {
        const int nb = std::min(e->batch_tasks, n - b0);
        cudaMemcpy(d_origin, origin.data() + b0, sizeof(int) * nb, cudaMemcpyHostToDevice);
        cudaMemcpy(d_zone, zones + b0, sizeof(int) * nb, cudaMemcpyHostToDevice);
        cudaEvent_t ea = e->get_event(), eb = e->get_event();
        cudaEventRecord(ea, e->st);
        rc = run_sssp_batch(e, f, nb, d_origin, d_zone, true, 0, origin.data() + b0, -1);
        cudaEventRecord(eb, e->st);
        if (rc) break;
        if (cudaStreamSynchronize(e->st) != cudaSuccess) { set_err("SSSP batch failed: %s", cudaGetErrorString(cudaGetLastError())); rc = -1; break; }
        float t = 0;
        cudaEventElapsedTime(&t, ea, eb);
        total_ms += t;
        e->free_events.push_back(ea); e->free_events.push_back(eb);
        if (tie_nodes) memcpy(tie_nodes + b0, e->h_tie_nodes.data(), sizeof(int) * nb);
        if (labels) cudaMemcpy(labels + (size_t)b0 * N, e->d_labels, sizeof(double) * (size_t)nb * N, cudaMemcpyDeviceToHost);
        if (pred_link || pred_node) {
            hp.resize((size_t)nb * N);
            cudaMemcpy(hp.data(), e->d_pred, sizeof(int) * (size_t)nb * N, cudaMemcpyDeviceToHost);
            for (size_t i = 0; i < (size_t)nb * N; ++i) {
                const int pl = hp[i];
                if (pred_link) pred_link[(size_t)b0 * N + i] = pl;
                if (pred_node) pred_node[(size_t)b0 * N + i] = pl >= 0 ? h_tail[pl] : DTL_NO_PRED;
            }
        }
    }
    cudaFree(d_origin);
    cudaFree(d_zone);
    if (rc) return rc;
    if (check_flags(e)) return -1;
    e->resolve_timing();
    cudaMemcpy(&s1, e->d_stats, sizeof(s1), cudaMemcpyDeviceToHost);
    if (stats) {
        stats[0] = total_ms;
        stats[1] = (double)(s1.counted_edges - s0.counted_edges);
        stats[2] = (double)(s1.relaxations - s0.relaxations);
        stats[3] = (double)(s1.pops - s0.pops);
        stats[4] = (double)(s1.rounds - s0.rounds);
        stats[5] = (double)(s1.replayed - s0.replayed);
        if (getenv("DTL_SSSP_PROFILE")) { // filled by a -DSSSP_PROFILE build of the shared-memory-queue kernel
            double d[8];
            for (int k = 0; k < 8; ++k) d[k] = (double)(s1.prof[k] - s0.prof[k]);
            fprintf(stderr, "sssp: spilled %.3g | thread-0 cycles: init %.3g expansion %.3g barrier %.3g re-bucket %.3g | passes %.3g entries %.3g "
                            "rounds %.3g | bundle kernel, inside a pass: until the label row arrived %.3g, until the atomics returned %.3g\n",
                    d[0], d[1], d[2], d[3], d[4], d[6], d[7], (double)(s1.rounds - s0.rounds), (double)(s1.stage[0] - s0.stage[0]), d[5]);
            if (s1.stage[3] > s0.stage[3])
                fprintf(stderr, "sssp: asynchronous bundle search, cycles per bundle: max (since engine creation) %.4g, mean %.4g over %.0f bundles\n",
                        (double)s1.stage[1], (double)(s1.stage[2] - s0.stage[2]) / (double)(s1.stage[3] - s0.stage[3]), (double)(s1.stage[3] - s0.stage[3]));
        }
    }
    e->trees_valid = false;
    return 0;
}

int dtl_sssp_trees(dtl_engine* e, int at, int tau, const double* cost, const int* zones, int n, double* labels, int* pred_node,
                   int* pred_link, int* tie_nodes, double stats[6])
{
    if (tau < 0 || tau >= e->T || at < 0 || at >= e->AT) { set_err("index out of range"); return -1; }
    CK(cudaSetDevice(e->dev));
    std::vector<int> origin(std::max(0, n));
    for (int i = 0; i < n; ++i) {
        if (zones[i] < 0 || zones[i] >= e->Z) { set_err("zone %d out of range", zones[i]); return -1; }
        origin[i] = e->h_zone_centroid[zones[i]];
    }
    return trees_on_star(e, e->fs[fs_index(e, at, tau)], cost, origin, zones, n, e->h_link_from, labels, pred_node, pred_link, tie_nodes, stats);
}

// Backward one-to-all search from destination nodes: NetworkForSP::optimal_backward_label_correcting_from_destination
// (routing.h:1009-1273).  The star of the search is every node's in-link list in insertion order
// (g_node_vector[].m_incoming_link_seq_no_vector: no agent-type filter), without the links that carry an outgoing-connector
// zone tag (routing.h:1104-1108) and without the links rt_allowed rules out (RT_allowed_use, routing.h:1119); the cost of a
// link is avg_travel_time + RT_waiting_time (routing.h:1148).  The same K1 kernels run on it -- a label is now the least cost
// from the node TO the destination, pred_node the next node downstream, pred_link the link to take.
int dtl_backward_trees(dtl_engine* e, int tau, const double* cost, const unsigned char* rt_allowed, const int* dest_nodes, int n,
                       double* labels, int* pred_node, int* pred_link, int* tie_nodes, double stats[6])
{
    if (tau < 0 || tau >= e->T) { set_err("period %d out of range", tau); return -1; }
    CK(cudaSetDevice(e->dev));
    const int L = e->L;
    for (int i = 0; i < n; ++i)
        if (dest_nodes[i] < 0 || dest_nodes[i] >= e->N) { set_err("destination node %d out of range", dest_nodes[i]); return -1; }
    if (e->bstar.empty()) e->bstar.resize(e->T);
    ForwardStar* f = &e->bstar[tau];
    ForwardStar masked; // a caller-supplied RT_allowed_use mask changes the star: built for this call only
    std::vector<double> h_tt(L);
    CK(cudaMemcpy(h_tt.data(), e->ls.tt + (size_t)tau * L, sizeof(double) * L, cudaMemcpyDeviceToHost));
    if (rt_allowed || !f->row_ptr) {
        std::vector<unsigned char> keep(L);
        for (int l = 0; l < L; ++l) keep[l] = e->h_link_tag[l] < 0 && (!rt_allowed || rt_allowed[l]);
        if (rt_allowed) f = &masked;
        FsHost h;
        build_star_host(e, *f, e->h_link_to.data(), e->h_link_from.data(), nullptr, keep.data(), h_tt.data(), h);
        if (upload_star(e, *f, h)) return -1; // (device memory of a masked star stays with the engine until dtl_destroy)
        if (!e->d_link_to && e->upload(&e->d_link_to, e->h_link_to.data(), (size_t)L)) return -1;
        f->d_tail = e->d_link_to;
    }
    std::vector<int> origin(dest_nodes, dest_nodes + std::max(0, n)), zones(std::max(1, n), 0);
    return trees_on_star(e, *f, cost ? cost : h_tt.data(), origin, zones.data(), n, e->h_link_to, labels, pred_node, pred_link, tie_nodes, stats);
}

int dtl_vdf(dtl_engine* e, int tau, const double* volume, double* tt, double* doc, double* voc, double* P, double* vt2)
{
    if (tau < 0 || tau >= e->T) { set_err("period %d out of range", tau); return -1; }
    CK(cudaSetDevice(e->dev));
    // volumes of the other periods: keep the current link volumes
    CK(cudaMemcpyAsync(e->d_volume_tmp, e->ls.link_volume, sizeof(double) * (size_t)e->T * e->L, cudaMemcpyDeviceToDevice, e->st));
    CK(cudaMemcpyAsync(e->d_volume_tmp + (size_t)tau * e->L, volume, sizeof(double) * e->L, cudaMemcpyHostToDevice, e->st));
    if (do_vdf(e, true, false, e->d_volume_tmp)) return -1;
    if (sync_stream(e)) return -1;
    e->resolve_timing();
    return dtl_get_link_state(e, tau, tt, nullptr, nullptr, nullptr, doc, voc, P, vt2);
}

int dtl_scatter(dtl_engine* e, int decay_k)
{
    CK(cudaSetDevice(e->dev));
    if (do_scatter(e, decay_k, true)) return -1;
    if (sync_stream(e)) return -1;
    e->resolve_timing();
    return 0;
}

int dtl_get_tree(dtl_engine* e, int task, double* label, int* pred_node, int* pred_link)
{
    if (!e->trees_valid) { set_err("trees are not kept: set dtl_options.keep_trees and run dtl_iteration first"); return -1; }
    if (task < 0 || task >= (int)e->task_at.size()) { set_err("task %d out of range", task); return -1; }
    CK(cudaSetDevice(e->dev));
    const int N = e->N;
    if (label) CK(cudaMemcpy(label, e->d_labels + (size_t)task * N, sizeof(double) * N, cudaMemcpyDeviceToHost));
    if (pred_node || pred_link) {
        std::vector<int> hp(N);
        CK(cudaMemcpy(hp.data(), e->d_pred + (size_t)task * N, sizeof(int) * N, cudaMemcpyDeviceToHost));
        for (int i = 0; i < N; ++i) {
            if (pred_link) pred_link[i] = hp[i];
            if (pred_node) pred_node[i] = hp[i] >= 0 ? e->h_link_from[hp[i]] : DTL_NO_PRED;
        }
    }
    return 0;
}

int dtl_set_bounded_search(dtl_engine* e, int on)
{
    const int was = e->bounded ? 1 : 0;
    e->bounded = on != 0;
    return was;
}

int dtl_set_overlap(dtl_engine* e, int on)
{
    const int was = e->overlap ? 1 : 0;
    e->overlap = on != 0 && e->st2 != nullptr;
    return was;
}

int dtl_get_timing(dtl_engine* e, double ms[8], long long counts[8], int reset)
{
    CK(cudaSetDevice(e->dev));
    if (sync_stream(e)) return -1;
    e->resolve_timing();
    if (ms) memcpy(ms, e->ms, sizeof(e->ms));
    if (counts) memcpy(counts, e->launches, sizeof(e->launches));
    if (reset) { memset(e->ms, 0, sizeof(e->ms)); memset(e->launches, 0, sizeof(e->launches)); }
    return 0;
}

int dtl_get_counters(dtl_engine* e, double c[10], int reset)
{
    CK(cudaSetDevice(e->dev));
    if (sync_stream(e)) return -1;
    SsspStats s{};
    unsigned long long ctr[4] = { 0 };
    CK(cudaMemcpy(&s, e->d_stats, sizeof(s), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(ctr, e->d_counters, sizeof(ctr), cudaMemcpyDeviceToHost));
    const double now[10] = { e->counted_edges_cached < 0 ? 0.0 : e->counted_edges_cached, (double)s.relaxations, (double)s.pops,
                             (double)s.rounds, e->tasks_done, (double)s.replayed, (double)ctr[0], (double)ctr[1], e->reruns,
                             (double)s.tie_candidates };
    if (c) {
        for (int i = 0; i < 10; ++i) c[i] = now[i] - e->ctr_base[i];
        c[0] = now[0]; // counted edges per iteration (static property of the trees), not cumulative
    }
    if (reset) memcpy(e->ctr_base, now, sizeof(now));
    return 0;
}

} // extern "C"
