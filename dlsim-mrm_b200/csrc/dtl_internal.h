// Internal device data layout and kernel launchers of the B200 DTALite engine.
// Everything here is structure-of-arrays in HBM; see DESIGN.md "Data layout".
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/dtalite_b200.h"

#define DTL_MAX_LABEL_COST 1.0e15  // unreachable label, reference DTA.h:11
#define DTL_NO_PRED (-9999)        // reference routing.h:683-685
#define DTL_FIX_SCALE 4294967296.0 // Q31.32 fixed point for link volumes
#define DTL_SPEED_SLOTS 300        // MAX_TIMEINTERVAL_PerDay, reference DTA.h:24

namespace dtl {

// One forward-star edge: everything a relaxation needs in a single 16-byte load.
// `to` carries the node seq no; bit 31 is set when the link has an outgoing-connector zone tag
// (reference routing.h:779-786), in which case link_tag[link] must equal the origin zone.
struct __align__(16) Edge {
    double cost;
    int to;
    int link;
};

// SSSP work-queue entry: node, the label it was pushed with, and the link that produced it.
struct __align__(16) QEntry {
    double lab;
    int node;
    int link;
};

// forward star of one (agent type, period): reference NetworkForSP::BuildNetwork, routing.h:257-352
struct ForwardStar {
    int n_edges = 0;
    int* row_ptr = nullptr;   // [N+1]
    Edge* edges = nullptr;    // [E] forward-star order == reference scan order
    int2* rows = nullptr;     // [N] {row_ptr[i], row_ptr[i+1]}: one 8-byte load per node
    int4* link_info = nullptr; // [L] {from node, row begin of the to-node, row end of the to-node, 0}: what the expansion of an
                              // entry reached over link l needs, in one 16-byte load
    int* in_ptr = nullptr;    // [N+1] reverse star (tie verification only)
    int2* in_edges = nullptr; // [E] {from node, forward edge index}
    int2* rin4 = nullptr;     // [N][4] the first four in-edges of every node {from node, forward edge index}, {-1, 0} padding;
                              // bit 30 of [v][3].y: the node has more than four (the rest is read through in_ptr / in_edges)
    int* link_edge = nullptr; // [L] forward edge index of a link, -1 when the agent type may not use it
    const int* d_tail = nullptr; // [L] the node a link leaves in the search direction; null: link_from (forward search)
    double delta = 1.0;       // current label window of the near/far SSSP
    bool open_tags = false;   // a node that is not a hidden centroid row has a zone-tagged out-link
    std::vector<int> h_row_ptr;
    std::vector<int> h_to;      // [E] head node of every forward-star edge (host copy for the origin-bundle planner)
    mutable std::vector<float> h_embed; // [2N] landmark-distance coordinates of every node (filled on first use)
};

struct VdfParams { // [T*L] each, reference CPeriod_VDF (VDF.h:369-466)
    double *fftt, *cap, *alpha, *beta, *vf, *vc, *nlanes, *cap_lane, *qdf, *preload, *penalty;
    double *Q_alpha, *Q_beta, *Q_cd, *Q_n, *Q_cp, *Q_s, *t2, *L_h, *start_h, *end_h, *plf;
    float *cycle_length, *red_time, *sat_flow;
    int* vdf_type;
};

struct LinkState { // [T*L] each
    unsigned long long* pce_fix;    // Q31.32 PCE volume sums (all-reduced across GPUs)
    unsigned long long* person_fix; // [T*(1+AT)*L]: slot 0 person volume, slot 1+at per agent type
    double *tt, *link_volume, *doc, *voc, *P, *vt2;
    // detail outputs, filled by the DETAIL variant of the VDF kernel
    double *avg_queue_speed, *avg_speed_bpr, *Q_mu, *Q_gamma, *t0, *t3, *severe_P;
};

struct ColumnPool {     // column CSR with fixed slots per OD cell
    int n_od = 0;       // OD cells owned by this rank
    int cap = 0;        // column slots per OD cell
    int* count;         // [n_od]
    int* node_sum;      // [n_od*cap]
    int* n_links;       // [n_od*cap]
    int* n_nodes;       // [n_od*cap]
    long long* offset;  // [n_od*cap] first path link in the arena
    double* volume;     // [n_od*cap]
    double* cost;       // [n_od*cap] path_gradient_cost
    double* time;       // [n_od*cap] path_travel_time
    int* seq;           // [n_od*cap] or null: path_seq_no of a column where it is not its slot index (only g_add_new_column of the
                        // route modification makes them differ, assignment.h:1042)
    int* arena;         // path links, origin -> destination order
    long long arena_cap = 0;
    unsigned long long* arena_used; // device bump pointer
};

struct SsspStats { // accumulated on device
    unsigned long long relaxations, pops, rounds, tie_candidates, counted_edges, overflow, tied_origins, replayed;
    unsigned long long prof[8]; // unused
    unsigned long long stage[8]; // unused
};

// ------------------------------------------------------------------ kernel launchers
struct VdfLaunch {
    int T, L, AT;
    VdfParams p;
    LinkState s;
    const double *toll, *lr_price; // [T*AT*L]
    const float* vot;              // [AT]
    int** link_edge;               // device array [AT*T] of link->edge maps
    Edge** edges;                  // device array [AT*T]
    double* block_sums;            // [n_blocks] partial sums of tt*link_volume
    double* sys_tt;                // device scalar
    const double* volume_override; // non-null: use this volume instead of pce+preload (KAT entry)
    bool detail;
    bool write_cost;
    int* error_flag;               // bit 2: a generalized link cost is negative (not supported by the label-correcting kernels)
};
void launch_vdf(const VdfLaunch& a, cudaStream_t st);
void launch_model_speed(int T, int L, int tau, const VdfParams& p, const LinkState& s, int first_slot, int n_slots,
                        float* out, cudaStream_t st);
int vdf_blocks(int n);

struct SsspLaunch {
    int N;
    const int* row_ptr;
    const int2* rows;        // [N] {row begin, row end}
    const int4* link_info;   // [L] {from node, row begin of the to-node, row end of the to-node, 0}
    const Edge* edges;
    const int* in_ptr;
    const int2* in_edges;
    const int2* rin4;        // [N][4] padded reverse star (ForwardStar::rin4)
    const int* link_tag;
    const int* link_from;
    int n_tasks;             // tasks in this batch
    int n_run;               // trees to compute in this launch (== n_tasks unless task_list is set)
    const int* task_list;    // optional: batch-local task indices to (re)compute
    const int* task_origin;  // [n_tasks] origin centroid node
    const int* task_zone;    // [n_tasks] origin zone seq no
    unsigned long long* labels; // [n_tasks][N] FP64 bit patterns
    int* pred_link;          // [n_tasks][N]
    int* tie_nodes;          // [n_tasks] out: verified tie nodes (>= 2 tight upstream nodes); -1 = work queue overflowed
    double delta;
    // per-CTA workspace: two near queues of cap_near entries and two far piles of cap_far entries
    QEntry* ws;              // [grid][ws_stride]
    size_t ws_stride;
    int cap_near, cap_far;
    int cap_smem;            // shared-memory-queue variant: entries per near queue (two buffers, 28 B per entry)
    int smq;                 // 1: launch the shared-memory-queue variant (few trees per SM)
    int far_trigger;         // far-pile size above which dead entries are dropped between rounds
    // destination-bounded search of the shared-memory-queue kernel (the loop's default; null pointers = complete search; see BundleLaunch::dest_ptr): the destinations with demand
    // of batch-local task i are od_dest[cells[k]] for k in [cell_ptr[i], cell_ptr[i + 1])
    const int* bd_cell_ptr;  // null: every search runs to the end
    const int* bd_cells;
    const int* bd_dest;
    int* tie_cand;           // [grid][tie_cap]
    int tie_cap;
    int* task_counter;       // device int, zeroed before launch
    SsspStats* stats;
    int grid, block;
};
void launch_sssp(const SsspLaunch& a, cudaStream_t st);     // global-memory queues: many trees per SM
void launch_sssp_smq(const SsspLaunch& a, cudaStream_t st); // shared-memory queues: few trees per SM
// exact serial emulation of the reference deque for origins whose tree has tied predecessors
void launch_sssp_replay(const SsspLaunch& a, const int* replay_tasks, int n_replay, int* status_ws, int* next_ws,
                        cudaStream_t st);
void launch_count_edges(const SsspLaunch& a, cudaStream_t st);

// origin-bundle variant (kernels_sssp_bundle.cu): B neighbouring origins share one search over node-major label rows
struct BundleLaunch {
    SsspLaunch s;            // network, tasks, output trees, delta, stats (queues of `s` are not used)
    int B;                   // origins per bundle: 4, 8 or 16
    int lpt;                 // origins per thread: 2 or 4
    int async;               // 1: search without round barriers (kernels_sssp_bundle_async.cu; cap_near = ring size, a power of two)
    int tagged;              // 1: zone-tagged out-links exist outside the hidden centroid rows (the search then checks tags per edge)
    int near_limit;          // async: entries queued or in flight above which the bundle gives up (<= ring size; test hook)
    int n_bundles;
    const int* bundle_tasks; // [n_bundles][B] batch-local task index, -1 = empty slot
    unsigned long long* bl;  // [n_bundles][N][B] labels, node-major
    int2* far;               // [grid][2][cap_far] far piles {node, float key}
    int cap_far;
    int cap_near;            // shared-memory near queue: entries per buffer
    int* bundle_counter;     // device int, zeroed before launch
    int grid, block;
    int ctas_per_sm;         // async: resident CTAs per SM the kernel is compiled for (1, or 2 / 3 / 4 with the smaller CTA sizes)
    int slim;                // 1: the label -> tree pass leaves predecessors node-major in bp and writes no per-tree arrays
    int fuse_pred;           // slim + async: the search kernel runs the label -> predecessor pass itself -- every CTA that has no
                             // (more) bundle to search takes tiles of the bundles that are finished (work[] below)
    int* work;               // fuse_pred: [2 + 2 * n_bundles] {finished count, finished bundles in order (-1 = not yet), tile counters,
                             // resident CTAs of the search kernel}
    // destination-bounded search (the loop's default, dtl_set_bounded_search; dest_ptr null = complete search): a bundle's search stops as soon as no queued node can still
    // improve the label of a destination the bundle's origins have demand to; labels beyond that bound are not final and the
    // label -> predecessor pass treats them as unreached
    const int* dest_ptr;     // [n_bundles + 1] into dest_idx, or null: every search runs to the end
    const int* dest_idx;     // node * B + slot of every (origin, destination with demand) of a bundle
    unsigned* bundle_bound;  // [n_bundles] out: scheduling key (high word of the label) below which the labels of the bundle are final
    unsigned* started_word;  // fuse_pred, optional: the CTA that completes the resident count writes started_tag here (pipelined loop:
    unsigned started_tag;    // the back-trace of the previous iteration waits for it on the side stream)
    int2* bp;                // [n_bundles][N][B] {predecessor link, predecessor node}, node-major (slim)
    int* batch_flag;         // slim: bit 0 = an order-dependent tie was found, bit 1 = a bundle overflowed
};
void launch_count_edges_nm(const BundleLaunch& a, cudaStream_t st);     // Graph500 edge count on the node-major labels of a bundle batch
void launch_sssp_bundle(const BundleLaunch& a, cudaStream_t st);        // search + label -> tree pass
void launch_sssp_bundle_finish(const BundleLaunch& a, cudaStream_t st); // the label -> tree pass alone
void launch_sssp_bundle_async_search(const BundleLaunch& a, cudaStream_t st);
void launch_bundle_pred_helpers(const BundleLaunch& a, int grid, cudaStream_t st); // extra CTAs for the fused label -> predecessor pass
size_t bundle_async_smem_bytes(int N, int ring);
size_t bundle_smem_bytes(int N, int cap_near); // dynamic shared memory of one CTA: two queue bits per node + two near queues
void sssp_occupancy(int block, int* max_ctas_per_sm, int* regs);

struct ColumnLaunch {
    int N, L, T, AT;
    ColumnPool pool;
    // OD cells of this rank (local index space)
    const int *od_dest_node, *od_at, *od_tau, *od_task; // [n_od]
    const double* od_vol;
    const int* link_from;
    // trees of the current batch
    const unsigned long long* labels;
    const int* pred_link;
    int task_first, task_count;        // batch range in local task numbering
    const int* batch_cells;            // OD cells (local) belonging to the batch tasks
    int n_batch_cells;
    const unsigned char* task_skip_backtrace; // [n_tasks] origin has no physical out-link (routing.h:428)
    // node-major trees of an origin-bundle batch (slim label -> tree pass); used instead of labels / pred_link when bp != nullptr
    const unsigned long long* bl; // [n_bundles][N][B]
    const int2* bp;               // [n_bundles][N][B] {predecessor link, predecessor node}
    const int* task_slot;         // [task_count] bundle * B + slot of a batch-local task
    int B;
    const int* batch_flag;        // non-zero: the batch is redone through the per-tree path, this launch does nothing
    int k;
    double* least_contrib;             // [n_od]
    int* error_flag;                   // bit0 arena overflow, bit1 column slots exhausted
    unsigned long long* path_nodes;    // counter
    int* path_len;                     // [n_od] nodes of the path found for the cell (may be null): key of sort_cells_by_path_length
};
void launch_backtrace(const ColumnLaunch& a, cudaStream_t st);
// induced delay, main_api.cpp:877-891: links without a scenario whose degree of congestion exceeds 3 get network_design_flag -2
void launch_induced_delay_flags(signed char* design_flag, const double* doc, size_t n, cudaStream_t st);
void launch_set_go(const int* flags, int n, unsigned* go, unsigned* start, unsigned tag, int start_when_clean, cudaStream_t st);
void sort_cells_by_path_length(const int* path_len, int* cells, const int* cell_ptr, int first, int n_tasks, cudaStream_t st);

struct ScatterLaunch {
    int L, T, AT;
    ColumnPool pool;
    const int *od_at, *od_tau;
    const double *pce, *occ; // [T*AT*L]
    const double *pce_uni, *occ_uni; // [T*AT] the value when every link of (tau, at) carries the same one, NaN otherwise
    unsigned long long* pce_fix;
    unsigned long long* person_fix; // may be null (skipped inside the CG loop)
    int decay_k;                    // <0: none
    const unsigned char* decay_at;  // [AT] or null: the decay applies only to columns of agent types with a non-zero entry (SA
                                    // stage: only users with real-time information are re-routed, assignment.h:150-176)
    const unsigned char* decay_cell; // [n_od] or null: the same per OD cell (message-sign users are only re-routed from an
                                    // information zone, assignment.h:160-170); takes precedence over decay_at
    unsigned long long* link_refs;  // counter
    int cell_stride;                // set by launch_scatter: warp w handles cell (w * cell_stride) % n_od
};
void launch_scatter(const ScatterLaunch& a, cudaStream_t st);

struct PoolUpdateLaunch {
    int L, T, AT;
    ColumnPool pool;
    const int *od_at, *od_tau;
    const double* od_vol;
    const double* tt;      // [T*L]
    const double* penalty; // [T*L]
    const double* toll;    // [T*AT*L]
    const float* vot;
    int n;                 // column-pool iteration index
    double* od_sums;       // [n_od][4]: gap, cost, time, demand contributions
    double2* link_cost;    // [T*AT*L] workspace: {travel time, generalized first-order cost} per (tau, at, link)
    // sensitivity-analysis stage (assignment.h:699-713, 828-846): flow only shifts in OD cells of agent types with real-time
    // information that have a column over a link with a supply-side lane change
    int sa;                          // 1: SA rules
    const signed char* design_flag;  // [T*L] -1 = lanes changed by the scenario
    const int* rt_info;              // [AT]
    unsigned char* od_impact;        // [n_od] sticky flag of the cell
};
void launch_pool_update(const PoolUpdateLaunch& a, cudaStream_t st);

// ODME (kernels_odme.cu): reference ODME.h:91-516, assignment.h:263-563
struct OdmeLaunch {
    int L, T, AT;
    ColumnPool pool;
    const int *od_at, *od_tau;
    const double* tt;               // [T*L] current link travel times
    const double* link_volume;      // [T*L] PCE volume + preload as the last VDF pass stored it
    const float *at_pce, *at_occ;   // [AT] agent-type PCE / person occupancy (float)
    unsigned long long* pce_fix;    // [T*L]
    unsigned long long* person_fix; // [T*(1+AT)*L]
    double* od_sums;                // [n_od][4] {demand, travel time, UE gap, -}
    const double* obs;              // [T*L] observed count, 0 = none
    const int* ub;                  // [T*L] upper_bound_flag
    double* est_dev;                // [T*L] est_count_dev
    int n_sensors;
    const int* sensor_idx;          // [n_sensors] tau * L + link, in (link, tau) order
    double* sensor_dev;             // [n_sensors] est_count_dev of the sensors, for the host-side summary row
};
void launch_odme_stats(const OdmeLaunch& a, cudaStream_t st);
void launch_odme_scatter(const OdmeLaunch& a, cudaStream_t st);
void launch_odme_dev(const OdmeLaunch& a, cudaStream_t st);
void launch_odme_update(const OdmeLaunch& a, cudaStream_t st);

// deterministic sum of n doubles with stride `stride` starting at `in` -> out[0]
void launch_sum(const double* in, long long n, int stride, double* out, cudaStream_t st);

} // namespace dtl
