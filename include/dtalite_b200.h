/* dtalite_b200.h -- C ABI of the B200-native DTALite traffic-assignment hot path.
 *
 * Two levels, both plain C (pointers + sizes, no C++/torch types):
 *
 *  (A) The reference's own entry point, unchanged:
 *        double network_assignment(int assignment_mode, int column_generation_iterations,
 *                                  int column_updating_iterations, int ODME_iterations,
 *                                  int sensitivity_analysis_iterations, int simulation_iterations,
 *                                  int number_of_memory_blocks);
 *      replaces reference src/cpp/config.h:19-28 / main_api.cpp:499.  Reads settings.csv,
 *      node.csv, link.csv and the demand file(s) from the current directory and writes
 *      link_performance.csv, route_assignment.csv and final_summary.csv there.
 *
 *  (B) An engine API for hosts that keep their own GMNS reader: the network image is passed
 *      as arrays (dtl_problem) and every stage of the hot loop is one call.  Each entry
 *      point names the reference function it replaces.
 *
 * All calls are synchronous with respect to the host unless stated.  Return value 0 = ok,
 * non-zero = error (dtl_last_error() has the message).  There is no CPU fallback: when no
 * CUDA device is usable every compute call fails.
 */
#ifndef DTALITE_B200_H
#define DTALITE_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------- (A) drop-in entry point ---------------------------------- */
double network_assignment(int assignment_mode, int column_generation_iterations, int column_updating_iterations,
                          int ODME_iterations, int sensitivity_analysis_iterations, int simulation_iterations,
                          int number_of_memory_blocks);
/* main() of the reference executable (flash_dta.cpp:140-246): reads settings.csv [assignment]
 * and calls network_assignment.  Returns the process exit code. */
int dtalite_main(void);

/* ---------------------------------- (B) engine API ---------------------------------- */

/* Network + demand image.  Same content as the reference's g_node_vector / g_link_vector /
 * g_zone_vector / column-pool headers after g_read_input_data and connector creation
 * (main_api.cpp:525-570), as structure-of-arrays.  Per-period arrays are [T][L], per-period-
 * per-agent-type arrays [T][AT][L], row-major.  All pointers are HOST pointers; they are copied. */
typedef struct dtl_problem {
    int n_nodes, n_links, n_zones, n_agent_types, n_periods;
    int memory_blocks;            /* number_of_memory_blocks after clamping (main_api.cpp:519) */
    const int* link_from;         /* [L] from_node_seq_no */
    const int* link_to;           /* [L] to_node_seq_no */
    const int* link_tag;          /* [L] zone_seq_no_for_outgoing_connector, -1 if none */
    const int* node_zone;         /* [N] zone seq no for centroid nodes, else -1 */
    const int* zone_centroid;     /* [Z] centroid node seq no */
    const double *fftt, *cap, *alpha, *beta, *vf, *vc, *nlanes, *cap_lane, *qdf, *preload, *penalty;
    const double *Q_alpha, *Q_beta, *Q_cd, *Q_n, *Q_cp, *Q_s, *t2, *L_h, *start_h, *end_h, *plf;
    const float *cycle_length, *red_time, *sat_flow;
    const int* vdf_type;          /* 0 = q_vdf, 1 = bpr_vdf (DTA.h:53) */
    const double *pce, *occ, *toll, *lr_price;
    const unsigned char* allowed; /* CLink::AllowAgentType (DTA.h:1352) evaluated per (tau, at, link) */
    const float* vot;             /* [AT] */
    int n_od;                     /* OD cells with od_volume > 0 sorted by (o, d, at, tau) */
    const int *od_o, *od_d, *od_at, *od_tau;
    const double* od_vol;
    const float* origin_demand;   /* [Z] */
    float total_demand_volume;
} dtl_problem;

typedef struct dtl_options {
    int device;                   /* CUDA device ordinal; -1 = current device */
    int max_columns_per_od;       /* capacity of each OD cell's column set (>= number of CG iterations) */
    int rank, world_size;         /* origin zones are sharded over world_size ranks in spatially contiguous blocks (any partition gives
                                   * the same integer link-volume sums; DTL_SHARD=mod selects the reference's zone % blocks rule) */
    double sssp_delta_factor;     /* label window of the SSSP kernel in units of the mean link cost; <=0 = default */
    int sssp_block_threads;       /* 0 = default */
    int sssp_ctas_per_sm;         /* 0 = default */
    size_t tree_batch_bytes;      /* memory budget for one batch of shortest-path trees; 0 = default */
    size_t column_arena_bytes;    /* memory budget for path-link storage of the column pool; 0 = default */
    int keep_trees;               /* 1 = keep every tree of the last iteration readable (tests) */
} dtl_options;

typedef struct dtl_engine dtl_engine;

void dtl_default_options(dtl_options* opt);
dtl_engine* dtl_create(const dtl_problem* p, const dtl_options* opt);
void dtl_destroy(dtl_engine* e);
const char* dtl_last_error(void);
const char* dtl_version(void);

/* Multi-GPU: one process per GPU.  Rank 0 obtains a 128-byte NCCL id, the launcher broadcasts it,
 * every rank calls dtl_comm_init.  After that the engine all-reduces the fixed-point link-volume
 * vector once per iteration. */
int dtl_comm_get_unique_id(unsigned char id128[128]);
/* id128 == NULL selects shard-only mode: no collective is issued and the caller sums the ranks' fixed-point
 * volumes itself (used to check the sharding on one GPU). */
int dtl_comm_init(dtl_engine* e, const unsigned char id128[128]);
/* Host-only: the (agent type, period, zone) shortest-path tasks rank `rank` of `world_size` owns, in canonical
 * order (g_assign_computing_tasks_to_memory_blocks, main_api.cpp:201-275, then the rank's block of zones).
 * Returns the number of tasks (may exceed capacity; only `capacity` entries are written). */
int dtl_plan_tasks(const dtl_problem* p, int rank, int world_size, int* at, int* tau, int* zone, int capacity);

/* Host-only: groups `n` origin nodes into bundles of up to B (2, 4, 8 or 16) neighbouring origins -- the groups the
 * origin-bundle shortest-path kernel searches together: a k-d split of a two-landmark hop-distance embedding of the
 * forward star (row_ptr[n_nodes + 1], head[row_ptr[n_nodes]]).  Writes n_bundles * B entries (indices into `origins`,
 * -1 = empty slot; only the first `capacity` are written) and returns n_bundles.  Every origin appears exactly once and
 * all bundles but at most one are full.  Trees do not depend on the grouping, only the kernel's speed does. */
int dtl_plan_bundles(const int* row_ptr, const int* head, int n_nodes, const int* origins, int n, int B, int* out, int capacity);

/* SSSP tasks (agent type, period, origin zone) owned by this rank, canonical (at, tau, zone) order */
int dtl_num_tasks(const dtl_engine* e);
int dtl_get_tasks(const dtl_engine* e, int* at, int* tau, int* zone);

/* One column-generation iteration: VDF -> volume scatter (+MSA decay) -> SSSP trees -> column update.
 * Replaces the loop body main_api.cpp:589-728.
 * row = { systemTT, least_systemTT, relative_gap, avg_travel_time, counted_edges } */
int dtl_iteration(dtl_engine* e, int k, double row[5]);
/* n iterations k0 .. k0+n-1 in one call: the same results as n calls of dtl_iteration (rows: [n][5], may be NULL), issued as a pipeline
 * -- the back-trace, sums and checks of iteration k run on a second stream beside the shortest-path search of iteration k+1, whose
 * trees go to a second set of buffers (main_api.cpp:589-728 is a plain loop; nothing in it forbids the overlap: the search of k+1
 * reads the link travel times, the back-trace of k writes the column pool).  Falls back to one iteration at a time where the pipeline
 * does not apply (kept trees, the SA stage, several tree batches per iteration). */
int dtl_iterations(dtl_engine* e, int k0, int n, double* rows);
/* One column-pool gradient-projection iteration; replaces
 * g_update_gradient_cost_and_assigned_flow_in_column_pool (assignment.h:565).
 * row = { avg_travel_time, total_gap, relative_gap_percent, total_system_demand } */
int dtl_column_update(dtl_engine* e, int n, double row[4]);
/* Final scatter without decay + VDF (main_api.cpp:748-759); returns system travel time in *sys_tt. */
int dtl_finalize(dtl_engine* e, int K, double* sys_tt);
/* K iterations + CU column-pool iterations + finalize in one call.
 * iter_rows [K][5], cu_rows [CU][4] may be NULL. */
int dtl_run_assignment(dtl_engine* e, int K, int CU, double* iter_rows, double* cu_rows);

/* Link state read-back for one period (arrays of length L, NULL = skip).  pce_volume and
 * person_volume are the exact fixed-point sums converted to double. */
int dtl_get_link_state(dtl_engine* e, int tau, double* tt, double* pce_volume, double* person_volume,
                       double* link_volume, double* doc, double* voc, double* P, double* vt2);
/* exact fixed-point (Q31.32) PCE volume sums, for bit-identity checks across GPU counts */
int dtl_get_volume_fixed(dtl_engine* e, int tau, long long* pce_volume_q32);
/* person volume of one agent type (person_volume_per_period_per_at, assignment.h:142) */
int dtl_get_person_volume_at(dtl_engine* e, int tau, int at, double* person_volume);
/* full QVDF outputs of the current link volumes, for link_performance.csv */
int dtl_get_link_detail(dtl_engine* e, int tau, double* avg_queue_speed, double* avg_speed_bpr, double* Q_mu,
                        double* Q_gamma, double* t0, double* t3, double* severe_congestion_P);
/* 5-minute speed profile (CPeriod_VDF::model_speed, VDF.h:277-319) of slots [first_slot, first_slot+n_slots) */
int dtl_get_model_speed(dtl_engine* e, int tau, int first_slot, int n_slots, float* speed /*[L][n_slots]*/);
/* generalized link cost of forward star (at, tau) as used by the last SSSP (routing.h:229-231); [L] */
int dtl_get_gen_cost(dtl_engine* e, int at, int tau, double* cost);

/* column pool read-back, columns in (od index, node_sum ascending) order of this rank's OD cells */
long long dtl_num_columns(dtl_engine* e);
long long dtl_num_column_links(dtl_engine* e);
int dtl_get_columns(dtl_engine* e, int* od_index, int* node_sum, int* seq_no, int* n_links, int* n_nodes,
                    double* volume, double* cost, double* time, int* links);

/* ---- ODME: OD demand estimation against link counts, Assignment::Demand_ODME (ODME.h:91-516) and its scatter
 * g_reset_and_update_link_volume_based_on_ODME_columns (assignment.h:263-563).  Runs on the column pool the assignment left.
 *   dtl_odme_configure  at_pce / at_occ [AT] = agent-type PCE / person occupancy as floats (assignment.h:308-309); n "link"
 *                       rows of measurement.csv in file order (period, link, count, upper_bound_flag; ODME.h:140-178)
 *   dtl_odme_scatter    column volumes -> link volumes with the ODME addend, VDF pass, count deviations; row = { link MAPE / 100
 *                       (the convergence gap), system MPE / 100, link MAE, avg travel time, UE gap per unit of demand, UE gap %,
 *                       total demand, links with a count }
 *   dtl_odme_update     the path-flow gradient step (ODME.h:262-488)
 *   dtl_odme_run        the whole loop with the reference's stopping rule; rows [n_iterations + 1][8]
 *   dtl_odme_get        per link of a period: count, upper-bound flag, count deviation
 * Volumes are integer (Q31.32) sums as in the assignment: independent of the visiting order and of the GPU count, within 2^-33
 * per addend of the reference's serial FP64 sums. */
int dtl_odme_configure(dtl_engine* e, const float* at_pce, const float* at_occ, int n, const int* tau, const int* link,
                       const float* count, const int* upper_bound_flag);
int dtl_odme_scatter(dtl_engine* e, int s, double row[8]);
int dtl_odme_update(dtl_engine* e);
int dtl_odme_run(dtl_engine* e, int n_iterations, double* rows, int* n_done);
int dtl_odme_get(dtl_engine* e, int tau, double* obs_count, int* upper_bound_flag, double* est_count_dev);

/* ---- sensitivity-analysis stage: stage II of network_assignment, main_api.cpp:771-1042 ----
 * Covered: supply-side "sa" lane changes (scenario_API.h:152-160), the re-routing of the agent types with real-time
 * information, and "dms" scenario rows with their information zones and the route modification (below); demand-side
 * scenario factors are not.
 * Call order: dtl_sa_configure (any time after dtl_create), the assignment (dtl_iteration..., dtl_finalize), dtl_sa_begin,
 * dtl_sa_iteration(0..S), dtl_sa_column_update(0..S-1).
 *   dtl_sa_configure     rt_info [AT] = real_time_information of the agent types (0 none, 1 real-time, 2 dms; DTA.h:486);
 *                        n_changes lane changes (period, link, lanes)
 *   dtl_sa_begin         applies the lane changes (main_api.cpp:779-794) and refreshes the travel times (:857)
 *   dtl_sa_iteration     one column re-generation iteration (:864-1034): K4, K3 (only columns of agent types with real-time
 *                        information are decayed, assignment.h:150-176), K1 + K2a for their origins; row = { sysTT, avg travel time }
 *   dtl_sa_column_update one column-pool iteration under the SA rules (assignment.h:699-713, 828-846); row as dtl_column_update
 *   dtl_sa_get_columns   per column, in dtl_get_columns order: the cell's impact flag and the per-iteration path time / volume
 * "dms" scenario rows (scenario_API.h:165-188; one GPU only):
 *   dtl_sa_set_dms       after dtl_sa_configure: n message signs (period, link, information zone = zone seq no of the link's
 *                        to-node, -1 if it is no zone of the model).  Message-sign users (rt_info 2) are then re-routed in the SA
 *                        stage from origins that are information zones (assignment.h:160-170, main_api.cpp:990-994)
 *   dtl_sa_route_modification   g_column_pool_route_modification (assignment.h:1077-1383, main_api.cpp:1036), between the last
 *                        dtl_sa_iteration and the first dtl_sa_column_update: splits the columns of message-sign users that pass a
 *                        sign and a lane change; *n_split (may be NULL) = columns that were split */
int dtl_sa_configure(dtl_engine* e, const int* rt_info, int n_changes, const int* sa_tau, const int* sa_link, const double* sa_lanes);
int dtl_sa_begin(dtl_engine* e, double* sys_tt);
int dtl_sa_iteration(dtl_engine* e, int it, double row[2]);
int dtl_sa_column_update(dtl_engine* e, int n, double row[4]);
int dtl_sa_set_dms(dtl_engine* e, int n, const int* dms_tau, const int* dms_link, const int* dms_zone);
int dtl_sa_route_modification(dtl_engine* e, int* n_split);
/* per column, in dtl_get_columns order: impacted_path_flag as dtl_sa_route_modification left it (1 = passes a lane change, 3 = the new,
 * diverted column; assignment.h:1179,1356) and the information_type of its OD cell (assignment.h:1360); either may be NULL */
int dtl_sa_get_route_flags(dtl_engine* e, unsigned char* path_impact, unsigned char* info_type);
int dtl_sa_get_columns(dtl_engine* e, unsigned char* od_impact, int* n_iterations, double* hist_time, double* hist_volume);

/* ---- kernel-level entry points (known-answer tests, benchmarks) ---- */

/* Batched many-source SSSP; replaces NetworkForSP::optimal_label_correcting (routing.h:633) for
 * n origin zones of forward star (at, tau).  cost: [L] host array of generalized link costs, or
 * NULL to use the engine's current costs.  Outputs (host, may be NULL): labels [n][N],
 * pred_node [n][N], pred_link [n][N]; tie_nodes [n] = nodes with >= 2 tight predecessors from
 * different upstream nodes (such origins are resolved by the serial replay kernel).
 * stats (may be NULL): { device ms, counted edges, relaxations, queue pops, rounds, replayed origins } */
int dtl_sssp_trees(dtl_engine* e, int at, int tau, const double* cost, const int* zones, int n, double* labels,
                   int* pred_node, int* pred_link, int* tie_nodes, double stats[6]);
/* Backward one-to-all search from n destination nodes; replaces
 * NetworkForSP::optimal_backward_label_correcting_from_destination (routing.h:1009-1273), the search the reference's
 * real-time-information layer runs per destination zone centroid (main_api.cpp:293-321, assignment.h:1386-1414).
 * The star is every node's in-link list without the links that carry an outgoing-connector zone tag (routing.h:1104) and
 * without the links whose rt_allowed[l] == 0 (RT_allowed_use, routing.h:1119; NULL = all allowed).  cost: [L] host
 * array of avg_travel_time + RT_waiting_time (routing.h:1148), or NULL for the engine's current travel times of period tau.
 * labels [n][N] = least cost from every node TO the destination, pred_node [n][N] = next node downstream,
 * pred_link [n][N] = link to take; tie_nodes and stats as for dtl_sssp_trees. */
int dtl_backward_trees(dtl_engine* e, int tau, const double* cost, const unsigned char* rt_allowed, const int* dest_nodes,
                       int n, double* labels, int* pred_node, int* pred_link, int* tie_nodes, double stats[6]);
/* Elementwise link performance; replaces CPeriod_VDF::calculate_travel_time_based_on_QVDF
 * (VDF.h:139) over all links of period tau for the given volumes [L]. */
int dtl_vdf(dtl_engine* e, int tau, const double* volume, double* tt, double* doc, double* voc, double* P,
            double* vt2);
/* Path->link volume accumulation of the current column pool; replaces
 * g_reset_and_update_link_volume_based_on_columns (assignment.h:53).  decay_k < 0 = no MSA decay. */
int dtl_scatter(dtl_engine* e, int decay_k);
/* trees of the last dtl_iteration (needs keep_trees): task index in dtl_get_tasks order */
int dtl_get_tree(dtl_engine* e, int task, double* label, int* pred_node, int* pred_link);

/* device-time accounting (CUDA events on the engine's stream), milliseconds since the last reset:
 * ms[0]=VDF  ms[1]=scatter  ms[2]=SSSP  ms[3]=tie replay  ms[4]=backtrace/columns  ms[5]=column pool
 * ms[6]=all-reduce  ms[7]=whole iterations;  counts[i] = kernel launches in each class */
int dtl_get_timing(dtl_engine* e, double ms[8], long long counts[8], int reset);
/* The assignment loop runs the scatter of iteration k (+ its all-reduce) on a second stream beside the shortest-path search of
 * the same iteration (the scatter only feeds the VDF pass of iteration k+1; main_api.cpp:606-699 has no such dependency either,
 * it is simply serial).  on = 0 puts every kernel back on one stream -- same results; for per-kernel timing (ms[1] is the
 * duration of the scatter ALONE only then).  Returns the previous setting. */
int dtl_set_overlap(dtl_engine* e, int on);
/* Destination-bounded shortest-path search, ON by default (DTL_COMPLETE_SEARCH=1 in the environment or on = 0 turns it off).  The
 * reference computes every one-to-all tree to the end (routing.h:633-939) although the assignment loop only back-traces the
 * destinations an origin has demand to (routing.h:428-520).  on = 1: inside the assignment loop a search (origin bundles, shared-memory
 * queues) stops as soon as nothing that is still queued can improve the label of such a destination -- exact: the paths, columns, link
 * volumes and gaps are the same bits; only the labels beyond the bound (which nothing reads) are not final.  on = 0: every search runs
 * to the end.  Calls that hand out trees (dtl_sssp_trees, keep_trees), the Graph500 edge count of the first iteration, the global-queue
 * kernel and the overflow re-runs always run complete searches.  Returns the previous setting. */
int dtl_set_bounded_search(dtl_engine* e, int on);
/* work counters since the last reset: { counted edges (Graph500 convention), relaxations, queue pops,
 * SSSP rounds, SSSP tasks, replayed origins, column-link references scattered, back-traced path nodes,
 * trees re-run with worst-case queues, tie candidates observed } */
int dtl_get_counters(dtl_engine* e, double c[10], int reset);
/* which K1 kernel the last shortest-path launch used: 0 = global-memory queues, 1 = shared-memory queues,
 * 2 = origin bundles, round-synchronous search (search + label->tree pass, two launches), 3 = origin bundles, search
 * without round barriers (the default for bundles); -1 before the first launch */
int dtl_sssp_kernel_used(const dtl_engine* e);
/* number of HBM bytes the engine holds */
size_t dtl_device_bytes(const dtl_engine* e);

/* ---- GMNS file contract (host side; no GPU needed) ---- */

/* Reads settings.csv, node.csv, link.csv and the [demand_file_list] files from `dir` exactly as
 * g_read_input_data / g_ReadDemandFileBasedOnDemandFileList do (input.h:2388, :392) and adds the
 * virtual connectors (main_api.cpp:534-570).  The returned image is owned by the library. */
typedef struct dtl_gmns dtl_gmns;
dtl_gmns* dtl_gmns_read(const char* dir, int number_of_memory_blocks);
const dtl_problem* dtl_gmns_problem(const dtl_gmns* g);
void dtl_gmns_free(dtl_gmns* g);
const char* dtl_gmns_last_error(void);
/* identifiers for writers / tests: node_id [N], zone_id [Z], link_id strings are not exposed */
int dtl_gmns_ids(const dtl_gmns* g, int* node_id, int* zone_id);
/* what the sensitivity-analysis stage needs beyond dtl_problem: rt_info [AT] = [agent_type] real_time_info (input.h:2652-2658:
 * 0 none, 1 real-time information, 2 dms) and the "sa" rows of supply_side_scenario.csv (scenario_API.h:152-160) as
 * (period, link, lanes) triples.  Returns the number of sa rows (at most `capacity` are copied) or -1. */
int dtl_gmns_scenario(const dtl_gmns* g, int* rt_info, int* sa_tau, int* sa_link, double* sa_lanes, int capacity);
/* the "dms" rows of supply_side_scenario.csv (scenario_API.h:165-188): period, link and the zone seq no of the link's to-node (the
 * information zone; -1 when its zone id is no zone of the model).  Returns their number (feed them to dtl_sa_set_dms). */
int dtl_gmns_dms(const dtl_gmns* g, int* dms_tau, int* dms_link, int* dms_zone, int capacity);
/* ODME inputs: the agent types' PCE / person occupancy as floats (assignment.h:308-309) and the "link" rows of measurement.csv
 * (ODME.h:100-178), one entry per (row, demand period) in file order: (period, link, count, upper_bound_flag).  Returns the
 * number of entries (at most `capacity` are copied) or -1. */
int dtl_gmns_measurements(const dtl_gmns* g, float* at_pce, float* at_occ, int* ms_tau, int* ms_link, float* ms_count, int* ms_ub,
                          int capacity);
/* settings.csv [assignment] as flash_dta.cpp:171-242 derives them: out = { CG iterations, CU iterations,
 * ODME iterations, SA iterations, simulation iterations, memory blocks } */
int dtl_gmns_assignment_settings(const char* dir, int out[6]);
/* link_performance.csv / route_assignment.csv / final_summary.csv writers (output.h:486-1459) */
int dtl_gmns_write_outputs(const dtl_gmns* g, dtl_engine* e, const char* dir);

#ifdef __cplusplus
}
#endif
#endif
