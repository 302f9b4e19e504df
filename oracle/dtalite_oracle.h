/* TEST INFRASTRUCTURE ONLY -- the oracle is the checker, never the product.
 *
 * Plain-C, single-threaded restatement of the DTALite traffic-assignment hot path of
 * asu-trans-ai-lab/DLSim-MRM (src/cpp).  Every function in dtalite_oracle.c cites the
 * reference file:line it follows.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg may load this library.
 *
 * Parity status: PINNED.  tests/test_oracle_pins.py checks it bit-for-bit against the
 * reference's own golden vectors (src/test_case_01_two_corridor/final_summary.csv:35-58,
 * dataset/4_101_corridor/final_summary.csv:150-177) and against state dumps of the
 * reference engine itself (oracle/_ref/dtalite_ref, built from the unmodified sources).
 */
#ifndef DTL_ORACLE_H
#define DTL_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_LABEL_COST 1.0e15 /* DTA.h:11 */
#define ORC_NO_PRED (-9999)       /* routing.h:683-685 */

/* Array image of the network exactly as the reference holds it after g_read_input_data +
 * connector creation (SURVEY.md section 9 "Graph image").  Per-period arrays are [T][L],
 * per-period-per-agent arrays are [T][AT][L], row-major. */
typedef struct orc_problem {
    int n_nodes, n_links, n_zones, n_agent_types, n_periods;
    int memory_blocks;            /* number_of_memory_blocks after clamping, main_api.cpp:519 */
    const int* link_from;         /* [L] from_node_seq_no */
    const int* link_to;           /* [L] to_node_seq_no */
    const int* link_tag;          /* [L] zone_seq_no_for_outgoing_connector, -1 if none */
    const int* node_zone;         /* [N] zone seq no for centroid nodes (CNode.zone_id>=1), else -1 */
    const int* zone_centroid;     /* [Z] centroid node seq no of each zone */
    /* CPeriod_VDF fields, VDF.h:369-466 */
    const double *fftt, *cap, *alpha, *beta, *vf, *vc, *nlanes, *cap_lane, *qdf, *preload, *penalty;
    const double *Q_alpha, *Q_beta, *Q_cd, *Q_n, *Q_cp, *Q_s, *t2, *L_h, *start_h, *end_h, *plf;
    const float *cycle_length, *red_time, *sat_flow;
    const int* vdf_type;          /* 0 = q_vdf, 1 = bpr_vdf, DTA.h:53 */
    const double *pce, *occ, *toll, *lr_price; /* [T][AT][L] */
    const unsigned char* allowed; /* [T][AT][L] CLink::AllowAgentType, DTA.h:1352 */
    const float* vot;             /* [AT] value_of_time (float, DTA.h:479) */
    /* demand: OD cells with od_volume > 0 sorted by (o, d, at, tau) */
    int n_od;
    const int *od_o, *od_d, *od_at, *od_tau;
    const double* od_vol;
    const float* origin_demand;   /* [Z] g_origin_demand_array */
    float total_demand_volume;    /* assignment.total_demand_volume (float) */
} orc_problem;

typedef struct orc_state orc_state;

orc_state* orc_create(const orc_problem* p);
void orc_destroy(orc_state* s);

/* number of SSSP tasks (agent type, period, origin zone) per iteration and their identity,
 * in canonical (at, tau, zone ascending) order */
int orc_num_tasks(const orc_state* s);
void orc_get_tasks(const orc_state* s, int* at, int* tau, int* zone);

/* When set (non-NULL), orc_iteration stores every task's tree: labels [n_tasks][N],
 * pred_node [n_tasks][N], pred_link [n_tasks][N] in canonical task order. */
void orc_set_tree_capture(orc_state* s, double* labels, int* pred_node, int* pred_link);

/* One column-generation iteration (main_api.cpp:589-728).
 * row = { sysTT, least_total, relative_gap, avg_travel_time, counted_edges } */
void orc_iteration(orc_state* s, int k, double row[5]);
/* One column-pool iteration (assignment.h:565-928).
 * row = { avg_travel_time, total_gap, relative_gap_percent, total_system_demand } */
void orc_column_update(orc_state* s, int n, double row[4]);
/* Final scatter (no decay) + VDF, main_api.cpp:748-759.  Returns system travel time. */
double orc_finalize(orc_state* s, int K);

/* ---- sensitivity-analysis stage (main_api.cpp:771-1042): supply-side "sa" lane changes + re-routing of the agent types
 * with real-time information.  Call order: orc_sa_configure, orc_sa_lane_change*, [stage 1/2 + orc_finalize], orc_sa_begin,
 * orc_sa_iteration(0..S), orc_sa_column_update(0..S-1). ---- */
void orc_sa_configure(orc_state* s, const int* rt_info /* [AT] real_time_information: 0 none, 1 rt, 2 dms */);
void orc_sa_lane_change(orc_state* s, int tau, int link, double lanes);
void orc_sa_dms_link(orc_state* s, int tau, int link, int has_zone, int zone); /* "dms" row: information zone of the link's to-node */
int orc_sa_route_modification(orc_state* s); /* g_column_pool_route_modification (assignment.h:1077-1383), after the last orc_sa_iteration */
/* per column in orc_get_columns order: impacted_path_flag; per column: information_type of its pool, od-impact flag of its cell */
void orc_get_column_flags(const orc_state* s, int* impacted, int* info_type, int* od_impact);
double orc_sa_begin(orc_state* s);
void orc_sa_iteration(orc_state* s, int it, double row[3]); /* { sysTT, avg travel time, least system TT of the re-routed origins } */
void orc_sa_column_update(orc_state* s, int n, double row[4]);

/* ---- ODME: Assignment::Demand_ODME, ODME.h:91-516, and its scatter
 * g_reset_and_update_link_volume_based_on_ODME_columns, assignment.h:263-563.  Loop of the reference: for s in 0..n-1:
 * gap = scatter(s); stop if s >= 5 and gap - prev_gap > 0.01; update(); then one more scatter(n). ---- */
void orc_odme_configure(orc_state* s, const float* at_pce, const float* at_occ);
void orc_odme_sensor(orc_state* s, int tau, int link, float count, int upper_bound_flag);
void orc_odme_scatter(orc_state* s, int iteration_no, double row[8]);
void orc_odme_update(orc_state* s);

/* state read-back; per-period arrays of length L */
void orc_get_link_state(const orc_state* s, int tau, double* tt, double* pce_volume, double* person_volume,
                        double* link_volume, double* doc, double* voc, double* P, double* vt2);
void orc_get_model_speed(const orc_state* s, int link, float* speed300);
void orc_get_gen_cost(const orc_state* s, int at, int tau, double* cost);

long long orc_num_columns(const orc_state* s);
long long orc_num_column_links(const orc_state* s);
/* columns in (o, d, at, tau, node_sum ascending) order; links concatenated in the same order */
void orc_get_columns(const orc_state* s, int* od_index, int* node_sum, int* seq_no, int* n_links, int* n_nodes,
                     double* volume, double* cost, double* time, int* links);

/* ---- stand-alone kernels (known-answer tests for the CUDA kernels) ---- */

/* forward star of (at, tau): CSR in the reference's scan order (routing.h:257-313) */
void orc_forward_star(const orc_problem* p, int at, int tau, int* row_ptr /*[N+1]*/, int* col_link, int* col_to,
                      int* n_edges);
/* deque label-correcting SSSP (routing.h:633-882).  Returns the number of pops;
 * *relaxations = out-edges examined after the connector filter. */
long long orc_sssp(const orc_problem* p, const int* row_ptr, const int* col_link, const int* col_to,
                   const double* cost /*[L]*/, int origin_zone, double* label, int* pred_node, int* pred_link,
                   long long* relaxations);
/* backward deque label-correcting search from a destination node over the in-link lists
 * (NetworkForSP::optimal_backward_label_correcting_from_destination, routing.h:1009-1273).
 * cost [L] = avg_travel_time + RT_waiting_time; rt_allowed [L] or NULL (RT_allowed_use).
 * label = least cost to the destination, pred_node = next node downstream.  Returns the pops. */
long long orc_backward_sssp(const orc_problem* p, const double* cost, const unsigned char* rt_allowed, int dest_node,
                            double* label, int* pred_node, int* pred_link);
/* VDF for one (link, period) slot given the volume to be assigned (VDF.h:139-367) */
typedef struct orc_vdf_out {
    double tt, doc, voc, P, vt2, avg_queue_speed, avg_speed_bpr, Q_mu, Q_gamma, t0, t3, severe_congestion_P,
        lane_based_D, avg_waiting_time;
} orc_vdf_out;
void orc_vdf_link(const orc_problem* p, int tau, int link, double volume, orc_vdf_out* out, float* model_speed300);

#ifdef __cplusplus
}
#endif
#endif
