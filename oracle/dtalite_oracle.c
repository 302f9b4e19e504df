/* TEST INFRASTRUCTURE ONLY -- see dtalite_oracle.h.  Plain C99, single-threaded, libm only.
 * Compile WITHOUT -ffast-math and WITHOUT FMA contraction (-ffp-contract=off): the reference
 * is built for baseline x86-64 where every FP64 operation is individually rounded. */
#include "dtalite_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define MAX_SPEED_SLOTS 300 /* MAX_TIMEINTERVAL_PerDay, DTA.h:24 */

static double dmax(double a, double b) { return a > b ? a : b; } /* macro semantics (a>b)?a:b */
static double dmin(double a, double b) { return a < b ? a : b; }

typedef struct {
    int node_sum, seq_no, n_links, n_nodes;
    double volume, cost, time, toll;
    int impacted;   /* impacted_path_flag, DTA.h:523 */
    int* links; /* origin -> destination order, DTA.h:551-559 */
} orc_column;

typedef struct {
    int n, cap;
    int info_type;  /* CColumnVector::information_type, DTA.h:698 */
    orc_column* c; /* ascending node_sum == std::map iteration order */
} orc_colvec;

typedef struct {
    int at, tau, n_origins;
    int* zones; /* m_origin_zone_seq_no_vector */
} orc_net;

struct orc_state {
    const orc_problem* p;
    int N, L, Z, AT, T;
    /* forward stars per (at,tau) */
    int **row_ptr, **col_link, **col_to;
    /* link state per tau */
    double **tt, **pce_vol, **person_vol, **link_volume, **doc, **voc, **P, **vt2;
    double** person_vol_at; /* [tau*AT+at][L] */
    double** gen_cost;      /* [at*T+tau][L] */
    float* model_speed;     /* [L][300] or NULL */
    /* networks / tasks */
    int n_nets, n_blocks;
    orc_net* nets;
    int n_tasks;
    int *task_at, *task_tau, *task_zone;
    /* column pool */
    orc_colvec* pool; /* [n_od] */
    int* od_first;    /* [Z+1] first OD index of each origin */
    /* SSSP scratch */
    double* label;
    int *pred_node, *pred_link, *status, *next;
    int *tmp_nodes, *tmp_links;
    int* n_out_unfiltered; /* [N] */
    /* capture */
    double* cap_label;
    int *cap_pn, *cap_pl;
    /* sensitivity-analysis stage (main_api.cpp:771-1042) */
    orc_problem pm;            /* the problem with this state's own fftt / cap / nlanes (the sa lane change edits them) */
    double *m_fftt, *m_cap, *m_nlanes;
    int* rt_info;              /* [AT] real_time_information, DTA.h:486 */
    double* sa_lanes;          /* [T][L] sa_lanes_change */
    signed char* design_flag;  /* [T][L] network_design_flag, VDF.h:419 */
    unsigned char* od_impact;  /* [n_od] OD_impact_flag && at_od_impacted_flag_map[at] of the cell (assignment.h:699-713) */
    /* ODME (ODME.h:91-516) */
    double* obs_count;         /* [T][L] VDF_period.obs_count, 0 = no data */
    int* ub_flag;              /* [T][L] upper_bound_flag (default 1, VDF.h:60) */
    double* est_dev;           /* [T][L] est_count_dev */
    float *at_pce, *at_occ;    /* [AT] agent-type PCE / OCC as the ODME scatter uses them (float, assignment.h:308-309) */
    /* dms scenario rows (scenario_API.h:165-188) */
    int* info_zone;            /* [Z] 1 = information zone (zone_seq_no_2_info_mapping) */
    int* dms_zone;             /* [T][L] zone seq no of the to-node of a dms link (-1: none / not a zone of the model) */
    int* dms_has_zone;         /* [T][L] the to-node of the dms link carries a zone id (node_seq_no_2_zone_id_mapping) */
};

/* ------------------------------------------------------------------------------------------ */
/* Forward star of one (agent type, period): NetworkForSP::BuildNetwork, routing.h:257-313.
 * A node's out-links are visited in insertion order (input.h:3638-3642, 1897-1903); links are
 * appended to g_link_vector in the same order, so insertion order == ascending link seq no. */
void orc_forward_star(const orc_problem* p, int at, int tau, int* row_ptr, int* col_link, int* col_to, int* n_edges)
{
    const int N = p->n_nodes, L = p->n_links;
    const unsigned char* allowed = p->allowed + ((size_t)tau * p->n_agent_types + at) * L;
    memset(row_ptr, 0, sizeof(int) * (N + 1));
    for (int l = 0; l < L; ++l)
        if (allowed[l]) row_ptr[p->link_from[l] + 1]++;
    for (int i = 0; i < N; ++i) row_ptr[i + 1] += row_ptr[i];
    int* cursor = (int*)malloc(sizeof(int) * (N + 1));
    memcpy(cursor, row_ptr, sizeof(int) * (N + 1));
    for (int l = 0; l < L; ++l)
        if (allowed[l]) {
            int e = cursor[p->link_from[l]]++;
            col_link[e] = l;
            col_to[e] = p->link_to[l];
        }
    free(cursor);
    if (n_edges) *n_edges = row_ptr[N];
}

/* ------------------------------------------------------------------------------------------ */
/* CPeriod_VDF::calculate_travel_time_based_on_QVDF, VDF.h:139-367 */
void orc_vdf_link(const orc_problem* p, int tau, int link, double volume, orc_vdf_out* o, float* model_speed)
{
    const size_t i = (size_t)tau * p->n_links + link;
    const double qdf = p->qdf[i], nlanes = p->nlanes[i], cap_lane = p->cap_lane[i];
    const double vc = p->vc[i], vf = p->vf[i], Q_alpha = p->Q_alpha[i], Q_beta = p->Q_beta[i];
    const double beta = p->beta[i], alpha = p->alpha[i], FFTT = p->fftt[i], BPR_cap = p->cap[i];
    const double Q_cd = p->Q_cd[i], Q_n = p->Q_n[i], Q_cp = p->Q_cp[i], Q_s = p->Q_s[i];
    const double L = p->L_h[i], t2 = p->t2[i], plf = p->plf[i];
    const double start_h = p->start_h[i], end_h = p->end_h[i];
    const float cycle_length = p->cycle_length[i], red_time = p->red_time[i], sat = p->sat_flow[i];

    double lane_based_D = dmax(0.0, volume) * qdf / dmax(0.000001, nlanes);              /* :145 */
    double DOC = lane_based_D / dmax(0.00001, cap_lane);                                  /* :149 */
    double avg_queue_speed = vc / (1.0 + Q_alpha * pow(DOC, Q_beta));                     /* :156 */
    double RTT = 1.0 / vc;                                                                /* :162 */
    double Q_n_current = Q_n;
    if (DOC < 1.0) {                                                                      /* :164-177 */
        double vf_alpha = (1.0 + Q_alpha) * vf / dmax(0.0001, vc) - 1.0;
        double vf_avg_speed = vf / (1.0 + vf_alpha * pow(DOC, beta));
        avg_queue_speed = vf_avg_speed;
        Q_n_current = beta;
        RTT = 1.0 / dmax(0.01, vf);
    }
    double VOC = volume / dmax(0.00001, BPR_cap);                                         /* :184 */
    double avg_speed_BPR = vf / (1.0 + alpha * pow(VOC, beta));                           /* :191 */
    double avg_tt = FFTT * vf / dmax(0.1, avg_speed_BPR);                                 /* :192 */
    if (p->vdf_type[i] == 0 /* q_vdf */) avg_tt = FFTT * vf / dmax(0.1, avg_queue_speed); /* :196-202 */
    if (cycle_length >= 1) {                                                              /* :204-210 */
        float s_bar = (float)(1.0 / 60.0 * red_time * red_time / (2 * cycle_length));
        double lambda = lane_based_D;
        float uniform_delay = (float)(s_bar / dmax(1 - lambda / sat, (double)0.1f));
        avg_tt = uniform_delay + FFTT;
    }
    double P = Q_cd * pow(DOC, Q_n_current);                                              /* :217 */
    double base = Q_cp * pow(P, Q_s) + 1.0;
    double vt2 = vc / dmax(0.001, base);                                                  /* :220 */
    double nonpeak_hourly_flow = 0;
    if (L - P >= 10.0 / 60.0)
        nonpeak_hourly_flow = (volume * (1 - plf)) / dmax(1.0, nlanes) / dmax(0.1, dmin(L - 1, L - P - 5.0 / 60.0));
    if (nonpeak_hourly_flow > cap_lane) nonpeak_hourly_flow = cap_lane;
    double t0 = t2 - 0.5 * P, t3 = t2 + 0.5 * P;                                          /* :240-241 */
    double Q_mu = dmin(cap_lane, lane_based_D / P);                                       /* :249 */
    double wt2 = 1.0 / vt2 - RTT;                                                         /* :256 */
    double Q_gamma = wt2 * 64 * Q_mu / pow(P, 4);                                         /* :260 */
    double td_w = 0, severe = 0;
    for (int t_in_min = (int)(start_h * 60); t_in_min <= end_h * 60; t_in_min += 5) {     /* :277-319 */
        double t = t_in_min / 60.0, td_queue, td_speed;
        if (t0 <= t && t <= t3) {
            td_queue = 0.25 * Q_gamma * pow((t - t0), 2) * pow((t - t3), 2);
            td_w = td_queue / dmax(0.001, Q_mu);
            td_speed = 1.0 / (td_w + RTT);
        } else if (t < t0) {
            double factor = (t - start_h) / dmax(0.001, t0 - start_h);
            td_speed = (1 - factor) * vf + factor * dmax(vc, avg_queue_speed);
        } else {
            double factor = (t - t3) / dmax(0.001, end_h - t3);
            td_speed = (1 - factor) * dmax(vc, avg_queue_speed) + (factor)*vf;
        }
        int t_interval = t_in_min / 5;
        if (model_speed && t_interval >= 0 && t_interval < MAX_SPEED_SLOTS) model_speed[t_interval] = (float)td_speed;
        if (td_speed < vf * 0.5) severe += 5.0 / 60;
    }
    o->tt = avg_tt; o->doc = DOC; o->voc = VOC; o->P = P; o->vt2 = vt2;
    o->avg_queue_speed = avg_queue_speed; o->avg_speed_bpr = avg_speed_BPR;
    o->Q_mu = Q_mu; o->Q_gamma = Q_gamma; o->t0 = t0; o->t3 = t3; o->severe_congestion_P = severe;
    o->lane_based_D = lane_based_D; o->avg_waiting_time = avg_tt - FFTT;
}

/* ------------------------------------------------------------------------------------------ */
/* NetworkForSP::optimal_label_correcting, routing.h:633-882 (deque / Pape label correcting) */
static long long sssp_core(int N, const int* row_ptr, const int* col_link, const int* col_to, const double* cost,
                           const int* link_tag, int origin_node, int origin_zone, double* label, int* pred_node,
                           int* pred_link, int* status, int* next, long long* relax_out)
{
    long long pops = 0, relax = 0;
    for (int i = 0; i < N; ++i) {                                                         /* :677-689 */
        status[i] = 0; label[i] = ORC_MAX_LABEL_COST; pred_link[i] = ORC_NO_PRED; pred_node[i] = ORC_NO_PRED;
    }
    if (row_ptr[origin_node + 1] - row_ptr[origin_node] == 0) { if (relax_out) *relax_out = 0; return 0; } /* :692 */
    label[origin_node] = 0.0;
    int front = origin_node, tail = origin_node;                                          /* :706-707 */
    next[origin_node] = -1;
    while (front != -1) {                                                                 /* :716 */
        int u = front;                                                                    /* :721-726 */
        front = next[u];
        next[u] = -1;
        status[u] = 2;
        ++pops;
        for (int e = row_ptr[u]; e < row_ptr[u + 1]; ++e) {                               /* :738 */
            int v = col_to[e], l = col_link[e];
            if (link_tag[l] >= 0 && link_tag[l] != origin_zone) continue;                 /* :779-786 */
            ++relax;
            double c = label[u] + cost[l];                                                /* :804 */
            if (c < label[v]) {                                                           /* :816 */
                label[v] = c; pred_node[v] = u; pred_link[v] = l;                         /* :827-831 */
                if (status[v] == 0) {                                                     /* :840-859 push back */
                    if (front == -1) { front = v; tail = v; next[v] = -1; }
                    else { next[tail] = v; next[v] = -1; tail = v; }
                    status[v] = 1;
                }
                if (status[v] == 2) {                                                     /* :861-879 push front */
                    if (front == -1) { next[v] = -1; front = v; tail = v; }
                    else { next[v] = front; front = v; }
                    status[v] = 1;
                }
            }
        }
    }
    if (relax_out) *relax_out = relax;
    return pops;
}

long long orc_sssp(const orc_problem* p, const int* row_ptr, const int* col_link, const int* col_to,
                   const double* cost, int origin_zone, double* label, int* pred_node, int* pred_link,
                   long long* relaxations)
{
    int N = p->n_nodes;
    int* status = (int*)malloc(sizeof(int) * N);
    int* next = (int*)malloc(sizeof(int) * N);
    long long pops = sssp_core(N, row_ptr, col_link, col_to, cost, p->link_tag, p->zone_centroid[origin_zone],
                               origin_zone, label, pred_node, pred_link, status, next, relaxations);
    free(status); free(next);
    return pops;
}

/* NetworkForSP::optimal_backward_label_correcting_from_destination, routing.h:1009-1273: the same deque rule, scanning the
 * in-links of the popped node (g_node_vector[].m_incoming_link_seq_no_vector: every link, in insertion order == ascending
 * link seq no, no agent-type filter) */
long long orc_backward_sssp(const orc_problem* p, const double* cost, const unsigned char* rt_allowed, int dest_node,
                            double* label, int* pred_node, int* pred_link)
{
    const int N = p->n_nodes, L = p->n_links;
    int* in_ptr = (int*)calloc((size_t)N + 1, sizeof(int));
    int* in_link = (int*)malloc(sizeof(int) * (size_t)(L > 0 ? L : 1));
    int* cursor = (int*)malloc(sizeof(int) * ((size_t)N + 1));
    int* status = (int*)malloc(sizeof(int) * (size_t)N);
    int* next = (int*)malloc(sizeof(int) * (size_t)N);
    for (int l = 0; l < L; ++l) in_ptr[p->link_to[l] + 1]++;
    for (int i = 0; i < N; ++i) in_ptr[i + 1] += in_ptr[i];
    memcpy(cursor, in_ptr, sizeof(int) * ((size_t)N + 1));
    for (int l = 0; l < L; ++l) in_link[cursor[p->link_to[l]]++] = l;
    long long pops = 0;
    for (int i = 0; i < N; ++i) {                                                         /* :1042-1051 */
        status[i] = 0; label[i] = ORC_MAX_LABEL_COST; pred_link[i] = ORC_NO_PRED; pred_node[i] = ORC_NO_PRED;
    }
    label[dest_node] = 0.0;                                                               /* :1055 */
    int front = dest_node, tail = dest_node;                                              /* :1064-1065 */
    next[dest_node] = -1;
    while (front != -1) {                                                                 /* :1074 */
        int to = front;                                                                   /* :1077-1082 */
        front = next[to];
        next[to] = -1;
        status[to] = 2;
        ++pops;
        for (int j = in_ptr[to]; j < in_ptr[to + 1]; ++j) {                               /* :1094 */
            int l = in_link[j], from = p->link_from[l];
            if (p->link_tag[l] >= 0) continue;                                            /* :1104-1108 */
            if (rt_allowed && !rt_allowed[l]) continue;                                   /* :1119 */
            double c = label[to] + cost[l];                                               /* :1148 */
            if (c < label[from]) {                                                        /* :1162 */
                label[from] = c; pred_node[from] = to; pred_link[from] = l;               /* :1170-1174 */
                if (status[from] == 0) {                                                  /* :1183-1201 push back */
                    if (front == -1) { front = from; tail = from; next[from] = -1; }
                    else { next[tail] = from; next[from] = -1; tail = from; }
                    status[from] = 1;
                }
                if (status[from] == 2) {                                                  /* :1203-1219 push front */
                    if (front == -1) { next[from] = -1; front = from; tail = from; }
                    else { next[from] = front; front = from; }
                    status[from] = 1;
                }
            }
        }
    }
    free(in_ptr); free(in_link); free(cursor); free(status); free(next);
    return pops;
}

/* ------------------------------------------------------------------------------------------ */
static double** alloc2(int n, int m)
{
    double** a = (double**)calloc(n, sizeof(double*));
    for (int i = 0; i < n; ++i) a[i] = (double*)calloc(m, sizeof(double));
    return a;
}
static void free2(double** a, int n) { if (!a) return; for (int i = 0; i < n; ++i) free(a[i]); free(a); }

orc_state* orc_create(const orc_problem* p)
{
    orc_state* s = (orc_state*)calloc(1, sizeof(orc_state));
    s->p = p; s->N = p->n_nodes; s->L = p->n_links; s->Z = p->n_zones; s->AT = p->n_agent_types; s->T = p->n_periods;
    const int N = s->N, L = s->L, Z = s->Z, AT = s->AT, T = s->T;
    s->row_ptr = (int**)calloc(AT * T, sizeof(int*));
    s->col_link = (int**)calloc(AT * T, sizeof(int*));
    s->col_to = (int**)calloc(AT * T, sizeof(int*));
    for (int at = 0; at < AT; ++at)
        for (int tau = 0; tau < T; ++tau) {
            int i = at * T + tau;
            s->row_ptr[i] = (int*)malloc(sizeof(int) * (N + 1));
            s->col_link[i] = (int*)malloc(sizeof(int) * (L > 0 ? L : 1));
            s->col_to[i] = (int*)malloc(sizeof(int) * (L > 0 ? L : 1));
            orc_forward_star(p, at, tau, s->row_ptr[i], s->col_link[i], s->col_to[i], NULL);
        }
    s->tt = alloc2(T, L); s->pce_vol = alloc2(T, L); s->person_vol = alloc2(T, L); s->link_volume = alloc2(T, L);
    s->doc = alloc2(T, L); s->voc = alloc2(T, L); s->P = alloc2(T, L); s->vt2 = alloc2(T, L);
    s->person_vol_at = alloc2(T * AT, L);
    s->gen_cost = alloc2(AT * T, L);
    if ((size_t)L * MAX_SPEED_SLOTS * sizeof(float) <= ((size_t)256 << 20))
        s->model_speed = (float*)calloc((size_t)L * MAX_SPEED_SLOTS, sizeof(float));

    /* g_assign_computing_tasks_to_memory_blocks, main_api.cpp:201-275 */
    const int B = p->memory_blocks;
    int cap_nets = AT * T * (Z < B ? Z : B);
    s->nets = (orc_net*)calloc(cap_nets > 0 ? cap_nets : 1, sizeof(orc_net));
    s->n_nets = 0;
    for (int at = 0; at < AT; ++at)
        for (int tau = 0; tau < T; ++tau) {
            int base = s->n_nets; /* PointerMatrx[at][tau][b] == nets[base + b] */
            for (int z = 0; z < Z; ++z) {
                if (z < B) {                                                              /* :220-239 */
                    orc_net* n = &s->nets[s->n_nets++];
                    n->at = at; n->tau = tau; n->n_origins = 0;
                    n->zones = (int*)malloc(sizeof(int) * (Z / (B > 0 ? B : 1) + 2));
                    n->zones[n->n_origins++] = z;
                } else if (p->origin_demand[z] > 0.001) {                                 /* :253-264 */
                    orc_net* n = &s->nets[base + z % B];
                    n->zones[n->n_origins++] = z;
                }
            }
        }
    s->n_blocks = s->n_nets < B ? s->n_nets : B;                                          /* main_api.cpp:645 */

    /* canonical task list: everything the iteration loop of main_api.cpp:667-697 actually visits */
    int cap_tasks = AT * T * Z;
    s->task_at = (int*)malloc(sizeof(int) * (cap_tasks + 1));
    s->task_tau = (int*)malloc(sizeof(int) * (cap_tasks + 1));
    s->task_zone = (int*)malloc(sizeof(int) * (cap_tasks + 1));
    unsigned char* visited = (unsigned char*)calloc((size_t)cap_tasks + 1, 1);
    for (int pid = 0; pid < s->n_blocks; ++pid)
        for (int blk = 0; blk < AT * T; ++blk) {
            int ni = blk * B + pid;
            if (ni >= s->n_nets) continue;
            orc_net* n = &s->nets[ni];
            for (int oi = 0; oi < n->n_origins; ++oi) visited[((size_t)n->at * T + n->tau) * Z + n->zones[oi]] = 1;
        }
    s->n_tasks = 0;
    for (int at = 0; at < AT; ++at)
        for (int tau = 0; tau < T; ++tau)
            for (int z = 0; z < Z; ++z)
                if (visited[((size_t)at * T + tau) * Z + z]) {
                    s->task_at[s->n_tasks] = at; s->task_tau[s->n_tasks] = tau; s->task_zone[s->n_tasks] = z;
                    s->n_tasks++;
                }
    free(visited);

    s->pool = (orc_colvec*)calloc(p->n_od > 0 ? p->n_od : 1, sizeof(orc_colvec));
    s->od_first = (int*)calloc(Z + 1, sizeof(int));
    for (int i = 0; i < p->n_od; ++i) s->od_first[p->od_o[i] + 1]++;
    for (int z = 0; z < Z; ++z) s->od_first[z + 1] += s->od_first[z];

    s->label = (double*)malloc(sizeof(double) * N);
    s->pred_node = (int*)malloc(sizeof(int) * N);
    s->pred_link = (int*)malloc(sizeof(int) * N);
    s->status = (int*)malloc(sizeof(int) * N);
    s->next = (int*)malloc(sizeof(int) * N);
    s->tmp_nodes = (int*)malloc(sizeof(int) * (N + 1));
    s->tmp_links = (int*)malloc(sizeof(int) * (N + 1));
    s->n_out_unfiltered = (int*)calloc(N, sizeof(int));
    for (int l = 0; l < L; ++l) s->n_out_unfiltered[p->link_from[l]]++;
    s->pm = *p;
    const size_t TL = (size_t)T * L;
    s->m_fftt = (double*)malloc(sizeof(double) * (TL + 1)); memcpy(s->m_fftt, p->fftt, sizeof(double) * TL);
    s->m_cap = (double*)malloc(sizeof(double) * (TL + 1)); memcpy(s->m_cap, p->cap, sizeof(double) * TL);
    s->m_nlanes = (double*)malloc(sizeof(double) * (TL + 1)); memcpy(s->m_nlanes, p->nlanes, sizeof(double) * TL);
    s->pm.fftt = s->m_fftt; s->pm.cap = s->m_cap; s->pm.nlanes = s->m_nlanes;
    s->rt_info = (int*)calloc(AT > 0 ? AT : 1, sizeof(int));
    s->sa_lanes = (double*)calloc(TL + 1, sizeof(double));
    s->design_flag = (signed char*)calloc(TL + 1, 1);
    s->od_impact = (unsigned char*)calloc((size_t)(p->n_od > 0 ? p->n_od : 1), 1);
    s->obs_count = (double*)calloc(TL + 1, sizeof(double));
    s->ub_flag = (int*)malloc(sizeof(int) * (TL + 1));
    for (size_t i = 0; i < TL; ++i) s->ub_flag[i] = 1;
    s->est_dev = (double*)calloc(TL + 1, sizeof(double));
    s->info_zone = (int*)calloc(Z > 0 ? Z : 1, sizeof(int));
    s->dms_zone = (int*)malloc(sizeof(int) * (TL + 1));
    s->dms_has_zone = (int*)calloc(TL + 1, sizeof(int));
    for (size_t i = 0; i < TL; ++i) s->dms_zone[i] = -1;
    s->at_pce = (float*)calloc(AT > 0 ? AT : 1, sizeof(float));
    s->at_occ = (float*)calloc(AT > 0 ? AT : 1, sizeof(float));
    for (int at = 0; at < AT; ++at) { s->at_pce[at] = 1; s->at_occ[at] = 1; }
    return s;
}

void orc_destroy(orc_state* s)
{
    if (!s) return;
    for (int i = 0; i < s->AT * s->T; ++i) { free(s->row_ptr[i]); free(s->col_link[i]); free(s->col_to[i]); }
    free(s->row_ptr); free(s->col_link); free(s->col_to);
    free2(s->tt, s->T); free2(s->pce_vol, s->T); free2(s->person_vol, s->T); free2(s->link_volume, s->T);
    free2(s->doc, s->T); free2(s->voc, s->T); free2(s->P, s->T); free2(s->vt2, s->T);
    free2(s->person_vol_at, s->T * s->AT); free2(s->gen_cost, s->AT * s->T);
    free(s->model_speed);
    for (int i = 0; i < s->n_nets; ++i) free(s->nets[i].zones);
    free(s->nets); free(s->task_at); free(s->task_tau); free(s->task_zone);
    for (int i = 0; i < s->p->n_od; ++i) {
        for (int j = 0; j < s->pool[i].n; ++j) free(s->pool[i].c[j].links);
        free(s->pool[i].c);
    }
    free(s->pool); free(s->od_first);
    free(s->label); free(s->pred_node); free(s->pred_link); free(s->status); free(s->next);
    free(s->tmp_nodes); free(s->tmp_links); free(s->n_out_unfiltered);
    free(s->m_fftt); free(s->m_cap); free(s->m_nlanes); free(s->rt_info); free(s->sa_lanes); free(s->design_flag); free(s->od_impact);
    free(s->info_zone); free(s->dms_zone); free(s->dms_has_zone);
    free(s->obs_count); free(s->ub_flag); free(s->est_dev); free(s->at_pce); free(s->at_occ);
    free(s);
}

int orc_num_tasks(const orc_state* s) { return s->n_tasks; }
void orc_get_tasks(const orc_state* s, int* at, int* tau, int* zone)
{
    memcpy(at, s->task_at, sizeof(int) * s->n_tasks);
    memcpy(tau, s->task_tau, sizeof(int) * s->n_tasks);
    memcpy(zone, s->task_zone, sizeof(int) * s->n_tasks);
}
void orc_set_tree_capture(orc_state* s, double* labels, int* pred_node, int* pred_link)
{
    s->cap_label = labels; s->cap_pn = pred_node; s->cap_pl = pred_link;
}

/* ------------------------------------------------------------------------------------------ */
/* update_link_travel_time_and_cost, assignment.h:197-259 + CLink::calculate_dynamic_VDFunction,
 * main_api.cpp:404-435 */
static double vdf_pass(orc_state* s)
{
    const orc_problem* p = &s->pm;
    for (int l = 0; l < s->L; ++l)
        for (int tau = 0; tau < s->T; ++tau) {
            double v = s->pce_vol[tau][l] + p->preload[(size_t)tau * s->L + l] + 0.0 /* sa_volume */; /* :413 */
            orc_vdf_out o;
            orc_vdf_link(p, tau, l, v, &o, s->model_speed ? s->model_speed + (size_t)l * MAX_SPEED_SLOTS : NULL);
            s->tt[tau][l] = o.tt; s->link_volume[tau][l] = v;
            s->doc[tau][l] = o.doc; s->voc[tau][l] = o.voc; s->P[tau][l] = o.P; s->vt2[tau][l] = o.vt2;
        }
    double total = 0;                                                                     /* :246-258 */
    for (int l = 0; l < s->L; ++l)
        for (int tau = 0; tau < s->T; ++tau) total += s->tt[tau][l] * s->link_volume[tau][l];
    return total;
}

/* g_reset_and_update_link_volume_based_on_columns, assignment.h:53-195 (stage-1 branch) */
static void scatter_pass_sa(orc_state* s, int iteration_index, int self_reducing, int sa);
static void scatter_pass(orc_state* s, int iteration_index, int self_reducing) { scatter_pass_sa(s, iteration_index, self_reducing, 0); }
static void scatter_pass_sa(orc_state* s, int iteration_index, int self_reducing, int sa)
{
    const orc_problem* p = s->p;
    const int L = s->L, AT = s->AT, T = s->T;
    for (int tau = 0; tau < T; ++tau) {                                                   /* :57-75 */
        memset(s->pce_vol[tau], 0, sizeof(double) * L);
        memset(s->person_vol[tau], 0, sizeof(double) * L);
        for (int at = 0; at < AT; ++at) memset(s->person_vol_at[tau * AT + at], 0, sizeof(double) * L);
    }
    if (iteration_index < 0) return;
    for (int at = 0; at < AT; ++at)                                                       /* :79 at, o, d, tau */
        for (int i = 0; i < p->n_od; ++i) {
            if (p->od_at[i] != at) continue;
            int tau = p->od_tau[i];
            orc_colvec* cv = &s->pool[i];
            const double* pce = p->pce + ((size_t)tau * AT + at) * L;
            const double* occ = p->occ + ((size_t)tau * AT + at) * L;
            for (int j = 0; j < cv->n; ++j) {
                orc_column* c = &cv->c[j];
                float f = (float)c->volume;                                               /* :87,127 */
                for (int nl = 0; nl < c->n_links; ++nl) {
                    int l = c->links[nl];
                    s->pce_vol[tau][l] += f * pce[l];                                     /* :140 */
                    s->person_vol[tau][l] += f * occ[l];                                  /* :141 */
                    s->person_vol_at[tau * AT + at][l] += f * occ[l];                     /* :142 */
                }
                /* :150-176 in the SA stage only users with real-time information are re-routed: a column of an agent type
                   without information keeps its volume, and so does a DMS user's whose origin is not an information zone
                   */
                if (sa && !(s->rt_info[at] == 1 || (s->rt_info[at] == 2 && s->info_zone[p->od_o[i]]))) continue;
                if (self_reducing)                                                        /* :182-186 */
                    c->volume = c->volume * ((double)iteration_index / (double)(iteration_index + 1));
            }
        }
}

static int find_od(const orc_state* s, int o, int d, int at, int tau)
{
    const orc_problem* p = s->p;
    int lo = s->od_first[o], hi = s->od_first[o + 1];
    while (lo < hi) { /* lower bound on (d, at, tau) */
        int mid = (lo + hi) / 2;
        int c = p->od_d[mid] - d;
        if (c == 0) c = p->od_at[mid] - at;
        if (c == 0) c = p->od_tau[mid] - tau;
        if (c < 0) lo = mid + 1; else hi = mid;
    }
    if (lo < s->od_first[o + 1] && p->od_d[lo] == d && p->od_at[lo] == at && p->od_tau[lo] == tau) return lo;
    return -1;
}

/* NetworkForSP::backtrace_shortest_path_tree, routing.h:413-629 (column-based mode) */
static double backtrace(orc_state* s, int at, int tau, int origin_zone, int k)
{
    const orc_problem* p = s->p;
    const int N = s->N, L = s->L;
    int origin_node = p->zone_centroid[origin_zone];
    double least = 0;
    if (s->n_out_unfiltered[origin_node] == 0) return 0;                                  /* :428 */
    float kp = (float)1 / (float)(k + 1);                                                 /* :435 */
    for (int i = 0; i < N; ++i) {                                                         /* :469 */
        if (p->node_zone[i] < 0) continue;                                                /* :471 zone_id >= 1 */
        if (i == origin_node) continue;                                                   /* :473 */
        int od = find_od(s, origin_zone, p->node_zone[i], at, tau);
        double ODvolume = od >= 0 ? p->od_vol[od] : 0.0;
        double volume = ODvolume * kp;                                                    /* :502 */
        if (!(ODvolume > 0.000001)) continue;                                             /* :505 */
        int ln = 0, ll = 0, node_sum = 0;
        int cur = i;
        if (s->pred_link[cur] >= 0) least += ODvolume * s->label[cur];                    /* :517-520 */
        while (cur >= 0 && cur < N) {                                                     /* :522-572 */
            s->tmp_nodes[ln++] = cur;
            int pl = s->pred_link[cur];
            node_sum = (int)((unsigned)node_sum + (unsigned)cur + (unsigned)pl);          /* :536 int wraparound */
            if (pl >= 0 && pl < L) s->tmp_links[ll++] = pl;
            cur = s->pred_node[cur];
        }
        if (ln < 2) continue;                                                             /* :584 */
        orc_colvec* cv = &s->pool[od];
        int pos = 0;
        while (pos < cv->n && cv->c[pos].node_sum < node_sum) ++pos;
        if (pos == cv->n || cv->c[pos].node_sum != node_sum) {                            /* :595-617 insert */
            if (cv->n == cv->cap) {
                cv->cap = cv->cap ? cv->cap * 2 : 4;
                cv->c = (orc_column*)realloc(cv->c, sizeof(orc_column) * cv->cap);
            }
            memmove(&cv->c[pos + 1], &cv->c[pos], sizeof(orc_column) * (cv->n - pos));
            orc_column* c = &cv->c[pos];
            memset(c, 0, sizeof(*c));
            c->node_sum = node_sum; c->seq_no = cv->n; c->volume = 0; c->toll = s->label[i];
            c->n_links = ll; c->n_nodes = ln;
            c->links = (int*)malloc(sizeof(int) * (ll > 0 ? ll : 1));
            for (int q = 0; q < ll; ++q) c->links[q] = s->tmp_links[ll - 1 - q];          /* DTA.h:551-559 */
            cv->n++;
        }
        cv->c[pos].volume += volume;                                                      /* :619 */
    }
    return least;
}

static int task_index(const orc_state* s, int at, int tau, int zone)
{
    for (int lo = 0, hi = s->n_tasks; lo < hi;) {
        int mid = (lo + hi) / 2;
        int c = s->task_at[mid] - at;
        if (c == 0) c = s->task_tau[mid] - tau;
        if (c == 0) c = s->task_zone[mid] - zone;
        if (c == 0) return mid;
        if (c < 0) lo = mid + 1; else hi = mid;
    }
    return -1;
}

/* the shortest-path region of an iteration: main_api.cpp:667-697 (stage 1) and :966-1018 (SA stage: only the origins of agent
 * types with real-time information are searched and back-traced, :987-994) */
static double sp_region(orc_state* s, int k, int sa, long long* counted_out)
{
    const orc_problem* p = s->p;
    const int N = s->N, L = s->L, AT = s->AT, T = s->T, B = p->memory_blocks;
    double least_total = 0;
    long long counted_edges = 0;
    for (int pid = 0; pid < s->n_blocks; ++pid)                                           /* :667 */
        for (int blk = 0; blk < AT * T; ++blk) {
            int ni = blk * B + pid;                                                       /* :671 */
            if (ni >= s->n_nets) continue;
            orc_net* n = &s->nets[ni];
            int fs = n->at * T + n->tau;
            double* cost = s->gen_cost[fs];
            float vot = p->vot[n->at];
            for (int oi = 0; oi < n->n_origins; ++oi) {
                int oz = n->zones[oi];
                if (sa && !(s->rt_info[n->at] == 1 || (s->rt_info[n->at] == 2 && s->info_zone[oz]))) continue; /* main_api.cpp:987-994 */
                /* UpdateGeneralizedLinkCost, routing.h:229-231 */
                for (int l = 0; l < L; ++l) {
                    size_t a = ((size_t)n->tau * AT + n->at) * L + l;
                    cost[l] = s->tt[n->tau][l] + p->penalty[(size_t)n->tau * L + l] + p->lr_price[a] + p->toll[a] / vot * 60;
                }
                sssp_core(N, s->row_ptr[fs], s->col_link[fs], s->col_to[fs], cost, p->link_tag, p->zone_centroid[oz], oz,
                          s->label, s->pred_node, s->pred_link, s->status, s->next, NULL);
                for (int i = 0; i < N; ++i) {
                    if (!(s->label[i] < ORC_MAX_LABEL_COST)) continue;
                    for (int e = s->row_ptr[fs][i]; e < s->row_ptr[fs][i + 1]; ++e) {
                        int tg = p->link_tag[s->col_link[fs][e]];
                        if (tg >= 0 && tg != oz) continue;
                        ++counted_edges;
                    }
                }
                if (s->cap_label) {
                    int ti = task_index(s, n->at, n->tau, oz);
                    memcpy(s->cap_label + (size_t)ti * N, s->label, sizeof(double) * N);
                    memcpy(s->cap_pn + (size_t)ti * N, s->pred_node, sizeof(int) * N);
                    memcpy(s->cap_pl + (size_t)ti * N, s->pred_link, sizeof(int) * N);
                }
                least_total += backtrace(s, n->at, n->tau, oz, k);                        /* :687-693 */
            }
        }
    if (counted_out) *counted_out = counted_edges;
    return least_total;
}

/* one column-generation iteration, main_api.cpp:589-728 */
void orc_iteration(orc_state* s, int k, double row[5])
{
    const orc_problem* p = s->p;
    double sysTT = vdf_pass(s);                                                           /* :606 */
    scatter_pass(s, k, 1);                                                                /* :618 */
    long long counted_edges = 0;
    double least_total = sp_region(s, k, 0, &counted_edges);
    double gap = (sysTT - least_total) / dmax(0.00001, least_total);                      /* :711 */
    float denom = 1 > p->total_demand_volume ? (float)1 : p->total_demand_volume;         /* :716 max(1, float) */
    row[0] = sysTT; row[1] = least_total; row[2] = gap; row[3] = sysTT / denom; row[4] = (double)counted_edges;
}

/* g_update_gradient_cost_and_assigned_flow_in_column_pool, assignment.h:565-928 (stage 2, no SA) */
static void column_update_impl(orc_state* s, int n, double row[4], int sa)
{
    const orc_problem* p = s->p;
    const int L = s->L, AT = s->AT;
    double total_gap = 0, total_cost = 0, total_time = 0, total_demand = 0;
    scatter_pass_sa(s, n, 0, sa);                                                         /* :575 */
    vdf_pass(s);                                                                          /* :593 */
    for (int i = 0; i < p->n_od; ++i) {                                                   /* o, d, at, tau */
        orc_colvec* cv = &s->pool[i];
        int at = p->od_at[i], tau = p->od_tau[i];
        double od_volume = p->od_vol[i];
        float vot = p->vot[at];
        double least_cost = 999999;                                                       /* :681 */
        int least_seq = -1, least_pos = -1;
        for (int j = 0; j < cv->n; ++j) {                                                 /* :690-767 */
            orc_column* c = &cv->c[j];
            double path_toll = 0, path_cost = 0, path_time = 0;
            if (sa && s->rt_info[at] == 1)                                                /* :699-713 */
                for (int q = 0; q < c->n_links; ++q)
                    if (s->design_flag[(size_t)tau * L + c->links[q]] == -1) { s->od_impact[i] = 1; break; }
            for (int q = 0; q < c->n_links; ++q) {
                int l = c->links[q];
                size_t a = ((size_t)tau * AT + at) * L + l;
                path_toll += p->toll[a];
                path_time += s->tt[tau][l];
                path_cost += s->tt[tau][l] + p->penalty[(size_t)tau * L + l] + p->toll[a] / vot * 60; /* DTA.h:1157 */
            }
            c->toll = path_toll; c->time = path_time; c->cost = path_cost;
            total_time += c->time * c->volume;                                            /* :746-747 */
            total_demand += c->volume;
            if (cv->n == 1) total_cost += c->cost * c->volume;                            /* :750-754 */
            if (path_cost < least_cost) { least_cost = path_cost; least_seq = c->seq_no; least_pos = j; } /* :757 */
        }
        if (cv->n >= 2) {                                                                 /* :769 */
            double switched = 0;
            for (int j = 0; j < cv->n; ++j) {
                orc_column* c = &cv->c[j];
                if (c->seq_no == least_seq) continue;                                     /* :775 */
                double diff = c->cost - least_cost;
                double rel = diff / dmax(0.0001, least_cost);                             /* :781 */
                total_gap += diff * c->volume;                                            /* :786-787 */
                total_cost += c->cost * c->volume;
                double step = 1.0 / (n + 2) * od_volume;                                  /* :802 */
                double prev = c->volume;
                double shift = step * dmax(0.0000, rel);                                  /* :808 */
                if (rel > 3 * 60) shift = c->volume;                                      /* :811 */
                if (shift > c->volume) shift = c->volume;                                 /* :816 */
                if (sa && !s->od_impact[i]) shift = 0;                                    /* :828-846 */
                c->volume = dmax(0.0, c->volume - shift);                                 /* :851 */
                switched += (prev - c->volume);                                           /* :859 */
            }
            if (least_seq != -1 && least_pos >= 0) {                                      /* :869-886 */
                cv->c[least_pos].volume += switched;
                total_cost += cv->c[least_pos].cost * cv->c[least_pos].volume;
            }
        }
    }
    row[0] = total_time / dmax(0.001, total_demand);                                      /* :907 */
    row[1] = total_gap;
    row[2] = total_gap * 100.0 / dmax(0.00001, total_cost);                               /* :911 */
    row[3] = total_demand;
}

void orc_column_update(orc_state* s, int n, double row[4]) { column_update_impl(s, n, row, 0); }

/* ---- sensitivity-analysis stage, main_api.cpp:771-1042 (supply-side `sa` lane changes and re-routing of the agent types
 * with real-time information; dms scenario rows / information zones / demand-side factors are not covered) ---- */
void orc_sa_configure(orc_state* s, const int* rt_info) { memcpy(s->rt_info, rt_info, sizeof(int) * s->AT); }

/* a supply_side_scenario.csv row of type "sa", scenario_API.h:152-160 */
void orc_sa_lane_change(orc_state* s, int tau, int link, double lanes)
{
    s->sa_lanes[(size_t)tau * s->L + link] = (double)(float)lanes;
    s->design_flag[(size_t)tau * s->L + link] = -1;
}

/* main_api.cpp:779-794 (the lane changes take effect) and :857 (VDF pass); returns the system travel time */
double orc_sa_begin(orc_state* s)
{
    for (int l = 0; l < s->L; ++l)
        for (int tau = 0; tau < s->T; ++tau) {
            size_t a = (size_t)tau * s->L + l;
            if (s->design_flag[a] != -1) continue;
            double old_lanes = s->m_nlanes[a];                                            /* :786 float old_lanes */
            float old_f = (float)old_lanes;
            s->m_nlanes[a] = s->sa_lanes[a];                                              /* :787 */
            s->m_cap[a] *= s->m_nlanes[a] / dmax(0.00001, (double)old_f);                 /* :788 */
            if (s->m_nlanes[a] < 0.01) s->m_fftt[a] = 1440;                               /* :789-790 */
        }
    return vdf_pass(s);
}

/* one column re-generation iteration, main_api.cpp:864-1034.  row = { sysTT, avg travel time, least_total } */
void orc_sa_iteration(orc_state* s, int it, double row[3])
{
    const orc_problem* p = s->p;
    /* induced delay, main_api.cpp:877-891: links without a scenario whose degree of congestion of the last VDF pass exceeds 3 get
       network_design_flag -2; the flag only matters where the code asks for "<= -1" (g_column_pool_route_modification) */
    for (int tau = 0; tau < s->T; ++tau)
        for (int l = 0; l < s->L; ++l)
            if (s->design_flag[(size_t)tau * s->L + l] == 0 && s->doc[tau][l] > 3) s->design_flag[(size_t)tau * s->L + l] = -2;
    double sysTT = vdf_pass(s);                                                           /* :897 */
    scatter_pass_sa(s, it, 1, 1);                                                         /* :913 */
    double least_total = sp_region(s, it, 1, NULL);                                       /* :966-1018 */
    float denom = 1 > p->total_demand_volume ? (float)1 : p->total_demand_volume;         /* :1027 */
    row[0] = sysTT; row[1] = sysTT / denom; row[2] = least_total;
}

/* a supply_side_scenario.csv row of type "dms", scenario_API.h:165-188: has_zone = the to-node of the link carries a zone id,
 * zone = its zone seq no (-1 when that id is not a zone of the model) */
void orc_sa_dms_link(orc_state* s, int tau, int link, int has_zone, int zone)
{
    size_t a = (size_t)tau * s->L + link;
    s->design_flag[a] = 2;
    s->dms_has_zone[a] = has_zone;
    s->dms_zone[a] = zone;
    if (zone >= 0) s->info_zone[zone] = 1;
}

static orc_column* colvec_insert(orc_colvec* cv, int node_sum, int* existed)
{
    int pos = 0;
    while (pos < cv->n && cv->c[pos].node_sum < node_sum) ++pos;
    if (pos < cv->n && cv->c[pos].node_sum == node_sum) { *existed = 1; return &cv->c[pos]; }
    if (cv->n == cv->cap) {
        cv->cap = cv->cap ? cv->cap * 2 : 4;
        cv->c = (orc_column*)realloc(cv->c, sizeof(orc_column) * cv->cap);
    }
    memmove(&cv->c[pos + 1], &cv->c[pos], sizeof(orc_column) * (cv->n - pos));
    memset(&cv->c[pos], 0, sizeof(orc_column));
    cv->c[pos].node_sum = node_sum;
    cv->n++;
    *existed = 0;
    return &cv->c[pos];
}

/* g_column_pool_route_modification, assignment.h:1077-1383 (one thread: the travel times of the information-zone columns are
 * accumulated in visiting order, :1219).  Returns the number of columns that were split. */
int orc_sa_route_modification(orc_state* s)
{
    const orc_problem* p = s->p;
    const int L = s->L, Z = s->Z, AT = s->AT, T = s->T;
    int n_split = 0;
    int* seq = (int*)malloc(sizeof(int) * (size_t)(2 * s->N + 8));
    for (int o = 0; o < Z; ++o)
        for (int d = 0; d < Z; ++d)
            for (int at = 0; at < AT; ++at) {
                if (s->rt_info[at] != 2) continue;                                        /* :1110 */
                for (int tau = 0; tau < T; ++tau) {
                    int cell = find_od(s, o, d, at, tau);
                    if (cell < 0) continue;                                               /* od_volume > 0, :1115 */
                    orc_colvec* cv = &s->pool[cell];
                    const signed char* flag = s->design_flag + (size_t)tau * L;
                    /* the std::map is walked in key order while g_add_new_column inserts into it: a new column with a larger
                       key is visited later in the same walk */
                    int have_key = 0, last_key = 0;
                    for (;;) {
                        int j = -1;
                        for (int q = 0; q < cv->n; ++q)
                            if (!have_key || cv->c[q].node_sum > last_key) { j = q; break; }
                        if (j < 0) break;
                        have_key = 1; last_key = cv->c[j].node_sum;
                        if (cv->c[j].volume < 0.00001) continue;                          /* :1136 */
                        int passing_info = 0, new_orig = -1, passing_cap = 0, n_seq = 0;
                        for (int nl = 0; nl < cv->c[j].n_links; ++nl) {                   /* :1141 */
                            int l = cv->c[j].links[nl];
                            size_t a = (size_t)tau * L + l;
                            if (!passing_info && flag[l] == 2 && s->dms_has_zone[a]) {    /* :1146-1150 */
                                s->od_impact[cell] = 1;                                   /* :1153-1154 */
                                if (s->dms_zone[a] < 0) continue;                         /* :1158-1159 */
                                passing_info = 1; new_orig = s->dms_zone[a];
                                n_seq = 0;
                                for (int nl2 = 0; nl2 <= nl; ++nl2) seq[n_seq++] = cv->c[j].links[nl2]; /* :1166-1170 */
                            }
                            if (flag[l] == -1) { passing_cap = 1; cv->c[j].impacted = 1; } /* :1176-1180 */
                        }
                        if (!(passing_cap && passing_info)) continue;                     /* :1184 */
                        int cell2 = find_od(s, new_orig, d, at, tau);
                        if (cell2 < 0 || s->pool[cell2].n == 0) { free(seq); return -1; } /* :1203-1208 the reference stops here */
                        orc_colvec* cv2 = &s->pool[cell2];
                        float non_div = 999999;                                           /* :1214 */
                        for (int q = 0; q < cv2->n; ++q) {                                /* :1215-1236 */
                            orc_column* c2 = &cv2->c[q];
                            for (int nl2 = 1; nl2 < c2->n_links; ++nl2) c2->time += s->tt[tau][c2->links[nl2]]; /* accumulates, :1219 */
                            int div = 1;
                            for (int nl2 = 1; nl2 < c2->n_links; ++nl2)
                                if (flag[c2->links[nl2]] <= -1) div = 0;
                            if (!div) {
                                float abs_br = 3, ratio_br = 0.15;
                                if (c2->time < non_div - abs_br && c2->time < (non_div * (1 - ratio_br))) non_div = c2->time;
                            }
                        }
                        int found = 0;
                        for (int q = 0; q < cv2->n; ++q) {                                /* :1238-1262 */
                            orc_column* c2 = &cv2->c[q];
                            int div = 1;
                            for (int nl2 = 1; nl2 < c2->n_links; ++nl2)
                                if (flag[c2->links[nl2]] <= -1) div = 0;
                            if (div && !found && c2->time < non_div) {
                                for (int nl2 = 1; nl2 < c2->n_links; ++nl2) seq[n_seq++] = c2->links[nl2];
                                found = 1;
                            }
                        }
                        if (found) {                                                      /* :1264, optional detour :1298-1325 */
                            float base_tt = 0, alt_tt = 0;
                            for (int nl0 = 0; nl0 < cv->c[j].n_links; ++nl0) base_tt += s->tt[tau][cv->c[j].links[nl0]];
                            for (int nl3 = 1; nl3 < n_seq; ++nl3) alt_tt += s->tt[tau][seq[nl3]];
                            float beta = -0.1;
                            float prob_base = expf(beta * base_tt) / (expf(beta * base_tt) + expf(beta * alt_tt));
                            float base_volume = cv->c[j].volume;
                            cv->c[j].volume = base_volume * prob_base;
                            /* g_add_new_column, assignment.h:1012-1045 */
                            int node_sum = 0;
                            for (int q = 0; q < n_seq; ++q) node_sum = (int)((unsigned)node_sum + (unsigned)seq[q] + (unsigned)p->link_to[seq[q]]);
                            int size_before = cv->n, existed = 0;
                            orc_column* nc = colvec_insert(cv, node_sum, &existed);
                            (void)existed;
                            nc->seq_no = size_before; /* "map[key].path_seq_no = map.size()", assignment.h:1042: the right operand is evaluated
                                                         first (C++17), i.e. BEFORE operator[] creates the entry -- two columns of a cell can
                                                         end up with the same path_seq_no, and the column-pool update then treats both as
                                                         "the least-cost path" (assignment.h:775) */
                            free(nc->links);
                            nc->links = (int*)malloc(sizeof(int) * (size_t)(n_seq > 0 ? n_seq : 1));
                            memcpy(nc->links, seq, sizeof(int) * (size_t)n_seq);
                            nc->n_links = n_seq; nc->n_nodes = n_seq + 1;
                            nc->volume = base_volume * (1 - prob_base);
                            nc->impacted = 3;                                             /* :1322 */
                            ++n_split;
                        }
                        cv2->info_type = 1;                                               /* :1360 */
                    }
                }
            }
    free(seq);
    return n_split;
}

/* g_update_gradient_cost_and_assigned_flow_in_column_pool with the SA flag (main_api.cpp:1040, assignment.h:565-928) */
void orc_sa_column_update(orc_state* s, int n, double row[4]) { column_update_impl(s, n, row, 1); }

/* ---- ODME: Assignment::Demand_ODME, ODME.h:91-516 ---- */
void orc_odme_configure(orc_state* s, const float* at_pce, const float* at_occ)
{
    memcpy(s->at_pce, at_pce, sizeof(float) * s->AT);
    memcpy(s->at_occ, at_occ, sizeof(float) * s->AT);
}

/* one "link" row of measurement.csv for one period, ODME.h:140-178: an actual count overwrites what is there, an upper bound
 * does not overwrite an existing record */
void orc_odme_sensor(orc_state* s, int tau, int link, float count, int upper_bound_flag)
{
    size_t a = (size_t)tau * s->L + link;
    if (s->obs_count[a] >= 1) {
        if (upper_bound_flag == 0) { s->obs_count[a] = count; s->ub_flag[a] = upper_bound_flag; }
    } else {
        s->obs_count[a] = count; s->ub_flag[a] = upper_bound_flag;
    }
}

/* g_reset_and_update_link_volume_based_on_ODME_columns, assignment.h:263-563 (one thread: the order of the sums is the
 * loop order).  row = { gap (link MAPE / 100), system_gap, link MAE, avg_tt, UE gap per unit of demand, UE gap %, demand, links with data } */
void orc_odme_scatter(orc_state* s, int iteration_no, double row[8])
{
    const orc_problem* p = s->p;
    const int L = s->L, AT = s->AT, T = s->T;
    float total_gap = 0, sub_total_gap_link_count = 0, sub_total_system_gap_count = 0;           /* :265-268 */
    double total_cost = 0, total_time = 0, total_demand = 0, total_ue_gap = 0;
    (void)iteration_no; (void)total_cost;
    for (int tau = 0; tau < T; ++tau) {                                                   /* :281-293 */
        memset(s->pce_vol[tau], 0, sizeof(double) * L);
        memset(s->person_vol[tau], 0, sizeof(double) * L);
        for (int at = 0; at < AT; ++at) memset(s->person_vol_at[tau * AT + at], 0, sizeof(double) * L);
    }
    for (int at = 0; at < AT; ++at) {                                                     /* :302 at, o, d, tau */
        float PCE_ratio = s->at_pce[at], OCC_ratio = s->at_occ[at];                       /* :308-309 */
        for (int i = 0; i < p->n_od; ++i) {
            if (p->od_at[i] != at) continue;
            int tau = p->od_tau[i];
            orc_colvec* cv = &s->pool[i];
            double least_cost = 999999;                                                   /* :347 */
            int least_seq = -1;
            for (int j = 0; j < cv->n; ++j) {                                             /* :370-409 */
                orc_column* c = &cv->c[j];
                total_demand += c->volume;
                double path_time = 0;
                for (int q = 0; q < c->n_links; ++q) path_time += s->tt[tau][c->links[q]];
                c->toll = 0; c->time = path_time;
                total_time += c->time * c->volume;
                if (cv->n == 1) break;                                                    /* :393-396 */
                if (path_time < least_cost) { least_cost = path_time; least_seq = c->seq_no; }
                total_cost += c->time * c->volume;
            }
            if (cv->n >= 2)                                                               /* :411-430 */
                for (int j = 0; j < cv->n; ++j) {
                    orc_column* c = &cv->c[j];
                    if (c->seq_no != least_seq) total_ue_gap += (c->time - least_cost) * c->volume;
                }
            for (int j = 0; j < cv->n; ++j) {                                             /* :433-457 */
                orc_column* c = &cv->c[j];
                float f = (float)c->volume;
                for (int q = 0; q < c->n_links; ++q) {
                    int l = c->links[q];
                    s->pce_vol[tau][l] += f * PCE_ratio;                                  /* float product, :449 */
                    s->person_vol[tau][l] += f * OCC_ratio;
                    s->person_vol_at[tau * AT + at][l] += f;                              /* :451 pure volume */
                }
            }
        }
    }
    int total_link_count = 0;
    vdf_pass(s);                                                                          /* :467 per link, same function */
    for (int l = 0; l < L; ++l)                                                           /* :465-515 */
        for (int tau = 0; tau < T; ++tau) {
            size_t a = (size_t)tau * L + l;
            if (!(s->obs_count[a] >= 1)) continue;
            double dev = s->pce_vol[tau][l] + s->pm.preload[a] - s->obs_count[a];         /* :476 */
            s->est_dev[a] = dev;
            if (s->ub_flag[a] == 0 || dev > 0) {                                          /* :489-507 */
                total_gap += abs((int)dev);                                               /* :491 abs() binds to the int overload here */
                sub_total_gap_link_count += fabs(dev / s->obs_count[a]);
                sub_total_system_gap_count += dev / s->obs_count[a];
            }
            total_link_count += 1;
        }
    int denom = total_link_count > 1 ? total_link_count : 1;
    row[0] = sub_total_gap_link_count / denom;                                            /* :560 */
    row[1] = sub_total_system_gap_count / denom;                                          /* :561 */
    row[2] = total_gap / denom;
    row[3] = total_time / dmax(0.1, total_demand);
    row[4] = total_ue_gap / dmax(0.00001, total_demand);
    row[5] = total_ue_gap / dmax(0.00001, total_time) * 100;
    row[6] = total_demand;
    row[7] = total_link_count;
}

/* the path-flow step of one ODME iteration, ODME.h:262-488 */
void orc_odme_update(orc_state* s)
{
    const orc_problem* p = s->p;
    const int L = s->L;
    for (int i = 0; i < p->n_od; ++i) {
        orc_colvec* cv = &s->pool[i];
        int tau = p->od_tau[i];
        for (int j = 0; j < cv->n; ++j) {                                                 /* :349-470 */
            orc_column* c = &cv->c[j];
            float path_gradient_cost = 0;
            for (int q = 0; q < c->n_links; ++q) {                                        /* :396-421 */
                size_t a = (size_t)tau * L + c->links[q];
                if (s->obs_count[a] >= 1) {
                    if (s->ub_flag[a] == 0) path_gradient_cost += s->est_dev[a];
                    if (s->ub_flag[a] == 1 && s->est_dev[a] > 0) path_gradient_cost += s->est_dev[a];
                }
            }
            c->cost = path_gradient_cost;                                                 /* :426 */
            float step_size = 0.05;                                                       /* :428 */
            double weight = 1;
            double ue_gap = 0; /* weight 0: never contributes */
            double change = step_size * (weight * c->cost + (1 - weight) * ue_gap);       /* :434 */
            float lo = c->volume * 0.1 * (-1), hi = c->volume * 0.1;                      /* :439-440 */
            if (change < lo) change = lo;
            if (change > hi) change = hi;
            c->volume = dmax(1.0, c->volume - change);                                    /* :449 */
        }
    }
}

double orc_finalize(orc_state* s, int K)
{
    scatter_pass(s, K, 0);                                                                /* main_api.cpp:752 */
    return vdf_pass(s);                                                                   /* :759 */
}

void orc_get_link_state(const orc_state* s, int tau, double* tt, double* pce_volume, double* person_volume,
                        double* link_volume, double* doc, double* voc, double* P, double* vt2)
{
    size_t n = sizeof(double) * s->L;
    if (tt) memcpy(tt, s->tt[tau], n);
    if (pce_volume) memcpy(pce_volume, s->pce_vol[tau], n);
    if (person_volume) memcpy(person_volume, s->person_vol[tau], n);
    if (link_volume) memcpy(link_volume, s->link_volume[tau], n);
    if (doc) memcpy(doc, s->doc[tau], n);
    if (voc) memcpy(voc, s->voc[tau], n);
    if (P) memcpy(P, s->P[tau], n);
    if (vt2) memcpy(vt2, s->vt2[tau], n);
}
void orc_get_model_speed(const orc_state* s, int link, float* speed300)
{
    if (s->model_speed) memcpy(speed300, s->model_speed + (size_t)link * MAX_SPEED_SLOTS, sizeof(float) * MAX_SPEED_SLOTS);
    else memset(speed300, 0, sizeof(float) * MAX_SPEED_SLOTS);
}
void orc_get_gen_cost(const orc_state* s, int at, int tau, double* cost)
{
    memcpy(cost, s->gen_cost[at * s->T + tau], sizeof(double) * s->L);
}

long long orc_num_columns(const orc_state* s)
{
    long long n = 0;
    for (int i = 0; i < s->p->n_od; ++i) n += s->pool[i].n;
    return n;
}
long long orc_num_column_links(const orc_state* s)
{
    long long n = 0;
    for (int i = 0; i < s->p->n_od; ++i)
        for (int j = 0; j < s->pool[i].n; ++j) n += s->pool[i].c[j].n_links;
    return n;
}
void orc_get_columns(const orc_state* s, int* od_index, int* node_sum, int* seq_no, int* n_links, int* n_nodes,
                     double* volume, double* cost, double* time, int* links)
{
    long long c = 0, lp = 0;
    for (int i = 0; i < s->p->n_od; ++i)
        for (int j = 0; j < s->pool[i].n; ++j) {
            const orc_column* col = &s->pool[i].c[j];
            od_index[c] = i; node_sum[c] = col->node_sum; seq_no[c] = col->seq_no;
            n_links[c] = col->n_links; n_nodes[c] = col->n_nodes;
            volume[c] = col->volume; cost[c] = col->cost; time[c] = col->time;
            memcpy(links + lp, col->links, sizeof(int) * col->n_links);
            lp += col->n_links; ++c;
        }
}
