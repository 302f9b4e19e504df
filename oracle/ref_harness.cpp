/* TEST INFRASTRUCTURE ONLY -- never linked into, imported by, or shipped with the product.
 *
 * Harness around the UNMODIFIED reference engine (asu-trans-ai-lab/DLSim-MRM, src/cpp).
 * The reference is one translation unit in practice (main_api.cpp #includes every header
 * with its function bodies, main_api.cpp:94-126), so this file #includes main_api.cpp and
 * adds its own main().  It is compiled by oracle/build_ref.sh into oracle/_ref/dtalite_ref
 * (git-ignored), and is used
 *   (1) to pin oracle/dtalite_oracle.c and the CUDA path: every hot-path state the reference
 *       holds in public members/globals is dumped to a tagged binary file, and
 *   (2) as the timed "reference" CPU arm of bench.py (phase wall-clock timers).
 *
 * Modes (CWD must hold the GMNS input files; the reference opens its log files in CWD at
 * static-initialisation time, DTA.h:858-876):
 *   dtalite_ref golden K CU B      -> calls network_assignment(1,K,CU,0,0,0,B) verbatim
 *                                      (main_api.cpp:499) and exits.
 *   dtalite_ref golden_dump K CU B out S [ODME] -> network_assignment(1,K,CU,0,S,0,B) verbatim (S sensitivity-analysis iterations:
 *                                      supply_side_scenario.csv lane changes, re-routing of the agent types with real-time
 *                                      information, main_api.cpp:771-1042), then dumps the state it leaves in the globals.
 *   dtalite_ref trace  K CU B out  -> re-drives stage 1 / stage 2 of network_assignment
 *                                      (main_api.cpp:505-759) by calling the reference's own
 *                                      functions in the same order, dumping state to <out>.
 * Environment knobs for trace mode:
 *   DTL_TRACE_TREES=0,1,2    iterations whose per-origin label/predecessor arrays are dumped
 *   DTL_TRACE_ORIGIN_STRIDE  dump every n-th origin only (default 1)
 *   DTL_TRACE_LIGHT=1        skip all per-iteration link-state / cost dumps (timing runs; tree dumps of
 *                            DTL_TRACE_TREES are still written, after the iteration's timed region)
 *   DTL_TRACE_NO_COLUMNS=1   skip the final dump of the column pool (timing runs: 1.4 GB on the config-5 grid)
 *   DTL_TRACE_BACKWARD=n     also dump the backward tree (optimal_backward_label_correcting_from_destination) of every n-th
 *                            zone centroid on the final travel times, per agent type and period
 *   DTL_TRACE_EDGE_ITERS=0   iterations that also run the Graph500-style edge count (default: iteration 0 only;
 *                            the count is a diagnostic walk over every tree and is kept out of timed iterations)
 */
#include "main_api.cpp"

#include <chrono>
#include <unistd.h>
#include <cstdint>
#include <set>

namespace {

FILE* g_dump = nullptr;

enum { DT_I32 = 0, DT_F64 = 1, DT_F32 = 2, DT_I64 = 3 };

void dump_rec(const std::string& name, int dtype, const void* data, uint64_t count)
{
    if (!g_dump) return;
    static const int esz[4] = { 4, 8, 4, 8 };
    uint32_t len = (uint32_t)name.size();
    uint8_t dt = (uint8_t)dtype;
    fwrite(&len, 4, 1, g_dump);
    fwrite(name.data(), 1, len, g_dump);
    fwrite(&dt, 1, 1, g_dump);
    fwrite(&count, 8, 1, g_dump);
    if (count) fwrite(data, esz[dtype], count, g_dump);
}
void dump_i32(const std::string& n, const std::vector<int>& v) { dump_rec(n, DT_I32, v.data(), v.size()); }
void dump_f64(const std::string& n, const std::vector<double>& v) { dump_rec(n, DT_F64, v.data(), v.size()); }

double now_s()
{
    using namespace std::chrono;
    return duration<double>(steady_clock::now().time_since_epoch()).count();
}

std::set<int> parse_int_set(const char* s)
{
    std::set<int> out;
    if (!s) return out;
    std::stringstream ss(s);
    std::string tok;
    while (std::getline(ss, tok, ','))
        if (!tok.empty()) out.insert(atoi(tok.c_str()));
    return out;
}

/* static image of the network exactly as the reference holds it after reading */
void dump_network()
{
    const int N = (int)g_node_vector.size();
    const int L = (int)g_link_vector.size();
    const int Z = (int)g_zone_vector.size();
    const int AT = (int)assignment.g_AgentTypeVector.size();
    const int T = (int)assignment.g_DemandPeriodVector.size();
    std::vector<int> meta = { N, L, Z, AT, T };
    dump_i32("meta", meta);

    std::vector<int> node_id(N), node_zone(N), node_zone_org(N);
    for (int i = 0; i < N; ++i) {
        node_id[i] = g_node_vector[i].node_id;
        node_zone[i] = g_node_vector[i].zone_id;
        node_zone_org[i] = g_node_vector[i].zone_org_id;
    }
    dump_i32("node_id", node_id);
    dump_i32("node_zone_id", node_zone);
    dump_i32("node_zone_org_id", node_zone_org);

    std::vector<int> zone_id(Z), zone_node(Z);
    for (int z = 0; z < Z; ++z) { zone_id[z] = g_zone_vector[z].zone_id; zone_node[z] = g_zone_vector[z].node_seq_no; }
    dump_i32("zone_id", zone_id);
    dump_i32("zone_node", zone_node);

    std::vector<int> from(L), to(L), tag(L), ltype(L), vdft(L);
    std::vector<double> length_m(L), lanes(L);
    for (int l = 0; l < L; ++l) {
        const CLink& k = g_link_vector[l];
        from[l] = k.from_node_seq_no; to[l] = k.to_node_seq_no;
        tag[l] = k.zone_seq_no_for_outgoing_connector; ltype[l] = k.link_type;
        vdft[l] = (int)k.vdf_type; length_m[l] = k.length_in_meter; lanes[l] = k.number_of_lanes;
    }
    dump_i32("link_from", from); dump_i32("link_to", to); dump_i32("link_tag", tag);
    dump_i32("link_type", ltype); dump_i32("link_vdf_type", vdft);
    dump_f64("link_length_m", length_m); dump_f64("link_lanes", lanes);

    for (int tau = 0; tau < T; ++tau) {
        std::string s = ":tau=" + std::to_string(tau);
        std::vector<double> fftt(L), cap(L), alpha(L), beta(L), vf(L), vc(L), nl(L), capl(L), qdf(L),
            preload(L), penalty(L), cyc(L), red(L), sat(L), qa(L), qb(L), qcd(L), qn(L), qcp(L), qs(L),
            t2(L), Lh(L), sth(L), enh(L), plf(L), tt0(L);
        std::vector<int> pvdf(L);
        for (int l = 0; l < L; ++l) {
            const CPeriod_VDF& v = g_link_vector[l].VDF_period[tau];
            fftt[l] = v.FFTT; cap[l] = v.BPR_period_capacity; alpha[l] = v.alpha; beta[l] = v.beta;
            vf[l] = v.vf; vc[l] = v.v_congestion_cutoff; nl[l] = v.nlanes; capl[l] = v.lane_based_ultimate_hourly_capacity;
            qdf[l] = v.queue_demand_factor; preload[l] = v.preload; penalty[l] = v.penalty;
            cyc[l] = v.cycle_length; red[l] = v.red_time; sat[l] = v.saturation_flow_rate;
            qa[l] = v.Q_alpha; qb[l] = v.Q_beta; qcd[l] = v.Q_cd; qn[l] = v.Q_n; qcp[l] = v.Q_cp; qs[l] = v.Q_s;
            t2[l] = v.t2; Lh[l] = v.L; sth[l] = v.starting_time_in_hour; enh[l] = v.ending_time_in_hour;
            plf[l] = v.peak_load_factor; pvdf[l] = (int)v.vdf_type;
            tt0[l] = g_link_vector[l].travel_time_per_period[tau];
        }
        dump_f64("vdf_fftt" + s, fftt); dump_f64("vdf_cap" + s, cap); dump_f64("vdf_alpha" + s, alpha);
        dump_f64("vdf_beta" + s, beta); dump_f64("vdf_vf" + s, vf); dump_f64("vdf_vc" + s, vc);
        dump_f64("vdf_nlanes" + s, nl); dump_f64("vdf_cap_lane" + s, capl); dump_f64("vdf_qdf" + s, qdf);
        dump_f64("vdf_preload" + s, preload); dump_f64("vdf_penalty" + s, penalty);
        dump_f64("vdf_cycle" + s, cyc); dump_f64("vdf_red" + s, red); dump_f64("vdf_sat" + s, sat);
        dump_f64("vdf_Q_alpha" + s, qa); dump_f64("vdf_Q_beta" + s, qb); dump_f64("vdf_Q_cd" + s, qcd);
        dump_f64("vdf_Q_n" + s, qn); dump_f64("vdf_Q_cp" + s, qcp); dump_f64("vdf_Q_s" + s, qs);
        dump_f64("vdf_t2" + s, t2); dump_f64("vdf_L" + s, Lh); dump_f64("vdf_start_h" + s, sth);
        dump_f64("vdf_end_h" + s, enh); dump_f64("vdf_plf" + s, plf);
        dump_i32("vdf_type" + s, pvdf); dump_f64("tt_initial" + s, tt0);
        for (int at = 0; at < AT; ++at) {
            std::string sa = s + ":at=" + std::to_string(at);
            std::vector<double> pce(L), occ(L), toll(L), lr(L);
            std::vector<int> allowed(L);
            for (int l = 0; l < L; ++l) {
                const CPeriod_VDF& v = g_link_vector[l].VDF_period[tau];
                pce[l] = v.pce[at]; occ[l] = v.occ[at]; toll[l] = v.toll[at]; lr[l] = v.LR_price[at];
                allowed[l] = g_link_vector[l].AllowAgentType(assignment.g_AgentTypeVector[at].agent_type, tau) ? 1 : 0;
            }
            dump_f64("vdf_pce" + sa, pce); dump_f64("vdf_occ" + sa, occ); dump_f64("vdf_toll" + sa, toll);
            dump_f64("vdf_lr_price" + sa, lr); dump_i32("link_allowed" + sa, allowed);
        }
    }
    std::vector<double> vot(AT);
    for (int at = 0; at < AT; ++at) vot[at] = assignment.g_AgentTypeVector[at].value_of_time;
    dump_f64("agent_vot", vot);

    /* demand: every OD cell with od_volume > 0, in o,d,at,tau order */
    std::vector<int> od_o, od_d, od_at, od_tau;
    std::vector<double> od_vol;
    for (int o = 0; o < Z; ++o) {
        int so = g_zone_vector[o].sindex; if (so < 0) continue;
        for (int d = 0; d < Z; ++d) {
            int sd = g_zone_vector[d].sindex; if (sd < 0) continue;
            for (int at = 0; at < AT; ++at) for (int tau = 0; tau < T; ++tau) {
                double v = assignment.g_column_pool[so][sd][at][tau].od_volume;
                if (v > 0) { od_o.push_back(o); od_d.push_back(d); od_at.push_back(at); od_tau.push_back(tau); od_vol.push_back(v); }
            }
        }
    }
    dump_i32("od_o", od_o); dump_i32("od_d", od_d); dump_i32("od_at", od_at); dump_i32("od_tau", od_tau);
    dump_f64("od_volume", od_vol);
    std::vector<double> tdv = { (double)assignment.total_demand_volume };
    dump_f64("total_demand_volume", tdv);
    std::vector<double> odem(Z);
    for (int z = 0; z < Z; ++z) odem[z] = assignment.g_origin_demand_array[z];
    dump_f64("origin_demand", odem);
}

void dump_link_state(const std::string& suffix)
{
    const int L = (int)g_link_vector.size();
    const int T = (int)assignment.g_DemandPeriodVector.size();
    for (int tau = 0; tau < T; ++tau) {
        std::vector<double> tt(L), pv(L), lv(L), perv(L), doc(L), voc(L), P(L), vt2(L);
        for (int l = 0; l < L; ++l) {
            tt[l] = g_link_vector[l].travel_time_per_period[tau];
            pv[l] = g_link_vector[l].PCE_volume_per_period[tau];
            perv[l] = g_link_vector[l].person_volume_per_period[tau];
            lv[l] = g_link_vector[l].VDF_period[tau].link_volume;
            doc[l] = g_link_vector[l].VDF_period[tau].DOC;
            voc[l] = g_link_vector[l].VDF_period[tau].VOC;
            P[l] = g_link_vector[l].VDF_period[tau].P;
            vt2[l] = g_link_vector[l].VDF_period[tau].vt2;
        }
        std::string s = suffix + ":tau=" + std::to_string(tau);
        dump_f64("tt" + s, tt); dump_f64("pce_volume" + s, pv); dump_f64("person_volume" + s, perv);
        dump_f64("vdf_link_volume" + s, lv); dump_f64("doc" + s, doc); dump_f64("voc" + s, voc);
        dump_f64("P" + s, P); dump_f64("vt2" + s, vt2);
    }
}

void dump_columns(const std::string& suffix)
{
    const int Z = (int)g_zone_vector.size();
    const int AT = (int)assignment.g_AgentTypeVector.size();
    const int T = (int)assignment.g_DemandPeriodVector.size();
    std::vector<int> co, cd, cat, ctau, csum, cseq, cnl, cnn, links;
    std::vector<double> cvol, ccost, ctime;
    for (int o = 0; o < Z; ++o) {
        int so = g_zone_vector[o].sindex; if (so < 0) continue;
        for (int d = 0; d < Z; ++d) {
            int sd = g_zone_vector[d].sindex; if (sd < 0) continue;
            for (int at = 0; at < AT; ++at) for (int tau = 0; tau < T; ++tau) {
                CColumnVector& cv = assignment.g_column_pool[so][sd][at][tau];
                if (!(cv.od_volume > 0)) continue;
                for (auto it = cv.path_node_sequence_map.begin(); it != cv.path_node_sequence_map.end(); ++it) {
                    co.push_back(o); cd.push_back(d); cat.push_back(at); ctau.push_back(tau);
                    csum.push_back(it->first); cseq.push_back(it->second.path_seq_no);
                    cnl.push_back(it->second.m_link_size); cnn.push_back(it->second.m_node_size);
                    cvol.push_back(it->second.path_volume);
                    ccost.push_back(it->second.path_gradient_cost); ctime.push_back(it->second.path_travel_time);
                    for (int i = 0; i < it->second.m_link_size; ++i) links.push_back(it->second.path_link_vector[i]);
                }
            }
        }
    }
    dump_i32("col_o" + suffix, co); dump_i32("col_d" + suffix, cd); dump_i32("col_at" + suffix, cat);
    dump_i32("col_tau" + suffix, ctau); dump_i32("col_node_sum" + suffix, csum); dump_i32("col_seq" + suffix, cseq);
    dump_i32("col_nlinks" + suffix, cnl); dump_i32("col_nnodes" + suffix, cnn); dump_f64("col_volume" + suffix, cvol);
    dump_f64("col_cost" + suffix, ccost); dump_f64("col_time" + suffix, ctime); dump_i32("col_links" + suffix, links);
}

int run_trace(int K, int CU, int B, const char* out_path)
{
    const bool light = getenv("DTL_TRACE_LIGHT") && atoi(getenv("DTL_TRACE_LIGHT")) != 0;
    const std::set<int> tree_iters = parse_int_set(getenv("DTL_TRACE_TREES"));
    const std::set<int> edge_iters = parse_int_set(getenv("DTL_TRACE_EDGE_ITERS") ? getenv("DTL_TRACE_EDGE_ITERS") : "0");
    const int stride = getenv("DTL_TRACE_ORIGIN_STRIDE") ? std::max(1, atoi(getenv("DTL_TRACE_ORIGIN_STRIDE"))) : 1;
    g_dump = fopen(out_path, "wb");
    if (!g_dump) { fprintf(stderr, "cannot open %s\n", out_path); return 2; }

    double t_io0 = now_s();
    /* same prologue as network_assignment, main_api.cpp:505-574 */
    assignment.g_number_of_column_generation_iterations = K;
    assignment.g_number_of_column_updating_iterations = CU;
    assignment.g_number_of_ODME_iterations = 0;
    assignment.g_number_of_sensitivity_analysis_iterations = 0;
    assignment.assignment_mode = dta;
    assignment.g_number_of_memory_blocks = std::min(std::max(1, B), (int)MAX_MEMORY_BLOCKS);
    g_read_input_data(assignment);
    g_ReadInformationConfiguration(assignment);
    g_ReadOutputFileConfiguration(assignment);
    g_ReadDemandFileBasedOnDemandFileList(assignment);
    {
        const int n_phys = (int)g_node_vector.size();
        for (int i = 0; i < n_phys; ++i) {
            int zid = g_node_vector[i].zone_org_id;
            if (zid < 1) continue;
            if (assignment.zone_id_to_seed_zone_id_mapping.count(zid)) zid = assignment.zone_id_to_seed_zone_id_mapping[zid];
            int centroid = assignment.g_node_id_to_seq_no_map[assignment.zone_id_to_centriod_node_id_mapping[zid]];
            int zseq = assignment.g_zoneid_to_zone_seq_no_mapping[zid];
            g_add_new_virtual_connector_link(g_node_vector[i].node_seq_no, centroid, g_node_vector[i].agent_type_str, -1);
            g_add_new_virtual_connector_link(centroid, g_node_vector[i].node_seq_no, g_node_vector[i].agent_type_str, zseq);
        }
    }
    g_load_supply_side_scenario_file(assignment);
    g_assign_computing_tasks_to_memory_blocks(assignment);
    double t_io = now_s() - t_io0;

    dump_network();

    const int L = (int)g_link_vector.size();
    const int N = assignment.g_number_of_nodes;
    const int nets = (int)g_NetworkForSP_vector.size();
    const int nblocks = std::min(nets, assignment.g_number_of_memory_blocks);
    const int n_at_tau = (int)(assignment.g_AgentTypeVector.size() * assignment.g_DemandPeriodVector.size());
    int nthreads = omp_get_max_threads();
    std::vector<double> th_sssp(nthreads, 0.0), th_bt(nthreads, 0.0);
    double t_vdf = 0, t_scatter = 0, t_region = 0, t_loop = 0, t_cu = 0;
    long long counted_edges = 0, n_sssp = 0;

    double t_loop0 = now_s();
    std::vector<double> iter_s;
    for (int k = 0; k < K; ++k) {
        const double t_iter0 = now_s();
        double total_distance = 0;
        double t0 = now_s();
        double sysTT = update_link_travel_time_and_cost(k, total_distance);              /* main_api.cpp:606 */
        double t1 = now_s(); t_vdf += t1 - t0;
        if (!light) dump_link_state(":after_vdf:k=" + std::to_string(k));
        t0 = now_s();
        g_reset_and_update_link_volume_based_on_columns(L, k, true, false);               /* main_api.cpp:618 */
        t1 = now_s(); t_scatter += t1 - t0;
        if (!light) dump_link_state(":after_scatter:k=" + std::to_string(k));

        /* per-(network, origin) least cost, summed afterwards in the single-thread order of
           main_api.cpp:667-697 so the result does not depend on the thread count */
        std::vector<std::vector<double> > least(nets);
        for (int n = 0; n < nets; ++n) least[n].assign(g_NetworkForSP_vector[n]->m_origin_node_vector.size(), 0.0);
        const bool dump_trees = tree_iters.count(k) > 0;
        const bool count_edges = edge_iters.count(k) > 0;
        std::vector<long long> th_edges(nthreads, 0);

        t0 = now_s();
#pragma omp parallel for
        for (int pid = 0; pid < nblocks; ++pid) {                                         /* main_api.cpp:667 */
            int tid = omp_get_thread_num();
            for (int blk = 0; blk < n_at_tau; ++blk) {
                int net = blk * assignment.g_number_of_memory_blocks + pid;
                if (net >= nets) continue;
                NetworkForSP* p = g_NetworkForSP_vector[net];
                for (int oi = 0; oi < (int)p->m_origin_node_vector.size(); ++oi) {
                    double a = omp_get_wtime();
                    p->optimal_label_correcting(pid, &assignment, k, oi, false, false, false);    /* :680 */
                    double b = omp_get_wtime();
                    least[net][oi] = p->backtrace_shortest_path_tree(assignment, k, oi, false);   /* :687 */
                    double c = omp_get_wtime();
                    th_sssp[tid] += b - a; th_bt[tid] += c - b;
                    /* Graph500-style edge count: forward-star edges of reachable nodes after filters (the reachable set does
                       not depend on the costs, so one counted iteration gives the per-iteration figure) */
                    long long e = 0;
                    int oz = p->m_origin_zone_seq_no_vector[oi];
                    for (int i = 0; count_edges && i < N; ++i) {
                        if (!(p->m_node_label_cost[i] < MAX_LABEL_COST)) continue;
                        const CNodeForwardStar& fs = p->NodeForwardStarArray[i];
                        for (int j = 0; j < fs.OutgoingLinkSize; ++j) {
                            int tg = p->m_link_outgoing_connector_zone_seq_no_array[fs.OutgoingLinkNoArray[j]];
                            if (tg >= 0 && tg != oz) continue;
                            ++e;
                        }
                    }
                    th_edges[tid] += e;
                    if (dump_trees && (oz % stride) == 0) {
#pragma omp critical
                        {
                            std::string s = ":k=" + std::to_string(k) + ":at=" + std::to_string(p->m_agent_type_no) +
                                            ":tau=" + std::to_string(p->m_tau) + ":z=" + std::to_string(oz);
                            dump_rec("label" + s, DT_F64, p->m_node_label_cost, N);
                            dump_rec("pred_node" + s, DT_I32, p->m_node_predecessor, N);
                            dump_rec("pred_link" + s, DT_I32, p->m_link_predecessor, N);
                        }
                    }
                }
            }
        }
        t1 = now_s(); t_region += t1 - t0;
        if (count_edges) {
            counted_edges = 0;
            for (int t = 0; t < nthreads; ++t) counted_edges += th_edges[t];
        }

        double least_total = 0;
        for (int pid = 0; pid < nblocks; ++pid)
            for (int blk = 0; blk < n_at_tau; ++blk) {
                int net = blk * assignment.g_number_of_memory_blocks + pid;
                if (net >= nets) continue;
                for (size_t oi = 0; oi < least[net].size(); ++oi) { least_total += least[net][oi]; ++n_sssp; }
            }
        if (!light) {
            /* generalized cost arrays as last refreshed by UpdateGeneralizedLinkCost (routing.h:229) */
            for (int blk = 0; blk < n_at_tau; ++blk) {
                int net = blk * assignment.g_number_of_memory_blocks;
                if (net >= nets) continue;
                NetworkForSP* p = g_NetworkForSP_vector[net];
                dump_rec("gen_cost:k=" + std::to_string(k) + ":at=" + std::to_string(p->m_agent_type_no) + ":tau=" + std::to_string(p->m_tau),
                         DT_F64, p->m_link_genalized_cost_array, L);
            }
        }
        double gap = (sysTT - least_total) / max(0.00001, least_total);                    /* main_api.cpp:711 */
        double avg_tt = sysTT / max(1, assignment.total_demand_volume);                    /* :716 */
        if (k == 0)
            assignment.summary_file << "Iteration, CPU Running Time, # of agents, Avg Travel Time(min),  Avg UE gap %" << endl;
        assignment.summary_file << k << "," << 0.0 << "," << assignment.total_demand_volume << "," << avg_tt << "," << gap * 100 << endl;
        std::vector<double> row = { (double)k, sysTT, least_total, gap, avg_tt };
        dump_f64("iter_row:k=" + std::to_string(k), row);
        iter_s.push_back(now_s() - t_iter0);
    }
    t_loop = now_s() - t_loop0;

    /* stage 2, assignment.h:1055-1075 */
    double t_cu0 = now_s();
    assignment.summary_file << "column updating" << endl;
    for (int n = 0; n < CU; ++n) {
        g_update_gradient_cost_and_assigned_flow_in_column_pool(assignment, n, false);
        if (!light) dump_link_state(":cu=" + std::to_string(n));
    }
    t_cu = now_s() - t_cu0;

    /* main_api.cpp:748-759 */
    g_reset_and_update_link_volume_based_on_columns(L, K, false, false);
    double total_distance = 0;
    double final_sysTT = update_link_travel_time_and_cost(K, total_distance);
    dump_link_state(":final");
    std::vector<double> fin = { final_sysTT };
    dump_f64("final_sysTT", fin);
    if (!(getenv("DTL_TRACE_NO_COLUMNS") && atoi(getenv("DTL_TRACE_NO_COLUMNS")) != 0)) dump_columns(":final");
    if (getenv("DTL_TRACE_BACKWARD")) {
        /* the real-time-information layer's per-destination search (main_api.cpp:293-321, assignment.h:1386-1414) on the final
           travel times: one NetworkForSP per (agent type, period), destination = every n-th zone centroid */
        const int bstride = std::max(1, atoi(getenv("DTL_TRACE_BACKWARD")));
        const int Zn = (int)g_zone_vector.size();
        for (int at = 0; at < (int)assignment.g_AgentTypeVector.size(); ++at)
            for (int tau = 0; tau < (int)assignment.g_DemandPeriodVector.size(); ++tau) {
                NetworkForSP* p = new NetworkForSP();
                p->m_agent_type_no = at; p->m_tau = tau;
                p->AllocateMemory(assignment.g_number_of_nodes, assignment.g_number_of_links, assignment.g_number_of_analysis_districts);
                float now_min = 1000.0f;
                for (int z = 0; z < Zn; z += bstride) {
                    p->m_RT_dest_node = g_zone_vector[z].node_seq_no; p->m_RT_dest_zone = z;
                    now_min += 10.0f + assignment.g_info_updating_freq_in_min; /* routing.h:1011: only runs when the clock has advanced */
                    p->optimal_backward_label_correcting_from_destination(0, &assignment, now_min, z, p->m_RT_dest_node, -1, 0);
                    std::string s = ":at=" + std::to_string(at) + ":tau=" + std::to_string(tau) + ":z=" + std::to_string(z);
                    dump_rec("bw_label" + s, DT_F64, p->m_node_label_cost, N);
                    dump_rec("bw_pred_node" + s, DT_I32, p->m_node_predecessor, N);
                    dump_rec("bw_pred_link" + s, DT_I32, p->m_link_predecessor, N);
                }
            }
    }
    assignment.summary_file.flush();

    double s_sssp = 0, s_bt = 0;
    for (int t = 0; t < nthreads; ++t) { s_sssp += th_sssp[t]; s_bt += th_bt[t]; }
    std::vector<double> tm = { t_io, t_loop, t_vdf, t_scatter, t_region, s_sssp, s_bt, t_cu, (double)nthreads,
                               (double)counted_edges, (double)n_sssp, (double)nblocks };
    dump_f64("timing", tm);
    fclose(g_dump); g_dump = nullptr;
    printf("\nREFTIMING {\"io_s\": %.6f, \"loop_s\": %.6f, \"vdf_s\": %.6f, \"scatter_s\": %.6f, \"sp_region_s\": %.6f, "
           "\"sssp_thread_s\": %.6f, \"backtrace_thread_s\": %.6f, \"cu_s\": %.6f, \"threads\": %d, \"blocks\": %d, "
           "\"counted_edges\": %lld, \"n_sssp\": %lld, \"K\": %d, \"CU\": %d, \"nodes\": %d, \"links\": %d, \"zones\": %d, \"iter_s\": [",
           t_io, t_loop, t_vdf, t_scatter, t_region, s_sssp, s_bt, t_cu, nthreads, nblocks, counted_edges, n_sssp, K, CU,
           N, L, (int)g_zone_vector.size());
    for (size_t i = 0; i < iter_s.size(); ++i) printf("%s%.6f", i ? ", " : "", iter_s[i]);
    printf("]}\n");
    fflush(stdout);
    return 0;
}

} // namespace

int main(int argc, char** argv)
{
    if (argc < 5) {
        fprintf(stderr, "usage: %s golden|trace|golden_dump K CU B [out.bin] [S]\n", argv[0]);
        return 2;
    }
    std::string mode = argv[1];
    int K = atoi(argv[2]), CU = atoi(argv[3]), B = atoi(argv[4]);
    if (mode == "golden") {
        network_assignment(1, K, CU, 0, 0, 0, B);
        return 0;
    }
    if (mode == "golden_dump") {
        /* the unmodified entry point with a sensitivity-analysis stage of S iterations (main_api.cpp:771-1042), then the state
           it leaves in the reference's globals: dtalite_ref golden_dump K CU B out.bin S */
        const char* out = argc >= 6 ? argv[5] : "ref_trace.bin";
        const int S = argc >= 7 ? atoi(argv[6]) : 0;
        const int ODME = argc >= 8 ? atoi(argv[7]) : 0; /* OD demand estimation iterations (measurement.csv, ODME.h:91-516) */
        network_assignment(1, K, CU, ODME, S, 0, B);
        g_dump = fopen(out, "wb");
        if (!g_dump) { fprintf(stderr, "cannot open %s\n", out); return 2; }
        dump_network();
        const int L = (int)g_link_vector.size(), AT = (int)assignment.g_AgentTypeVector.size(), T = (int)assignment.g_DemandPeriodVector.size();
        std::vector<int> rt(AT);
        for (int at = 0; at < AT; ++at) rt[at] = assignment.g_AgentTypeVector[at].real_time_information;
        dump_i32("agent_rt_info", rt);
        std::vector<double> apce(AT), aocc(AT);
        for (int at = 0; at < AT; ++at) { apce[at] = assignment.g_AgentTypeVector[at].PCE; aocc[at] = assignment.g_AgentTypeVector[at].OCC; }
        dump_f64("agent_pce", apce); dump_f64("agent_occ", aocc);
        for (int tau = 0; tau < T; ++tau) {
            std::vector<int> flag(L);
            std::vector<double> lanes(L), nl(L), cap(L), fftt(L);
            for (int l = 0; l < L; ++l) {
                const CPeriod_VDF& v = g_link_vector[l].VDF_period[tau];
                flag[l] = v.network_design_flag; lanes[l] = v.sa_lanes_change; nl[l] = v.nlanes; cap[l] = v.BPR_period_capacity; fftt[l] = v.FFTT;
            }
            std::string sfx = ":tau=" + std::to_string(tau);
            std::vector<double> obs(L), dev(L);
            std::vector<int> ub(L);
            for (int l = 0; l < L; ++l) {
                const CPeriod_VDF& v = g_link_vector[l].VDF_period[tau];
                obs[l] = v.obs_count; dev[l] = v.est_count_dev; ub[l] = v.upper_bound_flag;
            }
            dump_f64("obs_count" + sfx, obs); dump_f64("est_count_dev" + sfx, dev); dump_i32("upper_bound_flag" + sfx, ub);
            dump_i32("design_flag" + sfx, flag); dump_f64("sa_lanes" + sfx, lanes);
            dump_f64("sa_nlanes_after" + sfx, nl); dump_f64("sa_cap_after" + sfx, cap); dump_f64("sa_fftt_after" + sfx, fftt);
        }
        dump_link_state(":final");
        dump_columns(":final");
        {   /* per-column and per-cell flags of the SA stage, in dump_columns order */
            const int Z = (int)g_zone_vector.size();
            std::vector<int> impacted, sa_flag, od_impact, at_impact;
            for (int o = 0; o < Z; ++o) {
                int so = g_zone_vector[o].sindex; if (so < 0) continue;
                for (int d = 0; d < Z; ++d) {
                    int sd = g_zone_vector[d].sindex; if (sd < 0) continue;
                    for (int at = 0; at < AT; ++at) for (int tau = 0; tau < T; ++tau) {
                        CColumnVector& cv = assignment.g_column_pool[so][sd][at][tau];
                        if (!(cv.od_volume > 0)) continue;
                        for (auto it = cv.path_node_sequence_map.begin(); it != cv.path_node_sequence_map.end(); ++it) {
                            impacted.push_back(it->second.impacted_path_flag);
                            sa_flag.push_back(it->second.b_sensitivity_analysis_flag ? 1 : 0);
                            od_impact.push_back(cv.OD_impact_flag);
                            at_impact.push_back(cv.at_od_impacted_flag_map.count(at) ? 1 : 0);
                        }
                    }
                }
            }
            dump_i32("col_impacted_path_flag:final", impacted); dump_i32("col_sa_flag:final", sa_flag);
            dump_i32("col_od_impact_flag:final", od_impact); dump_i32("col_at_od_impacted:final", at_impact);
        }
        fclose(g_dump); g_dump = nullptr;
        fflush(NULL);
        _exit(0);
    }
    if (mode == "trace") {
        const char* out = argc >= 6 ? argv[5] : "ref_trace.bin";
        int rc = run_trace(K, CU, B, out);
        /* skip static destructors of the reference's globals (slow, and not under test) */
        fflush(NULL);
        _exit(rc);
    }
    fprintf(stderr, "unknown mode %s\n", mode.c_str());
    return 2;
}
