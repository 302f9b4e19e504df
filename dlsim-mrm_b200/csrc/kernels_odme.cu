// OD demand estimation: the device side of Assignment::Demand_ODME (reference ODME.h:91-516) and of its
// scatter g_reset_and_update_link_volume_based_on_ODME_columns (assignment.h:263-563).
//   * odme_path_time_kernel / odme_cell_kernel: path travel times on the CURRENT link times and the per-cell terms of the
//     summary row (demand, travel time, user-equilibrium gap; assignment.h:347-430), summed afterwards by the fixed-tree
//     reduction of the engine (the reference adds them inside a parallel loop: its own sums depend on the thread count);
//   * odme_scatter_kernel: column volumes -> Q31.32 link volumes with the ODME addend -- the FLOAT product of the float path
//     volume and the agent type's float PCE (assignment.h:308-309, 449), not the per-link double PCE of the assignment scatter;
//   * odme_dev_kernel: est_count_dev = volume + preload - count on the links that carry a count (assignment.h:476);
//   * odme_update_kernel: one thread per column: the gradient is the float sum of the count deviations along the path in path
//     order (ODME.h:396-421), the step 0.05 x gradient clamped to +-10 % of the path volume, the volume floored at 1
//     (ODME.h:428-449).  Columns do not interact, so the result does not depend on the order they are visited in.
#include "dtl_internal.h"

namespace dtl {

__global__ void __launch_bounds__(128) odme_path_time_kernel(OdmeLaunch a)
{
    const size_t slot = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const ColumnPool& P = a.pool;
    if (slot >= (size_t)P.n_od * P.cap) return;
    const int cell = (int)(slot / P.cap);
    if ((int)(slot - (size_t)cell * P.cap) >= P.count[cell]) return;
    const double* tt = a.tt + (size_t)a.od_tau[cell] * a.L;
    const int n = P.n_links[slot];
    const int* links = P.arena + P.offset[slot];
    double t = 0;
    for (int q = 0; q < n; ++q) t += tt[links[q]];                      // assignment.h:378-383, path order
    P.time[slot] = t;
}

__global__ void __launch_bounds__(128) odme_cell_kernel(OdmeLaunch a)
{
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= a.pool.n_od) return;
    const ColumnPool& P = a.pool;
    const int cnt = P.count[cell];
    const size_t base = (size_t)cell * P.cap;
    double demand = 0, time = 0, ue = 0;
    // the reference walks the std::map (ascending node_sum); the least-cost column is the first strict minimum in that order
    double least = 999999;
    int least_j = -1, least_key = 0;
    for (int j = 0; j < cnt; ++j) {
        const double v = P.volume[base + j], t = P.time[base + j];
        demand += v;                                                    // :372
        time += t * v;                                                  // :390
        if (cnt == 1) break;                                            // :393-396
        const int key = P.node_sum[base + j];
        if (t < 999999 && (least_j < 0 || t < least || (t == least && key < least_key))) { least = t; least_j = j; least_key = key; }
    }
    if (cnt >= 2)
        for (int j = 0; j < cnt; ++j)
            if (j != least_j) ue += (P.time[base + j] - least) * P.volume[base + j]; // :419-427
    double* s = a.od_sums + (size_t)cell * 4;
    s[0] = demand; s[1] = time; s[2] = ue; s[3] = 0.0;
}

__global__ void __launch_bounds__(256) odme_scatter_kernel(OdmeLaunch a)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= a.pool.n_od) return;
    const int cell = warp;
    const ColumnPool& P = a.pool;
    const int cnt = P.count[cell];
    const int at = a.od_at[cell], tau = a.od_tau[cell];
    const float pce = a.at_pce[at], occ = a.at_occ[at];
    unsigned long long* acc = a.pce_fix + (size_t)tau * a.L;
    unsigned long long* per = a.person_fix + (size_t)tau * (1 + a.AT) * a.L;
    unsigned long long* per_at = per + (size_t)(1 + at) * a.L;
    const size_t base = (size_t)cell * P.cap;
    for (int j = 0; j < cnt; ++j) {
        const float f = (float)P.volume[base + j];                      // assignment.h:435
        if (f == 0.0f) continue;
        const unsigned long long x = (unsigned long long)__double2ll_rn((double)__fmul_rn(f, pce) * DTL_FIX_SCALE); // :449 float product
        const unsigned long long y = (unsigned long long)__double2ll_rn((double)__fmul_rn(f, occ) * DTL_FIX_SCALE); // :450
        const unsigned long long z = (unsigned long long)__double2ll_rn((double)f * DTL_FIX_SCALE);                 // :451 pure volume
        const int n = P.n_links[base + j];
        const int* links = P.arena + P.offset[base + j];
        for (int q = lane; q < n; q += 32) {
            const int l = links[q];
            atomicAdd(&acc[l], x);
            atomicAdd(&per[l], y);
            atomicAdd(&per_at[l], z);
        }
    }
}

__global__ void __launch_bounds__(128) odme_dev_kernel(OdmeLaunch a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n_sensors) return;
    const int idx = a.sensor_idx[i];                                    // tau * L + link
    const double dev = a.link_volume[idx] - a.obs[idx];                 // assignment.h:476: (PCE volume + preload) - obs_count
    a.est_dev[idx] = dev;
    a.sensor_dev[i] = dev;
}

__global__ void __launch_bounds__(128) odme_update_kernel(OdmeLaunch a)
{
    const size_t slot = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const ColumnPool& P = a.pool;
    if (slot >= (size_t)P.n_od * P.cap) return;
    const int cell = (int)(slot / P.cap);
    if ((int)(slot - (size_t)cell * P.cap) >= P.count[cell]) return;
    const size_t off = (size_t)a.od_tau[cell] * a.L;
    const int n = P.n_links[slot];
    const int* links = P.arena + P.offset[slot];
    float g = 0;                                                        // float path_gradient_cost, ODME.h:277
    for (int q = 0; q < n; ++q) {                                       // ODME.h:396-421
        const size_t i = off + links[q];
        if (a.obs[i] >= 1) {
            const double dev = a.est_dev[i];
            const int ub = a.ub[i];
            if (ub == 0) g = (float)((double)g + dev);
            if (ub == 1 && dev > 0) g = (float)((double)g + dev);
        }
    }
    P.cost[slot] = (double)g;                                           // ODME.h:426
    const float step = 0.05f;                                           // ODME.h:428
    const double vol = P.volume[slot];
    double change = (double)step * (1.0 * (double)g + (1 - 1.0) * 0.0); // ODME.h:431-434 (the UE-gap term has weight 0)
    const float lo = (float)(vol * 0.1 * (-1)), hi = (float)(vol * 0.1); // ODME.h:439-440
    if (change < (double)lo) change = (double)lo;
    if (change > (double)hi) change = (double)hi;
    const double nv = vol - change;
    P.volume[slot] = 1.0 > nv ? 1.0 : nv;                               // ODME.h:449 max(1.0, ...)
}

void launch_odme_stats(const OdmeLaunch& a, cudaStream_t st)
{
    if (a.pool.n_od == 0) return;
    const size_t n_slots = (size_t)a.pool.n_od * a.pool.cap;
    odme_path_time_kernel<<<(unsigned)((n_slots + 127) / 128), 128, 0, st>>>(a);
    odme_cell_kernel<<<(a.pool.n_od + 127) / 128, 128, 0, st>>>(a);
}

void launch_odme_scatter(const OdmeLaunch& a, cudaStream_t st)
{
    if (a.pool.n_od == 0) return;
    odme_scatter_kernel<<<(a.pool.n_od + 7) / 8, 256, 0, st>>>(a);
}

void launch_odme_dev(const OdmeLaunch& a, cudaStream_t st)
{
    if (a.n_sensors == 0) return;
    odme_dev_kernel<<<(a.n_sensors + 127) / 128, 128, 0, st>>>(a);
}

void launch_odme_update(const OdmeLaunch& a, cudaStream_t st)
{
    if (a.pool.n_od == 0) return;
    const size_t n_slots = (size_t)a.pool.n_od * a.pool.cap;
    odme_update_kernel<<<(unsigned)((n_slots + 127) / 128), 128, 0, st>>>(a);
}

} // namespace dtl
