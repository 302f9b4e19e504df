// K4: elementwise link performance (BPR / queue VDF) over (period, link), fused with
//   * the fixed-point -> FP64 conversion of the all-reduced PCE volume,
//   * the system-travel-time reduction of update_link_travel_time_and_cost (assignment.h:246-258),
//   * the generalized-cost refresh of every forward star (routing.h:229-231), which the reference
//     repeats once per origin although the result does not depend on the origin.
// Arithmetic follows CPeriod_VDF::calculate_travel_time_based_on_QVDF (reference VDF.h:139-367) and
// CLink::calculate_dynamic_VDFunction (main_api.cpp:404-435) operation by operation in FP64; the file
// is compiled with -fmad=false so that no multiply-add is contracted (the reference is baseline x86-64).
// Roofline: HBM, 160 B per (link, period).
#include "dtl_internal.h"

namespace dtl {

__device__ __forceinline__ double mx(double a, double b) { return a > b ? a : b; } // (a>b)?a:b macro semantics
__device__ __forceinline__ double mn(double a, double b) { return a < b ? a : b; }

struct VdfCore {
    double tt, doc, voc, P, vt2, avg_queue_speed, avg_speed_bpr, Q_mu, Q_gamma, t0, t3, RTT, lane_based_D;
};

__device__ __forceinline__ void vdf_core(const VdfParams& p, size_t i, double volume, VdfCore& o)
{
    const double qdf = p.qdf[i], nlanes = p.nlanes[i], cap_lane = p.cap_lane[i];
    const double vc = p.vc[i], vf = p.vf[i], Q_alpha = p.Q_alpha[i], Q_beta = p.Q_beta[i];
    const double beta = p.beta[i], alpha = p.alpha[i], FFTT = p.fftt[i], BPR_cap = p.cap[i];
    const double Q_cd = p.Q_cd[i], Q_n = p.Q_n[i], Q_cp = p.Q_cp[i], Q_s = p.Q_s[i];
    const float cycle_length = p.cycle_length[i];

    double lane_based_D = mx(0.0, volume) * qdf / mx(0.000001, nlanes); // VDF.h:145
    double DOC = lane_based_D / mx(0.00001, cap_lane);                   // :149
    double avg_queue_speed = vc / (1.0 + Q_alpha * pow(DOC, Q_beta));    // :156
    double RTT = 1.0 / vc;                                               // :162
    double Q_n_current = Q_n;
    if (DOC < 1.0) {                                                     // :164-177 free-flow regime
        double vf_alpha = (1.0 + Q_alpha) * vf / mx(0.0001, vc) - 1.0;
        avg_queue_speed = vf / (1.0 + vf_alpha * pow(DOC, beta));
        Q_n_current = beta;
        RTT = 1.0 / mx(0.01, vf);
    }
    double VOC = volume / mx(0.00001, BPR_cap);                          // :184
    double avg_speed_BPR = vf / (1.0 + alpha * pow(VOC, beta));          // :191
    double avg_tt = FFTT * vf / mx(0.1, avg_speed_BPR);                  // :192
    if (p.vdf_type[i] == 0) avg_tt = FFTT * vf / mx(0.1, avg_queue_speed); // :196-202 q_vdf
    if (cycle_length >= 1) {                                             // :204-210 signal override
        const float red_time = p.red_time[i], sat = p.sat_flow[i];
        float s_bar = (float)(1.0 / 60.0 * red_time * red_time / (2 * cycle_length));
        double lambda = lane_based_D;
        float uniform_delay = (float)(s_bar / mx(1 - lambda / sat, (double)0.1f));
        avg_tt = uniform_delay + FFTT;
    }
    double P = Q_cd * pow(DOC, Q_n_current);                             // :217
    double base = Q_cp * pow(P, Q_s) + 1.0;
    double vt2 = vc / mx(0.001, base);                                   // :220
    o.tt = avg_tt; o.doc = DOC; o.voc = VOC; o.P = P; o.vt2 = vt2;
    o.avg_queue_speed = avg_queue_speed; o.avg_speed_bpr = avg_speed_BPR;
    o.t0 = p.t2[i] - 0.5 * P; o.t3 = p.t2[i] + 0.5 * P;                  // :240-241
    o.Q_mu = mn(cap_lane, lane_based_D / P);                             // :249
    double wt2 = 1.0 / vt2 - RTT;                                        // :256
    o.Q_gamma = wt2 * 64 * o.Q_mu / pow(P, 4);                           // :260
    o.RTT = RTT; o.lane_based_D = lane_based_D;
}

// time-dependent speed of one 5-minute slot, VDF.h:277-319
__device__ __forceinline__ double td_speed_at(double t, const VdfCore& c, double start_h, double end_h, double vf,
                                              double vc)
{
    if (c.t0 <= t && t <= c.t3) {
        double td_queue = 0.25 * c.Q_gamma * pow((t - c.t0), 2) * pow((t - c.t3), 2);
        double td_w = td_queue / mx(0.001, c.Q_mu);
        return 1.0 / (td_w + c.RTT);
    } else if (t < c.t0) {
        double factor = (t - start_h) / mx(0.001, c.t0 - start_h);
        return (1 - factor) * vf + factor * mx(vc, c.avg_queue_speed);
    }
    double factor = (t - c.t3) / mx(0.001, end_h - c.t3);
    return (1 - factor) * mx(vc, c.avg_queue_speed) + (factor)*vf;
}

constexpr int VDF_THREADS = 256;
int vdf_blocks(int n) { return (n + VDF_THREADS - 1) / VDF_THREADS; }

template <bool DETAIL>
__global__ void __launch_bounds__(VDF_THREADS) vdf_kernel(VdfLaunch a)
{
    __shared__ double red[VDF_THREADS];
    const int n = a.T * a.L;
    const int i = blockIdx.x * VDF_THREADS + threadIdx.x;
    double contrib = 0.0;
    if (i < n) {
        const int tau = i / a.L, l = i - tau * a.L;
        double volume;
        if (a.volume_override) volume = a.volume_override[i];
        else {
            double pce_vol = (double)(long long)a.s.pce_fix[i] * (1.0 / DTL_FIX_SCALE); // exact: power-of-two scale
            volume = pce_vol + a.p.preload[i] + 0.0; /* + sa_volume, main_api.cpp:413 */
        }
        VdfCore c;
        vdf_core(a.p, (size_t)i, volume, c);
        a.s.tt[i] = c.tt; a.s.link_volume[i] = volume;
        a.s.doc[i] = c.doc; a.s.voc[i] = c.voc; a.s.P[i] = c.P; a.s.vt2[i] = c.vt2;
        contrib = c.tt * volume; // assignment.h:246-258
        if (DETAIL) {
            a.s.avg_queue_speed[i] = c.avg_queue_speed; a.s.avg_speed_bpr[i] = c.avg_speed_bpr;
            a.s.Q_mu[i] = c.Q_mu; a.s.Q_gamma[i] = c.Q_gamma; a.s.t0[i] = c.t0; a.s.t3[i] = c.t3;
            const double start_h = a.p.start_h[i], end_h = a.p.end_h[i], vf = a.p.vf[i], vc = a.p.vc[i];
            double severe = 0;
            for (int t_in_min = (int)(start_h * 60); t_in_min <= end_h * 60; t_in_min += 5) {
                double sp = td_speed_at(t_in_min / 60.0, c, start_h, end_h, vf, vc);
                if (sp < vf * 0.5) severe += 5.0 / 60;
            }
            a.s.severe_P[i] = severe;
        }
        if (a.write_cost) {
            const double pen = a.p.penalty[i];
            for (int at = 0; at < a.AT; ++at) {
                const int fs = at * a.T + tau;
                const int e = a.link_edge[fs][l];
                if (e >= 0) {
                    const size_t ai = ((size_t)tau * a.AT + at) * a.L + l;
                    // routing.h:229-231: tt + penalty + LR_price + toll / VOT(float) * 60
                    const double gc = c.tt + pen + a.lr_price[ai] + a.toll[ai] / a.vot[at] * 60;
                    a.edges[fs][e].cost = gc;
                    // K1 orders labels by their FP64 bit pattern, which is only valid for non-negative sums: a negative (or NaN)
                    // generalized cost -- a negative penalty, toll or LR price larger than the travel time -- is refused
                    if (!(gc >= 0.0)) atomicOr(a.error_flag, 4);
                }
            }
        }
    }
    red[threadIdx.x] = contrib;
    __syncthreads();
    for (int s = VDF_THREADS / 2; s > 0; s >>= 1) { // fixed tree: bit-reproducible for a given (T, L)
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) a.block_sums[blockIdx.x] = red[0];
}

__global__ void __launch_bounds__(1024) sum_kernel(const double* in, long long n, int stride, double* out)
{
    __shared__ double red[1024];
    double s = 0;
    for (long long i = threadIdx.x; i < n; i += 1024) s += in[i * stride];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int w = 512; w > 0; w >>= 1) {
        if (threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = red[0];
}

void launch_sum(const double* in, long long n, int stride, double* out, cudaStream_t st)
{
    sum_kernel<<<1, 1024, 0, st>>>(in, n, stride, out);
}

void launch_vdf(const VdfLaunch& a, cudaStream_t st)
{
    const int n = a.T * a.L;
    const int blocks = vdf_blocks(n);
    if (blocks == 0) return;
    if (a.detail) vdf_kernel<true><<<blocks, VDF_THREADS, 0, st>>>(a);
    else vdf_kernel<false><<<blocks, VDF_THREADS, 0, st>>>(a);
    launch_sum(a.block_sums, blocks, 1, a.sys_tt, st);
}

__global__ void model_speed_kernel(int L, int tau, VdfParams p, LinkState s, int first_slot, int n_slots, float* out)
{
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= L) return;
    const size_t i = (size_t)tau * L + l;
    VdfCore c;
    vdf_core(p, i, s.link_volume[i], c);
    const double start_h = p.start_h[i], end_h = p.end_h[i], vf = p.vf[i], vc = p.vc[i];
    for (int t_in_min = (int)(start_h * 60); t_in_min <= end_h * 60; t_in_min += 5) {
        int slot = t_in_min / 5;
        if (slot < first_slot || slot >= first_slot + n_slots || slot >= DTL_SPEED_SLOTS) continue;
        out[(size_t)l * n_slots + (slot - first_slot)] = (float)td_speed_at(t_in_min / 60.0, c, start_h, end_h, vf, vc);
    }
}

void launch_model_speed(int T, int L, int tau, const VdfParams& p, const LinkState& s, int first_slot, int n_slots,
                        float* out, cudaStream_t st)
{
    (void)T;
    if (L == 0) return;
    model_speed_kernel<<<(L + 127) / 128, 128, 0, st>>>(L, tau, p, s, first_slot, n_slots, out);
}

} // namespace dtl
