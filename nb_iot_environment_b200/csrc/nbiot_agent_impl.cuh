/*
 * nbiot_agent_impl.cuh -- the model-based NPRACH agent on the device (include/nbiot_agent.h), part of the one
 * translation unit of libnbiot_b200.so (included at the end of nbiot_kernels.cu).
 *
 * One thread per cell: the agent runs once per 3000 simulated subframes per cell on ~300 bytes of its decision
 * record, so the work is a few hundred table lookups per cell -- HBM-bound on the records (1.6 KB each), microseconds
 * for thousands of cells; the point of having it on the device is that the NPRACH-control loop
 * (agent -> nbiot_advance -> records -> agent) never crosses PCIe. Floating-point operations follow
 * agent_nprach.py:62-146 in order (the library is built with -fmad=false), so actions and estimates are
 * bit-identical to the reference agent on the same inputs (tests/test_mb_agent.py).
 */
#ifndef NBIOT_AGENT_IMPL_CUH
#define NBIOT_AGENT_IMPL_CUH

#include "nbiot_agent.h"

struct NbMbTablesDev {
    const int16_t *ml_est; const double *cfg_rate; const uint8_t *cfg_conf; const uint8_t *rate_conf;
    int32_t cfg_n[3]; int32_t period_list[6];
    double m3_det[15]; uint8_t m3_ce[15]; uint8_t th_default[2];
    double min_value, max_value, step;
    int32_t buffer_size, no_th;
};

struct nbiot_mb_agent {
    nbiot_handle_t *env;
    NbMbTablesDev T;
    nbiot_mb_state_t *state;
    int16_t *ml_est; double *cfg_rate; uint8_t *cfg_conf, *rate_conf;
    unsigned long long *faults;
};

/* CPython float floor division, float_divmod (Objects/floatobject.c): the agent indexes its table with
 * int((rate - min_value) // step) (agent_nprach.py:102), which is not floor(x / step) in the last place */
__device__ __forceinline__ double nb_py_floordiv(double vx, double wx)
{
    double mod = fmod(vx, wx);
    double div = (vx - mod) / wx;
    if (mod != 0.0) { if ((wx < 0) != (mod < 0)) { mod += wx; div -= 1.0; } }
    if (div != 0.0) { double f = floor(div); if (div - f > 0.5) f += 1.0; return f; }
    return copysign(0.0, vx / wx);
}

__device__ __forceinline__ void nb_mb_configurator(const NbMbTablesDev &T, const double m3[3], const double rates[3], uint8_t conf[6])
{                                                               /* configurator.configurator / find_conf (:161-196) */
    for (int ce = 0; ce < 3; ++ce) {
        double target = rates[ce] / m3[ce];
        int n = T.cfg_n[ce], j = n - 1;
        for (int q = 0; q < n; ++q) if (T.cfg_rate[ce * NBIOT_MB_K + q] > target) { j = q; break; }
        conf[ce] = T.cfg_conf[(ce * NBIOT_MB_K + j) * 2]; conf[ce + 3] = T.cfg_conf[(ce * NBIOT_MB_K + j) * 2 + 1];
    }
}

__global__ void nbiot_mb_init_kernel(nbiot_mb_state_t *st, int n, double margin, uint8_t th0, uint8_t th1)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    nbiot_mb_state_t s;                                          /* agent_nprach.py:44-57 */
    s.msg3_detection[0] = s.msg3_detection[1] = s.msg3_detection[2] = 0.9;
    s.lambda_estimate = 0.01; s.margin = margin; s.n = 0; s.k = 0;
    for (int j = 0; j < 6; ++j) s.conf[j] = 2;
    s.th_conf[0] = th0; s.th_conf[1] = th1;
    st[i] = s;
}

__global__ void nbiot_mb_act_kernel(NbMbTablesDev T, const nbiot_record_t *rec, const int32_t *service, int hist_cap, nbiot_mb_state_t *st, int n,
                                    const double *margin, uint8_t *actions, double *samples, unsigned long long *faults)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const nbiot_record_t *r = rec + i;
    if (r->state != NB_ST_NPRACH_UPDATE) return;                /* the agent only acts in its own state (controller.py:283-288) */
    nbiot_mb_state_t s = st[i];
    if (margin) s.margin = margin[i];                           /* set_parameter */
    s.n += 1;
    double rates[3];
    double lambda = 0.0;
    for (int ce = 0; ce < 3; ++ce) {                            /* agent_nprach.py:75-86 */
        int period = T.period_list[s.conf[ce + 3]];
        int cnt = r->n_occ[ce] < NBIOT_MAX_OCC ? r->n_occ[ce] : NBIOT_MAX_OCC;
        double a = 0.0;
        if (cnt > 0) {                                          /* estimate_incoming_traffic (detection_model.py:45-55) */
            const int16_t *ml = T.ml_est + (size_t)s.conf[ce] * NBIOT_MB_D * NBIOT_MB_C;
            long long sum = 0;
            for (int o = 0; o < cnt; ++o) {
                int d = r->det_hist[ce][o], c = r->col_hist[ce][o];
                if (d < 0 || d >= NBIOT_MB_D || c < 0 || c >= NBIOT_MB_C) {
                    atomicAdd(faults, 1ull);
                    d = d < 0 ? 0 : d >= NBIOT_MB_D ? NBIOT_MB_D - 1 : d; c = c < 0 ? 0 : c >= NBIOT_MB_C ? NBIOT_MB_C - 1 : c;
                }
                sum += ml[d * NBIOT_MB_C + c];
            }
            a = (double)sum / (double)cnt;
        }
        double lam = a / (double)period;
        lambda = lambda + lam;
        rates[ce] = s.margin * lam;
    }
    s.lambda_estimate = lambda;
    double arrival_rate = (0.0 + rates[0]) + rates[1] + rates[2];      /* Python sum(): left to right from 0 */
    uint8_t conf[6], th[2];
    if (s.k < T.buffer_size) {                                  /* still building the model (:90-97) */
        s.k += r->n_incoming;
        double acc[3] = { 0.0, 0.0, 0.0 }, mar[3] = { 0.0, 0.0, 0.0 };
        for (int b = 0; b < 14; ++b) { int ce = T.m3_ce[b]; acc[ce] = acc[ce] + r->distribution[b] * T.m3_det[b]; mar[ce] = mar[ce] + r->distribution[b]; }
        acc[0] = acc[0] + r->distribution[14] * T.m3_det[14]; mar[0] = mar[0] + r->distribution[14];
        const double dflt[3] = { 0.92, 0.9, 0.8 };
        for (int ce = 0; ce < 3; ++ce) s.msg3_detection[ce] = mar[ce] > 0 ? acc[ce] / mar[ce] : dflt[ce];
        nb_mb_configurator(T, s.msg3_detection, rates, conf);
        th[0] = s.th_conf[0]; th[1] = s.th_conf[1];
    } else if (T.no_th) {                                       /* NPRACH parameters only (:99-101) */
        nb_mb_configurator(T, s.msg3_detection, rates, conf);
        th[0] = s.th_conf[0]; th[1] = s.th_conf[1];
    } else {                                                    /* thresholds and NPRACH parameters from the table (:103-113) */
        int row;
        int stored = 0;
        if (arrival_rate >= T.min_value && arrival_rate <= T.max_value) {
            double q = nb_py_floordiv(arrival_rate - T.min_value, T.step);
            long long idx = (long long)q + 1;
            row = idx < NBIOT_MB_RATE_ROWS - 1 ? (int)idx : NBIOT_MB_RATE_ROWS - 1;
        } else if (arrival_rate < T.min_value) row = 0;
        else { row = NBIOT_MB_RATE_ROWS - 1; stored = 1; }
        const uint8_t *e = T.rate_conf + row * 8;
        th[0] = e[0]; th[1] = e[1];
        for (int j = 0; j < 6; ++j) conf[j] = e[2 + j];
        if (stored) {                                           /* only this branch remembers its choice (:109-113) */
            for (int j = 0; j < 6; ++j) s.conf[j] = conf[j];
            s.th_conf[0] = th[0]; s.th_conf[1] = th[1];
        }
    }
    st[i] = s;
    uint8_t *out = actions + (size_t)i * 8;
    out[0] = th[0]; out[1] = th[1];
    for (int j = 0; j < 6; ++j) out[2 + j] = conf[j];
    if (samples) {                                              /* agent_nprach.py:115-124 */
        int nd = r->departures, lim = nd < hist_cap ? nd : hist_cap;
        long long sum = 0;
        for (int q = 0; q < lim; ++q) sum += service[(size_t)i * hist_cap + q];
        samples[(size_t)i * 4 + 0] = (double)nd;
        samples[(size_t)i * 4 + 1] = r->NPRACH_occupation;
        samples[(size_t)i * 4 + 2] = nd > 0 ? (double)sum / (double)nd : 0.0;
        samples[(size_t)i * 4 + 3] = s.margin;
    }
}

extern "C" {

int nbiot_mb_agent_create(nbiot_handle_t *env, const nbiot_mb_tables_t *t, int buffer_size, int no_th, double margin,
                          void *state_dev, void *stream, nbiot_mb_agent_t **out)
{
    if (!env || !t || !state_dev || !out || !t->ml_est || !t->cfg_rate || !t->cfg_conf || !t->rate_conf) return set_err(NBIOT_E_INVALID, "nbiot_mb_agent_create: null argument");
    for (int ce = 0; ce < 3; ++ce) if (t->cfg_n[ce] < 1 || t->cfg_n[ce] > NBIOT_MB_K) return set_err(NBIOT_E_INVALID, "nbiot_mb_agent_create: bad rate list length");
    CK(cudaSetDevice(env->device));
    cudaStream_t st = (cudaStream_t)stream;
    nbiot_mb_agent_t *a = new nbiot_mb_agent_t();
    memset(a, 0, sizeof *a);
    a->env = env; a->state = (nbiot_mb_state_t *)state_dev;
    size_t n_ml = (size_t)4 * NBIOT_MB_D * NBIOT_MB_C;
    CK(cudaMalloc(&a->ml_est, n_ml * sizeof(int16_t)));
    CK(cudaMalloc(&a->cfg_rate, 3 * NBIOT_MB_K * sizeof(double)));
    CK(cudaMalloc(&a->cfg_conf, 3 * NBIOT_MB_K * 2));
    CK(cudaMalloc(&a->rate_conf, NBIOT_MB_RATE_ROWS * 8));
    CK(cudaMalloc(&a->faults, sizeof(unsigned long long)));
    CK(cudaMemsetAsync(a->faults, 0, sizeof(unsigned long long), st));
    CK(cudaMemcpyAsync(a->ml_est, t->ml_est, n_ml * sizeof(int16_t), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(a->cfg_rate, t->cfg_rate, 3 * NBIOT_MB_K * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(a->cfg_conf, t->cfg_conf, 3 * NBIOT_MB_K * 2, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(a->rate_conf, t->rate_conf, NBIOT_MB_RATE_ROWS * 8, cudaMemcpyHostToDevice, st));
    NbMbTablesDev &T = a->T;
    T.ml_est = a->ml_est; T.cfg_rate = a->cfg_rate; T.cfg_conf = a->cfg_conf; T.rate_conf = a->rate_conf;
    for (int k = 0; k < 3; ++k) T.cfg_n[k] = t->cfg_n[k];
    for (int k = 0; k < 6; ++k) { T.period_list[k] = t->period_list[k]; if (T.period_list[k] <= 0) return set_err(NBIOT_E_INVALID, "nbiot_mb_agent_create: bad period list"); }
    for (int k = 0; k < 15; ++k) { T.m3_det[k] = t->m3_det[k]; T.m3_ce[k] = t->m3_ce[k]; if (T.m3_ce[k] > 2) return set_err(NBIOT_E_INVALID, "nbiot_mb_agent_create: bad CE class"); }
    T.th_default[0] = t->th_default[0]; T.th_default[1] = t->th_default[1];
    T.min_value = t->min_value; T.max_value = t->max_value; T.step = t->step;
    T.buffer_size = buffer_size; T.no_th = no_th ? 1 : 0;
    nbiot_mb_init_kernel<<<(env->n_inst + 127) / 128, 128, 0, st>>>(a->state, env->n_inst, margin, T.th_default[0], T.th_default[1]);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));                              /* the host tables may go away after this call */
    env->launches += 1;
    *out = a;
    return 0;
}

int nbiot_mb_agent_destroy(nbiot_mb_agent_t *a)
{
    if (!a) return 0;
    cudaSetDevice(a->env->device);
    cudaFree(a->ml_est); cudaFree(a->cfg_rate); cudaFree(a->cfg_conf); cudaFree(a->rate_conf); cudaFree(a->faults);
    delete a;
    return 0;
}

int nbiot_mb_agent_act(nbiot_mb_agent_t *a, const double *margin_dev, uint8_t *actions_dev, double *samples_dev, void *stream)
{
    if (!a || !actions_dev) return set_err(NBIOT_E_INVALID, "nbiot_mb_agent_act: null argument");
    nbiot_handle_t *h = a->env;
    CK(cudaSetDevice(h->device));
    nbiot_mb_act_kernel<<<(h->n_inst + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        a->T, (const nbiot_record_t *)(h->state + h->L.off_rec), (const int32_t *)(h->state + h->L.off_service), h->L.hist_cap,
        a->state, h->n_inst, margin_dev, actions_dev, samples_dev, a->faults);
    CK(cudaGetLastError());
    h->launches += 1;
    return 0;
}

int nbiot_mb_agent_faults(nbiot_mb_agent_t *a, long long *count)
{
    if (!a || !count) return set_err(NBIOT_E_INVALID, "nbiot_mb_agent_faults: null argument");
    CK(cudaSetDevice(a->env->device));
    CK(cudaDeviceSynchronize());
    unsigned long long v = 0;
    CK(cudaMemcpy(&v, a->faults, sizeof v, cudaMemcpyDeviceToHost));
    *count = (long long)v;
    return 0;
}

}   /* extern "C" */

#endif /* NBIOT_AGENT_IMPL_CUH */
