/*
 * nbiot_agent.h -- C ABI of the model-based NPRACH agent for a batch of cells.
 *
 * Replaces, for n_inst cells at once and without leaving the device, the reference's
 *   NPRACH_THAgent(agent_3, metrics, buffer_size, margin, no_th)      agent_nprach.py:41-59
 *   NPRACH_THAgent.set_parameter(margin)                              agent_nprach.py:59-60
 *   NPRACH_THAgent.get_action(obs, r, info, action) -> th_conf + conf agent_nprach.py:62-146
 * which runs once per NPRACH update (3000 subframes) per cell on the info dict of the NPRACH_update decision
 * (NPRACH_detection / NPRACH_collision per-occasion lists, distribution, incoming) -- here read straight from the
 * decision records of an nbiot_handle_t. The model behind it (ML load estimator over P(n, s, c, r), Pareto rate lists
 * of the configurator, msg3 detection estimate, rate -> configuration table) is passed as flat tables
 * (tools/make_agent_tables.py); the call sequence a maintainer would bind is in INTEGRATION.md.
 *
 * Same conventions as nbiot.h: plain pointers, 0 / negative NBIOT_E_* return codes, stream-ordered calls,
 * caller-owned state (n_inst * sizeof(nbiot_mb_state_t) bytes of device memory: saving it checkpoints the agents).
 */
#ifndef NBIOT_AGENT_H
#define NBIOT_AGENT_H

#include <stdint.h>
#include "nbiot.h"

#ifdef __cplusplus
extern "C" {
#endif

#define NBIOT_MB_D 128          /* ml_est is int16[4][NBIOT_MB_D][NBIOT_MB_C]: detections x collisions per occasion */
#define NBIOT_MB_C 64
#define NBIOT_MB_K 32           /* capacity of one CE level's rate list */
#define NBIOT_MB_RATE_ROWS 186  /* rows of model/rate_to_conf.json */

/* per-cell agent memory (agent_nprach.py:44-57) */
typedef struct nbiot_mb_state {
    double msg3_detection[3];
    double lambda_estimate;     /* read by NPRACH_agent_wrapper as an observation (wrappers.py:311) */
    double margin;              /* beta */
    int32_t n, k;               /* calls so far; loss samples seen while the model is being built */
    uint8_t conf[6], th_conf[2];
} nbiot_mb_state_t;

/* host-side description of the model tables (all pointers are HOST pointers, copied at creation) */
typedef struct nbiot_mb_tables {
    const int16_t *ml_est;      /* [4][NBIOT_MB_D][NBIOT_MB_C]  ml_based_estimator(d, c, N_sc_list[i]) (detection_model.py:25-43) */
    const double *cfg_rate;     /* [3][NBIOT_MB_K] ascending rates of the Pareto lists (configurator.py:104-158), +inf padded */
    const uint8_t *cfg_conf;    /* [3][NBIOT_MB_K][2] (nsc index, period index) */
    int32_t cfg_n[3];
    double m3_det[15];          /* msg3_detection_estimation coefficients (random_access_model.py:86-131) */
    uint8_t m3_ce[15];
    uint8_t pad0;
    const uint8_t *rate_conf;   /* [NBIOT_MB_RATE_ROWS][8] (th_C1 idx, th_C0 idx, nsc0..2, period0..2) */
    double min_value, max_value, step;      /* agent_nprach.py:21-23 */
    int32_t period_list[6];     /* parameters.period_list */
    uint8_t th_default[2];      /* control_default_values th_C1 / th_C0 as indices (agent_nprach.py:29-30) */
    uint8_t pad1[6];
} nbiot_mb_tables_t;

typedef struct nbiot_mb_agent nbiot_mb_agent_t;

/* NPRACH_THAgent.__init__ for every cell of `env`; state_dev: nbiot_mb_state_t[n_inst], initialised here */
int nbiot_mb_agent_create(nbiot_handle_t *env, const nbiot_mb_tables_t *tables, int buffer_size, int no_th, double margin,
                          void *state_dev, void *stream, nbiot_mb_agent_t **out);
int nbiot_mb_agent_destroy(nbiot_mb_agent_t *a);

/* set_parameter + get_action for every cell whose last decision is an NPRACH_update (others are left untouched):
 *   margin_dev  : double[n_inst] or NULL (keep the current margins)
 *   actions_dev : uint8[n_inst][8] = th_C1, th_C0, sc_C0..2, period_C0..2 -- the external action of agent 3, ready for nbiot_advance
 *   samples_dev : double[n_inst][4] or NULL: departures, NPRACH_occupation, mean service time, beta (agent_nprach.py:115-124)
 * A detection / collision count outside the estimator table (cannot happen with <= 2 carriers) is clamped and counted. */
int nbiot_mb_agent_act(nbiot_mb_agent_t *a, const double *margin_dev, uint8_t *actions_dev, double *samples_dev, void *stream);
/* synchronises; *count = table-range violations so far (0 in a healthy run) */
int nbiot_mb_agent_faults(nbiot_mb_agent_t *a, long long *count);

#ifdef __cplusplus
}
#endif

#endif /* NBIOT_AGENT_H */
