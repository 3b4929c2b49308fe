/*
 * nbiot_conf.h -- construction-time configuration of one batch of simulated cells.
 * Fields mirror the `conf` dict read by the reference's create_system
 * (system/system_creator.py:24-44) plus the carrier count that the reference keeps in
 * a module global (system/parameters.py:16,96-100) and the fixed capacities that
 * replace the reference's unbounded Python containers.
 */
#ifndef NBIOT_CONF_H
#define NBIOT_CONF_H

#include <stdint.h>

/* the fields every event handler may read: the simulation kernels keep a by-value copy of this prefix in shared memory */
#define NBIOT_CONF_HOT_FIELDS \
    double ratio;             /* conf['ratio']: share of uniform (vs beta-burst) devices       */ \
    double max_d;             /* conf['max_d'] cell range in km; <= 0 means default (10 km)    */ \
    int32_t M;                /* conf['M']: devices per cell                                    */ \
    int32_t buffer_lo;        /* conf['buffer_range'][0]                                        */ \
    int32_t buffer_hi;        /* conf['buffer_range'][1]                                        */ \
    int32_t reward_criteria;  /* NB_RW_* (conf['reward_criteria'])                             */ \
    int32_t sort_by_loss;     /* conf['sort_criterium'] == 'loss' (default 't_connection')      */ \
    int32_t sc_adjustment;    /* default 1 */ \
    int32_t mcs_automatic;    /* default 1 */ \
    int32_t ce_mcs_automatic; /* default 1 */ \
    int32_t tx_all_buffer;    /* default 1 */ \
    int32_t random_traffic;   /* conf['random'] */ \
    int32_t n_carriers;       /* 1 or 2 (parameters.N_carriers) */ \
    /* capacities (CUDA path only; the oracle grows its containers) */ \
    int32_t ue_cap;           /* live-UE slots per instance (<= M suffices: closed population)  */ \
    int32_t grid_w;           /* carrier look-ahead ring, columns, power of two                 */ \
    int32_t hist_cap;         /* entries of incoming/delays/service_times kept per period       */

typedef struct nbiot_conf_hot { NBIOT_CONF_HOT_FIELDS } nbiot_conf_hot_t;

typedef struct nbiot_conf {
    NBIOT_CONF_HOT_FIELDS
    /* ---- read on rare paths only (arrivals, departures) ---- */
    int32_t stat_cap;         /* conf['statistics']: entries of the per-departure histories kept per instance
                                 (PerfMonitor.delay/connection/access/th_history, perf_monitor.py:61-84); 0 = off */
    int32_t clustered;        /* conf['clustered'] (population.py:57-63, :75-101)               */
    int32_t cluster_lo;       /* conf['cluster_ues'][0]                                         */
    int32_t cluster_hi;       /* conf['cluster_ues'][1]                                         */
    double cluster_prob;      /* conf['cluster_prob']                                           */
    double cluster_range;     /* conf['cluster_range'] in km (divided by the cell range when used) */
    int32_t trace_cap;        /* entries of the per-cell sample ring (below) kept per instance; 0 = off                   */
    int32_t trace_mask;       /* which samples are recorded: NBIOT_TR_* bits. conf['traces'] sets the NPRACH and msg3 samples
                                 (PerfMonitor.nprach_sample / rar_sample, perf_monitor.py:61-69, :111-140), conf['statistics'] the
                                 RAR-window samples, the three resource histories and the departures
                                 (rar_window_sample, register_*_resources, npusch_delivered; :71-79, :125-129, :161-202)  */
    int32_t capture_effect;   /* Channel(capture_effect=True) (channel.py:124-126, :158-159, :200-219): a preamble chosen by several UEs
                                 can still be detected (states CAPTURE / COLLIDED), msg3 grants then carry several contenders.
                                 Not reachable through create_system (system_creator.py:51); a constructor argument here.     */
    int32_t kernel;           /* which kernel runs the LONG advances (many NodeB steps per launch; launches of one or two steps always run on the
                                 warp-per-cell kernel): NBIOT_KERNEL_AUTO by batch size (thread-per-cell from 32 768 cells on), or forced.
                                 Results are bit-identical; lightly loaded cells favour the thread-per-cell kernel (2.9e9 against 1.4e9
                                 subframe-instances/s at 131 072 cells of configs[0] traffic), cells with hundreds of live UEs the
                                 warp-per-cell one (1.01e9 against 0.87e9 at 65 536 cells of the configs[1] scenario)                 */
} nbiot_conf_t;
enum { NBIOT_KERNEL_AUTO = 0, NBIOT_KERNEL_WARP = 1, NBIOT_KERNEL_THREAD = 2, NBIOT_KERNEL_THREAD_VOTED = 3 };

/* one sample of the ring (16 bytes), in the order the reference appends to its lists */
enum { NBIOT_TR_NPRACH = 1, NBIOT_TR_MSG3 = 2, NBIOT_TR_RARWIN = 4, NBIOT_TR_RES_MSG3 = 8, NBIOT_TR_RES_NPUSCH = 16, NBIOT_TR_RES_NPRACH = 32,
       NBIOT_TR_DEPART = 64 };
typedef struct nbiot_trace_rec {
    int32_t t;                /* time of the sample                                                                         */
    int32_t a, b;             /* NPRACH: n_ues, detections | MSG3: detection, error + 2 * collision | RARWIN: RAR_in + (RAR_attmp << 16),
                                 RAR_sent + (RAR_ids << 16) | RES_*: subcarriers x subframes, 0 | DEPART: 0, 0               */
    uint8_t type, ce;         /* NBIOT_TR_* bit, CE level                                                                   */
    uint16_t pad;
} nbiot_trace_rec_t;

#endif /* NBIOT_CONF_H */
