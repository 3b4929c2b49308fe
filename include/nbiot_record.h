/*
 * nbiot_record.h -- the fixed-size decision record: everything the reference returns
 * from NodeB.step()/reset() at a decision epoch, i.e. the reward plus the `info` dict
 * assembled by NodeB.get_node_info (reference system/node_b.py:278-418, schema in
 * SURVEY.md App. D). One record per cell instance; the CUDA path, the CPU oracle and
 * the golden fixtures generated from the reference all use this layout.
 *
 * Variable-length info lists are bounded here (capacities below); the true length is
 * always stored, entries beyond the capacity are dropped and NBIOT_FAULT_LIST_TRUNC is
 * raised in `fault` (never silent). The unbounded per-period lists `incoming`, `delays`
 * and `service_times` (node_b.py:403-407) live in a side buffer of `hist_cap` entries
 * per instance; the record carries their true lengths.
 */
#ifndef NBIOT_RECORD_H
#define NBIOT_RECORD_H

#include <stdint.h>

#define NBIOT_MAX_OCC 48    /* NPRACH occasions per CE per 3000-sf update period (<= 3000/80 + 1) */
#define NBIOT_STEP_LIST 16  /* receptions / errors / unfit / access entries per scheduling epoch  */
#define NBIOT_STEP_SPILL 48 /* further entries of each of those lists kept in a side buffer of the state (nbiot_spill_ptr): at least this many,
                               hist_cap when that is larger (info['access'] grows by one per connection for as long as a cell stays in
                               RAR windows: hundreds of entries under RA congestion). A list is only cut -- and NBIOT_FAULT_LIST_TRUNC
                               raised -- beyond NBIOT_STEP_LIST + that */

/* per-instance fault bits: a faulted instance is frozen (time stops advancing) */
#define NBIOT_FAULT_UE_SLOTS 0x1u     /* more live UEs than ue_cap                          */
#define NBIOT_FAULT_LOOKAHEAD 0x2u    /* carrier look-ahead beyond the grid ring (grid_w)    */
#define NBIOT_FAULT_ACTION 0x4u       /* action the reference raises on (sc=2, rar_Imcs>2)   */
#define NBIOT_FAULT_MCS_KEY 0x8u      /* loss 171.4 has no loss_tbs_to_mcs row (App. C #9)   */
#define NBIOT_FAULT_EVENT_SLOTS 0x10u /* more pending system events than slots               */
#define NBIOT_FAULT_LIST_TRUNC 0x20u  /* an info list was truncated (simulation unaffected)  */
#define NBIOT_FAULT_STALL 0x40u       /* an external step took NBIOT_STALL_STEPS NodeB steps without reaching a state of the external
                                         agent (e.g. a UE-selection agent over a cell whose NPRACH configuration connects nobody): the
                                         reference's Controller.to_next_step (controller.py:262-288) loops on for ever; the cell stops */
#define NBIOT_FAULT_FATAL_MASK 0x5Fu
#define NBIOT_STALL_STEPS 65536       /* NodeB steps inside one external step (hundreds to ~1500 in the reference's scenarios) */

typedef struct nbiot_record {
    /* ---- doubles ---- */
    double reward;                 /* perf_monitor.get_reward() of the last NodeB.step        */
    double loss[4];                /* Scheduling: path loss of the selectable UEs             */
    double sinr[4];                /* Scheduling: clip(sinr,-40,30)+40 for retransmissions    */
    double detection_ratios[3];    /* NPRACH_update ...                                       */
    double colision_ratios[3];
    double msg3_detection[3];
    double msg3_sent[3];
    double msg3_received[3];
    double distribution[15];
    double RA_occupation, NPUSCH_occupation, NPRACH_occupation, av_delay;
    /* ---- int32 ---- */
    int32_t time;
    int32_t state;                 /* NB_ST_* */
    int32_t total_ues;
    int32_t fault;
    int32_t n_sel;                 /* len(info['ues'])                                        */
    int32_t ues[4];
    int32_t connection_time[4];
    int32_t buffer[4];
    int32_t n_rx, n_err, n_unfit, n_acc;   /* true lengths of the four per-epoch collections  */
    int32_t rx_id[NBIOT_STEP_LIST], rx_delay[NBIOT_STEP_LIST];      /* info['receptions']      */
    int32_t err_id[NBIOT_STEP_LIST];                                /* info['errors']          */
    int32_t unfit_id[NBIOT_STEP_LIST];                              /* info['unfit']           */
    int32_t acc_id[NBIOT_STEP_LIST], acc_delay[NBIOT_STEP_LIST];    /* info['access']          */
    int32_t ues_per_CE[3], RAR_in[3], RAR_sent[3], RAR_ids[3], RAR_detected[3], RAR_failed[3];
    int32_t valid_CE_control, NPDCCH_sf_left, total_departures;
    int32_t action[3];             /* info['action'] == info['final_action'] (App. C #11)     */
    int32_t nprach_conf[16];       /* NPRACH_items order (parameters.py:136-153)              */
    int32_t preambles[3];
    int32_t departures;            /* len(delays) == len(service_times)                       */
    int32_t n_incoming;            /* len(incoming)                                           */
    int32_t n_occ[3];              /* len(NPRACH_detection[ce])                               */
    int32_t draws_lo, draws_hi;    /* random counters consumed so far (parity aid)            */
    int32_t events;                /* event pops so far (diagnostic; lazily retired events excluded) */
    /* ---- int16 / uint8 ---- */
    int16_t det_hist[3][NBIOT_MAX_OCC];   /* info['NPRACH_detection'] */
    int16_t col_hist[3][NBIOT_MAX_OCC];   /* info['NPRACH_collision'] */
    uint8_t carrier_busy[40];      /* carrier_state[j] = 1 - carrier_busy[j]/12               */
} nbiot_record_t;

#endif /* NBIOT_RECORD_H */
