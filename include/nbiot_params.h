/*
 * nbiot_params.h -- constants and 3GPP index tables of the simulated NB-IoT cell.
 * Values restate reference system/parameters.py (line numbers cited per block),
 * system/carrier.py:22-41 and system/channel.py:24-52. Shared by the CUDA path and
 * the CPU oracle (they are data, not algorithm).
 */
#ifndef NBIOT_PARAMS_H
#define NBIOT_PARAMS_H

#include <stdint.h>

#if defined(__CUDACC__) && !defined(NBIOT_HOST_TABLES)
#define NB_CONST static __constant__ const
#else
#define NB_CONST static const
#endif

/* parameters.py:14-18 */
#define NB_N_USERS 4
#define NB_HORIZON 40
#define NB_MSG3_TBS 88
/* parameters.py:24-31, channel.py:33 (n_ = 10**((F+N)/10), value printed by CPython) */
#define NB_NOISE_DBM (-121.4)
#define NB_TX_PW 23.0
#define NB_GMAX 15.0
#define NB_SIGMA_F 8.0
#define NB_MAX_RANGE 10.0
#define NB_NOISE_FIG 3.0
#define NB_NOISE_MW 0x1.96dae986609a5p-40
#define NB_SRP 35.0                        /* parameters.py:36 */
/* parameters.py:59,251-257 */
#define NB_NPDCCH_SF 8
#define NB_NPDCCH_PERIOD 30
#define NB_NPRACH_UPDATE_PERIOD 3000
/* ue_generator.py:17-18 */
#define NB_T_BETA 10000
#define NB_T_UNIFORM 60000
#define NB_N_ACTIONS 23                    /* parameters.py:187 */

/* control_items: index of each item in the shared 23-int action (parameters.py:159-185) */
enum {
    NB_A_CARRIER = 0, NB_A_ID = 1, NB_A_IMCS = 2, NB_A_NRU = 3, NB_A_DELAY = 4, NB_A_SC = 5,
    NB_A_NREP = 6, NB_A_BACKOFF = 7, NB_A_RAR_WINDOW = 8, NB_A_MAC_TIMER = 9, NB_A_TRANSMAX = 10,
    NB_A_PANCHOR = 11, NB_A_PERIOD_C0 = 12, NB_A_REP_C0 = 13, NB_A_SC_C0 = 14, NB_A_PERIOD_C1 = 15,
    NB_A_REP_C1 = 16, NB_A_SC_C1 = 17, NB_A_PERIOD_C2 = 18, NB_A_REP_C2 = 19, NB_A_SC_C2 = 20,
    NB_A_TH_C1 = 21, NB_A_TH_C0 = 22
};

/* NodeB decision states (node_b.py:28); numbering is this project's */
enum { NB_ST_SCHEDULING = 0, NB_ST_NPRACH_UPDATE = 1, NB_ST_RAR_WINDOW = 2, NB_ST_RAR_WINDOW_END = 3, NB_ST_IDLE = 4 };

/* UE states (user.py:19) */
enum { NB_UE_RAO = 0, NB_UE_BACKOFF = 1, NB_UE_RA = 2, NB_UE_RAR = 3, NB_UE_CAPTURE = 4, NB_UE_COLLIDED = 5,
       NB_UE_CONREQ = 6, NB_UE_CONNECTED = 7, NB_UE_DATA = 8, NB_UE_DONE = 9 };

/* reward criteria (perf_monitor.py:29-38) */
enum { NB_RW_THROUGHPUT = 0, NB_RW_AVERAGE_DELAY = 1, NB_RW_ACCUMULATED_DELAY = 2, NB_RW_USERS = 3,
       NB_RW_INVD_USERS = 4, NB_RW_LOG_INVD_USERS = 5, NB_RW_AVERAGE_TOTAL_DELAY = 6 };

/* parameters.py:37 ; entry 14 (0 dBm) switches a level off */
NB_CONST int16_t nb_threshold_list[15] = { -115, -114, -113, -112, -111, -110, -109, -108, -106, -104, -102, -100, -98, -96, 0 };
#define NB_MIN_TH (-115)
#define NB_NO_REP_TH (-109)                /* parameters.py:71 */
#define NB_NO_REP_MSG3_TH (-100)           /* parameters.py:93 */
/* thr_to_n_rep, keys -116..-110 (parameters.py:61-69): index th+116 */
NB_CONST uint8_t nb_thr_to_n_rep[7] = { 3, 3, 3, 2, 2, 1, 1 };
/* msg_3_conf N_rep, keys -116..-101 (parameters.py:74-91), I_tbs=2, N_ru=1 throughout: index th+116 */
NB_CONST uint8_t nb_msg3_conf_nrep[16] = { 16, 16, 16, 16, 16, 8, 8, 8, 8, 4, 4, 4, 4, 2, 2, 2 };

NB_CONST uint8_t nb_npdcch_sf_per_ce[3] = { 1, 2, 4 };                       /* parameters.py:252 */
NB_CONST uint8_t nb_ce_for_sf_left[9] = { 0, 0, 1, 1, 2, 2, 2, 2, 2 };      /* parameters.py:253 */

NB_CONST uint8_t nb_imcs_to_itbs[7] = { 0, 1, 2, 3, 4, 6, 8 };              /* parameters.py:276 */
NB_CONST uint8_t nb_itbs_to_row[9] = { 0, 1, 2, 3, 4, 255, 5, 255, 6 };     /* parameters.py:277 (Itbs_to_Imcs) */
NB_CONST uint8_t nb_rar_imcs_to_nru[3] = { 4, 3, 1 };                       /* parameters.py:281 */
NB_CONST uint8_t nb_n_rep_list[8] = { 1, 2, 4, 8, 16, 32, 64, 128 };        /* parameters.py:284 */
NB_CONST uint8_t nb_n_ru_list[8] = { 1, 2, 3, 4, 5, 6, 8, 10 };             /* parameters.py:285 */
NB_CONST uint8_t nb_sf_delay_list[4] = { 8, 16, 32, 64 };                   /* parameters.py:286 */
/* effective N_sc_list after the second assignment (parameters.py:340): sc=2 -> 9 is invalid */
NB_CONST uint8_t nb_n_sc_list[4] = { 3, 6, 9, 12 };
/* N_sc_selection_list (parameters.py:260-265), rows for N_sc = 1,3,6,12 */
NB_CONST uint8_t nb_sc_fallback[4][3] = { { 12, 6, 3 }, { 12, 6, 1 }, { 12, 3, 1 }, { 6, 3, 1 } };
/* lowest subcarrier of every aligned group of 1 / 3 / 6 / 12 subcarriers (offsets_per_sc, carrier.py:29-34), as a 12-bit mask */
NB_CONST uint16_t nb_sc_heads[4] = { 0xFFFu, 0x249u, 0x041u, 0x001u };

/* TBS_table rows for I_tbs = 0,1,2,3,4,6,8 (parameters.py:291-310) */
NB_CONST uint16_t nb_tbs_table[7][8] = {
    { 16, 32, 56, 88, 120, 152, 208, 256 },   { 24, 56, 88, 144, 176, 208, 256, 344 },
    { 32, 72, 144, 176, 208, 256, 328, 424 }, { 40, 104, 176, 208, 256, 328, 440, 568 },
    { 56, 120, 208, 256, 328, 408, 552, 680 }, { 88, 176, 256, 392, 504, 600, 808, 1000 },
    { 120, 256, 392, 536, 680, 808, 1096, 1384 } };
/* TBS_list = sorted set of all entries (parameters.py:300-301); column index of loss_tbs_to_mcs */
NB_CONST uint16_t nb_tbs_list[30] = { 16, 24, 32, 40, 56, 72, 88, 104, 120, 144, 152, 176, 208, 256, 328, 344,
                                      392, 408, 424, 440, 504, 536, 552, 568, 600, 680, 808, 1000, 1096, 1384 };
#define NB_N_TBS 30
#define NB_MCS_LOSS_ROWS 500               /* loss 121.4 .. 171.3 step 0.1 (loss_tbs_to_mcs.p) */

/* parameters.py:318-335 */
NB_CONST int32_t nb_backoff_list[12] = { 0, 256, 512, 1024, 4096, 8192, 16384, 32768, 65536, 131072, 262144, 524288 };
NB_CONST uint8_t nb_rar_window_list[8] = { 2, 3, 4, 5, 6, 7, 8, 10 };
NB_CONST uint8_t nb_mac_timer_list[8] = { 1, 2, 3, 4, 8, 16, 32, 64 };
NB_CONST uint8_t nb_transmax_list[11] = { 3, 4, 5, 6, 7, 8, 10, 20, 50, 100, 200 };
NB_CONST int16_t nb_nprach_period_list[7] = { 80, 160, 240, 320, 640, 1280, 2560 };
/* ceil(5.6 * N_rep) for N_rep = 1..128 (parameters.py:343, carrier.py:22) */
NB_CONST int16_t nb_nprach_nsf_list[8] = { 6, 12, 23, 45, 90, 180, 359, 717 };
/* probability_anchor denominators (parameters.py:332-335): p = den ? 1.0/den : 0 */
NB_CONST uint8_t nb_panchor_den[16] = { 0, 16, 15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2 };

/* NPRACH detection sigmoid (x0, k) per N_rep = 1,2,4,...,128 (channel.py:43-52) */
NB_CONST double nb_nprach_x0[8] = { -26.65585077, -28.25585118, -29.85584994, -31.45585117,
                                    -33.05585139, -34.65584975, -36.25585128, -37.85585128 };
NB_CONST double nb_nprach_k[8] = { 0.32267874, 0.32267874, 0.32267875, 0.32267874,
                                   0.32267859, 0.32267876, 0.3226787, 0.3226787 };

/* NPRACH_items (parameters.py:136-153) -> index in the 23-int action, and their control_default_values */
NB_CONST uint8_t nb_nprach_item_idx[16] = { NB_A_BACKOFF, NB_A_RAR_WINDOW, NB_A_MAC_TIMER, NB_A_TRANSMAX, NB_A_PANCHOR, NB_A_TH_C1, NB_A_TH_C0,
    NB_A_PERIOD_C0, NB_A_REP_C0, NB_A_SC_C0, NB_A_PERIOD_C1, NB_A_REP_C1, NB_A_SC_C1, NB_A_PERIOD_C2, NB_A_REP_C2, NB_A_SC_C2 };
NB_CONST uint8_t nb_nprach_default_conf[16] = { 3, 6, 2, 4, 12, 8, 11, 3, 0, 1, 2, 0, 1, 3, 2, 1 };

/* LUT2.csv geometry (channel.py:61-79): 56 curves (7 I_tbs x 8 N_rep) x 501 SNR points, BLER * 2^15 */
#define NB_LUT_SNR_PTS 501
#define NB_LUT_CURVES 56
#define NB_LUT_SCALE 32768.0

#endif /* NBIOT_PARAMS_H */
