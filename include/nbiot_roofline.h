/*
 * nbiot_roofline.h -- the frozen constants of the algorithmic-bytes model:
 *
 *   B_alg = 2*H_HOT + 2*B_COL*C + Lbar*U_HOT + rho_ev*(2*H_EVT + 2*K_UE*U_FULL)   bytes per subframe-instance
 *
 * Lbar (time-mean live UEs per cell) and rho_ev (events per subframe) are outputs of the simulation
 * and are measured on the device in every bench run. The model describes a simulator that touches
 * every live UE's hot record once per subframe; bench.py and DESIGN.md both read the numbers from here.
 */
#ifndef NBIOT_ROOFLINE_H
#define NBIOT_ROOFLINE_H

#define NBIOT_RL_H_HOT 128    /* clock / FSM / NPRACH resource records, read + written every subframe */
#define NBIOT_RL_B_COL 2      /* one 12-bit grid column retired + one generated, per carrier          */
#define NBIOT_RL_U_HOT 12     /* state / list / CE / wake / order stamp of one live UE                 */
#define NBIOT_RL_U_FULL 72    /* a full UE record                                                      */
#define NBIOT_RL_H_EVT 512    /* cold header part touched when an event fires                          */
#define NBIOT_RL_K_UE 2       /* UEs touched per event                                                 */

#endif /* NBIOT_ROOFLINE_H */
