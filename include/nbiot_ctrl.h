/*
 * nbiot_ctrl.h -- the multi-agent controller as data: write tables derived from the reference's agent
 * dictionaries (control_agents.py:29-71) by nb_iot_environment_b200/controller.py and applied to the
 * shared 23-int action vector next to the simulation (reference controller.py:234-288).
 */
#ifndef NBIOT_CTRL_H
#define NBIOT_CTRL_H

#include <stdint.h>

/* internal agents are fixed-action agents (the reference's DummyAgent, control_agents.py:83-113);
 * a write table holds one value per action index, -1 = leave the entry alone */
typedef struct nbiot_ctrl {
    uint32_t ext_state_mask;          /* bit s set: an external decision is needed in NodeB state s */
    int32_t ext_k;                    /* number of action items the external agent sets          */
    uint8_t ext_a_mask[24];           /* their indices in the 23-vector                           */
    int8_t ext_post[24];              /* writes of the `next` chain after the external action     */
    int8_t state_writes[4][24];       /* fixed actions of the internal agents of each NodeB state */
    uint8_t init_action[24];          /* Controller._initialize_action result                     */
    int8_t full_writes[4][24];        /* the same with EVERY agent of the state, the external one's fixed action included:
                                         what Controller.run_until / single_step apply (controller.py:156-195), used by
                                         advances without an external action */
} nbiot_ctrl_t;

#endif /* NBIOT_CTRL_H */
