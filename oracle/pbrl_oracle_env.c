/* CPU oracle, batched env layer -- TEST INFRASTRUCTURE (see pbrl_oracle.h).
 *
 * Restates the env glue of the reference on top of pbrl_oracle.c:
 *   reset   /root/reference/agents_env/agent_env_template.py:88-112  (qpos0, zero qvel, pipeline_init)
 *   step    /root/reference/agents_env/agent_mimic.py:58-91 and variants :94-271
 *           (PD law once per env step, then n_frames x mjx.step with the torque held as ctrl)
 *   replay  /root/reference/agents_env/agent_replay_motion.py:51-83  (pipeline_init at given qpos/qvel)
 *   obs     /root/reference/agents_env/agent_env_template.py:151-155 ([qpos, qvel])
 * The brax `pipeline_state` fields the PD laws read (qM, qfrc_bias) belong to the LAST forward,
 * i.e. to the state that entered the last substep.  This layer keeps that
 * state as (qpos_lag, qvel_lag) and re-evaluates M / bias from it -- identical values, 276 B of
 * state instead of 4.8 KB.
 *
 * Arrays are env-major rows of REAL with the logical widths nq, nv, nu (no padding).
 */
#include "pbrl_oracle.h"

#include <math.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    REAL *qpos, *qvel, *time, *warm, *ctrl, *qpos_lag, *qvel_lag; /* state, updated in place */
    REAL *obs, *reward, *done, *metrics;                          /* outputs */
} OrcEnvBuffers;

static void load_env(const OrcModel* m, const OrcEnvBuffers* s, int e, OrcData* d) {
    int nq = m->nq[0], nv = m->nv[0], nu = m->nu[0];
    memset(d, 0, sizeof(*d));
    memcpy(d->qpos, s->qpos + (size_t)e * nq, nq * sizeof(REAL));
    memcpy(d->qvel, s->qvel + (size_t)e * nv, nv * sizeof(REAL));
    memcpy(d->qacc_warmstart, s->warm + (size_t)e * nv, nv * sizeof(REAL));
    memcpy(d->ctrl, s->ctrl + (size_t)e * nu, nu * sizeof(REAL));
    d->time[0] = s->time[e];
}
static void write_obs(const OrcModel* m, const OrcEnvBuffers* s, int e, const OrcData* d) {
    int nq = m->nq[0], nv = m->nv[0];
    if (!s->obs) return;
    memcpy(s->obs + (size_t)e * (nq + nv), d->qpos, nq * sizeof(REAL));
    memcpy(s->obs + (size_t)e * (nq + nv) + nq, d->qvel, nv * sizeof(REAL));
}

void orc_env_reset(const OrcModel* m, int N, OrcEnvBuffers* s) {
    int nq = m->nq[0], nv = m->nv[0], nu = m->nu[0];
    for (int e = 0; e < N; e++) {
        OrcData d;
        orc_reset_data(m, &d);
        memcpy(s->qpos + (size_t)e * nq, d.qpos, nq * sizeof(REAL));
        memcpy(s->qpos_lag + (size_t)e * nq, d.qpos, nq * sizeof(REAL));
        memset(s->qvel + (size_t)e * nv, 0, nv * sizeof(REAL));
        memset(s->qvel_lag + (size_t)e * nv, 0, nv * sizeof(REAL));
        memset(s->warm + (size_t)e * nv, 0, nv * sizeof(REAL));
        memset(s->ctrl + (size_t)e * nu, 0, nu * sizeof(REAL));
        s->time[e] = 0;
        write_obs(m, s, e, &d);
        if (s->reward) for (int k = 0; k < 3; k++) s->reward[3 * e + k] = 0;
        if (s->done) s->done[e] = 0;
        if (s->metrics) for (int k = 0; k < 3; k++) s->metrics[3 * e + k] = 0;
    }
}

/* pd_mode: 0 action is the torque, 1 standard, 2 feedback, 3 stable.  pd_time NULL -> state time.
 * ctrl_mask bit a set -> actuator a takes the new torque, else keeps the previous ctrl
 * (agent_mimic.py:122-125,173-179). */
void orc_env_step(const OrcModel* m, int N, int n_frames, int pd_mode, int action_cols, const REAL* action,
                  const REAL* kp, const REAL* kd, REAL pd_dt, const REAL* pd_time, unsigned ctrl_mask,
                  OrcEnvBuffers* s, int nthreads) {
    int nq = m->nq[0], nv = m->nv[0], nu = m->nu[0];
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    (void)nthreads;
#pragma omp parallel for schedule(static)
    for (int e = 0; e < N; e++) {
        OrcData d, lag;
        REAL tau[ORC_MAXU];
        load_env(m, s, e, &d);
        if (pd_mode >= 2) { /* the PD law reads qM / qfrc_bias of the last forward */
            memset(&lag, 0, sizeof(lag));
            memcpy(lag.qpos, s->qpos_lag + (size_t)e * nq, nq * sizeof(REAL));
            memcpy(lag.qvel, s->qvel_lag + (size_t)e * nv, nv * sizeof(REAL));
            orc_forward(m, &lag);
        }
        orc_pd_torque(m, &lag, pd_mode, action_cols, action + (size_t)e * nu * action_cols, d.qpos, d.qvel, kp, kd,
                      pd_time ? pd_time[e] : d.time[0], pd_dt, tau);
        for (int a = 0; a < nu; a++) if ((ctrl_mask >> a) & 1u) d.ctrl[a] = tau[a];
        for (int f = 0; f < n_frames; f++) {
            if (f == n_frames - 1) {
                memcpy(s->qpos_lag + (size_t)e * nq, d.qpos, nq * sizeof(REAL));
                memcpy(s->qvel_lag + (size_t)e * nv, d.qvel, nv * sizeof(REAL));
            }
            orc_step(m, &d);
        }
        memcpy(s->qpos + (size_t)e * nq, d.qpos, nq * sizeof(REAL));
        memcpy(s->qvel + (size_t)e * nv, d.qvel, nv * sizeof(REAL));
        memcpy(s->warm + (size_t)e * nv, d.qacc_warmstart, nv * sizeof(REAL));
        memcpy(s->ctrl + (size_t)e * nu, d.ctrl, nu * sizeof(REAL));
        s->time[e] = d.time[0];
        write_obs(m, s, e, &d);
        if (s->reward) for (int k = 0; k < 3; k++) s->reward[3 * e + k] = 0; /* agent_mimic.py:84 */
        if (s->done) s->done[e] = 1;                                           /* agent_mimic.py:72 */
    }
}

/* replay: pipeline_init at the given qpos/qvel; returns obs and, optionally, body poses */
void orc_env_replay(const OrcModel* m, int N, const REAL* qpos, const REAL* qvel, REAL* obs, REAL* reward,
                    REAL* done, REAL* xpos, REAL* xquat, int nthreads) {
    int nq = m->nq[0], nv = m->nv[0], nb = m->nbody[0];
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    (void)nthreads;
#pragma omp parallel for schedule(static)
    for (int e = 0; e < N; e++) {
        OrcData d;
        memset(&d, 0, sizeof(d));
        memcpy(d.qpos, qpos + (size_t)e * nq, nq * sizeof(REAL));
        memcpy(d.qvel, qvel + (size_t)e * nv, nv * sizeof(REAL));
        orc_forward(m, &d);
        if (obs) {
            memcpy(obs + (size_t)e * (nq + nv), d.qpos, nq * sizeof(REAL));
            memcpy(obs + (size_t)e * (nq + nv) + nq, d.qvel, nv * sizeof(REAL));
        }
        if (reward) for (int k = 0; k < 3; k++) reward[3 * e + k] = 0;
        if (done) done[e] = 1;
        if (xpos) memcpy(xpos + (size_t)e * nb * 3, d.xpos, nb * 3 * sizeof(REAL));
        if (xquat) memcpy(xquat + (size_t)e * nb * 4, d.xquat, nb * 4 * sizeof(REAL));
    }
}

/* batched PD law (the separate-controller path of testing_pds_env3.ipynb cell 21); qM_out optional (nv x nv) */
void orc_env_pd_torque(const OrcModel* m, int N, int pd_mode, int action_cols, const REAL* action, const REAL* qpos,
                       const REAL* qvel, const REAL* qpos_lag, const REAL* qvel_lag, const REAL* kp, const REAL* kd,
                       const REAL* time, REAL pd_dt, REAL* tau, REAL* qM_out, REAL* bias_out) {
    int nq = m->nq[0], nv = m->nv[0], nu = m->nu[0];
    for (int e = 0; e < N; e++) {
        OrcData lag;
        memset(&lag, 0, sizeof(lag));
        memcpy(lag.qpos, qpos_lag + (size_t)e * nq, nq * sizeof(REAL));
        memcpy(lag.qvel, qvel_lag + (size_t)e * nv, nv * sizeof(REAL));
        orc_forward(m, &lag);
        orc_pd_torque(m, &lag, pd_mode, action_cols, action + (size_t)e * nu * action_cols, qpos + (size_t)e * nq,
                      qvel + (size_t)e * nv, kp, kd, time[e], pd_dt, tau + (size_t)e * nu);
        if (qM_out)
            for (int i = 0; i < nv; i++)
                for (int j = 0; j < nv; j++) qM_out[((size_t)e * nv + i) * nv + j] = lag.qM[i * ORC_MAXV + j];
        if (bias_out) memcpy(bias_out + (size_t)e * nv, lag.qfrc_bias, nv * sizeof(REAL));
    }
}
