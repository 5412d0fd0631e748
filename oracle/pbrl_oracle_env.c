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

/* ======================================================================================================
 * Mimic mode (build-defined; SURVEY.md App. A.11, "parity unpinned": the reference's reward is zeros(3)).
 * Spec shared with the CUDA epilogue (physics-based-rl_b200/csrc/pbrl_mimic.cuh), restated here densely:
 *   reference pose at clip time t: frame pair (i, i+1), blend u; root position lerp (+ per-cycle xy offset for
 *   wrapped clips), root and 3-dof joint quaternions slerp, 1-dof angles and qvel lerp;
 *   reward = w_p exp(-k_p e_p) + w_v exp(-k_v e_v) + w_e exp(-k_e e_e) + w_c exp(-k_c e_c);
 *   done = fall (listed geoms penetrating the floor, if enabled for the clip) or end of a non-wrapping clip.
 * ==================================================================================================== */
#include "../include/pbrl.h"

static void q_mul(const REAL* a, const REAL* b, REAL* r) {
    REAL w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
    REAL x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
    REAL y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
    REAL z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
    r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
static void q_rot(const REAL* q, const REAL* v, REAL* r) { /* R(q) v */
    REAL w = q[0], x = q[1], y = q[2], z = q[3];
    REAL m[9] = {w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y),
                 2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x),
                 2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z};
    REAL a = m[0] * v[0] + m[1] * v[1] + m[2] * v[2], b = m[3] * v[0] + m[4] * v[1] + m[5] * v[2], c = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
    r[0] = a; r[1] = b; r[2] = c;
}
static void q_slerp(const REAL* a, const REAL* b, REAL u, REAL* r) {
    REAL d = a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3], sg = 1, wa, wb, n = 0;
    if (d < 0) { d = -d; sg = -1; }
    if (d > (REAL)0.9995) { wa = 1 - u; wb = u; }
    else {
        REAL th = (REAL)acos((double)d), s = (REAL)sin((double)th);
        wa = (REAL)sin((double)((1 - u) * th)) / s; wb = (REAL)sin((double)(u * th)) / s;
    }
    for (int k = 0; k < 4; k++) { r[k] = wa * a[k] + sg * wb * b[k]; n += r[k] * r[k]; }
    n = (REAL)sqrt((double)n);
    for (int k = 0; k < 4; k++) r[k] /= n;
}
static REAL q_angle(const REAL* a, const REAL* b) { /* geodesic angle between unit quaternions */
    REAL d = a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3];
    d = d < 0 ? -d : d;
    if (d > 1) d = 1;
    return 2 * (REAL)acos((double)d);
}
static void q_to_euler_xyz(const REAL* q, REAL* e) { /* intrinsic X-Y-Z, as euler_from_quaternion(..., 'rxyz') */
    REAL w = q[0], x = q[1], y = q[2], z = q[3];
    REAL m00 = 1 - 2 * (y * y + z * z), m01 = 2 * (x * y - z * w), m02 = 2 * (x * z + y * w);
    REAL m12 = 2 * (y * z - x * w), m22 = 1 - 2 * (x * x + y * y);
    REAL cy = (REAL)sqrt((double)(m22 * m22 + m12 * m12));
    e[0] = (REAL)atan2((double)-m12, (double)m22);
    e[1] = (REAL)atan2((double)m02, (double)cy);
    e[2] = (REAL)atan2((double)-m01, (double)m00);
}

typedef struct {
    REAL root_pos[3], root_quat[4], jquat[8][4], hinge[ORC_MAXU], qvel[ORC_MAXV];
    REAL qpos[ORC_MAXQ];                   /* MuJoCo layout, Euler angles from the interpolated quaternions */
    REAL xpos[ORC_MAXBODY][3], xquat[ORC_MAXBODY][4], com[3], ee[4][3];
} OrcRefPose;

static void mimic_reference(const OrcModel* m, const PbrlMimicDesc* mm, int clip, REAL t, OrcRefPose* r) {
    int nq = m->nq[0], nv = m->nv[0], nb = m->nbody[0];
    int first = mm->clip_first[clip], F = mm->clip_nframes[clip];
    REAL T = (REAL)mm->clip_duration[clip], dt = (REAL)mm->clip_dt[clip], cyc = 0, tc = t;
    if (mm->clip_wrap[clip]) { cyc = (REAL)floor((double)(t / T)); tc = t - cyc * T; }
    else if (tc > T) tc = T;
    if (tc < 0) tc = 0;
    int fi = (int)(tc / dt);
    if (fi > F - 2) fi = F - 2;
    REAL u = tc / dt - (REAL)fi;
    if (u < 0) u = 0;
    if (u > 1) u = 1;
    const float* f0 = mm->frames + (size_t)(first + fi) * PBRL_FRAME_STRIDE;
    const float* f1 = f0 + PBRL_FRAME_STRIDE;
    const float* fa = mm->frames + (size_t)first * PBRL_FRAME_STRIDE;
    const float* fz = mm->frames + (size_t)(first + F - 1) * PBRL_FRAME_STRIDE;
    for (int k = 0; k < 3; k++) r->root_pos[k] = (1 - u) * (REAL)f0[k] + u * (REAL)f1[k];
    for (int k = 0; k < 2; k++) r->root_pos[k] += cyc * ((REAL)fz[k] - (REAL)fa[k]);
    REAL qa[4], qb[4];
    for (int k = 0; k < 4; k++) { qa[k] = (REAL)f0[3 + k]; qb[k] = (REAL)f1[3 + k]; }
    q_slerp(qa, qb, u, r->root_quat);
    for (int j = 0; j < 8; j++) {
        for (int k = 0; k < 4; k++) { qa[k] = (REAL)f0[69 + 4 * j + k]; qb[k] = (REAL)f1[69 + 4 * j + k]; }
        q_slerp(qa, qb, u, r->jquat[j]);
    }
    for (int i = 0; i < nv; i++) r->qvel[i] = (1 - u) * (REAL)f0[35 + i] + u * (REAL)f1[35 + i];
    /* qpos: root, then per body either Euler angles of its quaternion or the lerped hinge angle */
    for (int k = 0; k < 3; k++) r->qpos[k] = r->root_pos[k];
    for (int k = 0; k < 4; k++) r->qpos[3 + k] = r->root_quat[k];
    int slot = 0;
    memset(r->xpos, 0, sizeof(r->xpos));
    memset(r->xquat, 0, sizeof(r->xquat));
    r->xquat[0][0] = 1;
    for (int b = 1; b < nb; b++) {
        int p = m->body_parent[b], ja = m->body_jntadr[b], qa0 = m->jnt_qposadr[ja];
        REAL bp[3] = {(REAL)m->body_pos[3 * b], (REAL)m->body_pos[3 * b + 1], (REAL)m->body_pos[3 * b + 2]}, t3[3], ql[4];
        if (m->jnt_type[ja] == 0) {
            for (int k = 0; k < 3; k++) r->xpos[b][k] = r->root_pos[k];
            for (int k = 0; k < 4; k++) r->xquat[b][k] = r->root_quat[k];
            continue;
        }
        q_rot(r->xquat[p], bp, t3);
        for (int k = 0; k < 3; k++) r->xpos[b][k] = r->xpos[p][k] + t3[k];
        if (m->body_jntnum[b] == 3) {
            q_to_euler_xyz(r->jquat[slot], r->qpos + qa0);
            for (int k = 0; k < 4; k++) ql[k] = r->jquat[slot][k];
            slot++;
        } else {
            REAL th = (1 - u) * (REAL)f0[qa0] + u * (REAL)f1[qa0], ax[3] = {(REAL)m->jnt_axis[3 * ja], (REAL)m->jnt_axis[3 * ja + 1], (REAL)m->jnt_axis[3 * ja + 2]};
            r->qpos[qa0] = th;
            r->hinge[b] = th;
            ql[0] = (REAL)cos(0.5 * (double)th);
            for (int k = 0; k < 3; k++) ql[1 + k] = ax[k] * (REAL)sin(0.5 * (double)th);
        }
        q_mul(r->xquat[p], ql, r->xquat[b]);
    }
    (void)nq;
    REAL mt = 0;
    r->com[0] = r->com[1] = r->com[2] = 0;
    for (int b = 1; b < nb; b++) {
        REAL ip[3] = {(REAL)m->body_ipos[3 * b], (REAL)m->body_ipos[3 * b + 1], (REAL)m->body_ipos[3 * b + 2]}, t3[3], ms = (REAL)m->body_mass[b];
        q_rot(r->xquat[b], ip, t3);
        for (int k = 0; k < 3; k++) r->com[k] += ms * (r->xpos[b][k] + t3[k]);
        mt += ms;
    }
    for (int k = 0; k < 3; k++) r->com[k] /= mt;
    for (int e = 0; e < mm->n_ee; e++) {
        REAL off[3] = {(REAL)mm->ee_offset[3 * e], (REAL)mm->ee_offset[3 * e + 1], (REAL)mm->ee_offset[3 * e + 2]}, t3[3];
        q_rot(r->xquat[mm->ee_body[e]], off, t3);
        for (int k = 0; k < 3; k++) r->ee[e][k] = r->xpos[mm->ee_body[e]][k] + t3[k];
    }
}

/* reward / done / metrics for the state in d (kinematics are re-evaluated here) at clip time t */
static void mimic_epilogue(const OrcModel* m, const PbrlMimicDesc* mm, OrcData* d, int clip, REAL t, REAL* reward, REAL* done, REAL* metrics) {
    int nb = m->nbody[0], nv = m->nv[0];
    OrcRefPose r;
    mimic_reference(m, mm, clip, t, &r);
    orc_forward(m, d); /* kinematics, com and floor distances of the final state */
    REAL e_p = 0, e_v = 0, e_e = 0, e_c = 0, a;
    a = q_angle(r.root_quat, d->xquat + 4);
    e_p += a * a;
    int slot = 0;
    for (int b = 2; b < nb; b++) {
        int p = m->body_parent[b], ja = m->body_jntadr[b];
        if (m->body_jntnum[b] == 3) {
            REAL pc[4] = {d->xquat[4 * p], -d->xquat[4 * p + 1], -d->xquat[4 * p + 2], -d->xquat[4 * p + 3]}, rel[4];
            q_mul(pc, d->xquat + 4 * b, rel);
            a = q_angle(r.jquat[slot], rel);
            slot++;
        } else {
            a = r.hinge[b] - d->qpos[m->jnt_qposadr[ja]];
        }
        e_p += a * a;
    }
    for (int i = 3; i < nv; i++) { a = r.qvel[i] - d->qvel[i]; e_v += a * a; }
    for (int e = 0; e < mm->n_ee; e++) {
        int b = mm->ee_body[e];
        REAL off[3] = {(REAL)mm->ee_offset[3 * e], (REAL)mm->ee_offset[3 * e + 1], (REAL)mm->ee_offset[3 * e + 2]}, t3[3];
        q_rot(d->xquat + 4 * b, off, t3);
        for (int k = 0; k < 3; k++) { a = r.ee[e][k] - (d->xpos[3 * b + k] + t3[k]); e_e += a * a; }
    }
    for (int k = 0; k < 3; k++) { a = r.com[k] - d->com[k]; e_c += a * a; }
    REAL rp = (REAL)exp(-(double)(mm->k_pose * e_p)), rv = (REAL)exp(-(double)(mm->k_vel * e_v));
    REAL re = (REAL)exp(-(double)(mm->k_ee * e_e)), rc = (REAL)exp(-(double)(mm->k_com * e_c));
    reward[0] = mm->w_pose * rp + mm->w_vel * rv + mm->w_ee * re + mm->w_com * rc;
    reward[1] = rp;
    reward[2] = re;
    int fall = 0, nc = (int)d->ncon[0];
    if (mm->clip_fall_enabled[clip])
        for (int c = 0; c < nc; c++)
            if (m->geom_type[(int)d->con_geom1[c]] == 0 /* floor */ && ((mm->fall_geom_mask >> (int)d->con_geom[c]) & 1u) && d->con_dist[c] < 0) fall = 1;
    int end = !mm->clip_wrap[clip] && t >= (REAL)mm->clip_duration[clip];
    *done = (fall || end) ? 1 : 0;
    metrics[0] += 1;
    metrics[1] = e_p;
    metrics[2] = (REAL)fall;
}

/* mimic reset: reference-state initialisation at clip time ref_t0 */
void orc_env_reset_mimic(const OrcModel* m, const PbrlMimicDesc* mm, int N, const int* motion_id, const REAL* ref_t0, OrcEnvBuffers* s) {
    int nq = m->nq[0], nv = m->nv[0], nu = m->nu[0];
    for (int e = 0; e < N; e++) {
        OrcRefPose r;
        OrcData d;
        mimic_reference(m, mm, motion_id[e], ref_t0[e], &r);
        memset(&d, 0, sizeof(d));
        memcpy(d.qpos, r.qpos, nq * sizeof(REAL));
        memcpy(d.qvel, r.qvel, nv * sizeof(REAL));
        memcpy(s->qpos + (size_t)e * nq, d.qpos, nq * sizeof(REAL));
        memcpy(s->qpos_lag + (size_t)e * nq, d.qpos, nq * sizeof(REAL));
        memcpy(s->qvel + (size_t)e * nv, d.qvel, nv * sizeof(REAL));
        memcpy(s->qvel_lag + (size_t)e * nv, d.qvel, nv * sizeof(REAL));
        memset(s->warm + (size_t)e * nv, 0, nv * sizeof(REAL));
        memset(s->ctrl + (size_t)e * nu, 0, nu * sizeof(REAL));
        s->time[e] = 0;
        write_obs(m, s, e, &d);
        if (s->reward) for (int k = 0; k < 3; k++) s->reward[3 * e + k] = 0;
        if (s->done) s->done[e] = 0;
        if (s->metrics) for (int k = 0; k < 3; k++) s->metrics[3 * e + k] = 0;
    }
}

/* mimic step: orc_env_step + the DeepMimic epilogue; stats[8] accumulates sum r, sum r^2, sum done, sum fall, sum e_p, n */
void orc_env_step_mimic(const OrcModel* m, const PbrlMimicDesc* mm, int N, int n_frames, int pd_mode, int action_cols,
                        const REAL* action, const REAL* kp, const REAL* kd, REAL pd_dt, const int* motion_id,
                        const REAL* ref_t0, OrcEnvBuffers* s, double* stats, int nthreads) {
    int nq = m->nq[0], nv = m->nv[0];
    orc_env_step(m, N, n_frames, pd_mode, action_cols, action, kp, kd, pd_dt, 0, 0x0FFFFFFFu, s, nthreads);
    for (int e = 0; e < N; e++) {
        OrcData d;
        memset(&d, 0, sizeof(d));
        memcpy(d.qpos, s->qpos + (size_t)e * nq, nq * sizeof(REAL));
        memcpy(d.qvel, s->qvel + (size_t)e * nv, nv * sizeof(REAL));
        REAL done;
        mimic_epilogue(m, mm, &d, motion_id[e], ref_t0[e] + s->time[e], s->reward + 3 * e, &done, s->metrics + 3 * e);
        s->done[e] = done;
        if (stats) {
            stats[0] += s->reward[3 * e]; stats[1] += s->reward[3 * e] * s->reward[3 * e]; stats[2] += done;
            stats[3] += s->metrics[3 * e + 2]; stats[4] += s->metrics[3 * e + 1]; stats[5] += 1;
        }
    }
}

/* interpolated reference pose, for tests */
void orc_mimic_reference_pose(const OrcModel* m, const PbrlMimicDesc* mm, int clip, REAL t, REAL* qpos, REAL* qvel, REAL* com, REAL* ee) {
    OrcRefPose r;
    mimic_reference(m, mm, clip, t, &r);
    memcpy(qpos, r.qpos, m->nq[0] * sizeof(REAL));
    memcpy(qvel, r.qvel, m->nv[0] * sizeof(REAL));
    memcpy(com, r.com, 3 * sizeof(REAL));
    memcpy(ee, r.ee, 12 * sizeof(REAL));
}
