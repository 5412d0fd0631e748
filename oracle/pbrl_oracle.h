/* CPU oracle for the humanoid-mimic env step -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this.  It restates, in plain C, the algorithm the reference executes through third-party
 * packages that are absent from /root/reference:
 *     brax==0.10.3 (requirements.txt:5)  envs/base.py PipelineEnv, mjx/pipeline.py init/step
 *     mujoco==3.1.3 (requirements.txt:6) + mujoco-mjx (unpinned): forward.py, smooth.py,
 *         passive.py, collision_primitive.py / collision_convex.py, constraint.py, solver.py
 * anchored on the reference's own call sites (agents_env/agent_mimic.py:58-91,
 * agents_env/agent_env_template.py:88-112, agents_env/agent_replay_motion.py:51-83,
 * agents_env/pds_controllers_agents.py:17-172) and pinned by the notebook goldens T0-T4
 * (tests/golden/mass_traces.npz).  Parity status: contact-free chain pinned by
 * goldens T1-T4 up to print 10 (T1: 13); contact rows -- floor and body-vs-body -- and the mimic reward have no
 * reference known-answer ("parity unpinned"): the hands first touch the thighs at T1 print 14, after the saturated PD
 * loop has already lost fp32 reproducibility (tests/test_self_collision.py).
 *
 * Compiled twice: -DORC_F32 (REAL=float) and default (REAL=double).
 */
#ifndef PBRL_ORACLE_H
#define PBRL_ORACLE_H

#ifdef ORC_F32
typedef float REAL;
#else
typedef double REAL;
#endif

#define ORC_MAXBODY 16
#define ORC_MAXQ 40
#define ORC_MAXV 36
#define ORC_MAXU 32
#define ORC_MAXJNT 32
#define ORC_MAXGEOM 20
#define ORC_MAXPAIR 128
#define ORC_MAXCON 136 /* 29 floor slots + the body-vs-body contacts in range: 74 one-contact pairs, 14 capsule-box pairs x 2, 1 box-box pair x 4 (135 slots) */
#define ORC_MAXEFC (ORC_MAXJNT + 4 * ORC_MAXCON)

/* Model: field names equal the attributes of ModelConst (physics-based-rl_b200/mjcf.py). */
#define ORC_MODEL_FIELDS(I, D)                                                                    \
    I(nbody, 1) I(nq, 1) I(nv, 1) I(nu, 1) I(njnt, 1) I(ngeom, 1) I(iterations, 1)                \
    I(ls_iterations, 1)                                                                            \
    D(timestep, 1) D(gravity, 3) D(solref, 2) D(solimp, 5) D(impratio, 1) D(tolerance, 1)         \
    D(ls_tolerance, 1) D(meaninertia, 1)                                                           \
    I(body_parent, ORC_MAXBODY) D(body_pos, ORC_MAXBODY * 3) D(body_mass, ORC_MAXBODY)            \
    D(body_ipos, ORC_MAXBODY * 3) D(body_inertia, ORC_MAXBODY * 9) I(body_jntadr, ORC_MAXBODY)    \
    I(body_jntnum, ORC_MAXBODY) I(body_dofadr, ORC_MAXBODY) I(body_dofnum, ORC_MAXBODY)           \
    D(body_invweight0, ORC_MAXBODY * 2)                                                            \
    I(jnt_type, ORC_MAXJNT) I(jnt_bodyid, ORC_MAXJNT) I(jnt_qposadr, ORC_MAXJNT)                  \
    I(jnt_dofadr, ORC_MAXJNT) D(jnt_axis, ORC_MAXJNT * 3) D(jnt_pos, ORC_MAXJNT * 3)              \
    I(jnt_limited, ORC_MAXJNT) D(jnt_range, ORC_MAXJNT * 2) D(jnt_margin, ORC_MAXJNT)             \
    I(dof_bodyid, ORC_MAXV) I(dof_parentid, ORC_MAXV) D(dof_armature, ORC_MAXV)                   \
    D(dof_damping, ORC_MAXV) D(dof_invweight0, ORC_MAXV) D(qpos0, ORC_MAXQ)                       \
    I(geom_type, ORC_MAXGEOM) I(geom_bodyid, ORC_MAXGEOM) D(geom_size, ORC_MAXGEOM * 3)           \
    D(geom_pos, ORC_MAXGEOM * 3) D(geom_quat, ORC_MAXGEOM * 4) D(geom_friction, ORC_MAXGEOM * 3)  \
    D(geom_margin, ORC_MAXGEOM) D(geom_gap, ORC_MAXGEOM)                                           \
    I(act_dofadr, ORC_MAXU) D(act_gear, ORC_MAXU) I(act_ctrllimited, ORC_MAXU)                    \
    D(act_ctrlrange, ORC_MAXU * 2)                                                                 \
    I(npair, 1) I(pair_geom1, ORC_MAXPAIR) I(pair_geom2, ORC_MAXPAIR)

typedef struct {
#define ORC_I(name, n) int name[n];
#define ORC_D(name, n) double name[n];
    ORC_MODEL_FIELDS(ORC_I, ORC_D)
#undef ORC_I
#undef ORC_D
} OrcModel;

/* Per-env data: every intermediate of SURVEY.md App. A is exposed for stage-wise parity. */
#define ORC_DATA_FIELDS(X)                                                                         \
    X(qpos, ORC_MAXQ) X(qvel, ORC_MAXV) X(time, 1) X(qacc_warmstart, ORC_MAXV) X(ctrl, ORC_MAXU)   \
    X(xpos, ORC_MAXBODY * 3) X(xquat, ORC_MAXBODY * 4) X(xmat, ORC_MAXBODY * 9)                    \
    X(xipos, ORC_MAXBODY * 3) X(ximat, ORC_MAXBODY * 9) X(xanchor, ORC_MAXJNT * 3)                 \
    X(xaxis, ORC_MAXJNT * 3) X(geom_xpos, ORC_MAXGEOM * 3) X(geom_xmat, ORC_MAXGEOM * 9)           \
    X(subtree_com, 3) X(cinert, ORC_MAXBODY * 10) X(crb, ORC_MAXBODY * 10) X(cdof, ORC_MAXV * 6)   \
    X(cdof_dot, ORC_MAXV * 6) X(cvel, ORC_MAXBODY * 6) X(cacc, ORC_MAXBODY * 6)                    \
    X(cfrc, ORC_MAXBODY * 6) X(qM, ORC_MAXV * ORC_MAXV) X(qL, ORC_MAXV * ORC_MAXV)                 \
    X(qfrc_bias, ORC_MAXV) X(qfrc_passive, ORC_MAXV) X(qfrc_actuator, ORC_MAXV)                    \
    X(qfrc_applied, ORC_MAXV) X(qfrc_smooth, ORC_MAXV) X(qacc_smooth, ORC_MAXV) X(qacc, ORC_MAXV)  \
    X(qfrc_constraint, ORC_MAXV) X(ncon, 1) X(con_dist, ORC_MAXCON) X(con_pos, ORC_MAXCON * 3)     \
    X(con_frame, ORC_MAXCON * 9) X(con_geom, ORC_MAXCON) X(con_mu, ORC_MAXCON)                     \
    X(con_geom1, ORC_MAXCON) X(nself, 1) X(nbox_skipped, 1)                                        \
    X(con_includemargin, ORC_MAXCON) X(nefc, 1) X(efc_J, ORC_MAXEFC * ORC_MAXV)                    \
    X(efc_pos, ORC_MAXEFC) X(efc_D, ORC_MAXEFC) X(efc_aref, ORC_MAXEFC) X(efc_enabled, ORC_MAXEFC) \
    X(efc_force, ORC_MAXEFC) X(solver_niter, 1) X(com, 3)

typedef struct {
#define ORC_X(name, n) REAL name[n];
    ORC_DATA_FIELDS(ORC_X)
#undef ORC_X
} OrcData;

#ifdef __cplusplus
extern "C" {
#endif
int orc_real_size(void);
int orc_model_size(void);
int orc_model_field(const char* name, int* offset, int* count, int* is_int);
int orc_data_size(void);
int orc_data_field(const char* name, int* offset, int* count);

void orc_reset_data(const OrcModel* m, OrcData* d);
void orc_forward(const OrcModel* m, OrcData* d);          /* mjx.forward   */
void orc_step(const OrcModel* m, OrcData* d);             /* mjx.step      */
void orc_pd_torque(const OrcModel* m, const OrcData* lag, int mode, int action_cols,
                   const REAL* action, const REAL* q, const REAL* qd, const REAL* kp,
                   const REAL* kd, REAL time, REAL dt, REAL* tau);
#ifdef __cplusplus
}
#endif
#endif
