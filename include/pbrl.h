/* pbrl.h -- C ABI of the B200-native batched humanoid-mimic simulator (libpbrl.so).
 *
 * The reference has no FFI for this path: its env step runs as Python glue over brax/MJX
 * (SURVEY.md section 8b).  The seams a replacement sits behind are therefore the ones the
 * reference itself uses, and every entry point below names the reference interface it replaces:
 *
 *   pbrl_reset      HumanoidDiff.reset            /root/reference/agents_env/agent_env_template.py:88-112
 *                   (brax pipeline_init = mjx.make_data + mjx.forward)
 *   pbrl_step       HumanoidTrain.step + variants /root/reference/agents_env/agent_mimic.py:58-91,94-271
 *                   (PD callback, then PipelineEnv.pipeline_step = n_frames x mjx.step, obs/reward/done)
 *   pbrl_replay     HumanoidReplay.step           /root/reference/agents_env/agent_replay_motion.py:51-83
 *   pbrl_replay_rollout  the frame loop around it /root/reference/testing_forward_env.ipynb cell 10 (write frame i, step, next)
 *   pbrl_autoreset  brax AutoResetWrapper.step    brax 0.10.3 envs/wrappers/training.py (imported by agent_mimic.py:14-15 via ppo)
 *   pbrl_balance_order   no counterpart: a scheduling hint for pbrl_step (which envs share a CTA)
 *   pbrl_step_launches   no counterpart: how many kernels a pbrl_step call launches (bench.py's gpu_launches)
 *   pbrl_pd_torque  standard/feedback/stable PD   /root/reference/agents_env/pds_controllers_agents.py:17-172
 *                   and the *_mjx signature twins /root/reference/agents_env/pds_controllers_mjx.py:17-191
 *   pbrl_debug_forward  fields of brax `pipeline_state` the reference reads back
 *                   (qM, qfrc_bias, ...; /root/reference/agents_env/pds_controllers_agents.py:65,69,86-90)
 *
 * Conventions: plain pointers and sizes, no torch types.  All array arguments of the device entry
 * points are DEVICE pointers to env-major fp32 rows with the logical widths (qpos 35, qvel 34,
 * ctrl 28, obs 69, reward 3, metrics 3; no padding).  Calls are stream-ordered on `stream`
 * (a cudaStream_t passed as void*), do not allocate and do not synchronise.  The model handle is
 * immutable after creation.  Return value: 0 on success, negative on error (text via
 * pbrl_last_error()).  There is no CPU fallback: every entry point fails if CUDA is unavailable.
 */
#ifndef PBRL_H
#define PBRL_H

#ifdef __cplusplus
extern "C" {
#endif

#define PBRL_NBODY 14
#define PBRL_NQ 35
#define PBRL_NV 34
#define PBRL_NU 28
#define PBRL_NJNT 29
#define PBRL_NGEOM 16
#define PBRL_OBS 69
#define PBRL_MAXCLIP 8
#define PBRL_FRAME_STRIDE 104 /* floats per motion-table frame (416 B, a 16-byte multiple for cp.async.bulk) */

/* pd_mode */
#define PBRL_PD_TORQUE 0   /* action IS the torque (step_custom_target_and_outside, agent_mimic.py:239-271) */
#define PBRL_PD_STANDARD 1 /* pds_controllers_agents.py:17-34  */
#define PBRL_PD_FEEDBACK 2 /* pds_controllers_agents.py:42-78  */
#define PBRL_PD_STABLE 3   /* pds_controllers_agents.py:137-172 */

/* reward_mode */
#define PBRL_REWARD_REFERENCE 0 /* reward = zeros(3), done = 1.0, obs = [qpos,qvel]  (agent_mimic.py:72,84,86) */
#define PBRL_REWARD_MIMIC 1     /* DeepMimic pose/vel/end-effector/CoM reward vs interpolated reference frames */

/* status bits written per env */
#define PBRL_STATUS_NAN 1
#define PBRL_STATUS_PIVOT 2
#define PBRL_STATUS_SOLVER_MAXITER 4
#define PBRL_STATUS_SELF_OVERFLOW 8 /* more body-vs-body contacts inside the margin than the kernel keeps (8): the deepest 8 were kept */

typedef struct PbrlModel PbrlModel;

/* Host-side model description.  Field names and meanings equal MuJoCo's mjModel / the attributes
 * of ModelConst (physics-based-rl_b200/mjcf.py); fixed capacities, doubles and ints only. */
#define PBRL_MAXBODY 16
#define PBRL_MAXQ 40
#define PBRL_MAXV 36
#define PBRL_MAXU 32
#define PBRL_MAXJNT 32
#define PBRL_MAXGEOM 20
#define PBRL_MAXPAIR 128
#define PBRL_MODEL_FIELDS(I, D)                                                                     \
    I(nbody, 1) I(nq, 1) I(nv, 1) I(nu, 1) I(njnt, 1) I(ngeom, 1) I(iterations, 1)                  \
    I(ls_iterations, 1) I(solver_rule, 1)                                                            \
    D(timestep, 1) D(gravity, 3) D(solref, 2) D(solimp, 5) D(impratio, 1) D(tolerance, 1)           \
    D(ls_tolerance, 1) D(meaninertia, 1)                                                             \
    I(body_parent, PBRL_MAXBODY) D(body_pos, PBRL_MAXBODY * 3) D(body_mass, PBRL_MAXBODY)           \
    D(body_ipos, PBRL_MAXBODY * 3) D(body_inertia, PBRL_MAXBODY * 9) I(body_jntadr, PBRL_MAXBODY)   \
    I(body_jntnum, PBRL_MAXBODY) I(body_dofadr, PBRL_MAXBODY) I(body_dofnum, PBRL_MAXBODY)          \
    D(body_invweight0, PBRL_MAXBODY * 2)                                                             \
    I(jnt_type, PBRL_MAXJNT) I(jnt_bodyid, PBRL_MAXJNT) I(jnt_qposadr, PBRL_MAXJNT)                 \
    I(jnt_dofadr, PBRL_MAXJNT) D(jnt_axis, PBRL_MAXJNT * 3) D(jnt_pos, PBRL_MAXJNT * 3)             \
    I(jnt_limited, PBRL_MAXJNT) D(jnt_range, PBRL_MAXJNT * 2) D(jnt_margin, PBRL_MAXJNT)            \
    I(dof_bodyid, PBRL_MAXV) I(dof_parentid, PBRL_MAXV) D(dof_armature, PBRL_MAXV)                  \
    D(dof_damping, PBRL_MAXV) D(dof_invweight0, PBRL_MAXV) D(qpos0, PBRL_MAXQ)                      \
    I(geom_type, PBRL_MAXGEOM) I(geom_bodyid, PBRL_MAXGEOM) D(geom_size, PBRL_MAXGEOM * 3)          \
    D(geom_pos, PBRL_MAXGEOM * 3) D(geom_quat, PBRL_MAXGEOM * 4) D(geom_friction, PBRL_MAXGEOM * 3) \
    D(geom_margin, PBRL_MAXGEOM) D(geom_gap, PBRL_MAXGEOM)                                           \
    I(act_dofadr, PBRL_MAXU) D(act_gear, PBRL_MAXU) I(act_ctrllimited, PBRL_MAXU)                   \
    D(act_ctrlrange, PBRL_MAXU * 2)                                                                  \
    I(npair, 1) I(pair_geom1, PBRL_MAXPAIR) I(pair_geom2, PBRL_MAXPAIR)

/* Constraint-solver options (mjOption: iterations, tolerance, ls_iterations, ls_tolerance) and
 * `solver_rule`, which selects how the primal Newton iteration on qacc stops:
 *   PBRL_SOLVER_EXACT (0, default)  iterate until a full Newton step with exact line search leaves the active set of the
 *       piecewise-quadratic cost unchanged -- the minimiser is then reached exactly; at most min(iterations, 12) iterations
 *       (PBRL_STATUS_SOLVER_MAXITER reports the cap); the line search runs to fp32 resolution.  `tolerance`, `ls_iterations`
 *       and `ls_tolerance` are not used.
 *   PBRL_SOLVER_MJX (1)             MuJoCo / MJX's rule: stop when tolerance > improvement / (meaninertia * max(1, nv)) or
 *       tolerance > |gradient| / (meaninertia * max(1, nv)), at most `iterations` iterations; the line search stops at
 *       |phi'| < tolerance * ls_tolerance * |search| * meaninertia * max(1, nv) or after `ls_iterations` iterations.
 * Both rules converge to the same (unique) minimiser; they differ in the stopping point. */
#define PBRL_SOLVER_EXACT 0
#define PBRL_SOLVER_MJX 1

typedef struct PbrlModelDesc {
#define PBRL_I(name, n) int name[n];
#define PBRL_D(name, n) double name[n];
    PBRL_MODEL_FIELDS(PBRL_I, PBRL_D)
#undef PBRL_I
#undef PBRL_D
} PbrlModelDesc;

/* Mimic-mode configuration: reference-motion table + reward/termination parameters.
 * Build-defined (the reference's reward is zeros(3)). */
typedef struct PbrlMimicDesc {
    int nclip;                /* <= PBRL_MAXCLIP */
    int nframes_total;
    const float* frames;      /* HOST pointer, [nframes_total][PBRL_FRAME_STRIDE]; layout in mocap.py */
    int clip_first[PBRL_MAXCLIP];
    int clip_nframes[PBRL_MAXCLIP];
    float clip_duration[PBRL_MAXCLIP]; /* sum of frame durations */
    float clip_dt[PBRL_MAXCLIP];       /* frame duration (uniform inside a DeepMimic clip) */
    int clip_wrap[PBRL_MAXCLIP];       /* "Loop": "wrap" */
    int clip_fall_enabled[PBRL_MAXCLIP];
    float w_pose, w_vel, w_ee, w_com;  /* 0.65, 0.1, 0.15, 0.1 */
    float k_pose, k_vel, k_ee, k_com;  /* 2, 0.1, 40, 10 */
    int n_ee;                          /* <= 4 end effectors */
    int ee_body[4];
    float ee_offset[12];               /* body-frame offsets */
    unsigned fall_geom_mask;           /* geoms whose floor penetration terminates the episode */
} PbrlMimicDesc;

/* Per-call device buffers.  State is updated in place. */
typedef struct PbrlBuffers {
    float* qpos;     /* [N][35] */
    float* qvel;     /* [N][34] */
    float* time;     /* [N]     */
    float* warm;     /* [N][34] qacc_warmstart */
    float* ctrl;     /* [N][28] last applied ctrl (pipeline_state.ctrl) */
    float* qpos_lag; /* [N][35] state that entered the last substep: reproduces the one-substep lag   */
    float* qvel_lag; /* [N][34] of pipeline_state.qM / qfrc_bias                 */
    float* obs;      /* [N][69] out */
    float* reward;   /* [N][3]  out */
    float* done;     /* [N]     out */
    float* metrics;  /* [N][3]  in/out: step_index, pose_error, fall (agent_env_template.py:93) */
    int* status;     /* [N]     out, may be NULL */
    const int* motion_id; /* [N] mimic mode, may be NULL */
    const float* ref_t0;  /* [N] clip time at reset, mimic mode, may be NULL */
    float* stats;    /* [8] accumulated with atomics: sum r, sum r^2, sum done, sum fall, sum pose_err, n, 0, 0; may be NULL */
    const int* env_order; /* [N] optional processing order (a permutation of 0..N-1), may be NULL.  Envs that are
                             neighbours in this order share a CTA; grouping envs of similar cost (e.g. by clip)
                             balances the CTA-wide phase barriers of pbrl_step.  Results do not depend on it. */
    int* cost;       /* [N] out, may be NULL: work estimate of the env step just taken (constraint-solver iterations of its most
                        expensive substep, dense-Hessian iterations weighted up); the input of pbrl_balance_order */
} PbrlBuffers;

int pbrl_model_desc_size(void);
int pbrl_model_desc_field(const char* name, int* offset, int* count, int* is_int);
const char* pbrl_last_error(void);
int pbrl_device_count(void);

int pbrl_model_create(const PbrlModelDesc* desc, const PbrlMimicDesc* mimic /* may be NULL */, int device,
                      PbrlModel** out);
void pbrl_model_destroy(PbrlModel* model);

int pbrl_reset(const PbrlModel* model, int N, int reward_mode, const PbrlBuffers* buf, void* stream);

/* Processing order for the next pbrl_step from the work estimates of the last one (PbrlBuffers.cost): envs sorted by
 * decreasing cost, so that the 16 envs that share a CTA (and its phase barriers) need similar numbers of solver iterations
 * and the expensive groups start first.  A counting sort on the device (one small launch); results of pbrl_step do not
 * depend on the order. */
int pbrl_balance_order(const PbrlModel* model, int N, const int* cost /* [N] */, int* env_order /* [N] out */, void* stream);

/* Auto-reset (brax.envs.wrappers.training.AutoResetWrapper.step, the `where(done, first_state, state)` part): every
 * env with done[e] != 0 restarts -- qpos0 / zero velocity in reference mode, the reference pose at clip time
 * next_ref_t0[e] in mimic mode (NULL: the env's current ref_t0; otherwise ref_t0[e] is overwritten) -- and its obs row
 * becomes the first observation of the new episode.  reward / done / metrics keep the values of the step that ended the
 * episode (the trainer reads them).  steps (may be NULL) is zeroed for the restarted envs. */
int pbrl_autoreset(const PbrlModel* model, int N, int reward_mode, const float* done /* [N] */,
                   const float* next_ref_t0 /* [N] or NULL */, int* steps /* [N] or NULL */, const PbrlBuffers* buf,
                   void* stream);

/* Batches of at least one group of 16 envs per SM that pass env_order, cost and status are stepped in two parts:
 * PD law + the first half of the substeps (4 of 8) in the caller's order, then the other substeps and the epilogue with the envs regrouped by
 * the work of the first part -- an env's solver work is predicted far better by its previous substep than by its previous
 * env step.  Below 16 groups per SM both parts run in one launch (an env that is through part A queues for part B by work
 * class; CTAs draw the 16 heaviest on offer); at and above, in two launches with a sort in between.  Results do not depend on
 * it (PBRL_REGROUP_AFTER=0 at model creation: always the whole step in one pass).  pbrl_step_launches tells how many kernels a
 * call launches. */
int pbrl_step_launches(const PbrlModel* model, int N, int n_frames, const PbrlBuffers* buf);
int pbrl_step(const PbrlModel* model, int N, int n_frames, int pd_mode, int action_cols /* 1 or 4 */,
              const float* action /* [N][28][action_cols] */, const float* kp /* [28] */,
              const float* kd /* [28] */, float pd_dt, const float* pd_time /* [N] or NULL = state time */,
              unsigned ctrl_mask /* bit a: actuator a takes the new torque */, int reward_mode,
              const PbrlBuffers* buf, void* stream);

int pbrl_replay(const PbrlModel* model, int N, const float* qpos, const float* qvel, float* obs,
                float* reward, float* done, float* xpos /* [N][14][3] or NULL */,
                float* xquat /* [N][14][4] or NULL */, void* stream);

/* Kinematic replay of a frame sequence in one launch: HumanoidReplay.step in the loop the reference's notebooks run
 * (testing_forward_env.ipynb cell 10: for every frame write data_pos[i] / data_vel[i] into the state, then step).
 * Env e replays frame (frame0[e] + k) mod F at step k = 0 .. K-1 (frame0 NULL: frame0[e] = e).  data_pos [F][35] and
 * data_vel [F][34] are device arrays.  Outputs are step-major: obs [K][N][69], reward [K][N][3] (zeros), done [K][N] (ones),
 * xpos [K][N][14][3], xquat [K][N][14][4]; any of them may be NULL.
 * When the N x K items outnumber the F clip rows (>= 4 items per row) the kinematics are evaluated once per clip row per call
 * into a frame table owned by the model and the items gather their rows from it; otherwise per item.  Same arithmetic, same
 * bits (PBRL_REPLAY_PER_ITEM=1 forces the per-item form).  When every output array is given and 16-byte aligned, the gathered
 * rows leave through the bulk-copy engine in blocks of 16 items (PBRL_REPLAY_BULK=0: ordinary stores; same bits). */
int pbrl_replay_rollout(const PbrlModel* model, int N, int K, const float* data_pos, const float* data_vel, int F,
                        const int* frame0, float* obs, float* reward, float* done, float* xpos, float* xquat, void* stream);

int pbrl_pd_torque(const PbrlModel* model, int N, int pd_mode, int action_cols, const float* action,
                   const float* qpos, const float* qvel, const float* qpos_lag, const float* qvel_lag,
                   const float* kp, const float* kd, const float* time, float pd_dt, float* tau /* [N][28] */,
                   float* qM /* [N][34][34] or NULL */, float* qfrc_bias /* [N][34] or NULL */, void* stream);

/* Stage intermediates of one mjx.forward at (qpos, qvel, ctrl, warm), for parity tests.
 * out is [N][PBRL_DBG_SIZE] floats with the layout below. */
#define PBRL_DBG_XPOS 0            /* 14*3  */
#define PBRL_DBG_XQUAT 42          /* 14*4  */
#define PBRL_DBG_COM 98            /* 3     */
#define PBRL_DBG_QM 101            /* 34*34 dense */
#define PBRL_DBG_BIAS 1257         /* 34 */
#define PBRL_DBG_QFRC_SMOOTH 1291  /* 34 */
#define PBRL_DBG_QACC_SMOOTH 1325  /* 34 */
#define PBRL_DBG_QACC 1359         /* 34 */
#define PBRL_DBG_QFRC_CONSTRAINT 1393 /* 34 */
#define PBRL_DBG_NCON_ACTIVE 1427  /* 1  */
#define PBRL_DBG_CON_DIST 1428     /* 32 */
#define PBRL_DBG_CON_POS 1460      /* 32*3 */
#define PBRL_DBG_NITER 1556        /* 1 */
#define PBRL_DBG_CON_FORCE 1557    /* 32*4 pyramid forces */
#define PBRL_DBG_LIM_FORCE 1685    /* 28 */
#define PBRL_DBG_NSELF 1713        /* 1: body-vs-body contacts inside the margin */
#define PBRL_DBG_SELF_DIST 1714    /* 8 */
#define PBRL_DBG_SELF_FORCE 1722   /* 8*4 pyramid forces */
#define PBRL_DBG_SIZE 1756
int pbrl_debug_forward(const PbrlModel* model, int N, const float* qpos, const float* qvel, const float* ctrl,
                       const float* warm, float* out, void* stream);

/* Dev builds only (-DPBRL_STAGE_PROF): cumulative cycles / hit counts per kernel stage, 64 slots each.  The product
 * build returns -1 ("not a profiling build"). */
int pbrl_debug_stage_profile(unsigned long long* cycles, unsigned long long* hits, int reset);

/* ---- host-buffer convenience layer: what a Python/JAX host that keeps its data on the CPU calls.
 * Owns device state for N envs; every call moves the inputs host->device, launches, moves the outputs device->host on the
 * handle's stream and returns after the results are in the host buffers.  action / obs_out / reward_out / done_out that are
 * page-locked, device-mapped host memory (cudaHostAlloc, cudaHostRegister, torch pin_memory) are read and written IN PLACE by
 * the kernel -- the rows cross PCIe while the other envs are being stepped; pageable buffers are staged through
 * cudaMemcpyAsync (PBRL_HOST_ZERO_COPY=0 forces the staging). */
typedef struct PbrlBatch PbrlBatch;
int pbrl_batch_create(const PbrlModel* model, int N, PbrlBatch** out);
void pbrl_batch_destroy(PbrlBatch* batch);
int pbrl_batch_reset_host(PbrlBatch* batch, int reward_mode, const int* motion_id /* [N] or NULL */,
                          const float* ref_t0 /* [N] or NULL */, float* obs_out /* [N][69] */);
int pbrl_batch_step_host(PbrlBatch* batch, int n_frames, int pd_mode, int action_cols, const float* action,
                         const float* kp, const float* kd, float pd_dt, int reward_mode, float* obs_out,
                         float* reward_out, float* done_out, float* stats_out /* [8] or NULL */);
int pbrl_batch_get_state_host(PbrlBatch* batch, float* qpos, float* qvel, float* time);

#ifdef __cplusplus
}
#endif
#endif /* PBRL_H */
