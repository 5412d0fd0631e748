// Mimic-mode epilogue of the env step: DeepMimic-style imitation reward, termination and statistics against
// reference frames interpolated from the motion table (build-defined; the reference returns reward = zeros(3),
// /root/reference/agents_env/agent_mimic.py:84 -- SURVEY.md G4 / App. A.11, "parity unpinned").
//
// The two table frames an env step needs (2 x 416 B, adjacent in HBM) are staged into the warp's shared-memory
// workspace with ONE 1-D bulk async copy (cp.async.bulk.shared::cluster.global, SASS UBLKCP) issued when the step
// starts and consumed after the last substep, so the copy is hidden behind the dynamics.
#pragma once
#include "pbrl_core.cuh"

namespace pbrl {

// ---- frame staging -------------------------------------------------------------------------------------------
PB_DEV void frame_lookup(const DevMimic& mm, int clip, float t, int& frame, float& u, float& cyc) {
    const float T = mm.clip_duration[clip], dt = mm.clip_dt[clip];
    float tc = t;
    cyc = 0.f;
    if (mm.clip_wrap[clip]) { cyc = floorf(t / T); tc = t - cyc * T; }
    else if (tc > T) tc = T;
    if (tc < 0.f) tc = 0.f;
    int fi = (int)(tc / dt);
    if (fi > mm.clip_nframes[clip] - 2) fi = mm.clip_nframes[clip] - 2;
    u = fminf(fmaxf(tc / dt - (float)fi, 0.f), 1.f);
    frame = mm.clip_first[clip] + fi;
}

#if defined(__CUDACC__) && !defined(PBRL_HOST_EMU)
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
#endif

PB_DEV void frames_init(Ws& w, int lane) {
#if defined(__CUDACC__) && !defined(PBRL_HOST_EMU)
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&w.fr_bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        w.fr_phase = 0;
    }
    __syncwarp();
#else
    (void)lane;
    w.fr_phase = 0;
#endif
}

// issue the copy of frames [frame, frame+1] -> w.fr (one elected lane; completion is tracked by the mbarrier)
PB_DEV void frames_prefetch(const float* frames, const DevMimic& mm, Ws& w, int clip, float t, int lane) {
    int frame;
    float u, cyc;
    frame_lookup(mm, clip, t, frame, u, cyc);
    const float* src = frames + (size_t)frame * PBRL_FRAME_STRIDE;
#if defined(__CUDACC__) && !defined(PBRL_HOST_EMU)
    __syncwarp();
    if (lane == 0) {
        w.fr_u = u; w.fr_cyc = cyc;
        const unsigned bytes = 2u * PBRL_FRAME_STRIDE * sizeof(float), bar = smem_u32(&w.fr_bar), dst = smem_u32(w.fr);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic reads of w.fr before the async write
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
    }
#else
    (void)lane;
    w.fr_u = u; w.fr_cyc = cyc;
    for (int i = 0; i < 2 * PBRL_FRAME_STRIDE; i++) w.fr[i] = src[i];
#endif
}

PB_DEV void frames_wait(Ws& w, int lane) {
#if defined(__CUDACC__) && !defined(PBRL_HOST_EMU)
    const unsigned bar = smem_u32(&w.fr_bar);
    const unsigned parity = (unsigned)w.fr_phase;
    unsigned done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    }
    __syncwarp();
    if (lane == 0) w.fr_phase ^= 1;
    __syncwarp();
#else
    (void)w; (void)lane;
#endif
}

// ---- small quaternion helpers -----------------------------------------------------------------------------------
PB_DEV void quat_slerp(const float* a, const float* b, float u, float* r) {
    float d = a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3], sg = 1.f, wa, wb;
    if (d < 0.f) { d = -d; sg = -1.f; }
    if (d > 0.9995f) { wa = 1.f - u; wb = u; }
    else {
        const float th = acosf(d), inv = 1.0f / sinf(th);
        wa = sinf((1.f - u) * th) * inv; wb = sinf(u * th) * inv;
    }
    for (int k = 0; k < 4; k++) r[k] = wa * a[k] + sg * wb * b[k];
    quat_normalize(r);
}
// geodesic angle between unit quaternions, 2 acos |a . b|, evaluated as 2 atan2(|vec(conj(a) b)|, |a . b|): the arc cosine loses
// half the digits near 1 (a tracked pose: angles of 1e-3 rad came out 3e-4 off in fp32 and dominated the reward's error)
PB_DEV float quat_angle(const float* a, const float* b) {
    const float d = fabsf(a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3]);
    const float vx = a[0] * b[1] - b[0] * a[1] - (a[2] * b[3] - a[3] * b[2]);
    const float vy = a[0] * b[2] - b[0] * a[2] - (a[3] * b[1] - a[1] * b[3]);
    const float vz = a[0] * b[3] - b[0] * a[3] - (a[1] * b[2] - a[2] * b[1]);
    return 2.f * atan2f(sqrtf(vx * vx + vy * vy + vz * vz), d);
}
PB_DEV void quat_to_euler_xyz(const float* q, float* e) {
    const float w = q[0], x = q[1], y = q[2], z = q[3];
    const float m00 = 1.f - 2.f * (y * y + z * z), m01 = 2.f * (x * y - z * w), m02 = 2.f * (x * z + y * w);
    const float m12 = 2.f * (y * z - x * w), m22 = 1.f - 2.f * (x * x + y * y);
    e[0] = atan2f(-m12, m22);
    e[1] = atan2f(m02, sqrtf(m22 * m22 + m12 * m12));
    e[2] = atan2f(-m01, m00);
}

// ---- interpolated reference pose from the staged frames: joint quaternions, velocities, body poses, com, end effectors
PB_DEVN void mimic_reference(const DevModel& m, const DevMimic& mm, const float* frames, Ws& w, int clip, int lane) {
    const float u = w.fr_u, cyc = w.fr_cyc;
    const float* f0 = w.fr;
    const float* f1 = w.fr + PBRL_FRAME_STRIDE;
    LANE_IF(j, 9) {  // slot 8 = root orientation
        const int o = j < 8 ? 69 + 4 * j : 3;
        quat_slerp(f0 + o, f1 + o, u, w.rq + 4 * j);
    }
    LANE_LOOP(i, NV) w.rvel[i] = (1.f - u) * f0[35 + i] + u * f1[35 + i];
    LANE_IF(k, 3) {
        float p = (1.f - u) * f0[k] + u * f1[k];
        if (k < 2 && cyc != 0.f) {
            const float* fa = frames + (size_t)mm.clip_first[clip] * PBRL_FRAME_STRIDE;
            const float* fz = frames + (size_t)(mm.clip_first[clip] + mm.clip_nframes[clip] - 1) * PBRL_FRAME_STRIDE;
            p += cyc * (fz[k] - fa[k]);
        }
        w.rpos[k] = p;
    }
    LANE_IF(b, NB) {
        if (b > 1 && m.body_jntnum[b] == 1) { const int qa = m.jnt_qposadr[m.body_jntadr[b]]; w.rhinge[b] = (1.f - u) * f0[qa] + u * f1[qa]; }
    }
    LANE_IF(z, 1) { w.rxpos[0] = w.rxpos[1] = w.rxpos[2] = 0.f; w.rxquat[0] = 1.f; w.rxquat[1] = w.rxquat[2] = w.rxquat[3] = 0.f; }
    WSYNC();
    for (int lv = 1; lv < m.nlevel; lv++) {
        const int s = w.lvl_start[lv], n = w.lvl_start[lv + 1] - s;
        LANE_IF(k, n) {
            const int b = m.level_body[s + k], p = m.body_parent[b];
            float pos[3], q[4];
            if (m.jnt_type[m.body_jntadr[b]] == 0) {
                for (int c = 0; c < 3; c++) pos[c] = w.rpos[c];
                for (int c = 0; c < 4; c++) q[c] = w.rq[32 + c];
            } else {
                float t[3], ql[4];
                quat_rotate(w.rxquat + S4 * p, m.body_pos[b], t);
                for (int c = 0; c < 3; c++) pos[c] = w.rxpos[S3 * p + c] + t[c];
                if (m.body_jntnum[b] == 3) {
                    for (int c = 0; c < 4; c++) ql[c] = w.rq[4 * mm.body_slot[b] + c];
                } else {
                    float sn, cs;
                    const int j = m.body_jntadr[b];
                    sincos_half_angle(0.5f * w.rhinge[b], &sn, &cs);
                    ql[0] = cs; ql[1] = m.jnt_axis[j][0] * sn; ql[2] = m.jnt_axis[j][1] * sn; ql[3] = m.jnt_axis[j][2] * sn;
                }
                quat_mul(w.rxquat + S4 * p, ql, q);
            }
            for (int c = 0; c < 3; c++) w.rxpos[S3 * b + c] = pos[c];
            for (int c = 0; c < 4; c++) w.rxquat[S4 * b + c] = q[c];
        }
        WSYNC();
    }
    const float inv = 1.0f / m.total_mass;
    for (int c = 0; c < 3; c++) {
        const float s = wsum(lane, NB, [&](int b) {
            float t[3];
            quat_rotate(w.rxquat + S4 * b, m.body_ipos[b], t);
            return m.body_mass[b] * (w.rxpos[S3 * b + c] + t[c]);
        });
        LANE_IF(z, 1) w.rcom[c] = s * inv;
    }
    LANE_IF(e, mm.n_ee) {
        float t[3];
        quat_rotate(w.rxquat + S4 * mm.ee_body[e], mm.ee_offset[e], t);
        for (int c = 0; c < 3; c++) w.ree[3 * e + c] = w.rxpos[S3 * mm.ee_body[e] + c] + t[c];
    }
    WSYNC();
}

// ---- reward / termination / metrics for the final state of the step (w.qpos, w.qvel) at clip time t
PB_DEVN void mimic_epilogue(const DevModel& m, const DevMimic& mm, const float* frames, Ws& w, const PbrlBuffers& B,
                            int env, int clip, float t, float* acc, int lane) {
    frames_wait(w, lane);
    mimic_reference(m, mm, frames, w, clip, lane);
    stage_fk(m, w, w.qpos, lane);
    stage_com(m, w, lane);
    stage_collide(m, w, lane);
    const float e_p = wsum(lane, NB, [&](int b) {
        if (b == 0) return 0.f;
        float a;
        if (b == 1) a = quat_angle(w.rq + 32, w.xquat + S4 * b);
        else if (m.body_jntnum[b] == 3) {
            const int p = m.body_parent[b];
            const float pc[4] = {w.xquat[S4 * p], -w.xquat[S4 * p + 1], -w.xquat[S4 * p + 2], -w.xquat[S4 * p + 3]};
            float rel[4];
            quat_mul(pc, w.xquat + S4 * b, rel);
            a = quat_angle(w.rq + 4 * mm.body_slot[b], rel);
        } else a = w.rhinge[b] - w.qpos[m.jnt_qposadr[m.body_jntadr[b]]];
        return a * a;
    });
    const float e_v = wsum(lane, NV, [&](int i) { const float a = i < 3 ? 0.f : w.rvel[i] - w.qvel[i]; return a * a; });
    const float e_e = wsum(lane, mm.n_ee, [&](int e) {
        float t3[3], s = 0.f;
        const int b = mm.ee_body[e];
        mat_vec(w.xmat + S9 * b, mm.ee_offset[e], t3);
        for (int c = 0; c < 3; c++) { const float a = w.ree[3 * e + c] - (w.xpos[S3 * b + c] + t3[c]); s += a * a; }
        return s;
    });
    const float e_c = (w.rcom[0] - w.com[0]) * (w.rcom[0] - w.com[0]) + (w.rcom[1] - w.com[1]) * (w.rcom[1] - w.com[1]) +
                      (w.rcom[2] - w.com[2]) * (w.rcom[2] - w.com[2]);
    const float nfall = mm.clip_fall[clip] ? wsum(lane, NC, [&](int c) {
        return (((mm.fall_geom_mask >> m.con_geom[c]) & 1u) && w.con_dist[c] < 0.f) ? 1.f : 0.f; }) : 0.f;
    const float rp = expf(-mm.k_pose * e_p), rv = expf(-mm.k_vel * e_v), re = expf(-mm.k_ee * e_e), rc = expf(-mm.k_com * e_c);
    const float r = mm.w_pose * rp + mm.w_vel * rv + mm.w_ee * re + mm.w_com * rc;
    const float fall = nfall > 0.f ? 1.f : 0.f;
    const float done = (fall > 0.f || (!mm.clip_wrap[clip] && t >= mm.clip_duration[clip])) ? 1.f : 0.f;
    LANE_IF(z, 1) {
        if (B.reward) { B.reward[(size_t)env * 3] = r; B.reward[(size_t)env * 3 + 1] = rp; B.reward[(size_t)env * 3 + 2] = re; }
        if (B.done) B.done[env] = done;
        if (B.metrics) { B.metrics[(size_t)env * 3] += 1.f; B.metrics[(size_t)env * 3 + 1] = e_p; B.metrics[(size_t)env * 3 + 2] = fall; }
    }
    acc[0] += r; acc[1] += r * r; acc[2] += done; acc[3] += fall; acc[4] += e_p; acc[5] += 1.f;
}

// ---- reference-state initialisation: state = interpolated reference pose at clip time t0
PB_DEVN void env_reset_mimic(const DevModel& m, const DevMimic& mm, const float* frames, Ws& w, const PbrlBuffers& B, int env, int lane) {
    const int clip = B.motion_id[env];
    frames_prefetch(frames, mm, w, clip, B.ref_t0[env], lane);
    frames_wait(w, lane);
    mimic_reference(m, mm, frames, w, clip, lane);
    LANE_IF(b, NB) {
        if (b == 1) {
            for (int c = 0; c < 3; c++) w.qpos[c] = w.rpos[c];
            for (int c = 0; c < 4; c++) w.qpos[3 + c] = w.rq[32 + c];
        } else if (b > 1) {
            const int qa = m.jnt_qposadr[m.body_jntadr[b]];
            if (m.body_jntnum[b] == 3) quat_to_euler_xyz(w.rq + 4 * mm.body_slot[b], w.qpos + qa);
            else w.qpos[qa] = w.rhinge[b];
        }
    }
    LANE_LOOP(i, NV) { w.qvel[i] = w.rvel[i]; w.warm[i] = 0.f; }
    WSYNC();
    store_row(B.qpos + (size_t)env * NQ, w.qpos, NQ, lane);
    store_row(B.qpos_lag + (size_t)env * NQ, w.qpos, NQ, lane);
    store_row(B.qvel + (size_t)env * NV, w.qvel, NV, lane);
    store_row(B.qvel_lag + (size_t)env * NV, w.qvel, NV, lane);
    store_row(B.warm + (size_t)env * NV, w.warm, NV, lane);
    store_row(B.ctrl + (size_t)env * NU, w.warm, NU, lane);
    if (B.obs) { store_row(B.obs + (size_t)env * PBRL_OBS, w.qpos, NQ, lane); store_row(B.obs + (size_t)env * PBRL_OBS + NQ, w.qvel, NV, lane); }
    LANE_IF(k, 3) { if (B.reward) B.reward[(size_t)env * 3 + k] = 0.f; if (B.metrics) B.metrics[(size_t)env * 3 + k] = 0.f; }
    LANE_IF(z, 1) { B.time[env] = 0.f; if (B.done) B.done[env] = 0.f; if (B.status) B.status[env] = 0; }
    WSYNC();
}


// ---- auto-reset of one env whose episode ended (brax AutoResetWrapper): state and obs restart, reward / done / metrics stay
PB_DEVN void env_autoreset(const DevModel& m, const DevMimic& mm, const float* frames, Ws& w, const StepArgs& a,
                           const float* next_ref_t0, int* steps, int env, int lane) {
    StepArgs r = a;
    r.buf.reward = nullptr; r.buf.done = nullptr; r.buf.metrics = nullptr; r.buf.status = nullptr;
    if (a.reward_mode == PBRL_REWARD_MIMIC) {
        if (next_ref_t0) {
            LANE_IF(z, 1) const_cast<float*>(a.buf.ref_t0)[env] = next_ref_t0[env];
            WSYNC();
        }
        env_reset_mimic(m, mm, frames, w, r.buf, env, lane);
    } else {
        env_reset(m, w, r, env, lane);
    }
    LANE_IF(z, 1) if (steps) steps[env] = 0;
    WSYNC();
}

}  // namespace pbrl
