// Host emulation of the CUDA kernel bodies -- TEST INFRASTRUCTURE ONLY.
//
// Compiles physics-based-rl_b200/csrc/pbrl_core.cuh with PBRL_HOST_EMU (lane loops run sequentially)
// so that the kernel LOGIC can be checked against the oracle on a machine without a GPU
// (`pytest -m "not gpu"`).  It cannot see races or shuffle bugs; those are covered by the `-m gpu`
// parity tests and compute-sanitizer.  Nothing in the product loads this library.
#define PBRL_HOST_EMU 1
#include <cstdlib>
#include "../../physics-based-rl_b200/csrc/pbrl_core.cuh"
#include "../../physics-based-rl_b200/csrc/pbrl_mimic.cuh"
#include "../../physics-based-rl_b200/csrc/pbrl_desc_reflect.h"

#include <cstdio>
#include <vector>

using namespace pbrl;

static DevModel g_model;
static DevMimic g_mimic;
static std::vector<float> g_frames;
static bool g_has_mimic = false;
static std::string g_err;

extern "C" {
const char* emu_last_error() { return g_err.c_str(); }
int emu_model_desc_size() { return (int)sizeof(PbrlModelDesc); }
int emu_model_desc_field(const char* n, int* o, int* c, int* i) { return model_desc_field(n, o, c, i); }
int emu_ws_bytes() { return (int)sizeof(Ws); }
int emu_model_bytes() { return (int)sizeof(DevModel); }
int emu_set_model(const PbrlModelDesc* d) {
    g_err = build_dev_model(*d, g_model);
    return g_err.empty() ? 0 : -1;
}
int emu_set_mimic(const PbrlMimicDesc* d) {
    g_err = build_dev_mimic(*d, g_model, g_mimic);
    if (!g_err.empty()) return -1;
    g_frames.assign(d->frames, d->frames + (size_t)d->nframes_total * PBRL_FRAME_STRIDE);
    g_has_mimic = true;
    return 0;
}
int emu_reset(int N, int reward_mode, const PbrlBuffers* buf) {
    std::vector<Ws> ws(1);
    init_ws_tables(g_model, ws[0], 0);
    frames_init(ws[0], 0);
    StepArgs a{};
    a.N = N; a.buf = *buf; a.reward_mode = reward_mode;
    for (int e = 0; e < N; e++) {
        if (reward_mode == PBRL_REWARD_MIMIC) env_reset_mimic(g_model, g_mimic, g_frames.data(), ws[0], a.buf, e, 0);
        else env_reset(g_model, ws[0], a, e, 0);
    }
    return 0;
}
int emu_autoreset(int N, int reward_mode, const float* done, const float* next_ref_t0, int* steps, const PbrlBuffers* buf) {
    std::vector<Ws> ws(1);
    init_ws_tables(g_model, ws[0], 0);
    frames_init(ws[0], 0);
    StepArgs a{};
    a.N = N; a.buf = *buf; a.reward_mode = reward_mode;
    if (reward_mode == PBRL_REWARD_MIMIC) { a.frames = g_frames.data(); a.mm = g_mimic; }
    for (int e = 0; e < N; e++)
        if (done[e] != 0.f) env_autoreset(g_model, g_mimic, g_frames.data(), ws[0], a, next_ref_t0, steps, e, 0);
    return 0;
}
int emu_step(int N, int n_frames, int pd_mode, int action_cols, const float* action, const float* kp, const float* kd,
             float pd_dt, const float* pd_time, unsigned ctrl_mask, int reward_mode, const PbrlBuffers* buf) {
    std::vector<Ws> ws(1);
    init_ws_tables(g_model, ws[0], 0);
    StepArgs a{};
    a.N = N; a.n_frames = n_frames; a.pd_mode = pd_mode; a.action_cols = action_cols; a.reward_mode = reward_mode;
    a.action = action; a.kp = kp; a.kd = kd; a.pd_time = pd_time; a.pd_dt = pd_dt; a.ctrl_mask = ctrl_mask; a.buf = *buf;
    if (reward_mode == PBRL_REWARD_MIMIC) { a.frames = g_frames.data(); a.mm = g_mimic; }
    frames_init(ws[0], 0);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    // PBRL_EMU_SPLIT=r: the step in two parts (PD law + r substeps for every env, then the rest), as pbrl_step does when it
    // regroups the envs inside the step
    const char* sp = getenv("PBRL_EMU_SPLIT");
    const int r = sp ? atoi(sp) : 0;
    if (r > 0 && r < n_frames && buf->cost && buf->status) {
        for (int e = 0; e < N; e++) env_step(g_model, ws[0], a, e, acc, 0, 0, r, false);
        for (int e = N - 1; e >= 0; e--) env_step(g_model, ws[0], a, e, acc, 0, r, n_frames, true);   // (another order: the parts are independent per env)
    } else {
        for (int e = 0; e < N; e++) env_step(g_model, ws[0], a, e, acc, 0, 0, n_frames, false);
    }
    if (buf->stats && reward_mode == PBRL_REWARD_MIMIC) for (int k = 0; k < 6; k++) buf->stats[k] += acc[k];
    return 0;
}
#ifdef PBRL_EMU_TRACE
void emu_set_trace(int* p) { g_emu_trace = p; }
long* emu_counters() { return g_emu_cnt; }
#endif
int emu_replay(int N, const float* qpos, const float* qvel, float* obs, float* reward, float* done, float* xpos, float* xquat) {
    std::vector<Ws> ws(1);
    init_ws_tables(g_model, ws[0], 0);
    for (int e = 0; e < N; e++) env_replay(g_model, ws[0], qpos, qvel, obs, reward, done, xpos, xquat, e, 0);
    return 0;
}
int emu_replay_item_fk(const float* qpos, float* xpos, float* xquat) {
    ReplayTables T;
    replay_tables_from(g_model, T);
    replay_item_fk<true>(T, qpos, xpos, xquat);
    return 0;
}
// the convex narrow-phase functions on their own (tests/test_self_collision.py): rows of [pos(3), dist, normal(3), -]
int emu_capsule_box(const float* c, const float* axis_hl, float r, const float* bc, const float* R, const float* half, float* out16) {
    capsule_box_dev(c, axis_hl, r, bc, R, half, out16);
    return 0;
}
int emu_box_box(const float* c1, const float* R1, const float* h1, const float* c2, const float* R2, const float* h2, float* out32) {
    float scratch[300];
    box_box_warp(c1, R1, h1, c2, R2, h2, out32, out32 + 8, scratch, 0);
    return 0;
}
int emu_pd_torque(int N, int pd_mode, int action_cols, const float* action, const float* qpos, const float* qvel,
                  const float* qpos_lag, const float* qvel_lag, const float* kp, const float* kd, const float* time,
                  float pd_dt, float* tau, float* qM, float* bias) {
    std::vector<Ws> ws(1);
    init_ws_tables(g_model, ws[0], 0);
    for (int e = 0; e < N; e++)
        env_pd(g_model, ws[0], pd_mode, action_cols, action, qpos, qvel, qpos_lag, qvel_lag, kp, kd, time, pd_dt, tau, qM, bias, e, 0);
    return 0;
}
int emu_debug_forward(int N, const float* qpos, const float* qvel, const float* ctrl, const float* warm, float* out) {
    std::vector<Ws> ws(1);
    init_ws_tables(g_model, ws[0], 0);
    for (int e = 0; e < N; e++) env_debug_forward(g_model, ws[0], qpos, qvel, ctrl, warm, out, e, 0);
    return 0;
}
}
