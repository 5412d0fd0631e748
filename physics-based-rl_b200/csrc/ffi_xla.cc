// XLA FFI shim: pbrl_reset / pbrl_step / pbrl_autoreset / pbrl_replay / pbrl_pd_torque as `jax.ffi` custom calls, so that the
// reference's JAX host code (`state = env.step(state, action)`, /root/reference/agents_env/agent_mimic.py:58-91, call site of
// the pipeline :82) keeps its dataflow while the work runs in libpbrl (include/pbrl.h).
//
// Dataflow.  XLA custom calls are functional: operands are read-only, results are fresh buffers.  Every state tensor the C
// ABI updates in place is therefore BOTH an operand and a result of the call, and the Python side declares
// `input_output_aliases` (operand index -> result index) so that XLA hands the handler the same
// buffer for both whenever the operand is dead after the call (the normal case inside `lax.scan` / a jitted step).  When XLA
// cannot alias (the caller still holds the old state), the handler first copies operand -> result on the stream and then
// runs the kernels on the results; it never writes to an operand.  Accumulators the kernels add to (`stats`) are zeroed
// here before the launch; `metrics` is state (in / out).
//
// STATUS: type-checked in this image against tests/stubs/xla/ffi/api/ffi.h (tests/test_host_side.py); NOT run -- jaxlib is
// not installed here.  The batch dimension is inside the call ([N, width] fp32 buffers); the model handle travels as an
// int64 attribute.  Build line: INTEGRATION.md section 4.
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define PBRL_HAVE_XLA_FFI 1
#endif
#endif

#ifdef PBRL_HAVE_XLA_FFI
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/pbrl.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;
using F32 = ffi::Buffer<ffi::F32>;
using S32 = ffi::Buffer<ffi::S32>;
template <class B> using Out = ffi::Result<B>;

namespace {

ffi::Error fail_if(int rc) {
    return rc == 0 ? ffi::Error::Success() : ffi::Error(ffi::ErrorCode::kInternal, pbrl_last_error());
}
ffi::Error cuda_error(cudaError_t e) { return ffi::Error(ffi::ErrorCode::kInternal, cudaGetErrorString(e)); }

// result <- operand unless XLA aliased them (input_output_aliases honoured): the kernels then update the result in place
template <class B>
cudaError_t adopt(cudaStream_t stream, const B& in, Out<B>& out) {
    if (static_cast<const void*>(out->typed_data()) == static_cast<const void*>(in.typed_data())) return cudaSuccess;
    return cudaMemcpyAsync(out->typed_data(), in.typed_data(), in.element_count() * sizeof(*in.typed_data()), cudaMemcpyDeviceToDevice, stream);
}
#define ADOPT(in, out)                                                   \
    do {                                                                 \
        cudaError_t e_ = adopt(stream, in, out);                         \
        if (e_ != cudaSuccess) return cuda_error(e_);                    \
    } while (0)

const PbrlModel* handle(int64_t model) { return reinterpret_cast<const PbrlModel*>(static_cast<intptr_t>(model)); }

// ------------------------------------------------------------------------------------------------ reset
// HumanoidDiff.reset (agent_env_template.py:88-112).  operands: motion_id [N], ref_t0 [N] (read in mimic mode only)
// results: qpos, qvel, time, warm, ctrl, qpos_lag, qvel_lag, obs, reward, done, metrics, status
ffi::Error ResetImpl(cudaStream_t stream, int64_t model, int32_t reward_mode, S32 motion_id, F32 ref_t0, Out<F32> qpos, Out<F32> qvel,
                     Out<F32> time, Out<F32> warm, Out<F32> ctrl, Out<F32> qpos_lag, Out<F32> qvel_lag, Out<F32> obs, Out<F32> reward,
                     Out<F32> done, Out<F32> metrics, Out<S32> status) {
    const int N = (int)qpos->dimensions()[0];
    PbrlBuffers b{};
    b.qpos = qpos->typed_data(); b.qvel = qvel->typed_data(); b.time = time->typed_data(); b.warm = warm->typed_data();
    b.ctrl = ctrl->typed_data(); b.qpos_lag = qpos_lag->typed_data(); b.qvel_lag = qvel_lag->typed_data();
    b.obs = obs->typed_data(); b.reward = reward->typed_data(); b.done = done->typed_data(); b.metrics = metrics->typed_data();
    b.status = status->typed_data(); b.motion_id = motion_id.typed_data(); b.ref_t0 = ref_t0.typed_data();
    return fail_if(pbrl_reset(handle(model), N, reward_mode, &b, stream));
}
}  // namespace
XLA_FFI_DEFINE_HANDLER_SYMBOL(PbrlResetFfi, ResetImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("model").Attr<int32_t>("reward_mode")
                                  .Arg<S32>().Arg<F32>()
                                  .Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>()
                                  .Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<S32>());

// ------------------------------------------------------------------------------------------------ step
// HumanoidTrain.step and its variants (agent_mimic.py:58-271).
// operands 0..3:  action [N,28(,4)], kp [28], kd [28], pd_time [N] (read when use_pd_time != 0)
// operands 4..11: qpos, qvel, time, warm, ctrl, qpos_lag, qvel_lag, metrics      -> aliased to results 0..7
// operands 12..14: motion_id [N], ref_t0 [N], env_order [N] (processing order; arange(N) is fine)
// results  0..7:  the new state (same order)      8..11: obs, reward, done, stats[8]
// results 12..14: status [N], cost [N], next env_order [N] (pbrl_balance_order of this step's work estimates)
// Python: input_output_aliases = {4: 0, 5: 1, 6: 2, 7: 3, 8: 4, 9: 5, 10: 6, 11: 7}
namespace {
ffi::Error StepImpl(cudaStream_t stream, int64_t model, int32_t n_frames, int32_t pd_mode, int32_t action_cols, float pd_dt,
                    int32_t reward_mode, int32_t ctrl_mask, int32_t use_pd_time, F32 action, F32 kp, F32 kd, F32 pd_time, F32 qpos,
                    F32 qvel, F32 time, F32 warm, F32 ctrl, F32 qpos_lag, F32 qvel_lag, F32 metrics, S32 motion_id, F32 ref_t0,
                    S32 env_order, Out<F32> qpos_o, Out<F32> qvel_o, Out<F32> time_o, Out<F32> warm_o, Out<F32> ctrl_o,
                    Out<F32> qpos_lag_o, Out<F32> qvel_lag_o, Out<F32> metrics_o, Out<F32> obs, Out<F32> reward, Out<F32> done,
                    Out<F32> stats, Out<S32> status, Out<S32> cost, Out<S32> next_order) {
    const int N = (int)qpos.dimensions()[0];
    ADOPT(qpos, qpos_o); ADOPT(qvel, qvel_o); ADOPT(time, time_o); ADOPT(warm, warm_o); ADOPT(ctrl, ctrl_o);
    ADOPT(qpos_lag, qpos_lag_o); ADOPT(qvel_lag, qvel_lag_o); ADOPT(metrics, metrics_o);
    cudaError_t e = cudaMemsetAsync(stats->typed_data(), 0, 8 * sizeof(float), stream);   // the kernel accumulates with atomics
    if (e != cudaSuccess) return cuda_error(e);
    PbrlBuffers b{};
    b.qpos = qpos_o->typed_data(); b.qvel = qvel_o->typed_data(); b.time = time_o->typed_data(); b.warm = warm_o->typed_data();
    b.ctrl = ctrl_o->typed_data(); b.qpos_lag = qpos_lag_o->typed_data(); b.qvel_lag = qvel_lag_o->typed_data();
    b.metrics = metrics_o->typed_data(); b.obs = obs->typed_data(); b.reward = reward->typed_data(); b.done = done->typed_data();
    b.stats = stats->typed_data(); b.status = status->typed_data(); b.cost = cost->typed_data();
    b.motion_id = motion_id.typed_data(); b.ref_t0 = ref_t0.typed_data(); b.env_order = env_order.typed_data();
    int rc = pbrl_step(handle(model), N, n_frames, pd_mode, action_cols, action.typed_data(), kp.typed_data(), kd.typed_data(), pd_dt,
                       use_pd_time ? pd_time.typed_data() : nullptr, static_cast<unsigned>(ctrl_mask), reward_mode, &b, stream);
    if (rc) return fail_if(rc);
    return fail_if(pbrl_balance_order(handle(model), N, cost->typed_data(), next_order->typed_data(), stream));
}
}  // namespace
XLA_FFI_DEFINE_HANDLER_SYMBOL(PbrlStepFfi, StepImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("model").Attr<int32_t>("n_frames").Attr<int32_t>("pd_mode")
                                  .Attr<int32_t>("action_cols").Attr<float>("pd_dt").Attr<int32_t>("reward_mode")
                                  .Attr<int32_t>("ctrl_mask").Attr<int32_t>("use_pd_time")
                                  .Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>()
                                  .Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>()
                                  .Arg<S32>().Arg<F32>().Arg<S32>()
                                  .Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>()
                                  .Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>()
                                  .Ret<S32>().Ret<S32>().Ret<S32>());

// ------------------------------------------------------------------------------------------------ auto-reset
// brax AutoResetWrapper.step (imported through ppo at agent_mimic.py:14-15): envs with done != 0 restart.
// operands 0..1: done [N], next_ref_t0 [N] (mimic mode: the clip phase a restarted env starts at)
// operands 2..11: qpos, qvel, time, warm, ctrl, qpos_lag, qvel_lag, obs, ref_t0, steps -> aliased to results 0..9
// operand 12: motion_id [N]
// Python: input_output_aliases = {2: 0, 3: 1, 4: 2, 5: 3, 6: 4, 7: 5, 8: 6, 9: 7, 10: 8, 11: 9}
namespace {
ffi::Error AutoresetImpl(cudaStream_t stream, int64_t model, int32_t reward_mode, F32 done, F32 next_ref_t0, F32 qpos, F32 qvel, F32 time,
                         F32 warm, F32 ctrl, F32 qpos_lag, F32 qvel_lag, F32 obs, F32 ref_t0, S32 steps, S32 motion_id, Out<F32> qpos_o,
                         Out<F32> qvel_o, Out<F32> time_o, Out<F32> warm_o, Out<F32> ctrl_o, Out<F32> qpos_lag_o, Out<F32> qvel_lag_o,
                         Out<F32> obs_o, Out<F32> ref_t0_o, Out<S32> steps_o) {
    const int N = (int)qpos.dimensions()[0];
    ADOPT(qpos, qpos_o); ADOPT(qvel, qvel_o); ADOPT(time, time_o); ADOPT(warm, warm_o); ADOPT(ctrl, ctrl_o);
    ADOPT(qpos_lag, qpos_lag_o); ADOPT(qvel_lag, qvel_lag_o); ADOPT(obs, obs_o); ADOPT(ref_t0, ref_t0_o); ADOPT(steps, steps_o);
    PbrlBuffers b{};
    b.qpos = qpos_o->typed_data(); b.qvel = qvel_o->typed_data(); b.time = time_o->typed_data(); b.warm = warm_o->typed_data();
    b.ctrl = ctrl_o->typed_data(); b.qpos_lag = qpos_lag_o->typed_data(); b.qvel_lag = qvel_lag_o->typed_data();
    b.obs = obs_o->typed_data(); b.motion_id = motion_id.typed_data();
    b.ref_t0 = ref_t0_o->typed_data();   // pbrl_autoreset stores the new clip phase of a restarted env here
    return fail_if(pbrl_autoreset(handle(model), N, reward_mode, done.typed_data(), next_ref_t0.typed_data(), steps_o->typed_data(), &b, stream));
}
}  // namespace
XLA_FFI_DEFINE_HANDLER_SYMBOL(PbrlAutoresetFfi, AutoresetImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("model").Attr<int32_t>("reward_mode")
                                  .Arg<F32>().Arg<F32>()
                                  .Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<S32>()
                                  .Arg<S32>()
                                  .Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<S32>());

// ------------------------------------------------------------------------------------------------ replay, PD laws (pure functions)
namespace {
ffi::Error ReplayImpl(cudaStream_t stream, int64_t model, F32 qpos, F32 qvel, Out<F32> obs, Out<F32> reward, Out<F32> done,
                      Out<F32> xpos, Out<F32> xquat) {
    const int N = (int)qpos.dimensions()[0];
    return fail_if(pbrl_replay(handle(model), N, qpos.typed_data(), qvel.typed_data(), obs->typed_data(), reward->typed_data(),
                               done->typed_data(), xpos->typed_data(), xquat->typed_data(), stream));
}
ffi::Error PdImpl(cudaStream_t stream, int64_t model, int32_t pd_mode, int32_t action_cols, float pd_dt, F32 action, F32 qpos, F32 qvel,
                  F32 qpos_lag, F32 qvel_lag, F32 kp, F32 kd, F32 time, Out<F32> tau, Out<F32> qM, Out<F32> qfrc_bias) {
    const int N = (int)qpos.dimensions()[0];
    return fail_if(pbrl_pd_torque(handle(model), N, pd_mode, action_cols, action.typed_data(), qpos.typed_data(), qvel.typed_data(),
                                  qpos_lag.typed_data(), qvel_lag.typed_data(), kp.typed_data(), kd.typed_data(), time.typed_data(), pd_dt,
                                  tau->typed_data(), qM->typed_data(), qfrc_bias->typed_data(), stream));
}
}  // namespace
XLA_FFI_DEFINE_HANDLER_SYMBOL(PbrlReplayFfi, ReplayImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Attr<int64_t>("model")
                                  .Arg<F32>().Arg<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(PbrlPdTorqueFfi, PdImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Attr<int64_t>("model")
                                  .Attr<int32_t>("pd_mode").Attr<int32_t>("action_cols").Attr<float>("pd_dt")
                                  .Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>()
                                  .Ret<F32>().Ret<F32>().Ret<F32>());
#else
// jaxlib headers not found: nothing to compile.
extern "C" int pbrl_ffi_xla_unavailable(void) { return 1; }
#endif
