// XLA FFI shim: exposes pbrl_step / pbrl_reset / pbrl_replay / pbrl_pd_torque as `jax.ffi` custom calls so that the
// reference's JAX host code keeps calling `env.step` while the work runs in libpbrl (include/pbrl.h).
//
// UNVERIFIED: jaxlib's headers (xla/ffi/api/ffi.h) do not exist in the build image, so this translation unit is
// compiled only when they are found (see INTEGRATION.md for the build line).  The batch dimension is inside the call:
// buffers are [N, width] fp32, the model handle travels as an int64 attribute.
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define PBRL_HAVE_XLA_FFI 1
#endif
#endif

#ifdef PBRL_HAVE_XLA_FFI
#include <cuda_runtime.h>

#include "../../include/pbrl.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;
using F32 = ffi::Buffer<ffi::F32>;
using S32 = ffi::Buffer<ffi::S32>;

static ffi::Error fail_if(int rc) {
    return rc == 0 ? ffi::Error::Success() : ffi::Error(ffi::ErrorCode::kInternal, pbrl_last_error());
}

// state buffers are passed as operands and aliased to the results by the Python side (input_output_aliases), so
// the kernels update them in place exactly as the C ABI specifies.
static ffi::Error StepImpl(cudaStream_t stream, int64_t model, int32_t n_frames, int32_t pd_mode, int32_t action_cols,
                           float pd_dt, int32_t reward_mode, F32 action, F32 kp, F32 kd, F32 qpos, F32 qvel, F32 time,
                           F32 warm, F32 ctrl, F32 qpos_lag, F32 qvel_lag, S32 motion_id, F32 ref_t0,
                           ffi::Result<F32> obs, ffi::Result<F32> reward, ffi::Result<F32> done, ffi::Result<F32> metrics,
                           ffi::Result<F32> stats) {
    const int N = (int)qpos.dimensions()[0];
    PbrlBuffers b{};
    b.qpos = qpos.typed_data(); b.qvel = qvel.typed_data(); b.time = time.typed_data(); b.warm = warm.typed_data();
    b.ctrl = ctrl.typed_data(); b.qpos_lag = qpos_lag.typed_data(); b.qvel_lag = qvel_lag.typed_data();
    b.obs = obs->typed_data(); b.reward = reward->typed_data(); b.done = done->typed_data();
    b.metrics = metrics->typed_data(); b.stats = stats->typed_data(); b.status = nullptr;
    b.motion_id = motion_id.typed_data(); b.ref_t0 = ref_t0.typed_data();
    return fail_if(pbrl_step(reinterpret_cast<const PbrlModel*>(model), N, n_frames, pd_mode, action_cols, action.typed_data(),
                             kp.typed_data(), kd.typed_data(), pd_dt, nullptr, 0x0FFFFFFFu, reward_mode, &b, stream));
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(PbrlStepFfi, StepImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("model").Attr<int32_t>("n_frames").Attr<int32_t>("pd_mode")
                                  .Attr<int32_t>("action_cols").Attr<float>("pd_dt").Attr<int32_t>("reward_mode")
                                  .Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>()
                                  .Arg<F32>().Arg<F32>().Arg<S32>().Arg<F32>()
                                  .Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>());

static ffi::Error ReplayImpl(cudaStream_t stream, int64_t model, F32 qpos, F32 qvel, ffi::Result<F32> obs,
                             ffi::Result<F32> reward, ffi::Result<F32> done, ffi::Result<F32> xpos, ffi::Result<F32> xquat) {
    const int N = (int)qpos.dimensions()[0];
    return fail_if(pbrl_replay(reinterpret_cast<const PbrlModel*>(model), N, qpos.typed_data(), qvel.typed_data(),
                               obs->typed_data(), reward->typed_data(), done->typed_data(), xpos->typed_data(),
                               xquat->typed_data(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(PbrlReplayFfi, ReplayImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Attr<int64_t>("model")
                                  .Arg<F32>().Arg<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>());

static ffi::Error PdImpl(cudaStream_t stream, int64_t model, int32_t pd_mode, int32_t action_cols, float pd_dt, F32 action,
                         F32 qpos, F32 qvel, F32 qpos_lag, F32 qvel_lag, F32 kp, F32 kd, F32 time, ffi::Result<F32> tau) {
    const int N = (int)qpos.dimensions()[0];
    return fail_if(pbrl_pd_torque(reinterpret_cast<const PbrlModel*>(model), N, pd_mode, action_cols, action.typed_data(),
                                  qpos.typed_data(), qvel.typed_data(), qpos_lag.typed_data(), qvel_lag.typed_data(),
                                  kp.typed_data(), kd.typed_data(), time.typed_data(), pd_dt, tau->typed_data(), nullptr,
                                  nullptr, stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(PbrlPdTorqueFfi, PdImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Attr<int64_t>("model")
                                  .Attr<int32_t>("pd_mode").Attr<int32_t>("action_cols").Attr<float>("pd_dt")
                                  .Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>()
                                  .Ret<F32>());
#else
// jaxlib headers not found: nothing to compile.
extern "C" int pbrl_ffi_xla_unavailable(void) { return 1; }
#endif
