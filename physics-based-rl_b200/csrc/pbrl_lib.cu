// libpbrl.so: CUDA kernels (sm_100a) + the C ABI declared in include/pbrl.h.
//
// One warp per environment; WARPS_PER_CTA warps per CTA, each with its own shared-memory workspace
// (pbrl_core.cuh: Ws).  Grids are sized to the resident capacity of the device (SM count x CTAs/SM) and
// every warp walks a grid-stride loop over environments, so any N runs with full SMs and no tail CTA.
#include <cuda_runtime.h>

#include <atomic>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "pbrl_core.cuh"
#include "pbrl_mimic.cuh"
#include "pbrl_desc_reflect.h"

using namespace pbrl;

namespace {

// One CTA of 16 warps per SM, all in one phase-lock group: the kernel is bound by instruction fetch (about 200 KB of
// straight-line code per substep against a 32 KB L1.5 I-cache), so a single instruction stream per SM wins.
#ifndef PBRL_WARPS_PER_CTA
#define PBRL_WARPS_PER_CTA 16
#endif
constexpr int WARPS_PER_CTA = PBRL_WARPS_PER_CTA;
constexpr int FUSED_CLASSES = 16;   // work classes of the fused two-part step (k_step)
constexpr int CTA_THREADS = 32 * WARPS_PER_CTA;
static_assert(WARPS_PER_CTA % PBRL_LOCK_GROUP == 0, "lock groups tile the CTA");

// last error: per thread (the message of the calling thread's own failed call), with the process-wide last message behind it for
// a caller that asks from another thread than the one that made the call
thread_local std::string g_error;
std::mutex g_error_mu;
std::string g_error_any;
thread_local std::string g_error_ret;
int fail(const std::string& msg) {
    g_error = msg;
    std::lock_guard<std::mutex> lk(g_error_mu);
    g_error_any = msg;
    return -1;
}
#define CUDA_OK(call)                                                                              \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return fail(std::string(#call) + ": " + cudaGetErrorString(e_));    \
    } while (0)

// Entry points launch on the model's device whatever the caller's current device is, and leave the caller's device as
// they found it (torch streams of device k are usable only while device k is current).
struct DeviceGuard {
    int prev = -1;
    bool ok = true, changed = false;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != device) { changed = true; if (cudaSetDevice(device) != cudaSuccess) ok = false; }
    }
    ~DeviceGuard() { if (changed) cudaSetDevice(prev); }
};
#define ON_DEVICE(m)                                                  \
    DeviceGuard guard_((m)->device);                                  \
    if (!guard_.ok) return fail("cannot select the model's CUDA device")

// dynamic shared memory of a CTA: [DevModel copy][Ws x WARPS_PER_CTA]
constexpr size_t MODEL_SMEM = (sizeof(DevModel) + 15) / 16 * 16;
__device__ __forceinline__ Ws& my_ws() {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    return *reinterpret_cast<Ws*>(smem_raw + MODEL_SMEM + (size_t)(threadIdx.x >> 5) * sizeof(Ws));
}
// every CTA works from its own shared-memory copy of the model tables
__device__ __forceinline__ const DevModel& stage_model(const DevModel* __restrict__ gm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const unsigned* src = reinterpret_cast<const unsigned*>(gm);
    unsigned* dst = reinterpret_cast<unsigned*>(smem_raw);
    for (int i = threadIdx.x; i < (int)(sizeof(DevModel) / 4); i += blockDim.x) dst[i] = src[i];
    __syncthreads();
    return *reinterpret_cast<const DevModel*>(smem_raw);
}
__device__ __forceinline__ int first_env() { return blockIdx.x * WARPS_PER_CTA + (threadIdx.x >> 5); }
__device__ __forceinline__ int env_stride() { return gridDim.x * WARPS_PER_CTA; }

// Groups of WARPS_PER_CTA envs are handed out dynamically (one atomic per group): with the balanced processing order the
// groups arrive most expensive first, so the CTAs finish within one cheap group of each other (longest-processing-time-first);
// a static round-robin gives CTA 0 the most expensive group of every pass.  `sched` = {next group, CTAs finished}: the last
// CTA to leave re-arms the pair for the next launch.
#ifdef PBRL_STEP_MAXNREG   // dev sweeps: cap k_step's registers below what the launch bound allows
#define PBRL_STEP_BOUNDS __maxnreg__(PBRL_STEP_MAXNREG)
#else
#define PBRL_STEP_BOUNDS __launch_bounds__(CTA_THREADS)
#endif
__global__ void PBRL_STEP_BOUNDS k_step(const DevModel* __restrict__ gm, StepArgs a, int* __restrict__ sched) {
    Ws& w = my_ws();
    const int lane = threadIdx.x & 31;
    const DevModel& sm = stage_model(gm);
    init_ws_tables(sm, w, lane);
    frames_init(w, lane);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    __shared__ int s_unit;
    const int ngroups = (a.N + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
    if (a.fused_parts > 1) {
        // ---- the env step in P parts inside ONE launch (regrouping without the drain of a launch boundary) -------------------
        // Part 0 (PD law + the first substeps) is handed out in groups of the caller's order, heaviest first, as below.  An env
        // that finished part p - 1 is appended to queue p under its work class (the solver work of its last substep); a CTA that
        // finds no part-0 group left draws the 16 heaviest envs on offer in the earliest queue that offers 16 (the longest
        // remaining work first), and whatever is left in a queue once every env has been through the part before it.
        // Part p runs the substeps [p n / P, (p + 1) n / P); the last one ends with the epilogue.
        // sched: [0] next part-0 group; per queue p = 1 .. P - 1 at base(p) = 4 + (p - 1)(1 + 2 NCLS): [base] envs appended,
        // [base + 1 + c] entries appended to class c, [base + 1 + NCLS + c] entries of class c taken.  Queue entries are env + 1
        // (0 = reserved, not yet written).  The host zeroes sched and the queues before the launch.
        constexpr int NCLS = FUSED_CLASSES;
        const int P = a.fused_parts;
        __shared__ int s_kind, s_envs[WARPS_PER_CTA];
        volatile int* vs = sched;
        const int wid = threadIdx.x >> 5;
        auto base = [&](int q) { return 4 + (q - 1) * (1 + 2 * NCLS); };
        auto queue = [&](int q) { return a.fused_queue + (size_t)(q - 1) * NCLS * a.N; };
        auto bound = [&](int q) { return (int)((long long)q * a.n_frames / P); };
        for (;;) {
            __syncthreads();
            if (threadIdx.x == 0) {
                int kind = 0;   // 0: done, 1 + p: a unit of part p
                const int u = vs[0] < ngroups ? atomicAdd(&sched[0], 1) : ngroups;
                if (u < ngroups) { kind = 1; s_unit = u; }
                else {
                    for (int spin = 0; kind == 0; spin++) {
                        bool finished = true;
                        for (int q = 1; q < P && kind == 0; q++) {
                            const int b = base(q);
                            const bool all = vs[b] >= a.N;   // (read first: what is in the queue now is at least what was there then)
                            int total = 0;
                            for (int c = 0; c < NCLS; c++) total += vs[b + 1 + c] - vs[b + 1 + NCLS + c];
                            if (!all || total > 0) finished = false;
                            if (!(total >= WARPS_PER_CTA || (all && total > 0))) continue;
                            int got = 0;
                            for (int c = NCLS - 1; c >= 0 && got < WARPS_PER_CTA; c--) {
                                for (;;) {
                                    const int h = vs[b + 1 + NCLS + c], avail = vs[b + 1 + c] - h, want = WARPS_PER_CTA - got;
                                    if (avail <= 0) break;
                                    const int take = avail < want ? avail : want;
                                    if (atomicCAS(&sched[b + 1 + NCLS + c], h, h + take) != h) continue;   // another CTA drew from this class: look again
                                    for (int k = 0; k < take; k++) {
                                        volatile int* slot = queue(q) + (size_t)c * a.N + h + k;
                                        int e;
                                        while ((e = *slot) == 0) {}   // reserved by its writer, lands within a few hundred cycles
                                        s_envs[got + k] = e - 1;
                                    }
                                    got += take;
                                    break;
                                }
                            }
                            if (got > 0) { for (int k = got; k < WARPS_PER_CTA; k++) s_envs[k] = -1; kind = 1 + q; }
                        }
                        if (kind != 0 || finished) break;   // finished: every env is through every part but the last and every queue is empty
                        __nanosleep(spin < 8 ? 200 : 1000);
                    }
                    __threadfence();   // the envs' state of the part before is visible before this CTA loads it
                }
                s_kind = kind;
            }
            __syncthreads();
            const int kind = s_kind;
            if (kind == 0) break;
            const int part = kind - 1;
            int env = -1;
            if (part == 0) {
                const int slot = s_unit * WARPS_PER_CTA + wid;
                if (slot < a.N) env = a.buf.env_order ? a.buf.env_order[slot] : slot;
            } else env = s_envs[wid];
            const int fb = bound(part), fe = bound(part + 1);
            if (env >= 0) {
                env_step(sm, w, a, env, acc, lane, fb, fe, part > 0);
                if (part + 1 < P) {
                    __threadfence();   // this env's state before its queue entry
                    if (lane == 0) {
                        int c = w.work;   // the work of the LAST substep of the part: the best predictor of the next one (+0.5 % over the maximum)
                        c = c < 0 ? 0 : (c >= NCLS ? NCLS - 1 : c);
                        const int b = base(part + 1);
                        const int at = atomicAdd(&sched[b + 1 + c], 1);
                        queue(part + 1)[(size_t)c * a.N + at] = env + 1;
                        __threadfence();
                        atomicAdd(&sched[b], 1);
                    }
                }
            } else {
                for (int f = fb; f < fe; f++) {
                    for (int k = 0; k < __popc(a.phase_lock & 39); k++) CTA_SYNC();
                    if (a.phase_lock & 16) while (CTA_OR(0)) {}
                    if (a.phase_lock & 8) CTA_SYNC();
                }
            }
        }
        if (a.buf.stats && a.reward_mode == PBRL_REWARD_MIMIC && lane < 6 && acc[5] > 0.f) {
            float v = acc[0];
#pragma unroll
            for (int k = 1; k < 6; k++) v = lane == k ? acc[k] : v;
            atomicAdd(a.buf.stats + lane, v);
        }
        return;
    }
    // A batch with fewer groups than CTAs (64 envs = 4 groups on 148 SMs) is handed out in `split` parts of WARPS_PER_CTA / split
    // envs per group, so that the idle SMs share the work: a CTA with 8 (4, 1) active warps steps its envs in 0.79 (0.6, 0.45) of
    // the time 16 take.  split = the largest power of two with split x groups <= gridDim; the other warps of such a CTA mirror
    // the barrier sequence like the warps past the last env.  (Splitting the unfilled LAST wave of a large batch the same way
    // measured 4 % slower at 8 192 envs: the CTAs finish staggered and the dynamic hand-out already fills the gaps.)
    const int rem = ngroups < (int)gridDim.x ? ngroups : 0;
    int split = 1;
    while (a.tail_split && rem > 0 && split < WARPS_PER_CTA && WARPS_PER_CTA % (2 * split) == 0 && 2 * split * rem <= (int)gridDim.x) split *= 2;
    const int nwhole = ngroups - rem, nunits = nwhole + rem * split, per_part = WARPS_PER_CTA / split;
    // every warp of the CTA makes the same passes so that the per-substep CTA barrier (phase lock) is legal
    for (;;) {
        __syncthreads();  // (the previous pass has read s_unit)
        if (threadIdx.x == 0) s_unit = atomicAdd(&sched[0], 1);
        __syncthreads();
        const int unit = s_unit;
        if (unit >= nunits) break;
        const int wid = threadIdx.x >> 5;
        int group = unit;
        bool mine = true;
        if (unit >= nwhole) {
            group = nwhole + (unit - nwhole) / split;
            mine = wid / per_part == (unit - nwhole) % split;
        }
        const int slot = group * WARPS_PER_CTA + wid;
        if (mine && slot < a.N) env_step(sm, w, a, a.buf.env_order ? a.buf.env_order[slot] : slot, acc, lane, a.f_begin, a.f_end, false);
        else {  // no env for this warp in this pass: mirror the barrier sequence of a substep (forward / solve_constraints / substep)
            for (int f = a.f_begin; f < a.f_end; f++) {
                for (int k = 0; k < __popc(a.phase_lock & 39); k++) CTA_SYNC();
                if (a.phase_lock & 16) while (CTA_OR(0)) {}
                if (a.phase_lock & 8) CTA_SYNC();
            }
        }
    }
    if (threadIdx.x == 0 && atomicAdd(&sched[1], 1) == (int)gridDim.x - 1) {  // every CTA has drawn its terminal index
        sched[0] = 0;
        sched[1] = 0;
        __threadfence();
    }
    if (a.buf.stats && a.reward_mode == PBRL_REWARD_MIMIC && lane < 6 && acc[5] > 0.f) {
        float v = acc[0];  // one atomic per (warp, statistic): lane k publishes statistic k
#pragma unroll
        for (int k = 1; k < 6; k++) v = lane == k ? acc[k] : v;
        atomicAdd(a.buf.stats + lane, v);
    }
}
__global__ void __launch_bounds__(CTA_THREADS) k_reset(const DevModel* __restrict__ gm, StepArgs a) {
    Ws& w = my_ws();
    const int lane = threadIdx.x & 31;
    const DevModel& sm = stage_model(gm);
    if (a.reward_mode == PBRL_REWARD_MIMIC && a.buf.motion_id) {
        init_ws_tables(sm, w, lane);
        frames_init(w, lane);
        for (int env = first_env(); env < a.N; env += env_stride()) env_reset_mimic(sm, a.mm, a.frames, w, a.buf, env, lane);
        return;
    }
    for (int env = first_env(); env < a.N; env += env_stride()) env_reset(sm, w, a, env, lane);
}
// counting sort of the env ids by decreasing work estimate (one CTA; N <= a few 100k)
constexpr int BO_THREADS = 1024, BO_BINS = 256;
__global__ void __launch_bounds__(BO_THREADS) k_balance_order(int N, const int* __restrict__ cost, int* __restrict__ order) {
    __shared__ int hist[BO_BINS], start[BO_BINS];
    for (int b = threadIdx.x; b < BO_BINS; b += BO_THREADS) hist[b] = 0;
    __syncthreads();
    for (int e = threadIdx.x; e < N; e += BO_THREADS) atomicAdd(&hist[min(max(cost[e], 0), BO_BINS - 1)], 1);
    __syncthreads();
    if (threadIdx.x == 0) {  // 256 bins: a serial scan is a few hundred cycles
        int acc = 0;
        for (int b = BO_BINS - 1; b >= 0; b--) { start[b] = acc; acc += hist[b]; }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < N; e += BO_THREADS) order[atomicAdd(&start[min(max(cost[e], 0), BO_BINS - 1)], 1)] = e;
}
__global__ void __launch_bounds__(CTA_THREADS)
k_autoreset(const DevModel* __restrict__ gm, StepArgs a, const float* done, const float* next_ref_t0, int* steps) {
    Ws& w = my_ws();
    const int lane = threadIdx.x & 31;
    const DevModel& sm = stage_model(gm);
    init_ws_tables(sm, w, lane);
    frames_init(w, lane);
    for (int env = first_env(); env < a.N; env += env_stride())
        if (done[env] != 0.f) env_autoreset(sm, a.mm, a.frames, w, a, next_ref_t0, steps, env, lane);
}
__global__ void __launch_bounds__(CTA_THREADS)
k_replay(const DevModel* __restrict__ gm, int N, const float* qpos, const float* qvel, float* obs, float* reward,
         float* done, float* xpos, float* xquat) {
    Ws& w = my_ws();
    const int lane = threadIdx.x & 31;
    const DevModel& sm = stage_model(gm);
    init_ws_tables(sm, w, lane);
    for (int env = first_env(); env < N; env += env_stride())
        env_replay(sm, w, qpos, qvel, obs, reward, done, xpos, xquat, env, lane);
}
// Replay rollout: HBM-bound (944 B out per (env, frame) item against ~70 warp instructions), so the mapping is the
// opposite of the dynamics kernels: one THREAD per item.  The item's body poses (xpos 42 | xquat 56 floats) are staged in a
// shared-memory tile with an odd row stride (conflict-free column access) and written out by the whole CTA as contiguous
// 16-byte vectors (the rows of RR_ITEMS consecutive items are contiguous in every output array); the observation rows are
// copied straight from the clip (a few KB, L1-resident).  The model tables are read through L1 with warp-uniform loads:
// no per-CTA model copy, so four CTAs (16 warps) fit an SM and hide the latency of the per-thread kinematic chain.
#ifndef PBRL_RR_ITEMS
#define PBRL_RR_ITEMS 32
#endif
constexpr int RR_ITEMS = PBRL_RR_ITEMS;    // items (= threads) per CTA
constexpr int RR_ROW = NB * 3 + NB * 4 + 1;  // 99 floats, odd
static_assert(RR_ROW % 2 == 1, "odd row stride keeps per-thread rows on distinct banks");
constexpr size_t RR_SMEM = sizeof(float) * RR_ITEMS * RR_ROW + sizeof(int) * RR_ITEMS;
// tile columns [col0, col0 + W) of `nitems` rows -> out rows of width W starting at item g0, as VEC-float vectors
// (W % VEC == 0: a vector never straddles two rows; the division by the constant W / VEC is a multiply-shift)
template <int W, int VEC>
__device__ __forceinline__ void rr_store(float* __restrict__ out, const float* tile, long long g0, int nitems, int col0) {
    static_assert(W % VEC == 0 && (VEC == 2 || VEC == 4), "vector width");
    if (!out) return;
    constexpr int PER_ROW = W / VEC;
    float* dst = out + g0 * W;  // 16-byte aligned: g0 is a multiple of RR_ITEMS
    for (int v = threadIdx.x; v < nitems * PER_ROW; v += RR_ITEMS) {
        const int row = v / PER_ROW, s = v - row * PER_ROW;
        const float* src = tile + row * RR_ROW + col0 + VEC * s;
        if constexpr (VEC == 4) __stcs(reinterpret_cast<float4*>(dst) + v, make_float4(src[0], src[1], src[2], src[3]));  // streaming: never re-read
        else __stcs(reinterpret_cast<float2*>(dst) + v, make_float2(src[0], src[1]));
    }
}
template <bool AXSPEC>
__global__ void __launch_bounds__(RR_ITEMS, 512 / RR_ITEMS)
k_replay_rollout(const __grid_constant__ ReplayTables T, int N, int K, const float* __restrict__ data_pos, const float* __restrict__ data_vel,
                 int F, const int* __restrict__ frame0, float* obs, float* reward, float* done, float* xpos, float* xquat) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* tile = reinterpret_cast<float*>(smem_raw);
    int* frame = reinterpret_cast<int*>(smem_raw + sizeof(float) * RR_ITEMS * RR_ROW);
    const long long total = (long long)N * K;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // per-lane source columns of an observation row [qpos 35 | qvel 34] (byte pointers into frame 0; a frame adds f * row bytes)
    const char* src_a = reinterpret_cast<const char*>(data_pos + lane);
    const bool b_in_qpos = lane < NQ - 32;
    const char* src_b = b_in_qpos ? reinterpret_cast<const char*>(data_pos + 32 + lane) : reinterpret_cast<const char*>(data_vel + lane - (NQ - 32));
    const unsigned stride_b = b_in_qpos ? NQ * 4 : NV * 4;
    const char* src_c = reinterpret_cast<const char*>(data_vel + 64 - NQ + lane);
    for (long long g0 = (long long)blockIdx.x * RR_ITEMS; g0 < total; g0 += (long long)gridDim.x * RR_ITEMS) {
        const long long g = g0 + threadIdx.x;
        const int nitems = total - g0 < RR_ITEMS ? (int)(total - g0) : RR_ITEMS;
        if (g < total) {
            const int k = (int)(g / N), e = (int)(g - (long long)k * N);
            const int f = ((frame0 ? frame0[e] : e) + k) % F;
            frame[threadIdx.x] = f;
            float* row = tile + threadIdx.x * RR_ROW;
            if (xpos || xquat) replay_item_fk<AXSPEC>(T, data_pos + (size_t)f * NQ, row, row + NB * 3);
        }
        __syncthreads();
        if (obs) {
            // obs = [qpos, qvel] of the item's frame (agent_env_template.py:151-155).  One warp per row: the frame index and the two
            // source rows are warp-uniform, a lane's column is fixed, so a row costs three loads (the clip is L1-resident) and
            // three 128-byte streaming stores; consecutive rows are contiguous in the output.
            static_assert(NQ == 35 && NV == 34 && PBRL_OBS == 69, "column split of the three 32-lane chunks");
            float* dst = obs + (g0 + warp) * PBRL_OBS + lane;
#pragma unroll 4
            for (int r = warp; r < nitems; r += RR_ITEMS / 32, dst += (RR_ITEMS / 32) * PBRL_OBS) {
                const int f = frame[r];
                const float a = __ldg(reinterpret_cast<const float*>(src_a + (size_t)f * (NQ * 4)));      // columns 0..31
                const float b = __ldg(reinterpret_cast<const float*>(src_b + (size_t)f * stride_b));      // columns 32..63: qpos 32..34, qvel 0..28
                float c = 0.f;
                if (lane < PBRL_OBS - 64) c = __ldg(reinterpret_cast<const float*>(src_c + (size_t)f * (NV * 4)));  // columns 64..68: qvel 29..33
                __stcs(dst, a);
                __stcs(dst + 32, b);
                if (lane < PBRL_OBS - 64) __stcs(dst + 64, c);
            }
        }
        rr_store<NB * 3, 2>(xpos, tile, g0, nitems, 0);
        rr_store<NB * 4, 4>(xquat, tile, g0, nitems, NB * 3);
        if (reward) for (int i = threadIdx.x; i < nitems * 3; i += RR_ITEMS) reward[g0 * 3 + i] = 0.f;   // agent_replay_motion.py: zeros(3)
        if (done && threadIdx.x < nitems) done[g0 + threadIdx.x] = 1.0f;
        __syncthreads();
    }
}
// Replay rollout through a frame table.  An item's poses depend on its clip frame only, and a rollout of N x K items walks a
// clip of F rows (39 for the walk clip, 697 for all seven): when F is much smaller than N x K the kinematics are evaluated once
// per FRAME per launch (k_replay_frames: one thread per frame, the same replay_item_fk arithmetic, so the rows are bit-identical
// to the per-item kernel's) and the rollout itself becomes what its byte count says it is, a row gather: k_replay_gather copies,
// per item, the clip row into obs and the table row into xpos / xquat, one warp per item, four items in flight per warp.
// The table (F x 100 floats, L1 / L2-resident) lives in a scratch buffer of the model.
constexpr int RT_ROW = 100;           // xpos 42 | 2 pad | xquat 56: 16-byte aligned halves
constexpr int RT_QUAT = 44;
constexpr int RB_PARTS = 16, RB_CLAIM_STRIDE = 32;   // block counters of a k_replay_gather_bulk launch: one per range of blocks, 128 bytes apart
template <bool AXSPEC>
__global__ void __launch_bounds__(64) k_replay_frames(const __grid_constant__ ReplayTables T, int F, const float* __restrict__ data_pos, float* __restrict__ tab, unsigned* claim) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    // the gather behind this kernel is launched programmatically dependent: its CTAs may become resident and run their prologue now;
    // they wait (griddepcontrol.wait) for this grid to have completed before they touch the counters or the table
    asm volatile("griddepcontrol.launch_dependents;");
    if (claim && f < RB_PARTS) claim[f * RB_CLAIM_STRIDE] = 0u;   // the block counters of the k_replay_gather_bulk launch behind this one (blockDim 64 >= RB_PARTS)
    if (f < F) replay_item_fk<AXSPEC>(T, data_pos + (size_t)f * NQ, tab + (size_t)f * RT_ROW, tab + (size_t)f * RT_ROW + RT_QUAT);
}
constexpr int RG_THREADS = 256, RG_UNROLL = 4;
// ALL: every output array is present (the fast path: no per-array tests, output pointers advance by a constant per pass and
// every store of the four items has an immediate offset)
template <bool ALL>
__global__ void __launch_bounds__(RG_THREADS)
k_replay_gather(int N, int K, const float* __restrict__ data_pos, const float* __restrict__ data_vel, int F, const int* __restrict__ frame0,
                const float* __restrict__ tab, float* obs, float* reward, float* done, float* xpos, float* xquat) {
    static_assert(NQ == 35 && NV == 34 && PBRL_OBS == 69 && NB * 3 == 42 && NB * 4 == 56, "lane / column split of the rows");
    const long long total = (long long)N * K;
    const int lane = threadIdx.x & 31;
    const long long warp = (long long)blockIdx.x * (RG_THREADS / 32) + (threadIdx.x >> 5), nwarps = (long long)gridDim.x * (RG_THREADS / 32);
    const bool b_in_qpos = lane < NQ - 32;
    const bool has_obs = ALL || obs, has_pos = ALL || xpos, has_quat = ALL || xquat;
    // (step, env) of the warp's first item and its increments per pass: no division inside the loop (a 64-bit division per item,
    // executed by every warp, cost more issue slots than the copy itself)
    const long long step = nwarps * RG_UNROLL;
    long long g0 = warp * RG_UNROLL;
    int e0 = (int)(g0 % N), kf0 = (int)((g0 / N) % F);
    const int de = (int)(step % N), dkf = (int)((step / N) % F);
    // per-lane source columns (element offsets inside a clip row / table row) and output cursors
    const int col_b = b_in_qpos ? 32 + lane : lane - (NQ - 32);
    const float* src_b = b_in_qpos ? data_pos : data_vel;
    const int stride_b = b_in_qpos ? NQ : NV;
    float* o_obs = obs ? obs + g0 * PBRL_OBS + lane : nullptr;
    float* o_pos = xpos ? xpos + g0 * 42 + lane : nullptr;
    float* o_quat = xquat ? xquat + g0 * 56 + lane : nullptr;
    float* o_rew = reward ? reward + g0 * 3 + lane : nullptr;
    float* o_done = done ? done + g0 + lane : nullptr;
    for (; g0 < total; g0 += step) {
        int f[RG_UNROLL];
        float oa[RG_UNROLL], ob[RG_UNROLL], oc[RG_UNROLL], pa[RG_UNROLL], pb[RG_UNROLL], qa[RG_UNROLL], qb[RG_UNROLL];
#pragma unroll
        for (int u = 0; u < RG_UNROLL; u++) {
            int e = e0 + u, kf = kf0;
            while (e >= N) { e -= N; kf = kf + 1 == F ? 0 : kf + 1; }   // (at most one pass unless N < RG_UNROLL)
            // past the end (the last pass of the last warps) the loads repeat a valid row; nothing is stored for them
            f[u] = (int)(((unsigned)(frame0 ? __ldg(frame0 + e) : e) + (unsigned)kf) % (unsigned)F);
        }
        e0 += de; kf0 += dkf;
        if (e0 >= N) { e0 -= N; kf0++; }
        if (kf0 >= F) kf0 -= F;
#pragma unroll
        for (int u = 0; u < RG_UNROLL; u++) {  // every load of the four items is in flight before the first store
            const float* tr = tab + f[u] * RT_ROW + lane;
            if (has_obs) {
                oa[u] = __ldg(data_pos + f[u] * NQ + lane);                                   // obs columns 0..31
                ob[u] = __ldg(src_b + f[u] * stride_b + col_b);                               // 32..63: qpos 32..34, qvel 0..28
                oc[u] = lane < PBRL_OBS - 64 ? __ldg(data_vel + f[u] * NV + 64 - NQ + lane) : 0.f;   // 64..68: qvel 29..33
            }
            if (has_pos) { pa[u] = __ldg(tr); pb[u] = lane < 42 - 32 ? __ldg(tr + 32) : 0.f; }
            if (has_quat) { qa[u] = __ldg(tr + RT_QUAT); qb[u] = lane < 56 - 32 ? __ldg(tr + RT_QUAT + 32) : 0.f; }
        }
        const bool full = g0 + RG_UNROLL <= total;
#pragma unroll
        for (int u = 0; u < RG_UNROLL; u++) {
            if (!full && g0 + u >= total) break;
            if (has_obs) {
                __stcs(o_obs + u * PBRL_OBS, oa[u]);
                __stcs(o_obs + u * PBRL_OBS + 32, ob[u]);
                if (lane < PBRL_OBS - 64) __stcs(o_obs + u * PBRL_OBS + 64, oc[u]);
            }
            if (has_pos) {
                __stcs(o_pos + u * 42, pa[u]);
                if (lane < 42 - 32) __stcs(o_pos + u * 42 + 32, pb[u]);
            }
            if (has_quat) {
                __stcs(o_quat + u * 56, qa[u]);
                if (lane < 56 - 32) __stcs(o_quat + u * 56 + 32, qb[u]);
            }
        }
        // reward (agent_replay_motion.py: zeros(3)) and done (ones) of the four items: 12 + 4 contiguous floats
        if (ALL || reward) { if (lane < (full ? 3 * RG_UNROLL : 3 * (int)(total - g0))) __stcs(o_rew, 0.f); }
        if (ALL || done) { if (lane < (full ? RG_UNROLL : (int)(total - g0))) __stcs(o_done, 1.0f); }
        if (ALL || obs) o_obs += step * PBRL_OBS;
        if (ALL || xpos) o_pos += step * 42;
        if (ALL || xquat) o_quat += step * 56;
        if (ALL || reward) o_rew += step * 3;
        if (ALL || done) o_done += step;
    }
}
// The same gather with the stores handed to the bulk-copy engine.  k_replay_gather is bound by its own store instructions (nine
// 4-byte-per-lane stores per item, rows that start on any 4-byte boundary: lg_throttle is its top stall, 3.4 sectors per request).
// Here a warp assembles the rows of RB_ITEMS consecutive items in shared memory -- consecutive items are contiguous in every
// output array, and with 16 items every block of every array starts and ends on a 64-byte boundary (16 x 276 / 168 / 224 / 12 /
// 4 bytes), so no 32-byte sector is written in part -- and one lane hands the five blocks to cp.async.bulk.global.shared::cta.
// reward (zeros) and done (ones) never change: their blocks are written to shared memory once per CTA.
// The last block of a rollout whose item count is not a multiple of RB_ITEMS leaves through ordinary stores.
#ifndef PBRL_REPLAY_BULK_DEFAULT
#define PBRL_REPLAY_BULK_DEFAULT 1
#endif
#ifndef PBRL_RB_WARPS   // warps per CTA, items whose loads are in flight together, resident CTAs per SM the registers are capped for
#define PBRL_RB_WARPS 8
#define PBRL_RB_CHUNK 8
#define PBRL_RB_CTAS 2
#endif
constexpr int RB_ITEMS = 16, RB_WARPS = PBRL_RB_WARPS, RB_THREADS = 32 * RB_WARPS, RB_CHUNK = PBRL_RB_CHUNK;
constexpr int RB_OBS = RB_ITEMS * PBRL_OBS, RB_POS = RB_ITEMS * 42, RB_QUAT = RB_ITEMS * 56, RB_STAGE = RB_OBS + RB_POS + RB_QUAT;
constexpr int RB_CONST = RB_ITEMS * 4;   // floats: reward block (3 per item) | done block (1 per item)
constexpr size_t RB_SMEM = sizeof(float) * (RB_CONST + RB_WARPS * RB_STAGE);
static_assert(RB_PARTS <= 64, "k_replay_frames zeroes the counters with the threads of its first CTA");
static_assert((RB_OBS * 4) % 16 == 0 && (RB_POS * 4) % 16 == 0 && (RB_QUAT * 4) % 16 == 0 && (RB_STAGE * 4) % 16 == 0 && (RB_CONST * 4) % 16 == 0 &&
              (RB_ITEMS * 3 * 4) % 16 == 0, "bulk copies move 16-byte multiples between 16-byte aligned addresses");
__device__ __forceinline__ void bulk_store(float* dst, const float* src_smem, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(__cvta_generic_to_global(dst)), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
__global__ void __launch_bounds__(RB_THREADS, PBRL_RB_CTAS)
k_replay_gather_bulk(int N, int K, const float* __restrict__ data_pos, const float* __restrict__ data_vel, int F, const int* __restrict__ frame0,
                     const float* __restrict__ tab, unsigned* claim, float* obs, float* reward, float* done, float* xpos, float* xquat) {
    static_assert(NQ == 35 && NV == 34 && PBRL_OBS == 69 && NB * 3 == 42 && NB * 4 == 56, "lane / column split of the rows");
    extern __shared__ __align__(128) unsigned char rb_smem[];
    float* s_const = reinterpret_cast<float*>(rb_smem);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float* s_obs = s_const + RB_CONST + wid * RB_STAGE;
    float* s_pos = s_obs + RB_OBS;
    float* s_quat = s_pos + RB_POS;
    if (threadIdx.x < RB_CONST) s_const[threadIdx.x] = threadIdx.x < RB_ITEMS * 3 ? 0.f : 1.0f;   // agent_replay_motion.py: zeros(3), ones
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    asm volatile("griddepcontrol.wait;" ::: "memory");   // k_replay_frames (same stream, just ahead) has written the table and zeroed the counters
    const long long total = (long long)N * K, nblocks = (total + RB_ITEMS - 1) / RB_ITEMS;
    const long long warp = (long long)blockIdx.x * RB_WARPS + wid;
    const bool b_in_qpos = lane < NQ - 32;
    // Blocks are claimed, not dealt: the SMs drain their write queues at different rates (an even deal left them busy for 77 - 96 %
    // of the launch).  One counter for the whole launch would be one address taking an atomic every 1.7 ns -- as fast as an L2 slice
    // serves them, and slower than that when the slice sits on the other die -- so the blocks are split into RB_PARTS ranges with a
    // counter each (128 bytes apart, zeroed by k_replay_frames); a warp claims from its own range and moves on to the next one
    // when that is used up.  The claim for the pass after this one is in flight while this one runs.
    const long long per_part = (nblocks + RB_PARTS - 1) / RB_PARTS;
    int part = (int)(warp % RB_PARTS), dry = 0;
    unsigned claimed = 0;
    if (lane == 0) claimed = atomicAdd(claim + part * RB_CLAIM_STRIDE, 1u);
    for (;;) {
        const unsigned mine = __shfl_sync(0xffffffffu, claimed, 0);
        const long long first = part * per_part, left = nblocks - first, size = left < per_part ? left : per_part;
        if ((long long)mine >= size) {   // this range is used up (for good): try the next one, stop after a full round
            if (++dry == RB_PARTS) break;
            part = part + 1 == RB_PARTS ? 0 : part + 1;
            if (lane == 0) claimed = atomicAdd(claim + part * RB_CLAIM_STRIDE, 1u);
            continue;
        }
        dry = 0;
        if (lane == 0) claimed = atomicAdd(claim + part * RB_CLAIM_STRIDE, 1u);
        const long long g0 = (first + mine) * RB_ITEMS;
        // (step, env) of the block's first item: one division pair per 16 items (32-bit whenever the rollout allows it)
        int e0, kf0;
        if (total <= 0x7fffffffll) { const unsigned g = (unsigned)g0, k = g / (unsigned)N; e0 = (int)(g - k * (unsigned)N); kf0 = (int)(k % (unsigned)F); }
        else { const long long k = g0 / N; e0 = (int)(g0 - k * N); kf0 = (int)(k % F); }
        // lane u < RB_ITEMS: the clip row of item g0 + u (past the end: a valid row that is never stored)
        int fu;
        {
            int e = e0 + (lane & (RB_ITEMS - 1)), kf = kf0;
            while (e >= N) { e -= N; kf = kf + 1 == F ? 0 : kf + 1; }
            fu = (int)(((unsigned)(frame0 ? __ldg(frame0 + e) : e) + (unsigned)kf) % (unsigned)F);
        }
#pragma unroll 1
        for (int c = 0; c < RB_ITEMS / RB_CHUNK; c++) {   // RB_CHUNK items' loads are in flight before their first shared-memory store
            float oa[RB_CHUNK], ob[RB_CHUNK], oc[RB_CHUNK], pa[RB_CHUNK], pb[RB_CHUNK], qa[RB_CHUNK], qb[RB_CHUNK];
#pragma unroll
            for (int u = 0; u < RB_CHUNK; u++) {
                const unsigned f = (unsigned)__shfl_sync(0xffffffffu, fu, c * RB_CHUNK + u);
                // three row pointers per item, every load at a constant offset from one of them
                const float* rp = data_pos + lane + (size_t)f * NQ;
                const float* rv = data_vel + lane + (size_t)f * NV;
                const float* tr = tab + lane + (size_t)f * RT_ROW;
                oa[u] = __ldg(rp);                                                            // obs columns 0..31
                ob[u] = __ldg(b_in_qpos ? rp + 32 : rv - (NQ - 32));                          // 32..63: qpos 32..34, qvel 0..28
                oc[u] = lane < PBRL_OBS - 64 ? __ldg(rv + (64 - NQ)) : 0.f;                   // 64..68: qvel 29..33
                pa[u] = __ldg(tr); pb[u] = lane < 42 - 32 ? __ldg(tr + 32) : 0.f;
                qa[u] = __ldg(tr + RT_QUAT); qb[u] = lane < 56 - 32 ? __ldg(tr + RT_QUAT + 32) : 0.f;
            }
            if (c == 0) {   // the engine has read the previous block's rows (waited for behind this block's first loads)
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncwarp();
            }
            float* so = s_obs + c * (RB_CHUNK * PBRL_OBS) + lane;
            float* sp = s_pos + c * (RB_CHUNK * 42) + lane;
            float* sq = s_quat + c * (RB_CHUNK * 56) + lane;
#pragma unroll
            for (int u = 0; u < RB_CHUNK; u++) {
                so[u * PBRL_OBS] = oa[u];
                so[u * PBRL_OBS + 32] = ob[u];
                if (lane < PBRL_OBS - 64) so[u * PBRL_OBS + 64] = oc[u];
                sp[u * 42] = pa[u];
                if (lane < 42 - 32) sp[u * 42 + 32] = pb[u];
                sq[u * 56] = qa[u];
                if (lane < 56 - 32) sq[u * 56 + 32] = qb[u];
            }
        }
        if (g0 + RB_ITEMS <= total) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the rows above, written through the generic proxy, before the engine reads them
            __syncwarp();
            if (lane == 0) {
                bulk_store(obs + g0 * PBRL_OBS, s_obs, RB_OBS * 4);
                bulk_store(xpos + g0 * 42, s_pos, RB_POS * 4);
                bulk_store(xquat + g0 * 56, s_quat, RB_QUAT * 4);
                bulk_store(reward + g0 * 3, s_const, RB_ITEMS * 3 * 4);
                bulk_store(done + g0, s_const + RB_ITEMS * 3, RB_ITEMS * 4);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        } else {   // the rollout's last, partial block
            __syncwarp();
            const int n = (int)(total - g0);
            for (int i = lane; i < n * PBRL_OBS; i += 32) __stcs(obs + g0 * PBRL_OBS + i, s_obs[i]);
            for (int i = lane; i < n * 42; i += 32) __stcs(xpos + g0 * 42 + i, s_pos[i]);
            for (int i = lane; i < n * 56; i += 32) __stcs(xquat + g0 * 56 + i, s_quat[i]);
            for (int i = lane; i < n * 3; i += 32) __stcs(reward + g0 * 3 + i, 0.f);
            for (int i = lane; i < n; i += 32) __stcs(done + g0 + i, 1.0f);
            __syncwarp();
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // shared memory stays allocated until the engine has read it
}
__global__ void __launch_bounds__(CTA_THREADS)
k_pd(const DevModel* __restrict__ gm, int N, int pd_mode, int action_cols, const float* action, const float* qpos,
     const float* qvel, const float* qpos_lag, const float* qvel_lag, const float* kp, const float* kd,
     const float* time, float pd_dt, float* tau, float* qM, float* bias) {
    Ws& w = my_ws();
    const int lane = threadIdx.x & 31;
    const DevModel& sm = stage_model(gm);
    init_ws_tables(sm, w, lane);
    for (int env = first_env(); env < N; env += env_stride())
        env_pd(sm, w, pd_mode, action_cols, action, qpos, qvel, qpos_lag, qvel_lag, kp, kd, time, pd_dt, tau, qM, bias, env, lane);
}
__global__ void __launch_bounds__(CTA_THREADS)
k_debug_forward(const DevModel* __restrict__ gm, int N, const float* qpos, const float* qvel, const float* ctrl,
                const float* warm, float* out) {
    Ws& w = my_ws();
    const int lane = threadIdx.x & 31;
    const DevModel& sm = stage_model(gm);
    init_ws_tables(sm, w, lane);
    for (int env = first_env(); env < N; env += env_stride()) env_debug_forward(sm, w, qpos, qvel, ctrl, warm, out, env, lane);
}

#if defined(PBRL_STAGE_PROF)
// dev microbenchmark: cycles per call of one stage function, `blockDim.x / 32` warps per CTA looping over resident code
__global__ void k_bench_stage(const DevModel* __restrict__ gm, int stage, int iters, unsigned long long* out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Ws& w = *reinterpret_cast<Ws*>(smem_raw + MODEL_SMEM + (size_t)(threadIdx.x >> 5) * sizeof(Ws));
    const int lane = threadIdx.x & 31;
    const DevModel& sm = stage_model(gm);
    init_ws_tables(sm, w, lane);
    LANE_LOOP(i, NQ) w.qpos[i] = sm.qpos0[i] + (i >= 7 ? 0.05f * ((i * 7) % 5 - 2) : 0.f);
    LANE_LOOP(i, NV) { w.qvel[i] = 0.1f * ((i * 3) % 7 - 3); w.warm[i] = 0.f; }
    LANE_LOOP(i, NU) w.ctrl[i] = 0.f;
    LANE_IF(z, 1) { w.status = 0; w.phase_lock = 0; w.qpos[2] = 0.9f; }
    __syncwarp();
    forward(sm, w, lane);  // a consistent workspace (contacts active: the feet touch the floor at z = 0.9)
    long long best = 1ll << 60, tot = 0;
    for (int it = 0; it < iters; it++) {
        if (stage == 0 || stage == 13) { LANE_LOOP(e, NM) w.LD[e] = w.M[e]; LANE_LOOP(i, NV) w.p[i] = 1.f + 0.01f * i; }
        __syncwarp();
        const long long t0 = clock64();
        switch (stage) {
            case 0: ldl_factor_solve(sm, w, w.p, lane); break;
            case 1: stage_fk(sm, w, w.qpos, lane); break;
            case 2: stage_com(sm, w, lane); break;
            case 3: stage_crb(sm, w, lane); break;
            case 4: stage_collide(sm, w, lane); break;
            case 5: stage_comvel(sm, w, w.qvel, lane); break;
            case 6: stage_bias(sm, w, lane); break;
            case 7: stage_rows(sm, w, w.qpos, w.qvel, lane); break;
            case 8: tree_Mx(sm, w, w.qacc, w.Mp, lane); break;
            case 9: break;
            case 10: rows_JTf(sm, w, w.Mp, w.grad, 0, lane); break;
            case 11: assemble_hessian(sm, w, lane); break;
            case 12: rows_Jx(sm, w, w.qacc, w.lim_jp, w.con_jp, w.sc_jp, lane); break;
            case 13: dense_factor_solve(sm, w, 0u, w.p, lane); break;
            case 14: stage_collide_self(sm, w, lane); break;
        }
        const long long t1 = clock64();
        if (stage == 1 || stage == 2) { __syncwarp(); }
        if (it > 0) { tot += t1 - t0; best = (t1 - t0) < best ? (t1 - t0) : best; }
    }
    if (lane == 0) { atomicAdd(out, (unsigned long long)tot); atomicMin(out + 1, (unsigned long long)best); atomicAdd(out + 2, (unsigned long long)(iters - 1)); }
}
#endif

}  // namespace

struct PbrlModel {
    int device = 0;
    int sm_count = 0;
    int ctas_per_sm = 1;
    int rr_ctas_per_sm = 1;
    int rb_ctas_per_sm = 0;   // k_replay_gather_bulk: resident CTAs per SM (0: the kernel does not fit, never used)
    bool axes_match = false;  // the model's hinge axes equal the compiled table topo::JNT_AXIS_KIND (replay kinematics specialisation)
    int phase_lock = 48;  // CTA barrier mask of k_step (pbrl_core.cuh: forward): the cooperative Newton loop's vote per iteration
                          // (16) and one barrier after the collision stage (32: a warp that ran the convex narrow phase rejoins
                          // the others before the velocity stages); PBRL_PHASE_LOCK overrides it at model creation (dev sweeps)
    int regroup_parts = 2;        // parts of a step regrouped inside one launch (equal shares of the substeps).  More parts: 16 384 envs
                                  // 3: +0.2 %, 4: -1.4 %; 4 096 envs 3: +1.8 %, 4: +2.3 %; 2 368 envs 4: -3 %; PBRL_REGROUP_PARTS
    int regroup_min_groups = 1;   // regroup only batches of at least that many groups of 16 envs per SM (4 096 envs, 1.7 per SM: +8 %;
                                  // 2 368, one per SM: +1 %); PBRL_REGROUP_MIN_GROUPS
    int regroup_fused = -1;  // 1: the two parts of a regrouped step in ONE launch (work-class queues); 0: two launches with a sort in
                             // between; -1: one launch below 16 groups per SM (where a launch boundary's drain costs more than the
                             // coarser grouping: +2.4 % at 16 384 envs), two at and above (-2 % at 65 536); PBRL_REGROUP_FUSED
    int regroup_every = 0;   // ... and again every that many substeps (0: once); PBRL_REGROUP_EVERY
    int regroup_after = -1;  // pbrl_step: 0 = never regroup the envs inside a step; otherwise the split point of the two-launch form (-1 = after
                             // half of the substeps, 4 of 8); the one-launch form splits the step into regroup_parts equal shares; PBRL_REGROUP_AFTER
    int tail_split = 1;   // k_step splits the groups of a batch smaller than the GPU over the idle SMs; PBRL_TAIL_SPLIT=0 turns it off (dev sweeps)
    int cost_mode = 0;    // which work estimate k_step reports (StepArgs::cost_mode); PBRL_COST_MODE overrides it (dev sweeps)
    DevModel host{};
    DevModel* dev = nullptr;
    DevMimic mimic{};
    bool has_mimic = false;
    float* frames = nullptr;
    // k_step's group counters: a ring of {next group, CTAs finished} pairs, one pair per launch in flight
    static constexpr int SCHED_RING = 32;
    int* sched = nullptr;
    mutable std::atomic<unsigned> sched_next{0};
    // pbrl_replay_rollout's frame tables: grow-only scratch, one per (clip pointer, row count) seen, so that rollouts of
    // different clips through one model -- also concurrent ones on different streams -- never share a table.  A rollout is
    // stream-ordered behind the previous one of the same clip on its stream.
    mutable std::mutex scratch_mu;
    struct FrameTab { const float* clip; int F; float* tab; };
    mutable std::vector<FrameTab> frame_tabs;
    float* frame_table(const float* clip, int F) const {
        std::lock_guard<std::mutex> lk(scratch_mu);
        for (auto& t : frame_tabs) if (t.clip == clip && t.F == F) return t.tab;
        float* p = nullptr;
        if (cudaMalloc(&p, (size_t)F * 100 * sizeof(float)) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        frame_tabs.push_back({clip, F, p});
        return p;
    }
    // block counters of k_replay_gather_bulk: a ring of slots, one per call in flight (zeroed in-stream by k_replay_frames)
    static constexpr int CLAIM_SLOTS = 8, CLAIM_SLOT_WORDS = RB_PARTS * RB_CLAIM_STRIDE;
    mutable unsigned* claim_buf = nullptr;
    mutable unsigned claim_next = 0;
    unsigned* claim_slot() const {
        std::lock_guard<std::mutex> lk(scratch_mu);
        if (!claim_buf && cudaMalloc(&claim_buf, CLAIM_SLOTS * CLAIM_SLOT_WORDS * sizeof(unsigned)) != cudaSuccess) { cudaGetLastError(); claim_buf = nullptr; return nullptr; }
        return claim_buf + (size_t)(claim_next++ % CLAIM_SLOTS) * CLAIM_SLOT_WORDS;
    }
    // pbrl_step's order for the second part of a regrouped step: a ring of grow-only buffers, one per call in flight
    static constexpr int ORDER_RING = 4;
    mutable int* order_bufs[ORDER_RING] = {nullptr, nullptr, nullptr, nullptr};
    mutable int order_cap[ORDER_RING] = {0, 0, 0, 0};
    mutable unsigned order_next = 0;
    int* regroup_order(int N) const {
        std::lock_guard<std::mutex> lk(scratch_mu);
        const int i = order_next++ % ORDER_RING;
        if (order_cap[i] < N) {
            // grow EVERY slot of the ring now: an allocation synchronises the device and costs milliseconds, so it must happen
            // on the first call (a caller's warm-up), never one slot at a time over the first four steps
            for (int k = 0; k < ORDER_RING; k++) {
                if (order_cap[k] >= N) continue;
                if (order_bufs[k]) cudaFree(order_bufs[k]);   // (cudaFree waits for the work that still reads it)
                order_bufs[k] = nullptr; order_cap[k] = 0;
                if (cudaMalloc(&order_bufs[k], (size_t)N * sizeof(int)) != cudaSuccess) { cudaGetLastError(); if (k == i) return nullptr; continue; }
                order_cap[k] = N;
            }
        }
        return order_bufs[i];
    }
};

namespace {
constexpr size_t SMEM_BYTES = MODEL_SMEM + sizeof(Ws) * WARPS_PER_CTA;

template <class K>
int prepare_kernel(K kernel, PbrlModel* m) {
    CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    int nb = 0;
    CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, CTA_THREADS, SMEM_BYTES));
    if (nb < 1) return fail("kernel does not fit on an SM");
    if (m->ctas_per_sm == 0 || nb < m->ctas_per_sm) m->ctas_per_sm = nb;
    return 0;
}
int grid_for(const PbrlModel* m, int N) {
    int need = (N + WARPS_PER_CTA - 1) / WARPS_PER_CTA, cap = m->sm_count * m->ctas_per_sm;
    return need < cap ? (need > 0 ? need : 1) : cap;
}
int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(std::string(what) + ": " + cudaGetErrorString(e));
    return 0;
}
}  // namespace

extern "C" {

const char* pbrl_last_error(void) {
    if (!g_error.empty()) return g_error.c_str();
    std::lock_guard<std::mutex> lk(g_error_mu);
    g_error_ret = g_error_any;
    return g_error_ret.c_str();
}
int pbrl_model_desc_size(void) { return (int)sizeof(PbrlModelDesc); }
int pbrl_model_desc_field(const char* name, int* offset, int* count, int* is_int) { return model_desc_field(name, offset, count, is_int); }
int pbrl_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int pbrl_model_create(const PbrlModelDesc* desc, const PbrlMimicDesc* mimic, int device, PbrlModel** out) {
    if (!desc || !out) return fail("null argument");
    *out = nullptr;
    int ndev = pbrl_device_count();
    if (ndev <= 0) return fail("no CUDA device: this library has no CPU path");
    if (device < 0 || device >= ndev) return fail("bad device index");
    DeviceGuard guard(device);
    if (!guard.ok) return fail("cannot select the CUDA device");
    cudaDeviceProp prop;
    CUDA_OK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail("libpbrl is built for sm_100a (B200) only");
    PbrlModel* m = new PbrlModel();
    // every failure below releases the handle and whatever it owns (pbrl_model_destroy frees dev / frames if set)
    auto bail = [&](const std::string& msg) { pbrl_model_destroy(m); return fail(msg); };
#define CREATE_OK(call)                                                                                      \
    do {                                                                                                     \
        cudaError_t e_ = (call);                                                                             \
        if (e_ != cudaSuccess) return bail(std::string(#call) + ": " + cudaGetErrorString(e_));              \
    } while (0)
    m->device = device;
    m->sm_count = prop.multiProcessorCount;
    m->ctas_per_sm = 0;
    { const char* e = getenv("PBRL_PHASE_LOCK"); if (e) m->phase_lock = atoi(e) & 63; }
    { const char* e = getenv("PBRL_TAIL_SPLIT"); if (e) m->tail_split = atoi(e) != 0; }
    { const char* e = getenv("PBRL_REGROUP_AFTER"); if (e) m->regroup_after = atoi(e); }
    { const char* e = getenv("PBRL_REGROUP_EVERY"); if (e) m->regroup_every = atoi(e); }
    { const char* e = getenv("PBRL_REGROUP_FUSED"); if (e) m->regroup_fused = atoi(e); }
    { const char* e = getenv("PBRL_REGROUP_PARTS"); if (e) m->regroup_parts = atoi(e); }
    { const char* e = getenv("PBRL_REGROUP_MIN_GROUPS"); if (e) m->regroup_min_groups = atoi(e) > 0 ? atoi(e) : 1; }
    { const char* e = getenv("PBRL_COST_MODE"); if (e) m->cost_mode = atoi(e); }
    std::string err = build_dev_model(*desc, m->host);
    if (!err.empty()) return bail(err);
    if (prepare_kernel(k_step, m) || prepare_kernel(k_reset, m) || prepare_kernel(k_replay, m) || prepare_kernel(k_pd, m) ||
        prepare_kernel(k_debug_forward, m) || prepare_kernel(k_autoreset, m)) { const std::string keep = g_error; return bail(keep); }
    {
        CREATE_OK(cudaFuncSetAttribute(k_replay_rollout<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RR_SMEM));
        CREATE_OK(cudaFuncSetAttribute(k_replay_rollout<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RR_SMEM));
        int nb = 0, nb2 = 0;
        CREATE_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_replay_rollout<true>, RR_ITEMS, RR_SMEM));
        CREATE_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb2, k_replay_rollout<false>, RR_ITEMS, RR_SMEM));
        if (nb < 1 || nb2 < 1) return bail("k_replay_rollout does not fit on an SM");
        m->rr_ctas_per_sm = nb < nb2 ? nb : nb2;
        CREATE_OK(cudaFuncSetAttribute(k_replay_gather_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RB_SMEM));
        CREATE_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&m->rb_ctas_per_sm, k_replay_gather_bulk, RB_THREADS, RB_SMEM));
        m->axes_match = true;
        for (int j = 0; j < NJ; j++) {
            const int kind = topo::JNT_AXIS_KIND[j];
            if (kind == 0) continue;
            for (int c = 0; c < 3; c++) {
                const float want = (c + 1 == (kind > 0 ? kind : -kind)) ? (kind > 0 ? 1.f : -1.f) : 0.f;
                if (m->host.jnt_axis[j][c] != want) m->axes_match = false;
            }
        }
    }
    CREATE_OK(cudaMalloc(&m->dev, sizeof(DevModel)));
    CREATE_OK(cudaMemcpy(m->dev, &m->host, sizeof(DevModel), cudaMemcpyHostToDevice));
    CREATE_OK(cudaMalloc(&m->sched, 2 * PbrlModel::SCHED_RING * sizeof(int)));
    CREATE_OK(cudaMemset(m->sched, 0, 2 * PbrlModel::SCHED_RING * sizeof(int)));
    if (mimic) {
        std::string merr = build_dev_mimic(*mimic, m->host, m->mimic);
        if (!merr.empty()) return bail(merr);
        size_t bytes = (size_t)mimic->nframes_total * PBRL_FRAME_STRIDE * sizeof(float);
        CREATE_OK(cudaMalloc(&m->frames, bytes));
        CREATE_OK(cudaMemcpy(m->frames, mimic->frames, bytes, cudaMemcpyHostToDevice));
        m->has_mimic = true;
    }
#undef CREATE_OK
    *out = m;
    return 0;
}

void pbrl_model_destroy(PbrlModel* m) {
    if (!m) return;
    DeviceGuard guard(m->device);
    if (m->dev) cudaFree(m->dev);
    if (m->frames) cudaFree(m->frames);
    if (m->sched) cudaFree(m->sched);
    for (auto& t : m->frame_tabs) cudaFree(t.tab);
    for (int* p : m->order_bufs) if (p) cudaFree(p);
    if (m->claim_buf) cudaFree(m->claim_buf);
    delete m;
}

static int check_buffers(const PbrlBuffers* b) {
    if (!b || !b->qpos || !b->qvel || !b->time || !b->warm || !b->ctrl || !b->qpos_lag || !b->qvel_lag)
        return fail("PbrlBuffers: state pointers must not be null");
    return 0;
}

int pbrl_reset(const PbrlModel* m, int N, int reward_mode, const PbrlBuffers* buf, void* stream) {
    if (!m) return fail("null model");
    ON_DEVICE(m);
    if (N <= 0) return 0;
    if (check_buffers(buf)) return -1;
    StepArgs a{};
    a.N = N; a.reward_mode = reward_mode; a.buf = *buf;
    if (reward_mode == PBRL_REWARD_MIMIC) {
        if (!m->has_mimic) return fail("the model was created without a mimic description");
        if (!buf->motion_id || !buf->ref_t0) return fail("mimic mode needs motion_id and ref_t0");
        a.frames = m->frames; a.mm = m->mimic;
    } else if (reward_mode != PBRL_REWARD_REFERENCE) return fail("unknown reward_mode");
    k_reset<<<grid_for(m, N), CTA_THREADS, SMEM_BYTES, (cudaStream_t)stream>>>(m->dev, a);
    return check_launch("pbrl_reset");
}

int pbrl_balance_order(const PbrlModel* m, int N, const int* cost, int* env_order, void* stream) {
    if (!m) return fail("null model");
    ON_DEVICE(m);
    if (N <= 0) return 0;
    if (!cost || !env_order) return fail("null argument");
    k_balance_order<<<1, BO_THREADS, 0, (cudaStream_t)stream>>>(N, cost, env_order);
    return check_launch("pbrl_balance_order");
}

int pbrl_autoreset(const PbrlModel* m, int N, int reward_mode, const float* done, const float* next_ref_t0, int* steps,
                   const PbrlBuffers* buf, void* stream) {
    if (!m) return fail("null model");
    ON_DEVICE(m);
    if (N <= 0) return 0;
    if (check_buffers(buf)) return -1;
    if (!done) return fail("null done");
    StepArgs a{};
    a.N = N; a.reward_mode = reward_mode; a.buf = *buf;
    if (reward_mode == PBRL_REWARD_MIMIC) {
        if (!m->has_mimic) return fail("the model was created without a mimic description");
        if (!buf->motion_id || !buf->ref_t0) return fail("mimic mode needs motion_id and ref_t0");
        a.frames = m->frames; a.mm = m->mimic;
    } else if (reward_mode != PBRL_REWARD_REFERENCE) return fail("unknown reward_mode");
    k_autoreset<<<grid_for(m, N), CTA_THREADS, SMEM_BYTES, (cudaStream_t)stream>>>(m->dev, a, done, next_ref_t0, steps);
    return check_launch("pbrl_autoreset");
}

// how many kernels pbrl_step launches for this batch (1, or the parts of a regrouped step plus the sorts between them)
int pbrl_step_launches(const PbrlModel* m, int N, int n_frames, const PbrlBuffers* buf) {
    if (!m || !buf || N <= 0) return 0;
    const int cap = m->sm_count * m->ctas_per_sm, r = m->regroup_after < 0 ? n_frames / 2 : m->regroup_after;
    if (!(r > 0 && r < n_frames && buf->env_order && buf->cost && buf->status && N >= m->regroup_min_groups * WARPS_PER_CTA * cap)) return 1;
    if (m->regroup_fused < 0 ? N < 16 * WARPS_PER_CTA * cap : m->regroup_fused != 0) return 1;
    const int every = m->regroup_every > 0 ? m->regroup_every : n_frames;
    int parts = 1;
    for (int f = r; f < n_frames; f += every) parts++;
    return 2 * parts - 1;
}

int pbrl_step(const PbrlModel* m, int N, int n_frames, int pd_mode, int action_cols, const float* action, const float* kp,
              const float* kd, float pd_dt, const float* pd_time, unsigned ctrl_mask, int reward_mode, const PbrlBuffers* buf,
              void* stream) {
    if (!m) return fail("null model");
    ON_DEVICE(m);
    if (N <= 0) return 0;
    if (check_buffers(buf)) return -1;
    if (!action) return fail("null action");
    if (pd_mode < 0 || pd_mode > 3) return fail("pd_mode must be 0..3");
    if (action_cols != 1 && action_cols != 4) return fail("action_cols must be 1 (targets/torques) or 4 (cubic coefficients)");
    if (pd_mode != PBRL_PD_TORQUE && (!kp || !kd)) return fail("kp/kd required for the PD modes");
    if (n_frames < 1) return fail("n_frames must be >= 1");
    StepArgs a{};
    if (reward_mode == PBRL_REWARD_MIMIC) {
        if (!m->has_mimic) return fail("the model was created without a mimic description");
        if (!buf->motion_id || !buf->ref_t0) return fail("mimic mode needs motion_id and ref_t0");
        a.frames = m->frames; a.mm = m->mimic;
    } else if (reward_mode != PBRL_REWARD_REFERENCE) return fail("unknown reward_mode");
    a.N = N; a.n_frames = n_frames; a.pd_mode = pd_mode; a.action_cols = action_cols; a.reward_mode = reward_mode;
    a.action = action; a.kp = kp; a.kd = kd; a.pd_time = pd_time; a.pd_dt = pd_dt; a.ctrl_mask = ctrl_mask; a.buf = *buf;
    a.phase_lock = m->phase_lock;
    a.cost_mode = m->cost_mode;
    a.tail_split = m->tail_split;
    // small batches: more CTAs than groups, so that the groups can be split over them (down to one env per CTA)
    const int cap = m->sm_count * m->ctas_per_sm, grid = m->tail_split ? (N < cap ? N : cap) : grid_for(m, N);
    auto launch = [&](int f_begin, int f_end) {
        a.f_begin = f_begin; a.f_end = f_end;
        int* sched = m->sched + 2 * (m->sched_next.fetch_add(1) % PbrlModel::SCHED_RING);
        k_step<<<grid, CTA_THREADS, SMEM_BYTES, (cudaStream_t)stream>>>(m->dev, a, sched);
    };
    // Regrouping inside the step: how much solver work an env needs is far better predicted by its previous SUBSTEP (94 % of
    // the dense-Hessian envs stay dense from one substep to the next) than by its previous env step (a quarter of them were not
    // dense then).  With `regroup_after` = r > 0 the step is two launches: PD law + r substeps in the caller's processing order,
    // then -- the envs re-sorted by the work of those substeps (into a scratch order of the model) -- the other substeps and
    // the epilogue.  The state crosses the boundary in the caller's buffers.  Needs env_order, cost and status.
    const int r = m->regroup_after < 0 ? n_frames / 2 : m->regroup_after;
    const bool regroup = r > 0 && r < n_frames && buf->env_order && buf->cost && buf->status && N >= m->regroup_min_groups * WARPS_PER_CTA * cap;
    const bool fused = m->regroup_fused < 0 ? N < 16 * WARPS_PER_CTA * cap : m->regroup_fused != 0;
    if (regroup && fused) {
        // the parts in one launch: scheduler block (256 ints) + per later part one queue of N entries per work class, zeroed per launch
        int P = m->regroup_parts > 1 ? m->regroup_parts : 2;
        if (P > n_frames) P = n_frames;
        if (P > 7) P = 7;
        const size_t ints = 256 + (size_t)(P - 1) * FUSED_CLASSES * N;
        if (ints > 0x7fffffffull) return fail("pbrl_step: batch too large for the regrouping queues");
        int* blk = m->regroup_order((int)ints);
        if (!blk) return fail("pbrl_step: cannot allocate the regrouping queues");
        CUDA_OK(cudaMemsetAsync(blk, 0, ints * sizeof(int), (cudaStream_t)stream));
        a.f_begin = 0; a.f_end = n_frames; a.fused_parts = P; a.fused_queue = blk + 256;
        k_step<<<grid, CTA_THREADS, SMEM_BYTES, (cudaStream_t)stream>>>(m->dev, a, blk);
        return check_launch("pbrl_step");
    }
    int* order2 = nullptr;
    if (regroup) order2 = m->regroup_order(N);
    if (order2) {
        const int every = m->regroup_every > 0 ? m->regroup_every : n_frames;
        int f0 = 0;
        for (int f1 = r; f0 < n_frames; f0 = f1, f1 = (f1 + every < n_frames ? f1 + every : n_frames)) {
            if (f0 > 0) {
                k_balance_order<<<1, BO_THREADS, 0, (cudaStream_t)stream>>>(N, buf->cost, order2);
                a.buf.env_order = order2;
            }
            launch(f0, f1);
        }
    } else {
        launch(0, n_frames);
    }
    return check_launch("pbrl_step");
}

int pbrl_replay(const PbrlModel* m, int N, const float* qpos, const float* qvel, float* obs, float* reward, float* done,
                float* xpos, float* xquat, void* stream) {
    if (!m) return fail("null model");
    ON_DEVICE(m);
    if (N <= 0) return 0;
    if (!qpos || !qvel) return fail("null qpos/qvel");
    k_replay<<<grid_for(m, N), CTA_THREADS, SMEM_BYTES, (cudaStream_t)stream>>>(m->dev, N, qpos, qvel, obs, reward, done, xpos, xquat);
    return check_launch("pbrl_replay");
}

int pbrl_replay_rollout(const PbrlModel* m, int N, int K, const float* data_pos, const float* data_vel, int F, const int* frame0,
                        float* obs, float* reward, float* done, float* xpos, float* xquat, void* stream) {
    if (!m) return fail("null model");
    ON_DEVICE(m);
    if (N <= 0 || K <= 0) return 0;
    if (!data_pos || !data_vel || F <= 0) return fail("null frame data");
    ReplayTables T;
    replay_tables_from(m->host, T);
    cudaStream_t st = (cudaStream_t)stream;
    // frame table when the items outnumber the clip's rows (and poses are asked for): see k_replay_gather
    const char* per_item_env = getenv("PBRL_REPLAY_PER_ITEM");   // (read per call: bench.py times both forms in one process)
    const bool per_item_only = per_item_env && atoi(per_item_env) != 0;
    if (!per_item_only && (long long)N * K >= 4ll * F && F < (1 << 24)) {   // (row offsets of the table are 32-bit)
        float* tab = nullptr;
        if (xpos || xquat) {
            tab = m->frame_table(data_pos, F);
            if (!tab) return fail("pbrl_replay_rollout: cannot allocate the frame table");
        }
        const long long need = ((long long)N * K + RG_UNROLL * (RG_THREADS / 32) - 1) / (RG_UNROLL * (RG_THREADS / 32));
        const long long cap2 = (long long)m->sm_count * 8;
        const unsigned grid2 = (unsigned)(need < cap2 ? need : cap2);
        // every output present and 16-byte aligned: the stores go through the bulk-copy engine (PBRL_REPLAY_BULK=0: ordinary stores)
        const char* bulk_env = getenv("PBRL_REPLAY_BULK");
        const bool bulk = (bulk_env ? atoi(bulk_env) != 0 : PBRL_REPLAY_BULK_DEFAULT) && m->rb_ctas_per_sm > 0 && obs && reward && done && xpos && xquat &&
                          (((uintptr_t)obs | (uintptr_t)reward | (uintptr_t)done | (uintptr_t)xpos | (uintptr_t)xquat) & 15) == 0;
        unsigned* claim = bulk ? m->claim_slot() : nullptr;
        if (bulk && !claim) return fail("pbrl_replay_rollout: cannot allocate the block counters");
        if (tab) {
            if (m->axes_match) k_replay_frames<true><<<(F + 63) / 64, 64, 0, st>>>(T, F, data_pos, tab, claim);
            else k_replay_frames<false><<<(F + 63) / 64, 64, 0, st>>>(T, F, data_pos, tab, claim);
        }
        if (bulk) {
            const long long needb = ((long long)N * K + RB_ITEMS * RB_WARPS - 1) / (RB_ITEMS * RB_WARPS);
            const long long capb = (long long)m->sm_count * m->rb_ctas_per_sm;
            // programmatic dependent launch: the CTAs take their places while k_replay_frames (a handful of CTAs) is still running
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)(needb < capb ? needb : capb));
            cfg.blockDim = dim3(RB_THREADS);
            cfg.dynamicSmemBytes = RB_SMEM;
            cfg.stream = st;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            const cudaError_t le = cudaLaunchKernelEx(&cfg, k_replay_gather_bulk, N, K, data_pos, data_vel, F, frame0, (const float*)tab, claim, obs,
                                                      reward, done, xpos, xquat);
            if (le != cudaSuccess) { cudaGetLastError(); return fail(std::string("pbrl_replay_rollout: ") + cudaGetErrorString(le)); }
            return check_launch("pbrl_replay_rollout");
        }
        if (obs && reward && done && xpos && xquat)
            k_replay_gather<true><<<grid2, RG_THREADS, 0, st>>>(N, K, data_pos, data_vel, F, frame0, tab, obs, reward, done, xpos, xquat);
        else
            k_replay_gather<false><<<grid2, RG_THREADS, 0, st>>>(N, K, data_pos, data_vel, F, frame0, tab, obs, reward, done, xpos, xquat);
        return check_launch("pbrl_replay_rollout");
    }
    const long long ctas = ((long long)N * K + RR_ITEMS - 1) / RR_ITEMS;
    const long long cap = (long long)m->sm_count * m->rr_ctas_per_sm;
    const unsigned grid = (unsigned)(ctas < cap ? ctas : cap);
    if (m->axes_match) k_replay_rollout<true><<<grid, RR_ITEMS, RR_SMEM, st>>>(T, N, K, data_pos, data_vel, F, frame0, obs, reward, done, xpos, xquat);
    else k_replay_rollout<false><<<grid, RR_ITEMS, RR_SMEM, st>>>(T, N, K, data_pos, data_vel, F, frame0, obs, reward, done, xpos, xquat);
    return check_launch("pbrl_replay_rollout");
}

int pbrl_pd_torque(const PbrlModel* m, int N, int pd_mode, int action_cols, const float* action, const float* qpos,
                   const float* qvel, const float* qpos_lag, const float* qvel_lag, const float* kp, const float* kd,
                   const float* time, float pd_dt, float* tau, float* qM, float* bias, void* stream) {
    if (!m) return fail("null model");
    ON_DEVICE(m);
    if (N <= 0) return 0;
    if (!action || !qpos || !qvel || !qpos_lag || !qvel_lag || !kp || !kd || !time || !tau) return fail("null argument");
    if (pd_mode < 0 || pd_mode > 3) return fail("pd_mode must be 0..3");
    if (action_cols != 1 && action_cols != 4) return fail("action_cols must be 1 or 4");
    k_pd<<<grid_for(m, N), CTA_THREADS, SMEM_BYTES, (cudaStream_t)stream>>>(m->dev, N, pd_mode, action_cols, action, qpos, qvel, qpos_lag,
                                                                          qvel_lag, kp, kd, time, pd_dt, tau, qM, bias);
    return check_launch("pbrl_pd_torque");
}

int pbrl_debug_forward(const PbrlModel* m, int N, const float* qpos, const float* qvel, const float* ctrl, const float* warm,
                       float* out, void* stream) {
    if (!m) return fail("null model");
    ON_DEVICE(m);
    if (N <= 0) return 0;
    if (!qpos || !qvel || !ctrl || !warm || !out) return fail("null argument");
    k_debug_forward<<<grid_for(m, N), CTA_THREADS, SMEM_BYTES, (cudaStream_t)stream>>>(m->dev, N, qpos, qvel, ctrl, warm, out);
    return check_launch("pbrl_debug_forward");
}

// dev builds (-DPBRL_STAGE_PROF) accumulate per-stage cycle counters; the product build reports "not a profiling build"
int pbrl_debug_stage_profile(unsigned long long* cycles, unsigned long long* hits, int reset) {
#if defined(PBRL_STAGE_PROF)
    CUDA_OK(cudaDeviceSynchronize());
    if (cycles) CUDA_OK(cudaMemcpyFromSymbol(cycles, g_stage_cycles, sizeof(unsigned long long) * 64));
    if (hits) CUDA_OK(cudaMemcpyFromSymbol(hits, g_stage_hits, sizeof(unsigned long long) * 64));
    if (reset) {
        unsigned long long z[64] = {0};
        CUDA_OK(cudaMemcpyToSymbol(g_stage_cycles, z, sizeof(z)));
        CUDA_OK(cudaMemcpyToSymbol(g_stage_hits, z, sizeof(z)));
    }
    return 0;
#else
    (void)cycles; (void)hits; (void)reset;
    return fail("not a profiling build (compile with -DPBRL_STAGE_PROF)");
#endif
}

#if defined(PBRL_STAGE_PROF)
int pbrl_debug_bench_stage(const PbrlModel* m, int stage, int warps_per_cta, int ctas, int iters, unsigned long long* out3) {
    unsigned long long* d = nullptr;
    unsigned long long init[3] = {0ull, ~0ull, 0ull};
    CUDA_OK(cudaMalloc(&d, sizeof(init)));
    CUDA_OK(cudaMemcpy(d, init, sizeof(init), cudaMemcpyHostToDevice));
    const size_t smem = MODEL_SMEM + sizeof(Ws) * warps_per_cta;
    CUDA_OK(cudaFuncSetAttribute(k_bench_stage, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_bench_stage<<<ctas, 32 * warps_per_cta, smem>>>(m->dev, stage, iters, d);
    CUDA_OK(cudaDeviceSynchronize());
    CUDA_OK(cudaMemcpy(out3, d, sizeof(init), cudaMemcpyDeviceToHost));
    cudaFree(d);
    return 0;
}
#endif

// ------------------------------------------------------------------------------------------ host-buffer layer
struct PbrlBatch {
    const PbrlModel* model;
    int N;
    cudaStream_t stream;
    PbrlBuffers buf;   // device
    float* action;     // device, [N][28][4]
    float* gains;      // device, kp[28] kd[28]
    int* motion_id;
    float* ref_t0;
    int* order;        // device, processing order of the next step (pbrl_balance_order)
    std::vector<void*> allocs;
    cudaEvent_t results_ready = nullptr;  // recorded after the last device->host copy of a step
    float gains_host[2 * NU] = {0};       // what b->gains holds (the gains are re-uploaded only when they change)
    bool gains_valid = false;
};

int pbrl_batch_create(const PbrlModel* m, int N, PbrlBatch** out) {
    if (!m || !out || N <= 0) return fail("bad argument");
    ON_DEVICE(m);
    PbrlBatch* b = new PbrlBatch();
    b->model = m; b->N = N;
    if (cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking) != cudaSuccess) { delete b; return fail("cudaStreamCreate failed"); }
    if (cudaEventCreateWithFlags(&b->results_ready, cudaEventDisableTiming) != cudaSuccess) { cudaStreamDestroy(b->stream); delete b; return fail("cudaEventCreate failed"); }
    bool failed = false;
    auto alloc = [&](size_t n) -> float* {
        void* p = nullptr;
        if (cudaMalloc(&p, n * sizeof(float)) != cudaSuccess) { cudaGetLastError(); failed = true; return nullptr; }
        cudaMemsetAsync(p, 0, n * sizeof(float), b->stream);
        b->allocs.push_back(p);
        return (float*)p;
    };
    PbrlBuffers& B = b->buf;
    B.qpos = alloc((size_t)N * NQ); B.qvel = alloc((size_t)N * NV); B.time = alloc(N); B.warm = alloc((size_t)N * NV);
    B.ctrl = alloc((size_t)N * NU); B.qpos_lag = alloc((size_t)N * NQ); B.qvel_lag = alloc((size_t)N * NV);
    B.obs = alloc((size_t)N * PBRL_OBS); B.reward = alloc((size_t)N * 3); B.done = alloc(N); B.metrics = alloc((size_t)N * 3);
    B.status = (int*)alloc(N); B.stats = alloc(8);
    B.cost = (int*)alloc(N); b->order = (int*)alloc(N);
    b->motion_id = (int*)alloc(N); b->ref_t0 = alloc(N);
    B.motion_id = nullptr; B.ref_t0 = nullptr;
    b->action = alloc((size_t)N * NU * 4); b->gains = alloc(2 * NU);
    if (failed) { pbrl_batch_destroy(b); return fail("cudaMalloc failed"); }
    {
        std::vector<int> iota(N);
        for (int i = 0; i < N; i++) iota[i] = i;
        if (cudaMemcpyAsync(b->order, iota.data(), N * sizeof(int), cudaMemcpyHostToDevice, b->stream) != cudaSuccess ||
            cudaStreamSynchronize(b->stream) != cudaSuccess) { pbrl_batch_destroy(b); return fail("pbrl_batch_create: upload failed"); }
        B.env_order = b->order;
    }
    *out = b;
    return 0;
}
void pbrl_batch_destroy(PbrlBatch* b) {
    if (!b) return;
    DeviceGuard guard(b->model->device);
    cudaStreamSynchronize(b->stream);
    for (void* p : b->allocs) if (p) cudaFree(p);
    if (b->results_ready) cudaEventDestroy(b->results_ready);
    cudaStreamDestroy(b->stream);
    delete b;
}
int pbrl_batch_reset_host(PbrlBatch* b, int reward_mode, const int* motion_id, const float* ref_t0, float* obs_out) {
    if (!b) return fail("null batch");
    ON_DEVICE(b->model);
    if (motion_id) { CUDA_OK(cudaMemcpyAsync(b->motion_id, motion_id, b->N * sizeof(int), cudaMemcpyHostToDevice, b->stream)); b->buf.motion_id = b->motion_id; }
    if (ref_t0) { CUDA_OK(cudaMemcpyAsync(b->ref_t0, ref_t0, b->N * sizeof(float), cudaMemcpyHostToDevice, b->stream)); b->buf.ref_t0 = b->ref_t0; }
    if (pbrl_reset(b->model, b->N, reward_mode, &b->buf, b->stream)) return -1;
    if (obs_out) CUDA_OK(cudaMemcpyAsync(obs_out, b->buf.obs, (size_t)b->N * PBRL_OBS * sizeof(float), cudaMemcpyDeviceToHost, b->stream));
    CUDA_OK(cudaStreamSynchronize(b->stream));
    return 0;
}
// Host buffers in, host buffers out.  Output rows the kernel only writes (obs, reward, done) go STRAIGHT to the caller's buffers
// when those are pinned, device-mapped host memory (cudaHostAlloc / cudaHostRegister; torch's pin_memory): k_step's row stores
// then travel over PCIe while the other envs are still being stepped, instead of 4.8 MB of copies after the kernel.  The action
// rows are read in place the same way (one 112-byte row per env step).  Pageable buffers take the staged copies.
static void* device_view_of_host(const void* p) {
    static const bool enabled = [] { const char* e = getenv("PBRL_HOST_ZERO_COPY"); return !(e && atoi(e) == 0); }();
    if (!enabled || !p) return nullptr;
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return at.type == cudaMemoryTypeHost ? at.devicePointer : nullptr;
}
int pbrl_batch_step_host(PbrlBatch* b, int n_frames, int pd_mode, int action_cols, const float* action, const float* kp,
                         const float* kd, float pd_dt, int reward_mode, float* obs_out, float* reward_out, float* done_out,
                         float* stats_out) {
    if (!b || !action) return fail("null argument");
    if (action_cols != 1 && action_cols != 4) return fail("action_cols must be 1 or 4");
    ON_DEVICE(b->model);
    const int N = b->N;
    const float* action_dev = static_cast<const float*>(device_view_of_host(action));
    if (!action_dev) {
        CUDA_OK(cudaMemcpyAsync(b->action, action, (size_t)N * NU * action_cols * sizeof(float), cudaMemcpyHostToDevice, b->stream));
        action_dev = b->action;
    }
    if (kp && kd && !(b->gains_valid && !memcmp(b->gains_host, kp, NU * sizeof(float)) && !memcmp(b->gains_host + NU, kd, NU * sizeof(float)))) {
        memcpy(b->gains_host, kp, NU * sizeof(float));
        memcpy(b->gains_host + NU, kd, NU * sizeof(float));
        CUDA_OK(cudaMemcpyAsync(b->gains, b->gains_host, 2 * NU * sizeof(float), cudaMemcpyHostToDevice, b->stream));
        b->gains_valid = true;
    }
    if (stats_out) CUDA_OK(cudaMemsetAsync(b->buf.stats, 0, 8 * sizeof(float), b->stream));
    PbrlBuffers buf = b->buf;
    float* obs_dev = static_cast<float*>(device_view_of_host(obs_out));
    float* reward_dev = static_cast<float*>(device_view_of_host(reward_out));
    float* done_dev = static_cast<float*>(device_view_of_host(done_out));
    if (obs_dev) buf.obs = obs_dev;
    if (reward_dev) buf.reward = reward_dev;
    if (done_dev) buf.done = done_dev;
    if (pbrl_step(b->model, N, n_frames, pd_mode, action_cols, action_dev, b->gains, b->gains + NU, pd_dt, nullptr, 0x0FFFFFFFu,
                  reward_mode, &buf, b->stream)) return -1;
    if (obs_out && !obs_dev) CUDA_OK(cudaMemcpyAsync(obs_out, b->buf.obs, (size_t)N * PBRL_OBS * sizeof(float), cudaMemcpyDeviceToHost, b->stream));
    if (reward_out && !reward_dev) CUDA_OK(cudaMemcpyAsync(reward_out, b->buf.reward, (size_t)N * 3 * sizeof(float), cudaMemcpyDeviceToHost, b->stream));
    if (done_out && !done_dev) CUDA_OK(cudaMemcpyAsync(done_out, b->buf.done, (size_t)N * sizeof(float), cudaMemcpyDeviceToHost, b->stream));
    if (stats_out) CUDA_OK(cudaMemcpyAsync(stats_out, b->buf.stats, 8 * sizeof(float), cudaMemcpyDeviceToHost, b->stream));
    CUDA_OK(cudaEventRecord(b->results_ready, b->stream));
    // the next step's processing order is computed behind the results: the caller does not wait for it (the next call's
    // launch is ordered after it on the stream)
    if (pbrl_balance_order(b->model, N, b->buf.cost, b->order, b->stream)) return -1;
    CUDA_OK(cudaEventSynchronize(b->results_ready));
    return 0;
}
int pbrl_batch_get_state_host(PbrlBatch* b, float* qpos, float* qvel, float* time) {
    if (!b) return fail("null batch");
    ON_DEVICE(b->model);
    if (qpos) CUDA_OK(cudaMemcpyAsync(qpos, b->buf.qpos, (size_t)b->N * NQ * sizeof(float), cudaMemcpyDeviceToHost, b->stream));
    if (qvel) CUDA_OK(cudaMemcpyAsync(qvel, b->buf.qvel, (size_t)b->N * NV * sizeof(float), cudaMemcpyDeviceToHost, b->stream));
    if (time) CUDA_OK(cudaMemcpyAsync(time, b->buf.time, (size_t)b->N * sizeof(float), cudaMemcpyDeviceToHost, b->stream));
    CUDA_OK(cudaStreamSynchronize(b->stream));
    return 0;
}

}  // extern "C"
