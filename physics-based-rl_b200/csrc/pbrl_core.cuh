// One-warp-per-environment physics core for the humanoid-mimic env step (sm_100a).
//
// Replaces what the reference runs through brax/MJX for HumanoidTrain.step
// (/root/reference/agents_env/agent_mimic.py:58-91 -> PipelineEnv.pipeline_step -> mjx.step) with a
// single fused kernel body: stable-PD prologue, n_frames substeps of
//   FK -> com/cinert/cdof -> CRB -> tree-sparse LDL -> com_vel + RNE -> actuation -> plane contacts +
//   joint limits -> primal Newton on qacc -> implicit-damping Euler -> integration,
// and the obs/reward epilogue.  Algorithm spec: SURVEY.md App. A (upstream MuJoCo/MJX semantics).
//
// Structure chosen for the GPU (differs on purpose from the dense textbook form of the oracle):
//   * one warp owns one env; all per-env data lives in that warp's shared-memory workspace `Ws`;
//   * every stage is a lane-parallel loop over bodies / dofs / matrix entries / contact slots followed
//     by a warp barrier; tree recursions run level by level, leaf-to-root sums as subtree gathers;
//   * M, M + h*D, M + KD*dt and the Newton Hessian M + J^T D J all share the branch-induced sparsity
//     of the kinematic tree (310 non-zeros) and are factorised as L^T D L without fill-in;
//   * contact Jacobians are never formed: J x and J^T f run through the tree (spatial velocity down,
//     spatial force up), and J^T D J is assembled like a composite-rigid-body inertia.
//
// The same source compiles for the host when PBRL_HOST_EMU is defined (lane loops become sequential
// loops).  That build exists only for tests/ (logic checks without a GPU); the product library is the
// CUDA build and has no CPU path.
#pragma once
#include <math.h>
#include <stddef.h>
#include <string.h>

#include "pbrl_model.h"
#include "pbrl_topo_gen.h"

#if defined(__CUDACC__) && !defined(PBRL_HOST_EMU)
#define PB_DEV __device__ __forceinline__
#define PB_DEVN __device__ __noinline__
#define PB_MEMBER static __device__ __forceinline__
#define LANE_LOOP(i, n) for (int i = lane; i < (n); i += 32)
#define LANE_IF(i, n) for (int i = lane, once_ = 1; once_ && i < (n); once_ = 0)  /* n <= 32: at most one pass */
#define WSYNC() __syncwarp()
// phase-lock barrier: the warps of one lock group (PBRL_LOCK_GROUP consecutive warps of the CTA) meet on a named barrier
#ifndef PBRL_LOCK_GROUP
#define PBRL_LOCK_GROUP 16
#endif
#define CTA_SYNC() asm volatile("bar.sync %0, %1;" ::"r"(1 + (int)(threadIdx.x >> 5) / PBRL_LOCK_GROUP), "r"(32 * PBRL_LOCK_GROUP) : "memory")
// Newton vote: OR of a predicate over the warps that iterate in step (PBRL_VOTE_GROUP consecutive warps; default: the CTA)
#ifndef PBRL_VOTE_GROUP
#define PBRL_VOTE_GROUP 16
#endif
__device__ __forceinline__ int vote_group_or(int p) {
    int r;
    asm volatile("{\n\t.reg .pred q, r;\n\tsetp.ne.b32 q, %1, 0;\n\tbarrier.cta.red.or.pred r, %2, %3, q;\n\tselp.b32 %0, 1, 0, r;\n\t}"
                 : "=r"(r) : "r"(p), "r"(8 + (int)(threadIdx.x >> 5) / PBRL_VOTE_GROUP), "r"(32 * PBRL_VOTE_GROUP) : "memory");
    return r;
}
#define CTA_OR(p) vote_group_or(p)
#else
#define PB_DEV static inline
#define PB_DEVN static
#define PB_MEMBER static inline
#define LANE_LOOP(i, n) for (int i = 0; i < (n); i++)
#define LANE_IF(i, n) for (int i = 0; i < (n); i++)
#define WSYNC() ((void)0)
#define CTA_SYNC() ((void)0)
#define CTA_OR(p) (p)
#endif

// Stage profiler (dev builds only, -DPBRL_STAGE_PROF): lane 0 of every warp adds the cycles since its previous mark
// to a global per-stage counter.  Read back through pbrl_debug_stage_profile().
#if defined(PBRL_STAGE_PROF) && defined(__CUDACC__) && !defined(PBRL_HOST_EMU)
__device__ unsigned long long g_stage_cycles[64];
__device__ unsigned long long g_stage_hits[64];
#define PROF(id)                                                                      \
    do {                                                                              \
        if (lane == 0) {                                                              \
            const long long t_ = clock64();                                           \
            atomicAdd(&g_stage_cycles[id], (unsigned long long)(t_ - w.prof_t));      \
            atomicAdd(&g_stage_hits[id], 1ull);                                       \
            w.prof_t = t_;                                                            \
        }                                                                             \
    } while (0)
#define PROF_START() do { if (lane == 0) w.prof_t = clock64(); } while (0)
#define PROF_COUNT(id) do { if (lane == 0) atomicAdd(&g_stage_hits[id], 1ull); } while (0)
#else
#define PROF_COUNT(id) ((void)0)
#define PROF(id) ((void)0)
#define PROF_START() ((void)0)
#endif

#ifndef PBRL_DENSE_WEIGHT
#define PBRL_DENSE_WEIGHT 3  // a Newton iteration with the dense Hessian, in units of a tree-sparse one (work estimate only)
#endif

namespace pbrl {
#if defined(PBRL_HOST_EMU) && defined(PBRL_EMU_TRACE)
static long g_emu_cnt[16] = {0};
#endif

// row strides of the per-body / per-dof arrays.  6-vectors are packed (24-byte rows moved as three 8-byte accesses: with
// consecutive rows in consecutive lanes a 64-bit access touches every bank pair once per half warp, where 8-float rows
// moved as 16 + 8 bytes replayed 2x / 4x on bank conflicts); spatial inertias live in 12-float rows (48 bytes: conflict-free
// for 128-bit accesses) and symmetric 6x6 blocks in 24-float rows.
constexpr int S3 = 3, S4 = 4, S6 = 6, S9 = 9, S10 = 12, S21 = 24;
constexpr int NBUF6 = 256;  // floats of stage scratch: 32 x 8 pair results of the narrow phase is the largest user
constexpr int SR = 4;        // contact points: float4 rows
constexpr int NDL = NV - 3;  // "dof lanes": lane l owns dof l + 3; the lane of dof 3 also emits the free translations 0..2
constexpr int NEWTON_MAXIT = 12;  // iteration cap of the exact-active-set rule (it needs <= 6 in practice; guards against fp32 cycling)
constexpr int LS_MAXIT = 24;      // line-search cap of the exact-active-set rule

struct Ws {
    // state
    float qpos[36], qvel[36], warm[36], ctrl[28];
    // kinematics.  xquat/xmat are dead once the contacts are found; the spatial-force scratch F reuses them
    float xpos[NB * S3], com[4], pad0_[2];
    union {
        struct { alignas(16) float xquat[NB * S4]; float xmat[NB * S9]; };
        struct { alignas(16) float F[NB * S6]; };
    };
    // spatial quantities about the centre of mass, [ang; lin]
    alignas(16) float cinert[NB * S10];
    alignas(16) float cdof[NV * S6];
    alignas(16) float cvel[NB * S6];
    alignas(16) float buf6[NBUF6];    // stage scratch; also the 14 x 16 staging rows of the factorisation
    union {
        alignas(16) float cacc[NB * S6];  // bias accelerations (com_vel -> rne)
        alignas(16) float V[NB * S6];     // body velocities of a generalised vector (solver, feedback PD)
    };
    // joint-space inertia (tree-sparse), factor, vectors
    float M[NM], Dinv[36], pad1_[2];
    union {
        alignas(16) float LD[NM];                     // factor workspace
        alignas(16) float fr[2 * PBRL_FRAME_STRIDE];  // mimic: the two reference frames of this env step (cp.async.bulk
    };                                                // target; issued after the last factorisation of the step)
    union {  // directly behind LD: the dense Hessian of the body-vs-body path (NHD floats from LD on) covers both
        struct { alignas(16) float K[NB * S21]; };                         // Newton Hessian assembly: 6x6 blocks per body
        struct {                                                          // mimic epilogue (after the last substep)
            float rq[9 * 4], rvel[36], rpos[4], rhinge[NB], rxpos[NB * S3], rxquat[NB * 5], rcom[4], ree[12], rerr[8];
        };
    };
    float fsmooth[36], qacc[36], grad[36], p[36];
    union { float Ma[36]; float tau[28]; };      // tau: PD law output, consumed before the solver runs
    union { float Mp[36]; float asmooth[36]; };  // asmooth: only formed when no row is active
    union { float fcon[36]; float bias[36]; };   // bias is folded into fsmooth before the solver writes fcon
    // PD prologue only: the state of the last forward (qpos_lag / qvel_lag rows) is staged in fsmooth / qacc
    // constraint rows: lane l owns limit row l and contact slot l (4 pyramid rows)
    // (aref is consumed when the solver starts: jar = J a - aref; J p of a search direction is formed later -> shared storage)
    float lim_sign[32], lim_D[32], lim_jar[32];
    union { float lim_aref[32]; float lim_jp[32]; };
    int con_on[NCP], act_bits[32];
    float con_dist[NCP], con_D[NCP], con_t1[NCP * 2];
    alignas(16) float con_r[NCP * 4];
    union { alignas(16) float con_aref[NCP * 4]; alignas(16) float con_jp[NCP * 4]; };
    alignas(16) float con_jar[NCP * 4];
    // body-vs-body contacts inside the margin (at most KS): lane l owns pyramid row l & 3 of contact l >> 2
    alignas(16) float sc_r[KS * 4];     // contact point - com, [3] = D
    alignas(16) float sc_n[KS * 4];     // normal (geom1 -> geom2), [3] = dist
    union { alignas(16) float sc_aref[KS * 4]; alignas(16) float sc_jp[KS * 4]; };
    alignas(16) float sc_jar[KS * 4];
    int sc_bb[KS];                      // body1 | body2 << 8
    int nsc, work, pad3_[2];   // work: solver work estimate of the substep (input of pbrl_balance_order)
    alignas(8) unsigned long long fr_bar;        // mbarrier the bulk copy completes on
    float fr_u, fr_cyc;
    int fr_phase, fr_pad;
    // warp-uniform loop table copied from the model once per kernel
    int lvl_start[MAXLEVEL + 1];
    int status, niter;
    int phase_lock;  // CTA barriers at stage-group boundaries keep the CTA's warps inside one I-cache window
    unsigned con_mask;   // contact slots with a constraint row this substep
    unsigned act_mask;   // ... with at least one active pyramid edge in the current Newton iterate
    unsigned kbody_mask; // bodies that own such a slot
    unsigned pad2_[2];
#if defined(PBRL_STAGE_PROF)
    long long prof_t;
#endif
};
static_assert(sizeof(Ws) % 16 == 0, "workspaces tile shared memory on 16-byte boundaries");
constexpr int NHD = NV * (NV + 1) / 2;  // packed lower triangle of the dense Hessian
static_assert(offsetof(Ws, K) + sizeof(float) * NB * S21 - offsetof(Ws, LD) >= sizeof(float) * NHD, "dense Hessian spans LD + K");
static_assert(4 * 36 <= NBUF6 && 32 * 8 <= NBUF6 && NV * S6 <= NBUF6 && 8 * NSG <= NB * S21 && NCP * S6 + KS * S6 <= NB * S21, "body-vs-body scratch");

PB_DEV float fast_rcp(float x) {
#if defined(__CUDACC__) && !defined(PBRL_HOST_EMU)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));  // MUFU.RCP, 1 ulp
    return r;
#else
    return 1.0f / x;
#endif
}

PB_DEV float fast_rsqrt(float x) {
#if defined(__CUDACC__) && !defined(PBRL_HOST_EMU)
    return rsqrtf(x);  // MUFU.RSQ, 2 ulp
#else
    return 1.0f / sqrtf(x);
#endif
}

// ---------------------------------------------------------------------------------- warp helpers
template <class Fn>
PB_DEV float wsum(int lane, int n, Fn f) {
    float s = 0.f;
    LANE_LOOP(i, n) s += f(i);
#if defined(__CUDACC__) && !defined(PBRL_HOST_EMU)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
#endif
    (void)lane;
    return s;
}

// two sums at once (the shuffle rounds of the two reductions overlap); f(i, a, b) accumulates item i into a and b
template <class Fn>
PB_DEV void wsum2(int lane, int n, Fn f, float& s0, float& s1) {
    float a = 0.f, b = 0.f;
    LANE_LOOP(i, n) f(i, a, b);
#if defined(__CUDACC__) && !defined(PBRL_HOST_EMU)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ta = __shfl_xor_sync(0xffffffffu, a, o), tb = __shfl_xor_sync(0xffffffffu, b, o);
        a += ta; b += tb;
    }
#endif
    (void)lane;
    s0 = a; s1 = b;
}
// sum of one value per lane
PB_DEV float wreduce(float s) {
#if defined(__CUDACC__) && !defined(PBRL_HOST_EMU)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
#endif
    return s;  // (host emulation: the lane loop already accumulated into one variable)
}
// bit mask of the lanes i < n whose predicate holds
template <class Fn>
PB_DEV unsigned wballot(int lane, int n, Fn f) {
#if defined(__CUDACC__) && !defined(PBRL_HOST_EMU)
    return __ballot_sync(0xffffffffu, lane < n && f(lane));
#else
    unsigned mk = 0;
    for (int i = 0; i < n; i++) mk |= f(i) ? 1u << i : 0u;
    (void)lane;
    return mk;
#endif
}
PB_DEV int first_bit(unsigned mk) {  // index of the lowest set bit (mk != 0)
#if defined(__CUDACC__) && !defined(PBRL_HOST_EMU)
    return __ffs((int)mk) - 1;
#else
    int i = 0;
    while (!((mk >> i) & 1u)) i++;
    return i;
#endif
}

PB_DEV float int_bits_as_float(int i) {
#if defined(__CUDACC__) && !defined(PBRL_HOST_EMU)
    return __int_as_float(i);
#else
    float f; memcpy(&f, &i, 4); return f;
#endif
}
PB_DEV int float_bits_as_int(float f) {
#if defined(__CUDACC__) && !defined(PBRL_HOST_EMU)
    return __float_as_int(f);
#else
    int i; memcpy(&i, &f, 4); return i;
#endif
}
PB_DEV int pop_count(unsigned mk) {
#if defined(__CUDACC__) && !defined(PBRL_HOST_EMU)
    return __popc(mk);
#else
    return __builtin_popcount(mk);
#endif
}

// ---------------------------------------------------------------------------------- row moves (16-byte aligned rows)
#if defined(__CUDACC__) && !defined(PBRL_HOST_EMU)
PB_DEV void ld4(const float* p, float* r) { const float4 a = *reinterpret_cast<const float4*>(p); r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w; }
PB_DEV void st4(float* p, const float* r) { *reinterpret_cast<float4*>(p) = make_float4(r[0], r[1], r[2], r[3]); }
PB_DEV void ld2(const float* p, float* r) { const float2 a = *reinterpret_cast<const float2*>(p); r[0] = a.x; r[1] = a.y; }
PB_DEV void st2(float* p, const float* r) { *reinterpret_cast<float2*>(p) = make_float2(r[0], r[1]); }
#else
PB_DEV void ld4(const float* p, float* r) { for (int k = 0; k < 4; k++) r[k] = p[k]; }
PB_DEV void st4(float* p, const float* r) { for (int k = 0; k < 4; k++) p[k] = r[k]; }
PB_DEV void ld2(const float* p, float* r) { r[0] = p[0]; r[1] = p[1]; }
PB_DEV void st2(float* p, const float* r) { p[0] = r[0]; p[1] = r[1]; }
#endif
PB_DEV void ld6(const float* p, float* r) { ld2(p, r); ld2(p + 2, r + 2); ld2(p + 4, r + 4); }   // 8-byte aligned rows (S6)
PB_DEV void st6(float* p, const float* r) { st2(p, r); st2(p + 2, r + 2); st2(p + 4, r + 4); }
PB_DEV void ld10(const float* p, float* r) { ld4(p, r); ld4(p + 4, r + 4); ld2(p + 8, r + 8); }
PB_DEV void st10(float* p, const float* r) { st4(p, r); st4(p + 4, r + 4); st2(p + 8, r + 8); }
PB_DEV void ld24(const float* p, float* r) { for (int k = 0; k < 6; k++) ld4(p + 4 * k, r + 4 * k); }
PB_DEV void st24(float* p, const float* r) { for (int k = 0; k < 6; k++) st4(p + 4 * k, r + 4 * k); }

// ---------------------------------------------------------------------------------- small math
PB_DEV void cross3(const float* a, const float* b, float* r) {
    float x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
    r[0] = x; r[1] = y; r[2] = z;
}
PB_DEV float dot3(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
PB_DEV float dot6(const float* a, const float* b) {
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3] + a[4] * b[4] + a[5] * b[5];
}
PB_DEV void quat_mul(const float* a, const float* b, float* r) {
    float w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
    float x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
    float y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
    float z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
    r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
PB_DEV void quat_normalize(float* q) {
    float n = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    if (n < 1e-30f) { q[0] = 1.f; q[1] = q[2] = q[3] = 0.f; return; }
    float inv = 1.0f / n;
    q[0] *= inv; q[1] *= inv; q[2] *= inv; q[3] *= inv;
}
PB_DEV void quat_to_mat(const float* q, float* m) {
    float w = q[0], x = q[1], y = q[2], z = q[3];
    m[0] = w * w + x * x - y * y - z * z; m[1] = 2.f * (x * y - w * z); m[2] = 2.f * (x * z + w * y);
    m[3] = 2.f * (x * y + w * z); m[4] = w * w - x * x + y * y - z * z; m[5] = 2.f * (y * z - w * x);
    m[6] = 2.f * (x * z - w * y); m[7] = 2.f * (y * z + w * x); m[8] = w * w - x * x - y * y + z * z;
}
PB_DEV void quat_rotate(const float* q, const float* v, float* r) {  // R(q) v
    float m[9];
    quat_to_mat(q, m);
    float x = m[0] * v[0] + m[1] * v[1] + m[2] * v[2], y = m[3] * v[0] + m[4] * v[1] + m[5] * v[2],
          z = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
    r[0] = x; r[1] = y; r[2] = z;
}
PB_DEV void mat_vec(const float* m, const float* v, float* r) {
    float x = m[0] * v[0] + m[1] * v[1] + m[2] * v[2], y = m[3] * v[0] + m[4] * v[1] + m[5] * v[2],
          z = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
    r[0] = x; r[1] = y; r[2] = z;
}
PB_DEV void matT_vec(const float* m, const float* v, float* r) {
    float x = m[0] * v[0] + m[3] * v[1] + m[6] * v[2], y = m[1] * v[0] + m[4] * v[1] + m[7] * v[2],
          z = m[2] * v[0] + m[5] * v[1] + m[8] * v[2];
    r[0] = x; r[1] = y; r[2] = z;
}
// 10-float spatial inertia about the com times a motion vector
PB_DEV void mul_inert(const float* i, const float* v, float* r) {
    r[0] = i[0] * v[0] + i[3] * v[1] + i[4] * v[2] - i[8] * v[4] + i[7] * v[5];
    r[1] = i[3] * v[0] + i[1] * v[1] + i[5] * v[2] + i[8] * v[3] - i[6] * v[5];
    r[2] = i[4] * v[0] + i[5] * v[1] + i[2] * v[2] - i[7] * v[3] + i[6] * v[4];
    r[3] = i[8] * v[1] - i[7] * v[2] + i[9] * v[3];
    r[4] = i[6] * v[2] - i[8] * v[0] + i[9] * v[4];
    r[5] = i[7] * v[0] - i[6] * v[1] + i[9] * v[5];
}
PB_DEV void cross_motion(const float* v, const float* s, float* r) {
    float a[3], b[3], c[3];
    cross3(v, s, a); cross3(v, s + 3, b); cross3(v + 3, s, c);
    r[0] = a[0]; r[1] = a[1]; r[2] = a[2]; r[3] = b[0] + c[0]; r[4] = b[1] + c[1]; r[5] = b[2] + c[2];
}
PB_DEV void cross_force(const float* v, const float* f, float* r) {
    float a[3], b[3], c[3];
    cross3(v, f, a); cross3(v + 3, f + 3, b); cross3(v, f + 3, c);
    r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2]; r[3] = c[0]; r[4] = c[1]; r[5] = c[2];
}
// symmetric 6x6 stored as 21 floats, row-major upper triangle
PB_DEV int sym21(int r, int c) { return r <= c ? r * 6 - r * (r - 1) / 2 + (c - r) : c * 6 - c * (c - 1) / 2 + (r - c); }

// sine / cosine of a hinge half angle: odd / even polynomials after a reduction to |r| <= pi / 2 (error < 1e-7), branch-free
PB_DEV void sincos_half_angle(float x, float* sn, float* cs) {
    // x = k pi + r, |r| <= pi / 2 (two-term Cody-Waite: exact enough for the |x| < ~100 a joint angle can reach), then
    // sin x = (-1)^k sin r, cos x = (-1)^k cos r: branch-free, no library fallback to inline at each of the 28 hinges
    const float k = rintf(x * 0.31830988618f);
    float r = fmaf(k, -3.14159274101f, x);
    r = fmaf(k, 8.74227766e-8f, r);   // pi = 3.14159274101 (fp32) - 8.74227766e-8
    const float sgn = ((int)k & 1) ? -1.f : 1.f;
    const float z = r * r;
    float s = fmaf(z, -2.5052108e-8f, 2.7557319e-6f);   // Taylor to r^13 / r^12: truncation below 1e-8 on the interval
    s = fmaf(s, z, -1.9841270e-4f); s = fmaf(s, z, 8.3333333e-3f); s = fmaf(s, z, -1.6666667e-1f);
    *sn = sgn * fmaf(r * z, s, r);
    float c = fmaf(z, 2.0876757e-9f, -2.7557319e-7f);
    c = fmaf(c, z, 2.4801587e-5f); c = fmaf(c, z, -1.3888889e-3f); c = fmaf(c, z, 4.1666667e-2f); c = fmaf(c, z, -0.5f);
    *cs = sgn * fmaf(c, z, 1.0f);
}
// ---------------------------------------------------------------------------------- kinematics
// Flat form: everything that does not depend on the parent pose runs lane-parallel over joints / bodies first (joint
// rotations, their product per body, the joint axes in the parent-side frame); the level recursion then only composes
// one quaternion and one offset per body.
PB_DEVN void stage_fk(const DevModel& m, Ws& w, const float* qpos, int lane) {
    float* jq = w.cdof;           // [NJ][4] rotation of each hinge about its axis           (cdof is written by stage_com)
    float* bq = w.cdof + 4 * NJ;  // [NB][4] product of a body's hinge rotations
    float* axl = w.buf6;          // [NJ] x S3: joint axis in the body's pre-joint frame here, in the world frame on return
    static_assert(4 * NJ + 4 * NB <= NV * S6 && S3 * NJ <= NBUF6, "fk scratch");
    LANE_IF(j, NJ) {
        if (m.jnt_type[j] != 0) {
            const int qa = m.jnt_qposadr[j];
            float sn, cs;
            sincos_half_angle(0.5f * (qpos[qa] - m.qpos0[qa]), &sn, &cs);
            jq[4 * j] = cs; jq[4 * j + 1] = m.jnt_axis[j][0] * sn; jq[4 * j + 2] = m.jnt_axis[j][1] * sn; jq[4 * j + 3] = m.jnt_axis[j][2] * sn;
        }
    }
    LANE_IF(z, 1) {
        w.xpos[0] = w.xpos[1] = w.xpos[2] = 0.f;
        w.xquat[0] = 1.f; w.xquat[1] = w.xquat[2] = w.xquat[3] = 0.f;
        for (int k = 0; k < 9; k++) w.xmat[k] = (k % 4 == 0) ? 1.f : 0.f;
    }
    WSYNC();
    LANE_IF(b, NB) {
        const int j0 = m.body_jntadr[b], nj = m.body_jntnum[b];
        if (b > 0 && m.jnt_type[j0] != 0) {
            float q[4] = {jq[4 * j0], jq[4 * j0 + 1], jq[4 * j0 + 2], jq[4 * j0 + 3]};
            for (int c = 0; c < 3; c++) axl[S3 * j0 + c] = m.jnt_axis[j0][c];
            for (int j = j0 + 1; j < j0 + nj; j++) {
                float qn[4];
                quat_rotate(q, m.jnt_axis[j], axl + S3 * j);
                quat_mul(q, jq + 4 * j, qn);
                for (int c = 0; c < 4; c++) q[c] = qn[c];
            }
            for (int c = 0; c < 4; c++) bq[4 * b + c] = q[c];
        }
    }
    WSYNC();
    for (int lv = 1; lv < m.nlevel; lv++) {
        int s = w.lvl_start[lv], n = w.lvl_start[lv + 1] - s;
        LANE_IF(k, n) {
            const int b = m.level_body[s + k], p = m.body_parent[b], j0 = m.body_jntadr[b];
            float pos[3], q[4];
            if (m.jnt_type[j0] == 0) {
                const int qa = m.jnt_qposadr[j0];
                for (int c = 0; c < 3; c++) pos[c] = qpos[qa + c];
                for (int c = 0; c < 4; c++) q[c] = qpos[qa + 3 + c];
                quat_normalize(q);  // mj_kinematics normalises the free quaternion before use ...
            } else {
                float t[3];
                mat_vec(w.xmat + S9 * p, m.body_pos[b], t);
                for (int c = 0; c < 3; c++) pos[c] = w.xpos[S3 * p + c] + t[c];
                quat_mul(w.xquat + S4 * p, bq + 4 * b, q);
            }
            quat_normalize(q);      // ... and every body quaternion after its joints
            for (int c = 0; c < 3; c++) w.xpos[S3 * b + c] = pos[c];
            for (int c = 0; c < 4; c++) w.xquat[S4 * b + c] = q[c];
            float R[9];
            quat_to_mat(q, R);
            for (int c = 0; c < 9; c++) w.xmat[S9 * b + c] = R[c];
        }
        WSYNC();
    }
    // joint axes to the world frame: the pre-joint frame of a hinge body is its parent's frame (body_quat = identity)
    LANE_IF(j, NJ) {
        if (m.jnt_type[j] == 0) { axl[S3 * j] = 0.f; axl[S3 * j + 1] = 0.f; axl[S3 * j + 2] = 1.f; }
        else {
            float a[3] = {axl[S3 * j], axl[S3 * j + 1], axl[S3 * j + 2]};
            mat_vec(w.xmat + S9 * m.body_parent[m.jnt_bodyid[j]], a, axl + S3 * j);
        }
    }
    WSYNC();
}

// com, cinert, cdof  (mjx smooth.com_pos)
PB_DEVN void stage_com(const DevModel& m, Ws& w, int lane) {
    float inv = 1.0f / m.total_mass;
    float* xip = w.cvel;  // scratch: world position of each body's inertial frame (cvel is not live yet)
    LANE_IF(b, NB) {
        float t[3];
        mat_vec(w.xmat + S9 * b, m.body_ipos[b], t);
        for (int c = 0; c < 3; c++) xip[S6 * b + c] = w.xpos[S3 * b + c] + t[c];
    }
    WSYNC();
    const float cx = wsum(lane, NB, [&](int b) { return m.body_mass[b] * xip[S6 * b]; }) * inv;
    const float cy = wsum(lane, NB, [&](int b) { return m.body_mass[b] * xip[S6 * b + 1]; }) * inv;
    const float cz = wsum(lane, NB, [&](int b) { return m.body_mass[b] * xip[S6 * b + 2]; }) * inv;
    LANE_IF(z, 1) { w.com[0] = cx; w.com[1] = cy; w.com[2] = cz; }
    LANE_IF(b, NB) {
        const float* R = w.xmat + S9 * b;
        const float* I = m.body_inertia[b];  // xx yy zz xy xz yz (body frame, about ipos)
        float Iw[6];
        if (I[3] == 0.f && I[4] == 0.f && I[5] == 0.f) {  // principal axes = body axes: Iw = sum_k I_k r_k r_k^T
            Iw[0] = I[0] * R[0] * R[0] + I[1] * R[1] * R[1] + I[2] * R[2] * R[2];
            Iw[1] = I[0] * R[3] * R[3] + I[1] * R[4] * R[4] + I[2] * R[5] * R[5];
            Iw[2] = I[0] * R[6] * R[6] + I[1] * R[7] * R[7] + I[2] * R[8] * R[8];
            Iw[3] = I[0] * R[0] * R[3] + I[1] * R[1] * R[4] + I[2] * R[2] * R[5];
            Iw[4] = I[0] * R[0] * R[6] + I[1] * R[1] * R[7] + I[2] * R[2] * R[8];
            Iw[5] = I[0] * R[3] * R[6] + I[1] * R[4] * R[7] + I[2] * R[5] * R[8];
        } else {
            float A[9] = {I[0], I[3], I[4], I[3], I[1], I[5], I[4], I[5], I[2]}, RA[9];
            for (int r = 0; r < 3; r++)
                for (int c = 0; c < 3; c++) RA[3 * r + c] = R[3 * r] * A[c] + R[3 * r + 1] * A[3 + c] + R[3 * r + 2] * A[6 + c];
            Iw[0] = RA[0] * R[0] + RA[1] * R[1] + RA[2] * R[2];
            Iw[1] = RA[3] * R[3] + RA[4] * R[4] + RA[5] * R[5];
            Iw[2] = RA[6] * R[6] + RA[7] * R[7] + RA[8] * R[8];
            Iw[3] = RA[0] * R[3] + RA[1] * R[4] + RA[2] * R[5];
            Iw[4] = RA[0] * R[6] + RA[1] * R[7] + RA[2] * R[8];
            Iw[5] = RA[3] * R[6] + RA[4] * R[7] + RA[5] * R[8];
        }
        float ms = m.body_mass[b], ox = xip[S6 * b] - cx, oy = xip[S6 * b + 1] - cy, oz = xip[S6 * b + 2] - cz;
        float* ci = w.cinert + S10 * b;
        ci[0] = Iw[0] + ms * (oy * oy + oz * oz);
        ci[1] = Iw[1] + ms * (ox * ox + oz * oz);
        ci[2] = Iw[2] + ms * (ox * ox + oy * oy);
        ci[3] = Iw[3] - ms * ox * oy;
        ci[4] = Iw[4] - ms * ox * oz;
        ci[5] = Iw[5] - ms * oy * oz;
        ci[6] = ms * ox; ci[7] = ms * oy; ci[8] = ms * oz; ci[9] = ms;
    }
    LANE_LOOP(i, NV) {
        int b = m.dof_bodyid[i], kind = m.dof_kind[i];
        float* cd = w.cdof + S6 * i;
        if (kind < 3) {
            for (int c = 0; c < 6; c++) cd[c] = 0.f;
            cd[3 + kind] = 1.f;
        } else {
            float ax[3], off[3];
            if (kind < 6) {
                for (int c = 0; c < 3; c++) ax[c] = w.xmat[S9 * b + 3 * c + (kind - 3)];
            } else {
                for (int c = 0; c < 3; c++) ax[c] = w.buf6[S3 * m.dof_jntid[i] + c];
            }
            off[0] = cx - w.xpos[S3 * b]; off[1] = cy - w.xpos[S3 * b + 1]; off[2] = cz - w.xpos[S3 * b + 2];
            cd[0] = ax[0]; cd[1] = ax[1]; cd[2] = ax[2];
            cross3(ax, off, cd + 3);
        }
    }
    WSYNC();
}

// ---------------------------------------------------------------------------------- dof-lane helpers
// Lane l < NDL owns dof i = l + 3.  The free-joint translations 0..2 have unit cdofs ([0; e_k]), so whatever they need is
// a component of a quantity the lane of dof 3 (same body) already holds; that lane emits them.
struct DofRow {  // the sparse row of dof i: i and its ancestors, nearest ancestor last in `an` order (depth 0 first)
    int i, d, adr;
    unsigned an[MAXANC / 4];
};
PB_DEV DofRow dof_row(const DevModel& m, int i) {
    DofRow r;
    r.i = i; r.d = m.dof_depth[i]; r.adr = m.dof_Madr[i];
    for (int q = 0; q < MAXANC / 4; q++) r.an[q] = reinterpret_cast<const unsigned*>(m.anc_dof[i])[q];
    return r;
}
PB_DEV int row_anc(const DofRow& r, int t) { return (int)((r.an[t >> 2] >> (8 * (t & 3))) & 0xffu); }  // ancestor at depth t < d

// sum of the 6-vector rows src[c] over the bodies c of the subtree of b (bodies are numbered depth first).  Branch-free:
// every lane walks the longest possible subtree with clamped indices and 0/1 weights, so all loads are in flight at once.
PB_DEV void subtree_gather6(const DevModel& m, const float* src, int b, float* f) {
    const int e = m.body_subtree_end[b];
    ld6(src + S6 * b, f);
#pragma unroll
    for (int k = 1; k < NB - 1; k++) {
        const bool on = b + k < e;
        float t[6];
        ld6(src + S6 * (on ? b + k : b), t);
        const float wt = on ? 1.f : 0.f;
        for (int q = 0; q < 6; q++) f[q] = fmaf(wt, t[q], f[q]);
    }
}
// generalised force of the spatial forces F (one per body): out_i = cdof_i . sum_{subtree(body_i)} F  (+ extra per dof)
template <class Extra>
PB_DEV void force_to_dofs(const DevModel& m, Ws& w, const float* F, float* out, Extra extra, int lane) {
    LANE_IF(l, NDL) {
        const int i = l + 3;
        float f[6], cd[6];
        subtree_gather6(m, F, m.dof_bodyid[i], f);
        ld6(w.cdof + S6 * i, cd);
        out[i] = dot6(cd, f) + extra(i);
        if (i == 3) for (int k = 0; k < 3; k++) out[k] = f[3 + k] + extra(k);
    }
    WSYNC();
}

// composite inertias + sparse M  (mjx smooth.crb).  One pass, lane = dof: the composite inertia of the dof's body is
// gathered over the subtree, applied to the dof's cdof in registers, and dotted with the cdofs of the row's columns.
PB_DEVN void stage_crb(const DevModel& m, Ws& w, int lane) {
    LANE_IF(l, NDL) {
        const DofRow r = dof_row(m, l + 3);
        const int b = m.dof_bodyid[r.i], e = m.body_subtree_end[b];
        float ci[10], cd[6], f[6];
        ld10(w.cinert + S10 * b, ci);
#pragma unroll
        for (int k = 1; k < NB - 1; k++) {  // branch-free subtree walk (see subtree_gather6)
            const bool on = b + k < e;
            float t[10];
            ld10(w.cinert + S10 * (on ? b + k : b), t);
            const float wt = on ? 1.f : 0.f;
            for (int q = 0; q < 10; q++) ci[q] = fmaf(wt, t[q], ci[q]);
        }
        ld6(w.cdof + S6 * r.i, cd);
        mul_inert(ci, cd, f);
        w.M[r.adr] = dot6(cd, f) + m.dof_armature[r.i];
#pragma unroll
        for (int t = 0; t < MAXANC; t++) {
            const bool on = t < r.d;
            float cj[6];
            ld6(w.cdof + S6 * (on ? row_anc(r, t) : r.i), cj);
            const float v = dot6(cj, f);
            if (on) w.M[r.adr + r.d - t] = v;
        }
        if (r.i == 3) {  // rows 0..2: cdof = [0; e_k]  ->  M[k][j] = (I_c [0; e_k])_{3+j}: the mass on the diagonal, 0 elsewhere
            for (int k = 0; k < 3; k++)
                for (int j = 0; j <= k; j++) w.M[m.dof_Madr[k] + (k - j)] = (j == k ? ci[9] + m.dof_armature[k] : 0.f);
        }
    }
    WSYNC();
}

// V_b = sum over the dofs on the root path of body b of vec6_i * x_i (lane = body; the path is the row of b's last dof).
// Branch-free over the longest path (clamped index, zero weight), so the loads of all terms overlap.
template <bool HASX>
PB_DEV void path_product6(const DevModel& m, const float* vec6, const float* x, int b, float* v) {
    const DofRow r = dof_row(m, m.body_dofadr[b] + m.body_dofnum[b] - 1);
    float t[6];
    ld6(vec6 + S6 * r.i, t);
    const float xi = HASX ? x[r.i] : 1.f;
    for (int k = 0; k < 6; k++) v[k] = t[k] * xi;
#pragma unroll
    for (int q = 0; q < MAXANC; q++) {
        const bool on = q < r.d;
        const int j = on ? row_anc(r, q) : r.i;
        const float xj = on ? (HASX ? x[j] : 1.f) : 0.f;
        ld6(vec6 + S6 * j, t);
        for (int k = 0; k < 6; k++) v[k] = fmaf(t[k], xj, v[k]);
    }
}

// body velocities, cdof_dot, bias accelerations  (mjx smooth.com_vel + first half of smooth.rne): three lane-parallel
// passes (bodies, dofs, bodies), each lane summing along its root path in registers.
PB_DEVN void stage_comvel(const DevModel& m, Ws& w, const float* qvel, int lane) {
    float* da = w.buf6;  // [NV] x S6: cdof_dot_i * qvel_i
    LANE_IF(b, NB) {
        float v[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (b > 0) path_product6<true>(m, w.cdof, qvel, b, v);
        st6(w.cvel + S6 * b, v);
    }
    WSYNC();
    LANE_LOOP(i, NV) {
        const int kind = m.dof_kind[i];
        float t[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (kind >= 3) {  // translations of the free joint carry no cdof_dot
            const int b = m.dof_bodyid[i], p = m.body_parent[b], d0 = m.body_dofadr[b];
            float v[6], cd[6];
            ld6(w.cvel + S6 * p, v);
            // velocity before this dof: parent body + the earlier dofs of the same body (free joint: the translations only)
            const int nprev = kind < 6 ? 3 : i - d0;
            for (int k = 0; k < nprev; k++) {
                const float qd = qvel[d0 + k];
                ld6(w.cdof + S6 * (d0 + k), cd);
                for (int c = 0; c < 6; c++) v[c] = fmaf(cd[c], qd, v[c]);
            }
            ld6(w.cdof + S6 * i, cd);
            cross_motion(v, cd, t);
            const float qd = qvel[i];
            for (int c = 0; c < 6; c++) t[c] *= qd;
        }
        st6(da + S6 * i, t);
    }
    WSYNC();
    LANE_IF(b, NB) {
        float a[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (b > 0) path_product6<false>(m, da, nullptr, b, a);
        a[3] -= m.gx; a[4] -= m.gy; a[5] -= m.gz;  // every body inherits -gravity from the world
        st6(w.cacc + S6 * b, a);
    }
    WSYNC();
}

// qfrc_bias  (second half of mjx smooth.rne)
PB_DEVN void stage_bias(const DevModel& m, Ws& w, int lane) {
    LANE_IF(b, NB) {
        float ci[10], a[6], v[6], t[6], u[6], f[6];
        ld10(w.cinert + S10 * b, ci);
        ld6(w.cacc + S6 * b, a);
        ld6(w.cvel + S6 * b, v);
        mul_inert(ci, a, t);
        mul_inert(ci, v, u);
        cross_force(v, u, f);
        for (int c = 0; c < 6; c++) f[c] = (b == 0) ? 0.f : t[c] + f[c];
        st6(w.F + S6 * b, f);
    }
    WSYNC();
    force_to_dofs(m, w, w.F, w.bias, [](int) { return 0.f; }, lane);
}

// ---------------------------------------------------------------------------------- tree-sparse L^T D L
// LD holds the unscaled rows: LD[adr(k)+a] = L[k, anc_a(k)] * D[k]; Dinv[k] = 1/D[k].
PB_DEV void init_ws_tables(const DevModel& m, Ws& w, int lane) {
    LANE_LOOP(l, MAXLEVEL + 1) w.lvl_start[l] = m.level_start[l];
    WSYNC();
}

// x <- (L^T D L)^-1 x for the tree-sparse matrix in w.LD (row k: k and its ancestors, MuJoCo mj_factorI order).
// LD is overwritten with the unscaled factor rows LD[adr(k)+a] = L[k, anc_a(k)] * D[k]; w.Dinv[k] = 1/D[k].
#if defined(__CUDACC__) && !defined(PBRL_HOST_EMU)
// Register-resident, shuffle-driven: no shared-memory round trip or warp barrier inside the elimination.
//   layout A (factor + forward substitution): lane t <-> depth t.  Row k lives in one register R[k]: lane t holds the
//     entry of k's ancestor at depth t, lane depth(k) the diagonal.  Eliminating k needs, per ancestor row, ONE broadcast
//     (L[k, anc]) -- the other factor of every rank-1 term is already in the right lane.  The right-hand side rides along
//     as X[slot] (slot = chain of the dof tree, lane = depth), so the forward substitution costs no extra latency.
//   layout B (back substitution): lane <-> dof.  Row k sits in lane k's registers, x_k is pushed level by level to all
//     descendants with one shuffle per level.
// Dofs are eliminated deepest first, so the five leaf chains of the humanoid interleave and hide each other's latency.
namespace ldl {
constexpr unsigned FULL = 0xffffffffu;
constexpr int NSTAGE = 14;  // rotating 16-float staging rows in Ws::buf6 (idle during the factorisation)
template <int K, int T>
struct AncRows {  // rows of K's ancestors at depths T-1 .. 0 lose L[K, anc] * row K; c[t] = row K's entry at depth t
    PB_MEMBER void update(float (&R)[NV], const float (&c)[16], const float rks) {  // rks = -row K / D[K], per lane
        if constexpr (T > 0) {
            constexpr int J = topo::anc_at(K, T - 1);
            R[J] = fmaf(c[T - 1], rks, R[J]);
            AncRows<K, T - 1>::update(R, c, rks);
        }
    }
};
template <int K, int S>
struct AncRhs {  // X of K's ancestors (found in slots S-1 .. 0) loses L[K, anc] * D[K] * s
    PB_MEMBER void run(float (&X)[NSLOT], const float rs, int lane) {
        if constexpr (S > 0) {
            constexpr unsigned MK = topo::anc_mask(K, S - 1);
            if constexpr (MK != 0u) { if ((MK >> lane) & 1u) X[S - 1] -= rs; }
            AncRhs<K, S - 1>::run(X, rs, lane);
        }
    }
};
template <int I>
struct Eliminate {
    // Row K is all-gathered through a 16-float staging row: ONE store, a warp barrier and (depth+1)/4 broadcast vector
    // loads give every lane the whole row (a chain of per-entry shuffles serialises on the shuffle scoreboard).
    PB_MEMBER void run(float (&R)[NV], float (&X)[NSLOT], bool& bad, float* stage, float* ld_lane, int lane) {
        constexpr int K = topo::elim_order(I), d = topo::depth(K), S = topo::slot(K);
        const float rk = R[K];
        float4* st = reinterpret_cast<float4*>(stage + 16 * (I % NSTAGE));
        if (lane < 16) reinterpret_cast<float*>(st)[lane] = rk;
        if (lane <= d) ld_lane[topo::madr(K) + d] = rk;  // final row K (the back substitution reads it lane-per-dof)
        float xs = 0.f;
        if constexpr (d > 0) xs = __shfl_sync(FULL, X[S], d);
        __syncwarp();
        float c[16];
#pragma unroll
        for (int q = 0; q < (d + 4) / 4; q++) {
            const float4 v = st[q];
            c[4 * q] = v.x; c[4 * q + 1] = v.y; c[4 * q + 2] = v.z; c[4 * q + 3] = v.w;
        }
        const float diag = c[d];
        bad |= !(diag > 0.f);
        if constexpr (d > 0) {
            const float rks = rk * -fast_rcp(diag);
            AncRows<K, d>::update(R, c, rks);
            AncRhs<K, NSLOT>::run(X, -rks * xs, lane);
        }
        if constexpr (I + 1 < NV) Eliminate<I + 1>::run(R, X, bad, stage, ld_lane, lane);
    }
};
template <int K>
struct RowIO {
    PB_MEMBER void load(float (&R)[NV], const float* ld_lane, int lane) {  // ld_lane = LD - lane
        // unconditional: lanes beyond the diagonal read whatever precedes the row (still inside the workspace) and
        // carry garbage that is never broadcast, never enters a valid lane and is never stored
        R[K] = ld_lane[topo::madr(K) + topo::depth(K)];
        if constexpr (K + 1 < NV) RowIO<K + 1>::load(R, ld_lane, lane);
    }
    PB_MEMBER void store(const float (&R)[NV], float* ld_lane, int lane) {
        constexpr int d = topo::depth(K);
        if (lane <= d) ld_lane[topo::madr(K) + d] = R[K];
        if constexpr (K + 1 < NV) RowIO<K + 1>::store(R, ld_lane, lane);
    }
};
template <int S>
struct SlotIO {
    PB_MEMBER void load(float (&X)[NSLOT], const float* x, int lane) {
        const int t = lane - topo::slot_fd(S);
        X[S] = (t >= 0 && t < topo::slot_len(S)) ? x[topo::slot_first(S) + t] : 0.f;
        if constexpr (S + 1 < NSLOT) SlotIO<S + 1>::load(X, x, lane);
    }
    PB_MEMBER void store(const float (&X)[NSLOT], float* x, int lane) {
        const int t = lane - topo::slot_fd(S);
        if (t >= 0 && t < topo::slot_len(S)) x[topo::slot_first(S) + t] = X[S];
        if constexpr (S + 1 < NSLOT) SlotIO<S + 1>::store(X, x, lane);
    }
};
template <int T>
struct BackLevel {  // every dof deeper than T receives L[k, anc_T(k)] * x[anc_T(k)]
    PB_MEMBER void run(const float (&cl)[MAXANC], const float (&ch)[MAXANC], const unsigned (&sl)[MAXANC / 4],
                       const unsigned (&sh)[MAXANC / 4], float& xl, float& xh) {
        if constexpr (T < topo::max_depth()) {
            const int a_lo = (sl[T / 4] >> (8 * (T % 4))) & 0xffu, a_hi = (sh[T / 4] >> (8 * (T % 4))) & 0xffu;
            float v_lo = __shfl_sync(FULL, xl, a_lo & 31), v_hi = __shfl_sync(FULL, xl, a_hi & 31);
            if constexpr (topo::hi_source_at(T)) {
                const float w_lo = __shfl_sync(FULL, xh, a_lo & 31), w_hi = __shfl_sync(FULL, xh, a_hi & 31);
                v_lo = a_lo >= 32 ? w_lo : v_lo;
                v_hi = a_hi >= 32 ? w_hi : v_hi;
            }
            xl = fmaf(-cl[T], v_lo, xl);  // cl[T] = 0 when the dof is not deeper than T
            xh = fmaf(-ch[T], v_hi, xh);
            BackLevel<T + 1>::run(cl, ch, sl, sh, xl, xh);
        }
    }
};
}  // namespace ldl

PB_DEVN void ldl_factor_solve(const DevModel& m, Ws& w, float* x, int lane) {
    static_assert(offsetof(Ws, LD) >= 32 * sizeof(float), "row loads reach up to 31 floats below LD");
    static_assert(sizeof(w.buf6) >= ldl::NSTAGE * 16 * sizeof(float) && offsetof(Ws, buf6) % 16 == 0 && sizeof(Ws) % 16 == 0, "staging rows");
    float R[NV], X[NSLOT];
    bool bad = false;
    ldl::RowIO<0>::load(R, w.LD - lane, lane);
    ldl::SlotIO<0>::load(X, x, lane);
    ldl::Eliminate<0>::run(R, X, bad, w.buf6, w.LD - lane, lane);
    ldl::SlotIO<0>::store(X, x, lane);  // x <- L^-T x
    if (bad && lane == 0) w.status |= PBRL_STATUS_PIVOT;
    __syncwarp();
    // back substitution, lane <-> dof (lanes 0 .. NV-33 also own dof lane + 32)
    const int klo = lane, khi = lane + 32 < NV ? lane + 32 : lane;
    const bool has_hi = lane + 32 < NV;
    const int dlo = m.dof_depth[klo], dhi = has_hi ? m.dof_depth[khi] : 0;
    const float* rlo = w.LD + m.dof_Madr[klo] + dlo;  // entry of the ancestor at depth t: rlo[-t]
    const float* rhi = w.LD + m.dof_Madr[khi] + dhi;
    const float ilo = fast_rcp(rlo[-dlo]), ihi = fast_rcp(rhi[-dhi]);  // 1 / D[k]: the diagonal sits at offset 0 of the row
    w.Dinv[klo] = ilo;
    if (has_hi) w.Dinv[khi] = ihi;
    float cl[MAXANC], ch[MAXANC];
#pragma unroll
    for (int t = 0; t < MAXANC; t++) {
        cl[t] = t < dlo ? rlo[-t] * ilo : 0.f;
        ch[t] = t < dhi ? rhi[-t] * ihi : 0.f;
    }
    unsigned sl[MAXANC / 4], sh[MAXANC / 4];
#pragma unroll
    for (int q = 0; q < MAXANC / 4; q++) {
        sl[q] = reinterpret_cast<const unsigned*>(m.anc_dof[klo])[q];
        sh[q] = reinterpret_cast<const unsigned*>(m.anc_dof[khi])[q];
    }
    float xl = x[klo] * ilo, xh = has_hi ? x[khi] * ihi : 0.f;  // x <- D^-1 L^-T x
    __syncwarp();
    ldl::BackLevel<0>::run(cl, ch, sl, sh, xl, xh);
    x[klo] = xl;
    if (has_hi) x[khi] = xh;
    __syncwarp();
}
#else
// host emulation (tests only): the same elimination written as plain loops over the model tables
PB_DEVN void ldl_factor_solve(const DevModel& m, Ws& w, float* x, int lane) {
    (void)lane;
    for (int k = NV - 1; k >= 1; k--) {
        const int d = m.dof_depth[k], adr = m.dof_Madr[k];
        const float inv = 1.0f / w.LD[adr];
        for (int a = 1; a <= d; a++) {
            const int ia = m.M_col[adr + a];
            for (int t = 0; t + a <= d; t++) w.LD[m.dof_Madr[ia] + t] -= (w.LD[adr + a] * inv) * w.LD[adr + a + t];
        }
    }
    for (int i = 0; i < NV; i++) {
        const float piv = w.LD[m.dof_Madr[i]];
        if (!(piv > 0.f)) w.status |= PBRL_STATUS_PIVOT;
        w.Dinv[i] = 1.0f / piv;
    }
    for (int k = NV - 1; k >= 1; k--) {
        const int d = m.dof_depth[k], adr = m.dof_Madr[k];
        const float s = x[k] * w.Dinv[k];
        for (int a = 1; a <= d; a++) x[m.M_col[adr + a]] -= w.LD[adr + a] * s;
    }
    for (int i = 0; i < NV; i++) x[i] *= w.Dinv[i];
    for (int k = 1; k < NV; k++) {
        const int d = m.dof_depth[k], adr = m.dof_Madr[k];
        float s = 0.f;
        for (int a = 1; a <= d; a++) s += w.LD[adr + a] * x[m.M_col[adr + a]];
        x[k] -= s * w.Dinv[k];
    }
}
#endif

// ---------------------------------------------------------------------------------- tree products
// out = M x, matrix-free (includes armature); leaves V[b] = spatial velocity of body b for the generalised velocity x
PB_DEVN void tree_Mx(const DevModel& m, Ws& w, const float* x, float* out, int lane) {
    LANE_IF(b, NB) {
        float v[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, f[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (b > 0) {
            float ci[10];
            path_product6<true>(m, w.cdof, x, b, v);
            ld10(w.cinert + S10 * b, ci);
            mul_inert(ci, v, f);
        }
        st6(w.V + S6 * b, v);
        st6(w.F + S6 * b, f);
    }
    WSYNC();
    force_to_dofs(m, w, w.F, out, [&](int i) { return m.dof_armature[i] * x[i]; }, lane);
}

// ---------------------------------------------------------------------------------- smooth dynamics
PB_DEVN void smooth_position(const DevModel& m, Ws& w, const float* qpos, int lane) {
    stage_fk(m, w, qpos, lane);
    stage_com(m, w, lane);
    stage_crb(m, w, lane);
}
PB_DEVN void smooth_velocity(const DevModel& m, Ws& w, const float* qvel, int lane) {
    stage_comvel(m, w, qvel, lane);
    stage_bias(m, w, lane);
}

// ---------------------------------------------------------------------------------- collision with the floor plane
// impedance of a row at penetration `pos` (mjx constraint._kbi)
// general solimp power: kept out of line (one copy of the powf code), MuJoCo's default power 2 takes the squares below
PB_DEVN float impedance_y_general(float x, float mid, float power) {
    if (x < mid) return (1.0f / powf(mid, power - 1.f)) * powf(x, power);
    return 1.f - (1.0f / powf(1.f - mid, power - 1.f)) * powf(1.f - x, power);
}
PB_DEV float impedance(const DevModel& m, float pos) {
    float x = fabsf(pos) / m.sol_width, y;
    if (m.sol_power != 2.f) y = impedance_y_general(x, m.sol_mid, m.sol_power);
    else if (x < m.sol_mid) y = (1.0f / m.sol_mid) * (x * x);
    else y = 1.f - (1.0f / (1.f - m.sol_mid)) * ((1.f - x) * (1.f - x));
    float imp = m.sol_dmin + y * (m.sol_dmax - m.sol_dmin);
    imp = fminf(fmaxf(imp, m.sol_dmin), m.sol_dmax);
    if (x > 1.f) imp = m.sol_dmax;
    return imp;
}
// first maximum of score[i] + (ok[i] ? 0 : -1e6), added in fp32 as mjx _manifold_points does (jax's default dtype): with every
// candidate masked out the sums collapse onto fp32's grid at 1e6 (spacing 0.0625) and that rounding decides the pick
PB_DEV int argmax_masked(const float* score, const bool* ok, int n) {
    int best = 0;
    float bs = 0.f;  // running best score (no indexed read of score[])
#pragma unroll
    for (int i = 0; i < n; i++) {
        const float s = ok[i] ? score[i] : score[i] + -1e6f;
        if (i == 0 || s > bs) { best = i; bs = s; }
    }
    return best;
}
// pyramid edge e of a contact: w = n + sg*mu*t, j6 = [r x w ; w]
PB_DEV void contact_j6(const float* r, const float* t1, float mu, int e, float* j6) {
    float sg = (e & 1) ? -mu : mu, wv[3];
    if (e < 2) { wv[0] = sg * t1[0]; wv[1] = sg * t1[1]; wv[2] = 1.f; }
    else { wv[0] = -sg * t1[1]; wv[1] = sg * t1[0]; wv[2] = 1.f; }  // t2 = n x t1, n = +z
    cross3(r, wv, j6);
    j6[3] = wv[0]; j6[4] = wv[1]; j6[5] = wv[2];
}

PB_DEVN void stage_collide(const DevModel& m, Ws& w, int lane) {
    LANE_IF(c, NC) {
        int b = m.con_body[c], kind = m.con_kind[c];
        const float* R = w.xmat + S9 * b;
        float gp[3], t[3], dist, p[3], t1[3] = {0.f, 1.f, 0.f};
        mat_vec(R, m.con_gpos[c], t);
        for (int k = 0; k < 3; k++) gp[k] = w.xpos[S3 * b + k] + t[k];
        if (kind == CON_SPHERE) {
            float r = m.con_size[c][0];
            dist = gp[2] - m.plane_z - r;
            p[0] = gp[0]; p[1] = gp[1]; p[2] = gp[2] - (r + 0.5f * dist);
        } else if (kind == CON_CAPSULE) {
            float zl[3] = {0.f, 0.f, 1.f}, al[3], ax[3], r = m.con_size[c][0], hl = m.con_size[c][1];
            quat_rotate(m.con_gquat[c], zl, al);
            mat_vec(R, al, ax);
            float bn = sqrtf(ax[0] * ax[0] + ax[1] * ax[1]);
            if (bn >= 0.5f) { t1[0] = ax[0] / bn; t1[1] = ax[1] / bn; t1[2] = 0.f; }
            float sg = m.con_sub[c] == 0 ? hl : -hl, ce[3] = {gp[0] + sg * ax[0], gp[1] + sg * ax[1], gp[2] + sg * ax[2]};
            dist = ce[2] - m.plane_z - r;
            p[0] = ce[0]; p[1] = ce[1]; p[2] = ce[2] - (r + 0.5f * dist);
        } else {  // box corner manifold (mjx plane_convex + _manifold_points)
            // vertices are generated from their index (bit 2-k of v = sign along axis k): no indexed local arrays
            const float* sz = m.con_size[c];
            auto vert = [&](int v, float* o) { o[0] = (v & 4) ? sz[0] : -sz[0]; o[1] = (v & 2) ? sz[1] : -sz[1]; o[2] = (v & 1) ? sz[2] : -sz[2]; };
            float Rl[9], Rg[9], sup[8], sc[16], pl[3], nl[3], d3[3], e3[3], ab[3], ac[3], bc[3], va[3], vb[3], vc[3], vv[3], smax;
            bool ok[16];
            quat_to_mat(m.con_gquat[c], Rl);
            for (int r = 0; r < 3; r++)
                for (int q = 0; q < 3; q++) Rg[3 * r + q] = R[3 * r] * Rl[q] + R[3 * r + 1] * Rl[3 + q] + R[3 * r + 2] * Rl[6 + q];
            d3[0] = -gp[0]; d3[1] = -gp[1]; d3[2] = m.plane_z - gp[2];
            matT_vec(Rg, d3, pl);
            nl[0] = Rg[6]; nl[1] = Rg[7]; nl[2] = Rg[8];
            smax = -1e30f;
#pragma unroll
            for (int v = 0; v < 8; v++) {
                vert(v, vv);
                for (int k = 0; k < 3; k++) d3[k] = pl[k] - vv[k];
                sup[v] = dot3(d3, nl);
                smax = fmaxf(smax, sup[v]);
            }
            float thr = fmaxf(0.f, smax - 1e-3f);
#pragma unroll
            for (int v = 0; v < 8; v++) { ok[v] = sup[v] > thr; ok[8 + v] = ok[v]; sc[v] = 0.f; }
            int ia = argmax_masked(sc, ok, 8);
            vert(ia, va);
#pragma unroll
            for (int v = 0; v < 8; v++) {
                vert(v, vv);
                for (int k = 0; k < 3; k++) d3[k] = va[k] - vv[k];
                sc[v] = dot3(d3, d3);
            }
            int ib = argmax_masked(sc, ok, 8);
            vert(ib, vb);
            for (int k = 0; k < 3; k++) e3[k] = va[k] - vb[k];
            cross3(nl, e3, ab);
#pragma unroll
            for (int v = 0; v < 8; v++) {
                vert(v, vv);
                for (int k = 0; k < 3; k++) d3[k] = va[k] - vv[k];
                sc[v] = fabsf(dot3(d3, ab));
            }
            int ic = argmax_masked(sc, ok, 8);
            vert(ic, vc);
            for (int k = 0; k < 3; k++) e3[k] = va[k] - vc[k];
            cross3(nl, e3, ac);
            for (int k = 0; k < 3; k++) e3[k] = vb[k] - vc[k];
            cross3(nl, e3, bc);
#pragma unroll
            for (int v = 0; v < 8; v++) {
                vert(v, vv);
                for (int k = 0; k < 3; k++) d3[k] = vb[k] - vv[k];
                sc[v] = fabsf(dot3(d3, bc));
                for (int k = 0; k < 3; k++) d3[k] = va[k] - vv[k];
                sc[8 + v] = fabsf(dot3(d3, ac));
            }
            int id = argmax_masked(sc, ok, 16) % 8;
            const int sub = m.con_sub[c], me = sub == 0 ? ia : sub == 1 ? ib : sub == 2 ? ic : id;
            const bool uniq = (sub < 1 || ia != me) && (sub < 2 || ib != me) && (sub < 3 || ic != me);
            vert(me, vv);
            for (int k = 0; k < 3; k++) d3[k] = pl[k] - vv[k];
            dist = uniq ? -dot3(d3, nl) : 1.f;
            mat_vec(Rg, vv, t);
            p[0] = gp[0] + t[0]; p[1] = gp[1] + t[1]; p[2] = gp[2] + t[2] - 0.5f * dist;
        }
        w.con_dist[c] = dist;
        for (int k = 0; k < 3; k++) w.con_r[SR * c + k] = p[k] - w.com[k];
        w.con_t1[2 * c] = t1[0]; w.con_t1[2 * c + 1] = t1[1];  // tangent 1 lies in the floor plane
    }
    WSYNC();
}

// ---------------------------------------------------------------------------------- body-vs-body contacts
// Sphere / capsule candidates (mjx collision_primitive.py: _sphere_sphere, sphere_capsule, capsule_capsule with
// math.closest_segment_point / closest_segment_to_segment_points) and sphere-box candidates
// (collision_convex.py: _sphere_convex).  MJX evaluates every candidate
// every step and keeps a row block per candidate; only contacts inside the margin have non-zero rows, so only those
// (at most KS) are kept here.  Capsule-box and box-box candidates are not evaluated (DevModel::nbox_pairs).
PB_DEV void seg_point(const float* a, const float* b, const float* pt, float* out) {
    const float ab[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, ap[3] = {pt[0] - a[0], pt[1] - a[1], pt[2] - a[2]};
    float t = dot3(ap, ab) * fast_rcp(dot3(ab, ab) + 1e-6f);
    t = fminf(fmaxf(t, 0.f), 1.f);
    for (int k = 0; k < 3; k++) out[k] = a[k] + t * ab[k];
}
PB_DEV float normalize_with_norm(const float* x, float* n) {
    const float ss = dot3(x, x), nrm = ss > 0.f ? ss * fast_rsqrt(ss) : 0.f, inv = fast_rcp(nrm + (nrm == 0.f ? 1e-6f : 0.f));
    for (int k = 0; k < 3; k++) n[k] = x[k] * inv;
    return nrm;
}
PB_DEV void seg_seg(const float* a0, const float* a1, const float* b0, const float* b1, float* best_a, float* best_b) {
    float va[3], vb[3], da[3], db[3], am[3], bm[3], tr[3], na[3], nb[3], e1[3], e2[3];
    for (int k = 0; k < 3; k++) { va[k] = a1[k] - a0[k]; vb[k] = b1[k] - b0[k]; }
    const float ha = 0.5f * normalize_with_norm(va, da), hb = 0.5f * normalize_with_norm(vb, db);
    for (int k = 0; k < 3; k++) { am[k] = a0[k] + da[k] * ha; bm[k] = b0[k] + db[k] * hb; tr[k] = am[k] - bm[k]; }
    const float dd = dot3(da, db), dat = dot3(da, tr), dbt = dot3(db, tr);
    float ta = (-dat + dd * dbt) * fast_rcp((1.f - dd * dd) + 1e-6f), tb = dbt + ta * dd;
    ta = fminf(fmaxf(ta, -ha), ha);
    tb = fminf(fmaxf(tb, -hb), hb);
    for (int k = 0; k < 3; k++) { best_a[k] = am[k] + da[k] * ta; best_b[k] = bm[k] + db[k] * tb; }
    seg_point(a0, a1, best_b, na);
    seg_point(b0, b1, best_a, nb);
    for (int k = 0; k < 3; k++) { e1[k] = na[k] - best_b[k]; e2[k] = nb[k] - best_a[k]; }
    const bool first = dot3(e1, e1) < dot3(e2, e2);
    for (int k = 0; k < 3; k++) { best_a[k] = first ? na[k] : best_a[k]; best_b[k] = first ? best_b[k] : nb[k]; }
}
// tangents of a contact normal (mjx math.make_frame): t1 from the y axis unless the normal is close to it, t2 = n x t1
PB_DEV void frame_tangents(const float* n, float* t1, float* t2) {
    float b[3] = {0.f, 0.f, 0.f};
    if (n[1] > -0.5f && n[1] < 0.5f) b[1] = 1.f; else b[2] = 1.f;
    const float dt = dot3(n, b);
    for (int k = 0; k < 3; k++) b[k] -= n[k] * dt;
    const float inv = fast_rsqrt(dot3(b, b));
    for (int k = 0; k < 3; k++) t1[k] = b[k] * inv;
    cross3(n, t1, t2);
}
// pyramid row e of body-vs-body contact s: w = n +- mu t, j6 = [r x w ; w]; J x = j6 . (V[body2] - V[body1])
PB_DEV void self_j6(const Ws& w, float mu, int s, int e, float* j6) {
    float n[4], r[4], t1[3], t2[3], wv[3];
    ld4(w.sc_n + 4 * s, n);
    ld4(w.sc_r + 4 * s, r);
    frame_tangents(n, t1, t2);
    const float sg = (e & 1) ? -mu : mu;
    for (int k = 0; k < 3; k++) wv[k] = n[k] + sg * (e < 2 ? t1[k] : t2[k]);
    cross3(r, wv, j6);
    j6[3] = wv[0]; j6[4] = wv[1]; j6[5] = wv[2];
}
PB_DEV float self_row_dot(const Ws& w, const float* j6, const float* vec6, int s) {
    const int bb = w.sc_bb[s];
    float v1[6], v2[6];
    ld6(vec6 + S6 * (bb & 0xff), v1);
    ld6(vec6 + S6 * (bb >> 8), v2);
    for (int k = 0; k < 6; k++) v2[k] -= v1[k];
    return dot6(j6, v2);
}

// ---- capsule / box against a box (mjx collision_convex.py: capsule_convex, convex_convex with _create_contact_manifold, _clip,
// _clip_edge_to_planes, _manifold_points.  Conventions where the published algorithm leaves a choice: box faces listed +x, -x, +y,
// -y, +z, -z with corners counter-clockwise seen from outside, the first maximum wins, contact positions midway between the two
// surfaces, manifold picks that coincide with an earlier one switched off like plane_convex does for repeated indices).  These candidates (14 capsule-box,
// 1 box-box) are far apart in almost every pose: a lane runs the out-of-line functions below only when its pair passes a
// bounding-volume test, and the warp skips them when no lane does.
PB_DEV void project_on_plane(const float* pt, const float* plane_pt, const float* n, float* out) {
    const float t[3] = {pt[0] - plane_pt[0], pt[1] - plane_pt[1], pt[2] - plane_pt[2]}, dd = dot3(t, n);
    for (int k = 0; k < 3; k++) out[k] = pt[k] - dd * n[k];
}
PB_DEV void seg_plane_point(const float* a, const float* b, const float* p0, const float* n, float* out) {
    const float ab[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, dd = dot3(p0, n), denom = dot3(n, ab);
    float t = (dd - dot3(n, a)) / (denom + (denom == 0.f ? 1e-6f : 0.f));
    t = fminf(fmaxf(t, 0.f), 1.f);
    for (int k = 0; k < 3; k++) out[k] = a[k] + t * ab[k];
}
// clips the edge (e0, e1) against four planes (point, outward normal); returns the mask of the clipped edge
PB_DEV bool clip_edge_4(const float* e0, const float* e1, const float (*ppt)[3], const float (*pn)[3], float (*out)[3]) {
    float best0 = 0.f, best1 = 0.f, n0[3] = {0.f, 0.f, 0.f}, n1[3] = {0.f, 0.f, 0.f};
    const float e01[3] = {e1[0] - e0[0], e1[1] - e0[1], e1[2] - e0[2]}, e10[3] = {-e01[0], -e01[1], -e01[2]};
    bool both = false;
    for (int k = 0; k < 4; k++) {
        const float t0[3] = {e0[0] - ppt[k][0], e0[1] - ppt[k][1], e0[2] - ppt[k][2]}, t1[3] = {e1[0] - ppt[k][0], e1[1] - ppt[k][1], e1[2] - ppt[k][2]};
        const bool in0 = dot3(t0, pn[k]) > 1e-6f, in1 = dot3(t1, pn[k]) > 1e-6f;
        both |= in0 && in1;
        float cand[3], q0[3], q1[3];
        seg_plane_point(e0, e1, ppt[k], pn[k], cand);
        for (int c = 0; c < 3; c++) { q0[c] = in0 ? cand[c] : e0[c]; q1[c] = in1 ? cand[c] : e1[c]; }
        const float u0[3] = {q0[0] - e0[0], q0[1] - e0[1], q0[2] - e0[2]}, u1[3] = {q1[0] - e1[0], q1[1] - e1[1], q1[2] - e1[2]};
        const float d0 = dot3(u0, e01), d1 = dot3(u1, e10);
        if (k == 0 || d0 > best0) { best0 = d0; for (int c = 0; c < 3; c++) n0[c] = q0[c]; }
        if (k == 0 || d1 > best1) { best1 = d1; for (int c = 0; c < 3; c++) n1[c] = q1[c]; }
    }
    bool mask = !both;
    for (int c = 0; c < 3; c++) { out[0][c] = mask ? n0[c] : e0[c]; out[1][c] = mask ? n1[c] : e1[c]; }
    const float dn[3] = {out[0][0] - out[1][0], out[0][1] - out[1][1], out[0][2] - out[1][2]};
    if (dot3(e10, dn) < 0.f) mask = false;  // the clipped end points crossed
    return mask;
}
// capsule (world centre c, axis x half length a, radius r) against a box (world centre bc, frame R, half sizes): two contact
// slots, written as [pos(3), dist, normal(3)] rows; the normal points from the capsule to the box.
// Register-only: after the face of least penetration is known the problem is rotated into a canonical frame in which that face
// is the +z face with corners (-,-) (+,-) (+,+) (-,+) (a proper rotation of the box frame: x' y' = the two other axes, swapped
// for a negative face), so that every array index below is a compile-time constant -- no local memory on this path.
struct V3 { float x, y, z; };
PB_DEV float pick3(const float* v, int i) { return i == 0 ? v[0] : (i == 1 ? v[1] : v[2]); }
PB_DEV float dotv(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
// closest point of the segment (a, b) to the plane through p0 with normal n (mjx _closest_segment_point_plane)
PB_DEV V3 seg_plane_v(V3 a, V3 b, V3 p0, V3 n) {
    const V3 ab = {b.x - a.x, b.y - a.y, b.z - a.z};
    const float dd = dotv(p0, n), denom = dotv(n, ab);
    float t = (dd - dotv(n, a)) / (denom + (denom == 0.f ? 1e-6f : 0.f));
    t = fminf(fmaxf(t, 0.f), 1.f);
    return V3{a.x + t * ab.x, a.y + t * ab.y, a.z + t * ab.z};
}
PB_DEVN void capsule_box_dev(const float* c, const float* a, float r, const float* bc, const float* R, const float* half, float* out) {
    const float dl[3] = {c[0] - bc[0], c[1] - bc[1], c[2] - bc[2]};
    float cp[3], ax[3], q0[3], q1[3];
    matT_vec(R, dl, cp);
    matT_vec(R, a, ax);
    for (int k = 0; k < 3; k++) { q0[k] = cp[k] - ax[k]; q1[k] = cp[k] + ax[k]; }
    float best = 0.f, bsg = 1.f;
    int bk = 0;
    bool separated = false;
#pragma unroll
    for (int f = 0; f < 6; f++) {  // support of face f: the lower end of the inflated segment against the face plane
        const int k = f >> 1;
        const float sg = (f & 1) ? -1.f : 1.f;
        float sup = fminf(sg * q0[k], sg * q1[k]) - r - half[k];
        if (sup >= 0.f) { sup = -1e12f; separated = true; }
        if (f == 0 || sup > best) { best = sup; bk = k; bsg = sg; }
    }
    // canonical frame: z' = sg * axis k, (x', y') = the other two axes in cyclic order, swapped for a negative face
    const int ia = bk == 2 ? 0 : bk + 1, ib = ia == 2 ? 0 : ia + 1, ix = bsg > 0.f ? ia : ib, iy = bsg > 0.f ? ib : ia;
    const V3 p0 = {pick3(q0, ix), pick3(q0, iy), bsg * pick3(q0, bk)}, p1 = {pick3(q1, ix), pick3(q1, iy), bsg * pick3(q1, bk)};
    const float hx = pick3(half, ix), hy = pick3(half, iy), hz = pick3(half, bk);
    // side planes of the face in edge order (edge i runs from corner i-1 to corner i): x' = -hx, y' = -hy, x' = +hx, y' = +hy
    const V3 pp[4] = {{-hx, hy, hz}, {-hx, -hy, hz}, {hx, -hy, hz}, {hx, hy, hz}};       // edge start points (corner i-1)
    const V3 pe[4] = {{-hx, -hy, hz}, {hx, -hy, hz}, {hx, hy, hz}, {-hx, hy, hz}};       // edge end points (corner i)
    // (edge x face normal, unnormalised like mjx: length of the edge)
    const V3 pn[4] = {{-2.f * hy, 0.f, 0.f}, {0.f, -2.f * hx, 0.f}, {2.f * hy, 0.f, 0.f}, {0.f, 2.f * hx, 0.f}};
    // mjx _clip_edge_to_planes on the segment
    const V3 e01 = {p1.x - p0.x, p1.y - p0.y, p1.z - p0.z}, e10 = {-e01.x, -e01.y, -e01.z};
    V3 n0 = p0, n1 = p1;
    float best0 = 0.f, best1 = 0.f;
    bool both = false;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const V3 t0 = {p0.x - pp[k].x, p0.y - pp[k].y, p0.z - pp[k].z}, t1 = {p1.x - pp[k].x, p1.y - pp[k].y, p1.z - pp[k].z};
        const bool in0 = dotv(t0, pn[k]) > 1e-6f, in1 = dotv(t1, pn[k]) > 1e-6f;
        both |= in0 && in1;
        const V3 cand = seg_plane_v(p0, p1, pp[k], pn[k]);
        const V3 c0 = in0 ? cand : p0, c1 = in1 ? cand : p1;
        const V3 u0 = {c0.x - p0.x, c0.y - p0.y, c0.z - p0.z}, u1 = {c1.x - p1.x, c1.y - p1.y, c1.z - p1.z};
        const float d0 = dotv(u0, e01), d1 = dotv(u1, e10);
        if (k == 0 || d0 > best0) { best0 = d0; n0 = c0; }
        if (k == 0 || d1 > best1) { best1 = d1; n1 = c1; }
    }
    bool mask = !both;
    V3 cl0 = mask ? n0 : p0, cl1 = mask ? n1 : p1;
    { const V3 dn = {cl0.x - cl1.x, cl0.y - cl1.y, cl0.z - cl1.z}; if (dotv(e10, dn) < 0.f) mask = false; }
    mask = mask && !separated;
    // face contacts: the clipped end points lowered by the radius, against the face plane z' = hz
    cl0.z -= r; cl1.z -= r;
    V3 lp0 = {cl0.x, cl0.y, 0.5f * (cl0.z + hz)}, lp1 = {cl1.x, cl1.y, 0.5f * (cl1.z + hz)};
    V3 ln0 = {0.f, 0.f, -1.f}, ln1 = {0.f, 0.f, -1.f};
    float pen0 = mask ? hz - cl0.z : -1.f, pen1 = mask ? hz - cl1.z : -1.f;
    // a potential edge contact: the edge of that face closest to the capsule's segment
    float bd = 0.f;
    V3 ec = {0.f, 0.f, 0.f}, cc = {0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        float e_[3], c_[3];
        const float s0[3] = {pp[k].x, pp[k].y, pp[k].z}, s1[3] = {pe[k].x, pe[k].y, pe[k].z}, t0[3] = {p0.x, p0.y, p0.z}, t1[3] = {p1.x, p1.y, p1.z};
        seg_seg(s0, s1, t0, t1, e_, c_);
        const float t[3] = {e_[0] - c_[0], e_[1] - c_[1], e_[2] - c_[2]}, dd = dot3(t, t);
        if (k == 0 || dd < bd) { bd = dd; ec = V3{e_[0], e_[1], e_[2]}; cc = V3{c_[0], c_[1], c_[2]}; }
    }
    {
        const float ea[3] = {cc.x - ec.x, cc.y - ec.y, cc.z - ec.z};
        float eax[3];
        const float ed = normalize_with_norm(ea, eax);
        if (r - ed > 0.f && ed > 0.f) {
            lp0 = V3{0.5f * (ec.x + (cc.x - eax[0] * r)), 0.5f * (ec.y + (cc.y - eax[1] * r)), 0.5f * (ec.z + (cc.z - eax[2] * r))};
            ln0 = V3{-eax[0], -eax[1], -eax[2]};
            pen0 = r - ed;
            pen1 = -1.f;
        }
    }
    // back to the box frame (component ix <- x', iy <- y', bk <- sg z'), then to the world
    auto emit = [&](V3 lp, V3 ln, float pen, float* o) {
        float pl[3], nl[3], pw[3], nw[3];
        for (int k = 0; k < 3; k++) {
            pl[k] = k == ix ? lp.x : (k == iy ? lp.y : bsg * lp.z);
            nl[k] = k == ix ? ln.x : (k == iy ? ln.y : bsg * ln.z);
        }
        mat_vec(R, pl, pw);
        mat_vec(R, nl, nw);
        for (int k = 0; k < 3; k++) { o[k] = pw[k] + bc[k]; o[4 + k] = nw[k]; }
        o[3] = -pen;
    };
    emit(lp0, ln0, pen0, out);
    emit(lp1, ln1, pen1, out + 8);
}
// index of the first maximum of score(i), i < n <= 32 (a serial scan with a strict comparison): one candidate per lane, butterfly
// on (score, index) pairs where the lower index wins a tie
template <class Fn>
PB_DEV int warg_max_first(int lane, int n, Fn score) {
#if defined(__CUDACC__) && !defined(PBRL_HOST_EMU)
    float s = lane < n ? score(lane) : -3.0e38f;
    int i = lane < n ? lane : 255;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float so = __shfl_xor_sync(0xffffffffu, s, o);
        const int io = __shfl_xor_sync(0xffffffffu, i, o);
        if (so > s || (so == s && io < i)) { s = so; i = io; }
    }
    return i;
#else
    int b = 0;
    float bs = 0.f;
    for (int i = 0; i < n; i++) { const float sc = score(i); if (i == 0 || sc > bs) { bs = sc; b = i; } }
    (void)lane;
    return b;
#endif
}
// corner i (counter-clockwise seen from outside) of face f of a box, faces listed +x, -x, +y, -y, +z, -z
PB_DEV void box_face_corner(const float* half, int f, int i, float* v) {
    const int k = f >> 1, a = k == 2 ? 0 : k + 1, b = a == 2 ? 0 : a + 1;
    const int j = (f & 1) ? ((4 - i) & 3) : i;  // reversed order for the negative face
    const float su = (j == 1 || j == 2) ? 1.f : -1.f, sv = (j >= 2) ? 1.f : -1.f;
    v[k] = ((f & 1) ? -1.f : 1.f) * half[k]; v[a] = su * half[a]; v[b] = sv * half[b];
}
// box against box: four contact slots [pos(3), dist, normal(3)], normal from box 1 to box 2; slot 0 goes to o0, slots 1..3 to
// ox, ox + 8, ox + 16.  Warp-cooperative (one call per pair, every lane of the env's warp takes part): the 15 separating axes,
// the 8 edge clips and the 16 manifold candidates each get a lane, the picks are warp arg-max reductions -- the serial form of the
// same arithmetic kept one lane busy for ~35 k cycles while the CTA waited at its next barrier.
// `scratch`: 256 floats of the caller's shared-memory workspace.
PB_DEV void box_box_warp(const float* c1, const float* R1, const float* h1, const float* c2, const float* R2, const float* h2, float* o0,
                         float* ox, float* scratch, int lane) {
    float (*ax)[3] = reinterpret_cast<float (*)[3]>(scratch);               // [15][3]
    float (*rv)[3] = reinterpret_cast<float (*)[3]>(scratch + 45);          // [4][3] reference face corners
    float (*iv)[3] = reinterpret_cast<float (*)[3]>(scratch + 57);          // [4][3] incident face corners
    float (*rp0)[3] = reinterpret_cast<float (*)[3]>(scratch + 69);         // side planes of the reference / incident face
    float (*rpn)[3] = reinterpret_cast<float (*)[3]>(scratch + 81);
    float (*ip0)[3] = reinterpret_cast<float (*)[3]>(scratch + 93);
    float (*ipn)[3] = reinterpret_cast<float (*)[3]>(scratch + 105);
    float (*cand)[3] = reinterpret_cast<float (*)[3]>(scratch + 117);       // [16][3]
    float (*onref)[3] = reinterpret_cast<float (*)[3]>(scratch + 165);      // [16][3]
    float (*rproj)[3] = reinterpret_cast<float (*)[3]>(scratch + 213);      // [4][3]
    int* cmask = reinterpret_cast<int*>(scratch + 225);                     // [16]
    float* bax = scratch + 241;                                             // [3] best axis, [3] its overlap
    static_assert(245 <= 256, "box-box scratch");
    auto slot = [&](int q) { return q == 0 ? o0 : ox + 8 * (q - 1); };
    const float dc[3] = {c2[0] - c1[0], c2[1] - c1[1], c2[2] - c1[2]};
    LANE_IF(a, 15) {
        float v[3];
        if (a < 3) { for (int k = 0; k < 3; k++) v[k] = R1[3 * k + a]; }
        else if (a < 6) { for (int k = 0; k < 3; k++) v[k] = R2[3 * k + a - 3]; }
        else {
            const int i = (a - 6) / 3, j = (a - 6) % 3;
            const float u1[3] = {R1[i], R1[3 + i], R1[6 + i]}, u2[3] = {R2[j], R2[3 + j], R2[6 + j]};
            cross3(u1, u2, v);
        }
        for (int k = 0; k < 3; k++) ax[a][k] = v[k];
    }
    WSYNC();
    // overlap along axis a (negative: separated) and the signed unit axis from box 1 towards box 2
    auto sat_axis = [&](int a, float* L) {
        const float nn = sqrtf(dot3(ax[a], ax[a]));
        const bool degenerate = nn < 1e-6f;
        for (int k = 0; k < 3; k++) L[k] = degenerate ? ax[a][k] : ax[a][k] / nn;
        float r1 = 0.f, r2 = 0.f;
        for (int i = 0; i < 3; i++) { r1 += h1[i] * fabsf(dot3(ax[i], L)); r2 += h2[i] * fabsf(dot3(ax[3 + i], L)); }
        const float pc = dot3(dc, L), d1 = r1 + r2 - pc, d2 = r1 + r2 + pc;
        const float sg = d1 > d2 ? -1.f : 1.f;
        for (int k = 0; k < 3; k++) L[k] *= sg;
        return degenerate ? 1e6f : fminf(d1, d2);
    };
    const int besti = warg_max_first(lane, 15, [&](int a) { float L[3]; return -sat_axis(a, L); });   // first minimum of the overlap
    LANE_IF(a, 15) if (a == besti) { float L[3]; sat_axis(a, L); for (int k = 0; k < 3; k++) bax[k] = L[k]; }
    WSYNC();
    const float bestax[3] = {bax[0], bax[1], bax[2]};
    LANE_IF(q, 4) {
        float* o = slot(q);
        for (int k = 0; k < 3; k++) { o[k] = 0.5f * (c1[k] + c2[k]); o[4 + k] = bestax[k]; }
        o[3] = 1.f;
    }
    WSYNC();
    if (besti >= 6) {  // edge against edge (warp-uniform)
        LANE_IF(z, 1) {
            const int i = (besti - 6) / 3, j = (besti - 6) % 3;
            float a0[3], a1[3], b0[3], b1[3], pa[3], pb[3];
            for (int k = 0; k < 3; k++) { a0[k] = c1[k]; b0[k] = c2[k]; }
            for (int q = 0; q < 3; q++) {
                if (q != i) { const float sq = dot3(ax[q], bestax) > 0.f ? 1.f : -1.f; for (int k = 0; k < 3; k++) a0[k] += sq * h1[q] * ax[q][k]; }
                if (q != j) { const float sq = dot3(ax[3 + q], bestax) > 0.f ? -1.f : 1.f; for (int k = 0; k < 3; k++) b0[k] += sq * h2[q] * ax[3 + q][k]; }
            }
            for (int k = 0; k < 3; k++) {
                a1[k] = a0[k] + h1[i] * ax[i][k]; a0[k] -= h1[i] * ax[i][k];
                b1[k] = b0[k] + h2[j] * ax[3 + j][k]; b0[k] -= h2[j] * ax[3 + j][k];
            }
            seg_seg(a0, a1, b0, b1, pa, pb);
            const float t[3] = {pb[0] - pa[0], pb[1] - pa[1], pb[2] - pa[2]};
            for (int k = 0; k < 3; k++) o0[k] = 0.5f * (pa[k] + pb[k]);
            o0[3] = dot3(t, bestax);
        }
        WSYNC();
        return;
    }
    // face against face: reference face = the face most aligned with the axis (every lane derives the same choice)
    float amax = 0.f, bmin = 0.f;
    int af = 0, bfc = 0;
    for (int f = 0; f < 6; f++) {
        const float sg = (f & 1) ? -1.f : 1.f, da = sg * dot3(ax[f >> 1], bestax), db = sg * dot3(ax[3 + (f >> 1)], bestax);
        if (f == 0 || da > amax) { amax = da; af = f; }
        if (f == 0 || db < bmin) { bmin = db; bfc = f; }
    }
    const bool ref1 = amax >= -bmin;  // the reference face belongs to box 1
    const float *rc = ref1 ? c1 : c2, *rm = ref1 ? R1 : R2, *ic = ref1 ? c2 : c1, *im = ref1 ? R2 : R1, *rh = ref1 ? h1 : h2, *ih = ref1 ? h2 : h1;
    const int rf = ref1 ? af : bfc, inf = ref1 ? bfc : af;
    float rn[3], inn[3];
    for (int k = 0; k < 3; k++) {   // face normals: +- the face's axis in the world
        rn[k] = ((rf & 1) ? -1.f : 1.f) * rm[3 * k + (rf >> 1)];
        inn[k] = ((inf & 1) ? -1.f : 1.f) * im[3 * k + (inf >> 1)];
    }
    LANE_IF(q, 4) {
        float lv[3];
        box_face_corner(rh, rf, q, lv);
        mat_vec(rm, lv, rv[q]);
        for (int k = 0; k < 3; k++) rv[q][k] += rc[k];
        box_face_corner(ih, inf, q, lv);
        mat_vec(im, lv, iv[q]);
        for (int k = 0; k < 3; k++) iv[q][k] += ic[k];
    }
    WSYNC();
    const float dpl = dot3(iv[0], inn), den = dot3(rn, inn);
    LANE_IF(q, 4) {
        float e[3];
        for (int k = 0; k < 3; k++) { rp0[q][k] = rv[(q + 3) & 3][k]; e[k] = rv[q][k] - rp0[q][k]; }
        cross3(e, rn, rpn[q]);
        for (int k = 0; k < 3; k++) { ip0[q][k] = iv[(q + 3) & 3][k]; e[k] = iv[q][k] - ip0[q][k]; }
        cross3(e, inn, ipn[q]);
        // reference corners moved onto the incident plane along the reference normal
        const float t = (dpl - dot3(rv[q], inn)) / (den + (den == 0.f ? 1e-6f : 0.f));
        for (int k = 0; k < 3; k++) rproj[q][k] = rv[q][k] + t * rn[k];
    }
    WSYNC();
    LANE_IF(l, 8) {  // lanes 0..3: incident edges against the reference side planes; 4..7: projected reference edges against the incident ones
        const int q = l & 3;
        float out2[2][3];
        const bool mk = l < 4 ? clip_edge_4(ip0[q], iv[q], rp0, rpn, out2) : clip_edge_4(rproj[(q + 3) & 3], rproj[q], ip0, ipn, out2);
        for (int k = 0; k < 3; k++) { cand[2 * l][k] = out2[0][k]; cand[2 * l + 1][k] = out2[1][k]; }
        cmask[2 * l] = cmask[2 * l + 1] = mk ? 1 : 0;
    }
    WSYNC();
    LANE_IF(q, 16) {
        const float t[3] = {cand[q][0] - rv[0][0], cand[q][1] - rv[0][1], cand[q][2] - rv[0][2]};
        project_on_plane(cand[q], rv[0], rn, onref[q]);
        cmask[q] = (cmask[q] && (-dot3(t, rn) > 1e-6f)) ? 1 : 0;  // behind the reference face
    }
    WSYNC();
    // mjx _manifold_points over the 16 candidates (scores: first maximum of score + (mask ? 0 : -1e6), added in fp32)
    const unsigned mk16 = wballot(lane, 16, [&](int i) { return cmask[i] != 0; });
    const int ia = mk16 ? first_bit(mk16) : 0;
    auto masked = [&](int i, float sc) { return sc + (cmask[i] ? 0.f : -1e6f); };
    auto dist2 = [&](int from, int i) { const float t[3] = {onref[from][0] - onref[i][0], onref[from][1] - onref[i][1], onref[from][2] - onref[i][2]}; return dot3(t, t); };
    const int ib = warg_max_first(lane, 16, [&](int i) { return masked(i, dist2(ia, i)); });
    float e3[3], ab[3], ac[3], bc[3];
    for (int k = 0; k < 3; k++) e3[k] = onref[ia][k] - onref[ib][k];
    cross3(rn, e3, ab);
    auto along = [&](int from, const float* dir, int i) {
        const float t[3] = {onref[from][0] - onref[i][0], onref[from][1] - onref[i][1], onref[from][2] - onref[i][2]};
        return fabsf(dot3(t, dir));
    };
    const int ic3 = warg_max_first(lane, 16, [&](int i) { return masked(i, along(ia, ab, i)); });
    for (int k = 0; k < 3; k++) e3[k] = onref[ia][k] - onref[ic3][k];
    cross3(rn, e3, ac);
    for (int k = 0; k < 3; k++) e3[k] = onref[ib][k] - onref[ic3][k];
    cross3(rn, e3, bc);
    const int d1 = warg_max_first(lane, 16, [&](int i) { return masked(i, along(ib, bc, i)); });
    const int d2 = warg_max_first(lane, 16, [&](int i) { return masked(i, along(ia, ac, i)); });
    const int id = masked(d2, along(ia, ac, d2)) > masked(d1, along(ib, bc, d1)) ? d2 : d1;
    LANE_IF(q, 4) {
        const int idx[4] = {ia, ib, ic3, id};
        const int p = q == 0 ? ia : (q == 1 ? ib : (q == 2 ? ic3 : id));
        bool unique = true;
        for (int r2 = 0; r2 < 3; r2++) {  // a pick that coincides with an earlier one (within 10 microns) is switched off
            if (r2 >= q) break;
            const float u[3] = {onref[p][0] - onref[idx[r2]][0], onref[p][1] - onref[idx[r2]][1], onref[p][2] - onref[idx[r2]][2]};
            unique = unique && !(dot3(u, u) < 1e-10f);
        }
        const float t[3] = {cand[p][0] - onref[p][0], cand[p][1] - onref[p][1], cand[p][2] - onref[p][2]};
        float* o = slot(q);
        o[3] = (cmask[p] && unique) ? dot3(t, rn) : 1.f;   // dist = -penetration = (incident - reference) . n_ref
        for (int k = 0; k < 3; k++) o[k] = 0.5f * (cand[p][k] + onref[p][k]);
    }
    WSYNC();
}

PB_DEVN void stage_collide_self(const DevModel& m, Ws& w, int lane) {
    if (m.nsp == 0) {
        LANE_IF(z, 1) w.nsc = 0;
        WSYNC();
        return;
    }
    float* gc = w.K;            // [NSG] x 4: world centre of each geom            (K is idle outside the Newton iteration)
    float* ga = w.K + 4 * NSG;  // [NSG] x 4: world axis x half length (zero for spheres)
    float* pr = w.buf6;         // [32] x 8: pos(3), dist, normal(3) of the pair a lane evaluated  (buf6 is idle here)
    float* px = w.K + 8 * NSG;  // [NXROW] x 8: the further contact slots of the capsule-box (1 each) / box-box (3 each) pairs
    constexpr int NXROW = 24;
    static_assert(8 * NSG + 8 * NXROW <= NB * S21 && NM >= 300, "narrow-phase scratch (box_box_warp works in the idle factor workspace LD)");
    // row of slot q >= 1 of convex pair e (counted from the first capsule-box pair)
    auto xrow = [&](int e, int q) { return e < m.nsp_cb ? e : m.nsp_cb + 3 * (e - m.nsp_cb) + (q - 1); };
    LANE_IF(g, NSG) {
        const int b = m.sg_body[g];
        const float* R = w.xmat + S9 * b;
        float t[3], a[3];
        mat_vec(R, m.sg_pos[g], t);
        mat_vec(R, m.sg_axis[g], a);
        for (int k = 0; k < 3; k++) { gc[4 * g + k] = w.xpos[S3 * b + k] + t[k]; ga[4 * g + k] = a[k] * m.sg_hl[g]; }
    }
    WSYNC();
    const int n_one = m.nsp_sc + m.nsp_cc + m.nsp_ss + m.nsp_sb;  // pairs with one contact slot; then capsule-box (2), box-box (4)
    int total = 0;
    // contacts of this round (one row per lane in `pr`, lanes of `act`) join the list; a full list keeps the deepest KS
    auto commit = [&](unsigned act) {
        LANE_IF(l, 32) {
            if ((act >> l) & 1u) {
                const int slot = total + pop_count(act & ((1u << l) - 1u));
                if (slot < KS) {
                    const float* o = pr + 8 * l;
                    for (int k = 0; k < 3; k++) { w.sc_r[4 * slot + k] = o[k] - w.com[k]; w.sc_n[4 * slot + k] = o[4 + k]; }
                    w.sc_r[4 * slot + 3] = 0.f;
                    w.sc_n[4 * slot + 3] = o[3];
                    w.sc_bb[slot] = float_bits_as_int(o[7]);
                }
            }
        }
        WSYNC();
        if (total + pop_count(act) > KS) {  // (warp-uniform, rare) list full: a deeper contact replaces the shallowest one kept
            LANE_IF(z, 1) {
                int skip = KS > total ? KS - total : 0;  // the first `skip` contacts of this round got slots above
                for (unsigned rest = act; rest != 0u; rest &= rest - 1u) {
                    if (skip > 0) { skip--; continue; }
                    const float* o = pr + 8 * first_bit(rest);
                    int worst = 0;
                    for (int c = 1; c < KS; c++) if (w.sc_n[4 * c + 3] > w.sc_n[4 * worst + 3]) worst = c;
                    if (o[3] < w.sc_n[4 * worst + 3]) {
                        for (int k = 0; k < 3; k++) { w.sc_r[4 * worst + k] = o[k] - w.com[k]; w.sc_n[4 * worst + k] = o[4 + k]; }
                        w.sc_n[4 * worst + 3] = o[3];
                        w.sc_bb[worst] = float_bits_as_int(o[7]);
                    }
                }
            }
        }
        total += pop_count(act);
        WSYNC();
    };
    // ---- pairs with one contact slot: 32 per pass, ordered so that a pass runs (mostly) one distance function
    for (int pass = 0; pass * 32 < n_one; pass++) {
        const unsigned act = wballot(lane, 32, [&](int l) {
            const int p = l + 32 * pass;
            if (p >= n_one) return false;
            const int g1 = m.sp_g1[p], g2 = m.sp_g2[p];
            const float *c1 = gc + 4 * g1, *c2 = gc + 4 * g2, *h1 = ga + 4 * g1, *h2 = ga + 4 * g2;
            float* o = pr + 8 * l;
            float p1[3], p2[3];
            if (p < m.nsp_sc) {  // sphere - capsule
                const float b0[3] = {c2[0] - h2[0], c2[1] - h2[1], c2[2] - h2[2]}, b1[3] = {c2[0] + h2[0], c2[1] + h2[1], c2[2] + h2[2]};
                for (int k = 0; k < 3; k++) p1[k] = c1[k];
                seg_point(b0, b1, c1, p2);
            } else if (p < m.nsp_sc + m.nsp_cc) {  // capsule - capsule
                const float a0[3] = {c1[0] - h1[0], c1[1] - h1[1], c1[2] - h1[2]}, a1[3] = {c1[0] + h1[0], c1[1] + h1[1], c1[2] + h1[2]};
                const float b0[3] = {c2[0] - h2[0], c2[1] - h2[1], c2[2] - h2[2]}, b1[3] = {c2[0] + h2[0], c2[1] + h2[1], c2[2] + h2[2]};
                seg_seg(a0, a1, b0, b1, p1, p2);
            } else if (p < m.nsp_sc + m.nsp_cc + m.nsp_ss) {  // sphere - sphere
                for (int k = 0; k < 3; k++) { p1[k] = c1[k]; p2[k] = c2[k]; }
            } else {  // sphere - box (mjx _sphere_convex on the six faces): p2 = closest point of the chosen face, as a zero-radius geom
                const float* R = w.xmat + S9 * m.sg_body[g2];
                const float* half = m.sg_axis[g2];
                const float dl0[3] = {c1[0] - c2[0], c1[1] - c2[1], c1[2] - c2[2]}, r = m.sg_r[g1];
                float q[3], pt[3], best = -1e12f;
                int bk = 0;
                float bsgn = 1.f;
                matT_vec(R, dl0, q);
                for (int k = 0; k < 3; k++) {
                    const float sp = q[k] - half[k] - r, sm = -q[k] - half[k] - r;   // support of the +k / -k face
                    if (sp < 0.f && sp > best) { best = sp; bk = k; bsgn = 1.f; }
                    if (sm < 0.f && sm > best) { best = sm; bk = k; bsgn = -1.f; }
                }
                for (int k = 0; k < 3; k++) pt[k] = fminf(fmaxf(q[k], -half[k]), half[k]);
                for (int k = 0; k < 3; k++) if (k == bk) pt[k] = bsgn * half[k];
                mat_vec(R, pt, p2);
                for (int k = 0; k < 3; k++) { p1[k] = c1[k]; p2[k] += c2[k]; }
            }
            const float r1 = m.sg_r[g1], r2 = p < m.nsp_sc + m.nsp_cc + m.nsp_ss ? m.sg_r[g2] : 0.f, dl[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
            float n[3];
            const float nrm = normalize_with_norm(dl, n);
            if (nrm == 0.f) { n[0] = 1.f; n[1] = 0.f; n[2] = 0.f; }
            const float dist = nrm - (r1 + r2);
            for (int k = 0; k < 3; k++) { o[k] = p1[k] + n[k] * (r1 + 0.5f * dist); o[4 + k] = n[k]; }
            o[3] = dist;
            o[7] = int_bits_as_float((int)m.sg_body[g1] | ((int)m.sg_body[g2] << 8));
            return dist < m.self_margin;
        });
        WSYNC();
        commit(act);
    }
    // ---- capsule-box (two slots) and box-box (four slots) pairs: lane e = pair n_one + e.  Off the hot path: a lane runs the
    // out-of-line convex functions only if its pair passes a bounding test, and the warp skips the section if no lane does.
    const int n_cvx = m.nsp_cb + m.nsp_bb;
    if (n_cvx > 0) {
        // box-box (foot against foot): separating-axis test along the six face normals, one lane per axis, with the contact margin.
        // A pair separated along a face normal has no contact inside the margin (the feet of a standing or walking humanoid are).
        const unsigned bb_axes = wballot(lane, 32, [&](int l) {
            const int q = l / 6, a = l - 6 * q;
            if (q >= m.nsp_bb) return false;
            const int p = n_one + m.nsp_cb + q, g1 = m.sp_g1[p], g2 = m.sp_g2[p];
            const float *R1 = w.xmat + S9 * m.sg_body[g1], *R2 = w.xmat + S9 * m.sg_body[g2], *h1 = m.sg_axis[g1], *h2 = m.sg_axis[g2];
            const float *Ra = a < 3 ? R1 : R2, *Rb = a < 3 ? R2 : R1, *ha = a < 3 ? h1 : h2, *hb = a < 3 ? h2 : h1;
            const int i = a < 3 ? a : a - 3;
            const float L[3] = {Ra[i], Ra[3 + i], Ra[6 + i]};   // column i: face normal
            const float d[3] = {gc[4 * g2] - gc[4 * g1], gc[4 * g2 + 1] - gc[4 * g1 + 1], gc[4 * g2 + 2] - gc[4 * g1 + 2]};
            float reach = ha[i] + m.self_margin;
            for (int j = 0; j < 3; j++) reach += hb[j] * fabsf(L[0] * Rb[j] + L[1] * Rb[3 + j] + L[2] * Rb[6 + j]);
            return fabsf(dot3(d, L)) > reach;   // separated along this axis
        });
        unsigned need_ = wballot(lane, 32, [&](int e) {
            if (e >= n_cvx) return false;
            const int p = n_one + e, g1 = m.sp_g1[p], g2 = m.sp_g2[p];
            if (e < m.nsp_cb) {  // the capsule's segment in the box frame: interval test per box axis, inflated by radius + margin
                const float *c1 = gc + 4 * g1, *c2 = gc + 4 * g2, *h1 = ga + 4 * g1, *R = w.xmat + S9 * m.sg_body[g2], *half = m.sg_axis[g2];
                const float dl[3] = {c1[0] - c2[0], c1[1] - c2[1], c1[2] - c2[2]}, pad = m.sg_r[g1] + m.self_margin;
                float cp[3], ax[3];
                matT_vec(R, dl, cp);
                matT_vec(R, h1, ax);
                bool overlap = true;
                for (int k = 0; k < 3; k++) overlap = overlap && fabsf(cp[k]) - fabsf(ax[k]) <= half[k] + pad;
                // ... and the three axes ax x e_k (the segment projects onto a point): a gap wider than pad along any unit
                // direction bounds the distance from below.  A diagonal shin or thigh over a foot passes the interval test
                // six times more often than it touches; these axes reject most of those.
                for (int k = 0; k < 3; k++) {
                    const int i = (k + 1) % 3, j = (k + 2) % 3;   // L = ax x e_k: L_i = ax_j, L_j = -ax_i, L_k = 0
                    const float gap = fabsf(cp[i] * ax[j] - cp[j] * ax[i]) - (half[i] * fabsf(ax[j]) + half[j] * fabsf(ax[i]));
                    if (gap > 0.f && gap * gap > pad * pad * (ax[i] * ax[i] + ax[j] * ax[j])) overlap = false;   // gap / |L| > pad
                }
                return overlap;
            }
            return ((bb_axes >> (6 * (e - m.nsp_cb))) & 63u) == 0u;   // not separated along any face normal
        });
        PROF_COUNT(62);
#if defined(PBRL_DEV_CONVEX_MASK)   // dev ablation: drop convex pairs after the prefilter
        need_ &= PBRL_DEV_CONVEX_MASK;
#endif
        const unsigned need = need_;
#if defined(PBRL_HOST_EMU) && defined(PBRL_EMU_TRACE)
        g_emu_cnt[0]++; if (need) g_emu_cnt[1]++; if (need & ((1u << m.nsp_cb) - 1u)) g_emu_cnt[2]++; if (need >> m.nsp_cb) g_emu_cnt[3]++;
        g_emu_cnt[4] += pop_count(need);
#endif
        if (need != 0u) {  // (warp-uniform, rare)
            PROF_COUNT(63);
            for (int e = m.nsp_cb; e < n_cvx; e++) {  // box-box pairs (the two feet): the whole warp works on one pair
                if (!((need >> e) & 1u)) continue;
                const int p = n_one + e, g1 = m.sp_g1[p], g2 = m.sp_g2[p];
                box_box_warp(gc + 4 * g1, w.xmat + S9 * m.sg_body[g1], m.sg_axis[g1], gc + 4 * g2, w.xmat + S9 * m.sg_body[g2], m.sg_axis[g2],
                             pr + 8 * e, px + 8 * xrow(e, 1), w.LD, lane);
            }
            const unsigned act = wballot(lane, 32, [&](int e) {
                if (!((need >> e) & 1u)) return false;
                const int p = n_one + e, g1 = m.sp_g1[p], g2 = m.sp_g2[p];
                const float bodies = int_bits_as_float((int)m.sg_body[g1] | ((int)m.sg_body[g2] << 8));
                float* o = pr + 8 * e;
                if (e < m.nsp_cb) {  // capsule-box: one lane per pair
                    float res[16];
                    capsule_box_dev(gc + 4 * g1, ga + 4 * g1, m.sg_r[g1], gc + 4 * g2, w.xmat + S9 * m.sg_body[g2], m.sg_axis[g2], res);
                    for (int k = 0; k < 7; k++) o[k] = res[k];
                    float* x = px + 8 * xrow(e, 1);
                    for (int k = 0; k < 7; k++) x[k] = res[8 + k];
                    x[7] = bodies;
                } else {
                    for (int q = 1; q < 4; q++) px[8 * xrow(e, q) + 7] = bodies;
                }
                o[7] = bodies;
                return o[3] < m.self_margin;
            });
            WSYNC();
#if defined(PBRL_HOST_EMU) && defined(PBRL_EMU_TRACE)
            if (act) g_emu_cnt[5]++; g_emu_cnt[6] += pop_count(act); if (act >> m.nsp_cb) g_emu_cnt[7]++;
#endif
            commit(act);
            for (int q = 1; q < 4; q++) {  // the further slots, one round per slot
                const unsigned more = wballot(lane, 32, [&](int e) {
                    if (!((need >> e) & 1u) || q >= (e < m.nsp_cb ? 2 : 4)) return false;
                    const float* x = px + 8 * xrow(e, q);
                    if (!(x[3] < m.self_margin)) return false;
                    float* o = pr + 8 * e;
                    for (int k = 0; k < 8; k++) o[k] = x[k];
                    return true;
                });
                WSYNC();
                if (more != 0u) commit(more);
            }
        }
    }
    LANE_IF(z, 1) {
        w.nsc = total < KS ? total : KS;
        if (total > KS) w.status |= PBRL_STATUS_SELF_OVERFLOW;
    }
    WSYNC();
}

// constraint rows: enabled flags, D, aref  (mjx constraint.make_constraint; needs cvel of the current qvel)
PB_DEVN void stage_rows(const DevModel& m, Ws& w, const float* qpos, const float* qvel, int lane) {
    LANE_IF(l, 32) {
        // joint limit row l
        w.lim_sign[l] = 0.f;
        if (l < NLIM) {
            int j = m.lim_jnt[l];
            float q = qpos[m.jnt_qposadr[j]], dlo = q - m.jnt_lo[j], dhi = m.jnt_hi[j] - q;
            float pos = fminf(dlo, dhi) - m.jnt_margin[j];
            if (pos < 0.f) {
                float sg = dlo < dhi ? 1.f : -1.f, imp = impedance(m, pos);
                float Rr = fmaxf(m.dof_invweight0[m.jnt_dofadr[j]] * (1.f - imp) / imp, 1e-15f);
                w.lim_sign[l] = sg;
                w.lim_D[l] = 1.0f / Rr;
                w.lim_aref[l] = -m.sol_b * (sg * qvel[m.jnt_dofadr[j]]) - m.sol_k * imp * pos;
            }
        }
        // contact slot l
        w.con_on[l] = 0;
        if (l < NC) {
            float pos = w.con_dist[l] - m.con_margin[l];
            if (pos < 0.f) {
                float imp = impedance(m, pos), j6[6];
                float Rr = fmaxf(m.con_invweight[l] * (1.f - imp) / imp, 1e-15f);
                w.con_on[l] = 1;
                w.con_D[l] = 1.0f / Rr;
                for (int e = 0; e < 4; e++) {
                    contact_j6(w.con_r + SR * l, w.con_t1 + 2 * l, m.con_mu[l], e, j6);
                    w.con_aref[4 * l + e] = -m.sol_b * dot6(j6, w.cvel + S6 * m.con_body[l]) - m.sol_k * imp * pos;
                }
            }
        }
    }
    WSYNC();
    const unsigned cm = wballot(lane, NC, [&](int c) { return w.con_on[c] != 0; });
    LANE_IF(z, 1) w.con_mask = cm;
    const int nsc = w.nsc;
    if (nsc > 0) {  // lane l: pyramid row l & 3 of body-vs-body contact l >> 2
        LANE_IF(l, 32) {
            if (l < 4 * nsc) {
                const int c = l >> 2, bb = w.sc_bb[c];
                const float pos = w.sc_n[4 * c + 3] - m.self_margin, imp = impedance(m, pos);
                const float invw = m.self_invw_scale * (m.body_invw[bb & 0xff] + m.body_invw[bb >> 8]);
                float j6[6];
                self_j6(w, m.self_mu, c, l & 3, j6);
                w.sc_aref[l] = -m.sol_b * self_row_dot(w, j6, w.cvel, c) - m.sol_k * imp * pos;
                // ([3] of the point row: the other three lanes of the contact load the row but use only its xyz)
                if ((l & 3) == 0) w.sc_r[l + 3] = 1.0f / fmaxf(invw * (1.f - imp) / imp, 1e-15f);
            }
        }
    }
    WSYNC();
}

// residual-like products for all rows from w.V = tree_vel(x): out = J x (limits into lim_out, contacts into con_out)
PB_DEV void rows_Jx(const DevModel& m, Ws& w, const float* x, float* lim_out, float* con_out, float* self_out, int lane) {
    LANE_IF(l, 32) {
        if (l < 4 * w.nsc) {
            float j6[6];
            self_j6(w, m.self_mu, l >> 2, l & 3, j6);
            self_out[l] = self_row_dot(w, j6, w.V, l >> 2);
        }
        if (l < NLIM && w.lim_sign[l] != 0.f) lim_out[l] = w.lim_sign[l] * x[m.jnt_dofadr[m.lim_jnt[l]]];
        if (l < NC && w.con_on[l]) {
            float j6[6];
            for (int e = 0; e < 4; e++) {
                contact_j6(w.con_r + SR * l, w.con_t1 + 2 * l, m.con_mu[l], e, j6);
                con_out[4 * l + e] = dot6(j6, w.V + S6 * m.con_body[l]);
            }
        }
    }
    WSYNC();
}

// out[i] -= (J^T f)_i with f_r = -D_r jar_r on active rows (jar<0); i.e. adds the constraint gradient D*jar*J
// mode 0: out = base - J^T f   (gradient),  mode 1: out = J^T f  (qfrc_constraint)
// Pass 1 (lane = contact slot): the spatial force of the slot's active pyramid edges.  With w_e = n +- mu t the four
// edge forces collapse to one force vector, so a slot costs one cross product.  Pass 2 (lane = dof): gather the slots
// of the dof's subtree (contact slots are sorted by body, a subtree is a bit mask) and project on the dof.
PB_DEVN void rows_JTf(const DevModel& m, Ws& w, const float* base, float* out, int mode, int lane) {
    float* Fc = w.K;  // [NCP] x S6 (the Hessian blocks are rebuilt after every gradient)
    static_assert(NCP * S6 <= NB * S21, "contact force scratch");
    LANE_IF(c, NC) {
        if ((w.con_mask >> c) & 1u) {
            float jar[4], r[4], f6[6];
            ld4(w.con_jar + 4 * c, jar);
            ld4(w.con_r + SR * c, r);
            const float D = w.con_D[c], mu = m.con_mu[c], t1x = w.con_t1[2 * c], t1y = w.con_t1[2 * c + 1];
            const float f0 = jar[0] < 0.f ? -D * jar[0] : 0.f, f1 = jar[1] < 0.f ? -D * jar[1] : 0.f;
            const float f2 = jar[2] < 0.f ? -D * jar[2] : 0.f, f3 = jar[3] < 0.f ? -D * jar[3] : 0.f;
            const float a = mu * (f0 - f1), b = mu * (f2 - f3);
            f6[3] = a * t1x - b * t1y; f6[4] = a * t1y + b * t1x; f6[5] = (f0 + f1) + (f2 + f3);
            cross3(r, f6 + 3, f6);
            st6(Fc + S6 * c, f6);
        }
    }
    const int nsc = w.nsc;
    float* Fs = w.K + NCP * S6;  // [KS] x S6: spatial force of each body-vs-body contact on its body2 (body1 gets the opposite)
    LANE_IF(c, KS) {
        if (c < nsc) {
            float jar[4], n[4], r[4], t1[3], t2[3], f6[6];
            ld4(w.sc_jar + 4 * c, jar);
            ld4(w.sc_n + 4 * c, n);
            ld4(w.sc_r + 4 * c, r);
            frame_tangents(n, t1, t2);
            const float D = r[3], mu = m.self_mu;
            const float f0 = jar[0] < 0.f ? -D * jar[0] : 0.f, f1 = jar[1] < 0.f ? -D * jar[1] : 0.f;
            const float f2 = jar[2] < 0.f ? -D * jar[2] : 0.f, f3 = jar[3] < 0.f ? -D * jar[3] : 0.f;
            const float fn = (f0 + f1) + (f2 + f3), a = mu * (f0 - f1), b = mu * (f2 - f3);
            for (int k = 0; k < 3; k++) f6[3 + k] = fn * n[k] + a * t1[k] + b * t2[k];
            cross3(r, f6 + 3, f6);
            st6(Fs + S6 * c, f6);
        }
    }
    WSYNC();
    const int lim0 = m.jnt_dofadr[m.lim_jnt[0]];  // limited hinges own contiguous dofs (checked at model creation)
    LANE_IF(l, NDL) {
        const int i = l + 3;
        float f[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, cd[6];
        unsigned mk = w.con_mask & m.body_subconmask[m.dof_bodyid[i]];
        while (mk) {
            const int c = first_bit(mk);
            mk &= mk - 1u;
            float t[6];
            ld6(Fc + S6 * c, t);
            for (int k = 0; k < 6; k++) f[k] += t[k];
        }
        const unsigned sub = m.body_submask[m.dof_bodyid[i]];
#pragma unroll 1   // warp-uniform trip count, zero for most envs: a real loop (unrolled and predicated it costs 470 instructions per call)
        for (int c = 0; c < nsc; c++) {  // +force where body2 hangs below this dof, -force where body1 does (both: they cancel)
            const int bb = w.sc_bb[c];
            const float sg = (float)((int)((sub >> (bb >> 8)) & 1u) - (int)((sub >> (bb & 0xff)) & 1u));
            float t[6];
            ld6(Fs + S6 * c, t);
            for (int k = 0; k < 6; k++) f[k] = fmaf(sg, t[k], f[k]);
        }
        ld6(w.cdof + S6 * i, cd);
        float jtf = dot6(cd, f);
        if (m.jnt_limited[m.dof_jntid[i]]) {
            const int r = i - lim0;
            if (w.lim_sign[r] != 0.f && w.lim_jar[r] < 0.f) jtf += w.lim_sign[r] * (-w.lim_D[r] * w.lim_jar[r]);
        }
        out[i] = mode == 0 ? base[i] - jtf : jtf;
        if (i == 3) for (int k = 0; k < 3; k++) out[k] = mode == 0 ? base[k] - f[3 + k] : f[3 + k];
    }
    WSYNC();
}

// LD <- M + J^T D_active J  (tree-sparse, assembled like a composite inertia).
// Pass 1 (lane = slot): W = sum over active edges of D w w^T, a symmetric 3x3 that has a closed form in the pyramid
// basis.  Pass 2 (lane = body): the 6x6 block of each body from its slots, K = [r]x W [r]x^T etc.  Pass 3 (lane = dof):
// gather the blocks of the subtree (bit mask of bodies that carry active contacts), apply to the cdof, fill the row.
PB_DEVN void assemble_hessian(const DevModel& m, Ws& w, int lane) {
    float* Wc = w.buf6;  // [NCP] x S6: xx yy zz xy xz yz
    static_assert(NCP * S6 <= NBUF6, "contact W scratch");
    const unsigned act = wballot(lane, NC, [&](int c) {
        return ((w.con_mask >> c) & 1u) && (w.con_jar[4 * c] < 0.f || w.con_jar[4 * c + 1] < 0.f || w.con_jar[4 * c + 2] < 0.f || w.con_jar[4 * c + 3] < 0.f);
    });
    LANE_IF(c, NC) {
        if ((act >> c) & 1u) {
            float jar[4], W[6];
            ld4(w.con_jar + 4 * c, jar);
            const float D = w.con_D[c], mu = m.con_mu[c], cs = w.con_t1[2 * c], sn = w.con_t1[2 * c + 1];
            const float a0 = jar[0] < 0.f ? D : 0.f, a1 = jar[1] < 0.f ? D : 0.f, a2 = jar[2] < 0.f ? D : 0.f, a3 = jar[3] < 0.f ? D : 0.f;
            const float B1 = mu * (a0 - a1), B2 = mu * (a2 - a3), C1 = mu * mu * (a0 + a1), C2 = mu * mu * (a2 + a3);
            W[0] = C1 * cs * cs + C2 * sn * sn; W[1] = C1 * sn * sn + C2 * cs * cs; W[2] = (a0 + a1) + (a2 + a3);
            W[3] = (C1 - C2) * cs * sn; W[4] = B1 * cs - B2 * sn; W[5] = B1 * sn + B2 * cs;
            st6(Wc + S6 * c, W);
        }
    }
    WSYNC();
    const unsigned kb = wballot(lane, NB, [&](int b) { return (act & m.body_conmask[b]) != 0u; });
    LANE_IF(b, NB) {
        unsigned mk = act & m.body_conmask[b];
        if (mk) {
            float k[24];
            for (int q = 0; q < 24; q++) k[q] = 0.f;
            while (mk) {
                const int c = first_bit(mk);
                mk &= mk - 1u;
                float W[6], r[4];
                ld6(Wc + S6 * c, W);
                ld4(w.con_r + SR * c, r);
                const float c0[3] = {W[0], W[3], W[4]}, c1[3] = {W[3], W[1], W[5]}, c2[3] = {W[4], W[5], W[2]};
                float g0[3], g1[3], g2[3];  // columns of G = [r]x W
                cross3(r, c0, g0); cross3(r, c1, g1); cross3(r, c2, g2);
                const float r0[3] = {g0[0], g1[0], g2[0]}, r1[3] = {g0[1], g1[1], g2[1]}, r2[3] = {g0[2], g1[2], g2[2]};
                float h0[3], h1[3], h2[3];  // rows of [r]x W [r]x^T
                cross3(r, r0, h0); cross3(r, r1, h1); cross3(r, r2, h2);
                k[0] += h0[0]; k[1] += h0[1]; k[2] += h0[2]; k[6] += h1[1]; k[7] += h1[2]; k[11] += h2[2];
                k[3] += r0[0]; k[4] += r0[1]; k[5] += r0[2]; k[8] += r1[0]; k[9] += r1[1]; k[10] += r1[2];
                k[12] += r2[0]; k[13] += r2[1]; k[14] += r2[2];
                k[15] += W[0]; k[16] += W[3]; k[17] += W[4]; k[18] += W[1]; k[19] += W[5]; k[20] += W[2];
            }
            st24(w.K + S21 * b, k);
        }
    }
    WSYNC();
    const int lim0 = m.jnt_dofadr[m.lim_jnt[0]];
    LANE_IF(l, NDL) {
        const DofRow r = dof_row(m, l + 3);
        unsigned bm = kb & m.body_submask[m.dof_bodyid[r.i]];
        float diag = 0.f;
        if (m.jnt_limited[m.dof_jntid[r.i]]) {
            const int q = r.i - lim0;
            if (w.lim_sign[q] != 0.f && w.lim_jar[q] < 0.f) diag = w.lim_D[q];
        }
        if (bm == 0u) {
            w.LD[r.adr] = w.M[r.adr] + diag;
#pragma unroll
            for (int a = 1; a <= MAXANC; a++) if (a <= r.d) w.LD[r.adr + a] = w.M[r.adr + a];
            if (r.i == 3) for (int e = 0; e < 6; e++) w.LD[e] = w.M[e];  // rows 0..2 are the first six entries
        } else {
            float k[24], cd[6], kc[6];
            for (int q = 0; q < 24; q++) k[q] = 0.f;
            while (bm) {
                const int c = first_bit(bm);
                bm &= bm - 1u;
                float t[24];
                ld24(w.K + S21 * c, t);
                for (int q = 0; q < 21; q++) k[q] += t[q];
            }
            ld6(w.cdof + S6 * r.i, cd);
            for (int a = 0; a < 6; a++) {
                float s = 0.f;
                for (int c = 0; c < 6; c++) s = fmaf(k[sym21(a, c)], cd[c], s);
                kc[a] = s;
            }
            w.LD[r.adr] = w.M[r.adr] + dot6(cd, kc) + diag;
#pragma unroll
            for (int t = 0; t < MAXANC; t++) {
                const bool on = t < r.d;
                float cj[6];
                ld6(w.cdof + S6 * (on ? row_anc(r, t) : r.i), cj);
                const float v = w.M[on ? r.adr + r.d - t : r.adr] + dot6(cj, kc);
                if (on) w.LD[r.adr + r.d - t] = v;
            }
            if (r.i == 3)  // rows 0..2: cdof = [0; e_k]  ->  the lin-lin block
                for (int a = 0; a < 3; a++)
                    for (int j = 0; j <= a; j++) w.LD[m.dof_Madr[a] + (a - j)] = w.M[m.dof_Madr[a] + (a - j)] + k[sym21(3 + j, 3 + a)];
        }
    }
    WSYNC();
}

// ---------------------------------------------------------------------------------- dense Hessian (body-vs-body rows)
// x <- H^-1 x with H = (tree-sparse matrix in w.LD) + sum over the active body-vs-body rows r of D_r u_r u_r^T, where
// u_r[i] = +-cdof_i . j6_r for the dofs above body2 (+) / body1 (-) and 0 above both.
PB_DEV int tri(int i, int j) { return i * (i + 1) / 2 + j; }  // i >= j
// the four pyramid rows of contact s over the dofs: U[36 e + i]
PB_DEV void self_rows_dense(const DevModel& m, Ws& w, int s, float* U, int lane) {
    const int bb = w.sc_bb[s];
    float n[4], r[4], t1[3], t2[3], jn[6], j1[6], j2[6];
    ld4(w.sc_n + 4 * s, n);
    ld4(w.sc_r + 4 * s, r);
    frame_tangents(n, t1, t2);
    cross3(r, n, jn); cross3(r, t1, j1); cross3(r, t2, j2);
    for (int k = 0; k < 3; k++) { jn[3 + k] = n[k]; j1[3 + k] = t1[k]; j2[3 + k] = t2[k]; }
    LANE_LOOP(i, NV) {
        const unsigned sub = m.body_submask[m.dof_bodyid[i]];
        const float sg = (float)((int)((sub >> (bb >> 8)) & 1u) - (int)((sub >> (bb & 0xff)) & 1u));
        float cd[6];
        ld6(w.cdof + S6 * i, cd);
        const float an = sg * dot6(cd, jn), a1 = sg * m.self_mu * dot6(cd, j1), a2 = sg * m.self_mu * dot6(cd, j2);
        U[i] = an + a1; U[36 + i] = an - a1; U[72 + i] = an + a2; U[108 + i] = an - a2;
    }
    LANE_IF(z, 2) for (int e = 0; e < 4; e++) U[36 * e + NV + z] = 0.f;  // pad columns 34, 35 (vector loads read them)
    WSYNC();
}
#if defined(__CUDACC__) && !defined(PBRL_HOST_EMU)
// Register-resident L^T D L (MuJoCo's elimination order, last dof first): lane i holds row i < 32 of the lower triangle in
// registers, rows 32 and 33 stay in shared memory (they are eliminated first and never touched again).  Each step
// broadcasts the pivot row through a 32-float staging row (double buffered); the right-hand side rides along (lane = dof),
// and the back substitution pushes x_i to the rows below with one shuffle per dof.  (A lane = column variant that
// broadcasts the pivot row with one shuffle per entry instead of the staging row measured 30 % slower: 10.2k vs 7.7k
// cycles per call for a lone warp -- shuffles serialise on their scoreboard.)
namespace dense {
constexpr unsigned FULL = 0xffffffffu;
template <int K>
struct Step {
    PB_MEMBER void run(float (&R)[32], float& xl, float& dinv_own, bool& bad, float* stage, int lane) {
        float* st = stage + 32 * (K & 1);
        if (lane == K) {
#pragma unroll
            for (int q = 0; q < (K + 4) / 4; q++) reinterpret_cast<float4*>(st)[q] = make_float4(R[4 * q], R[4 * q + 1], R[4 * q + 2], R[4 * q + 3]);
        }
        const float xk = __shfl_sync(FULL, xl, K);
        __syncwarp();
        float c[32];
#pragma unroll
        for (int q = 0; q < (K + 4) / 4; q++) {
            const float4 v = reinterpret_cast<const float4*>(st)[q];
            c[4 * q] = v.x; c[4 * q + 1] = v.y; c[4 * q + 2] = v.z; c[4 * q + 3] = v.w;
        }
        const float d = c[K], inv = fast_rcp(d);
        bad |= !(d > 0.f);
        if (lane == K) dinv_own = inv;
        if constexpr (K > 0) {
            const float li = lane < K ? st[lane] * inv : 0.f;  // L[K, lane]
#pragma unroll
            for (int j = 0; j < K; j++) R[j] = fmaf(-li, c[j], R[j]);  // entries right of the diagonal of a row are never read
            xl = fmaf(-li, xk, xl);
            Step<K - 1>::run(R, xl, dinv_own, bad, stage, lane);
        }
    }
};
template <int I>
struct Back {  // x_I is final: every row below receives L[k, I] x_I
    PB_MEMBER void run(const float (&R)[32], float& xl, float dinv_own, int lane) {
        if constexpr (I < 31) {
            const float xi = __shfl_sync(FULL, xl, I);
            if (lane > I) xl = fmaf(-R[I] * dinv_own, xi, xl);
            Back<I + 1>::run(R, xl, dinv_own, lane);
        }
    }
};
}  // namespace dense

PB_DEVN void dense_factor_solve(const DevModel& m, Ws& w, unsigned sact, float* x, int lane) {
    static_assert(NV == 34, "rows 32 and 33 are handled outside the register tile");
    float* H = w.LD;  // packed lower triangle across LD + K (scatter target; rows 32, 33 stay here)
    {
        float ent[(NM + 31) / 32];
#pragma unroll
        for (int k = 0; k < (NM + 31) / 32; k++) { const int e = lane + 32 * k; ent[k] = e < NM ? w.LD[e] : 0.f; }
        __syncwarp();
        for (int i = lane; i < NHD; i += 32) H[i] = 0.f;
        __syncwarp();
#pragma unroll
        for (int k = 0; k < (NM + 31) / 32; k++) { const int e = lane + 32 * k; if (e < NM) H[tri(m.M_row[e], m.M_col[e])] = ent[k]; }
        __syncwarp();
    }
    float R[32];
    {
        const float* row = H + tri(lane, 0);
#pragma unroll
        for (int j = 0; j < 32; j++) R[j] = j <= lane ? row[j] : 0.f;
    }
    float* H32 = H + tri(32, 0);  // 33 entries
    float* H33 = H + tri(33, 0);  // 34 entries
    float* U = w.buf6;            // [4] x 36: the four pyramid rows of one contact; the staging rows of the factorisation later
    const int nsc = w.nsc;
#pragma unroll 1
    for (int s = 0; s < nsc; s++) {
        const unsigned a4 = (sact >> (4 * s)) & 15u;
        if (a4 == 0u) continue;
        __syncwarp();
        self_rows_dense(m, w, s, U, lane);
        const float D = w.sc_r[4 * s + 3];
        float ui[4], u32[4], u33[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const float a = ((a4 >> e) & 1u) ? D : 0.f;
            ui[e] = a * U[36 * e + lane]; u32[e] = a * U[36 * e + 32]; u33[e] = a * U[36 * e + 33];
        }
#pragma unroll
        for (int q = 0; q < 8; q++) {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const float4 v = reinterpret_cast<const float4*>(U + 36 * e)[q];
                acc[0] = fmaf(ui[e], v.x, acc[0]); acc[1] = fmaf(ui[e], v.y, acc[1]); acc[2] = fmaf(ui[e], v.z, acc[2]); acc[3] = fmaf(ui[e], v.w, acc[3]);
            }
#pragma unroll
            for (int t = 0; t < 4; t++) R[4 * q + t] += acc[t];
        }
        for (int j = lane; j < NV; j += 32) {  // rows 32 / 33 in shared memory: column j
            float a32 = 0.f, a33 = 0.f;
#pragma unroll
            for (int e = 0; e < 4; e++) { const float uj = U[36 * e + j]; a32 = fmaf(u32[e], uj, a32); a33 = fmaf(u33[e], uj, a33); }
            if (j <= 32) H32[j] += a32;
            H33[j] += a33;
        }
    }
    __syncwarp();
    // eliminate dof 33, then dof 32 (rows in shared memory), then the register rows 31 .. 0
    bool bad = false;
    float xl = x[lane];
    float x32 = x[32];
    const float x33 = x[33];
    const float d33 = H33[33], inv33 = fast_rcp(d33);
    bad |= !(d33 > 0.f);
    {
        const float li = H33[lane] * inv33;
#pragma unroll
        for (int j = 0; j < 32; j++) R[j] = fmaf(-li, H33[j], R[j]);
        xl = fmaf(-li, x33, xl);
        const float l32 = H33[32] * inv33;
        x32 = fmaf(-l32, x33, x32);
        __syncwarp();
        for (int j = lane; j <= 32; j += 32) H32[j] = fmaf(-l32, H33[j], H32[j]);
        __syncwarp();
    }
    const float d32 = H32[32], inv32 = fast_rcp(d32);
    bad |= !(d32 > 0.f);
    {
        const float li = H32[lane] * inv32;
#pragma unroll
        for (int j = 0; j < 32; j++) R[j] = fmaf(-li, H32[j], R[j]);
        xl = fmaf(-li, x32, xl);
    }
    float dinv_own = 0.f;
    __syncwarp();
    dense::Step<31>::run(R, xl, dinv_own, bad, U, lane);
    // x <- D^-1 x, then L x = z from dof 0 upwards
    xl *= dinv_own;
    dense::Back<0>::run(R, xl, dinv_own, lane);
    float s32 = H32[lane] * xl, s33 = H33[lane] * xl;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s32 += __shfl_xor_sync(0xffffffffu, s32, o); s33 += __shfl_xor_sync(0xffffffffu, s33, o); }
    const float y32 = (x32 - s32) * inv32;                      // x32 holds L^-T b; z = D^-1 y; minus the rows above
    const float y33 = (x33 - s33 - H33[32] * y32) * inv33;
    __syncwarp();
    x[lane] = xl;
    if (lane == 0) { x[32] = y32; x[33] = y33; if (bad) w.status |= PBRL_STATUS_PIVOT; }
    __syncwarp();
}
#else
// host emulation (tests only): packed lower triangle, L D L^T column by column with the forward substitution riding along
PB_DEVN void dense_factor_solve(const DevModel& m, Ws& w, unsigned sact, float* x, int lane) {
    float* H = w.LD;
    {
        float ent[(NM + 31) / 32];
#pragma unroll
        for (int k = 0; k < (NM + 31) / 32; k++) ent[k] = 0.f;
#if defined(__CUDACC__) && !defined(PBRL_HOST_EMU)
#pragma unroll
        for (int k = 0; k < (NM + 31) / 32; k++) { const int e = lane + 32 * k; if (e < NM) ent[k] = w.LD[e]; }
        WSYNC();
        LANE_LOOP(i, NHD) H[i] = 0.f;
        WSYNC();
#pragma unroll
        for (int k = 0; k < (NM + 31) / 32; k++) { const int e = lane + 32 * k; if (e < NM) H[tri(m.M_row[e], m.M_col[e])] = ent[k]; }
#else
        float tmp[NM];
        for (int e = 0; e < NM; e++) tmp[e] = w.LD[e];
        for (int i = 0; i < NHD; i++) H[i] = 0.f;
        for (int e = 0; e < NM; e++) H[tri(m.M_row[e], m.M_col[e])] = tmp[e];
        (void)ent;
#endif
        WSYNC();
    }
    float* U = w.buf6;  // [4] x 36: the four pyramid rows of one contact
    const int nsc = w.nsc;
    for (int s = 0; s < nsc; s++) {
        const unsigned a4 = (sact >> (4 * s)) & 15u;
        if (a4 == 0u) continue;
        self_rows_dense(m, w, s, U, lane);
        const float D = w.sc_r[4 * s + 3];
        LANE_LOOP(i, NV) {
            float ui[4];
            for (int e = 0; e < 4; e++) ui[e] = ((a4 >> e) & 1u) ? D * U[36 * e + i] : 0.f;
            if (ui[0] != 0.f || ui[1] != 0.f || ui[2] != 0.f || ui[3] != 0.f) {
                float* row = H + tri(i, 0);
                for (int j = 0; j <= i; j++)
                    row[j] += ui[0] * U[j] + ui[1] * U[36 + j] + ui[2] * U[72 + j] + ui[3] * U[108 + j];
            }
        }
        WSYNC();
    }
    bool bad = false;
    for (int k = 0; k < NV; k++) {
        const float d = H[tri(k, k)], inv = 1.0f / d, xk = x[k];
        bad |= !(d > 0.f);
        LANE_LOOP(i, NV) {
            if (i > k) {
                float* row = H + tri(i, 0);
                const float lik = row[k] * inv;  // column k keeps L * D; only the columns right of it are touched
                for (int j = k + 1; j <= i; j++) row[j] = fmaf(-lik, H[tri(j, k)], row[j]);
                x[i] = fmaf(-lik, xk, x[i]);
            }
        }
        LANE_IF(z, 1) w.Dinv[k] = inv;
        WSYNC();
    }
    LANE_LOOP(i, NV) x[i] *= w.Dinv[i];
    WSYNC();
    for (int k = NV - 1; k > 0; k--) {
        const float xk = x[k];
        const float* row = H + tri(k, 0);
        LANE_LOOP(i, NV) if (i < k) x[i] = fmaf(-row[i] * w.Dinv[i], xk, x[i]);
        WSYNC();
    }
    if (bad) { LANE_IF(z, 1) w.status |= PBRL_STATUS_PIVOT; }
    WSYNC();
}

#endif

// ---------------------------------------------------------------------------------- Newton solver on qacc
// minimise 1/2 (a-a_s)^T M (a-a_s) + sum_r 1/2 D_r min(0, J_r a - aref_r)^2   
// MJX_RULE: the stop rule is a template parameter, so that the default (exact active set) carries none of the tolerance rule's work
template <bool MJX_RULE>
PB_DEVN void solve_constraints(const DevModel& m, Ws& w, int lane) {
    const unsigned lim_rows = wballot(lane, NLIM, [&](int l) { return w.lim_sign[l] != 0.f; });
    w.niter = 0;
    w.work = 0;
    const int nsr = 4 * w.nsc;  // body-vs-body pyramid rows (lane l: row l)
    const bool constrained = (lim_rows | w.con_mask) != 0u || nsr > 0;
    const bool coop = (w.phase_lock & 16) != 0;  // iterate in step with the other warps of the CTA (shared I-cache window)
    if (!constrained) {  // unconstrained: qacc = qacc_smooth = M^-1 qfrc_smooth
        LANE_LOOP(e, NM) w.LD[e] = w.M[e];
        WSYNC();
        ldl_factor_solve(m, w, w.asmooth, lane);
        PROF(10);
        LANE_LOOP(i, NV) { w.qacc[i] = w.asmooth[i]; w.fcon[i] = 0.f; }
        WSYNC();
        PROF(11);
        if (!coop) return;
    } else {
        // with active rows the optimum does not need qacc_smooth: the cost is written with qfrc_smooth
        // start from the warm start
        LANE_LOOP(i, NV) w.qacc[i] = w.warm[i];
        WSYNC();
        tree_Mx(m, w, w.qacc, w.Ma, lane);
        rows_Jx(m, w, w.qacc, w.lim_jar, w.con_jar, w.sc_jar, lane);
        LANE_IF(l, 32) {
            if (l < nsr) w.sc_jar[l] -= w.sc_aref[l];
            if (l < NLIM && w.lim_sign[l] != 0.f) w.lim_jar[l] -= w.lim_aref[l];
            if (l < NC && w.con_on[l]) for (int e = 0; e < 4; e++) w.con_jar[4 * l + e] -= w.con_aref[4 * l + e];
            int bits = 0;
            if (l < NLIM && w.lim_sign[l] != 0.f) bits |= (w.lim_jar[l] < 0.f) ? 1 : 0;
            if (l < NC && w.con_on[l]) for (int e = 0; e < 4; e++) bits |= (w.con_jar[4 * l + e] < 0.f) ? (2 << e) : 0;
            if (l < nsr) bits |= (w.sc_jar[l] < 0.f) ? 64 : 0;
            w.act_bits[l] = bits;
        }
        WSYNC();
        PROF(12);
    }
    int it = 0, ndense = 0;
    bool running = constrained;
    // stop rule (include/pbrl.h: solver_rule): exact active set (default) or MuJoCo / MJX's tolerance-scaled improvement / gradient
    constexpr bool mjx_rule = MJX_RULE;
    const int maxit = mjx_rule ? m.solver_iterations : (m.solver_iterations < NEWTON_MAXIT ? m.solver_iterations : NEWTON_MAXIT);
    const int ls_maxit = mjx_rule ? m.ls_iterations : LS_MAXIT;
    for (;;) {
        const bool go = running && it < maxit;
        PROF_COUNT(60);
        if (go) PROF_COUNT(61);
        if (coop) { if (!CTA_OR(go ? 1 : 0)) break; }
        else if (!go) break;
        if (!go) continue;
        // gradient and Hessian for the current active set
        LANE_LOOP(i, NV) w.Mp[i] = w.Ma[i] - w.fsmooth[i];
        WSYNC();
        rows_JTf(m, w, w.Mp, w.grad, 0, lane);
        PROF(13);
        if (mjx_rule) {  // mjx solver.solve cond: gradient / (meaninertia * max(1, nv)) < tolerance -> converged (also before the first iteration)
            const float g2 = wsum(lane, NV, [&](int i) { return w.grad[i] * w.grad[i]; });
            if (sqrtf(g2) < m.solver_tol) { running = false; continue; }
        }
        assemble_hessian(m, w, lane);
        PROF(14);
        LANE_LOOP(i, NV) w.p[i] = -w.grad[i];
        WSYNC();
        // active body-vs-body rows couple two branches of the tree: their Hessian is not tree-sparse -> dense factorisation
        const unsigned sact = wballot(lane, 32, [&](int l) { return l < nsr && w.sc_jar[l] < 0.f; });
        if (sact != 0u) { dense_factor_solve(m, w, sact, w.p, lane); ndense++; PROF_COUNT(58); }
        else ldl_factor_solve(m, w, w.p, lane);
        PROF(15);
        tree_Mx(m, w, w.p, w.Mp, lane);
        rows_Jx(m, w, w.p, w.lim_jp, w.con_jp, w.sc_jp, lane);
        PROF(17);
        // exact line search on the piecewise-quadratic cost along p (safeguarded Newton on phi')
        float g0, pMp;
        wsum2(lane, NV, [&](int i, float& a, float& b) { const float pi = w.p[i]; a = fmaf(w.Ma[i] - w.fsmooth[i], pi, a); b = fmaf(w.Mp[i], pi, b); }, g0, pMp);
        float alpha = 1.f, lo = 0.f, hi = -1.f;
        float d1_0 = 0.f, gtol = 0.f;
        if (mjx_rule) gtol = m.ls_gtol * sqrtf(wsum(lane, NV, [&](int i) { return w.p[i] * w.p[i]; }));
        for (int ls = 0; ls < ls_maxit; ls++) {
            float d1, d2;  // first and second derivative of the row costs at alpha
            wsum2(lane, 32, [&](int l, float& a, float& b) {
                if (l < nsr) {
                    const float jp = w.sc_jp[l], x = w.sc_jar[l] + alpha * jp;
                    if (x < 0.f) { const float dj = w.sc_r[(l & ~3) + 3] * jp; a = fmaf(dj, x, a); b = fmaf(dj, jp, b); }
                }
                if (l < NLIM && w.lim_sign[l] != 0.f) {
                    const float jp = w.lim_jp[l], x = w.lim_jar[l] + alpha * jp;
                    if (x < 0.f) { const float dj = w.lim_D[l] * jp; a = fmaf(dj, x, a); b = fmaf(dj, jp, b); }
                }
                if (l < NC && w.con_on[l]) {
                    float jar[4], jp[4];
                    ld4(w.con_jar + 4 * l, jar);
                    ld4(w.con_jp + 4 * l, jp);
                    const float D = w.con_D[l];
                    for (int e = 0; e < 4; e++) {
                        const float x = jar[e] + alpha * jp[e];
                        if (x < 0.f) { const float dj = D * jp[e]; a = fmaf(dj, x, a); b = fmaf(dj, jp[e], b); }
                    }
                }
            }, d1, d2);
            float dphi = g0 + alpha * pMp + d1, ddphi = pMp + d2;
            if (ls == 0) d1_0 = fabsf(g0) + fabsf(pMp) + fabsf(d1);
            if (mjx_rule ? fabsf(dphi) < gtol : fabsf(dphi) <= 1e-6f * d1_0) break;
            if (dphi < 0.f) lo = alpha; else hi = alpha;
            float an = alpha - dphi / ddphi;
            if (!(an > lo) || (hi > 0.f && !(an < hi))) an = hi > 0.f ? 0.5f * (lo + hi) : 2.f * alpha;
            if (an == alpha) break;
            alpha = an;
        }
        PROF(18);
        // step, refresh residuals, detect active-set change
        LANE_LOOP(i, NV) { w.qacc[i] += alpha * w.p[i]; w.Ma[i] += alpha * w.Mp[i]; }
        float drow = 0.f;  // change of the row costs 1/2 D min(0, jar)^2 along the step (rule 1: cost improvement)
        auto row_cost = [](float D, float x0, float x1) { const float a = fminf(x0, 0.f), b = fminf(x1, 0.f); return 0.5f * D * (b * b - a * a); };
        const unsigned changed = wballot(lane, 32, [&](int l) {
            int bits = 0;
            if (l < nsr) {
                const float x0 = w.sc_jar[l], x1 = x0 + alpha * w.sc_jp[l];
                w.sc_jar[l] = x1;
                if (mjx_rule) drow += row_cost(w.sc_r[(l & ~3) + 3], x0, x1);
                bits |= (x1 < 0.f) ? 64 : 0;
            }
            if (l < NLIM && w.lim_sign[l] != 0.f) {
                const float x0 = w.lim_jar[l], x1 = x0 + alpha * w.lim_jp[l];
                w.lim_jar[l] = x1;
                if (mjx_rule) drow += row_cost(w.lim_D[l], x0, x1);
                bits |= (x1 < 0.f) ? 1 : 0;
            }
            if (l < NC && w.con_on[l])
                for (int e = 0; e < 4; e++) {
                    const float x0 = w.con_jar[4 * l + e], x1 = x0 + alpha * w.con_jp[4 * l + e];
                    w.con_jar[4 * l + e] = x1;
                    if (mjx_rule) drow += row_cost(w.con_D[l], x0, x1);
                    bits |= (x1 < 0.f) ? (2 << e) : 0;
                }
            // act_bits[l] still holds the active set this iteration's Hessian was built from
            const bool ch = bits != w.act_bits[l];
            w.act_bits[l] = bits;
            return ch;
        });
        WSYNC();
        it++;
        if (mjx_rule) {  // improvement = cost(a) - cost(a + alpha p) = -(alpha g0 + alpha^2 pMp / 2 + change of the row costs)
            const float improvement = -(alpha * g0 + 0.5f * alpha * alpha * pMp + wreduce(drow));
            if (improvement < m.solver_tol) running = false;
        } else if (changed == 0u) running = false;  // unchanged active set: the Newton step minimised the exact quadratic, i.e. the optimum is reached
        PROF(19);
    }
    if (!constrained) return;
    if (it >= maxit && running) w.status |= PBRL_STATUS_SOLVER_MAXITER;
    w.niter = it;
    w.work = it + PBRL_DENSE_WEIGHT * ndense;  // work estimate: an iteration with the dense Hessian counts more
    rows_JTf(m, w, w.Mp, w.fcon, 1, lane);
    PROF(20);
}

// ---------------------------------------------------------------------------------- one mjx.forward + Euler
// forward: everything up to qacc / qfrc_constraint at (w.qpos, w.qvel, w.ctrl, w.warm)
// Phase lock: bit k of w.phase_lock places a CTA barrier at stage-group boundary k of every substep
// (0 substep start, 1 after the position stage, 2 after the constraint rows, 3 after the solver, 5 after the collision stage;
// bit 4 is the cooperative Newton vote of solve_constraints).  The barriers keep
// the CTA's warps inside one instruction-cache window; every warp of a CTA executes the same number of them.
PB_DEVN void forward(const DevModel& m, Ws& w, int lane) {
    const int lock = w.phase_lock;
    if (lock & 1) CTA_SYNC();
    PROF(27);
    stage_fk(m, w, w.qpos, lane);
    PROF(2);
    stage_com(m, w, lane);
    PROF(3);
    stage_crb(m, w, lane);
    PROF(4);
    if (lock & 2) CTA_SYNC();
    PROF(28);
    stage_collide(m, w, lane);
    stage_collide_self(m, w, lane);
    PROF(5);
    if (lock & 32) CTA_SYNC();  // a warp that ran the convex narrow phase rejoins the others before the velocity stages
    PROF(31);
    stage_comvel(m, w, w.qvel, lane);
    PROF(6);
    stage_bias(m, w, lane);
    PROF(7);
    stage_rows(m, w, w.qpos, w.qvel, lane);
    PROF(8);
    if (lock & 4) CTA_SYNC();
    PROF(29);
    LANE_LOOP(i, NV) {
        float f = -m.dof_damping[i] * w.qvel[i] - w.bias[i];
        int a = m.dof_act[i];
        if (a >= 0) f += m.dof_gear[i] * fminf(fmaxf(w.ctrl[a], m.dof_ctrl_lo[i]), m.dof_ctrl_hi[i]);
        w.fsmooth[i] = f;
        w.asmooth[i] = f;
    }
    PROF(9);
    if (m.solver_rule == PBRL_SOLVER_MJX) solve_constraints<true>(m, w, lane);
    else solve_constraints<false>(m, w, lane);
}

// mjx.step: forward, warmstart <- qacc, Euler with implicit joint damping, integrate
PB_DEVN void substep(const DevModel& m, Ws& w, float& time, int lane) {
    forward(m, w, lane);
    PROF_COUNT(40 + (w.niter < 15 ? w.niter : 15));
    if (w.phase_lock & 8) CTA_SYNC();
    PROF(30);
    float h = m.timestep;
    LANE_LOOP(i, NV) { w.warm[i] = w.qacc[i]; w.p[i] = w.fsmooth[i] + w.fcon[i]; }
    LANE_LOOP(e, NM) w.LD[e] = w.M[e] + ((m.M_row[e] == m.M_col[e]) ? h * m.dof_damping[m.M_row[e]] : 0.f);
    WSYNC();
    ldl_factor_solve(m, w, w.p, lane);
    PROF(21);
    LANE_LOOP(i, NV) w.qvel[i] += h * w.p[i];
    WSYNC();
    LANE_IF(j, NJ) {
        int qa = m.jnt_qposadr[j], da = m.jnt_dofadr[j];
        if (m.jnt_type[j] == 0) {
            for (int k = 0; k < 3; k++) w.qpos[qa + k] += h * w.qvel[da + k];
            float wx = w.qvel[da + 3], wy = w.qvel[da + 4], wz = w.qvel[da + 5];
            float nrm = sqrtf(wx * wx + wy * wy + wz * wz), ax[3] = {1.f, 0.f, 0.f}, sn, cs, ql[4], qn[4];
            if (nrm > 1e-30f) { ax[0] = wx / nrm; ax[1] = wy / nrm; ax[2] = wz / nrm; }
            sincosf(0.5f * h * nrm, &sn, &cs);
            ql[0] = cs; ql[1] = ax[0] * sn; ql[2] = ax[1] * sn; ql[3] = ax[2] * sn;
            quat_mul(w.qpos + qa + 3, ql, qn);
            quat_normalize(qn);
            for (int k = 0; k < 4; k++) w.qpos[qa + 3 + k] = qn[k];
        } else {
            w.qpos[qa] += h * w.qvel[da];
        }
    }
    WSYNC();
    time += h;
    PROF(23);
}

// ---------------------------------------------------------------------------------- PD laws
// /root/reference/agents_env/pds_controllers_agents.py:17-34 (standard), :42-78 (feedback), :83-172 (stable).
// M and bias are evaluated at (qpos_lag, qvel_lag): the state of the last forward.
// lag_q / lag_v: the state of the last forward, staged in the workspace by the caller.
PB_DEVN void pd_torque(const DevModel& m, Ws& w, int pd_mode, int action_cols, const float* action, const float* kp,
                       const float* kd, float time, float pd_dt, const float* lag_q, const float* lag_v, int lane) {
    if (pd_mode == PBRL_PD_TORQUE) {
        LANE_IF(a, NU) w.tau[a] = action[a * action_cols];
        WSYNC();
        return;
    }
    // target angles -> w.grad[0..NU)
    LANE_IF(a, NU) {
        if (action_cols == 4) {
            float t = pd_mode == PBRL_PD_STABLE ? time + pd_dt : time;
            const float* c = action + 4 * a;
            w.grad[a] = c[0] + c[1] * t + c[2] * (t * t) + c[3] * (t * t * t);
        } else {
            w.grad[a] = action[a];
        }
    }
    WSYNC();
    const int qo = NQ - NU, vo = NV - NU;
    if (pd_mode == PBRL_PD_STANDARD) {
        LANE_IF(a, NU) w.tau[a] = -kp[a] * (w.qpos[qo + a] - w.grad[a]) + (-kd[a] * w.qvel[vo + a]);
        WSYNC();
        return;
    }
    smooth_position(m, w, lag_q, lane);
    smooth_velocity(m, w, lag_v, lane);
    if (pd_mode == PBRL_PD_FEEDBACK) {
        LANE_LOOP(i, NV) w.p[i] = i < vo ? 0.f : -kp[i - vo] * (w.qpos[qo + i - vo] - w.grad[i - vo]) + (-kd[i - vo] * w.qvel[i]);
        WSYNC();
        tree_Mx(m, w, w.p, w.Mp, lane);
        LANE_IF(a, NU) w.tau[a] = w.Mp[vo + a] + w.bias[vo + a];
        WSYNC();
        return;
    }
    // stable PD: qdd = (M + KD dt)^-1 (-C - KP e - KD qd);  tau = -kp e - kd (qd + qdd dt)
    LANE_LOOP(i, NV) {
        float e = 0.f, kpi = 0.f, kdi = 0.f;
        if (i >= vo) { e = (w.qpos[qo + i - vo] + w.qvel[i] * pd_dt) - w.grad[i - vo]; kpi = kp[i - vo]; kdi = kd[i - vo]; }
        w.Mp[i] = e;
        w.p[i] = -w.bias[i] - kpi * e - kdi * w.qvel[i];
    }
    LANE_LOOP(e, NM) {
        int r = m.M_row[e];
        w.LD[e] = w.M[e] + ((r == m.M_col[e] && r >= vo) ? kd[r - vo] * pd_dt : 0.f);
    }
    WSYNC();
    ldl_factor_solve(m, w, w.p, lane);
    LANE_IF(a, NU) w.tau[a] = -kp[a] * w.Mp[vo + a] - kd[a] * (w.qvel[vo + a] + w.p[vo + a] * pd_dt);
    WSYNC();
}

}  // namespace pbrl

// =====================================================================================================
// Env-level bodies: one call = one env handled by one warp.  Global-memory rows are read and written
// with lane-strided (coalesced) accesses; everything in between stays in the warp's workspace.
// =====================================================================================================
namespace pbrl {

#if defined(PBRL_HOST_EMU) && defined(PBRL_EMU_TRACE)
static int* g_emu_trace = nullptr;  // dev study only: (niter, dense iterations) per (env, substep)
#endif
struct StepArgs {
    int N, n_frames, pd_mode, action_cols, reward_mode;
    int f_begin, f_end;  // the substeps of the env step this launch runs: [0, n_frames) is the whole step (see env_step)
    int fused_parts;     // > 1: one launch runs the step in that many parts (equal shares of the substeps), regrouping in between
    int* fused_queue;    // [fused_parts - 1][FUSED_CLASSES][N] env ids + 1 by work class (zeroed by the host before the launch)
    int phase_lock;  // bit mask of CTA barriers per substep (see forward); every warp of a CTA makes the same number of passes
    int tail_split;  // k_step: split the groups of a batch smaller than the GPU over the idle SMs
    int cost_mode;   // work estimate written to buf.cost: 0 most expensive substep, 1 sum over the substeps, 2 last substep
    const float *action, *kp, *kd, *pd_time;
    float pd_dt;
    unsigned ctrl_mask;
    PbrlBuffers buf;
    const float* frames;  // device motion table (mimic mode)
    DevMimic mm;
};

PB_DEV void load_row(float* dst, const float* src, int n, int lane) { LANE_LOOP(i, n) dst[i] = src[i]; }
// loads that bypass L1: rows another CTA wrote earlier in this launch (the fused two-part step).  A 140-byte row shares its
// 128-byte lines with its neighbours, and this SM may hold such a line from before the row was written.
#if defined(__CUDACC__) && !defined(PBRL_HOST_EMU)
PB_DEV float ld_fresh(const float* p) { return __ldcg(p); }
PB_DEV int ld_fresh(const int* p) { return __ldcg(p); }
#else
PB_DEV float ld_fresh(const float* p) { return *p; }
PB_DEV int ld_fresh(const int* p) { return *p; }
#endif
PB_DEV void load_row_fresh(float* dst, const float* src, int n, int lane) { LANE_LOOP(i, n) dst[i] = ld_fresh(src + i); }
PB_DEV void store_row(float* dst, const float* src, int n, int lane) { LANE_LOOP(i, n) dst[i] = src[i]; }

PB_DEV void write_obs_reference(const StepArgs& a, Ws& w, int env, float done_value, int lane) {
    const PbrlBuffers& B = a.buf;
    if (B.obs) { store_row(B.obs + (size_t)env * PBRL_OBS, w.qpos, NQ, lane); store_row(B.obs + (size_t)env * PBRL_OBS + NQ, w.qvel, NV, lane); }
    LANE_IF(k, 3) if (B.reward) B.reward[(size_t)env * 3 + k] = 0.f;      // agent_mimic.py:84
    LANE_IF(z, 1) if (B.done) B.done[env] = done_value;                  // agent_mimic.py:72
}

PB_DEV void frames_prefetch(const float* frames, const DevMimic& mm, Ws& w, int clip, float t, int lane);
PB_DEVN void mimic_epilogue(const DevModel& m, const DevMimic& mm, const float* frames, Ws& w, const PbrlBuffers& B,
                            int env, int clip, float t, float* acc, int lane);

// HumanoidTrain.step (agent_mimic.py:58-91): PD law once, n_frames substeps with the torque held.
// A launch runs the substeps [f_begin, f_end) of the env step: the PD law belongs to the part that starts at substep 0, the
// observation / reward epilogue to the part that ends at n_frames; in between the state lives in the caller's buffers
// (qpos, qvel, warm, ctrl, time, status, cost).  One launch with [0, n_frames) is the whole step.
// `fresh`: the state rows were written earlier in this launch by another CTA (fused two-part step): load them past L1.
PB_DEVN void env_step(const DevModel& m, Ws& w, const StepArgs& a, int env, float* acc, int lane, int f_begin, int f_end, bool fresh) {
    const PbrlBuffers& B = a.buf;
    const bool mimic = a.reward_mode == PBRL_REWARD_MIMIC;
    const bool first = f_begin == 0, last = f_end == a.n_frames;
    const int clip = mimic ? B.motion_id[env] : 0;
    PROF_START();
    if (fresh) {
        load_row_fresh(w.qpos, B.qpos + (size_t)env * NQ, NQ, lane);
        load_row_fresh(w.qvel, B.qvel + (size_t)env * NV, NV, lane);
        load_row_fresh(w.warm, B.warm + (size_t)env * NV, NV, lane);
        load_row_fresh(w.ctrl, B.ctrl + (size_t)env * NU, NU, lane);
    } else {
        load_row(w.qpos, B.qpos + (size_t)env * NQ, NQ, lane);
        load_row(w.qvel, B.qvel + (size_t)env * NV, NV, lane);
        load_row(w.warm, B.warm + (size_t)env * NV, NV, lane);
        load_row(w.ctrl, B.ctrl + (size_t)env * NU, NU, lane);
    }
    float* lag_q = w.fsmooth;  // dead until the first substep
    float* lag_v = w.qacc;
    if (first && a.pd_mode >= PBRL_PD_FEEDBACK) {
        load_row(lag_q, B.qpos_lag + (size_t)env * NQ, NQ, lane);
        load_row(lag_v, B.qvel_lag + (size_t)env * NV, NV, lane);
    }
    LANE_IF(z, 1) w.status = first ? 0 : (fresh ? ld_fresh(B.status + env) : B.status[env]);   // (a step in parts needs the status buffer: checked by the caller)
    WSYNC();
    float time = fresh ? ld_fresh(B.time + env) : B.time[env];
    PROF(0);
    if (first) {
        pd_torque(m, w, a.pd_mode, a.action_cols, a.action + (size_t)env * NU * a.action_cols, a.kp, a.kd,
                  a.pd_time ? a.pd_time[env] : time, a.pd_dt, lag_q, lag_v, lane);
        LANE_IF(k, NU) if ((a.ctrl_mask >> k) & 1u) w.ctrl[k] = w.tau[k];
        WSYNC();
    }
    LANE_IF(z, 1) w.phase_lock = a.phase_lock;
    WSYNC();
    PROF(1);
    int cost = first ? 0 : (fresh ? ld_fresh(B.cost + env) : B.cost[env]), cost_sum = cost;
    for (int f = f_begin; f < f_end; f++) {
        if (f == a.n_frames - 1) {  // the state entering the last forward is what the next PD law evaluates M / bias at
            store_row(B.qpos_lag + (size_t)env * NQ, w.qpos, NQ, lane);
            store_row(B.qvel_lag + (size_t)env * NV, w.qvel, NV, lane);
        }
        substep(m, w, time, lane);
#if defined(PBRL_HOST_EMU) && defined(PBRL_EMU_TRACE)
        if (g_emu_trace) { g_emu_trace[2 * ((size_t)env * a.n_frames + f)] = w.niter; g_emu_trace[2 * ((size_t)env * a.n_frames + f) + 1] = (w.work - w.niter) / PBRL_DENSE_WEIGHT; }
#endif
        cost = cost > w.work ? cost : w.work;  // the most expensive substep decides how long the CTA's barriers wait for this env
        cost_sum += w.work;
    }
    if (a.cost_mode == 1) cost = cost_sum;
    else if (a.cost_mode == 2) cost = w.work;
    // the two reference frames at the clip time the step ends on, into the (now dead) factor workspace; they land during the stores
    if (mimic && last) frames_prefetch(a.frames, a.mm, w, clip, B.ref_t0[env] + time, lane);
    float bad = wsum(lane, NQ + NV, [&](int i) { float v = i < NQ ? w.qpos[i] : w.qvel[i - NQ]; return (v == v && fabsf(v) < 1e30f) ? 0.f : 1.f; });
    store_row(B.qpos + (size_t)env * NQ, w.qpos, NQ, lane);
    store_row(B.qvel + (size_t)env * NV, w.qvel, NV, lane);
    store_row(B.warm + (size_t)env * NV, w.warm, NV, lane);
    if (first) store_row(B.ctrl + (size_t)env * NU, w.ctrl, NU, lane);
    LANE_IF(z, 1) {
        B.time[env] = time;
        if (B.status) B.status[env] = w.status | (bad > 0.f ? PBRL_STATUS_NAN : 0);
        if (B.cost) B.cost[env] = cost;
    }
    PROF(24);
    if (!last) { WSYNC(); return; }
    if (mimic) {
        if (B.obs) { store_row(B.obs + (size_t)env * PBRL_OBS, w.qpos, NQ, lane); store_row(B.obs + (size_t)env * PBRL_OBS + NQ, w.qvel, NV, lane); }
        mimic_epilogue(m, a.mm, a.frames, w, B, env, clip, B.ref_t0[env] + time, acc, lane);
    } else {
        write_obs_reference(a, w, env, 1.0f, lane);
    }
    WSYNC();
    PROF(25);
}

// HumanoidDiff.reset (agent_env_template.py:88-112): qpos0, zero velocity
PB_DEVN void env_reset(const DevModel& m, Ws& w, const StepArgs& a, int env, int lane) {
    const PbrlBuffers& B = a.buf;
    LANE_LOOP(i, NQ) w.qpos[i] = m.qpos0[i];
    LANE_LOOP(i, NV) w.qvel[i] = 0.f;
    WSYNC();
    store_row(B.qpos + (size_t)env * NQ, w.qpos, NQ, lane);
    store_row(B.qpos_lag + (size_t)env * NQ, w.qpos, NQ, lane);
    store_row(B.qvel + (size_t)env * NV, w.qvel, NV, lane);
    store_row(B.qvel_lag + (size_t)env * NV, w.qvel, NV, lane);
    store_row(B.warm + (size_t)env * NV, w.qvel, NV, lane);
    store_row(B.ctrl + (size_t)env * NU, w.qvel, NU, lane);
    LANE_IF(z, 1) { B.time[env] = 0.f; if (B.status) B.status[env] = 0; }
    LANE_IF(k, 3) if (B.metrics) B.metrics[(size_t)env * 3 + k] = 0.f;
    write_obs_reference(a, w, env, 0.0f, lane);
    WSYNC();
}

// HumanoidReplay.step (agent_replay_motion.py:51-83): kinematics at the given pose, obs = [qpos, qvel]
PB_DEVN void env_replay(const DevModel& m, Ws& w, const float* qpos, const float* qvel, float* obs, float* reward,
                        float* done, float* xpos, float* xquat, int env, int lane) {
    load_row(w.qpos, qpos + (size_t)env * NQ, NQ, lane);
    load_row(w.qvel, qvel + (size_t)env * NV, NV, lane);
    WSYNC();
    if (xpos || xquat) stage_fk(m, w, w.qpos, lane);
    if (obs) { store_row(obs + (size_t)env * PBRL_OBS, w.qpos, NQ, lane); store_row(obs + (size_t)env * PBRL_OBS + NQ, w.qvel, NV, lane); }
    LANE_IF(k, 3) if (reward) reward[(size_t)env * 3 + k] = 0.f;
    LANE_IF(z, 1) if (done) done[env] = 1.0f;
    if (xpos) LANE_LOOP(i, NB * 3) xpos[(size_t)env * NB * 3 + i] = w.xpos[i];
    if (xquat) LANE_LOOP(i, NB * 4) xquat[(size_t)env * NB * 4 + i] = w.xquat[S4 * (i / 4) + (i & 3)];
    WSYNC();
}

// Kinematic replay of one (env, frame) item by ONE thread (k_replay_rollout): the same kinematics as stage_fk, walked body by
// body in parents-first order; poses are written to / read back from the item's own row (xpos: 14 x 3, xquat: 14 x 4).
// Thread-serial, so the scalar cost matters: half-angle sine / cosine by odd / even polynomials after a reduction to
// |r| <= pi / 2 (error < 1e-7), reciprocal square roots for the normalisations, vectors rotated by
// the quaternion directly (v + w t + u x t, t = 2 u x v) instead of through a rotation matrix.
PB_DEV void quat_normalize_fast(float* q) {
    const float ss = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
    if (ss < 1e-30f) { q[0] = 1.f; q[1] = q[2] = q[3] = 0.f; return; }
    const float inv = fast_rsqrt(ss);
    q[0] *= inv; q[1] *= inv; q[2] *= inv; q[3] *= inv;
}
// The float tables the item kinematics read.  k_replay_rollout receives them as a kernel parameter (constant bank, uniform
// loads: no L1 traffic); the body tree itself is compile-time (pbrl_topo_gen.h), so the walk is fully unrolled, every qpos
// load has an immediate offset and the poses of the ancestors live in registers.
struct ReplayTables {
    float body_pos[NB][3], jnt_axis[NJ][3], qpos0[NQ + 1];
};
inline void replay_tables_from(const DevModel& m, ReplayTables& t) {  // host side
    for (int b = 0; b < NB; b++) for (int c = 0; c < 3; c++) t.body_pos[b][c] = m.body_pos[b][c];
    for (int j = 0; j < NJ; j++) for (int c = 0; c < 3; c++) t.jnt_axis[j][c] = m.jnt_axis[j][c];
    for (int i = 0; i <= NQ; i++) t.qpos0[i] = m.qpos0[i];
}
namespace replay {
// q <- q (x) rot(axis K, half-angle cosine c, sine s), K = 1 / 2 / 3: a coordinate axis (the sign of the axis is folded into s)
template <int K>
PB_DEV void quat_mul_axis(float (&q)[4], float c, float s) {
    const float w = q[0], x = q[1], y = q[2], z = q[3];
    if constexpr (K == 1) { q[0] = w * c - x * s; q[1] = w * s + x * c; q[2] = y * c + z * s; q[3] = z * c - y * s; }
    else if constexpr (K == 2) { q[0] = w * c - y * s; q[1] = x * c - z * s; q[2] = w * s + y * c; q[3] = z * c + x * s; }
    else { q[0] = w * c - z * s; q[1] = x * c + y * s; q[2] = y * c - x * s; q[3] = w * s + z * c; }
}
// ql <- product of the rotations of hinges J .. JEND-1 of a body (J0: its first hinge).
// AXSPEC: the model's hinge axes equal topo::JNT_AXIS_KIND (checked at model creation) -> axis-aligned hinges compose cheaply
template <int J, int J0, int JEND, bool AXSPEC>
struct Hinges {
    PB_MEMBER void run(const ReplayTables& T, const float* qpos, float (&ql)[4]) {
        if constexpr (J < JEND) {
            float sn, cs;
            sincos_half_angle(0.5f * (qpos[6 + J] - T.qpos0[6 + J]), &sn, &cs);
            constexpr int kind = AXSPEC ? topo::JNT_AXIS_KIND[J] : 0;
            if constexpr (kind != 0) {
                constexpr int k = kind > 0 ? kind : -kind;
                const float s = kind > 0 ? sn : -sn;
                if constexpr (J == J0) { ql[0] = cs; ql[1] = k == 1 ? s : 0.f; ql[2] = k == 2 ? s : 0.f; ql[3] = k == 3 ? s : 0.f; }
                else quat_mul_axis<k>(ql, cs, s);
            } else {
                float jq[4], qn[4];
                jq[0] = cs; jq[1] = T.jnt_axis[J][0] * sn; jq[2] = T.jnt_axis[J][1] * sn; jq[3] = T.jnt_axis[J][2] * sn;
                if constexpr (J == J0) { for (int c = 0; c < 4; c++) ql[c] = jq[c]; }
                else { quat_mul(ql, jq, qn); for (int c = 0; c < 4; c++) ql[c] = qn[c]; }
            }
            Hinges<J + 1, J0, JEND, AXSPEC>::run(T, qpos, ql);
        }
    }
};
template <int B, bool AXSPEC>
struct Body {
    PB_MEMBER void run(const ReplayTables& T, const float* qpos, float (&xp)[NB][3], float (&xq)[NB][4], float* xpos, float* xquat) {
        constexpr int P = topo::BODY_PARENT[B], J0 = topo::BODY_JNTADR[B], NJB = topo::BODY_JNTNUM[B];
        float q[4];
        if constexpr (J0 == 0) {  // the free joint
            for (int c = 0; c < 3; c++) xp[B][c] = qpos[c];
            for (int c = 0; c < 4; c++) q[c] = qpos[3 + c];
        } else {
            float t[3], ut[3], ql[4];
            const float* v = T.body_pos[B];
            cross3(xq[P] + 1, v, t);
            for (int c = 0; c < 3; c++) t[c] *= 2.f;
            cross3(xq[P] + 1, t, ut);
            for (int c = 0; c < 3; c++) xp[B][c] = xp[P][c] + (v[c] + xq[P][0] * t[c] + ut[c]);
            Hinges<J0, J0, J0 + NJB, AXSPEC>::run(T, qpos, ql);
            quat_mul(xq[P], ql, q);
        }
        quat_normalize_fast(q);
        for (int c = 0; c < 4; c++) xq[B][c] = q[c];
        for (int c = 0; c < 3; c++) xpos[3 * B + c] = xp[B][c];
        for (int c = 0; c < 4; c++) xquat[4 * B + c] = q[c];
        if constexpr (B + 1 < NB) Body<B + 1, AXSPEC>::run(T, qpos, xp, xq, xpos, xquat);
    }
};
}  // namespace replay
template <bool AXSPEC>
PB_DEV void replay_item_fk(const ReplayTables& T, const float* qpos, float* xpos, float* xquat) {
    float xp[NB][3], xq[NB][4];
    xp[0][0] = xp[0][1] = xp[0][2] = 0.f;
    xq[0][0] = 1.f; xq[0][1] = xq[0][2] = xq[0][3] = 0.f;
    for (int c = 0; c < 3; c++) xpos[c] = 0.f;
    xquat[0] = 1.f; xquat[1] = xquat[2] = xquat[3] = 0.f;
    replay::Body<1, AXSPEC>::run(T, qpos, xp, xq, xpos, xquat);
}

// separate-controller path (testing_pds_env3.ipynb cell 21): tau, and optionally dense qM / qfrc_bias
PB_DEVN void env_pd(const DevModel& m, Ws& w, int pd_mode, int action_cols, const float* action, const float* qpos,
                    const float* qvel, const float* qpos_lag, const float* qvel_lag, const float* kp, const float* kd,
                    const float* time, float pd_dt, float* tau, float* qM, float* bias, int env, int lane) {
    load_row(w.qpos, qpos + (size_t)env * NQ, NQ, lane);
    load_row(w.qvel, qvel + (size_t)env * NV, NV, lane);
    float* lag_q = w.fsmooth;
    float* lag_v = w.qacc;
    load_row(lag_q, qpos_lag + (size_t)env * NQ, NQ, lane);
    load_row(lag_v, qvel_lag + (size_t)env * NV, NV, lane);
    LANE_IF(z, 1) w.status = 0;
    WSYNC();
    if (qM || bias) {
        smooth_position(m, w, lag_q, lane);
        smooth_velocity(m, w, lag_v, lane);
        if (qM) {
            LANE_LOOP(i, NV * NV) qM[(size_t)env * NV * NV + i] = 0.f;
            WSYNC();
            LANE_LOOP(e, NM) {
                int r = m.M_row[e], c = m.M_col[e];
                qM[(size_t)env * NV * NV + r * NV + c] = w.M[e];
                qM[(size_t)env * NV * NV + c * NV + r] = w.M[e];
            }
        }
        if (bias) store_row(bias + (size_t)env * NV, w.bias, NV, lane);
        WSYNC();
    }
    pd_torque(m, w, pd_mode, action_cols, action + (size_t)env * NU * action_cols, kp, kd, time[env], pd_dt, lag_q, lag_v, lane);
    store_row(tau + (size_t)env * NU, w.tau, NU, lane);
    WSYNC();
}

// one mjx.forward with every stage dumped (layout: PBRL_DBG_* in pbrl.h)
PB_DEVN void env_debug_forward(const DevModel& m, Ws& w, const float* qpos, const float* qvel, const float* ctrl,
                               const float* warm, float* out, int env, int lane) {
    load_row(w.qpos, qpos + (size_t)env * NQ, NQ, lane);
    load_row(w.qvel, qvel + (size_t)env * NV, NV, lane);
    load_row(w.ctrl, ctrl + (size_t)env * NU, NU, lane);
    load_row(w.warm, warm + (size_t)env * NV, NV, lane);
    LANE_IF(z, 1) { w.status = 0; w.phase_lock = 0; }
    WSYNC();
    float* o = out + (size_t)env * PBRL_DBG_SIZE;
    LANE_LOOP(i, PBRL_DBG_SIZE) o[i] = 0.f;
    WSYNC();
    stage_fk(m, w, w.qpos, lane);  // body poses are dumped first: the workspace reuses xquat/xmat after the contact stage
    LANE_LOOP(i, NB * 3) o[PBRL_DBG_XPOS + i] = w.xpos[i];
    LANE_LOOP(i, NB * 4) o[PBRL_DBG_XQUAT + i] = w.xquat[S4 * (i / 4) + (i & 3)];
    WSYNC();
    forward(m, w, lane);
    LANE_IF(i, 3) o[PBRL_DBG_COM + i] = w.com[i];
    LANE_LOOP(e, NM) {
        int r = m.M_row[e], c = m.M_col[e];
        o[PBRL_DBG_QM + r * NV + c] = w.M[e];
        o[PBRL_DBG_QM + c * NV + r] = w.M[e];
    }
    LANE_LOOP(i, NV) {
        float act = 0.f;
        const int a = m.dof_act[i];
        if (a >= 0) act = m.dof_gear[i] * fminf(fmaxf(w.ctrl[a], m.dof_ctrl_lo[i]), m.dof_ctrl_hi[i]);
        o[PBRL_DBG_BIAS + i] = -m.dof_damping[i] * w.qvel[i] + act - w.fsmooth[i];  // qfrc_bias recovered from qfrc_smooth
        o[PBRL_DBG_QFRC_SMOOTH + i] = w.fsmooth[i];
        o[PBRL_DBG_QACC + i] = w.qacc[i];
        o[PBRL_DBG_QFRC_CONSTRAINT + i] = w.fcon[i];
    }
    WSYNC();
    // qacc_smooth is only formed by the step when no row is active; the dump always shows it
    LANE_LOOP(i, NV) w.asmooth[i] = w.fsmooth[i];
    LANE_LOOP(e, NM) w.LD[e] = w.M[e];
    WSYNC();
    ldl_factor_solve(m, w, w.asmooth, lane);
    LANE_LOOP(i, NV) o[PBRL_DBG_QACC_SMOOTH + i] = w.asmooth[i];
    float nact = wsum(lane, NC, [&](int c) { return w.con_on[c] ? 1.f : 0.f; });
    LANE_IF(z, 1) { o[PBRL_DBG_NCON_ACTIVE] = nact; o[PBRL_DBG_NITER] = (float)w.niter; }
    LANE_IF(c, NC) {
        o[PBRL_DBG_CON_DIST + c] = w.con_dist[c];
        for (int k = 0; k < 3; k++) o[PBRL_DBG_CON_POS + 3 * c + k] = w.con_r[SR * c + k] + w.com[k];
        for (int e = 0; e < 4; e++)
            o[PBRL_DBG_CON_FORCE + 4 * c + e] = (w.con_on[c] && w.con_jar[4 * c + e] < 0.f) ? -w.con_D[c] * w.con_jar[4 * c + e] : 0.f;
    }
    LANE_IF(l, NLIM) o[PBRL_DBG_LIM_FORCE + l] = (w.lim_sign[l] != 0.f && w.lim_jar[l] < 0.f) ? -w.lim_D[l] * w.lim_jar[l] : 0.f;
    LANE_IF(z, 1) o[PBRL_DBG_NSELF] = (float)w.nsc;
    LANE_IF(l, 32) {
        if (l < 4 * w.nsc) {
            if ((l & 3) == 0) o[PBRL_DBG_SELF_DIST + (l >> 2)] = w.sc_n[l + 3];
            o[PBRL_DBG_SELF_FORCE + l] = w.sc_jar[l] < 0.f ? -w.sc_r[(l & ~3) + 3] * w.sc_jar[l] : 0.f;
        }
    }
    WSYNC();
}

}  // namespace pbrl
