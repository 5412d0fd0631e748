// Device-side model constants (fp32) and their construction from the C-ABI PbrlModelDesc.
//
// The kernels are specialised to the dimensions of the reference's humanoid
// (/root/reference/models/final_humanoid.xml: 14 bodies, nq 35, nv 34, nu 28, 16 geoms; SURVEY.md
// App. B) but read the topology (parents, levels, ancestor lists, contact slots) as data.
#pragma once
#include <cmath>
#include <cstring>
#include <string>

#include "../../include/pbrl.h"
#include "pbrl_topo_gen.h"

namespace pbrl {

constexpr int NB = PBRL_NBODY;
constexpr int NQ = PBRL_NQ;
constexpr int NV = PBRL_NV;
constexpr int NU = PBRL_NU;
constexpr int NJ = PBRL_NJNT;
constexpr int NG = PBRL_NGEOM;
constexpr int NM = 310;   // non-zeros of the tree-sparse joint-space inertia (row i: i and its ancestors)
constexpr int NLIM = 28;  // limited hinges
constexpr int NC = 29;    // plane contact slots: 5 spheres + 8 capsules x 2 + 2 boxes x 4
constexpr int NCP = 32;   // padded
constexpr int NSG = 15;   // geoms that can touch each other (everything but the floor)
constexpr int NSP = 96;   // candidate pairs between bodies (SC 32 + CC 24 + SS 8 + SB 10 + CB 14 + BB 1 = 89)
constexpr int KS = 8;     // body-vs-body contacts kept per env and substep (inside the margin at the same time)
constexpr int NFAC = 1432;  // rank-1 update operations of one tree-sparse L^T D L factorisation
constexpr int MAXANC = 12;  // deepest dof chain (ancestors of one dof); checked against the compiled tree below
constexpr int MAXLEVEL = 8;
constexpr int MAXDEPTH = 16;

// ---- compile-time view of the dof tree (pbrl_topo_gen.h) ------------------------------------------------------
// The factorisation / solve code is unrolled over the dofs, so depths, row addresses, ancestor lists and the chain
// decomposition below are immediates.  A "slot" is a maximal chain of the dof tree (a dof continues its parent's
// chain iff it is the parent's first child): dofs of one slot have distinct, consecutive depths and consecutive
// indices, and two dofs of equal depth never share a slot.
#if defined(__CUDACC__)
#define PB_CX __host__ __device__ constexpr
#else
#define PB_CX constexpr
#endif
namespace topo {
PB_CX int depth(int k) { int d = 0; while (PARENT[k] >= 0) { k = PARENT[k]; d++; } return d; }
PB_CX int madr(int k) { int a = 0; for (int i = 0; i < k; i++) a += depth(i) + 1; return a; }
PB_CX int fac_total() { int a = 0; for (int i = 0; i < NDOF; i++) a += depth(i) * (depth(i) + 1) / 2; return a; }
PB_CX int max_depth() { int d = 0; for (int k = 0; k < NDOF; k++) d = depth(k) > d ? depth(k) : d; return d; }
PB_CX int depth_count(int lv) { int n = 0; for (int k = 0; k < NDOF; k++) n += depth(k) == lv; return n; }
PB_CX int depth_first(int lv) { int n = 0; for (int l = 0; l < lv; l++) n += depth_count(l); return n; }
PB_CX int anc_at(int k, int t) { while (depth(k) > t) k = PARENT[k]; return k; }  // ancestor of k at depth t <= depth(k)
PB_CX bool first_child(int k) { for (int j = PARENT[k] + 1; j < k; j++) if (PARENT[j] == PARENT[k]) return false; return true; }
PB_CX int slot(int k) {
    if (PARENT[k] < 0) return 0;
    if (first_child(k)) return slot(PARENT[k]);
    int n = 0;
    for (int j = 1; j <= k; j++) if (PARENT[j] >= 0 && !first_child(j)) n++;
    return n;
}
PB_CX int nslot() { int n = 0; for (int k = 0; k < NDOF; k++) n = slot(k) + 1 > n ? slot(k) + 1 : n; return n; }
PB_CX int slot_first(int s) { for (int k = 0; k < NDOF; k++) if (slot(k) == s) return k; return -1; }
PB_CX int slot_len(int s) { int n = 0; for (int k = 0; k < NDOF; k++) n += slot(k) == s; return n; }
PB_CX int slot_fd(int s) { return depth(slot_first(s)); }
PB_CX bool slots_are_chains() {
    int roots = 0;
    for (int k = 0; k < NDOF; k++) {
        roots += PARENT[k] < 0;
        if (k != slot_first(slot(k)) + depth(k) - slot_fd(slot(k))) return false;
    }
    return roots == 1;
}
// lanes (= depths) at which the ancestors of k live in slot s
PB_CX unsigned anc_mask(int k, int s) { unsigned m = 0; for (int t = 0; t < depth(k); t++) if (slot(anc_at(k, t)) == s) m |= 1u << t; return m; }
// i-th dof in elimination order: deepest first (descendants before ancestors; independent chains interleave)
PB_CX int elim_order(int i) {
    int n = 0;
    for (int d = max_depth(); d >= 0; d--)
        for (int k = NDOF - 1; k >= 0; k--)
            if (depth(k) == d) { if (n == i) return k; n++; }
    return -1;
}
PB_CX bool hi_source_at(int t) { for (int k = 0; k < NDOF; k++) if (depth(k) > t && anc_at(k, t) >= 32) return true; return false; }
static_assert(NDOF == PBRL_NV && madr(NDOF - 1) + depth(NDOF - 1) + 1 == NM && fac_total() == NFAC, "pbrl_topo_gen.h does not match the kernel dimensions");
static_assert(max_depth() <= MAXANC && max_depth() < 32 && NDOF <= 64, "dof chains too deep for the lane-per-depth factorisation");
static_assert(slots_are_chains(), "the chain decomposition expects depth-first dof numbering with one root");
}  // namespace topo
constexpr int NSLOT = topo::nslot();

enum { DOF_FREE_T0 = 0, DOF_FREE_R0 = 3, DOF_HINGE = 6 };
enum { CON_SPHERE = 0, CON_CAPSULE = 1, CON_BOX = 3 };

typedef unsigned char u8;
typedef signed char i8;
typedef unsigned short u16;

// Compact (~6.5 KB): every CTA keeps a copy in shared memory, so the lane-indexed table lookups of the stages cost a
// shared-memory access instead of an L1-miss / L2 round trip (the L1 is only ~28 KB next to 227 KB of workspaces).
struct DevModel {
    float timestep, gx, gy, gz;
    float sol_k, sol_b, sol_dmin, sol_dmax, sol_width, sol_mid, sol_power, meaninertia;
    float total_mass, plane_z;
    // constraint-solver options (include/pbrl.h: solver_rule)
    int solver_rule, solver_iterations, ls_iterations;
    float solver_tol;   // tolerance * meaninertia * max(1, nv): bound on the cost improvement and on |gradient| (rule 1)
    float ls_gtol;      // tolerance * ls_tolerance * meaninertia * max(1, nv): bound on |phi'| / |search| (rule 1)
    int nlevel;
    int ndepth;
    float body_pos[NB][3], body_mass[NB], body_ipos[NB][3], body_inertia[NB][6];
    float jnt_axis[NJ][3], jnt_lo[NJ], jnt_hi[NJ], jnt_margin[NJ];
    float dof_armature[NV], dof_damping[NV], dof_invweight0[NV], dof_gear[NV], dof_ctrl_lo[NV], dof_ctrl_hi[NV];
    float con_gpos[NCP][3], con_gquat[NCP][4], con_size[NCP][3], con_margin[NCP], con_mu[NCP], con_invweight[NCP];
    float qpos0[NQ + 1];
    // sparse inertia layout: entry e of row i holds M[i, anc_a(i)] at e = dof_Madr[i] + a (a = 0: diagonal)
    u16 dof_Madr[NV];
    u8 M_row[NM], M_col[NM];
    alignas(4) u8 anc_dof[NV][MAXANC];  // anc_dof[k][t]: ancestor of dof k at depth t < depth(k) (back substitution sources)
    u8 level_start[MAXLEVEL + 1], level_body[NB], depth_start[MAXDEPTH + 1], depth_dof[NV];
    u8 body_parent[NB], body_subtree_end[NB], body_jntadr[NB], body_jntnum[NB], body_dofadr[NB], body_dofnum[NB];
    u8 body_conadr[NB], body_connum[NB], body_nchild[NB], body_child[NB][4];
    alignas(4) u8 body_anc[NB][4];
    unsigned body_conmask[NB], body_subconmask[NB], body_submask[NB];  // contact slots of the body / of its subtree; bodies of its subtree  // self, parent, grandparent, ... padded with the world body (its partial sums are zero)
    u8 jnt_type[NJ], jnt_qposadr[NJ], jnt_dofadr[NJ], jnt_limited[NJ], jnt_bodyid[NJ];
    u8 dof_bodyid[NV], dof_jntid[NV], dof_depth[NV], dof_kind[NV];
    i8 dof_parentid[NV], dof_act[NV];
    u8 lim_jnt[NLIM];
    u8 con_body[NCP], con_kind[NCP], con_sub[NCP], con_geom[NCP];
    // body-vs-body candidates: geom table + pair list ordered [sphere-capsule][capsule-capsule][sphere-sphere][sphere-box]
    // so that the 32 lanes of a pass run (mostly) the same distance function
    int nsp, nsp_sc, nsp_cc, nsp_ss;   // pairs in total / sphere-capsule / capsule-capsule / sphere-sphere (0 pairs: self-collision off)
    int nsp_sb, nsp_cb, nsp_bb;        // sphere-box / capsule-box (2 contact slots each) / box-box (4 contact slots) pairs
    int nbox_pairs;                    // capsule-box + box-box candidates
    float box_bound[NSG];              // boxes: radius of the bounding sphere (prefilter of the capsule-box / box-box pairs)
    float self_margin, self_mu, self_invw_scale;  // includemargin, sliding friction, (1 + mu^2) 2 mu^2 / impratio
    float body_invw[NB];               // body_invweight0[:, 0]
    float sg_pos[NSG][3], sg_axis[NSG][3], sg_r[NSG], sg_hl[NSG];  // (boxes: sg_axis holds the half sizes, the box frame is the body frame)
    u8 sg_body[NSG], pad_sg_;
    u8 sp_g1[NSP], sp_g2[NSP];
};
static_assert(sizeof(DevModel) % 4 == 0, "DevModel is copied to shared memory word by word");

// ---- mimic configuration on the device --------------------------------------------------------
struct DevMimic {
    int nclip, nframes_total;
    int clip_first[PBRL_MAXCLIP], clip_nframes[PBRL_MAXCLIP], clip_wrap[PBRL_MAXCLIP], clip_fall[PBRL_MAXCLIP];
    float clip_duration[PBRL_MAXCLIP], clip_dt[PBRL_MAXCLIP];
    float w_pose, w_vel, w_ee, w_com, k_pose, k_vel, k_ee, k_com;
    int n_ee, ee_body[4];
    float ee_offset[4][3];
    unsigned fall_geom_mask;
    int body_slot[NB];  // index of the body's 3-dof joint quaternion in a table frame, -1 for 1-dof / root / world
};

inline std::string build_dev_mimic(const PbrlMimicDesc& d, const DevModel& m, DevMimic& mm) {
    std::memset(&mm, 0, sizeof(mm));
    if (d.nclip < 1 || d.nclip > PBRL_MAXCLIP || !d.frames || d.nframes_total < 2) return "bad mimic description";
    mm.nclip = d.nclip; mm.nframes_total = d.nframes_total;
    for (int c = 0; c < d.nclip; c++) {
        if (d.clip_nframes[c] < 2 || d.clip_first[c] < 0 || d.clip_first[c] + d.clip_nframes[c] > d.nframes_total) return "clip range outside the table";
        if (!(d.clip_dt[c] > 0.f) || !(d.clip_duration[c] > 0.f)) return "clip duration / dt must be positive";
        mm.clip_first[c] = d.clip_first[c]; mm.clip_nframes[c] = d.clip_nframes[c];
        mm.clip_wrap[c] = d.clip_wrap[c]; mm.clip_fall[c] = d.clip_fall_enabled[c];
        mm.clip_duration[c] = d.clip_duration[c]; mm.clip_dt[c] = d.clip_dt[c];
    }
    mm.w_pose = d.w_pose; mm.w_vel = d.w_vel; mm.w_ee = d.w_ee; mm.w_com = d.w_com;
    mm.k_pose = d.k_pose; mm.k_vel = d.k_vel; mm.k_ee = d.k_ee; mm.k_com = d.k_com;
    if (d.n_ee < 0 || d.n_ee > 4) return "n_ee must be 0..4";
    mm.n_ee = d.n_ee;
    for (int e = 0; e < d.n_ee; e++) {
        if (d.ee_body[e] < 1 || d.ee_body[e] >= NB) return "bad end-effector body";
        mm.ee_body[e] = d.ee_body[e];
        for (int k = 0; k < 3; k++) mm.ee_offset[e][k] = d.ee_offset[3 * e + k];
    }
    mm.fall_geom_mask = d.fall_geom_mask;
    int slot = 0;
    for (int b = 0; b < NB; b++) {
        mm.body_slot[b] = -1;
        if (b > 1 && m.body_jntnum[b] == 3) mm.body_slot[b] = slot++;
        else if (b > 1 && m.body_jntnum[b] != 1) return "mimic mode needs 1- or 3-hinge bodies";
    }
    if (slot != 8) return "mimic mode expects eight 3-dof joints";
    return "";
}

inline std::string build_dev_model(const PbrlModelDesc& d, DevModel& m) {
    std::memset(&m, 0, sizeof(m));
    if (d.nbody[0] != NB || d.nq[0] != NQ || d.nv[0] != NV || d.nu[0] != NU || d.njnt[0] != NJ || d.ngeom[0] != NG)
        return "model dimensions differ from the humanoid the kernels are specialised to (14 bodies, nq 35, nv 34, nu 28, 29 joints, 16 geoms)";
    m.timestep = (float)d.timestep[0];
    m.gx = (float)d.gravity[0]; m.gy = (float)d.gravity[1]; m.gz = (float)d.gravity[2];
    {   // mjx constraint._kbi constants
        double timeconst = std::fmax(d.solref[0], 2.0 * d.timestep[0]), dampratio = d.solref[1];
        double dmin = std::fmin(std::fmax(d.solimp[0], 1e-4), 0.9999), dmax = std::fmin(std::fmax(d.solimp[1], 1e-4), 0.9999);
        double k = 1.0 / (dmax * dmax * timeconst * timeconst * dampratio * dampratio), b = 2.0 / (dmax * timeconst);
        if (d.solref[0] <= 0) k = -d.solref[0] / (dmax * dmax);
        if (d.solref[1] <= 0) b = -d.solref[1] / dmax;
        m.sol_k = (float)k; m.sol_b = (float)b; m.sol_dmin = (float)dmin; m.sol_dmax = (float)dmax;
        m.sol_width = (float)std::fmax(1e-15, d.solimp[2]);
        m.sol_mid = (float)std::fmin(std::fmax(d.solimp[3], 1e-4), 0.9999);
        m.sol_power = (float)std::fmax(1.0, d.solimp[4]);
    }
    m.meaninertia = (float)d.meaninertia[0];
    if (d.solver_rule[0] != PBRL_SOLVER_EXACT && d.solver_rule[0] != PBRL_SOLVER_MJX) return "solver_rule must be 0 (exact active set) or 1 (MJX tolerance rule)";
    if (d.iterations[0] < 1 || d.ls_iterations[0] < 1) return "iterations and ls_iterations must be >= 1";
    m.solver_rule = d.solver_rule[0];
    m.solver_iterations = d.iterations[0];
    m.ls_iterations = d.ls_iterations[0];
    m.solver_tol = (float)(d.tolerance[0] * d.meaninertia[0] * (NV > 1 ? NV : 1));
    m.ls_gtol = (float)(d.tolerance[0] * d.ls_tolerance[0] * d.meaninertia[0] * (NV > 1 ? NV : 1));
    int level[NB] = {0};
    for (int b = 0; b < NB; b++) {
        m.body_parent[b] = d.body_parent[b];
        if (b > 0) {
            if (d.body_parent[b] >= b) return "bodies must be ordered parents-first";
            level[b] = level[d.body_parent[b]] + 1;
        }
        m.body_jntadr[b] = d.body_jntadr[b]; m.body_jntnum[b] = d.body_jntnum[b];
        m.body_dofadr[b] = d.body_dofadr[b]; m.body_dofnum[b] = d.body_dofnum[b];
        if (b > 0 && d.body_dofnum[b] == 0) return "every moving body needs at least one dof";
        m.body_mass[b] = (float)d.body_mass[b];
        m.total_mass += (float)d.body_mass[b];
        for (int k = 0; k < 3; k++) { m.body_pos[b][k] = (float)d.body_pos[3 * b + k]; m.body_ipos[b][k] = (float)d.body_ipos[3 * b + k]; }
        const double* I = d.body_inertia + 9 * b;
        m.body_inertia[b][0] = (float)I[0]; m.body_inertia[b][1] = (float)I[4]; m.body_inertia[b][2] = (float)I[8];
        m.body_inertia[b][3] = (float)I[1]; m.body_inertia[b][4] = (float)I[2]; m.body_inertia[b][5] = (float)I[5];
    }
    // depth-first order check + subtree ends
    for (int b = 0; b < NB; b++) {
        int e = b + 1;
        while (e < NB) {
            int a = e;
            while (a > b) a = d.body_parent[a];
            if (a != b) break;
            e++;
        }
        m.body_subtree_end[b] = e;
        for (int c = e; c < NB; c++) {  // no later body may descend from b
            int a = c;
            while (a > b) a = d.body_parent[a];
            if (a == b && b != 0) return "bodies must be in depth-first order";
        }
    }
    for (int b = 0; b < NB; b++) {
        int a = b;
        for (int k = 0; k < 4; k++) { m.body_anc[b][k] = (u8)a; a = a > 0 ? d.body_parent[a] : 0; }
        if (a != 0) return "body tree deeper than 4 levels";
    }
    for (int b = 1; b < NB; b++) {
        int p = d.body_parent[b];
        if (m.body_nchild[p] >= 4) return "at most 4 children per body";
        m.body_child[p][m.body_nchild[p]++] = b;
    }
    int nl = 0;
    for (int b = 0; b < NB; b++) nl = level[b] + 1 > nl ? level[b] + 1 : nl;
    if (nl > MAXLEVEL) return "tree too deep";
    m.nlevel = nl;
    int pos = 0;
    for (int l = 0; l < nl; l++) {
        m.level_start[l] = pos;
        for (int b = 0; b < NB; b++) if (level[b] == l) m.level_body[pos++] = b;
    }
    m.level_start[nl] = pos;
    int nlim = 0;
    for (int j = 0; j < NJ; j++) {
        m.jnt_type[j] = d.jnt_type[j]; m.jnt_qposadr[j] = d.jnt_qposadr[j]; m.jnt_dofadr[j] = d.jnt_dofadr[j];
        m.jnt_limited[j] = d.jnt_limited[j]; m.jnt_bodyid[j] = d.jnt_bodyid[j];
        for (int k = 0; k < 3; k++) {
            m.jnt_axis[j][k] = (float)d.jnt_axis[3 * j + k];
            if (d.jnt_pos[3 * j + k] != 0.0) return "joint anchors must sit at the body origin (jnt_pos = 0)";
        }
        m.jnt_lo[j] = (float)d.jnt_range[2 * j]; m.jnt_hi[j] = (float)d.jnt_range[2 * j + 1];
        m.jnt_margin[j] = (float)d.jnt_margin[j];
        if (d.jnt_type[j] == 0 && j != 0) return "only the first joint may be free";
        if (d.jnt_type[j] != 0 && d.jnt_type[j] != 3) return "only free and hinge joints";
        if (d.jnt_limited[j]) { if (nlim >= NLIM) return "too many limited joints"; m.lim_jnt[nlim++] = j; }
    }
    if (nlim != NLIM) return "expected 28 limited hinges";
    for (int l = 0; l < NLIM; l++)  // the kernels map dof -> limit row by offset
        if (d.jnt_dofadr[m.lim_jnt[l]] != d.jnt_dofadr[m.lim_jnt[0]] + l || d.jnt_type[m.lim_jnt[l]] != 3)
            return "limited hinges must own contiguous dofs";
    if (d.jnt_type[0] != 0) return "joint 0 must be the free root joint";
    int madr = 0, maxdepth = 0;
    for (int i = 0; i < NV; i++) {
        m.dof_bodyid[i] = d.dof_bodyid[i]; m.dof_parentid[i] = d.dof_parentid[i];
        m.dof_armature[i] = (float)d.dof_armature[i]; m.dof_damping[i] = (float)d.dof_damping[i];
        m.dof_invweight0[i] = (float)d.dof_invweight0[i];
        m.dof_depth[i] = d.dof_parentid[i] < 0 ? 0 : m.dof_depth[d.dof_parentid[i]] + 1;
        if (d.dof_parentid[i] >= i) return "dof parents must precede";
        maxdepth = m.dof_depth[i] > maxdepth ? m.dof_depth[i] : maxdepth;
        m.dof_Madr[i] = madr;
        for (int j = i; j >= 0; j = d.dof_parentid[j]) {
            if (madr >= NM) return "sparse inertia layout overflow";
            m.M_row[madr] = i; m.M_col[madr] = j; madr++;
        }
        m.dof_act[i] = -1; m.dof_gear[i] = 0; m.dof_ctrl_lo[i] = -1e30f; m.dof_ctrl_hi[i] = 1e30f;
    }
    if (madr != NM) return "sparse inertia layout size differs from 310";
    for (int i = 0; i < NV; i++)
        if (d.dof_parentid[i] != topo::PARENT[i]) return "the dof tree differs from the one the kernels were compiled for (csrc/pbrl_topo_gen.h)";
    for (int b = 1; b < NB; b++)
        if (d.body_parent[b] != topo::BODY_PARENT[b] || d.body_jntadr[b] != topo::BODY_JNTADR[b] || d.body_jntnum[b] != topo::BODY_JNTNUM[b])
            return "the body tree differs from the one the kernels were compiled for (csrc/pbrl_topo_gen.h)";
    for (int j = 1; j < NJ; j++) if (d.jnt_qposadr[j] != 6 + j || d.jnt_type[j] != 3) return "joints 1.. must be hinges reading qpos[6 + j]";
    if (maxdepth + 1 > MAXDEPTH) return "dof chains too deep";
    m.ndepth = maxdepth + 1;
    pos = 0;
    for (int l = 0; l <= maxdepth; l++) {
        m.depth_start[l] = pos;
        for (int i = 0; i < NV; i++) if (m.dof_depth[i] == l) m.depth_dof[pos++] = i;
    }
    m.depth_start[maxdepth + 1] = pos;
    for (int j = 0; j < NJ; j++) {
        int n = d.jnt_type[j] == 0 ? 6 : 1;
        for (int k = 0; k < n; k++) {
            m.dof_jntid[d.jnt_dofadr[j] + k] = j;
            m.dof_kind[d.jnt_dofadr[j] + k] = d.jnt_type[j] == 0 ? k : DOF_HINGE;
        }
    }
    for (int k = 0; k < NV; k++) {
        int chain[MAXDEPTH], n = 0;
        for (int j = d.dof_parentid[k]; j >= 0; j = d.dof_parentid[j]) chain[n++] = j;  // nearest first
        for (int t = 0; t < MAXANC; t++) m.anc_dof[k][t] = (u8)(t < n ? chain[n - 1 - t] : 0);
    }
    for (int a = 0; a < NU; a++) {
        int i = d.act_dofadr[a];
        if (m.dof_act[i] >= 0) return "one actuator per dof";
        m.dof_act[i] = a; m.dof_gear[i] = (float)d.act_gear[a];
        if (d.act_ctrllimited[a]) { m.dof_ctrl_lo[i] = (float)d.act_ctrlrange[2 * a]; m.dof_ctrl_hi[i] = (float)d.act_ctrlrange[2 * a + 1]; }
    }
    // contact slots: every non-plane geom against the floor plane (self-collisions: SURVEY.md section 8f N2, not built)
    int plane = -1;
    for (int g = 0; g < NG; g++) if (d.geom_type[g] == 0) { if (plane >= 0) return "one plane only"; plane = g; }
    if (plane < 0 || d.geom_bodyid[plane] != 0) return "the floor plane must belong to the world body";
    if (d.geom_quat[4 * plane] != 1.0 || d.geom_pos[3 * plane] != 0 || d.geom_pos[3 * plane + 1] != 0) return "floor must be z-up at the origin";
    m.plane_z = (float)d.geom_pos[3 * plane + 2];
    int nc = 0;
    for (int b = 0; b < NB; b++) {
        m.body_conadr[b] = nc;
        for (int g = 0; g < NG; g++) {
            if (g == plane || d.geom_bodyid[g] != b) continue;
            int t = d.geom_type[g], kind, n;
            if (t == 2) { kind = CON_SPHERE; n = 1; } else if (t == 3) { kind = CON_CAPSULE; n = 2; } else if (t == 6) { kind = CON_BOX; n = 4; } else return "geom type";
            double mu = std::fmax(d.geom_friction[3 * plane], d.geom_friction[3 * g]);
            double margin = std::fmax(d.geom_margin[plane], d.geom_margin[g]) - std::fmax(d.geom_gap[plane], d.geom_gap[g]);
            double tw = d.body_invweight0[2 * 0] + d.body_invweight0[2 * b];
            double invw = (tw + mu * mu * tw) * 2.0 * mu * mu / d.impratio[0];
            for (int s = 0; s < n; s++) {
                if (nc >= NC) return "too many contact slots";
                m.con_body[nc] = b; m.con_kind[nc] = kind; m.con_sub[nc] = s; m.con_geom[nc] = g;
                for (int k = 0; k < 3; k++) { m.con_gpos[nc][k] = (float)d.geom_pos[3 * g + k]; m.con_size[nc][k] = (float)d.geom_size[3 * g + k]; }
                for (int k = 0; k < 4; k++) m.con_gquat[nc][k] = (float)d.geom_quat[4 * g + k];
                m.con_margin[nc] = (float)margin; m.con_mu[nc] = (float)mu; m.con_invweight[nc] = (float)invw;
                nc++;
            }
        }
        m.body_connum[b] = nc - m.body_conadr[b];
    }
    if (nc != NC) return "expected 29 plane contact slots";
    for (int b = 0; b < NB; b++) {
        m.body_conmask[b] = m.body_connum[b] ? (((m.body_connum[b] >= 32 ? 0u : (1u << m.body_connum[b])) - 1u) << m.body_conadr[b]) : 0u;
        m.body_subconmask[b] = 0u; m.body_submask[b] = 0u;
    }
    for (int b = 0; b < NB; b++)
        for (int c = b; c < m.body_subtree_end[b]; c++) { m.body_subconmask[b] |= m.body_conmask[c]; m.body_submask[b] |= 1u << c; }
    // body-vs-body candidates among spheres and capsules (MJX evaluates every candidate every step)
    {
        int gi[PBRL_MAXGEOM], nsg = 0;
        for (int g = 0; g < NG; g++) {
            gi[g] = -1;
            if (g == plane) continue;
            if (nsg >= NSG) return "too many geoms";
            gi[g] = nsg;
            m.sg_body[nsg] = (u8)d.geom_bodyid[g];
            m.sg_r[nsg] = (float)d.geom_size[3 * g];
            m.sg_hl[nsg] = d.geom_type[g] == 3 ? (float)d.geom_size[3 * g + 1] : 0.f;
            const double* q = d.geom_quat + 4 * g;  // local z axis of the geom frame
            const double ax[3] = {2 * (q[1] * q[3] + q[0] * q[2]), 2 * (q[2] * q[3] - q[0] * q[1]), q[0] * q[0] - q[1] * q[1] - q[2] * q[2] + q[3] * q[3]};
            for (int k = 0; k < 3; k++) { m.sg_pos[nsg][k] = (float)d.geom_pos[3 * g + k]; m.sg_axis[nsg][k] = (float)ax[k]; }
            if (d.geom_type[g] == 6) {
                if (q[0] != 1.0) return "box geoms must be aligned with their body frame";
                for (int k = 0; k < 3; k++) m.sg_axis[nsg][k] = (float)d.geom_size[3 * g + k];
                m.box_bound[nsg] = (float)std::sqrt(d.geom_size[3 * g] * d.geom_size[3 * g] + d.geom_size[3 * g + 1] * d.geom_size[3 * g + 1] +
                                                    d.geom_size[3 * g + 2] * d.geom_size[3 * g + 2]);
            }
            nsg++;
        }
        for (int b = 0; b < NB; b++) m.body_invw[b] = (float)d.body_invweight0[2 * b];
        int np = d.npair[0];
        if (np < 0 || np > PBRL_MAXPAIR) return "bad npair";
        bool first = true;
        for (int pass = 0; pass < 6; pass++) {  // 0: sphere-capsule, 1: capsule-capsule, 2: sphere-sphere, 3: sphere-box, 4: capsule-box, 5: box-box
            for (int q = 0; q < np; q++) {
                const int g1 = d.pair_geom1[q], g2 = d.pair_geom2[q];
                if (g1 < 0 || g2 < 0 || g1 >= NG || g2 >= NG) return "pair geom out of range";
                const int t1 = d.geom_type[g1], t2 = d.geom_type[g2];
                if (g1 == plane || g2 == plane) continue;  // floor pairs are the contact slots above
                if ((t1 == 6 || (t2 == 6 && t1 != 2)) && pass == 0) m.nbox_pairs++;
                const int kind = (t1 == 2 && t2 == 3) ? 0 : (t1 == 3 && t2 == 3) ? 1 : (t1 == 2 && t2 == 2) ? 2 : (t1 == 2 && t2 == 6) ? 3 :
                                 (t1 == 3 && t2 == 6) ? 4 : (t1 == 6 && t2 == 6) ? 5 : -1;
                if (kind < 0) return "pair geoms must be ordered by type (sphere before capsule)";
                if (kind != pass) continue;
                if (m.nsp >= NSP) return "too many body-vs-body pairs";
                const double mu = std::fmax(d.geom_friction[3 * g1], d.geom_friction[3 * g2]);
                const double margin = std::fmax(d.geom_margin[g1], d.geom_margin[g2]) - std::fmax(d.geom_gap[g1], d.geom_gap[g2]);
                if (first) { m.self_mu = (float)mu; m.self_margin = (float)margin; first = false; }
                else if ((float)mu != m.self_mu || (float)margin != m.self_margin) return "body-vs-body pairs must share friction and margin";
                m.sp_g1[m.nsp] = (u8)gi[g1]; m.sp_g2[m.nsp] = (u8)gi[g2];
                m.nsp++;
            }
            if (pass == 0) m.nsp_sc = m.nsp;
            if (pass == 1) m.nsp_cc = m.nsp - m.nsp_sc;
            if (pass == 2) m.nsp_ss = m.nsp - m.nsp_sc - m.nsp_cc;
            if (pass == 3) m.nsp_sb = m.nsp - m.nsp_sc - m.nsp_cc - m.nsp_ss;
            if (pass == 4) m.nsp_cb = m.nsp - m.nsp_sc - m.nsp_cc - m.nsp_ss - m.nsp_sb;
            if (pass == 5) m.nsp_bb = m.nsp - m.nsp_sc - m.nsp_cc - m.nsp_ss - m.nsp_sb - m.nsp_cb;
        }
        if (m.nsp_cb + 3 * m.nsp_bb > 24) return "too many capsule-box / box-box pairs for the narrow-phase scratch";
        if (m.nsp_bb > 1) return "at most one box-box pair (its manifold construction owns the warp's scratch rows)";
        m.self_invw_scale = (1.f + m.self_mu * m.self_mu) * 2.f * m.self_mu * m.self_mu / (float)d.impratio[0];
    }
    if (d.jnt_dofadr[0] != 0) return "the free joint must own dofs 0..5";
    for (int i = 0; i < NQ; i++) m.qpos0[i] = (float)d.qpos0[i];
    return "";
}

}  // namespace pbrl
