/* CPU oracle -- TEST INFRASTRUCTURE (see pbrl_oracle.h for scope, sources and parity status).
 *
 * Dense, stage-by-stage restatement of what `mjx.step` does for the reference's humanoid
 * (SURVEY.md App. A.2-A.10).  Deliberately the "textbook" formulation (dense 34x34 Cholesky,
 * dense constraint Jacobian, exact break-point line search) so that it shares no structure with
 * the CUDA kernels it checks (tree-sparse LDL, matrix-free contact Jacobians).
 */
#include "pbrl_oracle.h"

#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#define NVMAX ORC_MAXV
#define R(x) ((REAL)(x))
#define ORC_MINVAL 1e-15
#define ORC_MINIMP 0.0001
#define ORC_MAXIMP 0.9999
enum { GEOM_PLANE = 0, GEOM_SPHERE = 2, GEOM_CAPSULE = 3, GEOM_BOX = 6 };
enum { JNT_FREE = 0, JNT_HINGE = 3 };

/* ---------------------------------------------------------------- reflection for ctypes */
int orc_real_size(void) { return (int)sizeof(REAL); }
int orc_model_size(void) { return (int)sizeof(OrcModel); }
int orc_data_size(void) { return (int)sizeof(OrcData); }
int orc_model_field(const char* name, int* offset, int* count, int* is_int) {
#define ORC_I(f, n) if (!strcmp(name, #f)) { *offset = (int)offsetof(OrcModel, f); *count = n; *is_int = 1; return 0; }
#define ORC_D(f, n) if (!strcmp(name, #f)) { *offset = (int)offsetof(OrcModel, f); *count = n; *is_int = 0; return 0; }
    ORC_MODEL_FIELDS(ORC_I, ORC_D)
#undef ORC_I
#undef ORC_D
    return -1;
}
int orc_data_field(const char* name, int* offset, int* count) {
#define ORC_X(f, n) if (!strcmp(name, #f)) { *offset = (int)offsetof(OrcData, f); *count = n; return 0; }
    ORC_DATA_FIELDS(ORC_X)
#undef ORC_X
    return -1;
}

/* ---------------------------------------------------------------- small math */
static REAL rsqrt_(REAL x) { return R(sqrt((double)x)); }
static void cross3(REAL* r, const REAL* a, const REAL* b) {
    REAL x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
    r[0] = x; r[1] = y; r[2] = z;
}
static REAL dot3(const REAL* a, const REAL* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static REAL dotn(const REAL* a, const REAL* b, int n) { REAL s = 0; for (int i = 0; i < n; i++) s += a[i] * b[i]; return s; }
static void quat_mul(REAL* r, const REAL* a, const REAL* b) {
    REAL w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
    REAL x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
    REAL y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
    REAL z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
    r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
static void quat_norm(REAL* q) {
    REAL n = rsqrt_(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    if (n < R(1e-30)) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }
    for (int i = 0; i < 4; i++) q[i] /= n;
}
static void quat_to_mat(REAL* m, const REAL* q) {
    REAL w = q[0], x = q[1], y = q[2], z = q[3];
    m[0] = w * w + x * x - y * y - z * z; m[1] = 2 * (x * y - w * z); m[2] = 2 * (x * z + w * y);
    m[3] = 2 * (x * y + w * z); m[4] = w * w - x * x + y * y - z * z; m[5] = 2 * (y * z - w * x);
    m[6] = 2 * (x * z - w * y); m[7] = 2 * (y * z + w * x); m[8] = w * w - x * x - y * y + z * z;
}
static void axis_angle_quat(REAL* q, const REAL* axis, REAL angle) {
    REAL s = R(sin(0.5 * (double)angle));
    q[0] = R(cos(0.5 * (double)angle)); q[1] = axis[0] * s; q[2] = axis[1] * s; q[3] = axis[2] * s;
}
static void mat_vec3(REAL* r, const REAL* m, const REAL* v) {
    REAL x = m[0] * v[0] + m[1] * v[1] + m[2] * v[2];
    REAL y = m[3] * v[0] + m[4] * v[1] + m[5] * v[2];
    REAL z = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
    r[0] = x; r[1] = y; r[2] = z;
}
static void matT_vec3(REAL* r, const REAL* m, const REAL* v) {
    REAL x = m[0] * v[0] + m[3] * v[1] + m[6] * v[2];
    REAL y = m[1] * v[0] + m[4] * v[1] + m[7] * v[2];
    REAL z = m[2] * v[0] + m[5] * v[1] + m[8] * v[2];
    r[0] = x; r[1] = y; r[2] = z;
}
/* spatial algebra, vectors are [ang; lin]   */
static void mul_inert_vec(REAL* r, const REAL* i, const REAL* v) {
    r[0] = i[0] * v[0] + i[3] * v[1] + i[4] * v[2] - i[8] * v[4] + i[7] * v[5];
    r[1] = i[3] * v[0] + i[1] * v[1] + i[5] * v[2] + i[8] * v[3] - i[6] * v[5];
    r[2] = i[4] * v[0] + i[5] * v[1] + i[2] * v[2] - i[7] * v[3] + i[6] * v[4];
    r[3] = i[8] * v[1] - i[7] * v[2] + i[9] * v[3];
    r[4] = i[6] * v[2] - i[8] * v[0] + i[9] * v[4];
    r[5] = i[7] * v[0] - i[6] * v[1] + i[9] * v[5];
}
static void cross_motion(REAL* r, const REAL* v, const REAL* s) {
    REAL a[3], b[3], c[3];
    cross3(a, v, s);
    cross3(b, v, s + 3);
    cross3(c, v + 3, s);
    r[0] = a[0]; r[1] = a[1]; r[2] = a[2];
    r[3] = b[0] + c[0]; r[4] = b[1] + c[1]; r[5] = b[2] + c[2];
}
static void cross_force(REAL* r, const REAL* v, const REAL* f) {
    REAL a[3], b[3], c[3];
    cross3(a, v, f);
    cross3(b, v + 3, f + 3);
    cross3(c, v, f + 3);
    r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2];
    r[3] = c[0]; r[4] = c[1]; r[5] = c[2];
}
/* dense Cholesky A = L L^T (lower), row stride NVMAX; returns 0 on success */
static int chol_factor(REAL* L, const REAL* A, int n) {
    for (int i = 0; i < n; i++)
        for (int j = 0; j <= i; j++) {
            REAL s = A[i * NVMAX + j];
            for (int k = 0; k < j; k++) s -= L[i * NVMAX + k] * L[j * NVMAX + k];
            if (i == j) {
                if (!(s > 0)) return -1;
                L[i * NVMAX + i] = rsqrt_(s);
            } else {
                L[i * NVMAX + j] = s / L[j * NVMAX + j];
            }
        }
    return 0;
}
static void chol_solve(const REAL* L, REAL* x, int n) {
    for (int i = 0; i < n; i++) {
        REAL s = x[i];
        for (int k = 0; k < i; k++) s -= L[i * NVMAX + k] * x[k];
        x[i] = s / L[i * NVMAX + i];
    }
    for (int i = n - 1; i >= 0; i--) {
        REAL s = x[i];
        for (int k = i + 1; k < n; k++) s -= L[k * NVMAX + i] * x[k];
        x[i] = s / L[i * NVMAX + i];
    }
}

/* ---------------------------------------------------------------- fwd_position */
static void kinematics(const OrcModel* m, OrcData* d) {
    int nb = m->nbody[0];
    memset(d->xpos, 0, sizeof(d->xpos));
    memset(d->xquat, 0, sizeof(d->xquat));
    d->xquat[0] = 1;
    quat_to_mat(d->xmat, d->xquat);
    for (int b = 1; b < nb; b++) {
        int p = m->body_parent[b];
        REAL pos[3], quat[4], bp[3] = {R(m->body_pos[3 * b]), R(m->body_pos[3 * b + 1]), R(m->body_pos[3 * b + 2])};
        mat_vec3(pos, d->xmat + 9 * p, bp);
        for (int k = 0; k < 3; k++) pos[k] += d->xpos[3 * p + k];
        memcpy(quat, d->xquat + 4 * p, sizeof(quat));
        for (int j = m->body_jntadr[b]; j < m->body_jntadr[b] + m->body_jntnum[b]; j++) {
            int qa = m->jnt_qposadr[j];
            REAL ax[3] = {R(m->jnt_axis[3 * j]), R(m->jnt_axis[3 * j + 1]), R(m->jnt_axis[3 * j + 2])};
            REAL jp[3] = {R(m->jnt_pos[3 * j]), R(m->jnt_pos[3 * j + 1]), R(m->jnt_pos[3 * j + 2])};
            if (m->jnt_type[j] == JNT_FREE) {
                for (int k = 0; k < 3; k++) pos[k] = d->qpos[qa + k];
                for (int k = 0; k < 4; k++) quat[k] = d->qpos[qa + 3 + k];
                quat_norm(quat);
                for (int k = 0; k < 3; k++) { d->xanchor[3 * j + k] = pos[k]; d->xaxis[3 * j + k] = (k == 2); }
            } else {
                REAL Rm[9], t[3], ql[4], qn[4];
                quat_to_mat(Rm, quat);
                mat_vec3(t, Rm, jp);
                for (int k = 0; k < 3; k++) d->xanchor[3 * j + k] = pos[k] + t[k];
                mat_vec3(d->xaxis + 3 * j, Rm, ax);
                axis_angle_quat(ql, ax, d->qpos[qa] - R(m->qpos0[qa]));
                quat_mul(qn, quat, ql);
                memcpy(quat, qn, sizeof(quat));
                quat_to_mat(Rm, quat);
                mat_vec3(t, Rm, jp);
                for (int k = 0; k < 3; k++) pos[k] = d->xanchor[3 * j + k] - t[k];
            }
        }
        quat_norm(quat);
        memcpy(d->xpos + 3 * b, pos, sizeof(pos));
        memcpy(d->xquat + 4 * b, quat, sizeof(quat));
        quat_to_mat(d->xmat + 9 * b, quat);
    }
    for (int b = 0; b < nb; b++) {
        REAL ip[3] = {R(m->body_ipos[3 * b]), R(m->body_ipos[3 * b + 1]), R(m->body_ipos[3 * b + 2])}, t[3];
        mat_vec3(t, d->xmat + 9 * b, ip);
        for (int k = 0; k < 3; k++) d->xipos[3 * b + k] = d->xpos[3 * b + k] + t[k];
        /* ximat holds the world-frame inertia tensor R I R^T (body_inertia is a full tensor) */
        const REAL* Rm = d->xmat + 9 * b;
        REAL RI[9];
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) {
                REAL s = 0;
                for (int k = 0; k < 3; k++) s += Rm[3 * r + k] * R(m->body_inertia[9 * b + 3 * k + c]);
                RI[3 * r + c] = s;
            }
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) {
                REAL s = 0;
                for (int k = 0; k < 3; k++) s += RI[3 * r + k] * Rm[3 * c + k];
                d->ximat[9 * b + 3 * r + c] = s;
            }
    }
    for (int g = 0; g < m->ngeom[0]; g++) {
        int b = m->geom_bodyid[g];
        REAL gp[3] = {R(m->geom_pos[3 * g]), R(m->geom_pos[3 * g + 1]), R(m->geom_pos[3 * g + 2])}, t[3];
        REAL gq[4] = {R(m->geom_quat[4 * g]), R(m->geom_quat[4 * g + 1]), R(m->geom_quat[4 * g + 2]), R(m->geom_quat[4 * g + 3])}, q[4];
        mat_vec3(t, d->xmat + 9 * b, gp);
        for (int k = 0; k < 3; k++) d->geom_xpos[3 * g + k] = d->xpos[3 * b + k] + t[k];
        quat_mul(q, d->xquat + 4 * b, gq);
        quat_to_mat(d->geom_xmat + 9 * g, q);
    }
}

static int dof_is_free_rot(const OrcModel* m, int i) {
    int j = -1;
    for (int k = 0; k < m->njnt[0]; k++) {
        int n = m->jnt_type[k] == JNT_FREE ? 6 : 1;
        if (i >= m->jnt_dofadr[k] && i < m->jnt_dofadr[k] + n) { j = k; break; }
    }
    return (m->jnt_type[j] == JNT_FREE) ? (i - m->jnt_dofadr[j]) : -1 - j; /* >=0: free component, <0: hinge -1-jnt */
}

static void com_pos(const OrcModel* m, OrcData* d) {
    int nb = m->nbody[0], nv = m->nv[0];
    REAL mt = 0, c[3] = {0, 0, 0};
    for (int b = 1; b < nb; b++) {
        REAL ms = R(m->body_mass[b]);
        mt += ms;
        for (int k = 0; k < 3; k++) c[k] += ms * d->xipos[3 * b + k];
    }
    for (int k = 0; k < 3; k++) { c[k] /= mt; d->subtree_com[k] = c[k]; d->com[k] = c[k]; }
    for (int b = 0; b < nb; b++) {
        REAL ms = R(m->body_mass[b]), o[3];
        const REAL* I = d->ximat + 9 * b;
        REAL* ci = d->cinert + 10 * b;
        for (int k = 0; k < 3; k++) o[k] = d->xipos[3 * b + k] - c[k];
        ci[0] = I[0] + ms * (o[1] * o[1] + o[2] * o[2]);
        ci[1] = I[4] + ms * (o[0] * o[0] + o[2] * o[2]);
        ci[2] = I[8] + ms * (o[0] * o[0] + o[1] * o[1]);
        ci[3] = I[1] - ms * o[0] * o[1];
        ci[4] = I[2] - ms * o[0] * o[2];
        ci[5] = I[5] - ms * o[1] * o[2];
        ci[6] = ms * o[0]; ci[7] = ms * o[1]; ci[8] = ms * o[2]; ci[9] = ms;
    }
    for (int i = 0; i < nv; i++) {
        int b = m->dof_bodyid[i], kind = dof_is_free_rot(m, i);
        REAL* cd = d->cdof + 6 * i;
        if (kind >= 0 && kind < 3) {
            for (int k = 0; k < 6; k++) cd[k] = 0;
            cd[3 + kind] = 1;
        } else {
            REAL ax[3], off[3];
            if (kind >= 3) {
                for (int k = 0; k < 3; k++) { ax[k] = d->xmat[9 * b + 3 * k + (kind - 3)]; off[k] = c[k] - d->xpos[3 * b + k]; }
            } else {
                int j = -1 - kind;
                for (int k = 0; k < 3; k++) { ax[k] = d->xaxis[3 * j + k]; off[k] = c[k] - d->xanchor[3 * j + k]; }
            }
            cd[0] = ax[0]; cd[1] = ax[1]; cd[2] = ax[2];
            cross3(cd + 3, ax, off);
        }
    }
}

static void crb(const OrcModel* m, OrcData* d) {
    int nb = m->nbody[0], nv = m->nv[0];
    memcpy(d->crb, d->cinert, sizeof(d->crb));
    for (int b = nb - 1; b > 0; b--) {
        int p = m->body_parent[b];
        if (p > 0) for (int k = 0; k < 10; k++) d->crb[10 * p + k] += d->crb[10 * b + k];
    }
    memset(d->qM, 0, sizeof(d->qM));
    for (int i = 0; i < nv; i++) {
        REAL buf[6];
        mul_inert_vec(buf, d->crb + 10 * m->dof_bodyid[i], d->cdof + 6 * i);
        for (int j = i; j >= 0; j = m->dof_parentid[j]) {
            REAL v = dotn(d->cdof + 6 * j, buf, 6);
            d->qM[i * NVMAX + j] = v;
            d->qM[j * NVMAX + i] = v;
        }
        d->qM[i * NVMAX + i] += R(m->dof_armature[i]);
    }
    chol_factor(d->qL, d->qM, nv);
}

/* ---------------------------------------------------------------- collision (geom vs floor plane) */
static void make_frame(REAL* fr, const REAL* a) { /* rows: a, b, a x b  (mjx math.make_frame) */
    REAL b[3] = {0, 0, 0}, dt, n;
    if (a[1] > R(-0.5) && a[1] < R(0.5)) b[1] = 1; else b[2] = 1;
    dt = dot3(a, b);
    for (int k = 0; k < 3; k++) b[k] -= a[k] * dt;
    n = rsqrt_(dot3(b, b));
    for (int k = 0; k < 3; k++) { b[k] /= n; fr[k] = a[k]; fr[3 + k] = b[k]; }
    cross3(fr + 6, a, b);
}
static void add_contact2(OrcData* d, int* nc, REAL dist, const REAL* pos, const REAL* frame, int geom1, int geom, REAL mu, REAL margin) {
    int c = (*nc)++;
    d->con_geom1[c] = (REAL)geom1;
    d->con_dist[c] = dist;
    memcpy(d->con_pos + 3 * c, pos, 3 * sizeof(REAL));
    memcpy(d->con_frame + 9 * c, frame, 9 * sizeof(REAL));
    d->con_geom[c] = (REAL)geom;
    d->con_mu[c] = mu;
    d->con_includemargin[c] = margin;
}
/* ---- sphere / capsule pairs (mjx collision_primitive.py: _sphere_sphere, sphere_capsule, capsule_capsule; math.py:
 *      normalize_with_norm, closest_segment_point, closest_segment_to_segment_points) */
static void closest_segment_point(REAL* out, const REAL* a, const REAL* b, const REAL* pt) {
    REAL ab[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, ap[3] = {pt[0] - a[0], pt[1] - a[1], pt[2] - a[2]};
    REAL t = dot3(ap, ab) / (dot3(ab, ab) + R(1e-6));
    t = t < 0 ? 0 : (t > 1 ? 1 : t);
    for (int k = 0; k < 3; k++) out[k] = a[k] + t * ab[k];
}
static REAL normalize_with_norm(REAL* n, const REAL* x) {
    REAL nrm = rsqrt_(dot3(x, x)), den = nrm + (nrm == 0 ? R(1e-6) : 0);
    for (int k = 0; k < 3; k++) n[k] = x[k] / den;
    return nrm;
}
static void closest_segment_to_segment(REAL* best_a, REAL* best_b, const REAL* a0, const REAL* a1, const REAL* b0, const REAL* b1) {
    REAL va[3], vb[3], dir_a[3], dir_b[3], a_mid[3], b_mid[3], trans[3], new_a[3], new_b[3], e1[3], e2[3];
    for (int k = 0; k < 3; k++) { va[k] = a1[k] - a0[k]; vb[k] = b1[k] - b0[k]; }
    REAL half_a = normalize_with_norm(dir_a, va) * R(0.5), half_b = normalize_with_norm(dir_b, vb) * R(0.5);
    for (int k = 0; k < 3; k++) { a_mid[k] = a0[k] + dir_a[k] * half_a; b_mid[k] = b0[k] + dir_b[k] * half_b; trans[k] = a_mid[k] - b_mid[k]; }
    REAL dd = dot3(dir_a, dir_b), dat = dot3(dir_a, trans), dbt = dot3(dir_b, trans), denom = 1 - dd * dd;
    REAL ta = (-dat + dd * dbt) / (denom + R(1e-6)), tb = dbt + ta * dd;
    ta = ta < -half_a ? -half_a : (ta > half_a ? half_a : ta);
    tb = tb < -half_b ? -half_b : (tb > half_b ? half_b : tb);
    for (int k = 0; k < 3; k++) { best_a[k] = a_mid[k] + dir_a[k] * ta; best_b[k] = b_mid[k] + dir_b[k] * tb; }
    closest_segment_point(new_a, a0, a1, best_b);
    closest_segment_point(new_b, b0, b1, best_a);
    for (int k = 0; k < 3; k++) { e1[k] = new_a[k] - best_b[k]; e2[k] = new_b[k] - best_a[k]; }
    if (dot3(e1, e1) < dot3(e2, e2)) memcpy(best_a, new_a, 3 * sizeof(REAL)); else memcpy(best_b, new_b, 3 * sizeof(REAL));
}
static void sphere_sphere(const REAL* p1, REAL r1, const REAL* p2, REAL r2, REAL* dist, REAL* pos, REAL* frame) {
    REAL dlt[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]}, n[3];
    REAL nrm = normalize_with_norm(n, dlt);
    if (nrm == 0) { n[0] = 1; n[1] = 0; n[2] = 0; }
    *dist = nrm - (r1 + r2);
    for (int k = 0; k < 3; k++) pos[k] = p1[k] + n[k] * (r1 + *dist * R(0.5));
    make_frame(frame, n);
}
/* sphere vs box: mjx collision_convex.py _sphere_convex on the box's six faces -- among the faces whose plane the sphere's
 * deepest point is behind (support < 0) take the one of least penetration, then the closest point of that face polygon to
 * the sphere centre (projection, or the nearest edge / corner when the projection falls outside) [M] */
static void sphere_box(const REAL* sp, REAL r, const REAL* bp, const REAL* bm, const double* half, REAL* dist, REAL* pos, REAL* frame) {
    REAL dlt[3] = {sp[0] - bp[0], sp[1] - bp[1], sp[2] - bp[2]}, p[3], pt[3], e[3], n[3], nw[3], pl[3], best = R(-1e12);
    int bk = 0, bs = 1;
    matT_vec3(p, bm, dlt);
    for (int k = 0; k < 3; k++)
        for (int sg = 1; sg >= -1; sg -= 2) {
            REAL sup = (REAL)sg * p[k] - R(half[k]) - r;
            if (sup >= 0) sup = R(-1e12);
            if (sup > best) { best = sup; bk = k; bs = sg; }
        }
    for (int k = 0; k < 3; k++) {
        REAL h = R(half[k]);
        pt[k] = p[k] < -h ? -h : (p[k] > h ? h : p[k]);
    }
    pt[bk] = (REAL)bs * R(half[bk]);
    for (int k = 0; k < 3; k++) e[k] = pt[k] - p[k];
    REAL nrm = normalize_with_norm(n, e);
    *dist = nrm - r;
    for (int k = 0; k < 3; k++) pl[k] = R(0.5) * (pt[k] + p[k] + n[k] * r);
    mat_vec3(nw, bm, n);
    mat_vec3(pos, bm, pl);
    for (int k = 0; k < 3; k++) pos[k] += bp[k];
    make_frame(frame, nw);
}
static int argmax_(const REAL* v, int n) { int b = 0; for (int i = 1; i < n; i++) if (v[i] > v[b]) b = i; return b; }
/* mjx _manifold_points adds -1e6 to the score of a masked-out candidate, in jax's default float32 (the reference never enables
 * x64): the sum lands on float32's grid at 1e6 (spacing 0.0625).  That rounding decides the picks when EVERY candidate is masked
 * out (a foot box hovering inside the margin without penetrating): scores below 0.03125 collapse onto -1e6 and the first index
 * wins.  The checker reproduces that one rounding whatever REAL is. */
static REAL mask_add(REAL s, int ok) { return ok ? s : (REAL)((float)s + -1e6f); }
/* ---- capsule / box against a box (mjx collision_convex.py: capsule_convex, convex_convex through _create_contact_manifold,
 *      _clip, _clip_edge_to_planes, _closest_segment_point_plane, _project_pt_onto_plane, _manifold_points) [M: restated from the
 *      published algorithm, mujoco-mjx 3.1.x; no golden exists for these rows -- parity unpinned].  Conventions of this
 *      restatement where the source could not be consulted: box faces are listed +x, -x, +y, -y, +z, -z with their corners
 *      counter-clockwise seen from outside; the first maximum / minimum wins every argmax / argmin; contact positions lie
 *      midway between the two surfaces; manifold picks that coincide with an earlier pick (within 10 microns) are switched off
 *      (dist = 1), as plane_convex does for repeated indices. */
static void box_face(const double* half, int f, REAL v[4][3], REAL* n) {
    int k = f / 2, a = (k + 1) % 3, b = (k + 2) % 3;
    REAL sg = (f & 1) ? R(-1) : R(1);
    static const int ccw[4][2] = {{-1, -1}, {1, -1}, {1, 1}, {-1, 1}};
    for (int i = 0; i < 4; i++) {
        int j = (f & 1) ? (4 - i) % 4 : i; /* reversed order for the negative face: still counter-clockwise from outside */
        v[i][k] = sg * R(half[k]); v[i][a] = (REAL)ccw[j][0] * R(half[a]); v[i][b] = (REAL)ccw[j][1] * R(half[b]);
    }
    n[0] = n[1] = n[2] = 0; n[k] = sg;
}
static void project_pt_onto_plane(REAL* out, const REAL* pt, const REAL* plane_pt, const REAL* n) {
    REAL t[3] = {pt[0] - plane_pt[0], pt[1] - plane_pt[1], pt[2] - plane_pt[2]}, dd = dot3(t, n);
    for (int k = 0; k < 3; k++) out[k] = pt[k] - dd * n[k];
}
static void closest_segment_point_plane(REAL* out, const REAL* a, const REAL* b, const REAL* p0, const REAL* n) {
    REAL ab[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, dd = dot3(p0, n), denom = dot3(n, ab);
    REAL t = (dd - dot3(n, a)) / (denom + (denom == 0 ? R(1e-6) : 0));
    t = t < 0 ? 0 : (t > 1 ? 1 : t);
    for (int k = 0; k < 3; k++) out[k] = a[k] + t * ab[k];
}
/* clips the edge (e0, e1) against `np` planes (point, outward normal); returns the mask of the clipped edge */
static int clip_edge_to_planes(const REAL* e0, const REAL* e1, REAL (*ppt)[3], REAL (*pn)[3], int np, REAL out[2][3]) {
    REAL cand[8][3], best0 = 0, best1 = 0, n0[3], n1[3], e01[3] = {e1[0] - e0[0], e1[1] - e0[1], e1[2] - e0[2]}, e10[3] = {-e01[0], -e01[1], -e01[2]};
    int in0[8], in1[8], both = 0;
    for (int k = 0; k < np; k++) {
        REAL t0[3] = {e0[0] - ppt[k][0], e0[1] - ppt[k][1], e0[2] - ppt[k][2]}, t1[3] = {e1[0] - ppt[k][0], e1[1] - ppt[k][1], e1[2] - ppt[k][2]};
        in0[k] = dot3(t0, pn[k]) > R(1e-6);
        in1[k] = dot3(t1, pn[k]) > R(1e-6);
        both |= in0[k] && in1[k];
        closest_segment_point_plane(cand[k], e0, e1, ppt[k], pn[k]);
    }
    for (int k = 0; k < np; k++) { /* the clipped point that lies farthest along the edge direction (the end point itself if nothing clips) */
        const REAL* q0 = in0[k] ? cand[k] : e0;
        const REAL* q1 = in1[k] ? cand[k] : e1;
        REAL t0[3] = {q0[0] - e0[0], q0[1] - e0[1], q0[2] - e0[2]}, t1[3] = {q1[0] - e1[0], q1[1] - e1[1], q1[2] - e1[2]};
        REAL d0 = dot3(t0, e01), d1 = dot3(t1, e10);
        if (k == 0 || d0 > best0) { best0 = d0; memcpy(n0, q0, sizeof(n0)); }
        if (k == 0 || d1 > best1) { best1 = d1; memcpy(n1, q1, sizeof(n1)); }
    }
    int mask = !both;
    if (!mask) { memcpy(n0, e0, sizeof(n0)); memcpy(n1, e1, sizeof(n1)); }
    memcpy(out[0], n0, sizeof(n0));
    memcpy(out[1], n1, sizeof(n1));
    REAL dn[3] = {n0[0] - n1[0], n0[1] - n1[1], n0[2] - n1[2]};
    if (dot3(e10, dn) < 0) mask = 0; /* the clipped end points crossed */
    return mask;
}
/* mjx _manifold_points: four well-spread points among the masked candidates */
static void manifold_points(REAL (*pt)[3], const int* mask, int n, const REAL* nrm, int idx[4]) {
    REAL sc[32], ab[3], ac[3], bc[3], e[3], t[3];
    for (int i = 0; i < n; i++) sc[i] = mask[i] ? 0 : R(-1e6);
    int a = argmax_(sc, n);
    for (int i = 0; i < n; i++) { for (int k = 0; k < 3; k++) t[k] = pt[a][k] - pt[i][k]; sc[i] = mask_add(dot3(t, t), mask[i]); }
    int b = argmax_(sc, n);
    for (int k = 0; k < 3; k++) e[k] = pt[a][k] - pt[b][k];
    cross3(ab, nrm, e);
    for (int i = 0; i < n; i++) { for (int k = 0; k < 3; k++) t[k] = pt[a][k] - pt[i][k]; sc[i] = mask_add(R(fabs((double)dot3(t, ab))), mask[i]); }
    int c = argmax_(sc, n);
    for (int k = 0; k < 3; k++) e[k] = pt[a][k] - pt[c][k];
    cross3(ac, nrm, e);
    for (int k = 0; k < 3; k++) e[k] = pt[b][k] - pt[c][k];
    cross3(bc, nrm, e);
    for (int i = 0; i < n; i++) {
        for (int k = 0; k < 3; k++) t[k] = pt[b][k] - pt[i][k];
        sc[i] = mask_add(R(fabs((double)dot3(t, bc))), mask[i]);
    }
    int d1 = argmax_(sc, n);
    REAL s1 = sc[d1];
    for (int i = 0; i < n; i++) {
        for (int k = 0; k < 3; k++) t[k] = pt[a][k] - pt[i][k];
        sc[i] = mask_add(R(fabs((double)dot3(t, ac))), mask[i]);
    }
    int d2 = argmax_(sc, n);
    idx[0] = a; idx[1] = b; idx[2] = c; idx[3] = sc[d2] > s1 ? d2 : d1;
}
/* capsule (world centre cpos, frame cmat, radius r, half length hl) against a box: two contacts, normal from the capsule to the box */
static void capsule_box(const REAL* cpos, const REAL* cmat, REAL r, REAL hl, const REAL* bpos, const REAL* bmat, const double* half,
                        REAL dist[2], REAL pos[2][3], REAL nrm[2][3]) {
    REAL dl[3] = {cpos[0] - bpos[0], cpos[1] - bpos[1], cpos[2] - bpos[2]}, axw[3] = {cmat[2], cmat[5], cmat[8]}, cp[3], ax[3], p0[3], p1[3];
    matT_vec3(cp, bmat, dl);
    matT_vec3(ax, bmat, axw);
    for (int k = 0; k < 3; k++) { p0[k] = cp[k] - ax[k] * hl; p1[k] = cp[k] + ax[k] * hl; }
    /* the face of least penetration among those the (inflated) segment reaches behind */
    REAL best = 0, fv[4][3], fn[3];
    int bf = 0, separated = 0; /* a face whose plane the inflated segment does not reach behind separates the two geoms: no face contact */
    for (int f = 0; f < 6; f++) {
        box_face(half, f, fv, fn);
        REAL a[3] = {p0[0] - fn[0] * r - fv[0][0], p0[1] - fn[1] * r - fv[0][1], p0[2] - fn[2] * r - fv[0][2]};
        REAL b[3] = {p1[0] - fn[0] * r - fv[0][0], p1[1] - fn[1] * r - fv[0][1], p1[2] - fn[2] * r - fv[0][2]};
        REAL sa = dot3(a, fn), sb = dot3(b, fn), sup = sa < sb ? sa : sb;
        if (sup >= 0) { sup = R(-1e12); separated = 1; }
        if (f == 0 || sup > best) { best = sup; bf = f; }
    }
    box_face(half, bf, fv, fn);
    REAL ep0[4][3], ep1[4][3], en[4][3], cl[2][3], lp[2][3], ln[2][3], pen[2];
    for (int i = 0; i < 4; i++) {
        REAL e[3];
        memcpy(ep0[i], fv[(i + 3) % 4], sizeof(ep0[i]));
        memcpy(ep1[i], fv[i], sizeof(ep1[i]));
        for (int k = 0; k < 3; k++) e[k] = ep1[i][k] - ep0[i][k];
        cross3(en[i], e, fn);
    }
    int mask = clip_edge_to_planes(p0, p1, ep0, en, 4, cl) && !separated;
    for (int i = 0; i < 2; i++) {
        REAL fp[3], t[3];
        for (int k = 0; k < 3; k++) cl[i][k] -= fn[k] * r;
        project_pt_onto_plane(fp, cl[i], fv[0], fn);
        for (int k = 0; k < 3; k++) { lp[i][k] = R(0.5) * (cl[i][k] + fp[k]); ln[i][k] = -fn[k]; t[k] = fp[k] - cl[i][k]; }
        pen[i] = mask ? dot3(t, fn) : R(-1);
    }
    /* a potential edge contact: the edge of that face closest to the capsule's segment */
    REAL bd = 0, ec[3], cc[3];
    for (int i = 0; i < 4; i++) {
        REAL e_[3], c_[3], t[3];
        closest_segment_to_segment(e_, c_, ep0[i], ep1[i], p0, p1);
        for (int k = 0; k < 3; k++) t[k] = e_[k] - c_[k];
        REAL dd = dot3(t, t);
        if (i == 0 || dd < bd) { bd = dd; memcpy(ec, e_, sizeof(ec)); memcpy(cc, c_, sizeof(cc)); }
    }
    REAL ea[3] = {cc[0] - ec[0], cc[1] - ec[1], cc[2] - ec[2]}, eax[3];
    REAL ed = normalize_with_norm(eax, ea);
    if (r - ed > 0 && ed > 0) {
        for (int k = 0; k < 3; k++) { lp[0][k] = R(0.5) * (ec[k] + (cc[k] - eax[k] * r)); ln[0][k] = -eax[k]; }
        pen[0] = r - ed;
        pen[1] = R(-1);
    }
    for (int i = 0; i < 2; i++) {
        dist[i] = -pen[i];
        mat_vec3(pos[i], bmat, lp[i]);
        for (int k = 0; k < 3; k++) pos[i][k] += bpos[k];
        mat_vec3(nrm[i], bmat, ln[i]);
    }
}
/* box against box: separating-axis test over the 6 face normals and the 9 edge cross products, then the face manifold
 * (reference face clipped against the incident one, four well-spread points) or the single edge-edge contact.
 * Up to four contacts, normal from box 1 to box 2; returns the number of contact slots filled (always 4). */
static void box_box(const REAL* c1, const REAL* m1, const double* h1, const REAL* c2, const REAL* m2, const double* h2,
                    REAL dist[4], REAL pos[4][3], REAL nrm[4][3]) {
    REAL ax[15][3], bestd = 0, bestax[3] = {1, 0, 0}, dc[3] = {c2[0] - c1[0], c2[1] - c1[1], c2[2] - c1[2]};
    int besti = -1;
    for (int i = 0; i < 3; i++) for (int k = 0; k < 3; k++) { ax[i][k] = m1[3 * k + i]; ax[3 + i][k] = m2[3 * k + i]; }
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) cross3(ax[6 + 3 * i + j], ax[i], ax[3 + j]);
    for (int a = 0; a < 15; a++) {
        REAL L[3], nn = rsqrt_(dot3(ax[a], ax[a]));
        int degenerate = nn < R(1e-6);
        for (int k = 0; k < 3; k++) L[k] = degenerate ? ax[a][k] : ax[a][k] / nn;
        REAL r1 = 0, r2 = 0;
        for (int i = 0; i < 3; i++) { r1 += R(h1[i]) * R(fabs((double)dot3(ax[i], L))); r2 += R(h2[i]) * R(fabs((double)dot3(ax[3 + i], L))); }
        REAL pc = dot3(dc, L), d1 = r1 + r2 - pc, d2 = r1 + r2 + pc; /* overlap if box 2 lies on the +L / -L side of box 1 */
        REAL dd = d1 < d2 ? d1 : d2, sg = d1 > d2 ? R(-1) : R(1);
        if (degenerate) dd = R(1e6);
        if (besti < 0 || dd < bestd) { bestd = dd; besti = a; for (int k = 0; k < 3; k++) bestax[k] = sg * L[k]; }
    }
    for (int q = 0; q < 4; q++) { dist[q] = 1; for (int k = 0; k < 3; k++) { pos[q][k] = R(0.5) * (c1[k] + c2[k]); nrm[q][k] = bestax[k]; } }
    if (besti >= 6) { /* edge-edge: the edge of each box that faces the other one, closest points between them */
        int i = (besti - 6) / 3, j = (besti - 6) % 3;
        REAL a0[3], a1[3], b0[3], b1[3], pa[3], pb[3], t[3];
        for (int k = 0; k < 3; k++) { a0[k] = c1[k]; b0[k] = c2[k]; }
        for (int q = 0; q < 3; q++) {
            if (q != i) { REAL s = dot3(ax[q], bestax) > 0 ? R(1) : R(-1); for (int k = 0; k < 3; k++) a0[k] += s * R(h1[q]) * ax[q][k]; }
            if (q != j) { REAL s = dot3(ax[3 + q], bestax) > 0 ? R(-1) : R(1); for (int k = 0; k < 3; k++) b0[k] += s * R(h2[q]) * ax[3 + q][k]; }
        }
        for (int k = 0; k < 3; k++) { a1[k] = a0[k] + R(h1[i]) * ax[i][k]; a0[k] -= R(h1[i]) * ax[i][k]; b1[k] = b0[k] + R(h2[j]) * ax[3 + j][k]; b0[k] -= R(h2[j]) * ax[3 + j][k]; }
        closest_segment_to_segment(pa, pb, a0, a1, b0, b1);
        for (int k = 0; k < 3; k++) { t[k] = pb[k] - pa[k]; pos[0][k] = R(0.5) * (pa[k] + pb[k]); }
        dist[0] = dot3(t, bestax);
        return;
    }
    /* face contact: reference face = the face most aligned with the axis (of box 1 towards +axis, of box 2 towards -axis) */
    REAL amax = 0, bmin = 0;
    int af = 0, bfc = 0;
    for (int f = 0; f < 6; f++) {
        REAL sg = (f & 1) ? R(-1) : R(1), da = sg * dot3(ax[f / 2], bestax), db = sg * dot3(ax[3 + f / 2], bestax);
        if (f == 0 || da > amax) { amax = da; af = f; }
        if (f == 0 || db < bmin) { bmin = db; bfc = f; }
    }
    int ref1 = amax >= -bmin; /* 1: the reference face belongs to box 1 */
    const REAL *rc = ref1 ? c1 : c2, *rm = ref1 ? m1 : m2, *ic = ref1 ? c2 : c1, *im = ref1 ? m2 : m1;
    const double *rh = ref1 ? h1 : h2, *ih = ref1 ? h2 : h1;
    REAL rv[4][3], rn[3], iv[4][3], in_[3], lv[4][3], ln[3];
    box_face(rh, ref1 ? af : bfc, lv, ln);
    for (int q = 0; q < 4; q++) { mat_vec3(rv[q], rm, lv[q]); for (int k = 0; k < 3; k++) rv[q][k] += rc[k]; }
    mat_vec3(rn, rm, ln);
    box_face(ih, ref1 ? bfc : af, lv, ln);
    for (int q = 0; q < 4; q++) { mat_vec3(iv[q], im, lv[q]); for (int k = 0; k < 3; k++) iv[q][k] += ic[k]; }
    mat_vec3(in_, im, ln);
    /* mjx _clip: incident edges against the reference side planes, then the (projected) reference edges against the incident side planes */
    REAL rp0[4][3], rpn[4][3], ip0[4][3], ipn[4][3], cand[16][3], out2[2][3], e[3];
    int cmask[16];
    for (int q = 0; q < 4; q++) {
        memcpy(rp0[q], rv[(q + 3) % 4], sizeof(rp0[q]));
        for (int k = 0; k < 3; k++) e[k] = rv[q][k] - rp0[q][k];
        cross3(rpn[q], e, rn);
        memcpy(ip0[q], iv[(q + 3) % 4], sizeof(ip0[q]));
        for (int k = 0; k < 3; k++) e[k] = iv[q][k] - ip0[q][k];
        cross3(ipn[q], e, in_);
    }
    for (int q = 0; q < 4; q++) {
        int mk = clip_edge_to_planes(ip0[q], iv[q], rp0, rpn, 4, out2);
        memcpy(cand[2 * q], out2[0], sizeof(out2[0])); memcpy(cand[2 * q + 1], out2[1], sizeof(out2[1]));
        cmask[2 * q] = cmask[2 * q + 1] = mk;
    }
    REAL rproj[4][3], dpl = dot3(iv[0], in_), den = dot3(rn, in_);
    for (int q = 0; q < 4; q++) { /* reference corners moved onto the incident plane along the reference normal */
        REAL t = (dpl - dot3(rv[q], in_)) / (den + (den == 0 ? R(1e-6) : 0));
        for (int k = 0; k < 3; k++) rproj[q][k] = rv[q][k] + t * rn[k];
    }
    for (int q = 0; q < 4; q++) {
        int mk = clip_edge_to_planes(rproj[(q + 3) % 4], rproj[q], ip0, ipn, 4, out2);
        memcpy(cand[8 + 2 * q], out2[0], sizeof(out2[0])); memcpy(cand[8 + 2 * q + 1], out2[1], sizeof(out2[1]));
        cmask[8 + 2 * q] = cmask[8 + 2 * q + 1] = mk;
    }
    REAL onref[16][3];
    for (int q = 0; q < 16; q++) {
        REAL t[3] = {cand[q][0] - rv[0][0], cand[q][1] - rv[0][1], cand[q][2] - rv[0][2]};
        project_pt_onto_plane(onref[q], cand[q], rv[0], rn);
        cmask[q] = cmask[q] && (-dot3(t, rn) > R(1e-6)); /* behind the reference face */
    }
    int idx[4];
    manifold_points(onref, cmask, 16, rn, idx);
    for (int q = 0; q < 4; q++) {
        int unique = 1, p = idx[q];
        for (int r2 = 0; r2 < q; r2++) { /* a pick that coincides with an earlier one (same candidate, or the shared corner of two clipped edges) is switched off */
            REAL u[3] = {onref[p][0] - onref[idx[r2]][0], onref[p][1] - onref[idx[r2]][1], onref[p][2] - onref[idx[r2]][2]};
            if (dot3(u, u) < R(1e-10)) unique = 0;
        }
        REAL t[3] = {cand[p][0] - onref[p][0], cand[p][1] - onref[p][1], cand[p][2] - onref[p][2]};
        REAL pen = -dot3(t, rn);
        dist[q] = (cmask[p] && unique) ? -pen : R(1);
        for (int k = 0; k < 3; k++) pos[q][k] = R(0.5) * (cand[p][k] + onref[p][k]);
    }
}
static void capsule_ends(const OrcModel* m, const OrcData* d, int g, REAL* e0, REAL* e1) {
    const REAL* gm = d->geom_xmat + 9 * g;
    for (int k = 0; k < 3; k++) {
        REAL s = gm[3 * k + 2] * R(m->geom_size[3 * g + 1]);
        e0[k] = d->geom_xpos[3 * g + k] - s; e1[k] = d->geom_xpos[3 * g + k] + s;
    }
}
static void plane_sphere(const REAL* n, const REAL* ppos, const REAL* c, REAL r, REAL* dist, REAL* pos) {
    REAL t[3] = {c[0] - ppos[0], c[1] - ppos[1], c[2] - ppos[2]};
    *dist = dot3(t, n) - r;
    for (int k = 0; k < 3; k++) pos[k] = c[k] - n[k] * (r + R(0.5) * (*dist));
}

static void collision(const OrcModel* m, OrcData* d) {
    int nc = 0, ng = m->ngeom[0];
    /* the floor is the only plane and the only world geom; every other geom collides with it */
    for (int pass = 0; pass < 3; pass++) { /* grouped per collision function like MJX: sphere, capsule, convex */
        for (int p = 0; p < ng; p++) {
            if (m->geom_type[p] != GEOM_PLANE) continue;
            const REAL* pm = d->geom_xmat + 9 * p;
            REAL n[3] = {pm[2], pm[5], pm[8]};
            const REAL* ppos = d->geom_xpos + 3 * p;
            for (int g = 0; g < ng; g++) {
                int t = m->geom_type[g];
                if (g == p || m->geom_bodyid[g] == m->geom_bodyid[p]) continue;
                if ((pass == 0 && t != GEOM_SPHERE) || (pass == 1 && t != GEOM_CAPSULE) || (pass == 2 && t != GEOM_BOX)) continue;
                REAL margin = R(fmax(m->geom_margin[p], m->geom_margin[g]) - fmax(m->geom_gap[p], m->geom_gap[g]));
                REAL mu = R(fmax(m->geom_friction[3 * p], m->geom_friction[3 * g]));
                const REAL* gp = d->geom_xpos + 3 * g;
                const REAL* gm = d->geom_xmat + 9 * g;
                REAL fr[9], dist, pos[3];
                if (t == GEOM_SPHERE) {
                    make_frame(fr, n);
                    plane_sphere(n, ppos, gp, R(m->geom_size[3 * g]), &dist, pos);
                    add_contact2(d, &nc, dist, pos, fr, p, g, mu, margin);
                } else if (t == GEOM_CAPSULE) {
                    REAL axis[3] = {gm[2], gm[5], gm[8]}, b[3], bn, na = dot3(n, axis);
                    for (int k = 0; k < 3; k++) b[k] = axis[k] - n[k] * na;
                    bn = rsqrt_(dot3(b, b));
                    if (bn < R(0.5)) {
                        b[0] = 0; b[1] = 0; b[2] = 0;
                        if (n[1] > R(-0.5) && n[1] < R(0.5)) b[1] = 1; else b[2] = 1;
                    } else {
                        for (int k = 0; k < 3; k++) b[k] /= bn;
                    }
                    for (int k = 0; k < 3; k++) { fr[k] = n[k]; fr[3 + k] = b[k]; }
                    cross3(fr + 6, n, b);
                    for (int s = 0; s < 2; s++) {
                        REAL sg = s == 0 ? R(1) : R(-1), c[3];
                        for (int k = 0; k < 3; k++) c[k] = gp[k] + sg * axis[k] * R(m->geom_size[3 * g + 1]);
                        plane_sphere(n, ppos, c, R(m->geom_size[3 * g]), &dist, pos);
                        add_contact2(d, &nc, dist, pos, fr, p, g, mu, margin);
                    }
                } else if (t == GEOM_BOX) {
                    /* mjx plane_convex + _manifold_points on the 8 box corners */
                    REAL vert[8][3], support[8], dmask[8], score[16], pl[3], nl[3], tmp[3], smax;
                    int idx[4];
                    for (int v = 0; v < 8; v++)
                        for (int k = 0; k < 3; k++) vert[v][k] = ((v >> (2 - k)) & 1 ? R(1) : R(-1)) * R(m->geom_size[3 * g + k]);
                    for (int k = 0; k < 3; k++) tmp[k] = ppos[k] - gp[k];
                    matT_vec3(pl, gm, tmp);
                    matT_vec3(nl, gm, n);
                    for (int v = 0; v < 8; v++) {
                        for (int k = 0; k < 3; k++) tmp[k] = pl[k] - vert[v][k];
                        support[v] = dot3(tmp, nl);
                    }
                    smax = support[argmax_(support, 8)];
                    for (int v = 0; v < 8; v++) dmask[v] = support[v] > R(fmax(0.0, (double)smax - 1e-3)) ? R(0) : R(-1e6);
                    int ia = argmax_(dmask, 8);
                    for (int v = 0; v < 8; v++) {
                        REAL s = 0;
                        for (int k = 0; k < 3; k++) s += (vert[ia][k] - vert[v][k]) * (vert[ia][k] - vert[v][k]);
                        score[v] = mask_add(s, dmask[v] == 0);
                    }
                    int ib = argmax_(score, 8);
                    REAL ab[3], ac[3], bc[3], e[3];
                    for (int k = 0; k < 3; k++) e[k] = vert[ia][k] - vert[ib][k];
                    cross3(ab, nl, e);
                    for (int v = 0; v < 8; v++) {
                        for (int k = 0; k < 3; k++) tmp[k] = vert[ia][k] - vert[v][k];
                        score[v] = mask_add(R(fabs((double)dot3(tmp, ab))), dmask[v] == 0);
                    }
                    int ic = argmax_(score, 8);
                    for (int k = 0; k < 3; k++) e[k] = vert[ia][k] - vert[ic][k];
                    cross3(ac, nl, e);
                    for (int k = 0; k < 3; k++) e[k] = vert[ib][k] - vert[ic][k];
                    cross3(bc, nl, e);
                    for (int v = 0; v < 8; v++) {
                        for (int k = 0; k < 3; k++) tmp[k] = vert[ib][k] - vert[v][k];
                        score[v] = mask_add(R(fabs((double)dot3(tmp, bc))), dmask[v] == 0);
                        for (int k = 0; k < 3; k++) tmp[k] = vert[ia][k] - vert[v][k];
                        score[8 + v] = mask_add(R(fabs((double)dot3(tmp, ac))), dmask[v] == 0);
                    }
                    int id = argmax_(score, 16) % 8;
                    idx[0] = ia; idx[1] = ib; idx[2] = ic; idx[3] = id;
                    make_frame(fr, n);
                    for (int q = 0; q < 4; q++) {
                        int unique = 1;
                        for (int r2 = 0; r2 < q; r2++) if (idx[r2] == idx[q]) unique = 0;
                        dist = unique ? -support[idx[q]] : R(1);
                        mat_vec3(tmp, gm, vert[idx[q]]);
                        for (int k = 0; k < 3; k++) pos[k] = gp[k] + tmp[k] - R(0.5) * dist * n[k];
                        add_contact2(d, &nc, dist, pos, fr, p, g, mu, margin);
                    }
                }
            }
        }
    }
    /* self-collision candidates.  MJX evaluates every candidate every step and multiplies
     * the rows of the distant ones by zero; only the contacts inside the margin get rows here (same optimum).  The
     * capsule-box and box-box pairs (mjx collision_convex.py capsule_convex / convex_convex) are not restated yet: they are counted. */
    int nplane = nc, nbox = 0;
    for (int q = 0; q < m->npair[0]; q++) {
        int g1 = m->pair_geom1[q], g2 = m->pair_geom2[q], t1 = m->geom_type[g1], t2 = m->geom_type[g2];
        if (t1 == GEOM_PLANE) continue;
        if (t2 == GEOM_BOX && t1 != GEOM_SPHERE) { /* capsule-box: 2 contact slots, box-box: 4 */
            REAL margin2 = R(fmax(m->geom_margin[g1], m->geom_margin[g2]) - fmax(m->geom_gap[g1], m->geom_gap[g2]));
            REAL mu2 = R(fmax(m->geom_friction[3 * g1], m->geom_friction[3 * g2])), dd[4], pp[4][3], nn[4][3], fr2[9];
            int nslot = t1 == GEOM_CAPSULE ? 2 : 4;
            if (t1 == GEOM_CAPSULE)
                capsule_box(d->geom_xpos + 3 * g1, d->geom_xmat + 9 * g1, R(m->geom_size[3 * g1]), R(m->geom_size[3 * g1 + 1]), d->geom_xpos + 3 * g2,
                            d->geom_xmat + 9 * g2, m->geom_size + 3 * g2, dd, pp, nn);
            else
                box_box(d->geom_xpos + 3 * g1, d->geom_xmat + 9 * g1, m->geom_size + 3 * g1, d->geom_xpos + 3 * g2, d->geom_xmat + 9 * g2,
                        m->geom_size + 3 * g2, dd, pp, nn);
            for (int q2 = 0; q2 < nslot; q2++)
                if (dd[q2] - margin2 < 0 && nc < ORC_MAXCON) { make_frame(fr2, nn[q2]); add_contact2(d, &nc, dd[q2], pp[q2], fr2, g1, g2, mu2, margin2); }
            continue;
        }
        REAL margin = R(fmax(m->geom_margin[g1], m->geom_margin[g2]) - fmax(m->geom_gap[g1], m->geom_gap[g2]));
        REAL mu = R(fmax(m->geom_friction[3 * g1], m->geom_friction[3 * g2]));
        REAL r1 = R(m->geom_size[3 * g1]), r2 = R(m->geom_size[3 * g2]), fr[9], dist, pos[3], a0[3], a1[3], b0[3], b1[3], pa[3], pb[3];
        if (t2 == GEOM_BOX) {
            sphere_box(d->geom_xpos + 3 * g1, r1, d->geom_xpos + 3 * g2, d->geom_xmat + 9 * g2, m->geom_size + 3 * g2, &dist, pos, fr);
        } else if (t1 == GEOM_SPHERE && t2 == GEOM_SPHERE) {
            sphere_sphere(d->geom_xpos + 3 * g1, r1, d->geom_xpos + 3 * g2, r2, &dist, pos, fr);
        } else if (t1 == GEOM_SPHERE && t2 == GEOM_CAPSULE) {
            capsule_ends(m, d, g2, b0, b1);
            closest_segment_point(pb, b0, b1, d->geom_xpos + 3 * g1);
            sphere_sphere(d->geom_xpos + 3 * g1, r1, pb, r2, &dist, pos, fr);
        } else {
            capsule_ends(m, d, g1, a0, a1);
            capsule_ends(m, d, g2, b0, b1);
            closest_segment_to_segment(pa, pb, a0, a1, b0, b1);
            sphere_sphere(pa, r1, pb, r2, &dist, pos, fr);
        }
        if (dist - margin < 0 && nc < ORC_MAXCON) add_contact2(d, &nc, dist, pos, fr, g1, g2, mu, margin);
    }
    d->ncon[0] = (REAL)nc;
    d->nself[0] = (REAL)(nc - nplane);
    d->nbox_skipped[0] = (REAL)nbox;
}

/* ---------------------------------------------------------------- constraints */
static void kbi(const OrcModel* m, REAL pos, REAL* k, REAL* b, REAL* imp) {
    double timeconst = fmax(m->solref[0], 2.0 * m->timestep[0]), dampratio = m->solref[1];
    double dmin = fmin(fmax(m->solimp[0], ORC_MINIMP), ORC_MAXIMP), dmax = fmin(fmax(m->solimp[1], ORC_MINIMP), ORC_MAXIMP);
    double width = fmax(ORC_MINVAL, m->solimp[2]), mid = fmin(fmax(m->solimp[3], ORC_MINIMP), ORC_MAXIMP);
    double power = fmax(1.0, m->solimp[4]);
    REAL kk = R(1.0 / (dmax * dmax * timeconst * timeconst * dampratio * dampratio));
    REAL bb = R(2.0 / (dmax * timeconst));
    if (m->solref[0] <= 0) kk = R(-m->solref[0] / (dmax * dmax));
    if (m->solref[1] <= 0) bb = R(-m->solref[1] / dmax);
    REAL x = R(fabs((double)pos)) / R(width), y;
    if (x < R(mid)) y = R(1.0 / pow(mid, power - 1)) * R(pow((double)x, power));
    else y = 1 - R(1.0 / pow(1 - mid, power - 1)) * R(pow((double)(1 - x), power));
    REAL im = R(dmin) + y * R(dmax - dmin);
    if (im < R(dmin)) im = R(dmin);
    if (im > R(dmax)) im = R(dmax);
    if (x > 1) im = R(dmax);
    *k = kk; *b = bb; *imp = im;
}
static void finish_row(const OrcModel* m, OrcData* d, int r, REAL pos, REAL invweight, int enabled) {
    int nv = m->nv[0];
    REAL k, b, imp, vel = dotn(d->efc_J + r * NVMAX, d->qvel, nv), Rr;
    kbi(m, pos, &k, &b, &imp);
    Rr = invweight * (1 - imp) / imp;
    if (Rr < R(ORC_MINVAL)) Rr = R(ORC_MINVAL);
    d->efc_pos[r] = pos;
    d->efc_D[r] = 1 / Rr;
    d->efc_aref[r] = -b * vel - k * imp * pos;
    d->efc_enabled[r] = (REAL)enabled;
}
/* translational Jacobian of a point fixed to `body` (3 x nv, rows stride NVMAX) */
static void jac_point(const OrcModel* m, const OrcData* d, int body, const REAL* p, REAL* jp) {
    int nv = m->nv[0];
    REAL off[3] = {p[0] - d->subtree_com[0], p[1] - d->subtree_com[1], p[2] - d->subtree_com[2]};
    memset(jp, 0, 3 * NVMAX * sizeof(REAL));
    if (body <= 0) return;
    int i = m->body_dofadr[body] + m->body_dofnum[body] - 1;
    while (m->body_dofnum[body] == 0) { body = m->body_parent[body]; if (body <= 0) return; i = m->body_dofadr[body] + m->body_dofnum[body] - 1; }
    (void)nv;
    for (; i >= 0; i = m->dof_parentid[i]) {
        REAL t[3];
        cross3(t, d->cdof + 6 * i, off);
        for (int k = 0; k < 3; k++) jp[k * NVMAX + i] = d->cdof[6 * i + 3 + k] + t[k];
    }
}
static void make_constraint(const OrcModel* m, OrcData* d) {
    int nv = m->nv[0], r = 0;
    memset(d->efc_J, 0, sizeof(d->efc_J));
    for (int j = 0; j < m->njnt[0]; j++) {
        if (!m->jnt_limited[j] || m->jnt_type[j] != JNT_HINGE) continue;
        REAL q = d->qpos[m->jnt_qposadr[j]];
        REAL dmin = q - R(m->jnt_range[2 * j]), dmax = R(m->jnt_range[2 * j + 1]) - q;
        REAL pos = (dmin < dmax ? dmin : dmax) - R(m->jnt_margin[j]);
        d->efc_J[r * NVMAX + m->jnt_dofadr[j]] = dmin < dmax ? R(1) : R(-1);
        finish_row(m, d, r, pos, R(m->dof_invweight0[m->jnt_dofadr[j]]), pos < 0);
        r++;
    }
    int nc = (int)d->ncon[0];
    for (int c = 0; c < nc; c++) {
        int g = (int)d->con_geom[c], b2 = m->geom_bodyid[g], b1 = m->geom_bodyid[(int)d->con_geom1[c]];
        REAL jp[3 * NVMAX], jp1[3 * NVMAX], jc[3 * NVMAX], mu = d->con_mu[c];
        REAL pos = d->con_dist[c] - d->con_includemargin[c];
        jac_point(m, d, b2, d->con_pos + 3 * c, jp);
        jac_point(m, d, b1, d->con_pos + 3 * c, jp1); /* zero for the world body (floor contacts) */
        for (int i = 0; i < 3 * NVMAX; i++) jp[i] -= jp1[i];
        for (int a = 0; a < 3; a++)
            for (int i = 0; i < nv; i++)
                jc[a * NVMAX + i] = d->con_frame[9 * c + 3 * a] * jp[i] + d->con_frame[9 * c + 3 * a + 1] * jp[NVMAX + i] +
                                    d->con_frame[9 * c + 3 * a + 2] * jp[2 * NVMAX + i];
        REAL t = R(m->body_invweight0[2 * b1]) + R(m->body_invweight0[2 * b2]);
        REAL invweight = (t + mu * mu * t) * 2 * mu * mu / R(m->impratio[0]);
        for (int e = 0; e < 4; e++) {
            int a = 1 + e / 2;
            REAL sg = (e & 1) ? R(-1) : R(1);
            for (int i = 0; i < nv; i++) d->efc_J[r * NVMAX + i] = jc[i] + sg * mu * jc[a * NVMAX + i];
            finish_row(m, d, r, pos, invweight, pos < 0);
            r++;
        }
    }
    d->nefc[0] = (REAL)r;
}

/* ---------------------------------------------------------------- fwd_velocity / actuation / acceleration */
static void com_vel(const OrcModel* m, OrcData* d) {
    int nb = m->nbody[0];
    memset(d->cvel, 0, sizeof(d->cvel));
    memset(d->cdof_dot, 0, sizeof(d->cdof_dot));
    for (int b = 1; b < nb; b++) {
        REAL v[6];
        memcpy(v, d->cvel + 6 * m->body_parent[b], sizeof(v));
        for (int j = m->body_jntadr[b]; j < m->body_jntadr[b] + m->body_jntnum[b]; j++) {
            int da = m->jnt_dofadr[j];
            if (m->jnt_type[j] == JNT_FREE) {
                for (int i = 0; i < 3; i++) for (int k = 0; k < 6; k++) v[k] += d->cdof[6 * (da + i) + k] * d->qvel[da + i];
                for (int i = 3; i < 6; i++) cross_motion(d->cdof_dot + 6 * (da + i), v, d->cdof + 6 * (da + i));
                for (int i = 3; i < 6; i++) for (int k = 0; k < 6; k++) v[k] += d->cdof[6 * (da + i) + k] * d->qvel[da + i];
            } else {
                cross_motion(d->cdof_dot + 6 * da, v, d->cdof + 6 * da);
                for (int k = 0; k < 6; k++) v[k] += d->cdof[6 * da + k] * d->qvel[da];
            }
        }
        memcpy(d->cvel + 6 * b, v, sizeof(v));
    }
}
static void rne(const OrcModel* m, OrcData* d) {
    int nb = m->nbody[0], nv = m->nv[0];
    memset(d->cacc, 0, sizeof(d->cacc));
    for (int k = 0; k < 3; k++) d->cacc[3 + k] = -R(m->gravity[k]);
    for (int b = 1; b < nb; b++) {
        REAL a[6], t[6], u[6];
        memcpy(a, d->cacc + 6 * m->body_parent[b], sizeof(a));
        for (int i = m->body_dofadr[b]; i < m->body_dofadr[b] + m->body_dofnum[b]; i++)
            for (int k = 0; k < 6; k++) a[k] += d->cdof_dot[6 * i + k] * d->qvel[i];
        memcpy(d->cacc + 6 * b, a, sizeof(a));
        mul_inert_vec(t, d->cinert + 10 * b, a);
        mul_inert_vec(u, d->cinert + 10 * b, d->cvel + 6 * b);
        cross_force(d->cfrc + 6 * b, d->cvel + 6 * b, u);
        for (int k = 0; k < 6; k++) d->cfrc[6 * b + k] += t[k];
    }
    for (int k = 0; k < 6; k++) d->cfrc[k] = 0;
    for (int b = nb - 1; b > 0; b--) {
        int p = m->body_parent[b];
        for (int k = 0; k < 6; k++) d->cfrc[6 * p + k] += d->cfrc[6 * b + k];
    }
    for (int i = 0; i < nv; i++) d->qfrc_bias[i] = dotn(d->cdof + 6 * i, d->cfrc + 6 * m->dof_bodyid[i], 6);
}
static void fwd_actuation(const OrcModel* m, OrcData* d) {
    memset(d->qfrc_actuator, 0, sizeof(d->qfrc_actuator));
    for (int a = 0; a < m->nu[0]; a++) {
        REAL c = d->ctrl[a];
        if (m->act_ctrllimited[a]) {
            if (c < R(m->act_ctrlrange[2 * a])) c = R(m->act_ctrlrange[2 * a]);
            if (c > R(m->act_ctrlrange[2 * a + 1])) c = R(m->act_ctrlrange[2 * a + 1]);
        }
        d->qfrc_actuator[m->act_dofadr[a]] += R(m->act_gear[a]) * c;
    }
}

/* ---------------------------------------------------------------- solver */
typedef struct { REAL a; REAL djp2; REAL c1; int on; } BreakPt;
static int cmp_bp(const void* x, const void* y) {
    REAL a = ((const BreakPt*)x)->a, b = ((const BreakPt*)y)->a;
    return a < b ? -1 : (a > b ? 1 : 0);
}
static REAL cost_of(const OrcModel* m, const OrcData* d, const REAL* a, int ne) {
    int nv = m->nv[0];
    REAL c = 0, da[NVMAX];
    for (int i = 0; i < nv; i++) da[i] = a[i] - d->qacc_smooth[i];
    for (int i = 0; i < nv; i++) c += R(0.5) * da[i] * dotn(d->qM + i * NVMAX, da, nv);
    for (int r = 0; r < ne; r++) {
        if (d->efc_enabled[r] == 0) continue;
        REAL jar = dotn(d->efc_J + r * NVMAX, a, nv) - d->efc_aref[r];
        if (jar < 0) c += R(0.5) * d->efc_D[r] * jar * jar;
    }
    return c;
}
static void solve_constraints(const OrcModel* m, OrcData* d) {
    int nv = m->nv[0], ne = (int)d->nefc[0], any = 0, it = 0;
    REAL a[NVMAX], grad[NVMAX], p[NVMAX], Mp[NVMAX], H[NVMAX * NVMAX], L[NVMAX * NVMAX];
    REAL jar[ORC_MAXEFC], jp[ORC_MAXEFC];
    BreakPt bp[ORC_MAXEFC];
    for (int r = 0; r < ne; r++) any |= d->efc_enabled[r] != 0;
    memcpy(d->qacc, d->qacc_smooth, sizeof(d->qacc));
    memset(d->qfrc_constraint, 0, sizeof(d->qfrc_constraint));
    memset(d->efc_force, 0, sizeof(d->efc_force));
    d->solver_niter[0] = 0;
    if (!any) return;
    REAL scale = R(1.0 / (m->meaninertia[0] * (nv > 1 ? nv : 1))), tol = R(m->tolerance[0]);
    memcpy(a, cost_of(m, d, d->qacc_warmstart, ne) < cost_of(m, d, d->qacc_smooth, ne) ? d->qacc_warmstart : d->qacc_smooth, sizeof(a));
    REAL cost = cost_of(m, d, a, ne);
    for (it = 0; it < m->iterations[0]; it++) {
        for (int i = 0; i < nv; i++) grad[i] = dotn(d->qM + i * NVMAX, a, nv) - d->qfrc_smooth[i];
        memcpy(H, d->qM, sizeof(H));
        for (int r = 0; r < ne; r++) {
            const REAL* J = d->efc_J + r * NVMAX;
            jar[r] = dotn(J, a, nv) - d->efc_aref[r];
            if (d->efc_enabled[r] == 0 || jar[r] >= 0) continue;
            for (int i = 0; i < nv; i++) {
                if (J[i] == 0) continue;
                grad[i] += d->efc_D[r] * jar[r] * J[i];
                for (int j = 0; j < nv; j++) H[i * NVMAX + j] += d->efc_D[r] * J[i] * J[j];
            }
        }
        REAL gn = rsqrt_(dotn(grad, grad, nv));
        if (scale * gn < tol) break;
        if (chol_factor(L, H, nv)) break;
        for (int i = 0; i < nv; i++) p[i] = -grad[i];
        chol_solve(L, p, nv);
        /* exact line search: walk the break points of the piecewise-linear derivative */
        REAL c0 = dotn(grad, p, nv), slope; /* phi'(alpha) = c0 + alpha*slope on the current segment */
        for (int i = 0; i < nv; i++) Mp[i] = dotn(d->qM + i * NVMAX, p, nv);
        slope = dotn(p, Mp, nv);
        int nb = 0;
        for (int r = 0; r < ne; r++) {
            if (d->efc_enabled[r] == 0) continue;
            jp[r] = dotn(d->efc_J + r * NVMAX, p, nv);
            int on = jar[r] < 0;
            if (on) slope += d->efc_D[r] * jp[r] * jp[r];
            if (jp[r] != 0) {
                REAL ab = -jar[r] / jp[r];
                if (ab > 0 || (jar[r] == 0 && jp[r] < 0)) {
                    bp[nb].a = ab > 0 ? ab : 0; bp[nb].djp2 = d->efc_D[r] * jp[r] * jp[r];
                    bp[nb].c1 = d->efc_D[r] * jar[r] * jp[r]; bp[nb].on = on; nb++;
                }
            }
        }
        qsort(bp, nb, sizeof(BreakPt), cmp_bp);
        REAL alpha = 0;
        int k = 0;
        for (;;) {
            REAL cand = slope > 0 ? -c0 / slope : R(1e30);
            if (k >= nb || cand <= bp[k].a) { alpha = cand; break; }
            if (bp[k].on) { c0 -= bp[k].c1; slope -= bp[k].djp2; } else { c0 += bp[k].c1; slope += bp[k].djp2; }
            k++;
        }
        if (!(alpha > 0) || alpha > R(1e20)) break;
        for (int i = 0; i < nv; i++) a[i] += alpha * p[i];
        REAL cnew = cost_of(m, d, a, ne);
        REAL improvement = cost - cnew;
        cost = cnew;
        if (scale * improvement < tol) { it++; break; }
    }
    d->solver_niter[0] = (REAL)it;
    memcpy(d->qacc, a, sizeof(a));
    for (int r = 0; r < ne; r++) {
        REAL jr = dotn(d->efc_J + r * NVMAX, a, nv) - d->efc_aref[r];
        if (d->efc_enabled[r] != 0 && jr < 0) {
            d->efc_force[r] = -d->efc_D[r] * jr;
            for (int i = 0; i < nv; i++) d->qfrc_constraint[i] += d->efc_J[r * NVMAX + i] * d->efc_force[r];
        }
    }
}

/* test hooks: the distance functions on their own (tests/test_self_collision.py) */
void orc_test_capsule_box(const REAL* cpos, const REAL* cmat, REAL r, REAL hl, const REAL* bpos, const REAL* bmat, const double* half,
                          REAL* dist, REAL* pos, REAL* nrm) {
    REAL dd[2], pp[2][3], nn[2][3];
    capsule_box(cpos, cmat, r, hl, bpos, bmat, half, dd, pp, nn);
    for (int i = 0; i < 2; i++) { dist[i] = dd[i]; for (int k = 0; k < 3; k++) { pos[3 * i + k] = pp[i][k]; nrm[3 * i + k] = nn[i][k]; } }
}
void orc_test_box_box(const REAL* c1, const REAL* m1, const double* h1, const REAL* c2, const REAL* m2, const double* h2, REAL* dist, REAL* pos, REAL* nrm) {
    REAL dd[4], pp[4][3], nn[4][3];
    box_box(c1, m1, h1, c2, m2, h2, dd, pp, nn);
    for (int i = 0; i < 4; i++) { dist[i] = dd[i]; for (int k = 0; k < 3; k++) { pos[3 * i + k] = pp[i][k]; nrm[3 * i + k] = nn[i][k]; } }
}
void orc_test_capsule_capsule(const REAL* a0, const REAL* a1, REAL r1, const REAL* b0, const REAL* b1, REAL r2, REAL* dist, REAL* pos, REAL* frame) {
    REAL pa[3], pb[3];
    closest_segment_to_segment(pa, pb, a0, a1, b0, b1);
    sphere_sphere(pa, r1, pb, r2, dist, pos, frame);
}
void orc_test_sphere_capsule(const REAL* c, REAL r1, const REAL* b0, const REAL* b1, REAL r2, REAL* dist, REAL* pos, REAL* frame) {
    REAL pb[3];
    closest_segment_point(pb, b0, b1, c);
    sphere_sphere(c, r1, pb, r2, dist, pos, frame);
}
void orc_test_sphere_box(const REAL* sp, REAL r, const REAL* bp, const REAL* bm, const double* half, REAL* dist, REAL* pos, REAL* frame) {
    sphere_box(sp, r, bp, bm, half, dist, pos, frame);
}

/* ---------------------------------------------------------------- public: forward / step */
void orc_reset_data(const OrcModel* m, OrcData* d) {
    memset(d, 0, sizeof(*d));
    for (int i = 0; i < m->nq[0]; i++) d->qpos[i] = R(m->qpos0[i]);
}
void orc_forward(const OrcModel* m, OrcData* d) {
    int nv = m->nv[0];
    kinematics(m, d);
    com_pos(m, d);
    crb(m, d);
    collision(m, d);
    make_constraint(m, d);
    com_vel(m, d);
    for (int i = 0; i < nv; i++) d->qfrc_passive[i] = -R(m->dof_damping[i]) * d->qvel[i];
    rne(m, d);
    fwd_actuation(m, d);
    for (int i = 0; i < nv; i++) {
        d->qfrc_smooth[i] = d->qfrc_passive[i] - d->qfrc_bias[i] + d->qfrc_actuator[i] + d->qfrc_applied[i];
        d->qacc_smooth[i] = d->qfrc_smooth[i];
    }
    chol_solve(d->qL, d->qacc_smooth, nv);
    solve_constraints(m, d);
}
void orc_step(const OrcModel* m, OrcData* d) {
    int nv = m->nv[0];
    REAL h = R(m->timestep[0]), Md[NVMAX * NVMAX], L[NVMAX * NVMAX], qacc[NVMAX];
    orc_forward(m, d);
    memcpy(d->qacc_warmstart, d->qacc, sizeof(d->qacc));
    /* Euler with implicit joint damping (mjx forward.euler) */
    int damped = 0;
    for (int i = 0; i < nv; i++) damped |= m->dof_damping[i] > 0;
    if (damped) {
        memcpy(Md, d->qM, sizeof(Md));
        for (int i = 0; i < nv; i++) { Md[i * NVMAX + i] += h * R(m->dof_damping[i]); qacc[i] = d->qfrc_smooth[i] + d->qfrc_constraint[i]; }
        chol_factor(L, Md, nv);
        chol_solve(L, qacc, nv);
    } else {
        memcpy(qacc, d->qacc, sizeof(qacc));
    }
    for (int i = 0; i < nv; i++) d->qvel[i] += h * qacc[i];
    for (int j = 0; j < m->njnt[0]; j++) {
        int qa = m->jnt_qposadr[j], da = m->jnt_dofadr[j];
        if (m->jnt_type[j] == JNT_FREE) {
            for (int k = 0; k < 3; k++) d->qpos[qa + k] += h * d->qvel[da + k];
            REAL* w = d->qvel + da + 3, nrm = rsqrt_(dot3(w, w)), ax[3] = {1, 0, 0}, ql[4], qn[4];
            if (nrm > R(1e-30)) { ax[0] = w[0] / nrm; ax[1] = w[1] / nrm; ax[2] = w[2] / nrm; }
            axis_angle_quat(ql, ax, h * nrm);
            quat_mul(qn, d->qpos + qa + 3, ql);
            quat_norm(qn);
            memcpy(d->qpos + qa + 3, qn, sizeof(qn));
        } else {
            d->qpos[qa] += h * d->qvel[da];
        }
    }
    d->time[0] += h;
}

/* ---------------------------------------------------------------- PD laws
 * /root/reference/agents_env/pds_controllers_agents.py:17-34 (standard), :42-78 (feedback),
 * :83-172 (stable).  `lag` supplies qM / qfrc_bias / qfrc_applied exactly like
 * `state.pipeline_state` does in the reference (fields of the LAST forward). */
void orc_pd_torque(const OrcModel* m, const OrcData* lag, int mode, int action_cols, const REAL* action,
                   const REAL* q, const REAL* qd, const REAL* kp, const REAL* kd, REAL time, REAL dt, REAL* tau) {
    int nv = m->nv[0], nu = m->nu[0], nq = m->nq[0], qo = nq - nu, vo = nv - nu;
    REAL target[ORC_MAXU];
    for (int a = 0; a < nu; a++) {
        if (action_cols == 4) {
            REAL t = (mode == 3) ? time + dt : time;
            const REAL* c = action + 4 * a;
            target[a] = c[0] + c[1] * t + c[2] * (t * t) + c[3] * (t * t * t);
        } else {
            target[a] = action[a];
        }
    }
    if (mode == 0) { for (int a = 0; a < nu; a++) tau[a] = action[a * action_cols]; return; }
    if (mode == 1) { /* standard */
        for (int a = 0; a < nu; a++) tau[a] = -kp[a] * (q[qo + a] - target[a]) + (-kd[a] * qd[vo + a]);
        return;
    }
    if (mode == 2) { /* feedback: (M [0;pd] + bias)[6:] */
        REAL pd[NVMAX];
        for (int i = 0; i < nv; i++) pd[i] = 0;
        for (int a = 0; a < nu; a++) pd[vo + a] = -kp[a] * (q[qo + a] - target[a]) + (-kd[a] * qd[vo + a]);
        for (int a = 0; a < nu; a++) tau[a] = dotn(lag->qM + (vo + a) * NVMAX, pd, nv) + lag->qfrc_bias[vo + a];
        return;
    }
    /* stable PD */
    REAL err[NVMAX], rhs[NVMAX], A[NVMAX * NVMAX], L[NVMAX * NVMAX];
    for (int i = 0; i < nv; i++) err[i] = 0;
    for (int a = 0; a < nu; a++) err[vo + a] = (q[qo + a] + qd[vo + a] * dt) - target[a];
    memcpy(A, lag->qM, sizeof(A));
    for (int a = 0; a < nu; a++) A[(vo + a) * NVMAX + vo + a] += kd[a] * dt;
    for (int i = 0; i < nv; i++) {
        REAL kpi = i >= vo ? kp[i - vo] : 0, kdi = i >= vo ? kd[i - vo] : 0;
        rhs[i] = -lag->qfrc_bias[i] - kpi * err[i] - kdi * qd[i] + lag->qfrc_applied[i];
    }
    chol_factor(L, A, nv);
    chol_solve(L, rhs, nv);
    for (int a = 0; a < nu; a++) tau[a] = -kp[a] * err[vo + a] - kd[a] * (qd[vo + a] + rhs[vo + a] * dt);
}
