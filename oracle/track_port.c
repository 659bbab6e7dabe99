/* Plain-C restatement of the reference's Track sweep -- TEST INFRASTRUCTURE and the Track half of
 * bench.py's CPU arm (kind "port": the reference itself cannot be compiled here).
 *
 *   SolveFrame / match gathering     /root/reference/cpp/tracker.cc:36-131
 *   GetRayObjectSpace, RayCast       /root/reference/cpp/ray_casting.h:53-63, ray_casting.cc:65-133
 *   ray / triangle                   /root/reference/cpp/ray_casting.h:125-179 (Moller-Trumbore)
 *   PnPProblem                       /root/reference/cpp/pnp/pnp_problem.h:13-142
 *   LevMarqDenseSolver               /root/reference/cpp/pnp/lev_marq.h:99-389
 *   SolvePnPIterative                /root/reference/cpp/pnp/solvers.cc:11-78
 *   robust losses                    /root/reference/cpp/pnp/robust_loss.h:47-104
 *   camera / pose / quaternion       /root/reference/cpp/pnp/types.h:18-198, pose.h:9-160, pnp/quaternion.h:11-20
 *
 * float32 like the reference (`using Float = float`), sums formed sequentially in residual order (the
 * reference with max_allowed_parallelism = 1).  Embree (third party, absent) is replaced by a median-split
 * BVH over the same triangles: nearest hit, tnear = 0; a hit on a masked triangle is a miss.
 * PARITY UNPINNED (no reference fixture exists for this path); cross-checked against the independent numpy
 * restatement (oracle/pnp.py, raycast.py, track.py) in tests/test_oracle_track_port.py.
 * Nothing under polychase_b200/ may link or call this file. */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    float fx, fy, cx, cy, aspect, width, height, convention; /* 0 = OpenGL, 1 = OpenCV */
    float q[4];                                               /* w, x, y, z */
    float t[3];
    float filled;
} orc_cam;

typedef struct {
    uint64_t max_iterations;
    int loss_type; /* 0 trivial, 1 huber, 2 cauchy */
    float loss_scale, gradient_tol, step_tol, initial_lambda, min_lambda, max_lambda;
} orc_bundle_opts;

typedef struct {
    uint64_t iterations;
    float initial_cost, cost, lambda;
    uint64_t invalid_steps;
    float step_norm, grad_norm;
} orc_bundle_stats;

/* ---- BVH (stand-in for Embree's rtcIntersect1) --------------------------------------------- */
typedef struct {
    float bmin[3], bmax[3];
    int first, count; /* count > 0: leaf over tri_order[first .. first+count); else first = left child, right = first+1 */
} bvh_node;

typedef struct {
    bvh_node* nodes;
    int num_nodes, cap_nodes;
    int* order;       /* triangle ids in leaf order */
    float* cen;       /* centroids, nt x 3 */
    float* tbmin;     /* per-triangle bounds */
    float* tbmax;
    const float* verts;
    const uint32_t* tris;
    float* verts_own;
    uint32_t* tris_own;
    int nt;
} orc_bvh;

static int g_axis;
static const float* g_cen;
static int cmp_axis(const void* a, const void* b) {
    const float x = g_cen[3 * *(const int*)a + g_axis], y = g_cen[3 * *(const int*)b + g_axis];
    return (x > y) - (x < y);
}

static void bvh_build_rec(orc_bvh* B, int node, int lo, int hi) {
    bvh_node* n = &B->nodes[node];
    float cmin[3] = {INFINITY, INFINITY, INFINITY}, cmax[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int k = 0; k < 3; k++) { n->bmin[k] = INFINITY; n->bmax[k] = -INFINITY; }
    for (int i = lo; i < hi; i++) {
        const int t = B->order[i];
        for (int k = 0; k < 3; k++) {
            if (B->tbmin[3 * t + k] < n->bmin[k]) n->bmin[k] = B->tbmin[3 * t + k];
            if (B->tbmax[3 * t + k] > n->bmax[k]) n->bmax[k] = B->tbmax[3 * t + k];
            if (B->cen[3 * t + k] < cmin[k]) cmin[k] = B->cen[3 * t + k];
            if (B->cen[3 * t + k] > cmax[k]) cmax[k] = B->cen[3 * t + k];
        }
    }
    int axis = 0;
    if (cmax[1] - cmin[1] > cmax[axis] - cmin[axis]) axis = 1;
    if (cmax[2] - cmin[2] > cmax[axis] - cmin[axis]) axis = 2;
    if (hi - lo <= 4 || !(cmax[axis] - cmin[axis] > 0.f)) {
        n->first = lo;
        n->count = hi - lo;
        return;
    }
    g_axis = axis;
    g_cen = B->cen;
    qsort(B->order + lo, (size_t)(hi - lo), sizeof(int), cmp_axis);
    const int mid = lo + (hi - lo) / 2;
    const int left = B->num_nodes;
    B->num_nodes += 2;
    B->nodes[node].first = left;
    B->nodes[node].count = 0;
    bvh_build_rec(B, left, lo, mid);
    bvh_build_rec(B, left + 1, mid, hi);
}

orc_bvh* orc_bvh_build(const float* verts, int nv, const uint32_t* tris, int nt) {
    orc_bvh* B = (orc_bvh*)calloc(1, sizeof(orc_bvh));
    B->nt = nt;
    B->verts_own = (float*)malloc(sizeof(float) * 3 * (size_t)nv);
    B->tris_own = (uint32_t*)malloc(sizeof(uint32_t) * 3 * (size_t)nt);
    memcpy(B->verts_own, verts, sizeof(float) * 3 * (size_t)nv);
    memcpy(B->tris_own, tris, sizeof(uint32_t) * 3 * (size_t)nt);
    B->verts = B->verts_own;
    B->tris = B->tris_own;
    B->order = (int*)malloc(sizeof(int) * (size_t)nt);
    B->cen = (float*)malloc(sizeof(float) * 3 * (size_t)nt);
    B->tbmin = (float*)malloc(sizeof(float) * 3 * (size_t)nt);
    B->tbmax = (float*)malloc(sizeof(float) * 3 * (size_t)nt);
    for (int i = 0; i < nt; i++) {
        B->order[i] = i;
        for (int k = 0; k < 3; k++) {
            const float a = verts[3 * tris[3 * i] + k], b = verts[3 * tris[3 * i + 1] + k], c = verts[3 * tris[3 * i + 2] + k];
            B->tbmin[3 * i + k] = fminf(a, fminf(b, c));
            B->tbmax[3 * i + k] = fmaxf(a, fmaxf(b, c));
            B->cen[3 * i + k] = (a + b + c) * (1.f / 3.f);
        }
    }
    B->cap_nodes = 2 * nt + 2;
    B->nodes = (bvh_node*)calloc((size_t)B->cap_nodes, sizeof(bvh_node));
    B->num_nodes = 1;
    bvh_build_rec(B, 0, 0, nt);
    return B;
}

void orc_bvh_free(orc_bvh* B) {
    if (!B) return;
    free(B->nodes); free(B->order); free(B->cen); free(B->tbmin); free(B->tbmax); free(B->verts_own); free(B->tris_own);
    free(B);
}

/* IntersectWithJac(ray, triangle) without the Jacobians, ray_casting.h:125-179 */
static int ray_tri(const float o[3], const float d[3], const float* p1, const float* p2, const float* p3, float* t_out,
                   float* u_out, float* v_out) {
    const float eps = 1e-10f;
    const float e1[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]}, e2[3] = {p3[0] - p1[0], p3[1] - p1[1], p3[2] - p1[2]};
    const float rx[3] = {d[1] * e2[2] - d[2] * e2[1], d[2] * e2[0] - d[0] * e2[2], d[0] * e2[1] - d[1] * e2[0]};
    const float det = e1[0] * rx[0] + e1[1] * rx[1] + e1[2] * rx[2];
    if (det > -eps && det < eps) return 0;
    const float inv = 1.0f / det;
    const float s[3] = {o[0] - p1[0], o[1] - p1[1], o[2] - p1[2]};
    const float u = inv * (s[0] * rx[0] + s[1] * rx[1] + s[2] * rx[2]);
    if (u < 0.f || u > 1.f) return 0;
    const float sx[3] = {s[1] * e1[2] - s[2] * e1[1], s[2] * e1[0] - s[0] * e1[2], s[0] * e1[1] - s[1] * e1[0]};
    const float v = inv * (d[0] * sx[0] + d[1] * sx[1] + d[2] * sx[2]);
    if (v < 0.f || u + v > 1.f) return 0;
    const float t = inv * (e2[0] * sx[0] + e2[1] * sx[1] + e2[2] * sx[2]);
    if (t < 0.f) return 0;
    *t_out = t; *u_out = u; *v_out = v;
    return 1;
}

/* nearest hit of one ray; returns the primitive id or -1 */
static int bvh_nearest(const orc_bvh* B, const float o[3], const float d[3], float* t_best, float* u_best, float* v_best) {
    int stack[128], sp = 0, best = -1;
    float tb = INFINITY, ub = 0, vb = 0;
    float inv[3];
    for (int k = 0; k < 3; k++) {
        const float x = fabsf(d[k]) < 1e-18f ? copysignf(1e-18f, d[k]) : d[k];
        inv[k] = 1.f / x;
    }
    stack[sp++] = 0;
    while (sp > 0) {
        const bvh_node* n = &B->nodes[stack[--sp]];
        float tmin = 0.f, tmax = tb;
        int miss = 0;
        for (int k = 0; k < 3 && !miss; k++) {
            float t0 = (n->bmin[k] - o[k]) * inv[k], t1 = (n->bmax[k] - o[k]) * inv[k];
            if (t0 > t1) { const float tt = t0; t0 = t1; t1 = tt; }
            t0 -= fabsf(t0) * 1e-4f + 1e-6f;                 /* conservative slab test */
            t1 += fabsf(t1) * 1e-4f + 1e-6f;
            if (t0 > tmin) tmin = t0;
            if (t1 < tmax) tmax = t1;
            if (tmin > tmax) miss = 1;
        }
        if (miss) continue;
        if (n->count > 0) {
            for (int i = n->first; i < n->first + n->count; i++) {
                const int tri = B->order[i];
                const uint32_t* ix = &B->tris[3 * tri];
                float t, u, v;
                if (ray_tri(o, d, &B->verts[3 * ix[0]], &B->verts[3 * ix[1]], &B->verts[3 * ix[2]], &t, &u, &v) &&
                    (t < tb || (t == tb && tri < best))) {
                    tb = t; ub = u; vb = v; best = tri;
                }
            }
        } else if (sp + 2 <= 128) {
            stack[sp++] = n->first;
            stack[sp++] = n->first + 1;
        }
    }
    *t_best = tb; *u_best = ub; *v_best = vb;
    return best;
}

/* ---- camera math ------------------------------------------------------------------------------ */
static void quat_to_matrix(const float q[4], float R[9]) { /* Eigen::Quaternionf::toRotationMatrix */
    const float w = q[0], x = q[1], y = q[2], z = q[3];
    const float tx = 2.f * x, ty = 2.f * y, tz = 2.f * z;
    const float twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x;
    const float tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1.f - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
    R[3] = txy + twz; R[4] = 1.f - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1.f - (txx + tyy);
}

static void quat_rotate(const float q[4], const float v[3], float out[3]) { /* Eigen: v + w*2(u x v) + u x 2(u x v) */
    const float ux = q[1], uy = q[2], uz = q[3], w = q[0];
    float cx = uy * v[2] - uz * v[1], cy = uz * v[0] - ux * v[2], cz = ux * v[1] - uy * v[0];
    cx += cx; cy += cy; cz += cz;
    out[0] = v[0] + w * cx + (uy * cz - uz * cy);
    out[1] = v[1] + w * cy + (uz * cx - ux * cz);
    out[2] = v[2] + w * cz + (ux * cy - uy * cx);
}

static void quat_step_post(const float q[4], const float w3[3], float out[4]) { /* quaternion.h:11-20 */
    const float angle = sqrtf(w3[0] * w3[0] + w3[1] * w3[1] + w3[2] * w3[2]);
    if (!(angle > 0.f)) { memcpy(out, q, 4 * sizeof(float)); return; }
    const float ax[3] = {w3[0] / angle, w3[1] / angle, w3[2] / angle};
    const float half = 0.5f * angle, c = cosf(half), s = sinf(half);
    const float bw = c, bx = s * ax[0], by = s * ax[1], bz = s * ax[2];
    const float aw = q[0], axx = q[1], ay = q[2], az = q[3];
    out[0] = aw * bw - axx * bx - ay * by - az * bz;
    out[1] = aw * bx + axx * bw + ay * bz - az * by;
    out[2] = aw * by + ay * bw + az * bx - axx * bz;
    out[3] = aw * bz + az * bw + axx * by - ay * bx;
}

static int invert4(const double m[16], double inv[16]) {
    double a[4][8];
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) { a[i][j] = m[4 * i + j]; a[i][4 + j] = i == j; }
    for (int c = 0; c < 4; c++) {
        int p = c;
        for (int r = c + 1; r < 4; r++)
            if (fabs(a[r][c]) > fabs(a[p][c])) p = r;
        if (a[p][c] == 0.0) return 0;
        if (p != c)
            for (int j = 0; j < 8; j++) { const double t = a[c][j]; a[c][j] = a[p][j]; a[p][j] = t; }
        const double d = a[c][c];
        for (int j = 0; j < 8; j++) a[c][j] /= d;
        for (int r = 0; r < 4; r++)
            if (r != c) {
                const double f = a[r][c];
                for (int j = 0; j < 8; j++) a[r][j] -= f * a[c][j];
            }
    }
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) inv[4 * i + j] = a[i][4 + j];
    return 1;
}

/* RayCast(accel_mesh, scene_transform, pos, check_mask) for n image positions under one camera
 * (ray_casting.cc:128-133, ray_casting.h:53-63).  Outputs: hit flag, object-space position, primitive. */
int orc_ray_cast(const orc_bvh* B, const float model[16], const orc_cam* cam, const float* pos, int n, const uint32_t* mask,
                 int check_mask, uint8_t* hit, float* pos_out, uint32_t* prim_out) {
    float R[9];
    quat_to_matrix(cam->q, R);
    double view[16] = {R[0], R[1], R[2], cam->t[0], R[3], R[4], R[5], cam->t[1], R[6], R[7], R[8], cam->t[2], 0, 0, 0, 1};
    double vm[16], inv[16];
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            double s = 0;
            for (int k = 0; k < 4; k++) s += view[4 * i + k] * (double)model[4 * k + j];
            vm[4 * i + j] = s;
        }
    if (!invert4(vm, inv)) return -1;
    float mat[16];
    for (int i = 0; i < 16; i++) mat[i] = (float)inv[i];
    const float o[3] = {mat[3], mat[7], mat[11]};
    const float sgn = cam->convention != 0.f ? 1.f : -1.f;     /* types.h:95-98 */
    for (int i = 0; i < n; i++) {
        const float dc[3] = {sgn * ((pos[2 * i] - cam->cx) / cam->fx), sgn * ((pos[2 * i + 1] - cam->cy) / cam->fy), sgn};
        const float d[3] = {mat[0] * dc[0] + mat[1] * dc[1] + mat[2] * dc[2], mat[4] * dc[0] + mat[5] * dc[1] + mat[6] * dc[2],
                            mat[8] * dc[0] + mat[9] * dc[1] + mat[10] * dc[2]};
        float t, u, v;
        int prim = bvh_nearest(B, o, d, &t, &u, &v);
        if (prim >= 0 && check_mask && mask && ((mask[prim >> 5] >> (prim & 31)) & 1u)) prim = -1;   /* ray_casting.cc:106-108 */
        hit[i] = prim >= 0;
        if (prim_out) prim_out[i] = prim >= 0 ? (uint32_t)prim : 0xFFFFFFFFu;
        if (prim >= 0) {
            const uint32_t* ix = &B->tris[3 * prim];
            const float w = 1.0f - u - v;                        /* geometry.h:17-19 */
            for (int k = 0; k < 3; k++)
                pos_out[3 * i + k] = w * B->verts[3 * ix[0] + k] + u * B->verts[3 * ix[1] + k] + v * B->verts[3 * ix[2] + k];
        } else {
            pos_out[3 * i] = pos_out[3 * i + 1] = pos_out[3 * i + 2] = 0.f;
        }
    }
    return 0;
}

/* ---- robust losses (robust_loss.h:47-104) --------------------------------------------------- */
typedef struct { int kind; float thr, sq, inv_sq; } loss_t;
static loss_t make_loss(int kind, float scale) {
    loss_t l = {kind, scale, scale * scale, 0.f};
    l.inv_sq = (float)(1.0 / (double)l.sq);
    return l;
}
static float loss_value(const loss_t* l, float r2) {
    if (l->kind == 0) return r2;
    if (l->kind == 1) {
        if (r2 <= l->sq) return r2;
        const float r = sqrtf(r2);
        return (float)((double)l->thr * (2.0 * (double)r - (double)l->thr));
    }
    return l->sq * log1pf(r2 * l->inv_sq);
}
static float loss_weight(const loss_t* l, float r2) {
    if (l->kind == 0) return 1.f;
    if (l->kind == 1) return r2 <= l->sq ? 1.f : l->thr / sqrtf(r2);
    const float w = 1.f / (1.f + r2 * l->inv_sq);
    return w > FLT_MIN ? w : FLT_MIN;
}

/* ---- PnPProblem + LevMarqDenseSolver ------------------------------------------------------------ */
typedef struct { float f_low, f_high, cx_low, cx_high, cy_low, cy_high; } bounds_t;
static bounds_t get_bounds(const orc_cam* c) { /* types.h:156-192 */
    const float min_fov = (float)(15.f * M_PI / 180), max_fov = (float)(160.f * M_PI / 180);
    const float tmin = tanf(min_fov / 2), tmax = tanf(max_fov / 2);
    bounds_t b;
    if (c->convention == 0.f) { b.f_low = -(c->width / 2.0f) / tmin; b.f_high = -(c->width / 2.0f) / tmax; }
    else { b.f_high = (c->width / 2.0f) / tmin; b.f_low = (c->width / 2.0f) / tmax; }
    b.cx_low = 0.f; b.cx_high = c->width; b.cy_low = 0.f; b.cy_high = c->height;
    return b;
}
static float clampf(float v, float lo, float hi) { return v < lo ? lo : (hi < v ? hi : v); }

typedef struct {
    const float* X; const float* x; const float* w; int m;
    int opt_f, opt_pp;
    bounds_t bounds;
    loss_t loss;
} pnp_t;

static void evaluate(const orc_cam* c, const float* X, const float* x, float r[2]) { /* pnp_problem.h:52-61 */
    float Z[3];
    quat_rotate(c->q, X, Z);
    Z[0] += c->t[0]; Z[1] += c->t[1]; Z[2] += c->t[2];
    const int behind = c->convention != 0.f ? Z[2] < 0.f : Z[2] > 0.f;
    if (behind) { r[0] = FLT_MAX; r[1] = FLT_MAX; return; }
    r[0] = c->fx * Z[0] / Z[2] + c->cx - x[0];
    r[1] = c->fy * Z[1] / Z[2] + c->cy - x[1];
}

static float total_cost(const pnp_t* P, const orc_cam* c) { /* lev_marq.h:316-356, kShouldNormalize = false */
    float cost = 0.f;
    for (int i = 0; i < P->m; i++) {
        const float wt = P->w ? P->w[i] : 1.f;
        if (wt == 0.f) continue;
        float r[2];
        evaluate(c, P->X + 3 * i, P->x + 2 * i, r);
        cost += wt * loss_value(&P->loss, r[0] * r[0] + r[1] * r[1]);
    }
    return cost;
}

static void build_normal_equations(const pnp_t* P, const orc_cam* c, const float R[9], float JtJ[81], float Jtr[9], float diag[9]) {
    memset(JtJ, 0, 81 * sizeof(float));
    memset(Jtr, 0, 9 * sizeof(float));
    for (int i = 0; i < P->m; i++) {                          /* lev_marq.h:231-297 */
        const float wt = P->w ? P->w[i] : 1.f;
        if (wt == 0.f) continue;
        const float* Z = P->X + 3 * i;
        const float Y[3] = {R[0] * Z[0] + R[1] * Z[1] + R[2] * Z[2] + c->t[0], R[3] * Z[0] + R[4] * Z[1] + R[5] * Z[2] + c->t[1],
                            R[6] * Z[0] + R[7] * Z[1] + R[8] * Z[2] + c->t[2]};   /* pose.h:75-96 */
        const float res[2] = {c->fx * Y[0] / Y[2] + c->cx - P->x[2 * i], c->fy * Y[1] / Y[2] + c->cy - P->x[2 * i + 1]};
        const float dz[2][3] = {{c->fx / Y[2], 0.f, -c->fx * Y[0] / (Y[2] * Y[2])}, {0.f, c->fy / Y[2], -c->fy * Y[1] / (Y[2] * Y[2])}};
        /* dRtZ_dR = R * Skew(-Z) */
        const float nz[3] = {-Z[0], -Z[1], -Z[2]};
        const float S[9] = {0, -nz[2], nz[1], nz[2], 0, -nz[0], -nz[1], nz[0], 0};
        float dR[9];
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) dR[3 * a + b] = R[3 * a] * S[b] + R[3 * a + 1] * S[3 + b] + R[3 * a + 2] * S[6 + b];
        float J[2][9];
        for (int r = 0; r < 2; r++) {
            for (int b = 0; b < 3; b++) J[r][b] = dz[r][0] * dR[b] + dz[r][1] * dR[3 + b] + dz[r][2] * dR[6 + b];
            for (int b = 0; b < 3; b++) J[r][3 + b] = dz[r][b];
            J[r][6] = J[r][7] = J[r][8] = 0.f;
        }
        if (P->opt_f) { J[0][6] = c->aspect * Y[0] / Y[2]; J[1][6] = Y[1] / Y[2]; }   /* types.h:88-92 */
        if (P->opt_pp) { J[0][7] = 1.f; J[1][8] = 1.f; }
        const float tw = wt * loss_weight(&P->loss, res[0] * res[0] + res[1] * res[1]);
        for (int a = 0; a < 9; a++)
            for (int b = 0; b <= a; b++) JtJ[9 * a + b] += tw * (J[0][a] * J[0][b] + J[1][a] * J[1][b]);
        for (int a = 0; a < 9; a++) Jtr[a] += J[0][a] * (tw * res[0]) + J[1][a] * (tw * res[1]);
    }
    for (int a = 0; a < 9; a++) diag[a] = fminf(fmaxf(JtJ[10 * a], 1e-6f), 1e32f);   /* :296 */
}

static int llt_solve9(const float A[81], const float b[9], float x[9]) { /* LLT<Lower> + solve; 0 = NumericalIssue */
    float L[81];
    memcpy(L, A, sizeof(L));
    for (int k = 0; k < 9; k++) {
        float d = L[10 * k];
        for (int j = 0; j < k; j++) d -= L[9 * k + j] * L[9 * k + j];
        if (!(d > 0.f)) return 0;
        d = sqrtf(d);
        L[10 * k] = d;
        for (int i = k + 1; i < 9; i++) {
            float s = L[9 * i + k];
            for (int j = 0; j < k; j++) s -= L[9 * i + j] * L[9 * k + j];
            L[9 * i + k] = s / d;
        }
    }
    float y[9];
    for (int i = 0; i < 9; i++) {
        float s = b[i];
        for (int j = 0; j < i; j++) s -= L[9 * i + j] * y[j];
        y[i] = s / L[10 * i];
    }
    for (int i = 8; i >= 0; i--) {
        float s = y[i];
        for (int j = i + 1; j < 9; j++) s -= L[9 * j + i] * x[j];
        x[i] = s / L[10 * i];
    }
    return 1;
}

static void pnp_step(const pnp_t* P, const orc_cam* c, const float dp[9], orc_cam* out) { /* pnp_problem.h:101-131 */
    *out = *c;
    quat_step_post(c->q, dp, out->q);
    for (int k = 0; k < 3; k++) out->t[k] = c->t[k] + dp[3 + k];
    if (P->opt_f) {
        out->fy = c->fy + dp[6];
        out->fx = out->fy * out->aspect;
        out->fy = clampf(out->fy, P->bounds.f_low, P->bounds.f_high);
        out->fx = clampf(out->fx, P->bounds.f_low, P->bounds.f_high);
    }
    if (P->opt_pp) {
        out->cx = clampf(c->cx + dp[7], P->bounds.cx_low, P->bounds.cx_high);
        out->cy = clampf(c->cy + dp[8], P->bounds.cy_low, P->bounds.cy_high);
    }
}

/* SolvePnPIterative, solvers.cc:11-78.  Returns 0, or -1 for bad arguments (rows < 3, unknown loss). */
int orc_solve_pnp(const float* X, const float* x, const float* w, int m, const orc_bundle_opts* o, float max_inlier_error,
                  int opt_f, int opt_pp, orc_cam* cam, orc_bundle_stats* st, float* inlier_ratio) {
    if (m < 3 || o->loss_type < 0 || o->loss_type > 2) return -1;
    pnp_t P = {X, x, w, m, opt_f && m > 3, opt_pp && m > 3, get_bounds(cam), make_loss(o->loss_type, o->loss_scale)};
    orc_cam params = *cam, params_new = *cam;
    float R[9];
    quat_to_matrix(params.q, R);
    memset(st, 0, sizeof(*st));                               /* lev_marq.h:132-228 */
    st->cost = total_cost(&P, &params);
    st->initial_cost = st->cost;
    st->grad_norm = -1.f;
    st->step_norm = -1.f;
    st->lambda = o->initial_lambda;
    float v = 2.0f;
    int rebuild = 1;
    float JtJ[81], Jtr[9], diag[9], step[9];
    for (st->iterations = 0; st->iterations < o->max_iterations; ++st->iterations) {
        if (rebuild) {
            build_normal_equations(&P, &params, R, JtJ, Jtr, diag);
            float g = 0.f;
            for (int a = 0; a < 9; a++) g += Jtr[a] * Jtr[a];
            st->grad_norm = sqrtf(g);
            if (st->grad_norm < o->gradient_tol) break;
        }
        float A[81];
        memcpy(A, JtJ, sizeof(A));
        for (int a = 0; a < 9; a++) A[10 * a] = diag[a] * (float)(1.0 + (double)st->lambda);   /* :301 */
        float sol[9];
        if (!llt_solve9(A, Jtr, sol)) {
            st->invalid_steps++;
            if (st->lambda == o->max_lambda) break;
            st->lambda = fminf(o->max_lambda, st->lambda * v);
            v = 2 * v;
            rebuild = 0;
            continue;
        }
        float sn = 0.f;
        for (int a = 0; a < 9; a++) { step[a] = -sol[a]; sn += step[a] * step[a]; }
        st->step_norm = sqrtf(sn);
        if (st->step_norm < o->step_tol) break;
        pnp_step(&P, &params, step, &params_new);
        const float cost_new = total_cost(&P, &params_new);
        if (cost_new < st->cost) {
            const float actual = cost_new - st->cost;
            /* step^T (2 Jtr + JtJ.selfadjointView<Lower>() * step), JtJ with the clamped, undamped diagonal */
            float expected = 0.f;
            for (int a = 0; a < 9; a++) {
                float s = 0.f;
                for (int b = 0; b < 9; b++) s += (a == b ? diag[a] : (a > b ? JtJ[9 * a + b] : JtJ[9 * b + a])) * step[b];
                expected += step[a] * (2.0f * Jtr[a] + s);
            }
            const float rho = actual / expected;
            if (rho > 0) {
                const float factor = (float)fmax(1.0 / 3.0, 1.0 - pow(2.0 * rho - 1.0, 3));
                st->lambda = clampf(st->lambda * factor, o->min_lambda, o->max_lambda);
            }
            params = params_new;
            quat_to_matrix(params.q, R);
            st->cost = cost_new;
            v = 2;
            rebuild = 1;
        } else {
            st->invalid_steps++;
            if (st->lambda == o->max_lambda) break;
            st->lambda = fminf(o->max_lambda, st->lambda * v);
            v = 2 * v;
            rebuild = 0;
        }
    }
    *cam = params;
    if (inlier_ratio) {                                       /* solvers.cc:30-47 */
        size_t inl = 0;
        if (max_inlier_error > 0.f)
            for (int i = 0; i < m; i++) {
                float r[2];
                evaluate(&params, X + 3 * i, x + 2 * i, r);
                if (r[0] * r[0] + r[1] * r[1] < max_inlier_error * max_inlier_error) inl++;
            }
        *inlier_ratio = (float)inl / (float)m;
    }
    return 0;
}

/* SolveFrame, tracker.cc:36-131: per posed source frame cast the matched keypoints, keep the hits, then solve from
 * `init`.  srcs_* are arrays of length nsrc.  Returns the number of matches (< 3: nothing solved, like nullopt). */
int orc_track_frame(const orc_bvh* B, const uint32_t* mask, const float model[16], int nsrc, const orc_cam* src_cams,
                    const float* const* src_kps, const uint32_t* const* src_idx, const float* const* src_tgt,
                    const int* src_rows, const orc_cam* init, const orc_bundle_opts* o, int opt_f, int opt_pp,
                    orc_cam* out, orc_bundle_stats* st, float* inlier_ratio) {
    size_t total = 0;
    for (int s = 0; s < nsrc; s++) total += (size_t)src_rows[s];
    float* X = (float*)malloc(sizeof(float) * 3 * (total + 1));
    float* x = (float*)malloc(sizeof(float) * 2 * (total + 1));
    int m = 0;
    for (int s = 0; s < nsrc; s++) {
        const int n = src_rows[s];
        float* kp = (float*)malloc(sizeof(float) * 2 * (size_t)(n + 1));
        uint8_t* hit = (uint8_t*)malloc((size_t)n + 1);
        float* pos = (float*)malloc(sizeof(float) * 3 * (size_t)(n + 1));
        for (int i = 0; i < n; i++) { kp[2 * i] = src_kps[s][2 * src_idx[s][i]]; kp[2 * i + 1] = src_kps[s][2 * src_idx[s][i] + 1]; }
        if (orc_ray_cast(B, model, &src_cams[s], kp, n, mask, 1, hit, pos, NULL) == 0)
            for (int i = 0; i < n; i++)
                if (hit[i]) {                                  /* tracker.cc:80-86 */
                    const float* p = pos + 3 * i;
                    for (int k = 0; k < 3; k++)
                        X[3 * m + k] = model[4 * k] * p[0] + model[4 * k + 1] * p[1] + model[4 * k + 2] * p[2] + model[4 * k + 3];
                    x[2 * m] = src_tgt[s][2 * i];
                    x[2 * m + 1] = src_tgt[s][2 * i + 1];
                    m++;
                }
        free(kp); free(hit); free(pos);
    }
    if (m >= 3) {
        *out = *init;
        orc_solve_pnp(X, x, NULL, m, o, 12.0f, opt_f, opt_pp, out, st, inlier_ratio);
    }
    free(X); free(x);
    return m;
}
