/* TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of the crowddynamics per-timestep agent update.
 *
 * Plain C, serial, IEEE fp64 with contraction disabled (build with -ffp-contract=off), mirroring the
 * evaluation order of the reference's numba code so that results agree with it to the last few ulp.
 * It is the checker for the CUDA path and the "port" CPU baseline timed by bench.py.  It is NOT part
 * of the product: nothing under crowddynamics_b200/ links, loads or calls it.
 *
 * Pinning: tests/test_oracle_vs_golden.py checks every function here against golden vectors produced
 * by the reference's own numba code (tests/golden/generate.py, run where /root/reference exists), and
 * against the reference's two known-answer cases (core/motion/tests/test_power_law_benchmark.py:13-63).
 * The block-list part restates the third-party `cell_lists` package (unpinned git dependency,
 * requirements.txt:5, source not available) from the in-tree spec core/block_list.py:28-52:
 * "parity unpinned" at that boundary.
 *
 * All file:line citations are relative to /root/reference/crowddynamics/.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SIGTH_SOC 3.0   /* core/interactions.py:45 */
#define F_SOC_MAX 2e3   /* core/motion/power_law.py:52 */
#define TAU_MAX 30.0    /* core/motion/power_law.py:243 */

/* Data contract: simulation/agents.py:447-457 via traits.py:158-219 (packed, no alignment). */
#pragma pack(push, 1)
typedef struct {
    uint8_t active, target_reached;
    int64_t target;
    uint8_t is_leader, is_follower;
    int64_t index_leader, familiar_exit;
    double radius, r_t, r_s, r_ts, mass, inertia_rot, target_velocity, target_angular_velocity;
    double position[2], velocity[2], target_direction[2], force[2], force_prev[2];
    double tau_adj, k_soc, tau_0, mu, kappa, damping, std_rand_force;
} agent_circular_t; /* 228 bytes */

typedef struct {
    double position_ls[2], position_rs[2];
    agent_circular_t c;
    double orientation, angular_velocity, target_orientation, torque, torque_prev, tau_rot, std_rand_torque;
} agent_three_circle_t; /* 316 bytes */

typedef struct { double p0[2], p1[2]; } obstacle_linear_t; /* core/structures.py:6-9 */
#pragma pack(pop)

_Static_assert(sizeof(agent_circular_t) == 228, "circular itemsize");
_Static_assert(sizeof(agent_three_circle_t) == 316, "three_circle itemsize");

#define ORACLE_OK 0
#define ORACLE_INVALID_TYPE 1

static inline agent_circular_t *circ(void *agents, int64_t itemsize, int64_t i) {
    char *p = (char *)agents + i * itemsize;
    return itemsize == 316 ? &((agent_three_circle_t *)p)->c : (agent_circular_t *)p;
}
static inline agent_three_circle_t *three(void *agents, int64_t i) {
    return (agent_three_circle_t *)((char *)agents + i * 316);
}

/* ---- core/vector2D.py ------------------------------------------------------------------- */
static inline double dot2(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1]; }     /* :118-131 */
static inline double cross2(const double *a, const double *b) { return a[0] * b[1] - a[1] * b[0]; }   /* :136-149 */
static inline double length2(const double *v) { return hypot(v[0], v[1]); }                           /* :100-113 */
static inline void truncate2(double *v, double l) {                                                    /* :167-187 */
    double vlen = length2(v);
    if (vlen > l) { double s = l / vlen; v[0] *= s; v[1] *= s; }
}
/* Python float modulo as numba lowers it (CPython float_rem semantics). */
static inline double py_mod(double x, double y) {
    double m = fmod(x, y);
    if (m != 0.0) { if ((y < 0) != (m < 0)) m += y; }
    else m = copysign(0.0, y);
    return m;
}
double oracle_wrap_to_pi(double rad) {                                                                 /* :8-36 */
    double rad_ = py_mod(rad, 2 * M_PI);
    if (rad < 0 && rad_ == M_PI) return -M_PI;
    else if (rad_ > M_PI) return rad_ - (2 * M_PI);
    else return rad_;
}

/* ---- core/motion/contact.py:14-48 ------------------------------------------------------ */
static inline void force_contact(double h, const double *n, const double *v, const double *t,
                                 double mu, double kappa, double damping, double *out) {
    double kvt = kappa * dot2(v, t);
    double dvn = damping * dot2(v, n);
    for (int k = 0; k < 2; ++k)
        out[k] = -h * (mu * n[k] - kvt * t[k]) + dvn * n[k];
}

/* ---- core/distance.py:19-47 ------------------------------------------------------------ */
static inline double distance_circles(const double *x0, double r0, const double *x1, double r1, double *n) {
    double x[2] = {x0[0] - x1[0], x0[1] - x1[1]};
    double d = length2(x);
    double r_tot = r0 + r1;
    double h = d - r_tot;
    if (d == 0.0) { n[0] = 0.0; n[1] = 0.0; }
    else { n[0] = x[0] / d; n[1] = x[1] / d; }
    return h;
}

/* ---- core/distance.py:55-105 (incl. the x0[j_min] quirk at :103) ------------------------ */
static double distance_three_circles(const double *const x0[3], const double r0[3],
                                     const double *const x1[3], const double r1[3],
                                     double *normal, double *r_moment0, double *r_moment1) {
    double h_min = NAN;
    int i_min = 0, j_min = 0;
    normal[0] = normal[1] = 0.0;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double n[2];
            double h = distance_circles(x0[i], r0[i], x1[j], r1[j], n);
            if (h < h_min || isnan(h_min)) { h_min = h; normal[0] = n[0]; normal[1] = n[1]; i_min = i; j_min = j; }
        }
    for (int k = 0; k < 2; ++k) {
        r_moment0[k] = x0[i_min][k] + r0[i_min] * normal[k] - x0[0][k];
        r_moment1[k] = x0[j_min][k] - r1[j_min] * normal[k] - x1[0][k];
    }
    return h_min;
}

/* ---- core/distance.py:110-146 ----------------------------------------------------------- */
static double distance_circle_line(const double *x, double r, const double *p0, const double *p1, double *n_iw) {
    double d[2] = {p1[0] - p0[0], p1[1] - p0[1]};
    double l_w = length2(d);
    double t_w[2] = {d[0] / l_w, d[1] / l_w};
    double n_w[2] = {-t_w[1], t_w[0]};               /* rotate90, vector2D.py:39-57 */
    double q0[2] = {x[0] - p0[0], x[1] - p0[1]};
    double q1[2] = {x[0] - p1[0], x[1] - p1[1]};
    double l_t = -dot2(t_w, q1) - dot2(t_w, q0);
    double d_iw;
    if (l_t > l_w) {
        d_iw = length2(q0); n_iw[0] = q0[0] / d_iw; n_iw[1] = q0[1] / d_iw;
    } else if (l_t < -l_w) {
        d_iw = length2(q1); n_iw[0] = q1[0] / d_iw; n_iw[1] = q1[1] / d_iw;
    } else {
        double l_n = dot2(n_w, q0);
        d_iw = fabs(l_n);
        double s = isnan(l_n) ? l_n : (l_n > 0) - (l_n < 0);   /* np.sign */
        n_iw[0] = s * n_w[0]; n_iw[1] = s * n_w[1];
    }
    return d_iw - r;
}

/* ---- core/distance.py:154-180 ----------------------------------------------------------- */
static double distance_three_circle_line(const double *const x[3], const double r[3], const double *p0,
                                         const double *p1, double *normal, double *r_moment) {
    double h_min = NAN;
    int i_min = 0;
    normal[0] = normal[1] = 0.0;
    for (int i = 0; i < 3; ++i) {
        double n[2];
        double h = distance_circle_line(x[i], r[i], p0, p1, n);
        if (h < h_min || isnan(h_min)) { h_min = h; normal[0] = n[0]; normal[1] = n[1]; i_min = i; }
    }
    for (int k = 0; k < 2; ++k) r_moment[k] = x[i_min][k] - r[i_min] * normal[k] - x[0][k];
    return h_min;
}

/* ---- core/motion/power_law.py:84-102 ---------------------------------------------------- */
static inline double magnitude(double tau, double tau_0) {
    return (2.0 / tau + 1.0 / tau_0) * exp(-tau / tau_0) / (tau * tau);
}

/* ---- core/motion/power_law.py:215-259 --------------------------------------------------- */
void oracle_force_social_circular(void *agents, int64_t itemsize, int64_t i, int64_t j, double *force_i, double *force_j) {
    const agent_circular_t *ai = circ(agents, itemsize, i), *aj = circ(agents, itemsize, j);
    force_i[0] = force_i[1] = force_j[0] = force_j[1] = 0.0;
    double x_rel[2] = {ai->position[0] - aj->position[0], ai->position[1] - aj->position[1]};
    double v_rel[2] = {ai->velocity[0] - aj->velocity[0], ai->velocity[1] - aj->velocity[1]};
    double r_tot = ai->radius + aj->radius;
    double a = dot2(v_rel, v_rel);
    double b = -dot2(x_rel, v_rel);
    double c = dot2(x_rel, x_rel) - r_tot * r_tot;
    double d = sqrt(b * b - a * c);
    if (isnan(d) || d == 0 || a == 0) return;
    double tau = (b - d) / a;
    if (tau <= 0 || tau > TAU_MAX) return;
    double mag_i = magnitude(tau, ai->tau_0), mag_j = magnitude(tau, aj->tau_0);
    for (int k = 0; k < 2; ++k) {
        double grad = (v_rel[k] - (v_rel[k] * b + x_rel[k] * a) / d) / a;   /* :107-126 */
        force_i[k] += -ai->mass * ai->k_soc * grad * mag_i;
        force_j[k] -= -aj->mass * aj->k_soc * grad * mag_j;
    }
    truncate2(force_i, F_SOC_MAX);
    truncate2(force_j, F_SOC_MAX);
}

/* ---- core/motion/power_law.py:264-363 (selection rule :324, no tau_max) ----------------- */
void oracle_force_social_three_circle(void *agents, int64_t i, int64_t j, double *force_i, double *force_j) {
    const agent_three_circle_t *ai = three(agents, i), *aj = three(agents, j);
    force_i[0] = force_i[1] = force_j[0] = force_j[1] = 0.0;
    double v_rel[2] = {ai->c.velocity[0] - aj->c.velocity[0], ai->c.velocity[1] - aj->c.velocity[1]};
    double a = dot2(v_rel, v_rel);
    if (a == 0) return;
    const double *x_i[3] = {ai->c.position, ai->position_ls, ai->position_rs};
    const double *x_j[3] = {aj->c.position, aj->position_ls, aj->position_rs};
    double r_i[3] = {ai->c.r_t, ai->c.r_s, ai->c.r_s};
    double r_j[3] = {aj->c.r_t, aj->c.r_s, aj->c.r_s};
    int contact_i = 0, contact_j = 0;
    double tau = NAN, b_min = NAN, d_min = NAN;
    for (int pi = 0; pi < 3; ++pi)
        for (int pj = 0; pj < 3; ++pj) {
            double x_rel[2] = {x_i[pi][0] - x_j[pj][0], x_i[pi][1] - x_j[pj][1]};
            double r_tot = r_i[pi] + r_j[pj];
            double b = -dot2(x_rel, v_rel);
            double c = dot2(x_rel, x_rel) - r_tot * r_tot;
            double d = sqrt(b * b - a * c);
            if (isnan(d) || d == 0) continue;
            double tau_new = (b - d) / a;
            if (isnan(tau) || (0 < tau_new && tau_new < tau)) {
                contact_i = pi; contact_j = pj; tau = tau_new; b_min = b; d_min = d;
            }
        }
    if (isnan(tau) || tau <= 0) return;
    double r_off_i[2] = {0, 0}, r_off_j[2] = {0, 0};
    if (contact_i == 1) {
        double phi = ai->orientation;
        r_off_i[0] += ai->c.r_ts * sin(phi); r_off_i[1] += ai->c.r_ts * -cos(phi);
    } else if (contact_i == 2) {
        double phi = ai->orientation;
        r_off_i[0] -= ai->c.r_ts * sin(phi); r_off_i[1] -= ai->c.r_ts * -cos(phi);
    }
    if (contact_j == 1) {
        double phi = aj->orientation;
        r_off_j[0] += aj->c.r_ts * sin(phi); r_off_j[1] += aj->c.r_ts * -cos(phi);
    } else if (contact_j == 2) {
        double phi = aj->orientation;
        r_off_j[0] -= aj->c.r_ts * sin(phi); r_off_j[1] -= aj->c.r_ts * -cos(phi);
    }
    double mag_i = magnitude(tau, ai->c.tau_0), mag_j = magnitude(tau, aj->c.tau_0);
    for (int k = 0; k < 2; ++k) {
        double x_rel = ai->c.position[k] - aj->c.position[k];
        double r_off = r_off_i[k] - r_off_j[k];
        double grad = (v_rel[k] - (a * (x_rel + 2 * r_off) + b_min * v_rel[k]) / d_min) / a;   /* :131-149 */
        force_i[k] += -ai->c.mass * ai->c.k_soc * grad * mag_i;
        force_j[k] -= -aj->c.mass * aj->c.k_soc * grad * mag_j;
    }
    truncate2(force_i, F_SOC_MAX);
    truncate2(force_j, F_SOC_MAX);
}

/* ---- core/interactions.py:53-70 --------------------------------------------------------- */
static void interaction_agent_agent_circular(int64_t i, int64_t j, void *agents) {
    agent_circular_t *ai = circ(agents, 228, i), *aj = circ(agents, 228, j);
    double n[2];
    double h = distance_circles(ai->position, ai->radius, aj->position, aj->radius, n);
    if (h < SIGTH_SOC) {
        double force_i[2], force_j[2];
        oracle_force_social_circular(agents, 228, i, j, force_i, force_j);
        if (h < 0) {
            double t[2] = {n[1], -n[0]};   /* rotate270, vector2D.py:60-78 */
            double v[2] = {ai->velocity[0] - aj->velocity[0], ai->velocity[1] - aj->velocity[1]};
            double fc[2];
            force_contact(h, n, v, t, ai->mu, ai->kappa, ai->damping, fc);
            force_i[0] += fc[0]; force_i[1] += fc[1];
            force_contact(h, n, v, t, aj->mu, aj->kappa, aj->damping, fc);
            force_j[0] -= fc[0]; force_j[1] -= fc[1];
        }
        ai->force[0] += force_i[0]; ai->force[1] += force_i[1];
        aj->force[0] += force_j[0]; aj->force[1] += force_j[1];
    }
}

/* ---- core/interactions.py:75-104 -------------------------------------------------------- */
static void interaction_agent_agent_three_circle(int64_t i, int64_t j, void *agents) {
    agent_three_circle_t *ai = three(agents, i), *aj = three(agents, j);
    const double *x_i[3] = {ai->c.position, ai->position_ls, ai->position_rs};
    const double *x_j[3] = {aj->c.position, aj->position_ls, aj->position_rs};
    double r_i[3] = {ai->c.r_t, ai->c.r_s, ai->c.r_s};
    double r_j[3] = {aj->c.r_t, aj->c.r_s, aj->c.r_s};
    double n[2], r_moment_i[2], r_moment_j[2];
    double h = distance_three_circles(x_i, r_i, x_j, r_j, n, r_moment_i, r_moment_j);
    if (h < SIGTH_SOC) {
        double force_i[2], force_j[2];
        oracle_force_social_three_circle(agents, i, j, force_i, force_j);
        if (h < 0) {
            double t[2] = {n[1], -n[0]};
            double v[2] = {ai->c.velocity[0] - aj->c.velocity[0], ai->c.velocity[1] - aj->c.velocity[1]};
            double fc[2];
            force_contact(h, n, v, t, ai->c.mu, ai->c.kappa, ai->c.damping, fc);
            force_i[0] += fc[0]; force_i[1] += fc[1];
            force_contact(h, n, v, t, aj->c.mu, aj->c.kappa, aj->c.damping, fc);
            force_j[0] -= fc[0]; force_j[1] -= fc[1];
        }
        ai->c.force[0] += force_i[0]; ai->c.force[1] += force_i[1];
        aj->c.force[0] += force_j[0]; aj->c.force[1] += force_j[1];
        ai->torque += cross2(r_moment_i, force_i);
        aj->torque += cross2(r_moment_j, force_j);
    }
}

/* ---- cell_lists.add_to_cells (restated; spec core/block_list.py:28-52) ------------------
 * cell = floor(p / c) per axis (lattice anchored at multiples of c), flat = (ix - ix_min) * ny + (iy - iy_min),
 * counts, offset = exclusive cumulative sum, points_indices = stable counting sort (ascending agent index in a cell).
 * Outputs are malloc'ed; caller frees with oracle_free.  grid[4] = {ix_min, iy_min, nx, ny}. */
typedef struct {
    int64_t n, ncell;
    int64_t grid[4];
    int64_t *cell_of_agent;   /* [n] flat cell id of every agent */
    int64_t *points_indices;  /* [n] */
    int64_t *cells_count;     /* [ncell] */
    int64_t *cells_offset;    /* [ncell] */
} oracle_cells_t;

void oracle_free_cells(oracle_cells_t *c) {
    free(c->cell_of_agent); free(c->points_indices); free(c->cells_count); free(c->cells_offset);
    memset(c, 0, sizeof(*c));
}

int oracle_add_to_cells(const void *agents, int64_t n, int64_t itemsize, double cell_size, oracle_cells_t *out) {
    memset(out, 0, sizeof(*out));
    if (itemsize != 228 && itemsize != 316) return ORACLE_INVALID_TYPE;
    out->n = n;
    if (n == 0) return ORACLE_OK;
    int64_t *ix = malloc(sizeof(int64_t) * n), *iy = malloc(sizeof(int64_t) * n);
    int64_t x_min = INT64_MAX, x_max = INT64_MIN, y_min = INT64_MAX, y_max = INT64_MIN;
    for (int64_t k = 0; k < n; ++k) {
        const agent_circular_t *a = circ((void *)agents, itemsize, k);
        ix[k] = (int64_t)floor(a->position[0] / cell_size);
        iy[k] = (int64_t)floor(a->position[1] / cell_size);
        if (ix[k] < x_min) x_min = ix[k];
        if (ix[k] > x_max) x_max = ix[k];
        if (iy[k] < y_min) y_min = iy[k];
        if (iy[k] > y_max) y_max = iy[k];
    }
    int64_t nx = x_max - x_min + 1, ny = y_max - y_min + 1, ncell = nx * ny;
    out->grid[0] = x_min; out->grid[1] = y_min; out->grid[2] = nx; out->grid[3] = ny;
    out->ncell = ncell;
    out->cell_of_agent = malloc(sizeof(int64_t) * n);
    out->points_indices = malloc(sizeof(int64_t) * n);
    out->cells_count = calloc(ncell, sizeof(int64_t));
    out->cells_offset = calloc(ncell, sizeof(int64_t));
    int64_t *fill = calloc(ncell, sizeof(int64_t));
    for (int64_t k = 0; k < n; ++k) {
        out->cell_of_agent[k] = (ix[k] - x_min) * ny + (iy[k] - y_min);
        out->cells_count[out->cell_of_agent[k]]++;
    }
    int64_t acc = 0;
    for (int64_t c = 0; c < ncell; ++c) { out->cells_offset[c] = acc; acc += out->cells_count[c]; }
    for (int64_t k = 0; k < n; ++k) {
        int64_t c = out->cell_of_agent[k];
        out->points_indices[out->cells_offset[c] + fill[c]++] = k;
    }
    free(fill); free(ix); free(iy);
    return ORACLE_OK;
}

/* ---- cell_lists.iter_nearest_neighbors (restated): every unordered pair in the same or an adjacent cell once,
 * as ordered (i, j): same cell i before j in points_indices; otherwise i in the cell with the smaller (ix, iy),
 * i.e. forward half stencil (0,+1), (+1,-1), (+1,0), (+1,+1).  visit(i, j, ctx). */
typedef void (*pair_fn)(int64_t, int64_t, void *);
static void for_each_pair(const oracle_cells_t *cl, pair_fn visit, void *ctx) {
    static const int sx[4] = {0, 1, 1, 1}, sy[4] = {1, -1, 0, 1};
    int64_t nx = cl->grid[2], ny = cl->grid[3];
    for (int64_t c = 0; c < cl->ncell; ++c) {
        int64_t n_c = cl->cells_count[c], o_c = cl->cells_offset[c];
        if (n_c == 0) continue;
        for (int64_t a = 0; a < n_c; ++a)
            for (int64_t b = a + 1; b < n_c; ++b)
                visit(cl->points_indices[o_c + a], cl->points_indices[o_c + b], ctx);
        int64_t x = c / ny, y = c % ny;
        for (int s = 0; s < 4; ++s) {
            int64_t x2 = x + sx[s], y2 = y + sy[s];
            if (x2 >= nx || y2 < 0 || y2 >= ny) continue;
            int64_t d = x2 * ny + y2, n_d = cl->cells_count[d], o_d = cl->cells_offset[d];
            for (int64_t a = 0; a < n_c; ++a)
                for (int64_t b = 0; b < n_d; ++b)
                    visit(cl->points_indices[o_c + a], cl->points_indices[o_d + b], ctx);
        }
    }
}

typedef struct { int64_t *i, *j; int64_t cap, count; } pair_sink_t;
static void sink_pair(int64_t i, int64_t j, void *ctx) {
    pair_sink_t *s = ctx;
    if (s->count < s->cap) { s->i[s->count] = i; s->j[s->count] = j; }
    s->count++;
}
/* Returns the number of candidate pairs (may exceed cap; only the first cap are written). */
int64_t oracle_neighbor_pairs(const void *agents, int64_t n, int64_t itemsize, double cell_size,
                              int64_t *out_i, int64_t *out_j, int64_t cap) {
    oracle_cells_t cl;
    if (oracle_add_to_cells(agents, n, itemsize, cell_size, &cl)) return -1;
    pair_sink_t s = {out_i, out_j, cap, 0};
    for_each_pair(&cl, sink_pair, &s);
    oracle_free_cells(&cl);
    return s.count;
}

static void visit_circular(int64_t i, int64_t j, void *ctx) { interaction_agent_agent_circular(i, j, ctx); }
static void visit_three(int64_t i, int64_t j, void *ctx) { interaction_agent_agent_three_circle(i, j, ctx); }

/* ---- core/interactions.py:191-205 ------------------------------------------------------- */
int oracle_agent_agent_block_list(void *agents, int64_t n, int64_t itemsize, double cell_size) {
    oracle_cells_t cl;
    if (oracle_add_to_cells(agents, n, itemsize, cell_size, &cl)) return ORACLE_INVALID_TYPE;
    for_each_pair(&cl, itemsize == 228 ? visit_circular : visit_three, agents);
    oracle_free_cells(&cl);
    return ORACLE_OK;
}

/* Brute force over all i<j pairs in index order (same pair kernels) -- used to show the block list loses no pair. */
int oracle_agent_agent_brute(void *agents, int64_t n, int64_t itemsize) {
    if (itemsize != 228 && itemsize != 316) return ORACLE_INVALID_TYPE;
    for (int64_t i = 0; i < n; ++i)
        for (int64_t j = i + 1; j < n; ++j)
            if (itemsize == 228) interaction_agent_agent_circular(i, j, agents);
            else interaction_agent_agent_three_circle(i, j, agents);
    return ORACLE_OK;
}

/* ---- core/interactions.py:107-141,169-186,208-214 ---------------------------------------- */
int oracle_agent_obstacle(void *agents, int64_t n, int64_t itemsize, const void *obstacles, int64_t n_obs) {
    const obstacle_linear_t *obs = obstacles;
    if (itemsize == 228) {
        for (int64_t i = 0; i < n; ++i)
            for (int64_t w = 0; w < n_obs; ++w) {
                agent_circular_t *a = circ(agents, 228, i);
                double nrm[2];
                double h = distance_circle_line(a->position, a->radius, obs[w].p0, obs[w].p1, nrm);
                if (h < 0) {
                    double t[2] = {nrm[1], -nrm[0]}, f[2];
                    force_contact(h, nrm, a->velocity, t, a->mu, a->kappa, a->damping, f);
                    a->force[0] += f[0]; a->force[1] += f[1];
                }
            }
    } else if (itemsize == 316) {
        for (int64_t i = 0; i < n; ++i)
            for (int64_t w = 0; w < n_obs; ++w) {
                agent_three_circle_t *a = three(agents, i);
                const double *x[3] = {a->c.position, a->position_ls, a->position_rs};
                double r[3] = {a->c.r_t, a->c.r_s, a->c.r_s};
                double nrm[2], r_moment[2];
                double h = distance_three_circle_line(x, r, obs[w].p0, obs[w].p1, nrm, r_moment);
                if (h < 0) {
                    double t[2] = {nrm[1], -nrm[0]}, f[2];
                    force_contact(h, nrm, a->c.velocity, t, a->c.mu, a->c.kappa, a->c.damping, f);
                    a->c.force[0] += f[0]; a->c.force[1] += f[1];
                    a->torque += cross2(r_moment, f);
                }
            }
    } else return ORACLE_INVALID_TYPE;
    return ORACLE_OK;
}

/* ---- core/motion/adjusting.py:18-51,56-95,101-121 (logic.py:89-94) ---------------------- */
int oracle_adjusting(void *agents, int64_t n, int64_t itemsize) {
    if (itemsize != 228 && itemsize != 316) return ORACLE_INVALID_TYPE;
    for (int64_t i = 0; i < n; ++i) {
        agent_circular_t *a = circ(agents, itemsize, i);
        double s = a->mass / a->tau_adj;
        for (int k = 0; k < 2; ++k)
            a->force[k] += s * (a->target_velocity * a->target_direction[k] - a->velocity[k]);
    }
    if (itemsize == 316)
        for (int64_t i = 0; i < n; ++i) {
            agent_three_circle_t *a = three(agents, i);
            a->torque += a->c.inertia_rot / a->tau_rot *
                         (oracle_wrap_to_pi(a->target_orientation - a->orientation) / M_PI * a->c.target_angular_velocity -
                          a->angular_velocity);
        }
    return ORACLE_OK;
}

/* ---- core/steering/orientation.py:17-21 (logic.py:258-261) ------------------------------ */
int oracle_orientation(void *agents, int64_t n, int64_t itemsize) {
    if (itemsize == 228) return ORACLE_OK;
    if (itemsize != 316) return ORACLE_INVALID_TYPE;
    for (int64_t i = 0; i < n; ++i) {
        agent_three_circle_t *a = three(agents, i);
        a->target_orientation = atan2(a->c.target_direction[1], a->c.target_direction[0]);
    }
    return ORACLE_OK;
}

/* ---- logic.py:149-165 + quickest_path.py:41-44 + navigation.py:60-78 ---------------------
 * indices = trunc((pos - (minx, miny)) / step) (astype(int64): toward zero); in-grid => e0 = (U[iy, ix], V[iy, ix]). */
int oracle_navigation(void *agents, int64_t n, int64_t itemsize, int64_t target, const double *U, const double *V,
                      int64_t ny, int64_t nx, double minx, double miny, double step) {
    if (itemsize != 228 && itemsize != 316) return ORACLE_INVALID_TYPE;
    for (int64_t i = 0; i < n; ++i) {
        agent_circular_t *a = circ(agents, itemsize, i);
        if (a->target != target) continue;
        double fx = (a->position[0] - minx) / step, fy = (a->position[1] - miny) / step;
        if (!(fabs(fx) < 9.0e18) || !(fabs(fy) < 9.0e18)) continue;   /* NaN/inf: numpy casts to INT64_MIN => outside */
        int64_t jx = (int64_t)fx, jy = (int64_t)fy;
        if (0 <= jy && jy < ny && 0 <= jx && jx < nx) {
            a->target_direction[0] = U[jy * nx + jx];
            a->target_direction[1] = V[jy * nx + jx];
        }
    }
    return ORACLE_OK;
}

/* ---- core/integrator.py:32-97 ----------------------------------------------------------- */
double oracle_adaptive_timestep(const void *agents, int64_t n, int64_t itemsize, double dt_min, double dt_max) {
    double v_max = 0.0, v0_max = -INFINITY;
    for (int64_t i = 0; i < n; ++i) {
        const agent_circular_t *a = circ((void *)agents, itemsize, i);
        double l = length2(a->velocity);
        if (l > v_max) v_max = l;
        if (a->target_velocity > v0_max || isnan(a->target_velocity)) v0_max = a->target_velocity;
    }
    if (v_max == 0.0) return dt_max;
    double dx_max = 1.1 * v0_max * dt_max;
    double dt = dx_max / v_max;
    if (dt > dt_max) return dt_max;
    else if (dt < dt_min) return dt_min;
    else return dt;
}

/* ---- core/integrator.py:167-193,209-256 + simulation/agents.py:473-486 ------------------- */
int oracle_velocity_verlet_integrator(void *agents, int64_t n, int64_t itemsize, double dt_min, double dt_max, double *dt_out) {
    if (itemsize != 228 && itemsize != 316) return ORACLE_INVALID_TYPE;
    double dt = oracle_adaptive_timestep(agents, n, itemsize, dt_min, dt_max);
    for (int64_t i = 0; i < n; ++i) {
        agent_circular_t *a = circ(agents, itemsize, i);
        for (int k = 0; k < 2; ++k) {
            double old_acc = a->force_prev[k] / a->mass;
            double new_acc = a->force[k] / a->mass;
            a->force_prev[k] = a->force[k];
            a->velocity[k] += (old_acc + new_acc) / 2 * dt;
            a->position[k] += a->velocity[k] * dt + new_acc / 2 * (dt * dt);
        }
    }
    if (itemsize == 316)
        for (int64_t i = 0; i < n; ++i) {
            agent_three_circle_t *a = three(agents, i);
            double old_acc = a->torque_prev / a->c.inertia_rot;
            double new_acc = a->torque / a->c.inertia_rot;
            a->torque_prev = a->torque;
            a->angular_velocity += (old_acc + new_acc) / 2 * dt;
            a->orientation += a->angular_velocity * dt + new_acc / 2 * (dt * dt);
            a->orientation = oracle_wrap_to_pi(a->orientation);
        }
    if (itemsize == 316)
        for (int64_t i = 0; i < n; ++i) {    /* shoulders */
            agent_three_circle_t *a = three(agents, i);
            double tx = sin(a->orientation), ty = -cos(a->orientation);   /* rotate270(unit_vector(phi)) */
            double ox = tx * a->c.r_ts, oy = ty * a->c.r_ts;
            a->position_ls[0] = a->c.position[0] - ox; a->position_ls[1] = a->c.position[1] - oy;
            a->position_rs[0] = a->c.position[0] + ox; a->position_rs[1] = a->c.position[1] + oy;
        }
    if (dt_out) *dt_out = dt;
    return ORACLE_OK;
}

/* ---- logic.py:59-64 --------------------------------------------------------------------- */
int oracle_reset(void *agents, int64_t n, int64_t itemsize) {
    if (itemsize != 228 && itemsize != 316) return ORACLE_INVALID_TYPE;
    for (int64_t i = 0; i < n; ++i) {
        agent_circular_t *a = circ(agents, itemsize, i);
        a->force[0] = a->force[1] = 0;
        if (itemsize == 316) three(agents, i)->torque = 0;
    }
    return ORACLE_OK;
}

/* One MultiAgentSimulation.update() (multiagent.py:51-55) in the Hallway post-order (examples/simulations.py:123-136),
 * Fluctuation / InsideDomain omitted: navigation -> orientation -> adjusting -> agent-agent -> agent-obstacle ->
 * integrator -> reset.  nav_* may describe n_targets fields laid out back to back (same shape). */
int oracle_step(void *agents, int64_t n, int64_t itemsize, const void *obstacles, int64_t n_obs,
                int64_t n_targets, const double *U, const double *V, int64_t ny, int64_t nx,
                double minx, double miny, double step, double cell_size, double dt_min, double dt_max, double *dt_out) {
    int rc;
    for (int64_t t = 0; t < n_targets; ++t)
        if ((rc = oracle_navigation(agents, n, itemsize, t, U + t * ny * nx, V + t * ny * nx, ny, nx, minx, miny, step))) return rc;
    if ((rc = oracle_orientation(agents, n, itemsize))) return rc;
    if ((rc = oracle_adjusting(agents, n, itemsize))) return rc;
    if ((rc = oracle_agent_agent_block_list(agents, n, itemsize, cell_size))) return rc;
    if ((rc = oracle_agent_obstacle(agents, n, itemsize, obstacles, n_obs))) return rc;
    if ((rc = oracle_velocity_verlet_integrator(agents, n, itemsize, dt_min, dt_max, dt_out))) return rc;
    return oracle_reset(agents, n, itemsize);
}

/* =====================================================================================================================
 * exit detection, herding, leader-follower (core/evacuation.py:137-174,
 * core/sensory_region.py:9-16, core/geom2D.py:38-59, core/steering/collective_motion.py:16-289).
 * ===================================================================================================================== */
#define NO_TARGET (-1)          /* simulation/agents.py:28 */
#define NO_LEADER (-1)          /* simulation/agents.py:29 */
#define MISSING_NEIGHBOR (-1)   /* collective_motion.py:13 */

int oracle_line_intersect(const double *x0, const double *x1, const double *y0, const double *y1) {   /* geom2D.py:38-59 */
    const double u[2] = {x1[0] - x0[0], x1[1] - x0[1]}, v[2] = {y1[0] - y0[0], y1[1] - y0[1]};
    const double b[2] = {y0[0] - x0[0], y0[1] - x0[1]};
    const double d = u[0] * v[1] - u[1] * v[0];
    if (d == 0) return 0;
    const double t0 = b[0] * v[1] - b[1] * v[0], t1 = b[0] * u[1] - b[1] * u[0];
    const double q0 = t0 / d, q1 = t1 / d;
    return 0 <= q0 && q0 <= 1 && 0 <= q1 && q1 <= 1;
}

int oracle_is_obstacle_between_points(const double *p0, const double *p1, const void *obstacles, int64_t n_obs) {   /* sensory_region.py:9-16 */
    const obstacle_linear_t *obs = obstacles;
    for (int64_t w = 0; w < n_obs; ++w)
        if (oracle_line_intersect(p0, p1, obs[w].p0, obs[w].p1)) return 1;
    return 0;
}

/* evacuation.py:137-174: closest door centre in detection range with a free line of sight; -1 / 0 if none */
int oracle_exit_detection(const double *center_door, int64_t n_doors, const void *agents, int64_t n, int64_t itemsize,
                          const void *obstacles, int64_t n_obs, double detection_range, int64_t *detected_exit, uint8_t *has_detected) {
    if (itemsize != 228 && itemsize != 316) return ORACLE_INVALID_TYPE;
    for (int64_t i = 0; i < n; ++i) {
        const double *p = circ((void *)agents, itemsize, i)->position;
        double distance = detection_range;
        detected_exit[i] = -1; has_detected[i] = 0;
        for (int64_t c = 0; c < n_doors; ++c) {
            if (oracle_is_obstacle_between_points(p, center_door + 2 * c, obstacles, n_obs)) continue;
            const double r[2] = {center_door[2 * c] - p[0], center_door[2 * c + 1] - p[1]};
            const double d = length2(r);
            if (d < distance) { distance = d; detected_exit[i] = c; has_detected[i] = 1; }
        }
    }
    return ORACLE_OK;
}

static inline void normalize2(const double *v, double *out) {   /* vector2D.py:152-163 */
    const double l = length2(v);
    if (l != 0) { out[0] = v[0] / l; out[1] = v[1] / l; } else { out[0] = v[0]; out[1] = v[1]; }
}

/* collective_motion.py:25-58; returns bit 0: agent 1 is heading away ..., bit 1: the second flag */
int oracle_herding_relationship(const double *x1, const double *x2, const double *v1, const double *v2, double phi) {
    if (length2(v1) == 0 || length2(v2) == 0) return 0;
    const double rel[2] = {x2[0] - x1[0], x2[1] - x1[1]};
    double e_rel[2], n1[2], n2[2];
    normalize2(rel, e_rel); normalize2(v1, n1); normalize2(v2, n2);
    const double c_i = dot2(e_rel, n1), c_j = -dot2(e_rel, n2);
    const double cos_phi = cos(phi);
    const int in_i = cos_phi < c_i && c_i < 1.0, in_j = cos_phi < c_j && c_j < 1.0;
    if (in_i) return in_j ? 0 : 1;
    return in_j ? 2 : 3;
}

typedef struct {
    const void *agents; int64_t itemsize; const void *obstacles; int64_t n_obs;
    int64_t k; int64_t *neighbors; double *distances, *distances_max;
} knn_ctx_t;

static void knn_set_neighbor(knn_ctx_t *c, int64_t i, int64_t j, double l) {   /* collective_motion.py:61-66 */
    double *d = c->distances + i * c->k;
    int64_t argmax = 0;
    for (int64_t s = 1; s < c->k; ++s) if (d[s] > d[argmax]) argmax = s;      /* np.argmax: first maximum */
    c->neighbors[i * c->k + argmax] = j;
    d[argmax] = l;
    double m = d[0];
    for (int64_t s = 1; s < c->k; ++s) if (d[s] > m) m = d[s];
    c->distances_max[i] = m;
}

static void knn_visit(int64_t i, int64_t j, void *ctx) {   /* collective_motion.py:94-108 */
    knn_ctx_t *c = ctx;
    const double *pi = circ((void *)c->agents, c->itemsize, i)->position, *pj = circ((void *)c->agents, c->itemsize, j)->position;
    if (oracle_is_obstacle_between_points(pi, pj, c->obstacles, c->n_obs)) return;
    const double r[2] = {pi[0] - pj[0], pi[1] - pj[1]};
    const double l = length2(r);
    if (l < c->distances_max[i]) knn_set_neighbor(c, i, j, l);
    if (l < c->distances_max[j]) knn_set_neighbor(c, j, i, l);
}

/* collective_motion.py:69-110 with the block list of :262-267 (cell_size = sight); neighbors[n][k], -1 = missing */
int oracle_find_nearest_neighbors(const void *agents, int64_t n, int64_t itemsize, double sight, int64_t k,
                                  const void *obstacles, int64_t n_obs, int64_t *neighbors) {
    if (itemsize != 228 && itemsize != 316) return ORACLE_INVALID_TYPE;
    if (k < 1) return ORACLE_OK;
    knn_ctx_t c = {agents, itemsize, obstacles, n_obs, k, neighbors, malloc(sizeof(double) * (n * k + 1)), malloc(sizeof(double) * (n + 1))};
    for (int64_t q = 0; q < n * k; ++q) { neighbors[q] = MISSING_NEIGHBOR; c.distances[q] = sight; }
    for (int64_t i = 0; i < n; ++i) c.distances_max[i] = sight;
    oracle_cells_t cl;
    int rc = oracle_add_to_cells(agents, n, itemsize, sight, &cl);
    if (!rc) { for_each_pair(&cl, knn_visit, &c); oracle_free_cells(&cl); }
    free(c.distances); free(c.distances_max);
    return rc;
}

static inline void weighted_average2(const double *e0, const double *e1, double w, double *out) {   /* vector2D.py:203-226 */
    out[0] = w * e0[0] + (1 - w) * e1[0]; out[1] = w * e0[1] + (1 - w) * e1[1];
}

/* collective_motion.py:113-154; new_direction[n][2], has_new_direction[n] */
int oracle_herding_interaction(const void *agents, int64_t n, int64_t itemsize, const uint8_t *is_herding, const int64_t *neighbors,
                               int64_t k, double weight_position, double phi, double *new_direction, uint8_t *has_new_direction) {
    if (itemsize != 228 && itemsize != 316) return ORACLE_INVALID_TYPE;
    for (int64_t i = 0; i < n; ++i) {
        new_direction[2 * i] = new_direction[2 * i + 1] = 0; has_new_direction[i] = 0;
        if (!is_herding[i]) continue;
        const agent_circular_t *a = circ((void *)agents, itemsize, i);
        double mean_position[2] = {0, 0}, mean_velocity[2] = {0, 0};
        int64_t num = 0;
        for (int64_t s = 0; s < k; ++s) {
            const int64_t j = neighbors[i * k + s];
            if (j == MISSING_NEIGHBOR) continue;
            const agent_circular_t *b = circ((void *)agents, itemsize, j);
            if (oracle_herding_relationship(a->position, b->position, a->velocity, b->velocity, phi) & 1) {
                mean_position[0] += b->position[0]; mean_position[1] += b->position[1];
                mean_velocity[0] += b->velocity[0]; mean_velocity[1] += b->velocity[1];
                num++;
            }
        }
        if (num > 0) {
            const double rel[2] = {mean_position[0] / num - a->position[0], mean_position[1] / num - a->position[1]};
            double e0[2], e1[2], w[2];
            normalize2(rel, e0); normalize2(mean_velocity, e1);
            weighted_average2(e0, e1, weight_position, w);
            normalize2(w, new_direction + 2 * i);
            has_new_direction[i] = 1;
        }
    }
    return ORACLE_OK;
}

typedef struct { double d; int64_t k; } dist_idx_t;
static int cmp_dist_idx(const void *a, const void *b) {
    const dist_idx_t *x = a, *y = b;
    if (x->d < y->d) return -1;
    if (x->d > y->d) return 1;
    return (x->k > y->k) - (x->k < y->k);   /* numba's argsort is not stable; ties (equal distances) are resolved by index here */
}

/* collective_motion.py:157-226; mutates target / index_leader of the followers in place, sequentially like the reference */
int oracle_leader_follower_interaction_brute(void *agents, int64_t n, int64_t itemsize, double weight_position, double phi,
                                             const void *obstacles, int64_t n_obs, double sight, double *new_direction, uint8_t *has_strategy) {
    if (itemsize != 228 && itemsize != 316) return ORACLE_INVALID_TYPE;
    int64_t n_lead = 0;
    int64_t *leaders = malloc(sizeof(int64_t) * (n + 1));
    for (int64_t i = 0; i < n; ++i) if (circ(agents, itemsize, i)->is_leader) leaders[n_lead++] = i;
    dist_idx_t *dist = malloc(sizeof(dist_idx_t) * (n_lead + 1));
    for (int64_t i = 0; i < n; ++i) { new_direction[2 * i] = new_direction[2 * i + 1] = 0; has_strategy[i] = 0; }
    for (int64_t i = 0; i < n; ++i) {
        agent_circular_t *a = circ(agents, itemsize, i);
        if (!a->is_follower) continue;
        int behind_obstacle = 0, heading_away = 0;
        for (int64_t k = 0; k < n_lead; ++k) {
            const agent_circular_t *b = circ(agents, itemsize, leaders[k]);
            const double r[2] = {a->position[0] - b->position[0], a->position[1] - b->position[1]};
            dist[k].d = length2(r); dist[k].k = k;
        }
        qsort(dist, n_lead, sizeof(dist_idx_t), cmp_dist_idx);
        for (int64_t q = 0; q < n_lead; ++q) {
            if (dist[q].d > sight) continue;
            const int64_t j = leaders[dist[q].k];
            const agent_circular_t *b = circ(agents, itemsize, j);
            if (oracle_is_obstacle_between_points(a->position, b->position, obstacles, n_obs)) {
                const int64_t leader = a->index_leader;
                if (leader != NO_LEADER && leader == j) {
                    behind_obstacle++;
                    a->target = circ(agents, itemsize, leader)->target;
                    has_strategy[i] = 1;
                    break;
                }
                continue;
            }
            if (oracle_herding_relationship(a->position, b->position, a->velocity, b->velocity, phi) & 1) {
                heading_away++;
                a->index_leader = j;
                a->target = NO_TARGET;
                const double rel[2] = {b->position[0] - a->position[0], b->position[1] - a->position[1]};
                double e0[2], e1[2], w[2];
                normalize2(rel, e0); normalize2(b->velocity, e1);
                weighted_average2(e0, e1, weight_position, w);
                normalize2(w, new_direction + 2 * i);
                has_strategy[i] = 1;
                break;
            }
        }
        if (behind_obstacle == 0 && heading_away == 0) {
            const int64_t leader = a->index_leader;
            if (leader != NO_LEADER) { a->target = circ(agents, itemsize, leader)->target; has_strategy[i] = 1; }
        }
    }
    free(dist); free(leaders);
    return ORACLE_OK;
}

/* collective_motion.py:229-243; direction[n][2] */
int oracle_leader_follower_interaction(void *agents, int64_t n, int64_t itemsize, const void *obstacles, int64_t n_obs,
                                       double sight, double phi, double weight_position_leader, double *direction) {
    uint8_t *has_strategy = malloc(n + 1);
    int rc = oracle_leader_follower_interaction_brute(agents, n, itemsize, weight_position_leader, phi, obstacles, n_obs, sight,
                                                      direction, has_strategy);
    if (!rc)
        for (int64_t i = 0; i < n; ++i) {
            agent_circular_t *a = circ(agents, itemsize, i);
            if (!has_strategy[i] && a->is_follower) a->target = a->familiar_exit;
        }
    free(has_strategy);
    return rc;
}

/* collective_motion.py:246-289; direction[n][2] */
int oracle_leader_follower_with_herding_interaction(void *agents, int64_t n, int64_t itemsize, const void *obstacles, int64_t n_obs,
                                                    double sight, int64_t size_nearest_other, double phi, double weight_position_herding,
                                                    double weight_position_leader, double weight_direction_leader, double *direction) {
    if (itemsize != 228 && itemsize != 316) return ORACLE_INVALID_TYPE;
    const double sight_leader = 20.0;
    int64_t *neighbors = malloc(sizeof(int64_t) * (n * size_nearest_other + 1));
    uint8_t *is_follower = malloc(n + 1), *has_direction = malloc(n + 1), *has_strategy = malloc(n + 1);
    double *dir_herding = malloc(sizeof(double) * (2 * n + 2)), *dir_leader = malloc(sizeof(double) * (2 * n + 2));
    int rc = oracle_find_nearest_neighbors(agents, n, itemsize, sight, size_nearest_other, obstacles, n_obs, neighbors);
    for (int64_t i = 0; i < n; ++i) is_follower[i] = circ(agents, itemsize, i)->is_follower;
    if (!rc) rc = oracle_herding_interaction(agents, n, itemsize, is_follower, neighbors, size_nearest_other, weight_position_herding,
                                             phi, dir_herding, has_direction);
    if (!rc) {
        for (int64_t i = 0; i < n; ++i) if (has_direction[i]) circ(agents, itemsize, i)->target = NO_TARGET;
        rc = oracle_leader_follower_interaction_brute(agents, n, itemsize, weight_position_leader, phi, obstacles, n_obs, sight_leader,
                                                      dir_leader, has_strategy);
    }
    if (!rc)
        for (int64_t i = 0; i < n; ++i) {
            agent_circular_t *a = circ(agents, itemsize, i);
            if (!(has_direction[i] || has_strategy[i]) && is_follower[i]) a->target = a->familiar_exit;
            double w[2];
            weighted_average2(dir_leader + 2 * i, dir_herding + 2 * i, weight_direction_leader, w);
            normalize2(w, direction + 2 * i);
        }
    free(neighbors); free(is_follower); free(has_direction); free(has_strategy); free(dir_herding); free(dir_leader);
    return rc;
}

/* =====================================================================================================================
 * InsideDomain / TargetReached (simulation/logic.py:343-387).  Both rest on
 * matplotlib.path.Path(vertices).contains_points(points) with radius 0.  matplotlib is a third-party dependency that is
 * absent here ("parity unpinned"): this restates the crossing-number rule of its point_in_path routine (src/_path.h, after
 * W. Randolph Franklin / Graphics Gems IV "CrossingsMultiplyTest"): an edge (v0, v1) is crossed when the flags
 * (v0.y >= y) and (v1.y >= y) differ and the intersection with the horizontal ray lies on the +x side; the polygon is
 * closed implicitly; the parity of the crossings decides.  Points exactly on an edge are the unpinned part.
 * ===================================================================================================================== */
int oracle_point_in_polygon(const double *vertices, int64_t n_vertices, double tx, double ty) {
    if (n_vertices < 3) return 0;
    int inside = 0;
    double vx0 = vertices[2 * (n_vertices - 1)], vy0 = vertices[2 * (n_vertices - 1) + 1];   /* implicit closing edge first */
    int yflag0 = vy0 >= ty;
    for (int64_t k = 0; k < n_vertices; ++k) {
        const double vx1 = vertices[2 * k], vy1 = vertices[2 * k + 1];
        const int yflag1 = vy1 >= ty;
        if (yflag0 != yflag1)
            if (((vy1 - ty) * (vx0 - vx1) >= (vx1 - tx) * (vy0 - vy1)) == yflag1) inside ^= 1;
        yflag0 = yflag1; vx0 = vx1; vy0 = vy1;
    }
    return inside;
}

/* InsideDomain.update (logic.py:351-357): active = contains(position); returns the number of agents whose flag changed */
int64_t oracle_inside_domain(void *agents, int64_t n, int64_t itemsize, const double *vertices, int64_t n_vertices) {
    if (itemsize != 228 && itemsize != 316) return -1;
    int64_t changed = 0;
    for (int64_t i = 0; i < n; ++i) {
        agent_circular_t *a = circ(agents, itemsize, i);
        const uint8_t now = (uint8_t)oracle_point_in_polygon(vertices, n_vertices, a->position[0], a->position[1]);
        changed += (a->active != 0) != (now != 0);
        a->active = now;
    }
    return changed;
}

/* TargetReached.update (logic.py:383-387) for one target polygon: reached_by |= contains(position); returns sum(reached_by) */
int64_t oracle_target_reached(const void *agents, int64_t n, int64_t itemsize, const double *vertices, int64_t n_vertices, uint8_t *reached_by) {
    if (itemsize != 228 && itemsize != 316) return -1;
    int64_t count = 0;
    for (int64_t i = 0; i < n; ++i) {
        const agent_circular_t *a = circ((void *)agents, itemsize, i);
        if (oracle_point_in_polygon(vertices, n_vertices, a->position[0], a->position[1])) reached_by[i] = 1;
        count += reached_by[i] != 0;
    }
    return count;
}
