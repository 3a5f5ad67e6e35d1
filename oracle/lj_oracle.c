/*
 * TEST INFRASTRUCTURE ONLY -- plain-C float64 restatement of mdsea's ContinuousPotentialSolver
 * step loop for the CUT-OFF / large-N variant, which the Python reference cannot run
 * (its pair table is O(N^2): mdsea/simulator.py:122-137; r_cutoff is a TODO: :48-49).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library (through oracle/oracle_c.py).  The product never links it.
 *
 * Build: `make -C oracle` (gcc -O2 -ffp-contract=off -pthread; the image's gcc has no libgomp, so the
 * all-cores timing leg uses a small pthread parallel-for).  -ffp-contract=off matters: the
 * bit-exact contracts below are defined on IEEE float64 operations WITHOUT fused multiply-add.
 *
 * What follows the reference (paths relative to /root/reference/):
 *   one-shot wrap              mdsea/simulator.py:156-166
 *   pair geometry, min image   mdsea/simulator.py:266-275  (dr = r_j - r_i; dr -= rint(dr/L)*L; true division)
 *   strict cutoff  dist < rc   mdsea/simulator.py:283
 *   LJ dV/dr and V             mdsea/potentials.py:43-48 and :15-19 (lambda = 4, t = sigma/sqrt(r*r))
 *   sign / scatter             mdsea/simulator.py:301-309  (a_i += unit*ff/m ; a_j -= unit*ff/m)
 *   verlet order               mdsea/simulator.py:537-544, advance :560-568
 *   energies                   mdsea/simulator.py:241-250  (PE = sum_pairs V / N ; KE = 0.5 m mean |v|^2)
 * Pinning: for r_cutoff >= sqrt(3) L / 2 every pair is inside the cutoff and this code must reproduce
 * the reference-generated golden fixtures (tests/test_oracle_c.py).
 *
 * Canonical definitions fixed by this repo -- the GPU must match them BIT FOR BIT:
 *   cell grid   nc = max(1, floor(L / (rc+skin))) per dim; cell_len = L / nc;
 *               wrapped w = r - floor(r / L) * L ; c_d = min(nc-1, (int)floor(w / cell_len));
 *               linear id = (c_z*nc + c_y)*nc + c_x ; order = stable sort by (cell id, particle id)
 *   list of i   ascending ids j != i with dx*dx + dy*dy + dz*dz < (rc+skin)^2, each d = minimage(r_j - r_i),
 *               sums evaluated left to right, no FMA
 *   rebuild     when max_i |minimage(r_i - r_i^build)|^2 > (skin/2)^2, checked after every drift
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

typedef struct {
    int64_t n;
    double  L, rc, skin, eps, sigma, mass, dt;
    /* state, SoA (3, n) */
    double *r, *v, *a, *rb; /* rb: positions at the last list build */
    /* cell list */
    int     nc;
    double  cell_len;
    int32_t *cell_of, *order, *cell_start; /* cell_start has nc^3+1 entries over `order` */
    /* neighbour list (CSR, full: both i->j and j->i) */
    int64_t *off;
    int32_t *idx;
    int64_t  cap;
    int      have_list;
    /* counters */
    int64_t rebuilds, pairs_unique, force_evals, list_checksum;
    double  pe_sum; /* sum over unique in-cutoff pairs of V at the last force evaluation */
    /* hard walls (simulator.py:168-179): wall_L > 0 means that L above is only the PERIOD of the cell grid / minimum image
     * (the product embeds a walled box [0, wall_L] in a periodic grid of period wall_L + 2 (rc + skin), include/mdsea_b200.h)
     * and that the boundary condition of a step is clamp + reflect at [radius, wall_L - radius] instead of the wrap */
    double  wall_L, radius, restitution;
} lj_sys;

static inline double minimg(double d, double L) { return d - rint(d / L) * L; }

/* ------------------------------------------------------------------------ pthread parallel-for */
static int g_threads = 0;
int oracle_num_threads(void) {
    if (g_threads <= 0) {
        long n = sysconf(_SC_NPROCESSORS_ONLN);
        g_threads = n > 0 ? (int)(n > 256 ? 256 : n) : 1;
    }
    return g_threads;
}
void oracle_set_threads(int n) { g_threads = n > 0 ? n : 1; }

typedef void (*range_fn)(int64_t lo, int64_t hi, int tid, void *ctx);
typedef struct { range_fn fn; void *ctx; int64_t lo, hi; int tid; } pf_job;
static void *pf_tramp(void *p) {
    pf_job *j = (pf_job *)p;
    j->fn(j->lo, j->hi, j->tid, j->ctx);
    return NULL;
}
static void parallel_for(int64_t n, range_fn fn, void *ctx) {
    int T = oracle_num_threads();
    if (n < 4096 || T == 1) { fn(0, n, 0, ctx); return; }
    if (T > 256) T = 256;
    pthread_t th[256];
    pf_job    jobs[256];
    const int64_t chunk = (n + T - 1) / T;
    for (int t = 0; t < T; ++t) {
        jobs[t].fn = fn; jobs[t].ctx = ctx; jobs[t].tid = t;
        jobs[t].lo = t * chunk; jobs[t].hi = (t + 1) * chunk < n ? (t + 1) * chunk : n;
        if (jobs[t].lo >= n) { jobs[t].lo = jobs[t].hi = n; }
        pthread_create(&th[t], NULL, pf_tramp, &jobs[t]);
    }
    for (int t = 0; t < T; ++t) pthread_join(th[t], NULL);
}

/* ---------------------------------------------------------------------------------------- cells */
int oracle_cell_grid(double L, double rlist) {
    int nc = (int)floor(L / rlist);
    return nc < 1 ? 1 : nc;
}

void oracle_cells(int64_t n, const double *r, double L, double rlist, int32_t *cell_of, int32_t *order,
                  int32_t *cell_start /* nc^3+1 or NULL */, int *nc_out) {
    const int    nc = oracle_cell_grid(L, rlist);
    const double cl = L / nc;
    const int64_t ncell = (int64_t)nc * nc * nc;
    int32_t *count = (int32_t *)calloc((size_t)ncell + 1, sizeof(int32_t));
    for (int64_t i = 0; i < n; ++i) {
        int c[3];
        for (int d = 0; d < 3; ++d) {
            const double x = r[d * n + i];
            const double w = x - floor(x / L) * L;
            int k = (int)floor(w / cl);
            if (k > nc - 1) k = nc - 1;
            if (k < 0) k = 0;
            c[d] = k;
        }
        cell_of[i] = (c[2] * nc + c[1]) * nc + c[0];
        count[cell_of[i] + 1]++;
    }
    for (int64_t k = 0; k < ncell; ++k) count[k + 1] += count[k];
    if (cell_start) memcpy(cell_start, count, sizeof(int32_t) * ((size_t)ncell + 1));
    /* counting sort in ascending particle id == stable sort by (cell id, particle id) */
    for (int64_t i = 0; i < n; ++i) order[count[cell_of[i]]++] = (int32_t)i;
    free(count);
    *nc_out = nc;
}

/* ---------------------------------------------------------------------------------- neighbours */
static int cmp_i32(const void *a, const void *b) {
    const int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
    return (x > y) - (x < y);
}

static inline int in_list(const double *r, int64_t n, int64_t i, int64_t j, double L, double rl2) {
    const double dx = minimg(r[j] - r[i], L);
    const double dy = minimg(r[n + j] - r[n + i], L);
    const double dz = minimg(r[2 * n + j] - r[2 * n + i], L);
    const double r2 = dx * dx + dy * dy + dz * dz;
    return r2 < rl2;
}

/* brute force O(N^2): the definition itself.  Two-pass: counts then fill. */
int64_t oracle_neighbors_brute(int64_t n, const double *r, double L, double rlist, int64_t *off, int32_t *idx, int64_t cap) {
    const double rl2 = rlist * rlist;
    int64_t total = 0;
    off[0] = 0;
    for (int64_t i = 0; i < n; ++i) {
        for (int64_t j = 0; j < n; ++j) {
            if (j == i) continue;
            if (in_list(r, n, i, j, L, rl2)) {
                if (idx && total < cap) idx[total] = (int32_t)j;
                ++total;
            }
        }
        off[i + 1] = total;
    }
    return total;
}

typedef struct {
    int64_t n; const double *r; double L, rl2; int nc; const int32_t *cell_of, *order, *cs; int32_t *cnt;
    const int64_t *off; int32_t *idx;
} nb_ctx;
static void nb_range(int64_t lo, int64_t hi, int tid, void *p) {
    nb_ctx *c = (nb_ctx *)p;
    const int nc = c->nc;
    const int64_t n = c->n;
    for (int64_t i = lo; i < hi; ++i) {
        const int cl = c->cell_of[i], cx = cl % nc, cy = (cl / nc) % nc, cz = cl / (nc * nc);
        int64_t w = c->idx ? c->off[i] : 0;
        int m = 0;
        for (int dz = -1; dz <= 1; ++dz)
            for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx) {
                    const int x = (cx + dx + nc) % nc, y = (cy + dy + nc) % nc, z = (cz + dz + nc) % nc;
                    const int cc = (z * nc + y) * nc + x;
                    for (int32_t k = c->cs[cc]; k < c->cs[cc + 1]; ++k) {
                        const int32_t j = c->order[k];
                        if (j != i && in_list(c->r, n, i, j, c->L, c->rl2)) {
                            if (c->idx) c->idx[w++] = j;
                            ++m;
                        }
                    }
                }
        if (c->idx) qsort(c->idx + c->off[i], (size_t)m, sizeof(int32_t), cmp_i32);
        else c->cnt[i] = m;
    }
}

/* cell-list version; identical output (rows ascending in j).  Needs nc >= 3. */
int64_t oracle_neighbors_cells(int64_t n, const double *r, double L, double rlist, int64_t *off, int32_t *idx, int64_t cap) {
    int nc;
    const int    nc0 = oracle_cell_grid(L, rlist);
    const int64_t ncell = (int64_t)nc0 * nc0 * nc0;
    int32_t *cell_of = (int32_t *)malloc(sizeof(int32_t) * n), *order = (int32_t *)malloc(sizeof(int32_t) * n);
    int32_t *cs = (int32_t *)malloc(sizeof(int32_t) * (ncell + 1));
    oracle_cells(n, r, L, rlist, cell_of, order, cs, &nc);
    if (nc < 3) {
        free(cell_of); free(order); free(cs);
        return oracle_neighbors_brute(n, r, L, rlist, off, idx, cap);
    }
    const double rl2 = rlist * rlist;
    int32_t *cnt = (int32_t *)calloc(n, sizeof(int32_t));
    nb_ctx c1 = {n, r, L, rl2, nc, cell_of, order, cs, cnt, off, NULL};
    parallel_for(n, nb_range, &c1); /* pass 1: counts */
    off[0] = 0;
    for (int64_t i = 0; i < n; ++i) off[i + 1] = off[i] + cnt[i];
    const int64_t total = off[n];
    if (idx && total <= cap) {
        c1.idx = idx;
        parallel_for(n, nb_range, &c1); /* pass 2: fill + sort rows */
    }
    free(cnt); free(cell_of); free(order); free(cs);
    return total;
}

/* -------------------------------------------------------------------------------------- system */
lj_sys *oracle_sys_create(int64_t n, double L, double rc, double skin, double eps, double sigma, double mass, double dt,
                          const double *r0, const double *v0) {
    lj_sys *s = (lj_sys *)calloc(1, sizeof(lj_sys));
    s->n = n; s->L = L; s->rc = rc; s->skin = skin; s->eps = eps; s->sigma = sigma; s->mass = mass; s->dt = dt;
    s->r = (double *)malloc(sizeof(double) * 3 * n);
    s->v = (double *)malloc(sizeof(double) * 3 * n);
    s->a = (double *)calloc(3 * n, sizeof(double));
    s->rb = (double *)malloc(sizeof(double) * 3 * n);
    memcpy(s->r, r0, sizeof(double) * 3 * n);
    memcpy(s->v, v0, sizeof(double) * 3 * n);
    s->off = (int64_t *)malloc(sizeof(int64_t) * (n + 1));
    s->cap = 0; s->idx = NULL; s->have_list = 0;
    return s;
}

void oracle_sys_destroy(lj_sys *s) {
    if (!s) return;
    free(s->r); free(s->v); free(s->a); free(s->rb); free(s->off); free(s->idx);
    free(s);
}

static void rebuild_list(lj_sys *s) {
    const double rl = s->rc + s->skin;
    int64_t total = oracle_neighbors_cells(s->n, s->r, s->L, rl, s->off, NULL, 0);
    if (total > s->cap) {
        free(s->idx);
        s->cap = total + total / 8 + 1024;
        s->idx = (int32_t *)malloc(sizeof(int32_t) * s->cap);
    }
    oracle_neighbors_cells(s->n, s->r, s->L, rl, s->off, s->idx, s->cap);
    memcpy(s->rb, s->r, sizeof(double) * 3 * s->n);
    s->have_list = 1;
    s->rebuilds++;
    int64_t cs = 0;
    for (int64_t k = 0; k < total; ++k) cs = cs * 1000003 + s->idx[k];
    s->list_checksum = s->list_checksum * 31 + cs + total;
}

typedef struct { const lj_sys *s; double lim; int flag[256]; } rb_ctx;
static void rb_range(int64_t lo, int64_t hi, int tid, void *p) {
    rb_ctx *c = (rb_ctx *)p;
    const lj_sys *s = c->s;
    const int64_t n = s->n;
    int flag = 0;
    for (int64_t i = lo; i < hi; ++i) {
        const double dx = minimg(s->r[i] - s->rb[i], s->L);
        const double dy = minimg(s->r[n + i] - s->rb[n + i], s->L);
        const double dz = minimg(s->r[2 * n + i] - s->rb[2 * n + i], s->L);
        if (dx * dx + dy * dy + dz * dz > c->lim) flag = 1;
    }
    c->flag[tid] = flag;
}
static int needs_rebuild(const lj_sys *s) {
    if (!s->have_list) return 1;
    rb_ctx c;
    memset(&c, 0, sizeof(c));
    c.s = s;
    c.lim = (0.5 * s->skin) * (0.5 * s->skin);
    parallel_for(s->n, rb_range, &c);
    int flag = 0;
    for (int t = 0; t < 256; ++t) flag |= c.flag[t];
    return flag;
}

/* update_acc with a strict cutoff: every unique pair (i<j) with dist < rc, evaluated like the reference:
 * dist = sqrt(r2); t = sigma / sqrt(dist*dist); ff = 4 eps dist (6 t^6 - 12 t^12) / (dist*dist);
 * acc_p = (dr / dist) * ff / m ; V = 4 eps (t^12 - t^6). */
typedef struct { lj_sys *s; double pe[256]; int64_t np[256]; } f_ctx;
static void f_range(int64_t lo, int64_t hi, int tid, void *p) {
    f_ctx *c = (f_ctx *)p;
    lj_sys *s = c->s;
    const int64_t n = s->n;
    const double  L = s->L, rc = s->rc, eps = s->eps, sig = s->sigma, m = s->mass;
    double  pe = 0.0;
    int64_t np = 0;
    for (int64_t i = lo; i < hi; ++i) {
        double ax = 0, ay = 0, az = 0;
        for (int64_t k = s->off[i]; k < s->off[i + 1]; ++k) {
            const int64_t j = s->idx[k];
            double dx = minimg(s->r[j] - s->r[i], L);
            double dy = minimg(s->r[n + j] - s->r[n + i], L);
            double dz = minimg(s->r[2 * n + j] - s->r[2 * n + i], L);
            const double dist = sqrt(dx * dx + dy * dy + dz * dz);
            if (!(dist < rc)) continue;
            const double aprs = dist * dist;
            const double t = sig / sqrt(aprs);
            const double t6 = pow(t, 6), t12 = pow(t, 12);
            const double ff = 4.0 * eps * dist * (6.0 * t6 - 12.0 * t12) / aprs;
            const double sc = ff / m / dist;
            ax += dx * sc; ay += dy * sc; az += dz * sc;
            if (j > i) { pe += 4.0 * eps * (t12 - t6); ++np; }
        }
        s->a[i] = ax; s->a[n + i] = ay; s->a[2 * n + i] = az;
    }
    c->pe[tid] = pe;
    c->np[tid] = np;
}
static void forces(lj_sys *s, int after_drift) {
    /* the skin trigger is evaluated once per drift (and when there is no list yet) */
    if (!s->have_list || (after_drift && needs_rebuild(s))) rebuild_list(s);
    f_ctx c;
    memset(&c, 0, sizeof(c));
    c.s = s;
    parallel_for(s->n, f_range, &c);
    double  pe = 0.0;
    int64_t np = 0;
    for (int t = 0; t < 256; ++t) { pe += c.pe[t]; np += c.np[t]; }
    s->pe_sum = pe;
    s->pairs_unique += np;
    s->force_evals++;
}

void oracle_sys_forces(lj_sys *s, double *acc_out, double *mean_pe) {
    forces(s, 1);
    if (acc_out) memcpy(acc_out, s->a, sizeof(double) * 3 * s->n);
    if (mean_pe) *mean_pe = s->pe_sum / (double)s->n;
}

/* one verlet step (simulator.py:560-568 with :537-544), two force evaluations like the reference. */
void oracle_sys_step(lj_sys *s, double *pe_out, double *ke_out) {
    const int64_t n = s->n, E = 3 * n;
    const double  L = s->L, dt = s->dt;
    if (s->wall_L > 0) {
        for (int64_t e = 0; e < E; ++e) { /* apply_hbc, simulator.py:168-179: two tests in this order */
            if (s->r[e] - s->radius < 0) { s->r[e] = s->radius; s->v[e] *= -s->restitution; }
            if (s->r[e] + s->radius > s->wall_L) { s->r[e] = s->wall_L - s->radius; s->v[e] *= -s->restitution; }
        }
    } else {
        for (int64_t e = 0; e < E; ++e) { /* apply_pbc */
            if (s->r[e] < 0) s->r[e] += L;
            if (s->r[e] > L) s->r[e] -= L;
        }
    }
    forces(s, 0);
    for (int64_t e = 0; e < E; ++e) {
        s->v[e] += 0.5 * s->a[e] * dt;
        s->r[e] += s->v[e] * dt;
    }
    forces(s, 1);
    double ke = 0.0;
    for (int64_t e = 0; e < E; ++e) {
        s->v[e] += 0.5 * s->a[e] * dt;
        ke += s->v[e] * s->v[e];
    }
    if (pe_out) *pe_out = s->pe_sum / (double)n;
    if (ke_out) *ke_out = 0.5 * s->mass * ke / (double)n;
}

/* switch a system created with L = wall_L + 2 (rc + skin) to hard walls */
void oracle_sys_set_walls(lj_sys *s, double wall_L, double radius, double restitution) {
    s->wall_L = wall_L; s->radius = radius; s->restitution = restitution;
}

void oracle_sys_get(const lj_sys *s, double *r, double *v, double *a) {
    if (r) memcpy(r, s->r, sizeof(double) * 3 * s->n);
    if (v) memcpy(v, s->v, sizeof(double) * 3 * s->n);
    if (a) memcpy(a, s->a, sizeof(double) * 3 * s->n);
}
void oracle_sys_counters(const lj_sys *s, int64_t *out /* rebuilds, pairs_unique, force_evals, checksum, nbr_entries */) {
    out[0] = s->rebuilds; out[1] = s->pairs_unique; out[2] = s->force_evals; out[3] = s->list_checksum;
    out[4] = s->have_list ? s->off[s->n] : 0;
}
/* current neighbour list (CSR over particle ids, rows ascending) */
int64_t oracle_sys_list(const lj_sys *s, int64_t *off, int32_t *idx) {
    if (!s->have_list) return 0;
    memcpy(off, s->off, sizeof(int64_t) * (s->n + 1));
    if (idx) memcpy(idx, s->idx, sizeof(int32_t) * s->off[s->n]);
    return s->off[s->n];
}
