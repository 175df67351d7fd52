/*
 * trace_driver.c -- TEST INFRASTRUCTURE (oracle side), not product code.
 *
 * A pthread harness written against the public turtle.h interface only. It
 * dlopen()s ANY library exporting that interface -- the unmodified reference
 * (oracle/_ref/libturtle_ref.so), the plain-C restatement (oracle/liboracle.so)
 * or the product's own scalar calls -- and drives it the way the reference's
 * examples do:
 *   - the canonical ray loop of examples/example-stepper.c:102-140, made total by
 *     the same stop rule as turtle_stepper_trace_batch (include/turtle_b200.h);
 *   - one stepper per thread over shared read-only maps, a mutex-locked stack
 *     (=> one turtle_client per stepper), examples/example-pthread.c:66-114;
 *   - turtle_stepper_reset() before every ray, which makes per-ray results
 *     independent of ray order.
 * It is used by tests/ as the parity checker and by bench.py for the
 * `cpu_baseline` / `--impl reference` legs.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ---- the slice of turtle.h that is driven (ref: include/turtle.h) ---------- */
struct turtle_map;
struct turtle_stack;
struct turtle_stepper;
struct turtle_projection;
struct turtle_map_info {
        int nx, ny;
        double x[2], y[2], z[2];
        const char * encoding;
};
typedef void turtle_function_t(void);
typedef void turtle_error_handler_t(int, turtle_function_t *, const char *);
typedef int turtle_stack_locker_t(void);

struct api {
        void * dl;
        void (*error_handler_set)(turtle_error_handler_t *);
        int (*map_create)(struct turtle_map **, const struct turtle_map_info *, const char *);
        void (*map_destroy)(struct turtle_map **);
        int (*map_fill)(struct turtle_map *, int, int, double);
        int (*map_node)(const struct turtle_map *, int, int, double *, double *, double *);
        int (*map_elevation)(const struct turtle_map *, double, double, double *, int *);
        int (*map_gradient)(const struct turtle_map *, double, double, double *, double *, int *);
        int (*stack_gradient)(struct turtle_stack *, double, double, double *, double *, int *);
        int (*stack_create)(struct turtle_stack **, const char *, int,
            turtle_stack_locker_t *, turtle_stack_locker_t *);
        void (*stack_destroy)(struct turtle_stack **);
        int (*stack_load)(struct turtle_stack *);
        int (*stack_elevation)(struct turtle_stack *, double, double, double *, int *);
        int (*stepper_create)(struct turtle_stepper **);
        int (*stepper_destroy)(struct turtle_stepper **);
        void (*stepper_geoid_set)(struct turtle_stepper *, struct turtle_map *);
        void (*stepper_range_set)(struct turtle_stepper *, double);
        void (*stepper_slope_set)(struct turtle_stepper *, double);
        void (*stepper_resolution_set)(struct turtle_stepper *, double);
        void (*stepper_reset)(struct turtle_stepper *);
        int (*stepper_add_layer)(struct turtle_stepper *);
        int (*stepper_add_flat)(struct turtle_stepper *, double);
        int (*stepper_add_map)(struct turtle_stepper *, struct turtle_map *, double);
        int (*stepper_add_stack)(struct turtle_stepper *, struct turtle_stack *, double);
        int (*stepper_step)(struct turtle_stepper *, double *, const double *, double *,
            double *, double *, double *, double *, int *);
        int (*stepper_position)(struct turtle_stepper *, double, double, double, int,
            double *, int *);
        void (*ecef_from_geodetic)(double, double, double, double *);
        void (*ecef_to_geodetic)(const double *, double *, double *, double *);
        void (*ecef_from_horizontal)(double, double, double, double, double *);
        void (*ecef_to_horizontal)(double, double, const double *, double *, double *);
        int (*projection_create)(struct turtle_projection **, const char *);
        void (*projection_destroy)(struct turtle_projection **);
        int (*projection_project)(const struct turtle_projection *, double, double,
            double *, double *);
        int (*projection_unproject)(const struct turtle_projection *, double, double,
            double *, double *);
};

enum { TD_ADD_LAYER = 0, TD_ADD_FLAT = 1, TD_ADD_MAP = 2, TD_ADD_STACK = 3 };
#define TD_MAX_OBJECTS 64
#define TD_MAX_OPS 64

struct td_op {
        int kind;
        int ref;
        double offset;
};

/* Same layout as struct turtle_trace_rule / turtle_trace_result (turtle_b200.h). */
struct td_rule {
        double altitude_min, altitude_max, length_max;
        int32_t max_steps, reserved;
};
struct td_result {
        double position[3];
        double altitude;
        double length[4];
        double total;
        int32_t n_steps, status;
        int32_t index[2];
        uint32_t medium_hash;
        int32_t n_changes;
};
typedef char td_result_is_96_bytes[(sizeof(struct td_result) == 96) ? 1 : -1];

/* Same layout as struct turtle_trace_crossing (turtle_b200.h): one medium change. */
struct td_crossing {
        double length; /* path length from the origin at which the new medium begins */
        int32_t from, to;
};

struct td_handle {
        struct api api;
        struct turtle_map * maps[TD_MAX_OBJECTS];
        int n_maps;
        struct turtle_stack * stacks[TD_MAX_OBJECTS];
        int n_stacks;
        struct td_op ops[TD_MAX_OPS];
        int n_ops;
        int geoid;
        double range, slope, resolution;
        char last_error[1024];
};

static pthread_mutex_t td_stack_mutex = PTHREAD_MUTEX_INITIALIZER;
static int td_lock(void) { return pthread_mutex_lock(&td_stack_mutex); }
static int td_unlock(void) { return pthread_mutex_unlock(&td_stack_mutex); }

static __thread char td_thread_error[1024];
static void td_on_error(int code, turtle_function_t * fn, const char * message)
{
        snprintf(td_thread_error, sizeof td_thread_error, "%s", message);
}

#define LOAD(field, name)                                                          \
        do {                                                                       \
                *(void **)(&h->api.field) = dlsym(h->api.dl, name);                \
                if (h->api.field == NULL) {                                        \
                        fprintf(stderr, "trace_driver: missing symbol %s in %s\n", \
                            name, path);                                           \
                        dlclose(h->api.dl);                                        \
                        free(h);                                                   \
                        return NULL;                                               \
                }                                                                  \
        } while (0)

struct td_handle * td_open(const char * path)
{
        struct td_handle * h = calloc(1, sizeof(*h));
        if (h == NULL) return NULL;
        h->api.dl = dlopen(path, RTLD_NOW | RTLD_LOCAL);
        if (h->api.dl == NULL) {
                fprintf(stderr, "trace_driver: %s\n", dlerror());
                free(h);
                return NULL;
        }
        LOAD(error_handler_set, "turtle_error_handler_set");
        LOAD(map_create, "turtle_map_create");
        LOAD(map_destroy, "turtle_map_destroy");
        LOAD(map_fill, "turtle_map_fill");
        LOAD(map_node, "turtle_map_node");
        LOAD(map_elevation, "turtle_map_elevation");
        LOAD(map_gradient, "turtle_map_gradient");
        LOAD(stack_gradient, "turtle_stack_gradient");
        LOAD(stack_create, "turtle_stack_create");
        LOAD(stack_destroy, "turtle_stack_destroy");
        LOAD(stack_load, "turtle_stack_load");
        LOAD(stack_elevation, "turtle_stack_elevation");
        LOAD(stepper_create, "turtle_stepper_create");
        LOAD(stepper_destroy, "turtle_stepper_destroy");
        LOAD(stepper_geoid_set, "turtle_stepper_geoid_set");
        LOAD(stepper_range_set, "turtle_stepper_range_set");
        LOAD(stepper_slope_set, "turtle_stepper_slope_set");
        LOAD(stepper_resolution_set, "turtle_stepper_resolution_set");
        LOAD(stepper_reset, "turtle_stepper_reset");
        LOAD(stepper_add_layer, "turtle_stepper_add_layer");
        LOAD(stepper_add_flat, "turtle_stepper_add_flat");
        LOAD(stepper_add_map, "turtle_stepper_add_map");
        LOAD(stepper_add_stack, "turtle_stepper_add_stack");
        LOAD(stepper_step, "turtle_stepper_step");
        LOAD(stepper_position, "turtle_stepper_position");
        LOAD(ecef_from_geodetic, "turtle_ecef_from_geodetic");
        LOAD(ecef_to_geodetic, "turtle_ecef_to_geodetic");
        LOAD(ecef_from_horizontal, "turtle_ecef_from_horizontal");
        LOAD(ecef_to_horizontal, "turtle_ecef_to_horizontal");
        LOAD(projection_create, "turtle_projection_create");
        LOAD(projection_destroy, "turtle_projection_destroy");
        LOAD(projection_project, "turtle_projection_project");
        LOAD(projection_unproject, "turtle_projection_unproject");
        h->api.error_handler_set(&td_on_error);
        h->geoid = -1;
        h->range = 1.;
        h->slope = 0.4;
        h->resolution = 1E-02;
        return h;
}

void td_close(struct td_handle * h)
{
        if (h == NULL) return;
        for (int i = 0; i < h->n_maps; i++) h->api.map_destroy(&h->maps[i]);
        for (int i = 0; i < h->n_stacks; i++) h->api.stack_destroy(&h->stacks[i]);
        dlclose(h->api.dl);
        free(h);
}

const char * td_last_error(struct td_handle * h)
{
        snprintf(h->last_error, sizeof h->last_error, "%s", td_thread_error);
        return h->last_error;
}

/* Create a map with turtle_map_create and fill every node from values[iy*nx+ix]. */
int td_map_create(struct td_handle * h, int nx, int ny, double x0, double x1, double y0,
    double y1, double z0, double z1, const char * projection, const double * values)
{
        if (h->n_maps >= TD_MAX_OBJECTS) return -1;
        struct turtle_map_info info = { nx, ny, { x0, x1 }, { y0, y1 }, { z0, z1 }, NULL };
        struct turtle_map * map = NULL;
        if ((h->api.map_create(&map, &info, projection) != 0) || (map == NULL)) return -1;
        for (int iy = 0; iy < ny; iy++)
                for (int ix = 0; ix < nx; ix++)
                        if (h->api.map_fill(map, ix, iy, values[(size_t)iy * nx + ix]) != 0) {
                                h->api.map_destroy(&map);
                                return -1;
                        }
        h->maps[h->n_maps] = map;
        return h->n_maps++;
}

/* Create a stack over a directory of tiles and load every tile up front. With
 * `locked` the stack gets mutex lock/unlock callbacks, so that each stepper wraps
 * it in its own client (ref: stepper.c:432-436). */
int td_stack_create(struct td_handle * h, const char * path, int locked)
{
        if (h->n_stacks >= TD_MAX_OBJECTS) return -1;
        struct turtle_stack * stack = NULL;
        if ((h->api.stack_create(&stack, path, 0, locked ? &td_lock : NULL,
                 locked ? &td_unlock : NULL) != 0) ||
            (stack == NULL))
                return -1;
        if (h->api.stack_load(stack) != 0) {
                h->api.stack_destroy(&stack);
                return -1;
        }
        h->stacks[h->n_stacks] = stack;
        return h->n_stacks++;
}

int td_geometry(struct td_handle * h, const struct td_op * ops, int n_ops, int geoid,
    double range, double slope, double resolution)
{
        if (n_ops > TD_MAX_OPS) return -1;
        memcpy(h->ops, ops, n_ops * sizeof(*ops));
        h->n_ops = n_ops;
        h->geoid = geoid;
        h->range = range;
        h->slope = slope;
        h->resolution = resolution;
        return 0;
}

static struct turtle_stepper * build_stepper(struct td_handle * h)
{
        struct turtle_stepper * s = NULL;
        if (h->api.stepper_create(&s) != 0) return NULL;
        if (h->geoid >= 0) h->api.stepper_geoid_set(s, h->maps[h->geoid]);
        h->api.stepper_slope_set(s, h->slope);
        h->api.stepper_resolution_set(s, h->resolution);
        h->api.stepper_range_set(s, h->range);
        for (int i = 0; i < h->n_ops; i++) {
                const struct td_op * op = &h->ops[i];
                int rc = 0;
                if (op->kind == TD_ADD_LAYER)
                        rc = h->api.stepper_add_layer(s);
                else if (op->kind == TD_ADD_FLAT)
                        rc = h->api.stepper_add_flat(s, op->offset);
                else if (op->kind == TD_ADD_MAP)
                        rc = h->api.stepper_add_map(s, h->maps[op->ref], op->offset);
                else
                        rc = h->api.stepper_add_stack(s, h->stacks[op->ref], op->offset);
                if (rc != 0) {
                        h->api.stepper_destroy(&s);
                        return NULL;
                }
        }
        return s;
}

/* ---- the ray loop -------------------------------------------------------------- */

static int finite3(const double * v) { return isfinite(v[0]) && isfinite(v[1]) && isfinite(v[2]); }

static void trace_one(struct td_handle * h, struct turtle_stepper * s, const double * position,
    const double * direction, const struct td_rule * rule, struct td_result * R,
    uint64_t * steps, struct td_crossing * crossings, int max_crossings)
{
        memset(R, 0x0, sizeof(*R));
        double pos[3] = { position[0], position[1], position[2] };
        if (!finite3(position) || !finite3(direction)) {
                memcpy(R->position, pos, sizeof pos);
                R->status = 4;
                R->index[0] = R->index[1] = -1;
                return;
        }
        h->api.stepper_reset(s);
        double altitude = 0.;
        int index[2] = { -1, -1 };
        h->api.stepper_step(s, pos, NULL, NULL, NULL, &altitude, NULL, NULL, index);
        uint32_t hash = (2166136261u ^ (uint32_t)(index[0] + 1)) * 16777619u;
        double total = 0.;
        int n = 0, changes = 0, status;
        for (;;) {
                if (index[0] < 0) {
                        status = 1;
                        break;
                } else if (!(altitude < rule->altitude_max) || !(altitude > rule->altitude_min)) {
                        status = 0;
                        break;
                } else if (total >= rule->length_max) {
                        status = 2;
                        break;
                } else if (n >= rule->max_steps) {
                        status = 3;
                        break;
                }
                const int m0 = index[0];
                double step = 0.;
                h->api.stepper_step(s, pos, direction, NULL, NULL, &altitude, NULL, &step, index);
                R->length[(m0 < 3) ? m0 : 3] += step;
                total += step;
                n++;
                if (index[0] != m0) {
                        if ((crossings != NULL) && (changes < max_crossings)) {
                                crossings[changes].length = total;
                                crossings[changes].from = m0;
                                crossings[changes].to = index[0];
                        }
                        changes++;
                        hash = (hash ^ (uint32_t)(index[0] + 1)) * 16777619u;
                }
        }
        memcpy(R->position, pos, sizeof pos);
        R->altitude = altitude;
        R->total = total;
        R->n_steps = n;
        R->status = status;
        R->index[0] = index[0];
        R->index[1] = index[1];
        R->medium_hash = hash;
        R->n_changes = changes;
        *steps += (uint64_t)n;
}

struct trace_job {
        struct td_handle * h;
        size_t n, first, stride;
        const double * position;
        const double * direction;
        const struct td_rule * rule;
        struct td_result * results;
        uint64_t steps;
        int failed;
        struct td_crossing * crossings; /* [n][max_crossings] or NULL */
        int max_crossings;
};

static void * trace_thread(void * arg)
{
        struct trace_job * job = arg;
        struct turtle_stepper * s = build_stepper(job->h);
        if (s == NULL) {
                job->failed = 1;
                return NULL;
        }
        for (size_t i = job->first; i < job->n; i += job->stride)
                trace_one(job->h, s, job->position + 3 * i, job->direction + 3 * i, job->rule,
                    job->results + i, &job->steps,
                    (job->crossings != NULL) ? job->crossings + i * (size_t)job->max_crossings : NULL,
                    job->max_crossings);
        job->h->api.stepper_destroy(&s);
        return NULL;
}

static double now_seconds(void)
{
        struct timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* Trace n rays on `threads` pthreads (thread t takes rays t, t+T, ...). Returns the
 * number of steps, or -1; *seconds = wall time from thread creation to join. */
long long td_trace_crossings(struct td_handle * h, size_t n, const double * position,
    const double * direction, const struct td_rule * rule, struct td_result * results,
    struct td_crossing * crossings, int max_crossings, int threads, double * seconds);

long long td_trace(struct td_handle * h, size_t n, const double * position,
    const double * direction, const struct td_rule * rule, struct td_result * results,
    int threads, double * seconds)
{
        return td_trace_crossings(h, n, position, direction, rule, results, NULL, 0, threads,
            seconds);
}

/* ... and the first `max_crossings` medium changes of every ray. */
long long td_trace_crossings(struct td_handle * h, size_t n, const double * position,
    const double * direction, const struct td_rule * rule, struct td_result * results,
    struct td_crossing * crossings, int max_crossings, int threads, double * seconds)
{
        if (threads < 1) threads = 1;
        struct trace_job * jobs = calloc(threads, sizeof(*jobs));
        pthread_t * tids = calloc(threads, sizeof(*tids));
        const double t0 = now_seconds();
        for (int t = 0; t < threads; t++) {
                struct trace_job j = { h, n, (size_t)t, (size_t)threads, position, direction,
                        rule, results, 0, 0, crossings, max_crossings };
                jobs[t] = j;
                if (threads == 1)
                        trace_thread(&jobs[t]);
                else
                        pthread_create(&tids[t], NULL, &trace_thread, &jobs[t]);
        }
        long long steps = 0;
        for (int t = 0; t < threads; t++) {
                if (threads > 1) pthread_join(tids[t], NULL);
                if (jobs[t].failed) steps = -1;
                if (steps >= 0) steps += (long long)jobs[t].steps;
        }
        if (seconds != NULL) *seconds = now_seconds() - t0;
        free(jobs);
        free(tids);
        return steps;
}

/* ---- particle walks: n_steps calls of turtle_stepper_step per particle with a
 * fresh direction each (BASELINE.json config 4). direction[(k*n + i)*3] is the
 * direction of particle i at iteration k. All outputs are [k*n + i] as well. */
struct walk_job {
        struct td_handle * h;
        size_t n, first, stride;
        int n_steps;
        double * position;
        const double * direction;
        double * step;
        double * altitude;
        int * index;
        int failed;
};

static void * walk_thread(void * arg)
{
        struct walk_job * job = arg;
        struct turtle_stepper * s = build_stepper(job->h);
        if (s == NULL) {
                job->failed = 1;
                return NULL;
        }
        for (size_t i = job->first; i < job->n; i += job->stride) {
                job->h->api.stepper_reset(s);
                double * pos = job->position + 3 * i;
                for (int k = 0; k < job->n_steps; k++) {
                        const size_t o = (size_t)k * job->n + i;
                        double step = 0., altitude = 0.;
                        int index[2] = { -1, -1 };
                        job->h->api.stepper_step(s, pos, job->direction + 3 * o, NULL, NULL,
                            &altitude, NULL, &step, index);
                        if (job->step != NULL) job->step[o] = step;
                        if (job->altitude != NULL) job->altitude[o] = altitude;
                        if (job->index != NULL) {
                                job->index[2 * o] = index[0];
                                job->index[2 * o + 1] = index[1];
                        }
                }
        }
        job->h->api.stepper_destroy(&s);
        return NULL;
}

int td_walk(struct td_handle * h, size_t n, int n_steps, double * position,
    const double * direction, double * step, double * altitude, int * index, int threads,
    double * seconds)
{
        if (threads < 1) threads = 1;
        struct walk_job * jobs = calloc(threads, sizeof(*jobs));
        pthread_t * tids = calloc(threads, sizeof(*tids));
        const double t0 = now_seconds();
        for (int t = 0; t < threads; t++) {
                struct walk_job j = { h, n, (size_t)t, (size_t)threads, n_steps, position,
                        direction, step, altitude, index, 0 };
                jobs[t] = j;
                if (threads == 1)
                        walk_thread(&jobs[t]);
                else
                        pthread_create(&tids[t], NULL, &walk_thread, &jobs[t]);
        }
        int rc = 0;
        for (int t = 0; t < threads; t++) {
                if (threads > 1) pthread_join(tids[t], NULL);
                if (jobs[t].failed) rc = -1;
        }
        if (seconds != NULL) *seconds = now_seconds() - t0;
        free(jobs);
        free(tids);
        return rc;
}

/* ---- single calls, vectorised for ctypes --------------------------------------- */

/* One turtle_stepper_step per entry on a FRESH (reset) stepper; direction may be NULL. */
int td_step(struct td_handle * h, size_t n, double * position, const double * direction,
    double * latitude, double * longitude, double * altitude, double * elevation,
    double * step, int * index)
{
        struct turtle_stepper * s = build_stepper(h);
        if (s == NULL) return -1;
        for (size_t i = 0; i < n; i++) {
                h->api.stepper_reset(s);
                double la, lo, al, el[2], st;
                int idx[2];
                h->api.stepper_step(s, position + 3 * i,
                    (direction != NULL) ? direction + 3 * i : NULL, &la, &lo, &al, el, &st, idx);
                if (latitude != NULL) latitude[i] = la;
                if (longitude != NULL) longitude[i] = lo;
                if (altitude != NULL) altitude[i] = al;
                if (elevation != NULL) {
                        elevation[2 * i] = el[0];
                        elevation[2 * i + 1] = el[1];
                }
                if (step != NULL) step[i] = st;
                if (index != NULL) {
                        index[2 * i] = idx[0];
                        index[2 * i + 1] = idx[1];
                }
        }
        h->api.stepper_destroy(&s);
        return 0;
}

int td_position(struct td_handle * h, size_t n, const double * latitude,
    const double * longitude, const double * height, int layer, double * position,
    int * data_index)
{
        struct turtle_stepper * s = build_stepper(h);
        if (s == NULL) return -1;
        for (size_t i = 0; i < n; i++)
                h->api.stepper_position(s, latitude[i], longitude[i], height[i], layer,
                    position + 3 * i, data_index + i);
        h->api.stepper_destroy(&s);
        return 0;
}

void td_ecef_to_geodetic(struct td_handle * h, size_t n, const double * ecef, double * latitude,
    double * longitude, double * altitude)
{
        for (size_t i = 0; i < n; i++)
                h->api.ecef_to_geodetic(ecef + 3 * i, latitude + i, longitude + i, altitude + i);
}

void td_ecef_from_geodetic(struct td_handle * h, size_t n, const double * latitude,
    const double * longitude, const double * elevation, double * ecef)
{
        for (size_t i = 0; i < n; i++)
                h->api.ecef_from_geodetic(latitude[i], longitude[i], elevation[i], ecef + 3 * i);
}

void td_ecef_from_horizontal(struct td_handle * h, size_t n, const double * latitude,
    const double * longitude, const double * azimuth, const double * elevation,
    double * direction)
{
        for (size_t i = 0; i < n; i++)
                h->api.ecef_from_horizontal(
                    latitude[i], longitude[i], azimuth[i], elevation[i], direction + 3 * i);
}

void td_ecef_to_horizontal(struct td_handle * h, size_t n, const double * latitude,
    const double * longitude, const double * direction, double * azimuth, double * elevation)
{
        for (size_t i = 0; i < n; i++)
                h->api.ecef_to_horizontal(
                    latitude[i], longitude[i], direction + 3 * i, azimuth + i, elevation + i);
}

int td_project(struct td_handle * h, const char * name, int inverse, size_t n, const double * a,
    const double * b, double * c, double * d)
{
        struct turtle_projection * p = NULL;
        if ((h->api.projection_create(&p, name) != 0) || (p == NULL)) return -1;
        for (size_t i = 0; i < n; i++) {
                if (inverse)
                        h->api.projection_unproject(p, a[i], b[i], c + i, d + i);
                else
                        h->api.projection_project(p, a[i], b[i], c + i, d + i);
        }
        h->api.projection_destroy(&p);
        return 0;
}

void td_map_elevation(struct td_handle * h, int map, size_t n, const double * x, const double * y,
    double * z, int * inside)
{
        for (size_t i = 0; i < n; i++)
                h->api.map_elevation(h->maps[map], x[i], y[i], z + i, inside + i);
}

void td_map_node(struct td_handle * h, int map, size_t n, const int * ix, const int * iy,
    double * x, double * y, double * z)
{
        for (size_t i = 0; i < n; i++)
                h->api.map_node(h->maps[map], ix[i], iy[i], x + i, y + i, z + i);
}

void td_stack_elevation(struct td_handle * h, int stack, size_t n, const double * latitude,
    const double * longitude, double * z, int * inside)
{
        for (size_t i = 0; i < n; i++)
                h->api.stack_elevation(h->stacks[stack], latitude[i], longitude[i], z + i,
                    inside + i);
}

/* gx / gy keep their input value where the reference leaves them untouched */
void td_map_gradient(struct td_handle * h, int map, size_t n, const double * x, const double * y,
    double * gx, double * gy, int * inside)
{
        for (size_t i = 0; i < n; i++)
                h->api.map_gradient(h->maps[map], x[i], y[i], gx + i, gy + i, inside + i);
}

void td_stack_gradient(struct td_handle * h, int stack, size_t n, const double * latitude,
    const double * longitude, double * glat, double * glon, int * inside)
{
        for (size_t i = 0; i < n; i++)
                h->api.stack_gradient(h->stacks[stack], latitude[i], longitude[i], glat + i,
                    glon + i, inside + i);
}

/* Raw object access for tests that drive the product's batch API on the same maps. */
void * td_map_pointer(struct td_handle * h, int map) { return h->maps[map]; }
void * td_stack_pointer(struct td_handle * h, int stack) { return h->stacks[stack]; }
