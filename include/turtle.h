/*
 * turtle.h -- C ABI of turtle-b200 for the DEM ray-stepping hot path.
 *
 * This header re-declares, with identical names, argument order and meaning, the
 * part of the public interface of niess/turtle v0.11 that lies on the stepping
 * path, so that existing C callers compile and link unchanged. Every prototype
 * cites the declaration it replaces as `ref: include/turtle.h:<line>` (lines in
 * the reference tree). The implementation behind it is new;
 * batched, GPU-resident entry points are declared in turtle_b200.h.
 *
 * turtle_map_load reads the five formats of the reference (.hgt .png .tif .grd .asc)
 * and turtle_map_dump writes .png and .tif, without libpng / libtiff (tb_io.cpp).
 */
#ifndef TURTLE_H
#define TURTLE_H
#ifdef __cplusplus
extern "C" {
#endif

#ifndef TURTLE_API
#define TURTLE_API
#endif

/* Return codes; numeric values are ABI (ref: include/turtle.h:35-62). */
enum turtle_return {
        TURTLE_RETURN_SUCCESS = 0,
        TURTLE_RETURN_BAD_ADDRESS,
        TURTLE_RETURN_BAD_EXTENSION,
        TURTLE_RETURN_BAD_FORMAT,
        TURTLE_RETURN_BAD_PROJECTION,
        TURTLE_RETURN_BAD_JSON,
        TURTLE_RETURN_DOMAIN_ERROR,
        TURTLE_RETURN_LIBRARY_ERROR,
        TURTLE_RETURN_LOCK_ERROR,
        TURTLE_RETURN_MEMORY_ERROR,
        TURTLE_RETURN_PATH_ERROR,
        TURTLE_RETURN_UNLOCK_ERROR,
        N_TURTLE_RETURNS
};

/* Opaque objects (ref: include/turtle.h:67-88). */
struct turtle_projection;
struct turtle_map;
struct turtle_stack;
struct turtle_client;
struct turtle_stepper;

/* Public map description (ref: include/turtle.h:93-106). */
struct turtle_map_info {
        int nx, ny;        /* number of nodes along x and y */
        double x[2];       /* closed x range of the grid */
        double y[2];       /* closed y range of the grid */
        double z[2];       /* elevation span mapped onto 16 bits */
        const char * encoding;
};

/* Callback types (ref: include/turtle.h:113,133-134,149). */
typedef void turtle_function_t(void);
typedef void turtle_error_handler_t(enum turtle_return code,
    turtle_function_t * function, const char * message);
typedef int turtle_stack_locker_t(void);

/* ---- error handling (ref: include/turtle.h:157,168,194) ------------------ */
TURTLE_API const char * turtle_error_function(turtle_function_t * function);
TURTLE_API turtle_error_handler_t * turtle_error_handler_get(void);
TURTLE_API void turtle_error_handler_set(turtle_error_handler_t * handler);

/* ---- projections (ref: include/turtle.h:235-333) ------------------------- */
TURTLE_API enum turtle_return turtle_projection_create(
    struct turtle_projection ** projection, const char * name);
TURTLE_API void turtle_projection_destroy(struct turtle_projection ** projection);
TURTLE_API enum turtle_return turtle_projection_configure(
    struct turtle_projection * projection, const char * name);
TURTLE_API const char * turtle_projection_name(
    const struct turtle_projection * projection);
TURTLE_API enum turtle_return turtle_projection_project(
    const struct turtle_projection * projection, double latitude,
    double longitude, double * x, double * y);
TURTLE_API enum turtle_return turtle_projection_unproject(
    const struct turtle_projection * projection, double x, double y,
    double * latitude, double * longitude);

/* ---- maps (ref: include/turtle.h:362-543) --------------------------------- */
TURTLE_API enum turtle_return turtle_map_create(struct turtle_map ** map,
    const struct turtle_map_info * info, const char * projection);
TURTLE_API void turtle_map_destroy(struct turtle_map ** map);
TURTLE_API enum turtle_return turtle_map_load(
    struct turtle_map ** map, const char * path); /* .hgt .png .tif .grd .asc */
TURTLE_API enum turtle_return turtle_map_dump(
    const struct turtle_map * map, const char * path); /* .png .tif (ref: :423) */
TURTLE_API enum turtle_return turtle_map_fill(
    struct turtle_map * map, int ix, int iy, double elevation);
TURTLE_API enum turtle_return turtle_map_node(const struct turtle_map * map,
    int ix, int iy, double * x, double * y, double * elevation);
TURTLE_API enum turtle_return turtle_map_elevation(
    const struct turtle_map * map, double x, double y, double * elevation,
    int * inside);
TURTLE_API enum turtle_return turtle_map_gradient(
    const struct turtle_map * map, double x, double y, double * gx, double * gy,
    int * inside);
TURTLE_API const struct turtle_projection * turtle_map_projection(
    const struct turtle_map * map);
TURTLE_API void turtle_map_meta(const struct turtle_map * map,
    struct turtle_map_info * info, const char ** projection);

/* ---- ECEF frames (ref: include/turtle.h:556-602) -------------------------- */
TURTLE_API void turtle_ecef_from_geodetic(
    double latitude, double longitude, double elevation, double ecef[3]);
TURTLE_API void turtle_ecef_to_geodetic(const double ecef[3], double * latitude,
    double * longitude, double * altitude);
TURTLE_API void turtle_ecef_from_horizontal(double latitude, double longitude,
    double azimuth, double elevation, double direction[3]);
TURTLE_API void turtle_ecef_to_horizontal(double latitude, double longitude,
    const double direction[3], double * azimuth, double * elevation);

/* ---- tile stacks (ref: include/turtle.h:637-750) -------------------------- */
TURTLE_API enum turtle_return turtle_stack_create(struct turtle_stack ** stack,
    const char * path, int stack_size, turtle_stack_locker_t * lock,
    turtle_stack_locker_t * unlock);
TURTLE_API void turtle_stack_destroy(struct turtle_stack ** stack);
TURTLE_API enum turtle_return turtle_stack_clear(struct turtle_stack * stack);
TURTLE_API enum turtle_return turtle_stack_load(struct turtle_stack * stack);
TURTLE_API enum turtle_return turtle_stack_elevation(
    struct turtle_stack * stack, double latitude, double longitude,
    double * elevation, int * inside);

TURTLE_API enum turtle_return turtle_stack_gradient(
    struct turtle_stack * stack, double latitude, double longitude,
    double * glat, double * glon, int * inside);

/* ---- stack clients (ref: include/turtle.h:773-842) ------------------------ */
TURTLE_API enum turtle_return turtle_client_create(
    struct turtle_client ** client, struct turtle_stack * stack);
TURTLE_API enum turtle_return turtle_client_destroy(
    struct turtle_client ** client);
TURTLE_API enum turtle_return turtle_client_clear(struct turtle_client * client);
TURTLE_API enum turtle_return turtle_client_elevation(
    struct turtle_client * client, double latitude, double longitude,
    double * elevation, int * inside);

/* ---- ECEF stepper (ref: include/turtle.h:859-1155) ------------------------ */
TURTLE_API enum turtle_return turtle_stepper_create(
    struct turtle_stepper ** stepper);
TURTLE_API enum turtle_return turtle_stepper_destroy(
    struct turtle_stepper ** stepper);
TURTLE_API void turtle_stepper_geoid_set(
    struct turtle_stepper * stepper, struct turtle_map * geoid);
TURTLE_API struct turtle_map * turtle_stepper_geoid_get(
    const struct turtle_stepper * stepper);
TURTLE_API void turtle_stepper_reset(struct turtle_stepper * stepper);
TURTLE_API void turtle_stepper_range_set(
    struct turtle_stepper * stepper, double range);
TURTLE_API double turtle_stepper_range_get(const struct turtle_stepper * stepper);
TURTLE_API double turtle_stepper_slope_get(const struct turtle_stepper * stepper);
TURTLE_API void turtle_stepper_slope_set(
    struct turtle_stepper * stepper, double slope);
TURTLE_API double turtle_stepper_resolution_get(
    const struct turtle_stepper * stepper);
TURTLE_API void turtle_stepper_resolution_set(
    struct turtle_stepper * stepper, double resolution);
TURTLE_API enum turtle_return turtle_stepper_add_layer(
    struct turtle_stepper * stepper);
TURTLE_API enum turtle_return turtle_stepper_add_stack(
    struct turtle_stepper * stepper, struct turtle_stack * stack, double offset);
TURTLE_API enum turtle_return turtle_stepper_add_map(
    struct turtle_stepper * stepper, struct turtle_map * map, double offset);
TURTLE_API enum turtle_return turtle_stepper_add_flat(
    struct turtle_stepper * stepper, double ground_level);
TURTLE_API enum turtle_return turtle_stepper_step(
    struct turtle_stepper * stepper, double * position,
    const double * direction, double * latitude, double * longitude,
    double * altitude, double * elevation, double * step, int * index);
TURTLE_API enum turtle_return turtle_stepper_position(
    struct turtle_stepper * stepper, double latitude, double longitude,
    double height, int layer_index, double * position, int * data_index);

#ifdef __cplusplus
}
#endif
#endif
