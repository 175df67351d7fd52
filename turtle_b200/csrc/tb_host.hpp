/*
 * tb_host.hpp -- host-side object model behind the opaque types of turtle.h.
 *
 * The reference keeps linked lists of heap nodes (stepper.h:45-110, stack.h:34-52,
 * map.h:41-71). Here every object is a plain C++ aggregate whose only job is to
 * describe a geometry that can be FLATTENED (tb::Geometry) and uploaded; the
 * scalar calls of turtle.h run the same tb_core.cuh functions on the host copy.
 */
#pragma once

#include <stdint.h>
#include <string>
#include <vector>

#include "tb_core.cuh"
#include "turtle.h"
#include "turtle_b200.h"

/* Mirror of a map on one CUDA device (created lazily by the batch calls). */
struct tb_device_mirror {
        int device = -1;
        uint16_t * nodes = nullptr;
        int pitch = 0;
        uint64_t version = 0;
        /* cell-packed second copy (tb::NodesPacked), built when the map asks for it */
        void * packed = nullptr;
        uint64_t packed_version = 0;
};

/* ref: struct turtle_projection, projection.h:37-47 */
struct turtle_projection {
        int type; /* -1 none, 0 Lambert, 1 UTM (enum projection_type) */
        double utm_longitude_0;
        int utm_hemisphere;
        int lambert_tag;
        char tag[64];
};

/* ref: struct turtle_map + turtle_map_meta, map.h:41-71. Nodes are stored row
 * major, SOUTH first, native endian whatever the file format was. */
struct turtle_map {
        int nx, ny;
        double x0, y0, z0, dx, dy, dz;
        int kind; /* tb::NodeKind */
        char encoding[8];
        struct turtle_projection projection;
        std::vector<uint16_t> nodes;
        struct turtle_stack * stack;
        uint64_t version; /* bumped by turtle_map_fill */
        int gather = 0;   /* turtle_map_gather_set: 0 row-major gathers, 1 cell-packed copy */
        std::vector<tb_device_mirror> mirrors;
};

/* ref: struct turtle_stack, stack.h:34-52 */
struct turtle_stack {
        int max_size;
        turtle_stack_locker_t * lock;
        turtle_stack_locker_t * unlock;
        double latitude_0, latitude_delta, longitude_0, longitude_delta;
        int latitude_n, longitude_n;
        std::string root;
        std::vector<std::string> path;        /* per cell, empty = no file */
        std::vector<struct turtle_map *> tile; /* per cell, NULL = not loaded */
        std::vector<tb::MapDesc> header;       /* per cell: shape and origin from the file
                                                * header (no nodes), read once at creation */
        std::vector<int> mru;                  /* loaded cells, most recent first */
};

/* ref: struct turtle_client, client.h:31-40 */
struct turtle_client {
        struct turtle_stack * stack;
};

struct tb_stepper_data {
        int kind; /* tb::DataKind */
        struct turtle_map * map;
        struct turtle_stack * stack;
        struct turtle_client * client; /* owned, when the stack has a lock */
        int transform;
};

struct tb_stepper_meta {
        int data;
        double offset;
};

struct tb_stepper_transform {
        std::string name;
        tb::ProjDesc proj;
};

/* A host-side flattening of a stepper: what gets uploaded by freeze. */
struct tb_flat_geometry {
        tb::Geometry G;
        std::vector<tb::MapDesc> maps;        /* nodes point to HOST memory */
        std::vector<struct turtle_map *> src; /* the map behind each descriptor, or NULL: */
        std::vector<std::string> file;        /* ... the tile file to ingest on the device */
        std::vector<tb::TileRec> tiles;
        size_t skipped = 0;                   /* tiles left out by the residency region */
};

/* ref: struct turtle_stepper, stepper.h:101-110 */
struct turtle_stepper {
        std::vector<std::vector<tb_stepper_meta> > layers; /* in adding order */
        std::vector<tb_stepper_data> data;
        std::vector<tb_stepper_transform> transforms;
        struct turtle_map * geoid;
        double local_range, slope_factor, resolution_factor;
        /* last sample (stepper.h:93-98) and local approximations (stepper.h:45-58) */
        tb::StepperState state;
        double lla[tb::LLA_ROWS_MAX]; /* tb::LlaView of stride 1, blocks of LLA_BLOCK_MAX */
        /* cached flattening */
        int dirty;
        tb_flat_geometry flat;
        enum turtle_return lookup_rc; /* first tile-load error of the running scalar call */
};

namespace tbh {

/* Error plumbing with the reference's message format (error.c:108-138). */
enum turtle_return raise(turtle_function_t * fn, enum turtle_return rc,
    const char * file, int line, const char * format, ...);

/* Make every tile of the stack resident on the host (turtle_stack_load). */
enum turtle_return stack_load_all(struct turtle_stack * stack,
    turtle_function_t * caller);

/* (Re)build stepper->flat (host nodes; stacks resolved on demand). */
enum turtle_return stepper_flatten(struct turtle_stepper * stepper,
    turtle_function_t * caller);
/* Flatten into `F`, for the host (load_tiles) or for a device residency plan. */
enum turtle_return flatten_into(struct turtle_stepper * stepper, tb_flat_geometry & F,
    turtle_function_t * caller, int load_tiles, const struct turtle_residency * region);

void projection_to_desc(const struct turtle_projection * p, tb::ProjDesc * d);

} /* namespace tbh */
