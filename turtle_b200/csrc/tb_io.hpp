/*
 * tb_io.hpp -- the on-disk DEM formats of the reference (src/turtle/io.c:54-71:
 * tif, grd, hgt, png, asc), read WITHOUT libpng / libtiff: PNG chunks + zlib inflate,
 * a baseline TIFF directory walk, and the two text grids.
 *
 * Every reader produces the same meta data and the same 16-bit node values as the
 * reference's reader of that format; what differs is the in-memory layout: nodes are
 * handed over in FILE order together with a `RawLayout` that says how the file stores
 * them (byte order, first row), so that the caller can normalise them either on the
 * host or -- for residency plans -- on the device (ingest_kernel, tb_kernels.cu).
 */
#pragma once

#include <stdint.h>
#include <string>
#include <vector>

#include "turtle.h"

namespace tbio {

/* What a reader learns from the file header (struct turtle_map_meta, map.h:41-58). */
struct Header {
        int nx = 0, ny = 0;
        double x0 = 0., y0 = 0., z0 = 0., dx = 0., dy = 0., dz = 0.;
        int kind = 0;            /* tb::NodeKind */
        std::string projection;  /* "" = geodetic */
        std::string encoding;    /* the file extension (io.c:95) */
};

/* How the 16-bit nodes are laid out in the `raw` buffer a reader returns. */
struct RawLayout {
        int big_endian = 0;  /* samples are big-endian (hgt, png) */
        int north_first = 0; /* first row of the buffer is the northernmost (hgt, png, tif) */
};

struct Error {
        enum turtle_return code = TURTLE_RETURN_SUCCESS;
        std::string message;
        const char * file = "src/turtle/io.c"; /* reference source the message belongs to */
};

/* Extension of `path` (after the last '.' of the file name), or NULL. */
const char * extension(const char * path);
/* 1 if the extension names one of the five formats. */
int known_extension(const char * ext);

/* Header only (what turtle_stack_create needs of each tile, stack.c:73-91). */
int read_header(const char * path, Header & header, Error & error);
/* Header + nodes in file order. Returns 0 on success. */
int read_map(const char * path, Header & header, RawLayout & layout,
    std::vector<uint16_t> & raw, Error & error);
/* In place: file order -> rows south first, native endian (the host layout of
 * struct turtle_map). */
void normalise(const Header & header, const RawLayout & layout, std::vector<uint16_t> & raw);

/* turtle_map_dump for `.png` (png16.c:456-546) and `.tif` (geotiff16.c:262-330).
 * `nodes`: south first, native endian, as decoded by `kind`. */
int write_map(const char * path, const Header & header, const std::vector<uint16_t> & nodes,
    Error & error);

} /* namespace tbio */
