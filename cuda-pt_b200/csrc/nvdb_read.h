// nvdb_read.h — reads one grid of a NanoVDB file (.nvdb, uncompressed, format major version 32) into the brick grid of
// include/cpts_format.h.  Own reader of the published container layout; no NanoVDB code is linked or included.
//
// Reference behaviour it stands in for: GridVolumeManager::from_vdb_file -> nanovdb::io::readGrid
// (src/volume/grid.cuh:155-170) followed by `getAccessor().getValue(Coord)` lookups (grid.cuh:51-92).  A NanoVDB leaf
// is an 8x8x8 block in x-major voxel order, which is exactly one brick.
#pragma once
#include "cpts_format.h"

#include <string>
#include <vector>

namespace cptb {

struct NvdbGrid {
    CptsGridChannel channel{}; // brick_offset / data_offset = 0: relative to `bricks` / `data` below
    std::vector<int32_t> bricks;
    std::vector<float> data;
    float value_min = 0, value_max = 0; // root extrema of the active values (first component for vector grids)
    double world_bbox[6] = {0, 0, 0, 0, 0, 0};
};

// ncomp: 1 = float grid, 3 = Vec3f grid; throws std::runtime_error on anything else (compressed files, other value
// types, other format versions)
void read_nvdb(const std::string &path, int ncomp, NvdbGrid &out);

} // namespace cptb
