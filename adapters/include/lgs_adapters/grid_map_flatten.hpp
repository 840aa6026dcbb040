/* grid_map_flatten.hpp -- reference GridMap<T> -> dense row-major doubles for lgs_grid_upload.
 *
 * Unallocated patches and unknown cells both become 0.0, which is what every reader on the
 * hot path gets from GridMap::Value(x, y, unknown) (grid_map/grid_map.hpp:859-873). */
#ifndef LGS_ADAPTERS_GRID_MAP_FLATTEN_HPP
#define LGS_ADAPTERS_GRID_MAP_FLATTEN_HPP

#include <algorithm>
#include <vector>

#include "my_lidar_graph_slam/grid_map/grid_map.hpp"

namespace LgsB200 {

template <typename MapT>
void FlattenGridMap(const MapT& map, std::vector<double>& dense)
{
    const int nx = map.NumOfGridCellsX(), ny = map.NumOfGridCellsY();
    const int patch = map.PatchSize();
    dense.assign(static_cast<std::size_t>(nx) * ny, 0.0);
    for (int py = 0; py < map.NumOfPatchesY(); ++py)
        for (int px = 0; px < map.NumOfPatchesX(); ++px) {
            if (!map.PatchIsAllocated(px, py))
                continue;
            const auto* cells = map.PatchAt(px, py).Data();   /* row-major y * size + x */
            for (int y = 0; y < patch; ++y) {
                double* row = dense.data() +
                    static_cast<std::size_t>(py * patch + y) * nx + px * patch;
                for (int x = 0; x < patch; ++x)
                    row[x] = cells[y * patch + x].Value();
            }
        }
}

} /* namespace LgsB200 */

#endif
