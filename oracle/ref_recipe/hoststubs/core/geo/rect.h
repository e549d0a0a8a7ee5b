// TEST INFRASTRUCTURE ONLY. core/geo/rect.h is included by voxel_collection.h but none of its names
// (2-d rectangles for the quadtree) are used by the 3-d voxelisation compiled here.
#pragma once
