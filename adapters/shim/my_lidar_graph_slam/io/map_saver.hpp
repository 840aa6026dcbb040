// Empty stand-in (TEST INFRASTRUCTURE ONLY): loop_detector_branch_bound.cpp:8 and
// loop_detector_grid_search.cpp:8 include io/map_saver.hpp without using it; the real
// header needs Boost.GIL/libpng, which this image does not have.
#pragma once
