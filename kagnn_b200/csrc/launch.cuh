// Kernel-launch seam of backward.cu.
//
// Product build (nvcc, the only thing libkagnn_b200.so is ever made from): KAGNN_LAUNCH is an ordinary <<< >>> launch.
//
// KAGNN_HOST_CHECK (defined only by tests/emul/build_emul.py, g++, output under tests/emul/_build/): the kernels of backward.cu
// use no shared memory, no barriers and no warp intrinsics, so running their threads one after another on the host is a valid
// schedule.  The CPU test-suite uses that to check the index arithmetic of those kernels in a container without a GPU.  It is
// test infrastructure: nothing under kagnn_b200/ loads it, and the product has no CPU path.
#pragma once
#ifdef KAGNN_HOST_CHECK
#include "host_check.h"            // tests/emul/host_check.h
#else
#include "common.cuh"
#define KAGNN_LAUNCH(kernel, grid, block, stream, ...) kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__)
// product build: try the shared-memory-tiled kernel (backward_tiled.cu) first; it declines shapes outside its limits
#define KAGNN_TRY_TILED(call)                              \
    do {                                                   \
        const int _trc = (call);                           \
        if (_trc != KAGNN_EUNSUPPORTED) return _trc;       \
    } while (0)
#endif
#ifndef KAGNN_TRY_TILED
#define KAGNN_TRY_TILED(call) do { } while (0)            // the serial host build checks the general kernels only
#endif
