// common.cuh -- error plumbing and small device helpers shared by the libb200kv translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>

#include "../../include/b200kv.h"

namespace b200kv {

void set_error(const std::string& msg);   // thread-local last error (api.cu)

#define B2_CHECK_CUDA(expr)                                                                      \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            ::b200kv::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));             \
            return -1;                                                                           \
        }                                                                                        \
    } while (0)

#define B2_REQUIRE(cond, msg)                                                                    \
    do {                                                                                         \
        if (!(cond)) {                                                                           \
            ::b200kv::set_error(std::string("invalid argument: ") + (msg));                      \
            return -2;                                                                           \
        }                                                                                        \
    } while (0)

// Kernel-parameter copy of the per-plane base pointers and quantiser constants
// (plane nl = kv * L + l).  Lives in the constant bank; indexed dynamically.
struct PlaneTable {
    const uint16_t* p[B200KV_MAX_PLANES];
    float maxq[B200KV_MAX_PLANES];       // bins // 2 - 1
};

// Fill a PlaneTable from a kv_desc + bins; returns 0 or <0 with error set.
int make_plane_table(const b200kv_kv_desc* kv, const float* key_bins, const float* value_bins, PlaneTable* out);

}  // namespace b200kv
