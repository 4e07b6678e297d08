// cdw_common.h — internal declarations shared by the translation units of libcdw_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/cdw.h"
#include "cdw_math.cuh"

namespace cdw {

void set_error(const char* fmt, ...);

#define CDW_CUDA(call)                                                                         \
    do {                                                                                       \
        cudaError_t _e = (call);                                                               \
        if (_e != cudaSuccess) {                                                               \
            cdw::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return CDW_ECUDA;                                                                  \
        }                                                                                      \
    } while (0)

#define CDW_REQUIRE(cond, msg)                                            \
    do {                                                                  \
        if (!(cond)) {                                                    \
            cdw::set_error("%s: %s (%s:%d)", __func__, msg, __FILE__, __LINE__); \
            return CDW_EINVAL;                                            \
        }                                                                 \
    } while (0)

}  // namespace cdw

struct cdw_system {
    cdw::SysDev dev;
    void* blob;          // single device allocation holding every table
    size_t blob_bytes;
    // optional registered lambda protocol (cdw_system_register_protocol): prebuilt effective-term tables per row
    const double* proto_lambda = nullptr;
    int proto_rows = 0;
    double proto_kT = 0.0;
    unsigned char* proto_tables = nullptr;
    size_t proto_stride = 0;
    int team = 0;        // lanes per replica pinned by cdw_system_set_team (0: chosen from the system)
};
