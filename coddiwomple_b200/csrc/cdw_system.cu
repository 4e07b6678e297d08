// cdw_system.cu — cdw_system_create/destroy: copy one replica's term tables to the device once, build the
// constraint clusters and the atom -> force-scratch-slot CSR the kernels use.
// Replaces: OpenMMPDFState.__init__ creating the OpenMM Context (coddiwomple/openmm/states.py:50-86).
#include <stdarg.h>
#include <string.h>

#include <algorithm>
#include <numeric>
#include <vector>

#include "cdw_common.h"
#include "cdw_pack.h"

namespace cdw {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

}  // namespace cdw

using namespace cdw;

extern "C" int cdw_version(void) { return CDW_VERSION; }
extern "C" const char* cdw_last_error(void) { return cdw::g_err; }

extern "C" int cdw_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    CDW_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    CDW_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    return CDW_OK;
}

extern "C" int cdw_system_create(const cdw_system_desc* d, cdw_system** out) {
    CDW_REQUIRE(d && out, "null argument");
    std::vector<unsigned char> blob;
    SysDev dev;
    std::string err;
    int rc = pack_system(d, blob, dev, err);
    if (rc != CDW_OK) {
        set_error("cdw_system_create: %s", err.c_str());
        return rc;
    }
    cdw_system* s = new cdw_system();
    s->blob_bytes = blob.size();
    cudaError_t e = cudaMalloc(&s->blob, s->blob_bytes);
    if (e != cudaSuccess) {
        set_error("cudaMalloc(%zu) failed: %s", s->blob_bytes, cudaGetErrorString(e));
        delete s;
        return CDW_ECUDA;
    }
    e = cudaMemcpy(s->blob, blob.data(), s->blob_bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        set_error("cudaMemcpy failed: %s", cudaGetErrorString(e));
        cudaFree(s->blob);
        delete s;
        return CDW_ECUDA;
    }
    rebase_system(dev, (const unsigned char*)s->blob);
    s->dev = dev;
    *out = s;
    return CDW_OK;
}

extern "C" int cdw_system_destroy(cdw_system* sys) {
    if (!sys) return CDW_OK;
    cudaFree(sys->blob);
    cudaFree(sys->proto_tables);
    delete sys;
    return CDW_OK;
}

extern "C" int cdw_system_n_atoms(const cdw_system* sys) { return sys ? sys->dev.n_atoms : CDW_EINVAL; }
extern "C" int cdw_system_n_lambda(const cdw_system* sys) { return sys ? sys->dev.n_lambda : CDW_EINVAL; }
