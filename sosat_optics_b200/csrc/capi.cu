// Library-level entry points of libsosat_b200.so: version, error text, device selection.
#include <cstring>

#include "common.cuh"

namespace sosat {

char *error_buffer() {
    static thread_local char buf[512] = {0};
    return buf;
}

int set_error(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(error_buffer(), 512, fmt, ap);
    va_end(ap);
    return code;
}

}  // namespace sosat

using namespace sosat;

extern "C" int sosat_abi_version(void) { return SOSAT_ABI_VERSION; }

extern "C" int sosat_build_flags(void) {
#ifdef SOSAT_EXPERIMENTS
    return SOSAT_BUILD_EXPERIMENTS;
#else
    return 0;
#endif
}

extern "C" const char *sosat_last_error(void) { return error_buffer(); }

extern "C" int sosat_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" int sosat_set_device(int device) {
    if (sosat_device_count() <= 0)
        return set_error(SOSAT_ERR_NO_DEVICE, "no CUDA device visible; libsosat_b200 has no CPU path");
    SOSAT_CUDA_OK(cudaSetDevice(device));
    return SOSAT_OK;
}

extern "C" int sosat_device_info(char *name, int *sm, int *n_sm, size_t *mem_bytes) {
    if (sosat_device_count() <= 0)
        return set_error(SOSAT_ERR_NO_DEVICE, "no CUDA device visible; libsosat_b200 has no CPU path");
    int dev = 0;
    SOSAT_CUDA_OK(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    SOSAT_CUDA_OK(cudaGetDeviceProperties(&prop, dev));
    if (name) {
        strncpy(name, prop.name, 255);
        name[255] = 0;
    }
    if (sm) *sm = prop.major * 10 + prop.minor;
    if (n_sm) *n_sm = prop.multiProcessorCount;
    if (mem_bytes) *mem_bytes = prop.totalGlobalMem;
    return SOSAT_OK;
}
