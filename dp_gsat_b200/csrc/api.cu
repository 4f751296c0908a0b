// Library-level entry points: version, error strings, device check.
#include "common.cuh"

extern "C" int gsatb_version(void) { return GSATB_VERSION; }

extern "C" const char* gsatb_strerror(int code) {
    switch (code) {
        case GSATB_OK: return "ok";
        case GSATB_EINVAL: return "invalid argument (null pointer, negative size or unsupported flag)";
        case GSATB_ESHAPE: return "unsupported shape (feature width must be a multiple of 4 and within limits)";
        case GSATB_EALIGN: return "pointer not 16-byte aligned";
        case GSATB_EWS_TOO_SMALL: return "workspace too small";
        case GSATB_EARCH: return "device is not sm_100-class (this library has no fallback path)";
        case GSATB_ENOT_SYMMETRIC: return "edge set is not symmetric";
        case GSATB_ELAUNCH: return "CUDA launch failed";
        default: return "unknown error";
    }
}

extern "C" int gsatb_check_device(void) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return GSATB_EARCH;
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return GSATB_EARCH;
    return major == 10 ? GSATB_OK : GSATB_EARCH;
}

const unsigned long long*& gsatb_step_counter_ref() {
    static const unsigned long long* ptr = nullptr;
    return ptr;
}

extern "C" int gsatb_set_step_counter(const uint64_t* dev_counter) {
    gsatb_step_counter_ref() = reinterpret_cast<const unsigned long long*>(dev_counter);
    return GSATB_OK;
}
