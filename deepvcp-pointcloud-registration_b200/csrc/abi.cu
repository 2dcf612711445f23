// ABI bookkeeping of libdvcp_b200.so.
#include "common.cuh"

extern "C" int dvcp_abi_version(void) { return DVCP_ABI_VERSION; }

extern "C" const char *dvcp_error_string(int code) {
    if (code == 0) return "ok";
    if (code == DVCP_E_ARG) return "dvcp: invalid argument (null pointer or non-positive size)";
    if (code == DVCP_E_UNSUPPORTED) return "dvcp: size outside what the kernels are built for";
    if (code == DVCP_E_WORKSPACE) return "dvcp: workspace too small";
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "dvcp: unknown error";
}
