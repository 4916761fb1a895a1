// ABI bookkeeping: version and error strings.
#include <stdio.h>

#include "rg_common.cuh"

extern "C" {

int rg_abi_version(void) { return RG_ABI_VERSION; }

const char *rg_strerror(int status) {
    switch (status) {
        case RG_OK: return "ok";
        case RG_ERR_BAD_ARG: return "bad argument (null pointer, negative size or inconsistent shapes)";
        case RG_ERR_UNSUPPORTED: return "unsupported hidden_dim (16/32/48/64) or attn_dim (<= 8)";
        case RG_ERR_WORKSPACE: return "workspace smaller than rg_workspace_bytes()";
        case RG_ERR_TOO_LARGE: return "an index space of this call exceeds 2^31-1: split the query batch";
        case RG_ERR_IO: return "file cannot be opened or mapped";
        case RG_ERR_PARSE: return "a line of a triples file does not hold exactly three names";
        case RG_ERR_UNKNOWN_NAME: return "a name of a triples file is missing from entity2id / relation2id";
        case RG_ERR_HOST: return "host allocation failed";
        default: break;
    }
    if (status <= RG_ERR_CUDA_BASE) return cudaGetErrorString((cudaError_t)(RG_ERR_CUDA_BASE - status));
    return "unknown rg_status";
}

}  // extern "C"
