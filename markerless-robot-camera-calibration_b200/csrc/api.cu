// api.cu — version / error strings of libb2me.
#include "common.cuh"

extern "C" int b2me_version(void) { return 100; }  // 0.1.0

extern "C" const char* b2me_strerror(int code) {
    switch (code) {
        case B2ME_OK: return "ok";
        case B2ME_EINVAL: return "invalid argument";
        case B2ME_EWORKSPACE: return "workspace or hash table too small";
        case B2ME_ELAUNCH: return "CUDA launch failed";
        case B2ME_EUNSUPPORTED: return "shape/dtype combination not supported by this kernel";
        default: return "unknown error";
    }
}
