// cabi.cu — library-level entry points and host-side argument plumbing of the C-ABI
// (include/sige_b200.h).  Replaces the dispatch half of reference sige/common.cpp.
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "common.cuh"

namespace sige {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: kernel launch failed: %s", what, cudaGetErrorString(e));
        return 2;
    }
    return 0;
}

// reference sige/common.cpp:25-34 (`broadcastable`): every dim is 1 or the full extent.
int make_bcast(const sige_bcast_t *in, int B, int C, int H, int W, const char *name, Bcast *out) {
    out->ptr = nullptr;
    out->sb = out->sc = out->sh = out->sw = 0;
    out->dtype = SIGE_F32;
    out->c_contig = 0;
    if (in == nullptr || in->ptr == nullptr) return 0;
    const int full[4] = {B, C, H, W};
    for (int d = 0; d < 4; ++d) {
        if (in->dims[d] != 1 && in->dims[d] != full[d]) {
            set_error("operand '%s' is not broadcastable: dim %d is %d, expected 1 or %d", name, d, in->dims[d],
                      full[d]);
            return 1;
        }
    }
    if (in->dtype < SIGE_F32 || in->dtype > SIGE_BF16) {
        set_error("operand '%s': unsupported dtype %d", name, in->dtype);
        return 1;
    }
    out->ptr = in->ptr;
    out->sb = in->dims[0] > 1 ? in->stride[0] : 0;
    out->sc = in->dims[1] > 1 ? in->stride[1] : 0;
    out->sh = in->dims[2] > 1 ? in->stride[2] : 0;
    out->sw = in->dims[3] > 1 ? in->stride[3] : 0;
    out->dtype = in->dtype;
    out->c_contig = (in->dims[1] > 1 && in->stride[1] == 1) ? 1 : 0;
    return 0;
}

}  // namespace sige

extern "C" {

const char *sige_last_error(void) { return sige::g_err; }

int sige_abi_version(void) { return SIGE_B200_ABI_VERSION; }

const char *sige_built_arch(void) {
#ifdef SIGE_BUILT_ARCH
    return SIGE_BUILT_ARCH;
#else
    return "unknown";
#endif
}

// reference sige/common.cpp:17-23
int sige_activation_from_name(const char *name) {
    if (name == nullptr) return -1;
    if (strcmp(name, "identity") == 0) return SIGE_ACT_IDENTITY;
    if (strcmp(name, "swish") == 0) return SIGE_ACT_SWISH;
    return -1;
}

}  // extern "C"
