// Host-only entry points of the C ABI (include/gm3d.h): version, error strings, workspace query.
#include "common.cuh"

namespace gm3d {
size_t fps_workspace_bytes(int B, int N);
size_t chamfer_workspace_bytes(int P);
size_t cloud_step_workspace_bytes(int P);
size_t learning_loss_workspace_bytes(int B);
}

GM3D_API int gm3d_abi_version(void) { return GM3D_ABI_VERSION; }

GM3D_API const char* gm3d_strerror(int code) {
    switch (code) {
        case GM3D_OK: return "success";
        case GM3D_EINVAL: return "invalid argument (shape, k > N, G > N or a required pointer is NULL)";
        case GM3D_ENOSUP: return "request not supported by this build (see include/gm3d.h)";
        case GM3D_EALIGN: return "pointer is not aligned as required";
        default: break;
    }
    if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
    return "unknown gm3d error";
}

GM3D_API size_t gm3d_workspace_bytes(int op, int B, int N, int G, int k) {
    (void)G;
    (void)k;
    switch (op) {
        case GM3D_OP_FPS:
        case GM3D_OP_GROUP: return gm3d::fps_workspace_bytes(B, N);
        case GM3D_OP_CHAMFER_FWD: return gm3d::chamfer_workspace_bytes(B);  // ticket + per-patch scratch
        case GM3D_OP_CLOUD_STEP: return gm3d::cloud_step_workspace_bytes(B);
        case GM3D_OP_LEARNING_LOSS: return gm3d::learning_loss_workspace_bytes(B);
        default: return 0;
    }
}
