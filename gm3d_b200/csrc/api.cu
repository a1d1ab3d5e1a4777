// Host-only entry points of the C ABI (include/gm3d.h): version, error strings, workspace query.
#include <string.h>

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

// ---- inter-GPU inboxes of the per-step statistics all-reduce (set-up time; see gm3d_step_reduce_t) ----------
static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");

GM3D_API int gm3d_peer_alloc(size_t bytes, void** ptr, unsigned char handle[64]) {
    if (!ptr || !handle || bytes == 0) return GM3D_EINVAL;
    void* d = nullptr;
    cudaError_t e = cudaMalloc(&d, bytes);
    if (e != cudaSuccess) return static_cast<int>(e);
    e = cudaMemset(d, 0, bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, d);
    if (e != cudaSuccess) {
        cudaFree(d);
        return static_cast<int>(e);
    }
    memcpy(handle, &h, 64);
    *ptr = d;
    return GM3D_OK;
}

GM3D_API int gm3d_peer_open(const unsigned char handle[64], void** ptr) {
    if (!ptr || !handle) return GM3D_EINVAL;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    void* d = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&d, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return static_cast<int>(e);
    *ptr = d;
    return GM3D_OK;
}

GM3D_API int gm3d_peer_close(void* ptr) {
    if (!ptr) return GM3D_EINVAL;
    const cudaError_t e = cudaIpcCloseMemHandle(ptr);
    return e == cudaSuccess ? GM3D_OK : static_cast<int>(e);
}

GM3D_API int gm3d_peer_free(void* ptr) {
    if (!ptr) return GM3D_EINVAL;
    const cudaError_t e = cudaFree(ptr);
    return e == cudaSuccess ? GM3D_OK : static_cast<int>(e);
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
