// capi_peer.cu -- CUDA IPC plumbing for the partitioned long signal (include/savgol_b200.h part 2).
//
// One process per GPU: a rank maps its ring neighbours' slices into its own address space once
// (cudaIpc*), after which savgol_apply_halo() can be handed halo pointers that live in the
// neighbours' HBM.  The 1D kernel's edge path then fetches the 2n halo samples with ordinary
// loads over NVLink / NVSwitch while it stages the first and last segment of the slice -- the halo
// "exchange" of SURVEY.md 8(e) costs no collective, no extra launch and no staging buffer.
#include <cuda.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstring>

#include "engine.h"


using sge::cuda_ok;

namespace {
// base address of the allocation that contains p (driver entry point resolved at run time: the
// library must load on machines without libcuda, e.g. for the CPU-side ABI tests)
bool allocation_base(const void* p, char** base)
{
    typedef CUresult (*GetRange)(CUdeviceptr*, size_t*, CUdeviceptr);
    static std::atomic<GetRange> s_fn{nullptr};
    GetRange fn = s_fn.load(std::memory_order_acquire);
    if (!fn) {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (!cuda_ok(cudaGetDriverEntryPoint("cuMemGetAddressRange", &f, cudaEnableDefault, &q), "cudaGetDriverEntryPoint") || !f ||
            q != cudaDriverEntryPointSuccess)
            return false;
        fn = reinterpret_cast<GetRange>(f);
        s_fn.store(fn, std::memory_order_release);
    }
    CUdeviceptr b = 0;
    size_t size = 0;
    if (fn(&b, &size, reinterpret_cast<CUdeviceptr>(p)) != CUDA_SUCCESS) return false;
    *base = reinterpret_cast<char*>(b);
    return true;
}
}  // namespace

extern "C" {

int savgol_b200_ipc_export(const void* dev_ptr, void* handle64, size_t* offset)
{
    static_assert(sizeof(cudaIpcMemHandle_t) == SAVGOL_B200_IPC_HANDLE_BYTES, "handle size");
    if (!dev_ptr || !handle64 || !offset) return -1;
    char* base = nullptr;
    if (!allocation_base(dev_ptr, &base)) {
        fprintf(stderr, "savgol_b200_ipc_export: not a device allocation\n");
        return -1;
    }
    cudaIpcMemHandle_t h;
    if (!cuda_ok(cudaIpcGetMemHandle(&h, base), "cudaIpcGetMemHandle")) return -1;
    std::memcpy(handle64, &h, sizeof h);
    *offset = static_cast<size_t>(static_cast<const char*>(dev_ptr) - base);
    return 0;
}

void* savgol_b200_ipc_open(const void* handle64, size_t offset)
{
    if (!handle64) return nullptr;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle64, sizeof h);
    void* p = nullptr;
    if (!cuda_ok(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle")) return nullptr;
    return static_cast<char*>(p) + offset;
}

int savgol_b200_ipc_close(void* mapped, size_t offset)
{
    if (!mapped) return 0;
    return cuda_ok(cudaIpcCloseMemHandle(static_cast<char*>(mapped) - offset), "cudaIpcCloseMemHandle") ? 0 : -1;
}

}  // extern "C"
