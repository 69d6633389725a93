// runtime.cu -- device context and small runtime helpers (see runtime.h).
#include "runtime.h"

namespace oemb200 {

Ctx::Ctx(const oemb200_opts *o) {
    memset(&st, 0, sizeof st);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev < 1)
        fail(OEMB200_ENODEVICE, "no CUDA device available (%s); liboem_b200 has no CPU fallback",
             e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (o && o->device >= 0) {
        if (o->device >= ndev) fail(OEMB200_EINVAL, "device %d out of range (%d devices)", o->device, ndev);
        OEM_CUDA(cudaSetDevice(o->device));
    }
    OEM_CUDA(cudaGetDevice(&device));
    cudaDeviceProp prop;
    OEM_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        fail(OEMB200_ENODEVICE, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major,
             prop.minor);
    num_sms = prop.multiProcessorCount;
    smem_optin = prop.sharedMemPerBlockOptin;
    if (o && o->stream) {
        stream = static_cast<cudaStream_t>(o->stream);
        own_stream = false;
    } else {
        OEM_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        own_stream = true;
    }
    if (o) { allreduce = o->allreduce; allreduce_ctx = o->allreduce_ctx; }
}

Ctx::~Ctx() {
    if (own_stream && stream) cudaStreamDestroy(stream);
}

void Ctx::all_reduce(double *dev_buf, int64_t count) {
    if (!allreduce) return;
    const int rc = allreduce(dev_buf, count, stream, allreduce_ctx);
    if (rc != 0) fail(OEMB200_ECOMM, "all-reduce callback failed with code %d", rc);
}

bool is_device_ptr(const void *p) {
    cudaPointerAttributes at;
    cudaError_t e = cudaPointerGetAttributes(&at, p);
    if (e != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

}  // namespace oemb200
