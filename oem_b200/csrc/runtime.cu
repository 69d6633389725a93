// runtime.cu -- device context and small runtime helpers (see runtime.h).
#include <cstdlib>
#include <map>
#include <vector>
#include <exception>
#include "runtime.h"

namespace oemb200 {

Ctx::Ctx(const oemb200_opts *o) {
    memset(&st, 0, sizeof st);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev < 1)
        fail(OEMB200_ENODEVICE, "no CUDA device available (%s); liboem_b200 has no CPU fallback",
             e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    OEM_CUDA(cudaGetDevice(&prev_device));
    if (o && o->device >= 0) {
        if (o->device >= ndev) fail(OEMB200_EINVAL, "device %d out of range (%d devices)", o->device, ndev);
        OEM_CUDA(cudaSetDevice(o->device));
    }
    OEM_CUDA(cudaGetDevice(&device));
    int major = 0, minor = 0, sms = 0, optin = 0;      // cudaGetDeviceProperties is milliseconds; these are microseconds
    OEM_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    OEM_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
    OEM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    OEM_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    if (major < 10)
        fail(OEMB200_ENODEVICE, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, major, minor);
    num_sms = sms;
    smem_optin = (size_t)optin;
    // NULL means the legacy default stream (like any CUDA library call without a stream argument): work is then
    // ordered after whatever the host program queued there (e.g. torch kernels that produced a device-resident X).
    stream = (o && o->stream) ? static_cast<cudaStream_t>(o->stream) : cudaStreamLegacy;
    own_stream = false;
    if (o) {
        allreduce = o->allreduce; allreduce_ctx = o->allreduce_ctx; comm = o->comm;
        rank = o->rank; world = o->world > 0 ? o->world : 1;
    }
    tm = new PhaseTimers(stream);
}

Ctx::~Ctx() {
    if (std::uncaught_exceptions() > 0) cudaDeviceSynchronize();   // unwinding: let queued kernels drain before buffers recycle
    delete tm;
    if (own_stream && stream) cudaStreamDestroy(stream);
    if (prev_device >= 0 && prev_device != device) cudaSetDevice(prev_device);   // leave the caller's device as we found it
}

void Ctx::finish() {
    OEM_CUDA(cudaStreamSynchronize(stream));
    if (tm) {
        tm->collect();
        delete tm;
        tm = new PhaseTimers(stream);
    }
}

// ---- per-thread, per-device free list of timing events ----
namespace {
thread_local std::map<int, std::vector<cudaEvent_t>> g_events;
}
cudaEvent_t event_acquire() {
    int dev = 0;
    cudaGetDevice(&dev);
    auto &fl = g_events[dev];
    if (!fl.empty()) {
        cudaEvent_t e = fl.back();
        fl.pop_back();
        return e;
    }
    cudaEvent_t e;
    OEM_CUDA(cudaEventCreate(&e));
    return e;
}
void event_release(cudaEvent_t e) {
    int dev = 0;
    cudaGetDevice(&dev);
    auto &fl = g_events[dev];
    if (fl.size() < 4096) fl.push_back(e);
    else cudaEventDestroy(e);
}

// ---- thread-local caching allocator ----
namespace {
struct Pool {
    std::multimap<size_t, void *> free_blocks;
    std::map<void *, size_t> live;
};
thread_local std::map<int, Pool> g_pools;      // one cache per device
thread_local std::map<void *, int> g_owner;    // live block -> device
Pool &cur_pool(int *dev_out = nullptr) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev_out) *dev_out = dev;
    return g_pools[dev];
}
}  // namespace

void *pool_alloc(size_t bytes) {
    int dev = 0;
    Pool &g_pool = cur_pool(&dev);
    const size_t sz = (bytes + 511) & ~size_t(511);
    auto it = g_pool.free_blocks.lower_bound(sz);
    if (it != g_pool.free_blocks.end() && it->first <= sz + sz / 8) {   // reuse a block at most 12.5% larger
        void *p = it->second;
        g_pool.live[p] = it->first;
        g_pool.free_blocks.erase(it);
        g_owner[p] = dev;
        return p;
    }
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, sz);
    if (e != cudaSuccess) {
        cudaGetLastError();
        pool_release_all();                       // drop the cache and retry once
        e = cudaMalloc(&p, sz);
        if (e != cudaSuccess) fail(OEMB200_ECUDA, "cudaMalloc of %.3f GB failed: %s", sz / 1e9, cudaGetErrorString(e));
    }
    g_pool.live[p] = sz;
    g_owner[p] = dev;
    return p;
}

void pool_free(void *p) {
    auto ow = g_owner.find(p);
    if (ow == g_owner.end()) { cudaFree(p); return; }
    Pool &g_pool = g_pools[ow->second];
    g_owner.erase(ow);
    auto it = g_pool.live.find(p);
    if (it == g_pool.live.end()) { cudaFree(p); return; }
    // Blocks up to OEMB200_POOL_MAX_GB (default 64) stay cached: repeated fits re-use the row-slab copy of the logistic
    // entry (16 GB at configs[3]) and the fold-sorted copy of xval.oem (40 GB at configs[2]) instead of paying a
    // cudaMalloc + cudaFree pair (each a device-wide synchronisation) per call.  pool_alloc drops the whole cache and
    // retries when the device runs out of memory; oemb200_release_cache() drops it on request.
    static const size_t keep_max = [] {
        const char *e = getenv("OEMB200_POOL_MAX_GB");
        const double gb = e ? atof(e) : 64.0;
        return (size_t)(gb * 1073741824.0);
    }();
    if (it->second > keep_max) cudaFree(p);
    else g_pool.free_blocks.emplace(it->second, p);
    g_pool.live.erase(it);
}

void pool_release_all() {
    for (auto &dp : g_pools) {
        for (auto &kv : dp.second.free_blocks) cudaFree(kv.second);
        dp.second.free_blocks.clear();
    }
}

bool Ctx::all_reduce_can_skip(int64_t count) const {
    if (comm) return comm_can_skip(comm, count);
    return allreduce == nullptr;
}

void Ctx::all_reduce(double *dev_buf, int64_t count, const int *skip, unsigned long long *t_acc) {
    if (comm) {
        comm_all_reduce(comm, dev_buf, count, stream, skip, t_acc);
    } else if (allreduce) {
        if (skip) fail(OEMB200_ECOMM, "a predicated all-reduce needs the in-library peer-memory transport");
        const int rc = allreduce(dev_buf, count, stream, allreduce_ctx);
        if (rc != 0) fail(OEMB200_ECOMM, "all-reduce callback failed with code %d", rc);
    } else {
        return;
    }
    st.allreduce_calls += 1;
    st.allreduce_doubles += count;
}

bool is_device_ptr(const void *p) {
    cudaPointerAttributes at;
    cudaError_t e = cudaPointerGetAttributes(&at, p);
    if (e != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

}  // namespace oemb200
