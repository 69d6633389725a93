// comm.cu -- the in-library communicator of row-sharded runs (include/oem_b200.h: oemb200_comm_*).
//
// The path has exactly one kind of exchange step: an in-place SUM all-reduce of FP64 sufficient statistics
//  -- the packed [Gram | X'y | column sums | counts] bundle once per fit, and the (p+1)-vector
// [sum(y - prob), X'(y - prob)] once per IRLS iteration of the logistic entry (the reference forms that sum in one
// process, src/oem_logistic_dense.h:970-1000).  Two transports, both ordered on the library's own CUDA stream:
//
//   * NCCL: libnccl.so.2 is dlopen()ed (the copy already mapped into the process -- torch's -- else the system one),
//     ncclAllReduce(double, sum).  Used for large buffers (the 8 MB bundle: bandwidth matters, NVLS applies).
//   * one-shot peer-memory kernel for vectors of <= OEMB200_P2P_MAX_DOUBLES on one NVLink / NVSwitch node: every
//     rank owns a mailbox [2 epochs][world][8192] doubles + flags, exported with cudaIpcGetMemHandle and mapped by
//     every peer.  One CTA stores the rank's vector into every peer's mailbox with 128-bit st.global over NVLink,
//     fences (system scope), raises its flag in every peer, waits for all flags of the epoch and sums the
//     mailboxes IN RANK ORDER -- every rank adds the same numbers in the same order, so the ranks stay
//     bit-identical (NCCL's ring / tree orders do not promise that) and the latency is one NVLink store round
//     (~2-3 us) instead of a NCCL launch (~20 us).  Two epochs of mailbox suffice: a rank can only be two
//     collectives ahead of a peer if that peer already finished reading the older one.
#include <dlfcn.h>
#include <unistd.h>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include "runtime.h"

namespace oemb200 {

// ---- the few NCCL entry points we need, resolved at run time (no link-time dependency) ----
namespace {
typedef struct { char internal[OEMB200_COMM_ID_BYTES]; } NcclUniqueId;      // ncclUniqueId, nccl.h
typedef int (*fn_get_version)(int *);
typedef int (*fn_get_unique_id)(NcclUniqueId *);
typedef int (*fn_comm_init_rank)(void **, int, NcclUniqueId, int);
typedef int (*fn_comm_destroy)(void *);
typedef int (*fn_all_reduce)(const void *, void *, size_t, int, int, void *, cudaStream_t);
typedef int (*fn_all_gather)(const void *, void *, size_t, int, void *, cudaStream_t);
typedef const char *(*fn_error_string)(int);
constexpr int kNcclChar = 0, kNcclFloat64 = 8, kNcclSum = 0;

struct NcclApi {
    void *handle = nullptr;
    int version = 0;
    std::string where;
    fn_get_version get_version = nullptr;
    fn_get_unique_id get_unique_id = nullptr;
    fn_comm_init_rank comm_init_rank = nullptr;
    fn_comm_destroy comm_destroy = nullptr;
    fn_all_reduce all_reduce = nullptr;
    fn_all_gather all_gather = nullptr;
    fn_error_string error_string = nullptr;
};

NcclApi &nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *env = getenv("OEMB200_NCCL_LIB");
        void *h = nullptr;
        if (env && *env) { h = dlopen(env, RTLD_NOW | RTLD_GLOBAL); api.where = env; }
        if (!h) { h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD); api.where = "libnccl.so.2 (already loaded in the process)"; }
        if (!h) { h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL); api.where = "libnccl.so.2"; }
        if (!h) { h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL); api.where = "libnccl.so"; }
        if (!h) return;
        api.handle = h;
        api.get_version = (fn_get_version)dlsym(h, "ncclGetVersion");
        api.get_unique_id = (fn_get_unique_id)dlsym(h, "ncclGetUniqueId");
        api.comm_init_rank = (fn_comm_init_rank)dlsym(h, "ncclCommInitRank");
        api.comm_destroy = (fn_comm_destroy)dlsym(h, "ncclCommDestroy");
        api.all_reduce = (fn_all_reduce)dlsym(h, "ncclAllReduce");
        api.all_gather = (fn_all_gather)dlsym(h, "ncclAllGather");
        api.error_string = (fn_error_string)dlsym(h, "ncclGetErrorString");
        if (api.get_version) api.get_version(&api.version);
    });
    return api;
}

NcclApi &need_nccl() {
    NcclApi &a = nccl_api();
    if (!a.handle || !a.get_unique_id || !a.comm_init_rank || !a.comm_destroy || !a.all_reduce || !a.all_gather)
        fail(OEMB200_ECOMM, "libnccl.so.2 could not be loaded (%s); set OEMB200_NCCL_LIB or pass an all-reduce callback",
             dlerror() ? dlerror() : "symbols missing");
    return a;
}

void nccl_check(int rc, const char *what) {
    if (rc == 0) return;
    NcclApi &a = nccl_api();
    fail(OEMB200_ECOMM, "%s failed: %s", what, a.error_string ? a.error_string(rc) : "NCCL error");
}
}  // namespace

constexpr int P2P_MAX_WORLD = 16;
constexpr int P2P_THREADS = 1024;

}  // namespace oemb200

// the opaque handle of the C ABI
struct oemb200_comm {
    void *nccl = nullptr;
    bool own_nccl = false;
    int rank = 0, world = 1, device = 0;
    // one-shot peer-memory transport
    bool p2p = false;
    void *mailbox_base = nullptr;                       // local allocation (cudaMalloc): data then flags
    void *peer_base[oemb200::P2P_MAX_WORLD] = {};       // every rank's allocation mapped here (own entry = mailbox_base)
    int64_t calls_nccl = 0, calls_p2p = 0;
};

namespace oemb200 {

namespace {
constexpr size_t kSlotDoubles = (size_t)OEMB200_P2P_MAX_DOUBLES;
inline size_t mailbox_data_bytes(int world) { return 2 * (size_t)world * kSlotDoubles * 8; }
// data | 2 x 16 flags | this rank's own epoch counter (never written by peers)
inline size_t mailbox_total_bytes(int world) { return mailbox_data_bytes(world) + 2 * (size_t)P2P_MAX_WORLD * 8 + 64; }

struct P2pArgs {
    double *peer_data[P2P_MAX_WORLD];
    unsigned long long *peer_flags[P2P_MAX_WORLD];
    int rank, world;
    unsigned long long *epoch_dev;     // this rank's count of EXECUTED collectives (skipped launches do not advance it)
    const int *skip;                   // optional: no-op when *skip != 0 (the flag is identical on every rank)
    unsigned long long *t_acc;         // optional: += this launch's duration in ns (device-side phase clock)
};

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// buf (count doubles, 16-byte aligned when count > 1) <- sum over ranks, in rank order.  One CTA.
__global__ void __launch_bounds__(P2P_THREADS, 1)
p2p_allreduce_kernel(const P2pArgs a, double *__restrict__ buf, int count) {
    if (a.skip && *a.skip) return;
    const unsigned long long t_in = (a.t_acc && threadIdx.x == 0) ? global_timer_ns() : 0ull;
    // the epoch lives on the device: ranks skip the same launches, so their counters agree, and two consecutive EXECUTED
    // collectives always use different mailbox slots even when predicated launches were enqueued between them
    const unsigned long long epoch = *a.epoch_dev + 1ull;
    const int slot = (int)(epoch & 1ull);
    const size_t my_off = ((size_t)slot * a.world + a.rank) * kSlotDoubles;
    const int pairs = count >> 1;
    // 1. push: my vector into slot [epoch & 1][my rank] of every rank's mailbox (NVLink stores; own mailbox included)
    for (int r = 0; r < a.world; ++r) {
        double *dst = a.peer_data[r] + my_off;
        for (int i = threadIdx.x; i < pairs; i += P2P_THREADS)
            reinterpret_cast<double2 *>(dst)[i] = reinterpret_cast<const double2 *>(buf)[i];
        if ((count & 1) && threadIdx.x == 0) dst[count - 1] = buf[count - 1];
    }
    __threadfence_system();
    __syncthreads();
    // 2. raise my flag in every rank
    if (threadIdx.x < a.world) st_release_sys(a.peer_flags[threadIdx.x] + slot * P2P_MAX_WORLD + a.rank, epoch);
    // 3. wait until every rank's vector of this epoch has landed here
    if (threadIdx.x < a.world) {
        const unsigned long long *f = a.peer_flags[a.rank] + slot * P2P_MAX_WORLD + threadIdx.x;
        const unsigned long long t0 = global_timer_ns();
        while (ld_acquire_sys(f) < epoch) {
            if (global_timer_ns() - t0 > 30ull * 1000000000ull) __trap();      // a peer died: fail loudly instead of hanging
        }
    }
    __syncthreads();
    // 4. sum the mailboxes in rank order (L1 bypassed: the lines are rewritten by peers every other epoch)
    const double *mine = a.peer_data[a.rank] + (size_t)slot * a.world * kSlotDoubles;
    for (int i = threadIdx.x; i < count; i += P2P_THREADS) {
        double s = __ldcg(mine + i);
        for (int r = 1; r < a.world; ++r) s += __ldcg(mine + (size_t)r * kSlotDoubles + i);
        buf[i] = s;
    }
    if (threadIdx.x == 0) {
        *a.epoch_dev = epoch;      // read again only by the next launch on this stream
        if (a.t_acc) *a.t_acc += global_timer_ns() - t_in;
    }
}

// Map every rank's mailbox.  Collective: the IPC handles travel through an ncclAllGather, the go / no-go decision
// through an ncclAllReduce, so all ranks end up with the same transport.
void p2p_setup(oemb200_comm *c) {
    if (c->world < 2 || c->world > P2P_MAX_WORLD) return;
    const char *env = getenv("OEMB200_COMM_P2P");
    const bool want = !(env && env[0] == '0');
    NcclApi &api = need_nccl();
    struct Card { cudaIpcMemHandle_t h; unsigned long long host; int ok; int pad; };
    static_assert(sizeof(Card) % 8 == 0, "Card must be 8-byte sized");
    Card mine;
    memset(&mine, 0, sizeof mine);
    char hn[256] = {0};
    gethostname(hn, sizeof hn - 1);
    unsigned long long hh = 1469598103934665603ull;
    for (const char *s = hn; *s; ++s) hh = (hh ^ (unsigned char)*s) * 1099511628211ull;
    mine.host = hh;
    bool local_ok = want;
    if (local_ok) {
        if (cudaMalloc(&c->mailbox_base, mailbox_total_bytes(c->world)) != cudaSuccess) { cudaGetLastError(); local_ok = false; }
    }
    if (local_ok) {
        cudaMemset(c->mailbox_base, 0, mailbox_total_bytes(c->world));
        if (cudaIpcGetMemHandle(&mine.h, c->mailbox_base) != cudaSuccess) { cudaGetLastError(); local_ok = false; }
    }
    mine.ok = local_ok ? 1 : 0;
    // all-gather the cards
    DBuf<Card> d_cards(c->world);
    OEM_CUDA(cudaMemcpy(d_cards.p + c->rank, &mine, sizeof mine, cudaMemcpyHostToDevice));
    nccl_check(api.all_gather(d_cards.p + c->rank, d_cards.p, sizeof(Card), kNcclChar, c->nccl, cudaStreamLegacy), "ncclAllGather");
    std::vector<Card> cards(c->world);
    OEM_CUDA(cudaMemcpy(cards.data(), d_cards.p, sizeof(Card) * c->world, cudaMemcpyDeviceToHost));
    bool ok = local_ok;
    for (int r = 0; r < c->world && ok; ++r) ok = cards[r].ok && cards[r].host == mine.host;
    if (ok) {
        for (int r = 0; r < c->world; ++r) {
            if (r == c->rank) { c->peer_base[r] = c->mailbox_base; continue; }
            if (cudaIpcOpenMemHandle(&c->peer_base[r], cards[r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                c->peer_base[r] = nullptr;
                ok = false;
                break;
            }
        }
    }
    // unanimous?
    DBuf<double> d_ok(1);
    const double okd = ok ? 1.0 : 0.0;
    OEM_CUDA(cudaMemcpy(d_ok.p, &okd, 8, cudaMemcpyHostToDevice));
    nccl_check(api.all_reduce(d_ok.p, d_ok.p, 1, kNcclFloat64, kNcclSum, c->nccl, cudaStreamLegacy), "ncclAllReduce");
    double tot = 0.0;
    OEM_CUDA(cudaMemcpy(&tot, d_ok.p, 8, cudaMemcpyDeviceToHost));
    c->p2p = (tot == (double)c->world);
    if (!c->p2p) {
        for (int r = 0; r < c->world; ++r)
            if (r != c->rank && c->peer_base[r]) { cudaIpcCloseMemHandle(c->peer_base[r]); c->peer_base[r] = nullptr; }
        if (c->mailbox_base) { cudaFree(c->mailbox_base); c->mailbox_base = nullptr; }
    }
}

void finish_create(oemb200_comm *c) {
    int prev = -1;
    OEM_CUDA(cudaGetDevice(&prev));
    if (c->device >= 0) OEM_CUDA(cudaSetDevice(c->device));
    OEM_CUDA(cudaGetDevice(&c->device));
    try {
        p2p_setup(c);
    } catch (...) {
        if (prev >= 0) cudaSetDevice(prev);
        throw;
    }
    if (prev >= 0 && prev != c->device) cudaSetDevice(prev);
}
}  // namespace

bool comm_can_skip(const oemb200_comm *c, int64_t count) {
    return !c || c->world <= 1 || (c->p2p && count <= OEMB200_P2P_MAX_DOUBLES);
}

void comm_all_reduce(oemb200_comm *c, double *dev_buf, int64_t count, cudaStream_t stream, const int *skip, unsigned long long *t_acc) {
    if (!c || c->world <= 1 || count <= 0) return;
    const bool aligned = count == 1 || (reinterpret_cast<uintptr_t>(dev_buf) & 15) == 0;
    if (skip && !(c->p2p && count <= OEMB200_P2P_MAX_DOUBLES && aligned))
        fail(OEMB200_ECOMM, "a predicated all-reduce needs the peer-memory transport and a 16-byte aligned buffer");
    if (c->p2p && count <= OEMB200_P2P_MAX_DOUBLES && aligned) {
        P2pArgs a;
        memset(&a, 0, sizeof a);
        for (int r = 0; r < c->world; ++r) {
            a.peer_data[r] = static_cast<double *>(c->peer_base[r]);
            a.peer_flags[r] = reinterpret_cast<unsigned long long *>(static_cast<char *>(c->peer_base[r]) + mailbox_data_bytes(c->world));
        }
        a.rank = c->rank; a.world = c->world; a.skip = skip; a.t_acc = t_acc;
        a.epoch_dev = reinterpret_cast<unsigned long long *>(static_cast<char *>(c->mailbox_base) + mailbox_data_bytes(c->world) +
                                                             2 * (size_t)P2P_MAX_WORLD * 8);
        p2p_allreduce_kernel<<<1, P2P_THREADS, 0, stream>>>(a, dev_buf, (int)count);
        OEM_CUDA(cudaGetLastError());
        c->calls_p2p += 1;
        return;
    }
    NcclApi &api = need_nccl();
    nccl_check(api.all_reduce(dev_buf, dev_buf, (size_t)count, kNcclFloat64, kNcclSum, c->nccl, stream), "ncclAllReduce");
    c->calls_nccl += 1;
}

bool comm_uses_p2p(const oemb200_comm *c, int64_t count) { return c && c->p2p && count <= OEMB200_P2P_MAX_DOUBLES; }

extern thread_local std::string g_last_error;
template <typename F>
static int comm_guarded(F &&f) {
    try {
        f();
        return OEMB200_OK;
    } catch (const Error &e) {
        g_last_error = e.what();
        cudaGetLastError();
        return e.code;
    } catch (const std::exception &e) {
        g_last_error = e.what();
        return OEMB200_EINVAL;
    }
}

}  // namespace oemb200

using namespace oemb200;

extern "C" {

int oemb200_comm_unique_id(void *id_out) {
    return comm_guarded([&] {
        if (!id_out) fail(OEMB200_EINVAL, "id_out is NULL");
        NcclApi &api = need_nccl();
        NcclUniqueId id;
        nccl_check(api.get_unique_id(&id), "ncclGetUniqueId");
        memcpy(id_out, &id, sizeof id);
    });
}

int oemb200_comm_create(const void *id, int rank, int world, int device, oemb200_comm **out) {
    return comm_guarded([&] {
        if (!id || !out) fail(OEMB200_EINVAL, "id / out must not be NULL");
        if (world < 1 || rank < 0 || rank >= world) fail(OEMB200_EINVAL, "bad rank %d of %d", rank, world);
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
            cudaGetLastError();
            fail(OEMB200_ENODEVICE, "no CUDA device available; liboem_b200 has no CPU fallback");
        }
        NcclApi &api = need_nccl();
        int prev = -1;
        OEM_CUDA(cudaGetDevice(&prev));
        if (device >= 0) OEM_CUDA(cudaSetDevice(device));
        oemb200_comm *c = new oemb200_comm();
        c->rank = rank; c->world = world; c->device = device;
        NcclUniqueId uid;
        memcpy(&uid, id, sizeof uid);
        const int rc = api.comm_init_rank(&c->nccl, world, uid, rank);
        if (prev >= 0) cudaSetDevice(prev);
        if (rc != 0) { delete c; nccl_check(rc, "ncclCommInitRank"); }
        c->own_nccl = true;
        try {
            finish_create(c);
        } catch (...) {
            api.comm_destroy(c->nccl);
            delete c;
            throw;
        }
        *out = c;
    });
}

int oemb200_comm_from_nccl(void *nccl_comm, int rank, int world, int device, oemb200_comm **out) {
    return comm_guarded([&] {
        if (!nccl_comm || !out) fail(OEMB200_EINVAL, "nccl_comm / out must not be NULL");
        if (world < 1 || rank < 0 || rank >= world) fail(OEMB200_EINVAL, "bad rank %d of %d", rank, world);
        need_nccl();
        oemb200_comm *c = new oemb200_comm();
        c->nccl = nccl_comm; c->own_nccl = false; c->rank = rank; c->world = world; c->device = device;
        try {
            finish_create(c);
        } catch (...) {
            delete c;
            throw;
        }
        *out = c;
    });
}

int oemb200_comm_destroy(oemb200_comm *c) {
    return comm_guarded([&] {
        if (!c) return;
        int prev = -1;
        cudaGetDevice(&prev);
        cudaSetDevice(c->device);
        cudaDeviceSynchronize();
        for (int r = 0; r < c->world && r < P2P_MAX_WORLD; ++r)
            if (r != c->rank && c->peer_base[r]) cudaIpcCloseMemHandle(c->peer_base[r]);
        if (c->mailbox_base) cudaFree(c->mailbox_base);
        if (c->own_nccl && c->nccl && nccl_api().comm_destroy) nccl_api().comm_destroy(c->nccl);
        if (prev >= 0) cudaSetDevice(prev);
        cudaGetLastError();
        delete c;
    });
}

int oemb200_comm_allreduce(oemb200_comm *c, double *dev_buf, int64_t count, void *stream, double *us_out) {
    return comm_guarded([&] {
        if (!c || !dev_buf || count < 0) fail(OEMB200_EINVAL, "comm / buffer missing");
        cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : cudaStreamLegacy;
        if (!us_out) { comm_all_reduce(c, dev_buf, count, s); return; }
        cudaEvent_t ea, eb;
        OEM_CUDA(cudaEventCreate(&ea)); OEM_CUDA(cudaEventCreate(&eb));
        OEM_CUDA(cudaEventRecord(ea, s));
        comm_all_reduce(c, dev_buf, count, s);
        OEM_CUDA(cudaEventRecord(eb, s));
        OEM_CUDA(cudaStreamSynchronize(s));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ea, eb);
        *us_out = ms * 1e3;
        cudaEventDestroy(ea); cudaEventDestroy(eb);
    });
}

int oemb200_comm_p2p_enabled(const oemb200_comm *c) { return (c && c->p2p) ? 1 : 0; }

}  // extern "C"
