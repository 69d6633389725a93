// ingest.cu -- host -> device ingest of column-major row blocks from PAGEABLE memory (an R matrix, a numpy array, the
// mmap'd .bk file of a bigmemory file-backed matrix: src/oem_big.cpp:52-64, R/big_oem.R:87-90).
//
// cudaMemcpy2DAsync from pageable memory goes through the driver's single staging thread and measured ~11 GB/s on
// the B200 boxes (profiles/r01_box_probe.json) against ~55 GB/s for pinned memory.  HostStager keeps a ring of
// pinned slots and a small pool of reader threads: the threads copy column segments of the source into a slot in
// parallel (that is also what faults the pages of a memory-mapped file in), the slot goes to the device with one
// asynchronous pinned copy, and the next slot is being filled while that DMA runs.  Pinned / registered sources skip
// the ring (is_pinned_host) and are copied in place.
#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>
#include "runtime.h"

namespace oemb200 {

bool is_pinned_host(const void *p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

class HostStager {
public:
    HostStager(int threads, size_t slot_bytes, int nslots) : slot_bytes_(slot_bytes) {
        slots_.resize(nslots);
        for (auto &s : slots_) {
            OEM_CUDA(cudaHostAlloc(&s.p, slot_bytes, cudaHostAllocDefault));
            OEM_CUDA(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
        }
        stop_ = false;
        for (int t = 0; t < threads; ++t) workers_.emplace_back([this] { work(); });
    }
    ~HostStager() {
        {
            std::lock_guard<std::mutex> g(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto &w : workers_) w.join();
        for (auto &s : slots_) {
            if (s.used) cudaEventSynchronize(s.done);
            cudaEventDestroy(s.done);
            cudaFreeHost(s.p);
        }
    }
    int threads() const { return (int)workers_.size(); }

    // rows [0, nr) x columns [0, p) of the column-major host block `src` (leading dimension ldx) -> device `dst`
    // (leading dimension ld_dst), ordered on `stream`.  Returns when the last slot has been queued (its DMA may still run).
    void copy_block(const double *src, int64_t ldx, int64_t nr, int p, double *dst, int64_t ld_dst, cudaStream_t stream) {
        const size_t col_bytes = (size_t)nr * 8;
        if (col_bytes > slot_bytes_) {
            // a single column segment larger than a slot: split the rows
            const int64_t rmax = (int64_t)(slot_bytes_ / 8);
            for (int64_t r = 0; r < nr; r += rmax)
                copy_block(src + r, ldx, std::min(rmax, nr - r), p, dst + r, ld_dst, stream);
            return;
        }
        const int kmax = (int)std::max<size_t>(1, slot_bytes_ / col_bytes);
        for (int j0 = 0; j0 < p; j0 += kmax) {
            const int k = std::min(kmax, p - j0);
            Slot &s = slots_[next_];
            next_ = (next_ + 1) % slots_.size();
            if (s.used) OEM_CUDA(cudaEventSynchronize(s.done));          // the slot's previous DMA has drained
            run_job(src + (size_t)j0 * ldx, ldx, nr, k, static_cast<double *>(s.p));
            OEM_CUDA(cudaMemcpy2DAsync(dst + (size_t)j0 * ld_dst, (size_t)ld_dst * 8, s.p, col_bytes, col_bytes, k,
                                       cudaMemcpyHostToDevice, stream));
            OEM_CUDA(cudaEventRecord(s.done, stream));
            s.used = true;
        }
    }

private:
    struct Slot { void *p = nullptr; cudaEvent_t done = nullptr; bool used = false; };
    struct Job { const double *src; int64_t ldx, nr; int k; double *dst; };

    void run_job(const double *src, int64_t ldx, int64_t nr, int k, double *dst) {
        {
            std::lock_guard<std::mutex> g(m_);
            job_ = Job{src, ldx, nr, k, dst};
            next_col_ = 0;
            pending_ = k;
            ++generation_;
        }
        cv_.notify_all();
        std::unique_lock<std::mutex> g(m_);
        done_cv_.wait(g, [this] { return pending_ == 0; });
    }
    void work() {
        uint64_t seen = 0;
        for (;;) {
            Job j;
            {
                std::unique_lock<std::mutex> g(m_);
                cv_.wait(g, [&] { return stop_ || generation_ != seen; });
                if (stop_) return;
                seen = generation_;
                j = job_;
            }
            for (;;) {
                int c;
                {
                    // column tickets are handed out under the lock and tied to the job generation: a worker that is
                    // late leaving job g can never take (or account for) a column of job g + 1
                    std::lock_guard<std::mutex> g(m_);
                    if (generation_ != seen || next_col_ >= j.k) break;
                    c = next_col_++;
                }
                memcpy(j.dst + (size_t)c * j.nr, j.src + (size_t)c * j.ldx, (size_t)j.nr * 8);
                std::lock_guard<std::mutex> g(m_);
                if (--pending_ == 0) done_cv_.notify_all();
            }
        }
    }

    size_t slot_bytes_;
    std::vector<Slot> slots_;
    size_t next_ = 0;
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_, done_cv_;
    Job job_{};
    int next_col_ = 0;
    int pending_ = 0;
    uint64_t generation_ = 0;
    bool stop_ = false;
};

namespace {
thread_local std::unique_ptr<HostStager> g_stager;
thread_local int g_stager_device = -1;
}

// reader threads: OEMB200_INGEST_THREADS, else min(12, hardware threads / ranks on this host).  (opts->ncores is NOT
// used for this: the reference's default ncores = 1 would serialise the ingest.)
static HostStager &stager_for(Ctx &cx) {
    int want = 0;
    if (const char *e = getenv("OEMB200_INGEST_THREADS")) want = atoi(e);
    if (want < 1) {
        const unsigned hc = std::thread::hardware_concurrency();
        want = (int)std::max(2u, std::min(12u, (hc ? hc : 8u) / (unsigned)std::max(1, cx.world)));
    }
    want = std::min(want, 32);
    if (!g_stager || g_stager->threads() != want || g_stager_device != cx.device) {
        g_stager.reset();
        g_stager.reset(new HostStager(want, (size_t)32 << 20, 8));
        g_stager_device = cx.device;
    }
    return *g_stager;
}

void release_host_stager() { g_stager.reset(); }

void h2d_block(Ctx &cx, const double *src, int64_t ldx, int64_t nr, int p, double *dst, int64_t ld_dst, cudaStream_t stream) {
    if (is_pinned_host(src)) {
        OEM_CUDA(cudaMemcpy2DAsync(dst, (size_t)ld_dst * 8, src, (size_t)ldx * 8, (size_t)nr * 8, p, cudaMemcpyHostToDevice, stream));
    } else {
        const auto t0 = std::chrono::steady_clock::now();
        stager_for(cx).copy_block(src, ldx, nr, p, dst, ld_dst, stream);
        cx.st.ms_ingest_wait += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    cx.st.h2d_bytes += nr * (int64_t)p * 8;
}

}  // namespace oemb200
