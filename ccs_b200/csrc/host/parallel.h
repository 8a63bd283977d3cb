// Chunked parallel-for over independent items (ZMWs / reads) for the host-side stage logic.
#pragma once
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <deque>
#include <memory>
#include <type_traits>
#include <ctime>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace ccs {

// Persistent worker pool behind parallel_for: the stage logic calls parallel_for ~70 times per lane and batch, and
// spawning + joining up to 16 threads per call cost ~0.15 host core-ms per ZMW.  Workers are created once (detached,
// never destroyed: the library may be unloaded by a host process at exit while they sleep on the condition variable).
class HostPool {
public:
    struct Job {
        std::atomic<int> next{0};
        std::atomic<int> active{0};          // helpers currently inside the work loop
        int n = 0, chunk = 1;
        const void* fn = nullptr;
        void (*call)(const void*, int) = nullptr;
        void work() {
            for (;;) {
                const int b = next.fetch_add(chunk);
                if (b >= n) break;
                const int e = std::min(n, b + chunk);
                for (int i = b; i < e; ++i) call(fn, i);
            }
        }
    };
    static HostPool& get() { static HostPool* p = new HostPool(); return *p; }
    void submit(const std::shared_ptr<Job>& job, int copies) {
        {
            std::lock_guard<std::mutex> g(mu_);
            for (int k = 0; k < copies; ++k) q_.push_back(job);
        }
        if (copies == 1) cv_.notify_one(); else cv_.notify_all();
    }
private:
    HostPool() {
        int n = (int)std::thread::hardware_concurrency();
        if (n < 4) n = 4;
        for (int t = 0; t < n; ++t) std::thread([this]() { loop(); }).detach();
    }
    void loop() {
        for (;;) {
            std::shared_ptr<Job> job;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&]() { return !q_.empty(); });
                job = std::move(q_.front());
                q_.pop_front();
            }
            // a helper that arrives after the work is gone never touches the caller's function object
            job->active.fetch_add(1);
            job->work();
            job->active.fetch_sub(1);
        }
    }
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<std::shared_ptr<Job>> q_;
};

// min_items_per_thread: small items (per-ZMW host logic) are not worth a helper for fewer than 8 of them; heavy items
// (a 64 KB zlib block) are.
template <class F>
inline void parallel_for(int n, int n_threads, F&& f, int min_items_per_thread = 8) {
    if (n <= 0) return;
    n_threads = std::max(1, std::min(n_threads, (n + min_items_per_thread - 1) / min_items_per_thread));
    if (n_threads == 1) { for (int i = 0; i < n; ++i) f(i); return; }
    auto job = std::make_shared<HostPool::Job>();
    job->n = n;
    job->chunk = std::max(1, std::min(16, n / (4 * n_threads)));
    job->fn = &f;
    job->call = [](const void* fn, int i) { (*static_cast<const std::remove_reference_t<F>*>(fn))(i); };
    HostPool::get().submit(job, n_threads - 1);
    job->work();
    // every chunk has been claimed; wait for the helpers still inside one (a claimed chunk implies active > 0, and
    // the increment is ordered before the claim)
    while (job->active.load() != 0) std::this_thread::yield();
}

// Host-side phase accounting (wall + process CPU time), enabled by CCS_B200_HOST_PROFILE=1 and printed to stderr by
// host_profile_dump().  Meaningful with one lane and one context (process CPU time is shared by all threads).
struct HostProf {
    struct Acc { double wall = 0, cpu = 0; long n = 0; };
    static bool on() { static const bool v = [] { const char* e = std::getenv("CCS_B200_HOST_PROFILE"); return e && e[0] == '1'; }(); return v; }
    static std::map<std::string, Acc>& table() { static std::map<std::string, Acc> t; return t; }
    static std::mutex& mu() { static std::mutex m; return m; }
    static double now_wall() { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
    static double now_cpu() { timespec t; clock_gettime(CLOCK_PROCESS_CPUTIME_ID, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
    static void dump() {
        if (!on()) return;
        std::lock_guard<std::mutex> g(mu());
        for (auto& kv : table())
            std::fprintf(stderr, "[host-profile] %-28s calls %6ld  wall %8.1f ms  cpu %8.1f ms\n", kv.first.c_str(), kv.second.n,
                         1e3 * kv.second.wall, 1e3 * kv.second.cpu);
        table().clear();
    }
};
struct HostPhase {
    const char* name; double w0 = 0, c0 = 0;
    explicit HostPhase(const char* n) : name(n) { if (HostProf::on()) { w0 = HostProf::now_wall(); c0 = HostProf::now_cpu(); } }
    ~HostPhase() {
        if (!HostProf::on()) return;
        const double w = HostProf::now_wall() - w0, c = HostProf::now_cpu() - c0;
        std::lock_guard<std::mutex> g(HostProf::mu());
        auto& a = HostProf::table()[name];
        a.wall += w; a.cpu += c; ++a.n;
    }
};

}  // namespace ccs
