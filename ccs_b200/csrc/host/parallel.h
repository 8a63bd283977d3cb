// Chunked parallel-for over independent items (ZMWs / reads) for the host-side stage logic.
#pragma once
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace ccs {

template <class F>
inline void parallel_for(int n, int n_threads, F&& f) {
    if (n <= 0) return;
    n_threads = std::max(1, std::min(n_threads, (n + 7) / 8));
    if (n_threads == 1) { for (int i = 0; i < n; ++i) f(i); return; }
    std::vector<std::thread> th;
    std::atomic<int> next(0);
    const int chunk = std::max(1, std::min(16, n / (4 * n_threads)));
    auto work = [&]() {
        for (;;) {
            const int b = next.fetch_add(chunk);
            if (b >= n) break;
            for (int i = b; i < std::min(n, b + chunk); ++i) f(i);
        }
    };
    for (int t = 1; t < n_threads; ++t) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
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
