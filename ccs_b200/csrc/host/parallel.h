// Chunked parallel-for over independent items (ZMWs / reads) for the host-side stage logic.
#pragma once
#include <algorithm>
#include <atomic>
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

}  // namespace ccs
