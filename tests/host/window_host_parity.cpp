// CPU parity test of the Polish Stage's windowing (ccs_b200/csrc/host/window_host.h: window layout, read slices from the
// mapping grid, core borders) against the oracle's restatement (oracle/pipeline_oracle.cpp make_windows / window_reads),
// on noisy forward / reverse / partial reads mapped by the oracle's aligner, plus unmapped reads, drafts below and above
// the split threshold and an empty window.  Built and run by tests/test_cpu_host.py::test_window_host_matches_oracle.
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
#include "../../ccs_b200/csrc/host/window_host.h"
#include "../../oracle/pipeline_oracle.h"

static int fail(const char* what, int trial) { std::fprintf(stderr, "MISMATCH: %s (trial %d)\n", what, trial); return 1; }

int main(int argc, char** argv) {
    const int trials = argc > 1 ? std::atoi(argv[1]) : 40;
    std::mt19937 rng(777u);
    long windows = 0, slices = 0, empties = 0;
    for (int trial = 0; trial < trials; ++trial) {
        const int nz = 1 + (int)(rng() % 3);
        const int W = 64 * (1 + (int)(rng() % 12)), OV = 64 * (int)(rng() % 3);
        std::vector<std::vector<uint8_t>> drafts(nz);
        std::vector<int32_t> zmw_read_off(1, 0);
        std::vector<int64_t> read_off(1, 0);
        std::vector<float> snr;
        std::vector<uint8_t> ok(nz, 1);
        std::vector<ccs::ReadMap> maps;
        std::vector<int32_t> grid;
        std::vector<int64_t> grid_off;
        std::vector<std::vector<oracle::ReadMapping>> omaps(nz);
        std::vector<std::vector<int>> olens(nz);
        for (int z = 0; z < nz; ++z) {
            const int J = 300 + (int)(rng() % 4000);
            drafts[z].resize(J);
            for (auto& b : drafts[z]) b = (uint8_t)(rng() & 3);
            for (int k = 0; k < 4; ++k) snr.push_back(8.f + (float)(rng() % 8));
            if (trial % 7 == 3 && z == 0) ok[z] = 0;
            const int nreads = 2 + (int)(rng() % 8);
            for (int r = 0; r < nreads; ++r) {
                // a noisy copy of a stretch of the draft: whole, a prefix, a suffix, or (trial % 5 == 4) only the first half
                int t0 = 0, t1 = J;
                const unsigned kind = rng() % 6;
                if (kind == 0) t1 = J / 2 + (int)(rng() % (J / 2));
                else if (kind == 1) t0 = (int)(rng() % (J / 2));
                if (trial % 5 == 4) t1 = std::min(t1, J / 2);
                std::vector<uint8_t> read;
                for (int i = t0; i < t1; ++i) {
                    const unsigned u = rng() % 1000;
                    if (u < 40) continue;
                    if (u < 90) read.push_back((uint8_t)(rng() & 3));
                    read.push_back(u < 110 ? (uint8_t)((drafts[z][i] + 1 + rng() % 3) & 3) : drafts[z][i]);
                }
                const bool junk = rng() % 12 == 0;
                if (junk) for (auto& b : read) b = (uint8_t)(rng() & 3);
                const int n = (int)read.size();
                const int strand = (int)(rng() & 1);
                // the oracle maps the ORIENTED read; native coordinates are flipped for reverse reads
                oracle::ReadMapping m = oracle::map_to_template(drafts[z].data(), J, read.data(), n);
                m.strand = strand;
                if (strand) { const int rs = n - m.rend, re = n - m.rstart; m.rstart = rs; m.rend = re; }
                if (m.mapped && (m.tend - m.tstart < 2 || m.rend - m.rstart < 2)) m.mapped = false;
                omaps[z].push_back(m);
                olens[z].push_back(n);
                ccs::ReadMap pm;
                pm.mapped = m.mapped; pm.strand = strand; pm.tstart = m.tstart; pm.tend = m.tend; pm.rstart = m.rstart;
                pm.rend = m.rend; pm.score = m.score;
                maps.push_back(pm);
                grid_off.push_back(m.grid.empty() ? -1 : (int64_t)grid.size());
                grid.insert(grid.end(), m.grid.begin(), m.grid.end());
                read_off.push_back(read_off.back() + n);
            }
            zmw_read_off.push_back((int32_t)maps.size());
        }
        grid.push_back(0);
        ccs::WindowParams wp;
        wp.size = W; wp.overlap = OV;
        ccs::WindowPlan P;
        ccs::build_window_plan(nz, zmw_read_off.data(), read_off.data(), snr.data(), drafts, ok.data(), maps.data(), grid.data(),
                               grid_off.data(), wp, P);
        for (int z = 0; z < nz; ++z) {
            const int w0 = P.zmw_win_off[z], w1 = P.zmw_win_off[z + 1];
            if (!ok[z]) { if (w1 != w0 || P.zmw_empty[z]) return fail("failed draft got windows", trial); continue; }
            const int J = (int)drafts[z].size();
            const std::vector<oracle::Window> ow = oracle::make_windows(J, W, OV);
            bool empty = false;
            std::vector<std::vector<oracle::WindowRead>> wr;
            for (const auto& w : ow) { wr.push_back(oracle::window_reads(w, omaps[z], olens[z])); if (wr.back().empty()) { empty = true; break; } }
            if (empty) {
                if (!P.zmw_empty[z] || w1 != w0) return fail("empty window not reported", trial);
                ++empties;
                continue;
            }
            if (P.zmw_empty[z] || w1 - w0 != (int)ow.size()) return fail("window count", trial);
            for (int k = 0; k < (int)ow.size(); ++k) {
                const int w = w0 + k;
                const oracle::Window& o = ow[k];
                if (P.win_zmw[w] != z || P.core_b[w] != o.c0 - o.a || P.core_e[w] != o.c1 - o.a) return fail("core borders", trial);
                if (P.tpl_off[w + 1] - P.tpl_off[w] != o.b - o.a) return fail("window length", trial);
                for (int j = o.a; j < o.b; ++j) if (P.tpl[P.tpl_off[w] + j - o.a] != drafts[z][j]) return fail("window bases", trial);
                for (int c = 0; c < 4; ++c) if (P.snr[4 * w + c] != snr[4 * z + c]) return fail("snr", trial);
                if (P.growth_min[w] != (ow.size() > 1 ? 128 : 512)) return fail("growth room", trial);
                const int r0 = P.win_read_off[w], r1 = P.win_read_off[w + 1];
                if (r1 - r0 != (int)wr[k].size()) return fail("reads of a window", trial);
                for (int x = 0; x < r1 - r0; ++x) {
                    const oracle::WindowRead& q = wr[k][x];
                    const int pr = zmw_read_off[z] + q.parent;
                    if (P.parent[r0 + x] != pr || P.strand[r0 + x] != q.strand || P.ts[r0 + x] != q.ts || P.te[r0 + x] != q.te ||
                        P.code_start[r0 + x] != read_off[pr] + q.ns || P.code_len[r0 + x] != q.ne - q.ns)
                        return fail("read slice", trial);
                    ++slices;
                }
                ++windows;
            }
        }
    }
    if (windows < trials || slices < 4 * trials) return fail("too few cases exercised", -1);
    std::printf("ok: %ld windows, %ld read slices, %ld empty-window ZMWs agree with the oracle\n", windows, slices, empties);
    return 0;
}
