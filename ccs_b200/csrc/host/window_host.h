// Windowing of the Polish Stage (host side; "Divide the subread-to-draft alignment into overlapping
// windows ... Windowing reduces the algorithm run time from quadratic to linear in the insert size",
// /root/reference/docs/how-does-ccs-work.md:57-61; "per-window consensus template sequences and base qualities are
// concatenated and overhangs, overlaps between adjacent windows, are trimmed", :108-110).
//
// A draft of at least 2 * size bases becomes floor(J / size) windows: cores [k*size, (k+1)*size), the last one running to
// the end of the draft, each padded by `overlap` bases on both sides.  A window is an Arrow problem of its own: template =
// the padded draft slice, reads = the slices of the mapped subreads that the subread -> draft alignment places on it
// (the traceback of the mapping records the read position at every kWindowGrid-th draft base, so borders lie on that
// grid).  Every window is polished independently -- thousands of short dependent chains instead of a dozen long ones per
// ZMW, and a late edit refills one window instead of the whole molecule -- and the polished cores are concatenated.
// Plain C++ (no CUDA): tests/host/window_host_parity.cpp checks it against the oracle's restatement.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>
#include "../cuda/poa_device.h"
#include "draft_host.h"

namespace ccs {

struct WindowParams {
    int32_t size = 1024;      // core length (multiple of kWindowGrid); 0 = never split
    int32_t overlap = 64;     // padding on both sides of a core (multiple of kWindowGrid)
};

inline int32_t round_to_window_grid(int32_t v) { return v <= 0 ? 0 : (v + kWindowGrid - 1) / kWindowGrid * kWindowGrid; }

struct WindowSpec { int32_t a, b, c0, c1; };      // padded draft range [a, b), core [c0, c1)

inline void make_window_specs(int32_t J, const WindowParams& wp, std::vector<WindowSpec>& out) {
    out.clear();
    if (wp.size <= 0 || J < 2 * wp.size) { out.push_back(WindowSpec{0, J, 0, J}); return; }
    const int32_t n = J / wp.size;
    for (int32_t k = 0; k < n; ++k) {
        WindowSpec w;
        w.c0 = k * wp.size;
        w.c1 = (k == n - 1) ? J : (k + 1) * wp.size;
        w.a = std::max(0, w.c0 - wp.overlap);
        w.b = std::min(J, w.c1 + wp.overlap);
        out.push_back(w);
    }
}

// The Polish Stage input of one chunk after windowing: every window is a ZMW of the Arrow engine.
struct WindowPlan {
    // per parent ZMW
    std::vector<int32_t> zmw_win_off;     // [nz+1] -> windows
    std::vector<uint8_t> zmw_empty;       // 1: a window of this ZMW has no read (EMPTY_WINDOW_DURING_POLISHING); no windows kept
    // per window
    std::vector<int32_t> win_zmw;         // parent ZMW
    std::vector<int32_t> core_b, core_e;  // core borders in window coordinates
    std::vector<int32_t> win_read_off;    // [nw+1] -> window reads
    std::vector<int64_t> tpl_off;         // [nw+1]
    std::vector<uint8_t> tpl;             // padded draft slices
    std::vector<float> snr;               // [nw*4]
    std::vector<int32_t> growth_min;      // template growth room: 512 for a whole-draft window, 128 otherwise
    // per window read
    std::vector<int32_t> parent;          // batch read index
    std::vector<int64_t> code_start;      // slice of the batch's codes (absolute offset), native orientation
    std::vector<int32_t> code_len, ts, te;
    std::vector<uint8_t> strand;
    int32_t n_windows() const { return (int32_t)win_zmw.size(); }
    int32_t n_reads() const { return (int32_t)parent.size(); }
};

// zmw_ok[z]: the Draft Stage passed ZMW z.  maps / grid / grid_off: DraftOutput.
inline void build_window_plan(int32_t nz, const int32_t* zmw_read_off, const int64_t* read_off, const float* snr,
                              const std::vector<std::vector<uint8_t>>& drafts, const uint8_t* zmw_ok, const ReadMap* maps,
                              const int32_t* grid, const int64_t* grid_off, const WindowParams& wp, WindowPlan& P) {
    P = WindowPlan();
    P.zmw_win_off.assign((size_t)nz + 1, 0);
    P.zmw_empty.assign((size_t)nz, 0);
    P.win_read_off.push_back(0);
    P.tpl_off.push_back(0);
    std::vector<WindowSpec> specs;
    for (int32_t z = 0; z < nz; ++z) {
        P.zmw_win_off[z] = P.n_windows();
        if (!zmw_ok[z]) continue;
        const std::vector<uint8_t>& d = drafts[z];
        const int32_t J = (int32_t)d.size();
        make_window_specs(J, wp, specs);
        const size_t w_mark = P.win_zmw.size(), r_mark = P.parent.size(), t_mark = P.tpl.size();
        bool empty = false;
        for (const WindowSpec& w : specs) {
            const size_t r_before = P.parent.size();
            for (int32_t r = zmw_read_off[z]; r < zmw_read_off[z + 1]; ++r) {
                const ReadMap& m = maps[r];
                if (!m.mapped) continue;
                if (std::min(m.tend, w.c1) - std::max(m.tstart, w.c0) < 1) continue;     // does not reach the core
                const int32_t lo = std::max(w.a, m.tstart), hi = std::min(w.b, m.tend);
                if (hi - lo < 2) continue;
                const int32_t n = (int32_t)(read_off[r + 1] - read_off[r]);
                // extents of the aligned part in the ORIENTED read (the grid's coordinates)
                const int32_t rs_o = m.strand ? n - m.rend : m.rstart, re_o = m.strand ? n - m.rstart : m.rend;
                const int32_t s_o = (lo == m.tstart) ? rs_o : grid[grid_off[r] + lo / kWindowGrid];
                const int32_t e_o = (hi == m.tend) ? re_o : grid[grid_off[r] + hi / kWindowGrid];
                if (e_o - s_o < 2 || s_o < 0 || e_o > n) continue;       // (the path passes every grid point inside its span)
                const int32_t ns = m.strand ? n - e_o : s_o;                             // native slice [ns, ns + len)
                P.parent.push_back(r);
                P.code_start.push_back(read_off[r] + ns);
                P.code_len.push_back(e_o - s_o);
                P.ts.push_back(lo - w.a);
                P.te.push_back(hi - w.a);
                P.strand.push_back((uint8_t)m.strand);
            }
            if (P.parent.size() == r_before) { empty = true; break; }
            P.win_zmw.push_back(z);
            P.core_b.push_back(w.c0 - w.a);
            P.core_e.push_back(w.c1 - w.a);
            P.growth_min.push_back(specs.size() > 1 ? 128 : 512);
            P.win_read_off.push_back((int32_t)P.parent.size());
            P.tpl.insert(P.tpl.end(), d.begin() + w.a, d.begin() + w.b);
            P.tpl_off.push_back((int64_t)P.tpl.size());
            P.snr.insert(P.snr.end(), snr + 4 * (size_t)z, snr + 4 * (size_t)z + 4);
        }
        if (empty) {      // roll the ZMW back
            P.zmw_empty[z] = 1;
            P.win_zmw.resize(w_mark); P.core_b.resize(w_mark); P.core_e.resize(w_mark); P.growth_min.resize(w_mark);
            P.win_read_off.resize(w_mark + 1); P.tpl_off.resize(w_mark + 1); P.snr.resize(4 * w_mark);
            P.tpl.resize(t_mark);
            P.parent.resize(r_mark); P.code_start.resize(r_mark); P.code_len.resize(r_mark); P.ts.resize(r_mark);
            P.te.resize(r_mark); P.strand.resize(r_mark);
        }
    }
    P.zmw_win_off[nz] = P.n_windows();
    P.tpl.push_back(0);      // never an empty buffer
}

}  // namespace ccs
