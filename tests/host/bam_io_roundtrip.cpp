// CPU round-trip test of the PacBio BAM reader / writer (ccs_b200/csrc/host/bam_io.*): a synthetic subreads.bam
// written with block-parallel deflate must read back record for record with block-parallel inflate, the bytes on
// disk must not depend on the number of threads, and a truncated file must deliver its complete leading ZMWs and
// then stop cleanly.  Built and run by tests/test_cpu_host.py::test_bam_io_round_trip.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <random>
#include <string>
#include <vector>
#include "../../ccs_b200/csrc/host/bam_io.h"

using namespace ccs;

struct Z { int hole; float snr[4]; std::vector<std::vector<uint8_t>> reads; std::vector<uint8_t> cx; std::vector<int> qs, qe; };

static std::vector<uint8_t> slurp(const std::string& p) {
    std::ifstream f(p, std::ios::binary);
    return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

static bool write_bam(const std::string& path, const std::vector<Z>& zs, int threads) {
    SubreadBamWriter w;
    if (!w.open(path, "m64001_261017_000000", true, threads)) return false;
    for (const Z& z : zs)
        for (size_t r = 0; r < z.reads.size(); ++r) {
            SubreadOut s{z.hole, z.qs[r], z.qe[r], z.snr, z.cx[r], z.reads[r].data(), (int32_t)z.reads[r].size()};
            w.write(s);
        }
    w.close();
    return true;
}

int main(int argc, char** argv) {
    const std::string dir = argc > 1 ? argv[1] : "/tmp";
    std::mt19937 rng(99u);
    std::vector<Z> zs;
    long bases = 0;
    for (int k = 0; k < 150; ++k) {
        Z z;
        z.hole = 4194368 + 7 * k;
        for (int c = 0; c < 4; ++c) z.snr[c] = 5.f + (float)(rng() % 1000) / 100.f;
        const int nr = 1 + (int)(rng() % 14);
        int q = 0;
        for (int r = 0; r < nr; ++r) {
            const int len = 1 + (int)(rng() % ((k % 9 == 0) ? 30000 : 6000));     // some reads span many BGZF blocks
            std::vector<uint8_t> codes(len);
            for (auto& c : codes) c = (uint8_t)(rng() % 12);
            z.reads.push_back(codes);
            z.cx.push_back((uint8_t)(rng() & 3));
            z.qs.push_back(q); z.qe.push_back(q + len); q += len + 40;
            bases += len;
        }
        zs.push_back(z);
    }
    const std::string p1 = dir + "/rt_t1.subreads.bam", p8 = dir + "/rt_t8.subreads.bam";
    if (!write_bam(p1, zs, 1) || !write_bam(p8, zs, 8)) { std::fprintf(stderr, "cannot write\n"); return 1; }
    const std::vector<uint8_t> b1 = slurp(p1), b8 = slurp(p8);
    if (b1 != b8) { std::fprintf(stderr, "MISMATCH: file bytes depend on the thread count (%zu vs %zu)\n", b1.size(), b8.size()); return 1; }
    for (int threads : {1, 8}) {
        SubreadBamReader rd;
        std::string err;
        if (!rd.open(p8, err, threads)) { std::fprintf(stderr, "open: %s\n", err.c_str()); return 1; }
        if (!rd.chemistry_ok() || rd.movie() != "m64001_261017_000000") { std::fprintf(stderr, "MISMATCH: header\n"); return 1; }
        ZmwSubreads got;
        for (const Z& z : zs) {
            if (!rd.next_zmw(got)) { std::fprintf(stderr, "MISMATCH: early EOF at hole %d\n", z.hole); return 1; }
            if (got.hole != z.hole || got.reads.size() != z.reads.size()) { std::fprintf(stderr, "MISMATCH: hole / read count at %d\n", z.hole); return 1; }
            for (int c = 0; c < 4; ++c) if (got.snr[c] != z.snr[c]) { std::fprintf(stderr, "MISMATCH: snr\n"); return 1; }
            for (size_t r = 0; r < z.reads.size(); ++r) {
                const Subread& s = got.reads[r];
                if (s.codes != z.reads[r] || s.cx != z.cx[r] || s.qs != z.qs[r] || s.qe != z.qe[r]) { std::fprintf(stderr, "MISMATCH: read %zu of hole %d\n", r, z.hole); return 1; }
            }
        }
        if (rd.next_zmw(got)) { std::fprintf(stderr, "MISMATCH: records after the last ZMW\n"); return 1; }
        if (!rd.error().empty()) { std::fprintf(stderr, "MISMATCH: sound file reported '%s'\n", rd.error().c_str()); return 1; }
    }
    // truncated copy: every ZMW delivered before the cut is COMPLETE, then the stream ends with an error (a cut file is
    // not a short file: the BGZF EOF marker is missing and the last block is damaged)
    {
        const std::string pt = dir + "/rt_trunc.subreads.bam";
        std::ofstream f(pt, std::ios::binary);
        f.write((const char*)b8.data(), (std::streamsize)(b8.size() * 3 / 5));
        f.close();
        SubreadBamReader rd;
        std::string err;
        if (!rd.open(pt, err, 8)) { std::fprintf(stderr, "open truncated: %s\n", err.c_str()); return 1; }
        ZmwSubreads got;
        size_t k = 0;
        while (rd.next_zmw(got)) {
            if (k >= zs.size() || got.hole != zs[k].hole) { std::fprintf(stderr, "MISMATCH: truncated stream out of order\n"); return 1; }
            if (got.reads.size() != zs[k].reads.size()) { std::fprintf(stderr, "MISMATCH: truncated stream delivered an incomplete ZMW\n"); return 1; }
            ++k;
        }
        if (k == 0 || k >= zs.size()) { std::fprintf(stderr, "MISMATCH: truncated stream delivered %zu of %zu ZMWs\n", k, zs.size()); return 1; }
        if (rd.error().empty()) { std::fprintf(stderr, "MISMATCH: truncated file read without an error\n"); return 1; }
    }
    // cut exactly at a block boundary (complete blocks, no EOF marker): still an error
    {
        size_t pos = 0, last = 0;
        while (pos + 18 <= b8.size()) { const size_t bs = (size_t)(b8[pos + 16] | (b8[pos + 17] << 8)) + 1; last = pos; pos += bs; }
        const std::string pt = dir + "/rt_noeof.subreads.bam";
        std::ofstream f(pt, std::ios::binary);
        f.write((const char*)b8.data(), (std::streamsize)last);     // drops the final (EOF marker) block
        f.close();
        SubreadBamReader rd;
        std::string err;
        if (!rd.open(pt, err, 2)) return 1;
        ZmwSubreads got;
        while (rd.next_zmw(got)) {}
        if (rd.error().find("EOF marker") == std::string::npos) { std::fprintf(stderr, "MISMATCH: missing EOF marker not reported ('%s')\n", rd.error().c_str()); return 1; }
    }
    // a corrupted payload byte fails the block's CRC32
    {
        std::vector<uint8_t> c = b8;
        c[c.size() / 2] ^= 0x5a;
        const std::string pt = dir + "/rt_crc.subreads.bam";
        std::ofstream f(pt, std::ios::binary);
        f.write((const char*)c.data(), (std::streamsize)c.size());
        f.close();
        SubreadBamReader rd;
        std::string err;
        if (rd.open(pt, err, 2)) {
            ZmwSubreads got;
            while (rd.next_zmw(got)) {}
            if (rd.error().empty()) { std::fprintf(stderr, "MISMATCH: corrupted block read without an error\n"); return 1; }
        }
    }
    // a record whose length fields lie (l_seq = 50 M in a 40-byte record) is rejected, not decoded
    {
        const std::string pt = dir + "/rt_badrec.subreads.bam";
        {
            BgzfWriter w;
            if (!w.open(pt, 1, 1)) return 1;
            const std::string text = "@HD\tVN:1.5\n@RG\tID:x\tPL:PACBIO\tDS:READTYPE=SUBREAD;BINDINGKIT=1;SEQUENCINGKIT=2;BASECALLERVERSION=3\tPU:m\n";
            std::vector<uint8_t> h = {'B', 'A', 'M', 1};
            auto w32 = [&](uint32_t x) { for (int b = 0; b < 4; ++b) h.push_back((x >> (8 * b)) & 255); };
            w32((uint32_t)text.size()); h.insert(h.end(), text.begin(), text.end()); w32(0);
            w32(40);                                   // block_size
            w32(0xffffffffu); w32(0xffffffffu);        // refID, pos
            h.push_back(2); h.push_back(255); h.push_back(0x48); h.push_back(0x12);   // l_read_name, mapq, bin
            h.push_back(0); h.push_back(0); h.push_back(4); h.push_back(0);           // n_cigar, flag
            w32(50000000u);                            // l_seq
            w32(0xffffffffu); w32(0xffffffffu); w32(0);
            h.push_back('a'); h.push_back(0);
            for (int k = 0; k < 6; ++k) h.push_back(0);
            w.write(h.data(), h.size());
            if (!w.close()) return 1;
        }
        SubreadBamReader rd;
        std::string err;
        if (!rd.open(pt, err, 1)) { std::fprintf(stderr, "open badrec: %s\n", err.c_str()); return 1; }
        ZmwSubreads got;
        if (rd.next_zmw(got) || rd.error().find("malformed BAM record") == std::string::npos) { std::fprintf(stderr, "MISMATCH: lying record accepted ('%s')\n", rd.error().c_str()); return 1; }
    }
    // .pbi: one entry per record, ZMW runs, and seeking to the first record of any ZMW continues the stream there
    {
        PbiIndex pbi;
        if (!pbi.read(p8 + ".pbi")) { std::fprintf(stderr, "MISMATCH: cannot read the .pbi\n"); return 1; }
        size_t n_rec = 0;
        for (const Z& z : zs) n_rec += z.reads.size();
        if (pbi.size() != n_rec) { std::fprintf(stderr, "MISMATCH: .pbi has %zu records, file has %zu\n", pbi.size(), n_rec); return 1; }
        const std::vector<int64_t> st = pbi.zmw_starts();
        if (st.size() != zs.size() + 1) { std::fprintf(stderr, "MISMATCH: .pbi ZMW runs\n"); return 1; }
        size_t rec = 0;
        for (size_t k = 0; k < zs.size(); ++k)
            for (size_t r = 0; r < zs[k].reads.size(); ++r, ++rec)
                if (pbi.hole[rec] != zs[k].hole || pbi.q_start[rec] != zs[k].qs[r] || pbi.q_end[rec] != zs[k].qe[r] || pbi.ctxt[rec] != zs[k].cx[r]) { std::fprintf(stderr, "MISMATCH: .pbi record %zu\n", rec); return 1; }
        SubreadBamReader rd;
        std::string err;
        if (!rd.open(p8, err, 4)) return 1;
        for (size_t k : {zs.size() - 1, (size_t)0, zs.size() / 2, (size_t)1, zs.size() / 3}) {
            if (!rd.seek_record((uint64_t)pbi.file_offset[st[k]])) { std::fprintf(stderr, "MISMATCH: seek to ZMW %zu\n", k); return 1; }
            ZmwSubreads got;
            for (size_t j = k; j < std::min(zs.size(), k + 3); ++j) {
                if (!rd.next_zmw(got) || got.hole != zs[j].hole || got.reads.size() != zs[j].reads.size()) { std::fprintf(stderr, "MISMATCH: ZMW %zu after seeking to %zu\n", j, k); return 1; }
                for (size_t r = 0; r < zs[j].reads.size(); ++r) if (got.reads[r].codes != zs[j].reads[r]) { std::fprintf(stderr, "MISMATCH: read after seek\n"); return 1; }
            }
        }
    }
    // --chunk i/N through select_chunk(): the chunks concatenate to the whole file, with the index and without it
    {
        auto chunk_holes = [&](const std::string& path, int i, int n, bool expect_index, std::vector<int>& holes) -> bool {
            SubreadBamReader rd;
            std::string err;
            if (!rd.open(path, err, 2)) return false;
            int64_t zb = 0, ze = 0;
            bool used = false;
            if (!select_chunk(rd, path, i, n, zb, ze, used, err) || used != expect_index) return false;
            // the loop of ccs_main.cpp
            ZmwSubreads z;
            int64_t z_index = 0;
            if (ze <= zb) return true;
            while (rd.next_zmw(z)) {
                if (z_index >= zb && z_index < ze) holes.push_back(z.hole);
                ++z_index;
                if (z_index >= ze) break;
            }
            return true;
        };
        const std::string pn = dir + "/rt_noindex.subreads.bam";
        { std::ofstream f(pn, std::ios::binary); f.write((const char*)b8.data(), (std::streamsize)b8.size()); }
        for (int n : {1, 2, 3, 7, 40}) {
            std::vector<int> with_idx, without_idx;
            for (int i = 1; i <= n; ++i) {
                if (!chunk_holes(p8, i, n, true, with_idx) || !chunk_holes(pn, i, n, false, without_idx)) { std::fprintf(stderr, "MISMATCH: chunk %d/%d failed\n", i, n); return 1; }
            }
            if (with_idx.size() != zs.size() || without_idx != with_idx) { std::fprintf(stderr, "MISMATCH: chunks of N=%d do not tile the file\n", n); return 1; }
            for (size_t k = 0; k < zs.size(); ++k) if (with_idx[k] != zs[k].hole) { std::fprintf(stderr, "MISMATCH: chunk order N=%d\n", n); return 1; }
        }
        // more chunks than ZMWs: single chunks agree between the two modes (most are empty)
        for (int i : {1, 57, 200, 399, 400}) {
            std::vector<int> a, b;
            if (!chunk_holes(p8, i, 400, true, a) || !chunk_holes(pn, i, 400, false, b) || a != b || a.size() > 1) { std::fprintf(stderr, "MISMATCH: chunk %d/400\n", i); return 1; }
        }
    }
    std::printf("ok: %zu ZMWs, %ld bases round-tripped; %zu bytes on disk, identical for 1 and 8 threads\n", zs.size(), bases, b8.size());
    return 0;
}
