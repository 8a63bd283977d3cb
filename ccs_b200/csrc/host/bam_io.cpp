// See bam_io.h.  BGZF: RFC1952 gzip members with a 'BC' extra field carrying the block size;
// BAM: little-endian records (SAM spec section 4).  PacBio conventions per SURVEY.md Appendix C.
#include "bam_io.h"
#include "parallel.h"
#include <zlib.h>
#include <algorithm>
#include <atomic>
#include <cstring>
#include <thread>

namespace ccs {

namespace {

inline uint16_t rd16(const uint8_t* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
inline uint32_t rd32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
inline void wr16(std::vector<uint8_t>& v, uint16_t x) { v.push_back(x & 255); v.push_back(x >> 8); }
inline void wr32(std::vector<uint8_t>& v, uint32_t x) { for (int k = 0; k < 4; ++k) v.push_back((x >> (8 * k)) & 255); }
inline void wrf(std::vector<uint8_t>& v, float f) { uint32_t u; std::memcpy(&u, &f, 4); wr32(v, u); }
inline void wrs(std::vector<uint8_t>& v, const std::string& s) { v.insert(v.end(), s.begin(), s.end()); }

const uint8_t kBgzfEof[28] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43, 0x02, 0,
                              0x1b, 0, 0x03, 0, 0, 0, 0, 0, 0, 0, 0, 0};
constexpr size_t kBlockData = 0xff00;   // uncompressed payload per BGZF block

void tag_i(std::vector<uint8_t>& v, const char* t, int32_t x) { v.push_back(t[0]); v.push_back(t[1]); v.push_back('i'); wr32(v, (uint32_t)x); }
void tag_f(std::vector<uint8_t>& v, const char* t, float x) { v.push_back(t[0]); v.push_back(t[1]); v.push_back('f'); wrf(v, x); }
void tag_Z(std::vector<uint8_t>& v, const char* t, const std::string& s) { v.push_back(t[0]); v.push_back(t[1]); v.push_back('Z'); wrs(v, s); v.push_back(0); }
void tag_Bf(std::vector<uint8_t>& v, const char* t, const float* x, int n) {
    v.push_back(t[0]); v.push_back(t[1]); v.push_back('B'); v.push_back('f'); wr32(v, (uint32_t)n);
    for (int k = 0; k < n; ++k) wrf(v, x[k]);
}

// core of an unmapped BAM record; returns the offset of the block_size field to patch afterwards
size_t begin_record(std::vector<uint8_t>& v, const std::string& name, int32_t l_seq) {
    const size_t at = v.size();
    wr32(v, 0);                                  // block_size (patched)
    wr32(v, (uint32_t)-1); wr32(v, (uint32_t)-1);   // refID, pos
    v.push_back((uint8_t)(name.size() + 1));     // l_read_name
    v.push_back(255);                            // mapq
    wr16(v, 4680);                               // bin of an unmapped read (reg2bin(-1,0))
    wr16(v, 0);                                  // n_cigar_op
    wr16(v, 4);                                  // flag: unmapped
    wr32(v, (uint32_t)l_seq);
    wr32(v, (uint32_t)-1); wr32(v, (uint32_t)-1); wr32(v, 0);   // next refID, next pos, tlen
    wrs(v, name); v.push_back(0);
    return at;
}

void put_seq(std::vector<uint8_t>& v, const uint8_t* bases, int32_t n) {
    static const uint8_t nib[4] = {1, 2, 4, 8};   // =ACMGRSVTWYHKDBN
    for (int32_t i = 0; i < n; i += 2) {
        const uint8_t hi = nib[bases[i] & 3], lo = (i + 1 < n) ? nib[bases[i + 1] & 3] : 0;
        v.push_back((uint8_t)((hi << 4) | lo));
    }
}

void end_record(std::vector<uint8_t>& v, size_t at) {
    const uint32_t bs = (uint32_t)(v.size() - at - 4);
    for (int k = 0; k < 4; ++k) v[at + k] = (bs >> (8 * k)) & 255;
}

std::string header_field(const std::string& line, const std::string& key) {   // "\tKEY:value"
    const size_t p = line.find("\t" + key + ":");
    if (p == std::string::npos) return "";
    const size_t b = p + key.size() + 2, e = line.find('\t', b);
    return line.substr(b, e == std::string::npos ? std::string::npos : e - b);
}

}  // namespace

// ---------------------------------------------------------------------------------------------
static int default_io_threads(int threads) {
    if (threads > 0) return threads;
    int hc = (int)std::thread::hardware_concurrency();
    if (hc < 1) hc = 1;
    return std::min(8, hc);
}
constexpr size_t kReadBatch = 64;        // compressed blocks per parallel inflate (<= 4 MiB of payload)

BgzfReader::~BgzfReader() { if (f_) std::fclose(f_); }

bool BgzfReader::open(const std::string& path, int threads) {
    f_ = std::fopen(path.c_str(), "rb");
    threads_ = default_io_threads(threads);
    blocks_.assign(kReadBatch, {}); comp_.assign(kReadBatch, {});
    n_batch_ = 0; cur_ = 0; pos_ = 0;
    failed_ = false; last_block_empty_ = false; error_.clear();
    return f_ != nullptr;
}

bool BgzfReader::refill_batch() {
    if (failed_) return false;
    size_t n = 0;
    std::vector<uint32_t> isize(kReadBatch, 0);
    std::vector<size_t> clen(kReadBatch, 0);
    uint8_t h[18];
    auto fail = [&](const char* why) { failed_ = true; if (error_.empty()) error_ = why; };
    while (n < kReadBatch) {
        const size_t got = std::fread(h, 1, 18, f_);
        if (got == 0) {                                                   // physical end of the file
            if (!last_block_empty_) fail("BGZF EOF marker missing: the file is truncated");
            break;
        }
        // a malformed or truncated block ends the stream AFTER the complete blocks in front of it have been delivered
        if (got != 18) { fail("truncated BGZF block header"); break; }
        if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4) || rd16(h + 10) != 6 || h[12] != 'B' || h[13] != 'C') { fail("malformed BGZF block header"); break; }
        const size_t bsize = (size_t)rd16(h + 16) + 1;
        if (bsize < 18 + 8) { fail("malformed BGZF block size"); break; }
        clen[n] = bsize - 18 - 8;
        comp_[n].resize(clen[n] + 8);
        if (std::fread(comp_[n].data(), 1, clen[n] + 8, f_) != clen[n] + 8) { fail("truncated BGZF block"); break; }
        isize[n] = rd32(comp_[n].data() + clen[n] + 4);
        if (isize[n] > 65536u) { fail("BGZF block claims more than 64 KiB of payload"); break; }
        last_block_empty_ = isize[n] == 0;
        ++n;
    }
    if (n == 0) return false;
    std::atomic<int> bad(0);
    parallel_for((int)n, threads_, [&](int k) {
        blocks_[k].resize(isize[k]);
        if (isize[k] == 0) return;                                        // empty block (EOF marker or flush point)
        z_stream zs;
        std::memset(&zs, 0, sizeof(zs));
        if (inflateInit2(&zs, -15) != Z_OK) { bad.store(1); return; }
        zs.next_in = comp_[k].data(); zs.avail_in = (uInt)clen[k];
        zs.next_out = blocks_[k].data(); zs.avail_out = (uInt)isize[k];
        const int rc = inflate(&zs, Z_FINISH);
        const bool full = zs.total_out == isize[k];
        inflateEnd(&zs);
        if (rc != Z_STREAM_END || !full) { bad.store(1); return; }
        const uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), blocks_[k].data(), (uInt)isize[k]);
        if (crc != rd32(comp_[k].data() + clen[k])) bad.store(2);
    }, /*min_items_per_thread=*/1);
    if (bad.load()) { fail(bad.load() == 2 ? "BGZF block CRC32 mismatch" : "BGZF block does not inflate"); return false; }
    n_batch_ = n;
    return true;
}

bool BgzfReader::fill() {
    for (;;) {
        if (cur_ + 1 < n_batch_) ++cur_;
        else { if (!refill_batch()) { n_batch_ = 0; cur_ = 0; pos_ = 0; return false; } cur_ = 0; }
        pos_ = 0;
        if (!blocks_[cur_].empty()) return true;
    }
}

bool BgzfReader::read(void* dst, size_t n) {
    uint8_t* d = (uint8_t*)dst;
    while (n > 0) {
        if ((n_batch_ == 0 || pos_ == blocks_[cur_].size()) && !fill()) return false;
        const std::vector<uint8_t>& blk = blocks_[cur_];
        const size_t k = std::min(n, blk.size() - pos_);
        std::memcpy(d, blk.data() + pos_, k);
        pos_ += k; d += k; n -= k;
    }
    return true;
}

bool BgzfReader::seek(uint64_t voffset) {
    if (!f_) return false;
    if (fseeko(f_, (off_t)(voffset >> 16), SEEK_SET) != 0) return false;
    n_batch_ = 0; cur_ = 0; pos_ = 0; failed_ = false; last_block_empty_ = false; error_.clear();
    const size_t within = (size_t)(voffset & 0xffff);
    if (!fill()) return within == 0;          // seeking to the very end is fine
    if (within > blocks_[cur_].size()) return false;
    pos_ = within;
    return true;
}

bool BgzfReader::eof() {
    if (n_batch_ != 0 && pos_ < blocks_[cur_].size()) return false;
    return !fill();
}

BgzfWriter::~BgzfWriter() { close(); }

bool BgzfWriter::open(const std::string& path, int level, int threads) {
    f_ = std::fopen(path.c_str(), "wb");
    level_ = level;
    threads_ = default_io_threads(threads);
    buf_.clear();
    uflushed_ = 0; cpos_ = 0; blocks_.clear(); io_error_ = false;
    return f_ != nullptr;
}

uint64_t BgzfWriter::voffset_of(uint64_t upos) const {
    // last block whose first uncompressed byte is <= upos
    size_t lo = 0, hi = blocks_.size();
    while (lo + 1 < hi) { const size_t mid = (lo + hi) / 2; if (blocks_[mid].first <= upos) lo = mid; else hi = mid; }
    if (blocks_.empty()) return 0;
    return (blocks_[lo].second << 16) | (upos - blocks_[lo].first);
}

void BgzfWriter::flush_block() {
    if (buf_.empty() || !f_) return;
    const size_t nb = (buf_.size() + kBlockData - 1) / kBlockData;
    if (comp_.size() < nb) comp_.resize(nb);
    std::vector<size_t> clen(nb, 0);
    std::vector<uint32_t> crc(nb, 0);
    std::atomic<int> bad(0);
    parallel_for((int)nb, threads_, [&](int k) {
        const size_t done = (size_t)k * kBlockData;
        const size_t n = std::min(kBlockData, buf_.size() - done);
        comp_[k].resize(compressBound((uLong)n) + 64);
        z_stream zs;
        std::memset(&zs, 0, sizeof(zs));
        if (deflateInit2(&zs, level_, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) { bad.store(1); return; }
        zs.next_in = buf_.data() + done; zs.avail_in = (uInt)n;
        zs.next_out = comp_[k].data(); zs.avail_out = (uInt)comp_[k].size();
        if (deflate(&zs, Z_FINISH) != Z_STREAM_END) bad.store(1);
        clen[k] = zs.total_out;
        deflateEnd(&zs);
        crc[k] = (uint32_t)crc32(crc32(0L, Z_NULL, 0), buf_.data() + done, (uInt)n);
    }, /*min_items_per_thread=*/1);
    if (bad.load()) io_error_ = true;
    for (size_t k = 0; k < nb; ++k) {
        const size_t n = std::min(kBlockData, buf_.size() - k * kBlockData);
        uint8_t h[18] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0, 0};
        const uint16_t bsize = (uint16_t)(clen[k] + 18 + 8 - 1);
        h[16] = bsize & 255; h[17] = bsize >> 8;
        blocks_.push_back({uflushed_ + k * kBlockData, cpos_});
        uint8_t t[8];
        for (int b = 0; b < 4; ++b) { t[b] = (crc[k] >> (8 * b)) & 255; t[4 + b] = ((uint32_t)n >> (8 * b)) & 255; }
        if (std::fwrite(h, 1, 18, f_) != 18 || std::fwrite(comp_[k].data(), 1, clen[k], f_) != clen[k] ||
            std::fwrite(t, 1, 8, f_) != 8) io_error_ = true;
        cpos_ += 18 + clen[k] + 8;
    }
    uflushed_ += buf_.size();
    buf_.clear();
}

void BgzfWriter::write(const void* src, size_t n) {
    const uint8_t* s = (const uint8_t*)src;
    buf_.insert(buf_.end(), s, s + n);
    if (buf_.size() >= 64 * kBlockData) flush_block();
}

bool BgzfWriter::close() {
    if (!f_) return !io_error_;
    flush_block();
    blocks_.push_back({uflushed_, cpos_});        // the EOF marker block: where a position at the very end maps to
    if (std::fwrite(kBgzfEof, 1, sizeof(kBgzfEof), f_) != sizeof(kBgzfEof)) io_error_ = true;
    if (std::fclose(f_) != 0) io_error_ = true;
    f_ = nullptr;
    return !io_error_;
}

// ---------------------------------------------------------------------------------------------
bool SubreadBamReader::open(const std::string& path, std::string& err, int threads) {
    if (!in_.open(path, threads)) { err = "cannot open " + path; return false; }
    uint8_t magic[4];
    if (!in_.read(magic, 4) || std::memcmp(magic, "BAM\1", 4) != 0) { err = path + " is not a BAM file"; return false; }
    uint8_t b4[4];
    if (!in_.read(b4, 4)) { err = "truncated header"; return false; }
    header_.resize(rd32(b4));
    if (!header_.empty() && !in_.read(&header_[0], header_.size())) { err = "truncated header"; return false; }
    while (!header_.empty() && header_.back() == 0) header_.pop_back();
    if (!in_.read(b4, 4)) { err = "truncated header"; return false; }
    for (uint32_t r = rd32(b4); r > 0; --r) {     // reference sequences (none in PacBio unaligned BAMs)
        if (!in_.read(b4, 4)) return false;
        std::vector<uint8_t> skip(rd32(b4) + 4);
        if (!in_.read(skip.data(), skip.size())) return false;
    }
    // read group: movie (PU), id, chemistry triple in DS
    size_t p = 0;
    while (p < header_.size()) {
        size_t e = header_.find('\n', p);
        if (e == std::string::npos) e = header_.size();
        const std::string line = header_.substr(p, e - p);
        if (line.compare(0, 3, "@RG") == 0 && rg_id_.empty()) {
            rg_id_ = header_field(line, "ID");
            movie_ = header_field(line, "PU");
            const std::string ds = header_field(line, "DS");
            chem_ok_ = ds.find("BINDINGKIT=") != std::string::npos && ds.find("SEQUENCINGKIT=") != std::string::npos &&
                       ds.find("BASECALLERVERSION=") != std::string::npos;
        }
        p = e + 1;
    }
    if (rg_id_.empty()) { err = "no @RG line in the BAM header"; return false; }
    return true;
}

// Every length field of a record is checked against the record before anything is decoded.  Returns the offset of the
// first tag, or 0 with err set.
static size_t validate_record(const uint8_t* r, size_t size, std::string& err) {
    auto bad = [&](const char* why) { err = std::string("malformed BAM record: ") + why; return (size_t)0; };
    if (size < 32) return bad("implausible record size");
    const int l_name = r[8];
    const int n_cigar = rd16(r + 12);
    const int32_t l_seq = (int32_t)rd32(r + 16);
    if (l_seq < 0) return bad("negative l_seq");
    const uint64_t fixed = 32ull + (uint64_t)l_name + 4ull * (uint64_t)n_cigar + ((uint64_t)l_seq + 1) / 2 + (uint64_t)l_seq;
    if (fixed > size) return bad("name / cigar / sequence lengths exceed the record");
    // walk the tags once: types and sizes must stay inside the record
    const uint8_t* p = r + fixed;
    const uint8_t* end = r + size;
    while (p + 3 <= end) {
        const char ty = (char)p[2];
        p += 3;
        size_t sz = 0;
        switch (ty) {
            case 'A': case 'c': case 'C': sz = 1; break;
            case 's': case 'S': sz = 2; break;
            case 'i': case 'I': case 'f': sz = 4; break;
            case 'Z': case 'H': {
                const void* z = std::memchr(p, 0, (size_t)(end - p));
                if (!z) return bad("unterminated string tag");
                sz = (size_t)((const uint8_t*)z - p) + 1;
                break;
            }
            case 'B': {
                if (p + 5 > end) return bad("truncated array tag");
                const char sub = (char)p[0];
                const uint32_t n = rd32(p + 1);
                const size_t es = (sub == 'c' || sub == 'C') ? 1 : ((sub == 's' || sub == 'S') ? 2 : 4);
                if ((uint64_t)es * n + 5 > (uint64_t)(end - p)) return bad("array tag exceeds the record");
                sz = 5 + es * n;
                break;
            }
            default: return bad("unknown tag type");
        }
        if (sz > (size_t)(end - p)) return bad("tag exceeds the record");
        p += sz;
    }
    return (size_t)fixed;
}

// hole number of a validated record: the zm tag, else the read name movie/zmw/qs_qe
static int32_t record_hole(const uint8_t* r, size_t size, size_t tags_off) {
    const uint8_t* p = r + tags_off;
    const uint8_t* end = r + size;
    while (p + 3 <= end) {
        const char t0 = (char)p[0], t1 = (char)p[1], ty = (char)p[2];
        p += 3;
        size_t sz = 0;
        switch (ty) {
            case 'A': case 'c': case 'C': sz = 1; break;
            case 's': case 'S': sz = 2; break;
            case 'i': case 'I': case 'f': sz = 4; break;
            case 'Z': case 'H': sz = std::strlen((const char*)p) + 1; break;
            case 'B': { const char sub = (char)p[0]; const uint32_t n = rd32(p + 1);
                        sz = 5 + (size_t)((sub == 'c' || sub == 'C') ? 1 : ((sub == 's' || sub == 'S') ? 2 : 4)) * n; break; }
            default: return 0;
        }
        if (t0 == 'z' && t1 == 'm') {
            if (ty == 'i' || ty == 'I') return (int32_t)rd32(p);
            if (ty == 's') return (int16_t)rd16(p);
            if (ty == 'S') return rd16(p);
            if (ty == 'c') return (int8_t)p[0];
            if (ty == 'C') return p[0];
        }
        p += sz;
    }
    const int l_name = r[8];
    const std::string name((const char*)r + 32, l_name > 0 ? l_name - 1 : 0);
    const size_t a = name.find('/'), b = name.find('/', a + 1);
    if (a != std::string::npos && b != std::string::npos) return std::atoi(name.substr(a + 1, b - a - 1).c_str());
    return 0;
}

bool SubreadBamReader::next_raw_record(std::vector<uint8_t>& rec, int32_t& hole) {
    uint8_t b4[4];
    if (!error_.empty()) return false;
    if (in_.eof()) return false;                 // clean end of the data (or a damaged container: in_.error())
    auto bad = [&](const char* why) { error_ = std::string("malformed BAM record: ") + why; return false; };
    if (!in_.read(b4, 4)) return bad("truncated record length");
    const uint32_t rec_size = rd32(b4);
    constexpr uint32_t kMaxRecord = 64u << 20;   // a subread of 20 M bases: far beyond any real polymerase read
    if (rec_size < 32 || rec_size > kMaxRecord) return bad("implausible record size");
    rec.resize(rec_size);
    if (!in_.read(rec.data(), rec.size())) return bad("truncated record");
    const size_t tags = validate_record(rec.data(), rec.size(), error_);
    if (!tags) return false;
    hole = record_hole(rec.data(), rec.size(), tags);
    return true;
}

bool decode_subread_record(const uint8_t* r, size_t size, Subread& s, std::vector<uint8_t>& pw, std::string& err) {
    if (!validate_record(r, size, err)) return false;
    auto bad = [&](const char* why) { err = std::string("malformed BAM record: ") + why; return false; };
    const int l_name = r[8];
    const int n_cigar = rd16(r + 12);
    const int32_t l_seq = (int32_t)rd32(r + 16);
    const uint8_t* p = r + 32;
    const std::string name((const char*)p, l_name > 0 ? l_name - 1 : 0);
    p += l_name + 4 * n_cigar;
    const uint8_t* seq = p;
    p += (l_seq + 1) / 2 + l_seq;               // packed bases + qualities
    s.hole = 0; s.qs = 0; s.qe = 0; s.cx = 0;
    s.snr[0] = s.snr[1] = s.snr[2] = s.snr[3] = 0.f;
    const uint8_t* pw8 = nullptr;                 // 8-bit pulse widths are used in place
    uint32_t n_pw = 0;
    const uint8_t* end = r + size;
    while (p + 3 <= end) {
        const char t0 = (char)p[0], t1 = (char)p[1], ty = (char)p[2];
        p += 3;
        auto is = [&](const char* t) { return t0 == t[0] && t1 == t[1]; };
        size_t sz = 0;
        switch (ty) {
            case 'A': case 'c': case 'C': sz = 1; break;
            case 's': case 'S': sz = 2; break;
            case 'i': case 'I': case 'f': sz = 4; break;
            case 'Z': case 'H': {
                const void* z = std::memchr(p, 0, (size_t)(end - p));
                if (!z) return bad("unterminated string tag");
                sz = (size_t)((const uint8_t*)z - p) + 1;
                break;
            }
            case 'B': {
                if (p + 5 > end) return bad("truncated array tag");
                const char sub = (char)p[0];
                const uint32_t n = rd32(p + 1);
                const size_t es = (sub == 'c' || sub == 'C') ? 1 : ((sub == 's' || sub == 'S') ? 2 : 4);
                if ((uint64_t)es * n + 5 > (uint64_t)(end - p)) return bad("array tag exceeds the record");
                if (is("sn") && sub == 'f' && n == 4) for (int k = 0; k < 4; ++k) { uint32_t u = rd32(p + 5 + 4 * k); std::memcpy(&s.snr[k], &u, 4); }
                if (is("pw")) {
                    n_pw = n;
                    if (es == 1) pw8 = p + 5;
                    else {
                        pw.resize(n);
                        for (uint32_t k = 0; k < n; ++k) pw[k] = (uint8_t)std::min<uint32_t>(255, es == 2 ? rd16(p + 5 + 2 * k) : rd32(p + 5 + 4 * k));
                        pw8 = pw.data();
                    }
                }
                sz = 5 + es * n;
                break;
            }
            default: return bad("unknown tag type");
        }
        if (sz > (size_t)(end - p)) return bad("tag exceeds the record");
        if (ty != 'B') {
            int32_t iv = 0;
            if (ty == 'i' || ty == 'I') iv = (int32_t)rd32(p);
            else if (ty == 's') iv = (int16_t)rd16(p); else if (ty == 'S') iv = rd16(p);
            else if (ty == 'c') iv = (int8_t)p[0]; else if (ty == 'C') iv = p[0];
            if (is("zm")) s.hole = iv;
            else if (is("qs")) s.qs = iv;
            else if (is("qe")) s.qe = iv;
            else if (is("cx")) s.cx = (uint8_t)iv;
        }
        p += sz;
    }
    if (s.hole == 0 && !name.empty()) {          // fall back to the read name movie/zmw/qs_qe
        const size_t a = name.find('/'), b = name.find('/', a + 1);
        if (a != std::string::npos && b != std::string::npos) s.hole = std::atoi(name.substr(a + 1, b - a - 1).c_str());
    }
    // Recursor::EncodeRead: code = 4 * (min(max(pw, 1), 3) - 1) + base; two bases per packed byte through a table
    struct Lut {
        uint8_t hi[256], lo[256], w4[256];
        Lut() {
            static const int8_t dec[16] = {-1, 0, 1, -1, 2, -1, -1, -1, 3, -1, -1, -1, -1, -1, -1, -1};   // =ACMGRSVTWYHKDBN
            for (int v = 0; v < 256; ++v) {
                hi[v] = (uint8_t)(dec[v >> 4] < 0 ? 0 : dec[v >> 4]);
                lo[v] = (uint8_t)(dec[v & 15] < 0 ? 0 : dec[v & 15]);
                w4[v] = (uint8_t)(4 * (std::min(std::max(v, 1), 3) - 1));
            }
        }
    };
    static const Lut lut;
    s.codes.resize(l_seq);
    uint8_t* out = s.codes.data();
    const int32_t n_w = (int32_t)std::min<uint32_t>(n_pw, (uint32_t)l_seq);   // bases beyond the pw array: width 1
    int32_t i = 0;
    for (; i + 1 < n_w; i += 2) {
        const uint8_t v = seq[i >> 1];
        out[i] = (uint8_t)(lut.w4[pw8[i]] + lut.hi[v]);
        out[i + 1] = (uint8_t)(lut.w4[pw8[i + 1]] + lut.lo[v]);
    }
    for (; i < l_seq; ++i) {
        const uint8_t v = seq[i >> 1];
        const uint8_t b = (i & 1) ? lut.lo[v] : lut.hi[v];
        out[i] = (uint8_t)((i < n_w ? lut.w4[pw8[i]] : 0) + b);
    }
    return true;
}

bool SubreadBamReader::next_record(Subread& s) {
    int32_t hole = 0;
    if (!next_raw_record(rec_, hole)) return false;
    return decode_subread_record(rec_.data(), rec_.size(), s, pw_, error_);
}

bool decode_zmw(const RawZmw& raw, ZmwSubreads& z, std::string& err) {
    z.reads.clear();
    z.hole = raw.hole;
    std::vector<uint8_t> pw;
    z.reads.resize(raw.n_records());
    for (size_t k = 0; k < raw.n_records(); ++k)
        if (!decode_subread_record(raw.data.data() + raw.rec_off[k], raw.rec_off[k + 1] - raw.rec_off[k], z.reads[k], pw, err)) return false;
    if (!z.reads.empty()) std::memcpy(z.snr, z.reads[0].snr, sizeof(z.snr));
    return true;
}

bool SubreadBamReader::next_zmw_raw(RawZmw& z) {
    z.data.clear(); z.rec_off.assign(1, 0);
    if (!have_pending_raw_) {
        if (!next_raw_record(pending_raw_, pending_raw_hole_)) return false;
        have_pending_raw_ = true;
    }
    z.hole = pending_raw_hole_;
    while (have_pending_raw_ && pending_raw_hole_ == z.hole) {
        z.data.insert(z.data.end(), pending_raw_.begin(), pending_raw_.end());
        z.rec_off.push_back((uint32_t)z.data.size());
        have_pending_raw_ = next_raw_record(pending_raw_, pending_raw_hole_);
    }
    // a damaged file ends the stream with an error; the ZMW being assembled may be missing subreads and is dropped
    if (!error().empty()) { z.data.clear(); z.rec_off.assign(1, 0); have_pending_raw_ = false; return false; }
    return true;
}

bool SubreadBamReader::seek_record(uint64_t voffset) {
    have_pending_ = false;
    have_pending_raw_ = false;
    return in_.seek(voffset);
}

bool select_chunk(SubreadBamReader& reader, const std::string& path, int chunk_i, int chunk_n, int64_t& z_begin,
                  int64_t& z_end, bool& used_index, std::string& err) {
    used_index = false;
    if (chunk_n < 1 || chunk_i < 1 || chunk_i > chunk_n) { err = "--chunk expects i/N with i in [1,N]"; return false; }
    PbiIndex pbi;
    if (pbi.read(path + ".pbi") && pbi.size() > 0) {
        const std::vector<int64_t> st = pbi.zmw_starts();
        const int64_t total = (int64_t)st.size() - 1;
        const int64_t zb = total * (chunk_i - 1) / chunk_n, ze = total * chunk_i / chunk_n;
        z_begin = 0;
        z_end = ze - zb;
        if (ze > zb && !reader.seek_record((uint64_t)pbi.file_offset[st[zb]])) {
            err = path + ".pbi does not match " + path;
            return false;
        }
        used_index = true;
        return true;
    }
    SubreadBamReader counter;
    if (!counter.open(path, err)) return false;
    ZmwSubreads z;
    int64_t total = 0;
    while (counter.next_zmw(z)) ++total;
    z_begin = total * (chunk_i - 1) / chunk_n;
    z_end = total * chunk_i / chunk_n;
    return true;
}

bool SubreadBamReader::next_zmw(ZmwSubreads& z) {
    z.reads.clear();
    if (!have_pending_) {
        if (!next_record(pending_)) return false;
        have_pending_ = true;
    }
    z.hole = pending_.hole;
    std::memcpy(z.snr, pending_.snr, sizeof(z.snr));
    while (have_pending_ && pending_.hole == z.hole) {
        z.reads.push_back(std::move(pending_));
        have_pending_ = next_record(pending_);
    }
    // a damaged file ends the stream with an error; the ZMW being assembled may be missing subreads and is dropped
    if (!error().empty()) { z.reads.clear(); have_pending_ = false; return false; }
    return true;
}

// ---------------------------------------------------------------------------------------------
bool CcsBamWriter::open(const std::string& path, const std::string& in_header, const std::string& movie,
                        const std::string& rg_id, const std::string& program_cl) {
    if (!out_.open(path)) return false;
    movie_ = movie; rg_ = rg_id;
    std::string text;
    size_t p = 0;
    while (p < in_header.size()) {
        size_t e = in_header.find('\n', p);
        if (e == std::string::npos) e = in_header.size();
        std::string line = in_header.substr(p, e - p);
        const size_t q = line.find("READTYPE=SUBREAD");
        if (q != std::string::npos) line.replace(q, 16, "READTYPE=CCS");
        if (!line.empty()) text += line + "\n";
        p = e + 1;
    }
    text += "@PG\tID:ccs\tPN:ccs\tVN:ccs-b200-0.1\tDS:Generate circular consensus sequences (ccs) from subreads.\tCL:" + program_cl + "\n";
    std::vector<uint8_t> h;
    wrs(h, std::string("BAM\1", 4));
    wr32(h, (uint32_t)text.size());
    wrs(h, text);
    wr32(h, 0);
    out_.write(h.data(), h.size());
    return true;
}

void CcsBamWriter::write(const CcsRecord& r, const char* suffix) {
    rec_.clear();
    const std::string name = movie_ + "/" + std::to_string(r.hole) + "/" + suffix;
    const size_t at = begin_record(rec_, name, r.len);
    put_seq(rec_, r.seq, r.len);
    rec_.insert(rec_.end(), r.qv, r.qv + r.len);
    tag_Z(rec_, "RG", rg_);                 // fixed per-read tags: RG zm np rq sn ec (docs/faq/bam-output.md:45-49)
    tag_i(rec_, "zm", r.hole);
    tag_i(rec_, "np", r.np);
    tag_f(rec_, "rq", r.rq);
    tag_Bf(rec_, "sn", r.snr, 4);
    tag_f(rec_, "ec", r.ec);
    end_record(rec_, at);
    out_.write(rec_.data(), rec_.size());
}

bool CcsBamWriter::close() { return out_.close(); }

bool SubreadBamWriter::open(const std::string& path, const std::string& movie, bool with_chemistry, int threads) {
    if (!out_.open(path, 1, threads)) return false;
    path_ = path; pbi_ = PbiIndex();
    movie_ = movie; rg_ = "b200sim0";
    std::string ds = "READTYPE=SUBREAD;Ipd:CodecV1=ip;PulseWidth:CodecV1=pw";
    if (with_chemistry) ds += ";BINDINGKIT=000-000-000;SEQUENCINGKIT=000-000-001;BASECALLERVERSION=0.0.0;FRAMERATEHZ=100.000000";
    const std::string text = "@HD\tVN:1.5\tSO:unknown\tpb:3.0.1\n@RG\tID:" + rg_ + "\tPL:PACBIO\tDS:" + ds + "\tPU:" + movie_ +
                             "\tPM:SEQUELII\n";
    std::vector<uint8_t> h;
    wrs(h, std::string("BAM\1", 4));
    wr32(h, (uint32_t)text.size());
    wrs(h, text);
    wr32(h, 0);
    out_.write(h.data(), h.size());
    return true;
}

void SubreadBamWriter::write(const SubreadOut& s) {
    rec_.clear();
    const std::string name = movie_ + "/" + std::to_string(s.hole) + "/" + std::to_string(s.qs) + "_" + std::to_string(s.qe);
    const size_t at = begin_record(rec_, name, s.len);
    std::vector<uint8_t> bases(s.len);
    for (int32_t i = 0; i < s.len; ++i) bases[i] = s.codes[i] & 3;
    put_seq(rec_, bases.data(), s.len);
    rec_.insert(rec_.end(), (size_t)s.len, (uint8_t)0xff);   // subreads carry no qualities
    tag_Z(rec_, "RG", rg_);
    tag_i(rec_, "zm", s.hole);
    tag_i(rec_, "qs", s.qs);
    tag_i(rec_, "qe", s.qe);
    tag_i(rec_, "cx", s.cx);
    tag_Bf(rec_, "sn", s.snr, 4);
    tag_f(rec_, "rq", 0.8f);
    rec_.push_back('p'); rec_.push_back('w'); rec_.push_back('B'); rec_.push_back('C'); wr32(rec_, (uint32_t)s.len);
    for (int32_t i = 0; i < s.len; ++i) rec_.push_back((uint8_t)((s.codes[i] >> 2) + 1));
    end_record(rec_, at);
    pbi_.rg_id.push_back(0x0b200510);             // numeric form of the read-group id
    pbi_.q_start.push_back(s.qs); pbi_.q_end.push_back(s.qe); pbi_.hole.push_back(s.hole);
    pbi_.read_qual.push_back(0.8f); pbi_.ctxt.push_back(s.cx);
    pbi_.file_offset.push_back((int64_t)out_.upos());
    out_.write(rec_.data(), rec_.size());
}

void SubreadBamWriter::close() {
    if (path_.empty()) return;
    out_.close();
    for (auto& o : pbi_.file_offset) o = (int64_t)out_.voffset_of((uint64_t)o);
    pbi_.write(path_ + ".pbi");
    path_.clear();
}

// ---------------------------------------------------------------------------------------------
bool PbiIndex::write(const std::string& path) const {
    BgzfWriter w;
    if (!w.open(path, 1, 1)) return false;
    std::vector<uint8_t> h;
    wrs(h, std::string("PBI\1", 4));
    wr32(h, 0x00030001u);                          // 3.0.1
    wr16(h, 0);                                    // basic section only
    wr32(h, (uint32_t)size());
    h.resize(h.size() + 18, 0);
    w.write(h.data(), h.size());
    const size_t n = size();
    if (n) {
        w.write(rg_id.data(), 4 * n); w.write(q_start.data(), 4 * n); w.write(q_end.data(), 4 * n);
        w.write(hole.data(), 4 * n); w.write(read_qual.data(), 4 * n); w.write(ctxt.data(), n);
        w.write(file_offset.data(), 8 * n);
    }
    w.close();
    return true;
}

bool PbiIndex::read(const std::string& path) {
    BgzfReader r;
    if (!r.open(path, 1)) return false;
    uint8_t h[32];
    if (!r.read(h, 32) || std::memcmp(h, "PBI\1", 4) != 0) return false;
    const uint32_t n = rd32(h + 10);
    rg_id.resize(n); q_start.resize(n); q_end.resize(n); hole.resize(n); read_qual.resize(n); ctxt.resize(n);
    file_offset.resize(n);
    if (n == 0) return true;
    return r.read(rg_id.data(), 4 * (size_t)n) && r.read(q_start.data(), 4 * (size_t)n) && r.read(q_end.data(), 4 * (size_t)n) &&
           r.read(hole.data(), 4 * (size_t)n) && r.read(read_qual.data(), 4 * (size_t)n) && r.read(ctxt.data(), n) &&
           r.read(file_offset.data(), 8 * (size_t)n);
}

std::vector<int64_t> PbiIndex::zmw_starts() const {
    std::vector<int64_t> st;
    for (size_t k = 0; k < size(); ++k) if (k == 0 || hole[k] != hole[k - 1]) st.push_back((int64_t)k);
    st.push_back((int64_t)size());
    return st;
}

}  // namespace ccs
