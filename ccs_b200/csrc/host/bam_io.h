// Minimal PacBio BAM I/O over zlib (BGZF blocks + BAM records), just what the `ccs` surface needs:
// read a *.subreads.bam grouped by ZMW, write an unaligned CCS BAM
// (/root/reference/docs/index.md:52-58; tags /root/reference/docs/faq/bam-output.md:9-30,45-49;
// SURVEY.md Appendix C).  No htslib / pbbam in this image, so the container format is restated here.
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

namespace ccs {

// BGZF blocks are independent gzip members: the reader pulls a batch of compressed blocks off the file and inflates
// them in parallel (host worker pool, parallel.h), the writer deflates the blocks of a flush in parallel and writes
// them in order -- at 3 k ZMW/s one GPU consumes ~0.8 GB/s of uncompressed subread records, about three zlib threads.
class BgzfReader {
public:
    ~BgzfReader();
    bool open(const std::string& path, int threads = 0);   // threads: 0 = min(8, hardware concurrency)
    bool read(void* dst, size_t n);          // false on EOF / error before n bytes
    bool eof();
private:
    bool fill();                             // advance to the next non-empty block
    bool refill_batch();                     // read + inflate the next batch of blocks
    FILE* f_ = nullptr;
    int threads_ = 1;
    std::vector<std::vector<uint8_t>> blocks_, comp_;
    size_t n_batch_ = 0, cur_ = 0;           // blocks in the batch, index of the current one
    size_t pos_ = 0;
    bool failed_ = false;                    // a malformed block was seen: stop after the blocks before it
};

class BgzfWriter {
public:
    ~BgzfWriter();
    bool open(const std::string& path, int level = 1, int threads = 0);
    void write(const void* src, size_t n);
    void close();                             // flushes and appends the BGZF EOF marker
private:
    void flush_block();
    FILE* f_ = nullptr;
    int level_ = 1, threads_ = 1;
    std::vector<uint8_t> buf_;
    std::vector<std::vector<uint8_t>> comp_;
};

struct Subread {
    int32_t hole = 0, qs = 0, qe = 0;
    float snr[4] = {0, 0, 0, 0};
    uint8_t cx = 0;
    std::vector<uint8_t> codes;               // 4*(min(pw,3)-1) + base, native orientation
};

struct ZmwSubreads {
    int32_t hole = 0;
    float snr[4] = {0, 0, 0, 0};
    std::vector<Subread> reads;
};

class SubreadBamReader {
public:
    // Opens and parses the header.  Fails (chemistry_ok() == false) if the read group lacks the
    // chemistry triple -- fatal in the reference too (docs/changelog.md:66, docs/faq/chemistry.md:7-10).
    bool open(const std::string& path, std::string& err, int threads = 0);   // threads: BGZF inflate workers
    bool next_zmw(ZmwSubreads& z);            // records of one hole number (consecutive in the file)
    const std::string& header_text() const { return header_; }
    const std::string& movie() const { return movie_; }
    const std::string& read_group_id() const { return rg_id_; }
    bool chemistry_ok() const { return chem_ok_; }
private:
    bool next_record(Subread& s);
    BgzfReader in_;
    std::string header_, movie_, rg_id_;
    bool chem_ok_ = false, have_pending_ = false;
    Subread pending_;
    std::vector<uint8_t> rec_, pw_;           // record / 16-bit pulse-width scratch, reused across records
};

struct CcsRecord {
    int32_t hole = 0, np = 0;
    float rq = 0, ec = 0, snr[4] = {0, 0, 0, 0};
    const uint8_t* seq = nullptr;             // bases 0..3
    const uint8_t* qv = nullptr;
    int32_t len = 0;
};

class CcsBamWriter {
public:
    // header derived from the input header: @RG DS:READTYPE=CCS (docs/faq/mode-heteroduplex-filtering.md:49-51)
    bool open(const std::string& path, const std::string& in_header, const std::string& movie, const std::string& rg_id,
              const std::string& program_cl);
    void write(const CcsRecord& r);           // name movie/zmw/ccs (docs/faq/mode-by-strand.md:11-14)
    void close();
private:
    BgzfWriter out_;
    std::string movie_, rg_;
    std::vector<uint8_t> rec_;
};

// Test / demo helper: writes a synthetic subreads.bam with the PacBio tags the reader consumes.
struct SubreadOut { int32_t hole, qs, qe; const float* snr; uint8_t cx; const uint8_t* codes; int32_t len; };
class SubreadBamWriter {
public:
    bool open(const std::string& path, const std::string& movie, bool with_chemistry = true, int threads = 0);
    void write(const SubreadOut& s);
    void close();
private:
    BgzfWriter out_;
    std::string movie_, rg_;
    std::vector<uint8_t> rec_;
};

}  // namespace ccs
