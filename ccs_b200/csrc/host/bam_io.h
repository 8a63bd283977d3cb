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
    bool seek(uint64_t voffset);             // BGZF virtual offset: compressed block start << 16 | offset in block
    // Empty while the stream is sound.  A malformed / truncated block, an inflate failure, a CRC32 or ISIZE mismatch, or
    // a file that ends without the 28-byte BGZF EOF marker sets it: end of data and a damaged file are different things.
    const std::string& error() const { return error_; }
private:
    bool fill();                             // advance to the next non-empty block
    bool refill_batch();                     // read + inflate the next batch of blocks
    FILE* f_ = nullptr;
    int threads_ = 1;
    std::vector<std::vector<uint8_t>> blocks_, comp_;
    size_t n_batch_ = 0, cur_ = 0;           // blocks in the batch, index of the current one
    size_t pos_ = 0;
    bool failed_ = false;                    // a malformed block was seen: stop after the blocks before it
    bool last_block_empty_ = false;          // the last block read from the file was an empty one (the EOF marker)
    std::string error_;
};

class BgzfWriter {
public:
    ~BgzfWriter();
    bool open(const std::string& path, int level = 1, int threads = 0);
    void write(const void* src, size_t n);
    bool close();                             // flushes and appends the BGZF EOF marker; false if any write failed
    bool ok() const { return !io_error_; }    // false once an fwrite / deflate / fclose failed (e.g. disk full)
    uint64_t upos() const { return uflushed_ + buf_.size(); }   // uncompressed bytes written so far
    // virtual offset of an uncompressed position (valid after close(); blocks are recorded as they are written)
    uint64_t voffset_of(uint64_t upos) const;
private:
    void flush_block();
    FILE* f_ = nullptr;
    int level_ = 1, threads_ = 1;
    std::vector<uint8_t> buf_;
    std::vector<std::vector<uint8_t>> comp_;
    bool io_error_ = false;
    uint64_t uflushed_ = 0, cpos_ = 0;        // uncompressed / compressed bytes already written to the file
    std::vector<std::pair<uint64_t, uint64_t>> blocks_;   // (first uncompressed byte, compressed offset) of every block
};

// PacBio BAM index (*.pbi), basic section only: per record the read group id, query start / end, hole number, read
// quality, local-context flags and the BGZF virtual offset -- what `ccs --chunk i/N` needs to jump to its share of the
// ZMWs without inflating the rest (/root/reference/docs/faq/parallelize.md:8-13).  Layout as published with the PacBio
// BAM format 3.0.1 (magic "PBI\1", version, flags, n_reads, 18 reserved bytes, then one column per field), BGZF-wrapped.
struct PbiIndex {
    std::vector<int32_t> rg_id, q_start, q_end, hole;
    std::vector<float> read_qual;
    std::vector<uint8_t> ctxt;
    std::vector<int64_t> file_offset;
    size_t size() const { return hole.size(); }
    bool write(const std::string& path) const;
    bool read(const std::string& path);
    // first record of every ZMW (runs of equal hole numbers), plus size() as the end sentinel
    std::vector<int64_t> zmw_starts() const;
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

// The records of one ZMW as they sit in the file (validated, not decoded): the reader thread only groups records by
// hole number; sequence / pulse-width decoding runs in the stage workers (decode_zmw), in parallel.
struct RawZmw {
    int32_t hole = 0;
    std::vector<uint8_t> data;                // records back to back (without their block_size words)
    std::vector<uint32_t> rec_off;            // n_records + 1 offsets into data
    size_t n_records() const { return rec_off.empty() ? 0 : rec_off.size() - 1; }
};

// Decodes one validated record (SEQ + pw -> emission codes, zm qs qe cx sn tags).  false + err on a malformed record.
bool decode_subread_record(const uint8_t* rec, size_t size, Subread& s, std::vector<uint8_t>& pw_scratch, std::string& err);
bool decode_zmw(const RawZmw& raw, ZmwSubreads& z, std::string& err);

class SubreadBamReader {
public:
    // Opens and parses the header.  Fails (chemistry_ok() == false) if the read group lacks the
    // chemistry triple -- fatal in the reference too (docs/changelog.md:66, docs/faq/chemistry.md:7-10).
    bool open(const std::string& path, std::string& err, int threads = 0);   // threads: BGZF inflate workers
    bool next_zmw_raw(RawZmw& z);             // the same grouping without decoding the records (see RawZmw)
    bool next_zmw(ZmwSubreads& z);            // records of one hole number (consecutive in the file); false at the end
                                              // of the data AND on a damaged file: check error() afterwards
    const std::string& error() const { return error_.empty() ? in_.error() : error_; }
    bool seek_record(uint64_t voffset);       // continue with the record that starts at this BGZF virtual offset
    const std::string& header_text() const { return header_; }
    const std::string& movie() const { return movie_; }
    const std::string& read_group_id() const { return rg_id_; }
    bool chemistry_ok() const { return chem_ok_; }
private:
    bool next_record(Subread& s);
    bool next_raw_record(std::vector<uint8_t>& rec, int32_t& hole);   // reads + validates one record, finds its hole number
    BgzfReader in_;
    std::string header_, movie_, rg_id_;
    bool chem_ok_ = false, have_pending_ = false;
    Subread pending_;
    std::vector<uint8_t> pending_raw_;
    int32_t pending_raw_hole_ = 0;
    bool have_pending_raw_ = false;
    std::string error_;
    std::vector<uint8_t> rec_, pw_;           // record / 16-bit pulse-width scratch, reused across records
};

// `--chunk i/N`: positions `reader` (already opened on `path`) so that the caller, reading ZMW after ZMW from there and
// counting them from 0, keeps exactly those with index in [z_begin, z_end).  With <path>.pbi the reader is moved to the
// chunk's first record and nothing before it is inflated; without it the ZMWs are counted in a separate first pass.
bool select_chunk(SubreadBamReader& reader, const std::string& path, int chunk_i, int chunk_n, int64_t& z_begin,
                  int64_t& z_end, bool& used_index, std::string& err);

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
    void write(const CcsRecord& r, const char* suffix = "ccs");   // name movie/zmw/ccs[/fwd|/rev] (docs/faq/mode-by-strand.md:11-14)
    bool close();                             // false if the file could not be written completely
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
    void close();                             // also writes <path>.pbi
private:
    BgzfWriter out_;
    std::string path_, movie_, rg_;
    std::vector<uint8_t> rec_;
    PbiIndex pbi_;                            // file_offset holds uncompressed positions until close()
};

}  // namespace ccs
