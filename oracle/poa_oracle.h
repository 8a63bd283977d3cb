// TEST INFRASTRUCTURE -- CPU restatement of the Draft Stage:
// FilterReads, k-mer orientation vote (the seeding half of SdpRangeFinder), SparsePoa
// (banded local sequence-to-DAG alignment, CommitAdd, FindConsensus) and the subread -> draft
// mapping.  PARITY UNPINNED: see arrow_oracle.h; behaviour anchored on
// /root/reference/docs/how-does-ccs-work.md:19-55 and SURVEY.md Appendix B; every rule below
// that the docs leave open (band rule, tie breaks, vertex placement) is fixed in DESIGN.md
// "Draft stage" and restated here independently of the product code in ccs_b200/.
#pragma once
#include <cstdint>
#include <vector>

namespace oracle {

constexpr int POA_MATCH = 3, POA_MISMATCH = -5, POA_INS = -4, POA_DEL = -4;
constexpr int POA_BAND = 64;
// band rule (DESIGN.md "Draft stage"): rows are taken in blocks of POA_BLOCK; the band of every row of a block is placed
// from the block's anchor row (the last row of the previous block)
constexpr int POA_BLOCK = 32;
constexpr int POA_ANCHOR_MIN = 20;     // an anchor row whose best score is lower carries no alignment yet
constexpr int POA_BAND_DECAY = 16;     // ... the band then moves this many cells back towards the read start
constexpr int POA_KMER = 11;
constexpr int POA_VOTE_BASES = 2048;   // orientation vote: k-mers of the read's first 2048 bases (DESIGN.md "Draft stage")

enum PoaMove : uint8_t { PM_STOP = 0, PM_MATCH = 1, PM_DEL = 2, PM_INS = 3 };

struct PathStep { int vertex; int readpos; uint8_t move; };   // MATCH: both; DEL: vertex only; INS: readpos only

struct PoaAlignment {
    int score = 0;
    std::vector<PathStep> path;   // start -> end
};

struct PoaGraph {
    // col: seed coordinate of the vertex -- its index for a seed vertex; a vertex added later takes the value of the
    // vertex it was placed behind (0 at the head)
    struct Vertex { uint8_t base; int nreads; int next, prev; std::vector<int> in, out; int col = 0; };
    std::vector<Vertex> v;
    int head = -1, tail = -1;
    std::vector<std::pair<int, int>> spans;   // (first, last) vertex of every threaded read
    int n_reads = 0;

    void add_first(const uint8_t* seq, int n);
    void order(std::vector<int>& ord, std::vector<int>& rank) const;
    PoaAlignment align(const uint8_t* seq, int n) const;
    void commit(const PoaAlignment& a, const uint8_t* seq);
    // consensus vertices in order; min_cov per SURVEY.md Appendix B
    std::vector<int> consensus(int min_cov) const;
};

// true if the reverse complement of `read` shares more K-mers with `ref` than `read` does
bool kmer_vote_reverse(const uint8_t* ref, int nref, const uint8_t* read, int n);

constexpr int WINDOW_GRID = 64;        // windowing (DESIGN.md "Windowing"): window borders lie on multiples of 64 template bases

struct ReadMapping {
    int strand = 0, tstart = 0, tend = 0, rstart = 0, rend = 0, score = 0;
    bool mapped = false;
    // grid[k] = number of bases of the ORIENTED read placed before template position k * WINDOW_GRID on the alignment
    // path (the base matched there, or the next one when the position is deleted); -1 where the path does not pass
    std::vector<int> grid;
};
// align read (already oriented by the vote) to a linear template; extents of the aligned part
ReadMapping map_to_template(const uint8_t* tpl, int J, const uint8_t* read_bases, int n);

}  // namespace oracle
