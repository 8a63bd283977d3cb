// Host-callable launchers of the Arrow kernels (implemented in the .cu files).
#pragma once
#include <cuda_runtime.h>
#include "arrow_device.h"

namespace ccs {

// order[n_items]: n_items is a multiple of 16; every aligned group of 16 entries (one CTA) holds reads of ONE ZMW,
// longest template first, padded with -1; a group's first entry is a real read
void launch_fill_alpha(const ArrowBatchView& V, const int32_t* order, int n_items, cudaStream_t stream);
void launch_fill_beta(const ArrowBatchView& V, const int32_t* order, int n_items, cudaStream_t stream);

// Recursor::EncodeRead on the device: job k encodes codes[src_off .. src_off + I) into the two row-code copies at
// rowcode[dst_off ..) (see arrow_pack.cu)
void launch_pack_rowcodes(const PackJob* jobs, int n_jobs, const uint8_t* codes, uint8_t* rowcode, cudaStream_t stream);

// delta[(zmw.delta_off + p) * kDeltaStride + slot]; INS total = slot[5+b] + slot[9+b].
// Ranges of one ZMW must be disjoint and non-touching.  generic = reference kernel (every mutation
// evaluated independently), used by the tests to cross-check the factored kernel.
void launch_score(const ArrowBatchView& V, const ScoreRange* ranges, int n_ranges, long long n_items, double* delta,
                  cudaStream_t stream, bool generic = false, int variant = 0);
// re-index delta rows after template edits: sites[] = edit positions in NEW coordinates (ascending per job),
// shifts[] = cumulative length change up to and including that edit
void launch_remap_delta(const RemapJob* jobs, int n_jobs, const int32_t* sites, const int32_t* shifts, double* delta,
                        double* scratch, cudaStream_t stream);
void launch_pick(const ArrowBatchView& V, const ScoreRange* ranges, int n_ranges, long long n_items, const double* delta,
                 Candidate* out, int cap, int* counter, cudaStream_t stream);
void launch_qv(const ArrowBatchView& V, const double* delta, uint8_t* qv, long long n_items, const ScoreRange* ranges,
               int n_ranges, cudaStream_t stream);

}  // namespace ccs
