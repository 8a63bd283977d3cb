// Host-callable launcher of the draft-stage aligner (poa_align.cu).
#pragma once
#include <cuda_runtime.h>
#include "poa_device.h"

namespace ccs {

// Runs the DP (one warp per task) and the traceback (one thread per task) on `stream`.
// hrows may be NULL when every task is linear; paths may be NULL when only extents are wanted.
void launch_poa_align(const PoaTask* tasks, int n_tasks, const uint8_t* vbase, const int32_t* pred_off,
                      const int32_t* preds, const uint8_t* reads, int32_t* lo, int32_t* besti, uint8_t* moves,
                      int32_t* hrows, uint8_t* paths, PoaResult* results, cudaStream_t stream);

}  // namespace ccs
