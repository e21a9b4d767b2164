// Launchers shared between translation units of libsuper_b200.so (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include "super_b200.h"

namespace sbi {

// data_term.cu: the J^T J pass of the frame loop (fixed-point target chosen on the device, regularisers in the trailing
// blocks, per-warp loss partials, LM decision in the last block).  adopt != 0: prologue (no decision).
int launch_jtj_fused(const SbLMFrame* f, int adopt, cudaStream_t st);
int jtj_fused_partials(int n_cap);

}  // namespace sbi
