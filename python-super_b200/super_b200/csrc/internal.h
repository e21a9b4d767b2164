// Launchers shared between translation units of libsuper_b200.so (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include "super_b200.h"

namespace sbi {

// data_term.cu: the two passes of the frame loop over the data term
int launch_eval_decide(const SbLMFrame* f, int adopt, cudaStream_t st);   // rows + keys + loss + LM decision (adopt: none)
int launch_gram(const SbLMFrame* f, cudaStream_t st, int which = 3);       // which: 1 = Gram records, 2 = scatter, 3 = both

}  // namespace sbi
