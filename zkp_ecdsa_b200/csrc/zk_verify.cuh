// zk_verify.cuh — batched verifier (verifySignatureList).  [under construction]
#pragma once
#include "zk_ops.cuh"

namespace zk {
// verifier tape: (2n+1) 32-byte GK drains, 78 index bytes (padded to 96), then 32-byte drains
ZK_LAYOUT_FN size_t verify_tape_len(int n, int reps) { return (size_t)32 * (2 * n + 1) + 96 + (size_t)32 * 25 * reps; }
}  // namespace zk
