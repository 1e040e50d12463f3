// kr_types.cuh -- plain types shared by the kernel headers (no device code here).
#pragma once
#include <stdint.h>

namespace kr {

typedef uint64_t u64;
typedef uint32_t u32;
typedef uint16_t u16;
typedef uint8_t u8;

constexpr int KR_INF = 9999;   // SequenceKmers.INFINITY

struct SeqMeta {   // one per sequence
    u32 size;      // number of distinct k-mers (SequenceKmers.size())
    u32 len;       // residues
    u64 hash;      // 64-bit hash of the residues (pre-filter for the equal-sequence rule)
    const u8 *res; // the residues themselves (exact check for the equal-sequence rule)
};

}  // namespace kr
