// align_costs.h — cost model of the alignment kernels as plain structs (shared by host launchers and kernels).
#pragma once
#include <cstdint>

namespace swb {

struct AlignCosts {
    int match, mismatch; // unary path: S = classes equal ? match : mismatch
    int first, next;     // gap costs (<= 0 in maximising form): length-k gap = first + (k-1)*next
    int ident;           // 1: symbols are compared as raw bytes (no class map)
    int negate;          // 1: write -H (Levenshtein distance)
};

// `byte_to_class` / `class_costs` of `NeedlemanWunschScores::new` (StringWars similarities/bench.rs:98-108, 658-662).
struct AlignTables {
    uint8_t byte_to_class[256];
    int8_t costs[32 * 32];   // costs[class(a)][class(b)]
    int8_t costs_t[32 * 32]; // transposed, used when the kernel swapped the pair so rows = shorter string
};

} // namespace swb
