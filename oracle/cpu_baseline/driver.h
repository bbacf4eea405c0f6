// driver.h — glue between the ISA instantiations and the threaded driver.
#pragma once
#include <atomic>
#include <cstdint>
#include <vector>
#include "../../include/frz_cuda.h"

struct FrzbJob {
    const uint8_t* needle; size_t needle_len; bool case_sensitive; int max_typos; size_t min_len; frz_scoring scoring;
    const uint8_t* bytes; const uint64_t* offsets; uint64_t n;
    std::atomic<uint64_t>* next_chunk; uint64_t n_chunks;
};
using FrzbWorker = void (*)(const FrzbJob&, std::vector<frz_match>&);

// One worker thread of match_list_parallel (src/matcher/parallel.rs:41-75): clone the matcher, claim
// 2048-item chunks, append matches in index order.
#define FRZB_DEFINE_WORKER(NAME, VTYPE)                                                               \
    void frzb_worker_##NAME(const FrzbJob& j, std::vector<frz_match>& local) {                        \
        SimdMatcher<VTYPE> m;                                                                         \
        m.init(j.needle, j.needle_len, j.case_sensitive, j.max_typos, j.min_len, j.scoring);          \
        for (;;) {                                                                                    \
            uint64_t ci = j.next_chunk->fetch_add(1, std::memory_order_relaxed);                      \
            if (ci >= j.n_chunks) break;                                                              \
            uint64_t lo = ci * 2048, hi = lo + 2048 < j.n ? lo + 2048 : j.n;                          \
            for (uint64_t i = lo; i < hi; i++) {                                                      \
                frz_match mt;                                                                         \
                if (m.match_one(j.bytes + j.offsets[i], (size_t)(j.offsets[i + 1] - j.offsets[i]), (uint32_t)i, &mt)) \
                    local.push_back(mt);                                                              \
            }                                                                                         \
        }                                                                                             \
    }

void frzb_worker_avx512(const FrzbJob&, std::vector<frz_match>&);
void frzb_worker_avx2(const FrzbJob&, std::vector<frz_match>&);
