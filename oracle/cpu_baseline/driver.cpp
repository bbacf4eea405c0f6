// driver.cpp — threaded match_list_parallel around the SIMD workers (src/matcher/parallel.rs:18-89,
// src/sort.rs:6-40, src/k_merge.rs:90-131).  MEASUREMENT INFRASTRUCTURE ONLY.
#include "driver.h"

#include <pthread.h>
#include <sched.h>

#include <algorithm>
#include <cstring>
#include <thread>

namespace {
bool has_avx512() {
    __builtin_cpu_init();
    return __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vbmi") &&
           __builtin_cpu_supports("bmi") && __builtin_cpu_supports("bmi2");
}
bool has_avx2() { __builtin_cpu_init(); return __builtin_cpu_supports("avx2") && __builtin_cpu_supports("bmi2"); }

// Worker ti runs on the ti-th CPU of the process's affinity mask (round-robin): rayon-style "one worker per core"
// placement, and the same placement on every call, which keeps the timing of the CPU arm repeatable on 2-socket hosts.
void pin_worker(size_t ti) {
    cpu_set_t all;
    CPU_ZERO(&all);
    if (sched_getaffinity(0, sizeof all, &all) != 0) return;
    const int ncpu = CPU_COUNT(&all);
    if (ncpu <= 0) return;
    int want = (int)(ti % (size_t)ncpu), seen = 0;
    for (int c = 0; c < CPU_SETSIZE; c++) {
        if (!CPU_ISSET(c, &all)) continue;
        if (seen++ == want) {
            cpu_set_t one;
            CPU_ZERO(&one);
            CPU_SET(c, &one);
            pthread_setaffinity_np(pthread_self(), sizeof one, &one);
            return;
        }
    }
}

void radix_sort_matches(std::vector<frz_match>& m) {  // src/sort.rs:6-40
    size_t n = m.size();
    std::vector<frz_match> b(n);
    uint32_t hist[256] = {0}, off[256] = {0};
    for (auto& x : m) hist[x.score & 0xFF]++;
    for (int i = 255; i >= 1; i--) off[i - 1] = off[i] + hist[i];
    for (auto& x : m) b[off[x.score & 0xFF]++] = x;
    memset(hist, 0, sizeof hist);
    for (auto& x : b) hist[(x.score >> 8) & 0xFF]++;
    off[255] = 0;
    for (int i = 255; i >= 1; i--) off[i - 1] = off[i] + hist[i];
    for (auto& x : b) m[off[(x.score >> 8) & 0xFF]++] = x;
}
uint32_t mpcb(const frz_scoring& s) {
    uint32_t bonus = std::max(s.delimiter_bonus, s.capitalization_bonus);
    uint32_t am = std::max((bonus + 1) / 2, bonus > s.gap_open_penalty ? bonus - s.gap_open_penalty : 0u);
    return am + s.matching_case_bonus;
}
uint32_t motb(const frz_scoring& s) {
    uint32_t bonus = std::max(s.delimiter_bonus, s.capitalization_bonus);
    uint32_t am = std::max((bonus + 1) / 2, bonus > s.gap_open_penalty ? bonus - s.gap_open_penalty : 0u);
    return bonus - am;
}
bool fits_u8(size_t n, const frz_scoring& s) {
    size_t mc = (size_t)s.match_score + s.mismatch_penalty;
    for (size_t v : {(size_t)s.gap_open_penalty, (size_t)s.gap_extend_penalty, (size_t)s.matching_case_bonus,
                     (size_t)s.capitalization_bonus, (size_t)s.delimiter_bonus, (size_t)s.prefix_bonus}) mc = std::max(mc, v);
    if (mc > 255 || 64 * (size_t)s.gap_extend_penalty + s.gap_open_penalty > 255) return false;
    return ((size_t)s.match_score + mpcb(s)) * n + motb(s) + s.prefix_bonus + s.mismatch_penalty <= 255;
}
}  // namespace

extern "C" const char* frzb_isa() { return has_avx512() ? "AVX-512BW+VBMI (64 x u8)" : has_avx2() ? "AVX2 (32 x u8)" : "none"; }

// Returns UINT64_MAX when the request is outside this restatement's scope (caller falls back to the scalar oracle).
extern "C" uint64_t frzb_match_list_parallel(const frz_pattern* patterns, size_t np, const frz_config* cfg, const uint8_t* bytes,
                                             const uint64_t* offsets, uint64_t n, int threads, frz_match* out, uint64_t cap) {
    if (np != 1 || patterns[0].negated || patterns[0].needle_len == 0) return UINT64_MAX;
    const frz_pattern& p = patterns[0];
    const int matching = p.matching >= 0 ? p.matching : cfg->matching;
    const int max_typos = p.max_typos >= 0 ? p.max_typos : cfg->max_typos;
    const int casing = p.casing >= 0 ? p.casing : cfg->casing;
    const frz_scoring sc = p.has_scoring ? p.scoring : cfg->scoring;
    if (matching != FRZ_MATCHING_FUZZY || max_typos > 1) return UINT64_MAX;
    bool cs = casing == FRZ_CASE_RESPECT;
    for (size_t i = 0; i < p.needle_len; i++) {
        if (p.needle[i] >= 0x80) return UINT64_MAX;
        if (casing == FRZ_CASE_SMART && p.needle[i] >= 'A' && p.needle[i] <= 'Z') cs = true;
    }
    if (!fits_u8(p.needle_len, sc)) return UINT64_MAX;
    const int lanes = cfg->emulate_lanes ? cfg->emulate_lanes : (has_avx512() ? 64 : 32);
    FrzbWorker worker = nullptr;
    if (lanes == 64 && has_avx512()) worker = frzb_worker_avx512;
    else if (lanes == 32 && has_avx2()) worker = frzb_worker_avx2;
    if (!worker) return UINT64_MAX;

    // thread clamp (src/matcher/parallel.rs:27)
    uint64_t t = std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)std::max(threads, 1), (n + 1999) / 2000));
    std::atomic<uint64_t> next{0};
    FrzbJob job{p.needle, p.needle_len, cs, max_typos,
                max_typos >= 0 ? (p.needle_len > (size_t)max_typos ? p.needle_len - max_typos : 0) : 0, sc,
                bytes, offsets, n, &next, (n + 2047) / 2048};
    const bool reversed = cfg->sort == FRZ_SORT_INDEX_DESC || cfg->sort == FRZ_SORT_SCORE_THEN_INDEX_DESC;
    const bool by_score = cfg->sort == FRZ_SORT_SCORE_THEN_INDEX_ASC || cfg->sort == FRZ_SORT_SCORE_THEN_INDEX_DESC;
    std::vector<std::vector<frz_match>> runs(t);
    auto body = [&](size_t ti) {
        std::vector<frz_match> local;  // thread-private while hot (no false sharing on the vector headers)
        local.reserve(4096);
        worker(job, local);
        // chunks are claimed in increasing order by each worker, so a run is already index-ascending
        if (reversed) std::reverse(local.begin(), local.end());
        if (by_score) radix_sort_matches(local);
        runs[ti] = std::move(local);
    };
    // all workers are spawned threads (the calling thread keeps its own affinity and only joins and merges)
    std::vector<std::thread> ths;
    for (size_t ti = 0; ti < t; ti++) ths.emplace_back([&body, ti] { pin_worker(ti); body(ti); });
    for (auto& th : ths) th.join();
    // k-way merge (src/k_merge.rs:90-131): binary heap of run cursors that carry their head Match inline,
    // advance-in-place + sift-down, and a bulk copy once a single run remains — as in the reference
    auto less = [&](const frz_match& a, const frz_match& b) {
        if (by_score) {
            if (a.score != b.score) return a.score > b.score;
            return reversed ? a.index > b.index : a.index < b.index;
        }
        return reversed ? a.index > b.index : a.index < b.index;
    };
    struct Cur { size_t run, pos; frz_match head; };
    std::vector<Cur> heap;
    heap.reserve(t);
    for (size_t r = 0; r < t; r++) if (!runs[r].empty()) heap.push_back({r, 0, runs[r][0]});
    auto sift_down = [&](size_t index) {
        size_t pos = index, child = 2 * pos + 1;
        while (child + 1 < heap.size()) {
            child += less(heap[child + 1].head, heap[child].head) ? 1 : 0;
            if (!less(heap[child].head, heap[pos].head)) return;
            std::swap(heap[pos], heap[child]);
            pos = child;
            child = 2 * pos + 1;
        }
        if (child < heap.size() && less(heap[child].head, heap[pos].head)) std::swap(heap[pos], heap[child]);
    };
    for (size_t i = heap.size() / 2; i-- > 0;) sift_down(i);
    uint64_t cnt = 0;
    while (heap.size() > 1) {
        if (cnt < cap) out[cnt] = heap[0].head;
        cnt++;
        const size_t run = heap[0].run, next = heap[0].pos + 1;
        if (next < runs[run].size()) { heap[0].pos = next; heap[0].head = runs[run][next]; }
        else { heap[0] = heap.back(); heap.pop_back(); }
        sift_down(0);
    }
    if (!heap.empty()) {
        const auto& r = runs[heap[0].run];
        for (size_t i = heap[0].pos; i < r.size(); i++) { if (cnt < cap) out[cnt] = r[i]; cnt++; }
    }
    return cnt;
}
