/* ffi_demo.c — a host program that uses ONLY include/frz_cuda.h (plain C, no CUDA headers, no Python, no torch):
 * the same calls a Rust `extern "C"` binding makes.
 *
 *   gcc -std=c11 -Iinclude examples/ffi_demo.c -Lfrizbee_b200 -lfrz_cuda -Wl,-rpath,$PWD/frizbee_b200 -o ffi_demo
 *   ./ffi_demo            # BASELINE.json configs[0]: needle "fBr" vs 5 haystacks → [Match{score 53, index 0}]
 *
 * Exit code 0: matched and printed; 3: no CUDA device (the library never falls back to the CPU). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "frz_cuda.h"

int main(void) {
    static const char* hay[] = {"fooBar", "foo_bar", "barfoo", "prelude", "println!"};
    enum { N = 5 };
    /* Arrow-style buffers: concatenated bytes + offsets */
    uint8_t bytes[256];
    uint64_t offsets[N + 1] = {0};
    for (int i = 0; i < N; i++) {
        size_t len = strlen(hay[i]);
        memcpy(bytes + offsets[i], hay[i], len);
        offsets[i + 1] = offsets[i] + len;
    }
    printf("frz ABI version %u\n", frz_abi_version());

    frz_config cfg;
    frz_config_default(&cfg);
    cfg.max_typos = 0;

    frz_pattern pat;
    memset(&pat, 0, sizeof pat);
    pat.needle = (const uint8_t*)"fBr";
    pat.needle_len = 3;
    pat.casing = -1; pat.unicode = -1; pat.matching = -1; pat.max_typos = -1;   /* inherit from the config */

    frz_matcher* m = NULL;
    frz_status st = frz_matcher_create(&pat, 1, &cfg, &m);
    if (st != FRZ_OK) { fprintf(stderr, "matcher: %s: %s\n", frz_status_str(st), frz_last_error()); return 1; }

    frz_corpus* corpus = NULL;
    st = frz_corpus_create(bytes, offsets, N, /*device=*/0, &corpus);
    if (st != FRZ_OK) {
        fprintf(stderr, "corpus: %s: %s\n", frz_status_str(st), frz_last_error());
        frz_matcher_destroy(m);
        return st == FRZ_ERR_NO_DEVICE || st == FRZ_ERR_CUDA ? 3 : 1;
    }

    frz_match out[N];
    uint64_t n_out = 0;
    st = frz_match_list(m, corpus, out, N, &n_out);
    if (st != FRZ_OK) { fprintf(stderr, "match_list: %s: %s\n", frz_status_str(st), frz_last_error()); return 1; }
    for (uint64_t i = 0; i < n_out; i++)
        printf("Match { score: %u, index: %u, exact: %s }  \"%s\"\n", out[i].score, out[i].index, out[i].exact ? "true" : "false",
               hay[out[i].index]);

    /* matched character offsets of the displayed rows (Matcher::match_list_indices) */
    uint32_t which[1] = {0}, idx[16], cnt[1];
    frz_match mi[1];
    st = frz_match_indices(m, corpus, which, 1, mi, idx, 16, cnt);
    if (st == FRZ_OK && cnt[0] != UINT32_MAX) {
        printf("indices of \"%s\":", hay[0]);
        for (uint32_t k = 0; k < cnt[0]; k++) printf(" %u", idx[k]);
        printf("\n");
    }
    frz_corpus_destroy(corpus);
    frz_matcher_destroy(m);
    return 0;
}
