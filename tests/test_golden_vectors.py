"""Replays tests/golden/reference_kats.json against the oracle, on every emulated backend where the reference requires
backend independence.  Each block cites its source and its provenance: 'reference-held' blocks are literal expectations
of the reference's own tests; the 'restatement-derived' block holds inputs on which the reference asserts only a property
(all backends agree) — the property is what is checked against the reference, the numbers are a regression pin."""
import json
import os

import pytest

import frizbee_b200 as F
from frizbee_b200.types import Config, SortStrategy
from oracle import pyoracle as O

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_kats.json"), encoding="utf-8"))
BACKENDS = [(8, 16), (16, 16), (32, 16), (16, 8), (32, 8), (64, 8)]


@pytest.mark.parametrize("lanes,bits", BACKENDS)
def test_sw_scores(lanes, bits):
    assert G["sw_score_whole_haystack_include_prefix"]["provenance"] == "reference-held"
    for needle, hay, want in G["sw_score_whole_haystack_include_prefix"]["cases"]:
        assert O.sw_score(needle, hay, lanes=lanes, score_bits=bits) == want, (needle, hay)
    # parity.rs pairs: what the REFERENCE asserts is backend independence (every backend == Scalar u16) ...
    derived = G["sw_score_parity_pairs_restatement_derived"]
    assert derived["provenance"] == "restatement-derived"
    for needle, hay, pinned in derived["cases"]:
        got = O.sw_score(needle, hay, lanes=lanes, score_bits=bits)
        assert got == O.sw_score(needle, hay, lanes=8, score_bits=16), (needle, hay)
        # ... the number is the oracle's own (regression pin only, not an independent confirmation)
        assert got == pinned, (needle, hay)
    for needle, hay, want in G["sw_score_unicode"]["cases"]:
        assert O.sw_score_unicode(needle, hay, lanes=lanes, score_bits=bits) == want, (needle, hay)


@pytest.mark.parametrize("lanes", [16, 32, 64])
def test_prefilter_windows(lanes):
    for needle, hay, k, want in G["prefilter_window"]["cases"]:
        assert list(O.prefilter(needle, hay, k, lanes)) == want, (needle, hay, k)
    for needle, hay, k, want in G["prefilter_unicode_window"]["cases"]:
        assert list(O.prefilter_unicode(needle, hay, k, lanes)) == want, (needle, hay, k)


def test_indices():
    for needle, hay, want in G["indices"]["cases"]:
        assert O.sw_indices(needle, hay)[1] == want, (needle, hay)
    for needle, hay, sp, want in G["indices_unicode"]["cases"]:
        assert O.sw_indices(needle, hay, sp, unicode=True)[1] == want, (needle, hay)


def test_match_lists():
    for needle, hs, k, order in G["match_list"]["cases"]:
        assert [m.index for m in O.match_list(needle, hs, Config(max_typos=k))] == order, needle
    for query, hs, want in G["multi_pattern"]["cases"]:
        pats = F.parse_query(query)
        assert [m.index for m in O.match_list(pats, hs, Config(sort=SortStrategy.IndexAsc))] == want, query
