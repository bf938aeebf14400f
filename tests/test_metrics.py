"""CPU: text metrics (model/evaluation/text.py:41-92 restated) on hand-computed cases."""
import math

from latex_ocr_b200 import metrics


def test_exact_match_and_edit_distance():
    refs = [["a", "b", "c"], ["x"], ["p", "q"]]
    hyps = [["a", "b", "c"], ["y"], ["p"]]
    assert metrics.exact_match_score(refs, hyps) == 1 / 3
    # levenshtein: 0 + 1 + 1 over max lens 3 + 1 + 2
    assert abs(metrics.edit_distance(refs, hyps) - (1 - 2 / 6)) < 1e-12
    assert metrics.levenshtein("kitten", "sitting") == 3
    assert metrics.exact_match_score([], []) == 0.0


def test_bleu4_corpus_level():
    ref = "the cat sat on the mat today".split()
    assert abs(metrics.bleu_score([ref], [ref]) - 1.0) < 1e-12
    hyp = "the cat sat on the mat".split()            # shorter: precisions all 1, brevity penalty exp(1 - 7/6)
    assert abs(metrics.bleu_score([ref], [hyp]) - math.exp(1 - 7 / 6)) < 1e-12
    # clipped counts: hyp "the the the the" vs ref with two "the" -> p1 = 2/4, no bigram match -> 0 (no smoothing)
    assert metrics.bleu_score(["the cat the dog".split()], ["the the the the".split()]) == 0.0
    # corpus level: counts are pooled over sentences before the geometric mean
    refs = ["a b c d e".split(), "f g h i j".split()]
    hyps = ["a b c d e".split(), "f g h x j".split()]
    p = [9 / 10, 6 / 8, 4 / 6, 2 / 4]
    want = math.exp(sum(0.25 * math.log(x) for x in p))
    assert abs(metrics.bleu_score(refs, hyps) - want) < 1e-12
    s = metrics.score(refs, hyps)
    assert set(s) == {"BLEU-4", "ExactMatchScore", "EditDistance"} and s["ExactMatchScore"] == 50.0
