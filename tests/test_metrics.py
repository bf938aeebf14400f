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
    # clipped counts: hyp "the the the the" vs ref with two "the" -> p1 = 2/4, no bigram match -> nltk 3.4.5's method0 puts
    # sys.float_info.min in place of the three zero precisions: (1/2)^(1/4) * (2.2e-308)^(3/4) ~ 1e-231, not exactly 0
    import sys
    b = metrics.bleu_score(["the cat the dog".split()], ["the the the the".split()])
    assert abs(b / (0.5 ** 0.25 * sys.float_info.min ** 0.75) - 1) < 1e-9 and 0 < b < 1e-200
    assert metrics.bleu_score(["a b".split()], ["c d".split()]) == 0          # no unigram match: exactly 0
    # corpus level: counts are pooled over sentences before the geometric mean
    refs = ["a b c d e".split(), "f g h i j".split()]
    hyps = ["a b c d e".split(), "f g h x j".split()]
    p = [9 / 10, 6 / 8, 4 / 6, 2 / 4]
    want = math.exp(sum(0.25 * math.log(x) for x in p))
    assert abs(metrics.bleu_score(refs, hyps) - want) < 1e-12
    s = metrics.score(refs, hyps)
    assert set(s) == {"BLEU-4", "ExactMatchScore", "EditDistance"} and s["ExactMatchScore"] == 50.0


def test_write_answers_and_score_files(tmp_path):
    rev = {0: "x", 1: "^", 2: "2", 3: "_END"}
    refs = [[0, 1, 2], [0]]
    hyps = [[[0, 1, 2, 3, 0], [2, 3]], [[0, 1, 3], [0, 3, 1]]]
    d = str(tmp_path) + "/answers/"
    files = metrics.write_answers(refs, hyps, rev, d, 3)
    assert files == [d + "ref.txt", d + "hyp_0.txt", d + "hyp_1.txt"]
    assert open(files[0]).read() == "x ^ 2\nx\n" and open(files[1]).read() == "x ^ 2\n2\n" and open(files[2]).read() == "x ^\nx\n"
    s = metrics.score_files(files[0], files[1])
    assert s["ExactMatchScore"] == 50.0 and abs(s["EditDistance"] - 75.0) < 1e-9
    assert metrics.truncate_end([5, 6, 3, 7], 3) == [5, 6] and metrics.truncate_end([5, 6], 3) == [5, 6]
