"""Text metrics of model/evaluation/text.py:41-92 restated without nltk / distance (both absent here): the second half
of the north-star metric ("BLEU / exact-match vs ref").  Host-side, pure Python.

* exact_match_score  — text.py:41-57.
* bleu_score         — text.py:60-73 = nltk.translate.bleu_score.corpus_bleu(weights=(.25,)*4), nltk 3.4.5 semantics:
  corpus-level modified n-gram precisions (clipped counts summed over the corpus), brevity penalty on summed lengths,
  no smoothing (a zero precision gives BLEU 0), geometric mean of the four precisions.  Parity unpinned (nltk cannot be
  imported here); checked against hand-computed cases in tests/test_metrics.py.
* edit_distance      — text.py:76-92 = 1 - sum(levenshtein)/sum(max(len)) with unit-cost Levenshtein (distance 0.1.3).
"""
import math
from collections import Counter


def exact_match_score(references, hypotheses):
    exact = sum(1 for r, h in zip(references, hypotheses) if list(r) == list(h))
    return exact / float(max(len(hypotheses), 1))


def _ngrams(seq, n):
    return Counter(tuple(seq[i:i + n]) for i in range(len(seq) - n + 1))


def bleu_score(references, hypotheses, max_n=4):
    num = [0] * max_n
    den = [0] * max_n
    hyp_len = ref_len = 0
    for ref, hyp in zip(references, hypotheses):
        hyp_len += len(hyp)
        ref_len += len(ref)                       # one reference per hypothesis -> closest length is its length
        for n in range(1, max_n + 1):
            h, r = _ngrams(hyp, n), _ngrams(ref, n)
            num[n - 1] += sum(min(c, r[g]) for g, c in h.items())
            den[n - 1] += max(1, sum(h.values()))
    if num[0] == 0:
        return 0.0
    if any(x == 0 for x in num):
        return 0.0                                # nltk: log(0) -> the geometric mean collapses to 0 (with a warning)
    bp = 1.0 if hyp_len > ref_len else (math.exp(1 - ref_len / hyp_len) if hyp_len > 0 else 0.0)
    return bp * math.exp(sum(0.25 * math.log(n / d) for n, d in zip(num, den)))


def levenshtein(a, b):
    a, b = list(a), list(b)
    if len(a) < len(b):
        a, b = b, a
    prev = list(range(len(b) + 1))
    for i, x in enumerate(a, 1):
        cur = [i]
        for j, y in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (x != y)))
        prev = cur
    return prev[-1]


def edit_distance(references, hypotheses):
    d = tot = 0.0
    for r, h in zip(references, hypotheses):
        d += levenshtein(r, h)
        tot += float(max(len(r), len(h)))
    return 1.0 - d / tot if tot else 1.0


def score(references, hypotheses):
    """text.py:33-38 (x100 like score_files)."""
    return {"BLEU-4": bleu_score(references, hypotheses) * 100, "ExactMatchScore": exact_match_score(references, hypotheses) * 100,
            "EditDistance": edit_distance(references, hypotheses) * 100}
