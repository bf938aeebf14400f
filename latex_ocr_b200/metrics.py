"""Text metrics and answer files of model/evaluation/text.py restated without nltk / distance (both absent here): the second
half of the north-star metric ("BLEU / exact-match vs ref").  Host-side, pure Python.

* exact_match_score  — text.py:41-57.
* bleu_score         — text.py:60-73 = nltk.translate.bleu_score.corpus_bleu(weights=(.25,)*4) with nltk 3.4.5 semantics
  (requirements.txt:4): corpus-level modified n-gram precisions (clipped counts summed over the corpus, denominator
  max(1, #n-grams)), brevity penalty on the summed lengths, returns 0 when there is no unigram match, and the default
  ``SmoothingFunction().method0``: a higher-order precision with a ZERO numerator is replaced by ``sys.float_info.min``
  (nltk warns), so such a corpus scores ~1e-77 * BP rather than exactly 0.  The log-sum uses math.fsum like nltk.
  Parity unpinned (nltk cannot be imported here); checked on hand-computed cases in tests/test_metrics.py.
* edit_distance      — text.py:76-92 = 1 - sum(levenshtein)/sum(max(len)) with unit-cost Levenshtein (Distance 0.1.3).
* truncate_end / write_answers / score_files / load_formulas — text.py:95-145, :12-38, utils/text.py:167-174.
"""
import math
import os
import sys
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
        return 0                                  # nltk: "if p_numerators[1] == 0: return 0"
    bp = 1.0 if hyp_len > ref_len else (math.exp(1 - ref_len / hyp_len) if hyp_len > 0 else 0.0)
    p = [(n / d) if n != 0 else sys.float_info.min for n, d in zip(num, den)]     # SmoothingFunction().method0
    return bp * math.exp(math.fsum(0.25 * math.log(x) for x in p))


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


def truncate_end(list_of_ids, id_end):
    """text.py:95-104: drop everything from the first END token on."""
    out = []
    for idx in list_of_ids:
        if idx == id_end:
            break
        out.append(idx)
    return out


def write_answers(references, hypotheses, rev_vocab, dir_name, id_end):
    """text.py:107-145: one ``ref.txt`` and one ``hyp_<i>.txt`` per hypothesis rank, one formula per line, tokens joined by a
    blank, each sequence truncated at its first END.  ``dir_name`` is used as a PREFIX exactly like the reference
    (``dir_name + "ref.txt"``), so it should end with a path separator.  Returns the list of file names."""
    def ids_to_str(ids):
        return " ".join(rev_vocab[int(idx)] for idx in truncate_end(ids, id_end))

    def write_file(file_name, list_of_list):
        with open(file_name, "w") as f:
            for l in list_of_list:
                f.write(ids_to_str(l) + "\n")

    if dir_name and not os.path.exists(dir_name):
        os.makedirs(dir_name)                                   # init_dir (utils/general.py)
    file_names = [dir_name + "ref.txt"]
    write_file(dir_name + "ref.txt", references)
    for i in range(len(hypotheses)):
        assert len(references) == len(hypotheses[i])
        write_file(dir_name + "hyp_{}.txt".format(i), hypotheses[i])
        file_names.append(dir_name + "hyp_{}.txt".format(i))
    return file_names


def load_formulas(filename):
    """utils/text.py:167-174: dict[line index] = stripped line."""
    formulas = dict()
    with open(filename) as f:
        for idx, line in enumerate(f):
            formulas[idx] = line.strip()
    return formulas


def score_files(path_ref, path_hyp):
    """text.py:12-38: reload two answer files, split on blanks, score (x100)."""
    formulas_ref = load_formulas(path_ref)
    formulas_hyp = load_formulas(path_hyp)
    assert len(formulas_ref) == len(formulas_hyp)
    refs = [ref.split(" ") for _, ref in formulas_ref.items()]
    hyps = [hyp.split(" ") for _, hyp in formulas_hyp.items()]
    return score(refs, hyps)
