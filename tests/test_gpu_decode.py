"""-m gpu: greedy and beam decoding through the C ABI vs the CPU restatement of the reference's decode loop
(oracle/ref_decode.py; parity unpinned — the TF graph cannot run here).  fp32 mode: token ids must match exactly."""
import pytest
import torch

from util import build_model

pytestmark = pytest.mark.gpu


def _setup(V=30, seed=4, N=3, H=32, W=80):
    from oracle import ref_model as rm
    pe, pd = rm.init_params(V, seed=seed)
    # make decoding non-degenerate: larger output layer so the argmax moves and END appears
    g = torch.Generator().manual_seed(seed)
    pd["fc.weight"] = (torch.rand(V, 512, generator=g) * 2 - 1) * 0.5
    pd["embedding.weight"] = (torch.rand(V, 512, generator=g) * 2 - 1) * 1.0
    img, _ = rm.synthetic_batch(N, H, W, V, 3, 5, seed=seed + 1)
    m = build_model(V, pe, pd, "fp32")
    enc = rm.encoder_forward(pe, img).reshape(N, -1, 512)
    return rm, pd, m, img, enc


def test_greedy_tokens_match_oracle():
    from latex_ocr_b200 import decode
    from oracle import ref_decode as rd
    V = 30
    rm, pd, m, img, enc = _setup(V)
    for end_id, L in ((V - 1, 6), (7, 12)):
        want = rd.greedy_decode(pd, enc, start_id=V - 2, end_id=end_id, max_iter=L + 1)
        got = decode.greedy_decode(m, img, start_id=V - 2, end_id=end_id, max_length_formula=L)
        assert got.shape == want.shape, (got.shape, want.shape)
        assert torch.equal(got, want)


@pytest.mark.parametrize("beam", [1, 3, 5])
def test_beam_tokens_match_oracle(beam):
    from latex_ocr_b200 import decode
    from oracle import ref_decode as rd
    V = 30
    rm, pd, m, img, enc = _setup(V, seed=6)
    for fin in ("reference", "backtrack"):
        want, wlp = rd.beam_decode(pd, enc, start_id=V - 2, end_id=5, beam=beam, max_iter=9, finalize=fin)
        got, glp = decode.beam_decode(m, img, start_id=V - 2, end_id=5, beam_size=beam, max_length_formula=8, finalize=fin)
        assert got.shape == want.permute(0, 2, 1).shape
        assert torch.equal(got, want.permute(0, 2, 1))
        assert (glp - wlp).abs().max().item() < 1e-3 * max(1.0, wlp.abs().max().item())


@pytest.mark.parametrize("gamma,prob", [(0.5, 1.0), (0.8, 0.6), (1.7, 0.5)])
def test_beam_diversity_penalty_matches_oracle(gamma, prob):
    """add_div_penalty (beam_search_decoder_cell.py:258-287) with the Bernoulli draws injected on both sides."""
    from latex_ocr_b200 import decode
    from oracle import ref_decode as rd
    V, beam, L = 30, 3, 8
    rm, pd, m, img, enc = _setup(V, seed=6)
    N = img.shape[0]
    u = torch.rand(L + 2, N * beam, V, generator=torch.Generator().manual_seed(77))
    want, wlp = rd.beam_decode(pd, enc, start_id=V - 2, end_id=5, beam=beam, max_iter=L + 1, div_gamma=gamma, div_prob=prob, div_u=u)
    base, _ = rd.beam_decode(pd, enc, start_id=V - 2, end_id=5, beam=beam, max_iter=L + 1)
    assert want.shape != base.shape or not torch.equal(want, base)          # the penalty changes the search at these settings
    got, glp = decode.beam_decode(m, img, start_id=V - 2, end_id=5, beam_size=beam, max_length_formula=L, div_gamma=gamma,
                                  div_prob=prob, div_u=u)
    assert got.shape == want.permute(0, 2, 1).shape
    assert torch.equal(got, want.permute(0, 2, 1))
    assert (glp - wlp).abs().max().item() < 1e-3 * max(1.0, wlp.abs().max().item())
    # in-kernel Philox draws: runs, is reproducible for a seed, and differs from the penalty-free search
    a1, _ = decode.beam_decode(m, img, V - 2, 5, beam, L, div_gamma=0.5, div_prob=0.5, div_seed=9)
    a2, _ = decode.beam_decode(m, img, V - 2, 5, beam, L, div_gamma=0.5, div_prob=0.5, div_seed=9)
    assert torch.equal(a1, a2)


def test_greedy_attention_export():
    """attention_mechanism.py:96-121: the per-step attention weights of a greedy decode, and their visualize_attention.py map."""
    from latex_ocr_b200 import decode
    from oracle import ref_decode as rd
    V = 30
    rm, pd, m, img, enc = _setup(V)
    ids, att = decode.greedy_decode(m, img, start_id=V - 2, end_id=V - 1, max_length_formula=6, return_attention=True)
    assert att.shape == (img.shape[0], ids.shape[1], enc.shape[1])
    assert (att.sum(dim=2) - 1).abs().max().item() < 1e-4
    # step 0 of the oracle: attention over enc with the initial hidden state
    h0, c0 = rm.init_hidden_state(pd, enc)
    _, a0 = rm.attention_forward(pd, enc, h0)
    assert (att[:, 0] - a0).abs().max().item() < 1e-5
    hh, ww = m.encoder.out_hw(img.shape[2], img.shape[3])
    maps = decode.attention_maps(att, hh, ww)
    assert maps.shape == (img.shape[0], ids.shape[1], hh, ww)
    r = 5
    assert abs(maps[0, 0, r // ww, r % ww].item() - (1 - att[0, 0, r].item()) * 255.0) < 1e-3


def test_predict_batch_surface():
    rm, pd, m, img, enc = _setup(30, seed=8)
    out = m.predict_batch(img, decoding="greedy")
    assert len(out) == 1 and len(out[0]) == img.shape[0]
    out = m.predict_batch(img, decoding="beam_search", beam_size=2)
    assert len(out) == 2 and all(len(h) == img.shape[0] for h in out)
    assert all(29 not in seq for seq in out[0])        # truncated at END


def test_train_epoch_and_evaluate_surface(tmp_path):
    """_run_train_epoch / evaluate with the reference's data conventions (lists of HWC uint8 arrays + token-id lists); answer files,
    save-on-best epoch checkpoints and resume-from-latest (base.py:33-69, :95-140)."""
    import numpy as np
    from latex_ocr_b200.data import SimpleVocab
    from latex_ocr_b200.img2seq import Img2SeqModel
    from latex_ocr_b200.lr_schedule import LRSchedule
    from util import Cfg
    rng = np.random.RandomState(0)
    V = 30
    vocab = SimpleVocab(V)
    data = [(rng.randint(0, 256, (32, 64 + 16 * (i % 2), 1)).astype(np.uint8), list(rng.randint(0, V - 3, 3 + i % 3))) for i in range(6)]
    # same-shape batches like the reference's bucketing (data_generator.py:84-122)
    data.sort(key=lambda d: d[0].shape)
    cfg = Cfg(batch_size=3, n_epochs=2, max_length_formula=6, decoding="greedy", dir_answers=str(tmp_path) + "/answers/")
    m = Img2SeqModel(cfg, dir_output=str(tmp_path), vocab=vocab, device="cuda", precision="fp32").build_train()
    sched = LRSchedule(lr_init=1e-3, apply_to=m)
    score = m.train(cfg, data, data, sched)
    assert score < 0 and np.isfinite(score)                      # negated perplexity
    s = m.last_epoch_stats
    assert {"BLEU-4", "ExactMatchScore", "EditDistance", "perplexity", "images_per_s"} <= set(s)
    # write_prediction wrote the answer files the scores were computed from (img2seq.py:215-254)
    import os
    assert sorted(os.listdir(str(tmp_path) + "/answers/")) == ["hyp_0.txt", "ref.txt"]
    assert len(open(str(tmp_path) + "/answers/ref.txt").read().splitlines()) == len(data)
    files, perp = m.write_prediction(cfg, data)
    assert perp < 0 and files[0].endswith("ref.txt")
    # train() saved a checkpoint on every new best score, keeping one (max_to_keep=1); a fresh model resumes from it
    ck = os.listdir(str(tmp_path) + "/model_weights")
    assert len(ck) == 1 and ck[0].startswith("model.cpkt-")
    m.save_session(7)                                            # current state; the older checkpoint is dropped
    ck = os.listdir(str(tmp_path) + "/model_weights")
    assert ck == ["model.cpkt-7"]
    m2 = Img2SeqModel(cfg, dir_output=str(tmp_path), vocab=vocab, device="cuda", precision="fp32").build_train()
    ep = m2.restore_latest()
    assert ep == int(ck[0].split("-")[1]) and m2.startepoch == ep
    assert torch.equal(m2.decoder.store.master, m.decoder.store.master) and torch.equal(m2.encoder.store.m, m.encoder.store.m)
    assert m2.predict_batch([d[0] for d in data[:3]]) == m.predict_batch([d[0] for d in data[:3]])
    # frozen encoder: fine_tune(False) leaves every conv weight untouched by the fused Adam
    before = m.encoder.store.master.clone()
    m.encoder.fine_tune(False)
    m._run_train_epoch(cfg, data, None, 0, None)
    assert torch.equal(before, m.encoder.store.master)


def test_bf16_greedy_token_match_rate_at_cfg5_widths():
    """BASELINE.json configs[4] shapes (height 64, widths 64..1024): greedy tokens of the bf16 / tcgen05 path against the fp32 CPU
    oracle.  bf16 rounding can legitimately flip an argmax whose top-2 margin is at rounding level, so the comparison is made
    where it is well defined: teacher-forced on the oracle's own tokens, on the steps whose fp32 top-2 logit margin exceeds
    5e-2 (logits are O(1)).  Stated bar: >= 0.99 of those steps agree; the free-running exact-match rate is reported."""
    from latex_ocr_b200 import decode
    from oracle import ref_decode as rd
    from oracle import ref_model as rm
    V, S = 500, 20
    pe, pd = rm.init_params(V, seed=21)
    g = torch.Generator().manual_seed(21)
    pd["fc.weight"] = (torch.rand(V, 512, generator=g) * 2 - 1) * 0.5         # non-degenerate argmax (as in _setup)
    pd["embedding.weight"] = (torch.rand(V, 512, generator=g) * 2 - 1) * 1.0
    m = build_model(V, pe, pd, "bf16", impl="tc")
    m.train_mode(False)
    agree = considered = exact = nseq = free_match = free_total = 0
    for W in (64, 128, 256, 512, 1024):
        img, _ = rm.synthetic_batch(2, 64, W, V, 3, 5, seed=100 + W)
        enc = rm.encoder_forward(pe, img).reshape(2, -1, 512)
        want = rd.greedy_decode(pd, enc, start_id=V - 2, end_id=V - 1, max_iter=S - 1)          # [2, S] fp32 oracle tokens
        n = want.shape[1]
        caps = torch.cat([torch.full((2, 1), V - 2, dtype=torch.long), want], dim=1)            # START + tokens
        lens = torch.full((2, 1), n + 1, dtype=torch.long)
        ref_logits, _, _, _, _ = rm.decoder_forward(pd, enc, caps, lens)                          # teacher-forced fp32 logits
        top2 = ref_logits.topk(2, dim=-1).values
        clear = (top2[..., 0] - top2[..., 1]) > 5e-2
        assert torch.equal(ref_logits.argmax(-1)[clear], want[clear])                             # the oracle agrees with itself
        enc_gpu = m.encoder(img.cuda())
        got_logits, _, _, _, _ = m.decoder(enc_gpu, caps.cuda(), lens)
        hit = got_logits.argmax(-1).cpu() == want
        agree += int((hit & clear).sum())
        considered += int(clear.sum())
        free = decode.greedy_decode(m, img, V - 2, V - 1, S - 2)[:, :n]
        eq = free == want[:, :free.shape[1]]
        free_match += int(eq.sum())
        free_total += eq.numel()
        exact += int(eq.all(dim=1).sum())
        nseq += 2
    rate = agree / max(considered, 1)
    print("cfg5 bf16 greedy: teacher-forced match on clear-margin steps %d/%d = %.4f ; free-running token match %.4f, exact sequences %d/%d"
          % (agree, considered, rate, free_match / free_total, exact, nseq))
    assert considered >= 0.5 * 10 * S and rate >= 0.99
