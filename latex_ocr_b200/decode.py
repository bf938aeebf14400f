"""Greedy / beam-search decoding on the B200 step kernels — the decode row of SURVEY.md §8 (a8).

Loop semantics follow the reference's TF decode loop, which is framework independent:
  dynamic_decode (model/components/dynamic_decode.py:38-61): run until every row is finished, at most
  max_length_formula + 2 steps (``finished |= time >= max_iter`` with max_iter = max_length_formula + 1, decoder.py:70);
  greedy (greedy_decoder_cell.py:53-66): id = argmax, finished |= id == END, rows keep decoding after END;
  beam (beam_search_decoder_cell.py:123-250): see include/latex_ocr_b200.h:lo_decoder_beam; ``finalize='reference'``
  reproduces the reference's identity finalize (its back-tracking loop never updates ``parents``, :237), ``'backtrack'``
  returns lineage-consistent hypotheses.
The torch flavour of the reference has no START symbol; the caller supplies ``start_id``.
"""
import ctypes

import torch

from . import _lib
from ._lib import check, stream_ptr


def _n_steps(fin_hist):
    """fin_hist [N, steps(, beam)] int32 -> number of loop iterations the reference would have executed."""
    N = fin_hist.shape[0]
    S = fin_hist.shape[1]
    allf = fin_hist.reshape(N, S, -1).all(dim=2).all(dim=0)          # [S]
    idx = torch.nonzero(allf)
    return int(idx[0]) + 1 if idx.numel() else S


def _prepare(model, img, rows_per_img, max_steps):
    enc = model.encoder.forward_raw(img.to(model.device).float(), need_grad=False)
    N = enc.shape[0]
    R = enc.shape[1] * enc.shape[2]
    enc_flat = enc.view(N, R, enc.shape[3])
    dec = model.decoder
    B = N * rows_per_img
    ws = dec.workspace(B, max_steps, R, need_grad=False)
    dec.sync_shadow()
    for t in range(max_steps):
        ws["bt"][t] = B
    a = dec.fill_args(ws, enc_flat, B, max_steps, R, has_dropout=False)
    a.rows_per_img = rows_per_img
    return a, ws, N


def greedy_decode(model, img, start_id, end_id, max_length_formula=150):
    """Returns token ids [N, steps] (CPU int64), steps as the reference's loop would have run."""
    L = _lib.lib()
    max_steps = max_length_formula + 2
    with torch.no_grad():
        a, ws, N = _prepare(model, img, 1, max_steps)
        dev = model.device
        tokens = torch.zeros(N, max_steps, dtype=torch.int64, device=dev)
        finished = torch.zeros(N, dtype=torch.int32, device=dev)
        hist = torch.zeros(N, max_steps, dtype=torch.int32, device=dev)
        check(L.lo_decoder_greedy_hist(ctypes.byref(a), int(start_id), int(end_id), max_steps, tokens.data_ptr(), finished.data_ptr(),
                                       hist.data_ptr(), stream_ptr()))
        n = _n_steps(hist.cpu())
        return tokens[:, :n].cpu()


def beam_decode(model, img, start_id, end_id, beam_size=5, max_length_formula=150, finalize="reference",
                div_gamma=1, div_prob=0):
    """Returns (ids [N, beam, steps], log_probs [N, beam]) on the CPU; hypothesis 0 is the one the reference scores
    (img2seq.py:210)."""
    if not (div_gamma == 1 or div_prob == 0):
        raise NotImplementedError("diversity penalty (beam_search_decoder_cell.py:258-287) is off in the shipped config and not implemented")
    if finalize not in ("reference", "backtrack"):
        raise NotImplementedError("finalize must be 'reference' or 'backtrack'")
    L = _lib.lib()
    max_steps = max_length_formula + 2
    with torch.no_grad():
        a, ws, N = _prepare(model, img, beam_size, max_steps)
        dev = model.device
        ids = torch.zeros(N, max_steps, beam_size, dtype=torch.int64, device=dev)
        parents = torch.zeros_like(ids)
        hist = torch.zeros(N, max_steps, beam_size, dtype=torch.int32, device=dev)
        logp = torch.zeros(N, beam_size, dtype=torch.float32, device=dev)
        check(L.lo_decoder_beam(ctypes.byref(a), int(start_id), int(end_id), max_steps, ids.data_ptr(), parents.data_ptr(),
                                hist.data_ptr(), logp.data_ptr(), stream_ptr()))
        n = _n_steps(hist.cpu())
        ids, parents = ids[:, :n].cpu(), parents[:, :n].cpu()
        if finalize == "backtrack":
            out = torch.zeros_like(ids)
            cur = torch.arange(beam_size).unsqueeze(0).repeat(N, 1)
            for t in range(n - 1, -1, -1):
                out[:, t] = torch.gather(ids[:, t], 1, cur)
                cur = torch.gather(parents[:, t], 1, cur)
            ids = out
        return ids.permute(0, 2, 1).contiguous(), logp.cpu()      # [N, beam, time] like img2seq.py:239-241


def truncate_end(list_of_ids, id_end):
    """model/evaluation/text.py:95-104."""
    out = []
    for seq in list_of_ids:
        seq = [int(t) for t in seq]
        out.append(seq[:seq.index(id_end)] if id_end in seq else seq)
    return out
