"""CPU: the extension's CPU definition (oracle/ref_ext.py) uses torch's own LSTM arithmetic — checked here against torch.nn.LSTM /
nn.LSTMCell modules loaded with the same parameters."""
import torch

from oracle import ref_ext as rx


def test_row_encoder_equals_nn_lstm_bidirectional():
    prow, p2 = rx.init_params_ext(seed=4, channels=16, hidden=8, D=16)
    feat = torch.randn(2, 3, 5, 16, generator=torch.Generator().manual_seed(1))
    got = rx.row_encoder_forward(prow, feat)
    lstm = torch.nn.LSTM(16, 8, bidirectional=True, batch_first=True)
    lstm.load_state_dict({k.split(".", 1)[1]: v for k, v in prow.items()})
    want, _ = lstm(feat.reshape(6, 5, 16))
    assert got.shape == (2, 3, 5, 16)
    assert (got.reshape(6, 5, 16) - want).abs().max().item() < 1e-6
    cell = torch.nn.LSTMCell(16, 16)
    cell.load_state_dict({k.split(".", 1)[1]: v for k, v in p2.items()})
    x = torch.randn(3, 4, 16)
    got2 = rx.lstm_seq(x, p2["cell.weight_ih"], p2["cell.weight_hh"], p2["cell.bias_ih"], p2["cell.bias_hh"])
    h = c = torch.zeros(3, 16)
    for t in range(4):
        h, c = cell(x[:, t], (h, c))
        assert (got2[:, t] - h).abs().max().item() < 1e-6
