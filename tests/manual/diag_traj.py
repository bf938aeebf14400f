"""GPU diagnostic: which tensors diverge from the CPU oracle over 3 Adam steps (tiny_eval)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import torch
from util import build_model, load_golden
from oracle import ref_model as rm
rec = load_golden(sys.argv[1] if len(sys.argv) > 1 else "tiny_eval"); c = rec["case"]
pe, pd = rm.init_params(c["V"], seed=c["pseed"])
img, formula = rm.synthetic_batch(c["B"], c["H"], c["W"], c["V"], c["tmin"], c["tmax"], seed=c["dseed"])
m = build_model(c["V"], pe, pd, "fp32", positional=c["positional"])
B, T = c["B"], formula.shape[1] - 1
st = {}
for step in range(3):
    neg, ge, gd, _ = rm.train_step(pe, pd, img, formula, st, positional=c["positional"])
    loss = m._step_body(img.cuda(), formula.cuda(), [T] * B, None); torch.cuda.synchronize()
    print("step", step, "oracle", neg, "gpu", -loss[0].item())
    rows = []
    for mod, g, p in ((m.decoder, gd, pd), (m.encoder, ge, pe)):
        for k, par in mod.named_parameters():
            gg = par.grad.detach().float().cpu()
            rows.append((k, (gg - g[k]).abs().max().item(), g[k].abs().max().item(), (par.detach().cpu() - p[k]).abs().max().item()))
    rows.sort(key=lambda r: -r[3])
    for r in rows[:8]:
        print("   %-34s grad abs err %.3e (max |g| %.3e)  param abs diff %.3e" % r)
