"""TEST INFRASTRUCTURE ONLY — generates tests/golden/host_pipeline.json and tests/golden/lr_schedule.json by running the
UNMODIFIED reference host-side helpers (model/utils/{image,general,text,lr_schedule,data_generator}.py,
model/evaluation/text.py) in the build container.  Run: ``python -m oracle.make_golden_host``.  Tests only READ the files.

Import shims (arithmetic-neutral): ``nltk`` / ``distance`` stub modules (only bleu / edit distance use them, not exercised
here), ``scipy.misc.imread`` (removed from scipy) replaced by a PIL reader returning the same uint8 arrays.
"""
import importlib
import json
import os
import sys
import tempfile
import types

import numpy as np

from oracle import ref_shim

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

LR_CASES = [
    dict(lr_init=1e-3, lr_min=1e-5, start_decay=6, end_decay=13, lr_warm=1e-4, end_warm=2),          # train.py:49-56 shape
    dict(lr_init=1e-3, lr_min=1e-4, start_decay=0, decay_rate=0.5, early_stopping=3),
    dict(lr_init=2e-3, lr_min=1e-4, start_decay=3, end_decay=9),
    dict(lr_init=1e-3),
]
LR_SCORES = [None, 1.0, 0.5, None, 0.5, 0.7, 0.6, None, 0.6, 0.5, 0.4, None, None, 0.3, None, None]

# synthetic dataset of the bucketing case: (height, width) per image in listing order, bucket_size 3
BUCKET_SHAPES = [(8, 16), (8, 32), (8, 16), (8, 16), (8, 64), (8, 16), (8, 32), (8, 16), (8, 32), (8, 32), (8, 16), (8, 64),
                 (8, 16), (8, 16), (8, 16), (8, 32), (8, 16)]
BUCKET_SIZE = 3


def lr_trace(cls, kw):
    s = cls(**kw)
    out = []
    for i, sc in enumerate(LR_SCORES):
        s.update(batch_no=i)
        if sc is not None:
            s.update(score=sc)
        out.append((s.lr, bool(s.stop_training)))
    return out


def synthetic_rgb(seed=7, h=9, w=13):
    return np.random.RandomState(seed).randint(0, 256, size=(h, w, 3)).astype(np.uint8)


def write_dataset(root, shapes):
    """PNG images of the given shapes + formulas + matching file; returns the three paths."""
    from PIL import Image
    os.makedirs(os.path.join(root, "images"), exist_ok=True)
    rs = np.random.RandomState(11)
    with open(os.path.join(root, "formulas.txt"), "w") as ff, open(os.path.join(root, "matching.txt"), "w") as fm:
        for i, (h, w) in enumerate(shapes):
            Image.fromarray(rs.randint(0, 256, size=(h, w, 3)).astype(np.uint8)).save(os.path.join(root, "images", "%d.png" % i))
            ff.write(" ".join("t%d" % int(x) for x in rs.randint(0, 9, size=int(rs.randint(1, 6)))) + "\n")
            fm.write("%d.png %d\n" % (i, i))
    return os.path.join(root, "formulas.txt"), os.path.join(root, "images") + "/", os.path.join(root, "matching.txt")


def load_reference_utils():
    for name in ("tensorflow", "nltk", "distance", "h5py"):
        ref_shim._stub(name)
    if "scipy.misc" not in sys.modules or not hasattr(sys.modules["scipy.misc"], "imread"):
        from PIL import Image
        import scipy
        m = types.ModuleType("scipy.misc")
        m.__spec__ = importlib.machinery.ModuleSpec("scipy.misc", loader=None)
        m.imread = lambda p: np.asarray(Image.open(p))
        sys.modules["scipy.misc"] = m
        scipy.misc = m
    if ref_shim.REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, ref_shim.REFERENCE_ROOT)
    mods = {}
    for k, name in (("image", "model.utils.image"), ("general", "model.utils.general"), ("text", "model.utils.text"),
                    ("lr", "model.utils.lr_schedule"), ("dg", "model.utils.data_generator"), ("ev", "model.evaluation.text")):
        mods[k] = importlib.import_module(name)
    return mods


def compute(mods):
    rec = {}
    rgb = synthetic_rgb()
    rec["greyscale"] = mods["image"].greyscale(rgb)[:, :, 0].tolist()
    imgs = [np.full((3, 4, 1), 7, np.uint8), np.full((5, 2, 1), 9, np.uint8)]
    rec["pad_batch_images"] = mods["image"].pad_batch_images(imgs)[:, :, :, 0].tolist()
    f, l = mods["text"].pad_batch_formulas([[4, 5, 6], [7]], 1, 2)
    rec["pad_batch_formulas"] = [np.asarray(f).tolist(), np.asarray(l).tolist()]
    rec["minibatches"] = [[list(x), list(y)] for x, y in mods["general"].minibatches(((i, -i) for i in range(7)), 3)]
    with tempfile.TemporaryDirectory() as d:
        pf, di, pm = write_dataset(d, BUCKET_SHAPES)
        gen = mods["dg"].DataGenerator(pf, di, pm, bucket=True, bucket_size=BUCKET_SIZE, img_prepro=mods["image"].greyscale)
        rec["bucket_order"] = [[p, int(i)] for p, i in gen._data_generator]
        rec["bucket_batches_shapes"] = [[list(im.shape) for im in xs] for xs, _ in mods["general"].minibatches(gen, BUCKET_SIZE)]
        rec["bucket_first_formula"] = next(iter(gen))[1]
        gen2 = mods["dg"].DataGenerator(pf, di, pm, max_len=3, max_iter=5)
        rec["maxlen_maxiter_formulas"] = [fm for _, fm in gen2]
        rev = {0: "x", 1: "^", 2: "2", 3: "_END"}
        files = mods["ev"].write_answers([[0, 1, 2], [0]], [[[0, 1, 2, 3, 0], [2, 3]], [[0, 1, 3], [0, 3, 1]]], rev, d + "/ans/", 3)
        rec["write_answers"] = [[os.path.basename(p), open(p).read()] for p in files]
    rec["truncate_end"] = mods["ev"].truncate_end([5, 6, 3, 7], 3)
    rec["exact_match"] = mods["ev"].exact_match_score([["a"], ["b", "c"], ["d"]], [["a"], ["b"], ["d"]])
    return rec


def main():
    if not ref_shim.reference_available():
        sys.exit("reference tree not available; golden files can only be regenerated in the build container")
    mods = load_reference_utils()
    with open(os.path.join(OUT, "host_pipeline.json"), "w") as f:
        json.dump(compute(mods), f)
    with open(os.path.join(OUT, "lr_schedule.json"), "w") as f:
        json.dump([lr_trace(mods["lr"].LRSchedule, kw) for kw in LR_CASES], f)
    print("wrote host_pipeline.json, lr_schedule.json")


if __name__ == "__main__":
    main()
