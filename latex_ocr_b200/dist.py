"""Data-parallel gradient synchronisation (new vs the reference, which is single-device): one process per
GPU, NCCL all-reduce (sum) of the two flat gradient buckets only — decoder bucket as soon as the decoder
backward ends (it overlaps the whole encoder backward), encoder bucket after it; Adam applies 1/world.
With equal per-rank batch and padded length the global loss gradient is the plain mean of rank gradients
(img2seq_torch.py:155,159 are means), so no other collective is needed.  Works with gloo on CPU tensors
for the world_size-2 tests."""
import torch
import torch.distributed as dist


class GradSync:
    def __init__(self, group=None, use_side_stream=True):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.group = group
        self.world_size = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self._handles = []
        self._side = None
        if use_side_stream and torch.cuda.is_available() and dist.get_backend(group) == "nccl":
            self._side = torch.cuda.Stream()

    def reduce_async(self, flat_grad):
        if self.world_size == 1:
            return
        if self._side is not None and not torch.cuda.is_current_stream_capturing():
            self._side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._side):
                dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=self.group)
            self._handles.append("side")
        else:
            self._handles.append(dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def layer_hook(self, encoder, min_bucket=262144):
        """Returns ``on_layer_grad(idx)`` for EncoderCNN.backward_raw: layers complete from the last conv to the first; their
        flat-store ranges are adjacent (descending), so consecutive small layers are merged until a bucket reaches
        ``min_bucket`` elements, and whatever is left goes out with the first layer."""
        first = encoder.layers[0][0]
        pending = []          # [lo, hi) of the not-yet-reduced tail, grows downwards

        def hook(idx):
            off, n = encoder.grad_range(idx)
            if pending and pending[0] != off + n:          # not adjacent (padding between entries is included by range math)
                lo, hi = pending[0], pending[1]
                self.reduce_async(encoder.store.grad[lo:hi])
                del pending[:]
            if not pending:
                pending.extend([off, off + n])
            else:
                pending[0] = off
            if pending[1] - pending[0] >= min_bucket or idx == first:
                self.reduce_async(encoder.store.grad[pending[0]:pending[1]])
                del pending[:]
        return hook

    def wait(self):
        for h in self._handles:
            if h == "side":
                torch.cuda.current_stream().wait_stream(self._side)
            elif h is not None:
                h.wait()
        self._handles = []


def attach(model, group=None):
    """Broadcast rank-0 parameters and enable gradient all-reduce in ``model._step_body``."""
    sync = GradSync(group)
    for store in (model.encoder.store, model.decoder.store):
        dist.broadcast(store.master, src=0, group=group)
        store.sync_shadow()
    model.dist = sync
    return sync
