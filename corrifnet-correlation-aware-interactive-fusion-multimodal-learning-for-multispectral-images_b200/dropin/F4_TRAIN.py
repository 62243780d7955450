"""Drop-in for the reference's F4_TRAIN.py: ``train_model`` / ``validate`` with the reference's
argument lists and text-file outputs (F4_TRAIN.py:39-208), driving the B200 kernels.

Single process (``python F2_MAIN.py``): one optimizer step per loader batch, exactly the reference's loop.

Data parallel (``torchrun --nproc-per-node G F2_MAIN.py``; new, the reference is single-device): this module
joins the NCCL process group itself (``ensure_distributed``), binds ``cuda:LOCAL_RANK`` and SHARDS the epoch at
micro-batch granularity: the loader's batches are consumed ``CORRIF_MICROBATCHES_PER_STEP``
at a time (default = world size; BASELINE configs[2] is 8 micro-batches of 8 = global batch 64), rank r runs
micro-batches r, r+G, ... of each group with local gradient accumulation, gradients are averaged by bucketed
all-reduces overlapped with the backward, and every rank applies the same Adam step.  A micro-batch is never
split: train-mode BatchNorm and the inter_attn batch-mixing view couple its samples.  Validation batches are
sharded the same way (after rank 0's BatchNorm running statistics have been broadcast).  Only rank 0 writes the
text files and checkpoints.

Other differences that do not change results: loss and Jaccard stay on the device and are read once per epoch
instead of twice per step (F4_TRAIN.py:64, F5_JACCARD2.py:12); inputs travel through pinned, double-buffered
staging (corrif_b200.staging.PinnedPipeline) one micro-batch ahead of the step that uses them; ``validate``
evaluates the model it is given instead of re-building it from the checkpoint it has just written (same
weights, F4_TRAIN.py:84,180).
"""
import os
import sys

import torch
import torch.distributed as dist
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from corrif_b200.metrics import Jaccard2  # noqa: E402
from corrif_b200.staging import PinnedPipeline  # noqa: E402
from corrif_b200.train import TrainStep, broadcast_module, shard_micro_batches  # noqa: E402

device = torch.device("cuda:%d" % int(os.environ.get("LOCAL_RANK", "0")) if torch.cuda.is_available() else "cpu")


def ensure_distributed():
    """Under torchrun (WORLD_SIZE > 1) join the process group and bind this rank's GPU; no-op otherwise."""
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and not dist.is_initialized():
        if device.type == "cuda":
            torch.cuda.set_device(device)
            dist.init_process_group("nccl", device_id=device)
        else:
            dist.init_process_group("gloo")
    return (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)


def _rank0():
    return int(os.environ.get("RANK", "0")) == 0 and ((not dist.is_initialized()) or dist.get_rank() == 0)


def _reduce_sums(values):
    """values: python/devices scalars -> float64 tensor of their sums over all ranks."""
    t = torch.stack([torch.as_tensor(v, dtype=torch.float64, device=device).reshape(()) for v in values])
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t)
    return t


def _batches(generator):
    """The loader's batches as an indexable plan.  A plain sequential DataLoader (what F2_MAIN builds,
    F2_MAIN.py:90,104-111) is re-sliced by index so a rank only materialises its own micro-batches; any other
    iterable is walked once per epoch and the foreign batches are skipped."""
    dl = generator
    if (isinstance(dl, torch.utils.data.DataLoader) and dl.batch_size is not None
            and isinstance(dl.sampler, torch.utils.data.SequentialSampler)):
        n, bs = len(dl.dataset), dl.batch_size
        spans = [(s, min(s + bs, n)) for s in range(0, n, bs) if not (dl.drop_last and s + bs > n)]
        collate = dl.collate_fn

        def fetch(i):
            lo, hi = spans[i]
            return collate([dl.dataset[j] for j in range(lo, hi)])
        return len(spans), fetch
    cache = list(generator)
    return len(cache), cache.__getitem__


class _Feeder:
    """Pinned, double-buffered H2D of this rank's micro-batches, one ahead of the consumer."""

    def __init__(self, fetch, order):
        self.fetch, self.order, self.k = fetch, order, 0
        self.pipe = PinnedPipeline(device) if device.type == "cuda" else None
        self._push()

    def _push(self):
        if self.k < len(self.order):
            im, ma = self.fetch(self.order[self.k])
            self.k += 1
            if self.pipe is None:
                self._cpu = (im, ma)
            else:
                self.pipe.prefetch([t if t.is_pinned() else t.pin_memory() for t in (im, ma)])

    def next(self):
        if self.pipe is None:
            out = self._cpu
            self._push()
            return out
        im, ma = self.pipe.get()
        self._push()
        return im, ma

    def done(self):
        if self.pipe is not None:
            self.pipe.release()


def train_model(n_epochs, trainloss, validationloss, accuracy, model, scheduler, lrFile, training_generator,
                optim, lim, trainFile, trainaccFile, trainepochFile, validation_generator, valFile,
                valaccFile, pathm, i, modeltype):
    if trainloss != "BCEWithLogitsLoss" or accuracy != "Jaccard":
        raise ValueError("only trainloss='BCEWithLogitsLoss' and accuracy='Jaccard' exist in the reference")
    rank, world = ensure_distributed()
    group = int(os.environ.get("CORRIF_MICROBATCHES_PER_STEP", str(world)))
    broadcast_module(model)
    # CUDA graphs of the model's forward / backward from the second optimizer step on (CORRIF_TRAIN_GRAPHS=0: stream
    # launches).  Capturing runs four extra passes over that step's first micro-batch, i.e. four extra momentum
    # updates of the BatchNorm running statistics early in epoch 0; train-mode results are unaffected.
    step = TrainStep(model, optim, lim=lim, graphs=os.environ.get("CORRIF_TRAIN_GRAPHS", "1") != "0")
    for epoch in range(n_epochs):
        model.train()
        scheduler.step()                                            # before any optimizer step, as :46
        if _rank0():
            print("Epoch:", epoch, "LR:", scheduler.get_last_lr())
            lrFile.write("Epoch:" + " " + str(epoch) + " " + "LR:" + " " + str(scheduler.get_last_lr()) + "\n")
            lrFile.write(str(scheduler.state_dict()) + "\n")
        n_batches, fetch = _batches(training_generator)
        plan = shard_micro_batches(n_batches, group, rank, world)
        feeder = _Feeder(fetch, [j for mine, _ in plan for j in mine])
        loss_sum, jac, pixels, count = 0.0, 0.0, 0, 0
        for mine, total in plan:
            # inputs of a multi-micro-batch step are cloned out of the two staging slots (a step with one local
            # micro-batch, the default, uses the slot directly)
            mbs = []
            for _ in mine:
                im, ma = feeder.next()
                mbs.append((im.clone(), ma.clone()) if len(mine) > 1 else (im, ma))
                if len(mine) > 1:
                    feeder.done()
            out = step(mbs, total_micro_batches=total)
            if len(mine) == 1:
                feeder.done()
            if mbs:
                loss_sum = loss_sum + out["loss_sum"]
                jac = jac + out["jaccard_sum"].reshape(())
                pixels += out["pixels"]
                count += len(mbs)
        sums = _reduce_sums([loss_sum, count, jac, pixels])         # the only host reads of the epoch
        training_loss, train_jac = (sums[0] / sums[1]).item(), (sums[2] / sums[3]).item()   # :74-77
        if _rank0():
            trainFile.write(str(training_loss) + "\n")
            trainaccFile.write(str(train_jac) + "\n")
            trainepochFile.write(str(epoch) + "\n")
            print("Training Jaccard:", train_jac, " (epoch:", epoch, ")")
            lrFile.write("Training loss:" + str(training_loss) + "\n")
            lrFile.write("Training accuracy:" + str(train_jac) + "\n")
            torch.save(model.state_dict(), os.path.join(pathm, "iremmodel{}.pt".format(i)))      # :84
        validate(validationloss, accuracy, validation_generator, valFile, valaccFile, lim, lrFile, pathm, i,
                 modeltype, model=model)
    if _rank0():
        torch.save(model.state_dict(), os.path.join(pathm, "Finaliremmodel{}.pt".format(i)))    # :86


def evaluate(model, generator, lim):
    """The eval loop shared by ``validate`` (F4_TRAIN.py:181-199) and ``test_model`` (F7_TEST2.py:131-176):
    mean BCE-with-logits loss over batches and pixel-weighted Jaccard2 of channel 0, batches sharded over ranks."""
    rank, world = ensure_distributed()
    broadcast_module(model, buffers_only=True)      # every rank evaluates with rank 0's running statistics
    was_training = model.training
    model.eval()
    n_batches, fetch = _batches(generator)
    mine = list(range(rank, n_batches, world))
    feeder = _Feeder(fetch, mine)
    loss_sum, jac, pixels = 0.0, 0.0, 0
    with torch.no_grad():
        for _ in mine:
            images, masks = feeder.next()
            outputs = model(images)
            loss_sum = loss_sum + F.binary_cross_entropy_with_logits(outputs, masks)
            load = len(masks) * lim * lim
            jac = jac + (Jaccard2(masks[:, 0].reshape(load, 1), outputs[:, 0].reshape(load, 1)) * load).reshape(())
            pixels += load
            feeder.done()
    model.train(was_training)
    sums = _reduce_sums([loss_sum, len(mine), jac, pixels])
    return (sums[0] / sums[1]).item(), (sums[2] / sums[3]).item()


def validate(validationloss, accuracy, validation_generator, valFile, valaccFile, lim, lrFile, pathm, i,
             modeltype, model=None):
    if model is None:       # reference behaviour: rebuild from the checkpoint of this epoch (:96-180)
        from mmvit4 import MMVit4
        model = MMVit4(num_cls=1).to(device)
        model.load_state_dict(torch.load(os.path.join(pathm, "iremmodel{}.pt".format(i)), map_location=device))
    val_loss, dni = evaluate(model, validation_generator, lim)
    if _rank0():
        valFile.write(str(val_loss) + "\n")
        valaccFile.write(str(dni) + "\n")
        print("Validation Jaccard:", dni)
        lrFile.write("Validation loss:" + str(val_loss) + "\n")
        lrFile.write("Validation accuracy:" + str(dni) + "\n")
    return val_loss, dni
