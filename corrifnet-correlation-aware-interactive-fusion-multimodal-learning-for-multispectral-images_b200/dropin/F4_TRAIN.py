"""Drop-in for the reference's F4_TRAIN.py: ``train_model`` / ``validate`` with the reference's
argument lists and text-file outputs (F4_TRAIN.py:39-208), driving the B200 fusion kernels.

Differences that do not change results: the device is ``cuda:LOCAL_RANK`` (one process per GPU under
torchrun) instead of the hard-wired ``cuda:0``; loss and Jaccard stay on the device and are read once
per epoch instead of twice per step; ``validate`` evaluates the model it is given instead of
re-building it from the checkpoint it has just written (same weights, F4_TRAIN.py:84,180); under
data parallelism gradients are averaged over ranks and only rank 0 writes files.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from corrif_b200.metrics import Jaccard2  # noqa: E402
from corrif_b200.train import TrainStep, broadcast_module  # noqa: E402

device = torch.device("cuda:%d" % int(os.environ.get("LOCAL_RANK", "0")) if torch.cuda.is_available() else "cpu")


def _rank0():
    return (not dist.is_initialized()) or dist.get_rank() == 0


def _global_mean(values):
    """values: list of device scalars -> python float of the mean over all ranks' entries."""
    t = torch.stack([v.reshape(()) for v in values]).double()
    s = torch.stack([t.sum(), torch.tensor(float(t.numel()), device=t.device, dtype=torch.float64)])
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(s)
    return (s[0] / s[1]).item()


def train_model(n_epochs, trainloss, validationloss, accuracy, model, scheduler, lrFile, training_generator,
                optim, lim, trainFile, trainaccFile, trainepochFile, validation_generator, valFile,
                valaccFile, pathm, i, modeltype):
    if trainloss != "BCEWithLogitsLoss" or accuracy != "Jaccard":
        raise ValueError("only trainloss='BCEWithLogitsLoss' and accuracy='Jaccard' exist in the reference")
    broadcast_module(model)
    step = TrainStep(model, optim, lim=lim)
    training_losses = []
    for epoch in range(n_epochs):
        model.train()
        scheduler.step()                                            # before any optimizer step, as :46
        if _rank0():
            print("Epoch:", epoch, "LR:", scheduler.get_last_lr())
            lrFile.write("Epoch:" + " " + str(epoch) + " " + "LR:" + " " + str(scheduler.get_last_lr()) + "\n")
            lrFile.write(str(scheduler.state_dict()) + "\n")
        losses, jac, pixels = [], None, 0
        for trainim, trainmas in training_generator:
            out = step((trainim.to(device, non_blocking=True), trainmas.to(device, non_blocking=True)))
            losses.append(out["loss"])
            jac = out["jaccard_sum"] if jac is None else jac + out["jaccard_sum"]
            pixels += out["pixels"]
        training_loss = _global_mean(losses)                        # the only host reads of the epoch
        tj = torch.stack([jac.reshape(()).double(), torch.tensor(float(pixels), device=jac.device, dtype=torch.float64)])
        if dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(tj)
        train_jac = (tj[0] / tj[1]).item()
        training_losses.append(training_loss)
        if _rank0():
            trainFile.write(str(training_loss) + "\n")
            trainaccFile.write(str(train_jac) + "\n")
            trainepochFile.write(str(epoch) + "\n")
            print("Training Jaccard:", train_jac, " (epoch:", epoch, ")")
            lrFile.write("Training loss:" + str(training_loss) + "\n")
            lrFile.write("Training accuracy:" + str(train_jac) + "\n")
            torch.save(model.state_dict(), os.path.join(pathm, "iremmodel{}.pt".format(i)))
        validate(validationloss, accuracy, validation_generator, valFile, valaccFile, lim, lrFile, pathm, i,
                 modeltype, model=model)
    if _rank0():
        torch.save(model.state_dict(), os.path.join(pathm, "Finaliremmodel{}.pt".format(i)))


def validate(validationloss, accuracy, validation_generator, valFile, valaccFile, lim, lrFile, pathm, i,
             modeltype, model=None):
    if model is None:       # reference behaviour: rebuild from the checkpoint of this epoch
        from mmvit4 import MMVit4
        model = MMVit4(num_cls=1).to(device)
        model.load_state_dict(torch.load(os.path.join(pathm, "iremmodel{}.pt".format(i))))
    was_training = model.training
    model.eval()
    losses, jac, pixels = [], None, 0
    with torch.no_grad():
        for valim, valmas in validation_generator:
            images, masks = valim.to(device, non_blocking=True), valmas.to(device, non_blocking=True)
            outputs = model(images)
            losses.append(F.binary_cross_entropy_with_logits(outputs, masks))
            load = len(masks) * lim * lim
            j = Jaccard2(masks[:, 0].reshape(load, 1), outputs[:, 0].reshape(load, 1)) * load
            jac = j if jac is None else jac + j
            pixels += load
    model.train(was_training)
    val_loss = float(np.mean([v.item() for v in losses]))
    dni = (jac / pixels).item()
    if _rank0():
        valFile.write(str(val_loss) + "\n")
        valaccFile.write(str(dni) + "\n")
        print("Validation Jaccard:", dni)
        lrFile.write("Validation loss:" + str(val_loss) + "\n")
        lrFile.write("Validation accuracy:" + str(dni) + "\n")
    return val_loss, dni
