"""Drop-in for the reference's F2_MAIN.py experiment driver (F2_MAIN.py:45-313), CorrIFNet (``MMVit4``) branch.

Honours the reference's contract: the 18-line positional config ``<experiments>/model{i}.txt`` (:61-83), CrossVal
split (:85), ``get_images4`` (:88), three sequential DataLoaders (:90-111), Adam/SGD + StepLR (:168-173), a
time-stamped result directory, the six text logs opened in the working directory (:179-190), ``train_model`` then
``test_model`` (:191-204) and the summary log (:258-291).  Plotting (:294-304) needs matplotlib and is skipped when
it is absent.

  python F2_MAIN.py                              one GPU
  torchrun --nproc-per-node 8 F2_MAIN.py         batch-sharded data parallel (see F4_TRAIN.py); every rank builds
                                                 the same model from the same seed, rank 0 writes the files

Environment: CORRIF_EXPERIMENTS (default ../../experiments), CORRIF_DSTL_ROOT (F8_IMAGES4), CORRIF_SYNTHETIC=N
(N synthetic DSTL-shaped tiles instead of the .mat files - there is no DSTL data offline), CORRIF_SEED.
"""
import datetime
import os
import sys

import numpy as np
import torch
from torch.optim.lr_scheduler import StepLR
from torch.utils.data import DataLoader

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from F3_DATASET import satellitedata  # noqa: E402
from F4_TRAIN import _rank0, device, ensure_distributed, train_model  # noqa: E402
from F7_TEST2 import test_model  # noqa: E402
from mmvit4 import MMVit4  # noqa: E402

CONFIG_FIELDS = (("trainSetSize", int), ("fno", int), ("fsiz", int), ("valRatio", float), ("miniBatchSize", int),
                 ("n_epochs", int), ("learnRate", float), ("optimizerType", str), ("trainloss", str),
                 ("validationloss", str), ("accuracy", str), ("initialization", str), ("step_size", int),
                 ("gamma", float), ("lim", int), ("modeltype", str), ("chindex", str), ("transfertype", str))


def read_config(path):
    """The 18 positional lines of model{i}.txt (F2_MAIN.py:61-83) -> dict."""
    with open(path) as f:
        lines = [line.rstrip() for line in f]
    if len(lines) < len(CONFIG_FIELDS):
        raise ValueError("%s: expected %d lines, found %d" % (path, len(CONFIG_FIELDS), len(lines)))
    return {name: cast(lines[k]) for k, (name, cast) in enumerate(CONFIG_FIELDS)}


def synthetic_tiles(n, lim, seed):
    """SURVEY.md section 8d synthetic inputs: zero-centred tiles, one binary mask replicated per modality."""
    g = torch.Generator().manual_seed(seed)
    images = torch.randn(n, 3, 3, lim, lim, generator=g)
    masks = (torch.rand(n, 1, 1, lim, lim, generator=g) < 0.3).float().repeat(1, 3, 1, 1, 1)
    return images, masks


def load_data(cfg):
    n = cfg["trainSetSize"]
    syn = int(os.environ.get("CORRIF_SYNTHETIC", "0"))
    if syn:
        tst = n // cfg["fsiz"]
        val = int((n - tst) * 0.1)
        ind = np.arange(n)
        tsind, vlind, trind = ind[:tst], ind[tst:tst + val], ind[tst + val:]
        images, masks = synthetic_tiles(n, cfg["lim"], int(os.environ.get("CORRIF_SEED", "0")))
        return images, masks, tsind, trind, vlind, (0.0, 0.0, 0.0)
    from F6_CROSSVAL import CrossVal
    from F8_IMAGES4 import get_images4
    tsind, trind, vlind = CrossVal(n, cfg["fno"], cfg["fsiz"])
    images, masks, r, g, b = get_images4(n, cfg["fno"], cfg["fsiz"], tsind, trind, vlind, cfg["chindex"])
    return images, masks, tsind, trind, vlind, (r, g, b)


def main(i=0):
    begin = datetime.datetime.now()
    ensure_distributed()
    data_folder = os.environ.get("CORRIF_EXPERIMENTS", os.path.join("../../experiments"))
    cfg = read_config(os.path.join(data_folder, "model{}.txt".format(i)))
    if cfg["modeltype"] != "MMVit4":
        raise ValueError("this drop-in carries the CorrIFNet branch only (modeltype 'MMVit4'), got %r" % cfg["modeltype"])
    images, masks, tsind, trind, vlind, means = load_data(cfg)
    params = {"batch_size": cfg["miniBatchSize"], "shuffle": False}                     # :90
    gens = [DataLoader(satellitedata(images[ix], masks[ix]), **params) for ix in (trind, vlind, tsind)]
    torch.manual_seed(int(os.environ.get("CORRIF_SEED", "0")))         # every rank builds the same initial model
    model = MMVit4().to(device)                                                         # :122-123
    if cfg["transfertype"] == "yestr":                                                  # :160-161
        model.load_state_dict(torch.load(os.path.join(data_folder, "2021_3_16_10_32.pt"), map_location=device))
    # 'notr' applies init_weights, which only touches nn.Conv2d (:134-157): a no-op for MMVit4
    if cfg["optimizerType"] == "Adam":
        optim = torch.optim.Adam(model.parameters(), cfg["learnRate"])
    elif cfg["optimizerType"] == "SGD":
        optim = torch.optim.SGD(model.parameters(), cfg["learnRate"])
    else:
        raise ValueError("optimizerType must be Adam or SGD")
    scheduler = StepLR(optim, cfg["step_size"], cfg["gamma"])
    d = datetime.datetime.now()
    pathm = os.path.join(data_folder, "{}_{}_{}_{}_{}_model{}".format(d.year, d.month, d.day, d.hour, d.minute, i))
    if _rank0():
        os.makedirs(pathm, exist_ok=True)
    names = ("lrFile", "trainaccFile", "valaccFile", "trainepochFile", "trainFile", "valFile", "testaccFile", "testFile")
    # the reference opens its logs in the working directory (:179-190); ranks > 0 write nowhere
    files = {n: open(n + ".txt" if _rank0() else os.devnull, "w") for n in names}
    train_model(cfg["n_epochs"], cfg["trainloss"], cfg["validationloss"], cfg["accuracy"], model, scheduler,
                files["lrFile"], gens[0], optim, cfg["lim"], files["trainFile"], files["trainaccFile"],
                files["trainepochFile"], gens[1], files["valFile"], files["valaccFile"], pathm, i, cfg["modeltype"])
    if torch.distributed.is_initialized():
        torch.distributed.barrier()                     # rank 0 has written Finaliremmodel{i}.pt
    test_model(gens[2], cfg["lim"], files["testFile"], files["testaccFile"], i, cfg["modeltype"], pathm, *means)
    for f in files.values():
        f.close()
    if _rank0():
        _summary(pathm, cfg, begin, len(vlind), len(trind), model)
    return pathm


def _summary(pathm, cfg, begin, n_val, n_train, model):
    a = datetime.datetime.now()
    test_acc = [float(x) for x in open("testaccFile.txt")]
    with open(os.path.join(pathm, "{}_{}_{}_{}_{}.txt".format(a.year, a.month, a.day, a.hour, a.minute)), "w") as log:
        log.write("Date:" + str(datetime.date.today()) + "\n")
        log.write("Ending Time:" + str(a.hour) + ":" + str(a.minute) + "\n")
        log.write("Starting Time:" + str(begin.hour) + ":" + str(begin.minute) + "\n")
        for label, key in (("Data set size:", "trainSetSize"), ("Fold number:", "fno"), ("Fold number:", "fsiz")):
            log.write(label + str(cfg[key]) + "\n")
        log.write("Number of validation images:" + str(n_val) + "\n")
        log.write("Number of training images:" + str(n_train) + "\n")
        log.write("Mini batch size:" + str(cfg["miniBatchSize"]) + "\n")
        log.write("Type of initialization:" + cfg["initialization"] + "\n")
        log.write("Test accuracy:" + str(test_acc) + "\n")
        for label, key in (("Learning rate:", "learnRate"), ("Model version:", "modeltype"),
                           ("Optimizer type:", "optimizerType"), ("Total number of epochs:", "n_epochs"),
                           ("Training loss function:", "trainloss"), ("Validation loss function:", "validationloss"),
                           ("Accuracy function:", "accuracy"), ("Channel index:", "chindex"),
                           ("Transfer:", "transfertype")):
            log.write(label + str(cfg[key]) + "\n")
        log.write("Model Summary:" + "\n" + str(model) + "\n")
        log.writelines(open("lrFile.txt").readlines())


if __name__ == "__main__":
    if torch.cuda.is_available():
        print(torch.cuda.get_device_name(0))
    main(0)
    if torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()
