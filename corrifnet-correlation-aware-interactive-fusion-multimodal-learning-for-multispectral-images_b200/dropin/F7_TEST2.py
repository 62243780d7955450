"""Drop-in for the reference's F7_TEST2.py: ``test_model`` with the reference's argument list and text outputs
(F7_TEST2.py:38-184) for the MMVit4 branch.

Reference behaviour kept: rebuild the model, load ``Finaliremmodel{i}.pt`` strictly (:126), eval mode, mean
BCE-with-logits loss over batches and pixel-weighted Jaccard2 of channel 0 (:167-183), one line each into
``testFile`` / ``testaccFile``.  Differences: the device is ``cuda:LOCAL_RANK`` instead of the hard-wired
``cuda:0`` (:35); the eval loop is the sharded, sync-free one of F4_TRAIN.evaluate; the first-batch figure and the
HSV overlay (:140-166, matplotlib / cv2 / F11_SEGPLOT: visualisation, out of scope) are written only when
matplotlib is importable - the three trMean arguments exist only for that overlay."""
import os

import torch

from F4_TRAIN import _rank0, device, ensure_distributed, evaluate


def test_model(test_generator, lim, testFile, testaccFile, i, modeltype, pathm, trMeanR, trMeanG, trMeanB,
               model=None):
    if modeltype != "MMVit4":
        raise ValueError("this drop-in carries the CorrIFNet (MMVit4) branch of test_model only, got %r" % (modeltype,))
    ensure_distributed()
    if model is None:
        from mmvit4 import MMVit4
        net = MMVit4(num_cls=1).to(device)                                   # :53-54
        net.load_state_dict(torch.load(os.path.join(pathm, "Finaliremmodel{}.pt".format(i)), map_location=device))
    else:
        net = model
    test_loss, dni = evaluate(net, test_generator, lim)
    if _rank0():
        _first_batch_figure(net, test_generator, pathm)
        testFile.write(str(test_loss) + "\n")                                # :181
        testaccFile.write(str(dni) + "\n")                                   # :182
        print("Test Jaccard:", dni)
    return test_loss, dni


def _first_batch_figure(net, test_generator, pathm):
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
    except ImportError:
        return
    testim, testmas = next(iter(test_generator))
    net.eval()
    with torch.no_grad():
        out = net(testim.to(device))
    fig = plt.figure()
    for k, (title, t) in enumerate((("Test Predicted Mask", out[0, 0, 0]), ("Ground Truth Mask", testmas[0, 0, 0])), 1):
        ax = fig.add_subplot(1, 2, k)
        ax.imshow(t.float().cpu().numpy(), cmap="gray")
        ax.set_title(title)
    fig.savefig(os.path.join(pathm, "mask_comparison.png"))
    plt.close(fig)
