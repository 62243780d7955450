"""Drop-in for the reference's F6_CROSSVAL.py (host-side index arithmetic, F6_CROSSVAL.py:5-37): fold ``fno`` of
``fsiz`` is the test set, the first 10 % of the remaining indices the validation set, the rest the training set,
all mapped through the fixed permutation file ``randInd{N}.txt`` (looked up in the working directory first, as
the reference does, then in ``CORRIF_RANDIND_DIR``)."""
import os

import numpy as np


def _permutation(N):
    name = "randInd{}.txt".format(N)
    for d in (".", os.environ.get("CORRIF_RANDIND_DIR", "")):
        p = os.path.join(d, name)
        if d is not None and os.path.exists(p):
            with open(p) as f:
                return np.asarray([int(line) for line in f if line.strip()])
    raise FileNotFoundError(name)


def CrossVal(N, fno, fsiz):
    ind = _permutation(N)
    fold = fno - 1
    tstsize = int(N / fsiz)
    if (fold + 1) * tstsize > N:        # wrap-around fold (unreachable for fno <= fsiz; kept for the contract)
        positions = np.concatenate((np.arange((fold * tstsize) % N, N), np.arange(0, ((fold + 1) * tstsize) % N)))
    else:
        positions = np.arange(fold * tstsize, (fold + 1) * tstsize)
    rest = np.setdiff1d(ind, positions)                     # sorted values, as np.setdiff1d returns them (:26)
    valsize = int((N - tstsize) * 0.1)                      # "val is always 10% of training set" (:29-30)
    return ind[positions], ind[rest[valsize:]], ind[rest[:valsize]]     # tsind, trind, vlind
