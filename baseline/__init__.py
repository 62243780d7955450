"""Measurement baselines (NOT product code): stock-PyTorch eager restatements used only by bench.py's
``eager_b200`` arm and by tests as the "plain PyTorch reference of the same op" on the GPU."""
