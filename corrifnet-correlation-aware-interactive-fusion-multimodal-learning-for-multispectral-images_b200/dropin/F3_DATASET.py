"""Drop-in for the reference's F3_DATASET.py: ``satellitedata(images, masks, transform=None)``."""
from torch.utils.data import Dataset


class satellitedata(Dataset):
    """Indexes two pre-loaded tensors: sample = (images[i] [3,3,H,W], masks[i] [3,1,H,W])."""

    def __init__(self, images, masks, transform=None):
        self.images, self.masks, self.transform = images, masks, transform

    def __len__(self):
        return len(self.images)

    def __getitem__(self, index):
        im, ma = self.images[index], self.masks[index]
        if self.transform:
            im, ma = self.transform(im), self.transform(ma)
        return im, ma
