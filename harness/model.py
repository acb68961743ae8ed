"""A stock-PyTorch SQ regressor with the architecture torch/models.py:172-204 describes (ResNet-18 trunk with a
one-channel stem, 512 -> 256 -> 256 LeakyReLU neck, four linear heads: size sigmoid(3), shape sigmoid(2), position
sigmoid(3), rotation L2-normalised (4)).  The CNN is outside the hot path and stays on stock PyTorch / cuDNN
(north-star); this module only exists so that BASELINE configs 3 and 4 can be run without the reference tree.
Random initialisation (no pretrained download)."""
import torch
import torch.nn as nn
from torchvision.models import resnet18


class SQRegressor(nn.Module):
    def __init__(self, width: int = 256):
        super().__init__()
        trunk = resnet18(weights=None)
        stem = trunk.conv1
        trunk.conv1 = nn.Conv2d(1, stem.out_channels, stem.kernel_size, stem.stride, stem.padding, bias=False)
        trunk.fc = nn.Sequential(nn.Linear(512, width), nn.LeakyReLU(), nn.Linear(width, width), nn.LeakyReLU())
        self.trunk = trunk
        self.size, self.shape, self.position, self.rotation = (nn.Linear(width, k) for k in (3, 2, 3, 4))

    def forward(self, depth_images: torch.Tensor, raw: bool = False) -> torch.Tensor:
        """(B,1,H,W) -> (B,12) rows [a(3) | e(2) | t(3) | q(4)] like torch/train.py:88-89.  raw=True returns the head
        outputs before sigmoid / normalisation, for ImplicitLoss.from_heads (the activations then run inside the loss)."""
        z = self.trunk(depth_images)
        q = self.rotation(z)
        if raw:
            return torch.cat([self.size(z), self.shape(z), self.position(z), q], dim=1)
        return torch.cat([torch.sigmoid(self.size(z)), torch.sigmoid(self.shape(z)), torch.sigmoid(self.position(z)),
                          q / q.norm(dim=-1, keepdim=True)], dim=1)
