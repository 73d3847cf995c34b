from rho_diffusion.models.unet_v2 import UNet as UNetv2
from rho_diffusion.models.conditioning import MultiEmbeddings

__all__ = ["UNetv2", "MultiEmbeddings"]
