from rho_diffusion.models.unet_v2 import UNet as UNetv2
from rho_diffusion.models.unet import UNet, UNetBlock2d, UNetBlock3d
from rho_diffusion.models.vit import VisionTransformer
from rho_diffusion.models.conditioning import MultiEmbeddings

__all__ = ["UNetv2", "UNet", "UNetBlock2d", "UNetBlock3d", "VisionTransformer", "MultiEmbeddings"]
