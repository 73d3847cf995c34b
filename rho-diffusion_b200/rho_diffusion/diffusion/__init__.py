from rho_diffusion.diffusion import schedule
from rho_diffusion.diffusion.abstract_diffusion import AbstractDiffusionPipeline
from rho_diffusion.diffusion.ddpm import DDPM
from rho_diffusion.diffusion.gaussian_diffusion import GaussianDiffusionPipeline
from rho_diffusion.diffusion.schedule import CosineBetaSchedule, LinearSchedule, SigmoidSchedule

__all__ = ["AbstractDiffusionPipeline", "DDPM", "GaussianDiffusionPipeline", "LinearSchedule", "CosineBetaSchedule", "SigmoidSchedule",
           "schedule"]
