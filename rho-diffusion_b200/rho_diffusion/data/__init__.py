from rho_diffusion.data.parameter_space import AbstractParameterSpace, DiscreteParameterSpace, sample_from_discrete_parameter_space
from rho_diffusion.data.deep_galaxy import DeepGalaxyDataset
from rho_diffusion.data.spectroscopy import SpectroscopyDataset
from rho_diffusion.data.synthetic import (DeviceLoader, GaussianBlobDataset, SphericalHarmonicDataset, compute_spherical_harmonic,
                                          make_spherical_grid)

__all__ = ["AbstractParameterSpace", "DiscreteParameterSpace", "sample_from_discrete_parameter_space", "SphericalHarmonicDataset", "DeepGalaxyDataset",
           "compute_spherical_harmonic", "make_spherical_grid", "GaussianBlobDataset", "DeviceLoader", "SpectroscopyDataset"]
