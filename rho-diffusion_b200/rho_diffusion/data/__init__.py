from rho_diffusion.data.parameter_space import AbstractParameterSpace, DiscreteParameterSpace, sample_from_discrete_parameter_space

__all__ = ["AbstractParameterSpace", "DiscreteParameterSpace", "sample_from_discrete_parameter_space"]
