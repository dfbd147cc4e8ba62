from .denoiser import Denoiser, DenoiserConfig, SigmaDistributionConfig
from .inner_model import InnerModelConfig
from .diffusion_sampler import DiffusionSampler, DiffusionSamplerConfig
