from .field_heads import SemanticFieldHead  # noqa: F401
from .ray_generators import OrthographicRayGenerator  # noqa: F401
from .ray_samplers import UniformLinDispPiecewiseSampler, UniformSamplerWithNoise  # noqa: F401
from .proposal_sampler import ProposalNetworkSampler  # noqa: F401
