"""Drop-in mirror of the reference's `models` package for the sampling path:
same module paths, class names, constructor arguments and state_dict keys."""
from .cross_attention import CrossAttentionFusion  # noqa: F401
from .diffusion import GaussianDiffusion  # noqa: F401
from .embedding import CondProjection, SinusoidalPosEmb, TimestepEmbedding  # noqa: F401
from .unet1d_ultimate import UNet1D_ultimate  # noqa: F401
from .unet1d import UNet1D  # noqa: F401
