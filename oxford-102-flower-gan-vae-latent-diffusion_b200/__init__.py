"""B200-native drop-in for the latent-DDPM sampling path of
ynyeh0221/Oxford-102-Flower-GAN-VAE-latent-diffusion (v2/model_train_test.py).

The directory name carries hyphens (it is the name the build contract fixes), so import it through the
shim at the repository root:  `import ldm_b200`.
"""
from .engine import default_precision, get_engine, invalidate, set_default_precision   # noqa: F401
from .modules import (CALayer, ClassEmbedding, ConditionalDenoiseDiffusion, ConditionalUNet, Decoder, Encoder,  # noqa: F401
                      LayerNorm2d, ResidualBlock, SimpleAutoencoder, SpatialAttention, SwitchSequential, Swish, TimeEmbedding,
                      UNetAttentionBlock, UNetResidualBlock,
                      euclidean_distance_loss, generate_class_samples, init_weights, load_autoencoder_checkpoint)
from .sharding import generate_sharded, shard_bounds                       # noqa: F401
from .io_utils import load_unet_checkpoint, parse_epoch_from_filename, save_image_grid, to_uint8, write_png   # noqa: F401
from . import v3                                                            # noqa: F401  (v3 multi-conditional denoiser)
from . import v4                                                            # noqa: F401  (v4 / v5 pixel-space diffusion)
from ._lib import LIB_PATH, LdmError                                        # noqa: F401

__all__ = ["ConditionalUNet", "ConditionalDenoiseDiffusion", "SimpleAutoencoder", "Decoder", "Encoder",
           "TimeEmbedding", "ClassEmbedding", "ResidualBlock", "CALayer", "SpatialAttention", "LayerNorm2d", "Swish",
           "UNetResidualBlock", "UNetAttentionBlock", "SwitchSequential", "generate_class_samples", "generate_sharded", "shard_bounds", "init_weights", "load_autoencoder_checkpoint",
           "get_engine", "invalidate", "set_default_precision", "default_precision", "LdmError", "LIB_PATH"]
