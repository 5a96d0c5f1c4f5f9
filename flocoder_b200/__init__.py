"""flocoder_b200 -- B200-native (sm_100a) latent flow-matching sampling path of flocoder.

Public surface mirrors the reference's for this path only:

    from flocoder_b200.unet import Unet                      # flocoder.unet.Unet
    from flocoder_b200.sampling import (warp_time, rk4_step, v_func_cfg,
                                        generate_latents_rk4, generate_latents, euler_sampler)
    from flocoder_b200.dist import generate_latents_sharded  # batch-sharded multi-GPU sampling

Everything is backed by the C-ABI library declared in ``include/flocoder_b200.h``; importing the
package is cheap and does not load the library, but any compute call raises if it is not built.
"""
from .unet import Unet  # noqa: F401
from .sampling import (  # noqa: F401
    warp_time, rk4_step, v_func_cfg, generate_latents_rk4, generate_latents, euler_sampler, time_grid,
)

__version__ = "0.1.0"
