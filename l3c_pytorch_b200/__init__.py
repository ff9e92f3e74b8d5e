"""l3c_pytorch_b200 -- B200-native (sm_100a) implementation of L3C's encode/decode hot path:
multi-scale conv encoder/decoder -> discretised-logistic-mixture CDFs -> range coder, behind the
reference's MultiscaleBlueprint / Bitcoding / torchac API.  See DESIGN.md and INTEGRATION.md.

Importing the package loads libl3c_b200.so (built in-tree by `python -m l3c_pytorch_b200.build`);
there is no CPU fallback.
"""
from . import _lib                                   # noqa: F401  (fails loudly if the .so is missing)
from . import config, engine, torchac                # noqa: F401
from .bitcoding import Bitcoding                     # noqa: F401
from .blueprint import MultiscaleBlueprint           # noqa: F401
from .coders import ArithmeticCoder                  # noqa: F401
from .codec import BatchCodec                        # noqa: F401
from .engine import set_conv_precision, get_conv_precision   # noqa: F401

__version__ = '0.1.0'
