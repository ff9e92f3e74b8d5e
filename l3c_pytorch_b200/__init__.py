"""l3c_pytorch_b200 -- B200-native (sm_100a) implementation of L3C's encode/decode hot path:
multi-scale conv encoder/decoder -> discretised-logistic-mixture CDFs -> range coder, behind the
reference's MultiscaleBlueprint / Bitcoding / torchac API.  See DESIGN.md and INTEGRATION.md.

Importing the package loads libl3c_b200.so (built in-tree by `python -m l3c_pytorch_b200.build`);
there is no CPU fallback.
"""
import os as _os

# A decode in flight queues hundreds of kernels on a dozen streams, most of them waiting on events of other
# streams.  With the default of 8 hardware work queues several streams share a queue and a kernel that is ready
# waits behind queued kernels of ANOTHER decode that are not (measured: the first range-decoder launch of a
# decode waited 36 ms for the other lane's RGB scale).  Must be set before the CUDA context exists.
_os.environ.setdefault('CUDA_DEVICE_MAX_CONNECTIONS', '32')

from . import _lib                                   # noqa: F401,E402  (fails loudly if the .so is missing)
from . import config, engine, torchac                # noqa: F401,E402
from .bitcoding import Bitcoding                     # noqa: F401,E402
from .blueprint import MultiscaleBlueprint           # noqa: F401,E402
from .coders import ArithmeticCoder                  # noqa: F401,E402
from .codec import BatchCodec                        # noqa: F401,E402
from .engine import set_conv_precision, get_conv_precision   # noqa: F401,E402

__version__ = '0.1.0'
