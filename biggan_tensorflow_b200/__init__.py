"""biggan_tensorflow_b200 -- B200-native (sm_100a) BigGAN G+D training step behind the operator
surface of david-jk/BigGAN-Tensorflow (ops.py / DiffAugment_tf.py).  Importing this package loads
libbiggan_b200.so and raises if it is missing: there is no CPU or library fallback."""
from . import _lib  # noqa: F401  (fails loudly when the CUDA library is absent)
from .tensor import BF16, F32, Tensor, no_grad  # noqa: F401
from .model import CONFIGS, Config  # noqa: F401
