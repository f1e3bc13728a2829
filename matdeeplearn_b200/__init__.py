"""B200-native message-passing engine behind the matdeeplearn.models operator
surface.  See DESIGN.md."""
import torch as _torch

__version__ = "0.1.0"

# fp32 contract of the path: the dense node-level layers that stay on library
# kernels (cuBLAS GEMMs, cuDNN GRU of MPNN -- forward AND backward) must not silently drop to
# TF32; the tensor-core contractions inside the engine use explicit 3xTF32 splitting instead.
_torch.backends.cuda.matmul.allow_tf32 = False
_torch.backends.cudnn.allow_tf32 = False
