"""vsrt -- Python (ctypes) binding of the B200-native functional ray-traversal library.

The product is the C-ABI shared library libvsrt.so (include/vsrt.h; hand-written sm_100a CUDA).  This
package only loads it; there is no Python or CPU implementation of the path behind it, and importing
`vsrt.api` on a machine without the built library raises."""
from . import _abi  # noqa: F401
