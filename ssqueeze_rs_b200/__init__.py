"""ssqueeze_rs_b200 -- B200 (sm_100a) engine behind the `ssqueeze._rs` API.

Layout: `csrc/` holds the CUDA kernels and the C ABI (`include/ssqcuda.h`),
`_lib.py` binds the shared library with ctypes, `_rs.py` mirrors the reference
module's callables (rust/src/lib.rs:23-35), `batch.py` is the multichannel
throughput path (device buffers, channel sharding across GPUs).
There is no CPU fallback: importing works anywhere the library is built, every
compute call needs a CUDA device.
"""
from . import _rs  # noqa: F401
from ._lib import SsqError, PanicException, lib_path  # noqa: F401

__all__ = ["_rs", "SsqError", "PanicException", "lib_path"]
