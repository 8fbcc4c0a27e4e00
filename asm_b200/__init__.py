"""asm_b200 -- B200 (sm_100a) implementation of ASM's convergence-diagnostic hot path.

The product is ``libasmgpu.so`` (CUDA kernels behind the C ABI of ``include/asmgpu.h``); this
package is the thin ctypes binding used by the tests and the benchmark, plus the host-side mirror
of the reference's criterion interface.  There is no CPU fallback: compute calls fail loudly
when the library or a B200 is missing.
"""
from ._lib import AsmGpuError, Encoder, GpuContext, lib, load_library  # noqa: F401
