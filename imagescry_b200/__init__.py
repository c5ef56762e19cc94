"""imagescry_b200 — the data-parallel sift path of imagescry on B200 (sm_100a).

Host-side mirror of the reference's interface for this path; every operator runs in
`lib/libimagescry_b200.so` (hand-written CUDA behind the C ABI of `include/imagescry_b200.h`).
There is no CPU or PyTorch fallback: operators raise if the library is missing or the tensors are
not on a CUDA device.
"""

__version__ = "0.1.0"
