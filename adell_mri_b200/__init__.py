"""adell_mri_b200 — B200-native volumetric augmentation hot path for adell-mri.

Only what the hot path needs lives here: ``csrc/`` (CUDA kernels + C ABI), the ctypes binding
(``_lib``), the host composer (``plan``), the launcher (``engine``), device-side intensity
statistics (``stats``), multi-GPU plumbing (``dist``) and the mirror of the reference's transform
surface (``transforms``, ``transform_factory``, ``collate``; batch fast path in ``pipelines``).
"""

__version__ = "0.2.0"
