"""adell_mri_b200 — B200-native volumetric augmentation hot path for adell-mri.

Only what the hot path needs lives here: ``csrc/`` (CUDA kernels + C ABI), the ctypes
binding (``_lib``), the host composer (``plan``), the launcher (``engine``), device-side
intensity statistics (``stats``) and the mirror of the reference's transform surface
(``monai_compat``, ``transforms``, ``transform_factory``).
"""

__version__ = "0.1.0"
