"""B200-native k-space -> image input stage for bonhchi/mri_acl_imagesegmentation_adsp.

Importing this package needs neither a GPU nor the built CUDA library; every compute
entry point loads ``csrc/libmriacl_recon.so`` on first use and raises if it is missing
(there is no CPU fallback).
"""
__version__ = "0.1.0"
