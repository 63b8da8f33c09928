"""sda_b200 -- B200-native drop-in for the data-parallel hot path of snipsco/sda.

The product is `libsda_b200.so` (hand-written sm_100a CUDA behind the C ABI of
`include/sda_b200.h`); this package is the host-side mirror of the reference's
`sda_client::crypto` trait surface on top of it.  Nothing here computes on the CPU.
"""
from ._lib import LIB_PATH, PROTOTYPES, load  # noqa: F401
from .crypto import (Context, CryptoModule, LinearMaskingScheme, LinearSecretSharingScheme,  # noqa: F401
                     MaskCombiner, SdaClientError, SecretMasker, SecretReconstructor, SecretUnmasker,
                     ShareCombiner, ShareGenerator)
from . import params  # noqa: F401
