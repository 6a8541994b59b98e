"""digipathai_b200 -- B200-native (sm_100a) implementation of DigiPathAI's ``getSegmentation`` hot path.

    from digipathai_b200.Segmentation import getSegmentation      # drop-in for DigiPathAI.Segmentation

Host orchestration is Python; every device computation goes through the C ABI in ``include/digipath_b200.h``
(``libdigipath_b200.so``, hand-written CUDA).  Importing this package does not load the library; the first use
of ``engine`` / ``Segmentation`` does, and fails loudly if it has not been built.
"""
__version__ = "0.1.0"
