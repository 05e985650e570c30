"""B200-native forward path of achraf-15/neural_image_compression behind the reference's module API.

    from neural_image_compression_b200.Models import JointAutoregressiveHierarchical
    from neural_image_compression_b200.RateDistortionLoss import rd_loss

Same class names, constructor arguments, ``state_dict`` keys and output dict as the reference's
``Models.py`` / ``Components.py`` / ``ContextModels.py`` / ``ParametersModels.py`` /
``EntropyModels.py`` / ``RateDistortionLoss.py``; the arithmetic runs in the hand-written sm_100a
kernels of ``libnic_b200.so`` (C ABI: include/nic.h).  No CPU path, no other GPU architecture.
"""
__version__ = "0.1.0"
