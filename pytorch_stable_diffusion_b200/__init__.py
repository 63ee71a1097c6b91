"""B200-native (sm_100a) Stable Diffusion sampling path.

Drop-in for the Python surface of dawmro/pytorch_stable_diffusion (sd/*.py): the same module
classes, state_dict keys, DDPMSampler and pipeline.generate(), with the arithmetic executed by
hand-written CUDA kernels reached through the C ABI declared in include/sdb200.h.
"""
__version__ = "0.1.0"
