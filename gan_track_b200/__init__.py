"""gan_track_b200 -- B200-native (sm_100a) implementation of the StyleGAN2-ADA training hot path that
ltronchin/Gan-track drives: the `torch_utils.ops` op surface (bias_act, upfirdn2d, conv2d_resample, conv2d_gradfix,
fma, grid_sample_gradfix, modulated_conv2d), the networks / loss / AugmentPipe that call it, and the data-parallel
training step.  Kernels live in `csrc/` behind the C ABI of `include/gantrack_b200.h`; there is no CPU path.
"""
__version__ = '0.1.0'
