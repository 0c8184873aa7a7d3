"""vtgaussian_slam_b200 -- B200-native (sm_100a) differentiable Gaussian-splatting hot path
of VTGaussian-SLAM: the rasteriser behind `diff_gaussian_rasterization` plus the fused
render -> loss -> backward -> Adam iteration of tracking and mapping.

The CUDA library (csrc/ -> lib/libvtgs_cuda.so) is loaded lazily on first use and there is
NO CPU fallback: calling any op without the built library or without a CUDA device raises.
"""
__version__ = "0.1.0"
