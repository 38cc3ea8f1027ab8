"""vs_seg_b200: B200-native (sm_100a) hot path of KCL-BMEIS/VS_Seg.

csrc/ + libvsseg_b200.so  hand-written CUDA kernels behind the C ABI of include/vsseg_b200.h
lib / tensors             ctypes binding and device-tensor descriptors
engine                    fused eval plan of UNet2d5_spvPA
sliding_window / parallel MONAI-compatible sliding-window inference, patch-index sharding
compat                    the handful of MONAI names the reference API mentions (Norm, Act, ...)
The reference-facing API lives in params/, VS_train.py and VS_inference.py at the repo root.
"""
__version__ = "0.1.0"
