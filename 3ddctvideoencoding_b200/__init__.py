"""dct3d-b200: the 3D-DCT video codec hot path (julianopiccoli/3dDCTVideoEncoding) on NVIDIA B200.

The product is libdct3d.so (CUDA, sm_100a; C ABI in include/dct3d.h).  This package holds its
in-tree build, a ctypes binding, and host-side mirrors of the reference's Encoder/Decoder
entry points.  The package name starts with a digit: import it with
importlib.import_module("3ddctvideoencoding_b200").
"""
__all__ = ["build", "_lib", "codec", "synth"]
