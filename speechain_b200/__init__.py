"""speechain_b200 — B200-native drop-in for SpeeChain's on-the-fly acoustic frontend.

Same module paths, class names, constructor keys and state-dict buffers as
``speechain.module.{frontend,augment,norm}`` of bagustris/speechain; the arithmetic runs in
hand-written sm_100a CUDA behind the C ABI of ``include/sfe.h`` (``csrc/libsfe.so``).
"""
__version__ = "0.1.0"
