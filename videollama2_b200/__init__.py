"""videollama2_b200 — B200 (sm_100a) engine for VideoLLaMA2's video->text prefill path.

Host code is Python/PyTorch (memory, streams, torch.distributed); all arithmetic runs in libvl2.so
(hand-written CUDA: tcgen05/TMEM/TMA GEMM + attention, fused row kernels) through the C-ABI in include/vl2.h.
"""
__version__ = "0.1.0"
