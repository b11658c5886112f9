"""maua_b200 -- Blackwell (sm_100a) replacement for the audio-reactive StyleGAN render path of maua.

Host side mirrors the reference interfaces (maua.GAN.wrappers, maua.audiovisual); all arithmetic on
the path runs in libmaua_b200.so (hand-written CUDA behind the C ABI of include/maua_b200.h).
"""
__version__ = "0.1.0"
