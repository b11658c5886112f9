"""Super-resolution: mirror of maua/super (the RealESRGAN x4 image model of SURVEY §8f N6)."""
