"""animeface_b200 -- B200-native StyleGAN2 training step behind the STomoya/animeface API."""
__version__ = '0.1.0'
