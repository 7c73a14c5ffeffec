"""Import stub: img-compression/vae_models.py imports tensorflow_compression at module level (line 119) for the
BLS2017 conv nets, which are out of scope; nothing here is ever called."""
