from .densenet import densenet121_unet_program, init_densenet_weights, DENSENET_BLOCKS  # noqa: F401
