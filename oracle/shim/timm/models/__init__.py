from . import regnet, resnest, vision_transformer_hybrid  # noqa: F401
