def register_mid_level_loss(arg=None, **kwargs):
    return arg if callable(arg) else (lambda f: f)
