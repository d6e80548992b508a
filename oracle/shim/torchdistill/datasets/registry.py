def _passthrough(arg=None, **kwargs):
    return arg if callable(arg) else (lambda f: f)


register_transform = register_collate_func = register_dataset = register_batch_sampler = _passthrough
