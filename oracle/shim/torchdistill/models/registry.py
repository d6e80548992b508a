MODEL_DICT = dict()


def register_model(arg=None, **kwargs):
    def _register(cls_or_func):
        MODEL_DICT[kwargs.get('key', cls_or_func.__name__)] = cls_or_func
        return cls_or_func
    return _register(arg) if callable(arg) else _register


def get_model(key, repo_or_dir=None, *args, **kwargs):
    if key in MODEL_DICT:
        return MODEL_DICT[key](*args, **kwargs)
    raise ValueError('model_name `{}` is not expected'.format(key))
