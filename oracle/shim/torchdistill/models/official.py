import torchvision


def get_image_classification_model(model_config, distributed=False):
    key = model_config['key']
    fn = torchvision.models.__dict__.get(key)
    if fn is None or not callable(fn):
        return None
    return fn(**model_config.get('kwargs', dict()))


def get_object_detection_model(model_config):
    fn = torchvision.models.detection.__dict__.get(model_config['key'])
    return None if fn is None else fn(**model_config.get('kwargs', dict()))


def get_semantic_segmentation_model(model_config):
    fn = torchvision.models.segmentation.__dict__.get(model_config['key'])
    return None if fn is None else fn(**model_config.get('kwargs', dict()))
