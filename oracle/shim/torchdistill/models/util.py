def redesign_model(org_model, model_config, model_label, model_type='original'):
    raise NotImplementedError('off the bottleneck path')
