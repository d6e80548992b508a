import os

import torch

from .constant import def_logger

logger = def_logger.getChild(__name__)


def is_main_process():
    return True


def get_world_size():
    return 1


def load_ckpt(ckpt_file_path, model=None, optimizer=None, lr_scheduler=None, strict=True):
    if ckpt_file_path is None or not os.path.isfile(str(ckpt_file_path)):
        logger.info('ckpt file is not found at `{}`'.format(ckpt_file_path))
        return None, None
    ckpt = torch.load(ckpt_file_path, map_location='cpu')
    if model is not None:
        model.load_state_dict(ckpt['model'] if 'model' in ckpt else ckpt, strict=strict)
    return ckpt.get('best_value', 0.0), ckpt.get('args', None)
