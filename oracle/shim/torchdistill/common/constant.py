import logging

def_logger = logging.getLogger('torchdistill')
