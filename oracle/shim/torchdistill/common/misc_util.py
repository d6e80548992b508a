import importlib
import inspect


def get_functions_as_dict(module_name):
    module = importlib.import_module(module_name)
    return {name: fn for name, fn in inspect.getmembers(module, inspect.isfunction)}


def get_classes_as_dict(module_name):
    module = importlib.import_module(module_name)
    return {name: c for name, c in inspect.getmembers(module, inspect.isclass)}
