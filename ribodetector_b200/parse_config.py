"""``config.json`` surface of the reference (``ribodetector/parse_config.py:10-92``): same keys
(``name``, ``n_gpu``, ``arch.type``, ``arch.args.*``, ``state_file.{mcc,recall}``), the same
reflection factory ``init_obj`` (the plugin seam: ``arch.type`` is resolved with ``getattr`` on the
model module) and the same logger format."""
import json
import logging
from collections import OrderedDict
from pathlib import Path


class ConfigParser:
    log_levels = {0: logging.WARNING, 1: logging.INFO, 2: logging.DEBUG}

    def __init__(self, config):
        self.config = config

    @classmethod
    def from_json(cls, config_json):
        with Path(config_json).open("rt") as handle:
            return cls(json.load(handle, object_hook=OrderedDict))

    def init_obj(self, name, module, *args, **kwargs):
        module_name = self[name]["type"]
        module_args = dict(self[name]["args"])
        assert all(k not in module_args for k in kwargs), "Overwriting kwargs given in config file is not allowed"
        module_args.update(kwargs)
        return getattr(module, module_name)(*args, **module_args)

    def __getitem__(self, name):
        return self.config[name]

    def get_logger(self, name, verbosity=2, logfile=None):
        handlers = [logging.StreamHandler()]
        if logfile is not None:
            handlers.append(logging.FileHandler(logfile, mode="w"))
        assert verbosity in self.log_levels, "verbosity option {} is invalid. Valid options are {}.".format(
            verbosity, self.log_levels.keys())
        logging.basicConfig(level=self.log_levels[verbosity], format="%(asctime)s : %(levelname)s  %(message)s",
                            datefmt="%Y-%m-%d %H:%M:%S", handlers=handlers, force=True)
        return logging.getLogger(name)
