"""Reader for the `.cf` model configs (`use parent.cf`, `dotted.key = <python literal>`, `#`
comments) that the reference parses with fjcommon.config_parser
(/root/reference/src/test/multiscale_tester.py:183, src/configs/ms/*.cf).  Returns an object with
attribute access (`cfg.q.C`, `cfg.enc.cls`, ...) as the model code expects."""
import ast
import os

CONFIG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'configs')


class Config(object):
    def __repr__(self):
        return 'Config(%s)' % ', '.join('%s=%r' % kv for kv in sorted(self.__dict__.items()))

    def all_params_and_values(self):
        return sorted(self.__dict__.items())


def _assign(root, dotted, value):
    *parents, leaf = dotted.split('.')
    node = root
    for name in parents:
        child = node.__dict__.get(name)
        if child is None:
            child = Config()
            setattr(node, name, child)
        node = child
    setattr(node, leaf, value)


def _read(root, path):
    with open(path) as f:
        for raw in f:
            line = raw.partition('#')[0].strip()
            if not line:
                continue
            if line.startswith('use '):
                _read(root, os.path.join(os.path.dirname(path), line[4:].strip()))
            elif line.startswith('constrain '):
                continue
            else:
                key, _, val = line.partition('=')
                _assign(root, key.strip(), ast.literal_eval(val.strip()))


def parse(path):
    """-> (config, path relative to cwd), mirroring fjcommon.config_parser.parse."""
    if not os.path.isfile(path):
        alt = os.path.join(CONFIG_DIR, path)
        if os.path.isfile(alt):
            path = alt
    cfg = Config()
    _read(cfg, path)
    return cfg, os.path.relpath(path)


def ms_config(name):
    """ms_config('cr') / 'cr_rgb_shared' / 'cr_rgb' -> parsed config shipped with the package."""
    if not name.endswith('.cf'):
        name += '.cf'
    return parse(os.path.join(CONFIG_DIR, 'ms', name))[0]
