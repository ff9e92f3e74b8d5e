"""Minimal re-statement of fjcommon.config_parser.parse for the reference's ms/*.cf files:
`use parent.cf` inheritance, `a.b = <python literal>`, `#` comments."""
import ast
import os


class _Node(object):
    def __repr__(self):
        return 'Config({})'.format(self.__dict__)

    def all_params_and_values(self):
        return sorted(self.__dict__.items())


def _set(root, dotted, value):
    parts = dotted.split('.')
    node = root
    for p in parts[:-1]:
        if not hasattr(node, p):
            setattr(node, p, _Node())
        node = getattr(node, p)
    setattr(node, parts[-1], value)


def _parse_into(root, path):
    base = os.path.dirname(path)
    with open(path) as f:
        for line in f:
            line = line.split('#', 1)[0].strip()
            if not line:
                continue
            if line.startswith('use '):
                _parse_into(root, os.path.join(base, line[4:].strip()))
                continue
            if line.startswith('constrain '):
                continue
            key, val = line.split('=', 1)
            _set(root, key.strip(), ast.literal_eval(val.strip()))


def parse(path):
    root = _Node()
    _parse_into(root, path)
    return root, os.path.relpath(path)
