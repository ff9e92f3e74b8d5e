"""`.partN` file naming for images that are coded as several crops
(/root/reference/src/bitcoding/part_suffix_helper.py:10-35)."""
import glob
import os
import re

_BASE = '.part'
_TAIL = re.compile(re.escape(_BASE) + r'(\d+)$')


def make_part_suffix(i):
    assert i >= 0, i
    return '%s%d' % (_BASE, i)


def contains_part_suffix(p):
    return _TAIL.search(p) is not None


def index_of_part_suffix(p):
    return int(_TAIL.search(p).group(1))


def iter_part_suffixes(pin):
    """All sibling part files of `pin`, ordered by part index."""
    assert os.path.isfile(pin) and contains_part_suffix(pin)
    stem = pin[:_TAIL.search(pin).start()] + _BASE
    found = [m for m in glob.glob(glob.escape(stem) + '*') if contains_part_suffix(m)]
    return sorted(found, key=index_of_part_suffix)


def existing_parts(pout):
    """Part files `pout.partN` that are already on disk (any N), ordered by part index."""
    found = [m for m in glob.glob(glob.escape(pout + _BASE) + '*') if contains_part_suffix(m)]
    return sorted(found, key=index_of_part_suffix)
