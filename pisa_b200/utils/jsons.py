"""JSON interchange in the reference's on-disk layout (pisa/utils/jsons.py:120-330).

``to_json(content, filename)`` writes what real PISA's ``jsons.to_json`` writes: objects are replaced by their
``serializable_state``, numpy arrays by nested lists, numpy scalars by python scalars; ``indent=2``; a ``.bz2``
extension compresses the text.  ``from_json(filename)`` reads either form back into ordered dicts.  This is the
hand-over point between a template produced here (``ContainerSet.get_mapset`` -> ``MapSet.to_json``) and the
reference's own analysis tooling (``MapSet.from_json``, map.py:2242-2262).
"""
import bz2
import json
import os
from collections import OrderedDict
from numbers import Integral, Real

import numpy as np

from pisa_b200.utils.units import Quantity

__all__ = ["to_json", "from_json", "dumps", "loads", "NumpyEncoder"]

JSON_EXTS = ("json",)
ZIP_EXTS = ("bz2",)


class NumpyEncoder(json.JSONEncoder):
    """jsons.py:283-330: serializable_state first, then quantities, arrays, numpy scalars."""

    def default(self, o):  # pylint: disable=method-hidden
        if hasattr(o, "serializable_state"):
            return o.serializable_state
        if isinstance(o, Quantity):
            return [self._plain(o.magnitude), str(o.units)]
        if isinstance(o, np.ndarray):
            return o.tolist()
        if isinstance(o, np.bool_):
            return bool(o)
        if isinstance(o, Integral):
            return int(o)
        if isinstance(o, Real):
            return float(o)
        if isinstance(o, (set, frozenset, tuple)):
            return list(o)
        return super().default(o)

    @staticmethod
    def _plain(x):
        return x.tolist() if isinstance(x, np.ndarray) else x


def dumps(content, indent=2, sort_keys=False):
    return json.dumps(content, indent=indent, cls=NumpyEncoder, sort_keys=sort_keys, allow_nan=True)


def loads(text):
    return json.loads(text, object_pairs_hook=OrderedDict)


def _ext(filename):
    return os.path.splitext(filename)[1].replace(".", "").lower()


def to_json(content, filename, indent=2, overwrite=True, warn=True, sort_keys=False):
    """jsons.py:196-277.  `warn` is accepted for signature compatibility (there is no logger here)."""
    if hasattr(content, "to_json"):
        return content.to_json(filename, indent=indent, overwrite=overwrite, warn=warn, sort_keys=sort_keys)
    ext = _ext(filename)
    if ext not in JSON_EXTS + ZIP_EXTS:
        raise ValueError("Unrecognized extension '%s' of file '%s': expected .json or .bz2" % (ext, filename))
    if os.path.exists(filename) and not overwrite:
        raise IOError("Refusing to overwrite existing path '%s'" % filename)
    data = dumps(content, indent=indent, sort_keys=sort_keys).encode()
    with open(filename, "wb") as f:
        f.write(bz2.compress(data) if ext == "bz2" else data)
    return None


def from_json(filename, cls=None):
    """jsons.py:120-193: a path or a name below the package's resources; `cls` is instantiated from the content."""
    from pisa_b200.utils.resources import find_resource
    path = filename if os.path.isfile(filename) else find_resource(filename)
    ext = _ext(path)
    if ext not in JSON_EXTS + ZIP_EXTS:
        raise ValueError("Unrecognized extension '%s' of file '%s': expected .json or .bz2" % (ext, filename))
    with open(path, "rb") as f:
        raw = f.read()
    content = loads((bz2.decompress(raw) if ext == "bz2" else raw).decode())
    if cls is None:
        return content
    if isinstance(content, dict):
        return cls(**content)
    if isinstance(content, (list, tuple)):
        return cls(*content)
    return cls(content)
