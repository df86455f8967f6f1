"""``find_resource``: locate a resource file like the reference does (pisa/utils/resources.py).

Search order: the path as given (absolute or relative to the CWD), each directory of the
colon-separated ``PISA_RESOURCES`` environment variable, then the resources shipped inside this
package (PREM tables, example configs)."""
import os

from pisa_b200 import RESOURCES_DIR

__all__ = ["find_resource", "resource_paths"]


def resource_paths():
    paths = []
    env = os.environ.get("PISA_RESOURCES", "")
    for p in env.split(":"):
        p = os.path.expandvars(os.path.expanduser(p.strip()))
        if p and os.path.isdir(p):
            paths.append(p)
    paths.append(RESOURCES_DIR)
    return paths


def find_resource(resource, fail=True):
    resource = os.path.expandvars(os.path.expanduser(str(resource)))
    if os.path.isabs(resource) or os.path.exists(resource):
        if os.path.exists(resource):
            return os.path.abspath(resource)
    else:
        for root in resource_paths():
            cand = os.path.join(root, resource)
            if os.path.exists(cand):
                return cand
    if fail:
        raise IOError('Could not find resource "%s" in %s' % (resource, resource_paths()))
    return None
