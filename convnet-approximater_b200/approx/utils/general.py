"""Small host helpers with the reference's names (src/approx/utils/general.py)."""
import os
import warnings


def to_2tuple(x) -> tuple:
    """int -> (x, x); tuple -> its first two entries (general.py:35-38)."""
    return x[:2] if isinstance(x, tuple) else (x, x)


def check_file(file, ext=None):
    """True iff `file` names an existing regular file (optionally with an extension in `ext`)."""
    if file is None:
        return False
    if not os.path.exists(file):
        warnings.warn(f"{file} is not exists")
        return False
    if not os.path.isfile(file):
        warnings.warn(f"{file} must be a file")
        return False
    return not ext or os.path.splitext(file)[1] in ext


def parse_path(path):
    """-> (dir, filename, extension)."""
    head, tail = os.path.split(path)
    name, extension = os.path.splitext(tail)
    return head, name, extension


def is_method_overridden(method, base_class, derived_class):
    assert isinstance(base_class, type), "base_class doesn't accept instance, Please pass class instead."
    derived = derived_class if isinstance(derived_class, type) else derived_class.__class__
    return getattr(derived, method) != getattr(base_class, method)
