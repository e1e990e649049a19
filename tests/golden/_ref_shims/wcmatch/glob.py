import fnmatch


def globfilter(names, pattern, **kw):  # minimal stand-in
    return [n for n in names if fnmatch.fnmatch(n, pattern)]


def translate(*a, **k):
    raise NotImplementedError


GLOBSTAR = BRACE = EXTGLOB = 0
