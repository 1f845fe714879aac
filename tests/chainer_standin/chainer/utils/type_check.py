"""Eager stand-in for chainer.utils.type_check: the comparisons written in check_type_forward evaluate to plain bools."""


class InvalidType(Exception):
    pass


class _TypeInfo(object):
    def __init__(self, a, name):
        self.dtype = a.dtype
        self.ndim = a.ndim
        self.shape = tuple(a.shape)
        self.name = name


class _TypeInfoTuple(tuple):
    def size(self):
        return len(self)


def get_types(data, name, accept_none):
    return _TypeInfoTuple(_TypeInfo(a, "%s[%d]" % (name, i)) for i, a in enumerate(data))


def expect(*conditions):
    for k, c in enumerate(conditions):
        if not c:
            raise InvalidType("type_check.expect: condition %d does not hold" % k)
